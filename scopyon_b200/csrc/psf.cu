// PSF tables: radial profile (Born-Wolf / Gaussian) and fixed-point summed-area tables.
//
// Reference: PointSpreadingFunction.__get / get_distribution / radial_to_cartesian
// (/root/reference/src/scopyon/_epifm.py:90-96, 133-134, 181-213, 98-126).
// The reference materialises a 1999x1999 fp64 Cartesian table per integer-nm depth and
// sums slices of it per pixel.  Here the table is integrated once into an int64
// summed-area table (SAT): every later pixel box-sum is four reads, exact in integer
// arithmetic, hence independent of summation order and of how spots are binned.
#include "scb_common.cuh"

#include <stdarg.h>

#include <atomic>

// ------------------------------------------------------------------ error plumbing
static thread_local char g_scb_error[512] = "";

void scb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_scb_error, sizeof(g_scb_error), fmt, ap);
    va_end(ap);
}

extern "C" const char *scb_last_error(void) { return g_scb_error; }
extern "C" int scb_version(void) { return SCB_VERSION; }

// Multiprocessor count of the current device, cached per device ordinal.
int scb_sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
extern "C" int scb_device_sm_count(void) { return scb_sm_count(); }

// Host-callable Philox for known-answer tests of the counter-based generator.
extern "C" void scb_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    Philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// ------------------------------------------------------------------ radial profile
// Born-Wolf, _epifm.py:181-213:
//   psf(r,z) = | sum_{n=1..100} J0(alpha r rho_n) exp(-2i z gamma rho_n^2) rho_n drho |^2 alpha^2/pi
// One CTA per depth key; the 100 depth-dependent phasors are staged in shared memory
// and every thread sums them against J0 for its radii.
constexpr int kRhoTerms = 100;

__global__ void __launch_bounds__(256)
born_wolf_radial_kernel(double wave_length, int n_radial, const double *__restrict__ depths,
                        double *__restrict__ radial) {
    __shared__ double yr[kRhoTerms], yi[kRhoTerms], rho[kRhoTerms];
    const double NA = 1.4;
    const double k = 2.0 * M_PI / wave_length;
    const double alpha = k * NA;
    const double gamma = k * (NA / 2) * (NA / 2);
    const double z = depths[blockIdx.x];
    const double drho = 1.0 / kRhoTerms;
    for (int n = threadIdx.x; n < kRhoTerms; n += blockDim.x) {
        double rn = (n + 1) * drho;
        double s, c;
        sincos(-2.0 * z * gamma * rn * rn, &s, &c);
        yr[n] = c * rn * drho;
        yi[n] = s * rn * drho;
        rho[n] = rn;
    }
    __syncthreads();
    for (int ir = threadIdx.x; ir < n_radial; ir += blockDim.x) {
        double r = (double)ir * 1e-9;  // arange(0, cutoff, 1e-9)[ir]
        double sr = 0.0, si = 0.0;
#pragma unroll 4
        for (int n = 0; n < kRhoTerms; ++n) {
            double j = j0(r * alpha * rho[n]);
            sr += yr[n] * j;
            si += yi[n] * j;
        }
        radial[(size_t)blockIdx.x * n_radial + ir] = (sr * sr + si * si) * (alpha * alpha / M_PI);
    }
}

// Gaussian, _epifm.py:133-134 (depth independent).
__global__ void gaussian_radial_kernel(double width, int n_radial, int n_keys,
                                       double *__restrict__ radial) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_radial * n_keys) return;
    double r = (double)(idx % n_radial) * 1e-9;
    double q = r / width;
    radial[idx] = exp(-0.5 * (q * q)) / (2 * M_PI * width * width);
}

extern "C" int scb_psf_radial_build(int psf_type, double wave_length, double radial_width,
                                    int n_radial, int n_keys, const double *d_depths,
                                    double *d_radial, void *stream) {
    SCB_REQUIRE(d_radial != nullptr, SCB_E_NULL, "scb_psf_radial_build: d_radial is NULL");
    SCB_REQUIRE(n_radial >= 2 && n_keys >= 1, SCB_E_INVALID,
                "scb_psf_radial_build: n_radial=%d n_keys=%d", n_radial, n_keys);
    cudaStream_t s = (cudaStream_t)stream;
    if (psf_type == SCB_PSF_GAUSSIAN) {
        SCB_REQUIRE(radial_width > 0, SCB_E_INVALID,
                    "fluorophore.radial_width must be given for Gaussian type fluorophore.");
        int total = n_radial * n_keys;
        gaussian_radial_kernel<<<scb_grid_for(total, 256), 256, 0, s>>>(radial_width, n_radial, n_keys,
                                                                        d_radial);
    } else if (psf_type == SCB_PSF_BORN_WOLF) {
        SCB_REQUIRE(d_depths != nullptr, SCB_E_NULL, "scb_psf_radial_build: d_depths is NULL");
        SCB_REQUIRE(wave_length > 0, SCB_E_INVALID, "scb_psf_radial_build: wave_length=%g", wave_length);
        born_wolf_radial_kernel<<<n_keys, 256, 0, s>>>(wave_length, n_radial, d_depths, d_radial);
    } else {
        SCB_REQUIRE(false, SCB_E_INVALID, "scb_psf_radial_build: unknown psf_type %d", psf_type);
    }
    SCB_CUDA_LAUNCH_CHECK("scb_psf_radial_build");
    return 0;
}

// ------------------------------------------------------------------ Cartesian sample
// T[a][b] of radial_to_cartesian (_epifm.py:118-126): linear interpolation of the radial
// profile at R = sqrt(da^2 + db^2) nm, R clamped to the last radial sample.  da, db are
// exact integers, sqrt is correctly rounded, and the interpolation uses explicit IEEE
// operations (no FMA contraction) so the C oracle reproduces it bit for bit.
__device__ __forceinline__ double table_sample(const double *__restrict__ prof, int n_radial,
                                               int da, int db) {
    const int c = n_radial - 1;
    double R = sqrt((double)(da * da + db * db));
    if (R > (double)c) R = (double)c;
    int k = (int)R;
    if (k > c - 1) k = c - 1;
    double t = __dsub_rn(R, (double)k);
    double lo = prof[k], hi = prof[k + 1];
    return __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), t));
}

// Deterministic sum of one table row (fixed tree), used only to pick the table's scale.
__global__ void __launch_bounds__(256)
table_rowsum_kernel(const double *__restrict__ radial, int n_radial, double *__restrict__ rowsum) {
    __shared__ double part[256];
    const int side = 2 * (n_radial - 1) + 1;
    const int key = blockIdx.y, a = blockIdx.x;
    const double *prof = radial + (size_t)key * n_radial;
    const int c = n_radial - 1;
    double acc = 0.0;
    for (int b = threadIdx.x; b < side; b += blockDim.x) acc += table_sample(prof, n_radial, a - c, b - c);
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) rowsum[(size_t)key * side + a] = part[0];
}

// scale_k = 2^(60 - ilogb(sum T)): the quantised table then sums to < 2^61 (+ rounding).
__global__ void __launch_bounds__(256)
table_scale_kernel(const double *__restrict__ rowsum, int side, double *__restrict__ scale,
                   double *__restrict__ inv_scale) {
    __shared__ double part[256];
    const int key = blockIdx.x;
    double acc = 0.0;
    for (int a = threadIdx.x; a < side; a += blockDim.x) acc += rowsum[(size_t)key * side + a];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double total = part[0];
        int e = (total > 0.0 && isfinite(total)) ? ilogb(total) : 0;
        scale[key] = ldexp(1.0, 60 - e);
        inv_scale[key] = ldexp(1.0, e - 60);
    }
}

// Row pass: quantise one table row and write its inclusive prefix sum into the plain
// scratch table P[a+1][1..] ((side+1)^2, row major).  Row 0 and column 0 of S are zero.
// Block-wide scan: per-thread serial chunk + warp shuffles + one shared pass; int64 adds are
// associative so the result is exact.
constexpr int kScanThreads = 256;
constexpr int kBuildBatch = 16;      // tables integrated per pass (scratch: 32 MB each)

__global__ void __launch_bounds__(kScanThreads)
sat_rows_kernel(const double *__restrict__ radial, int n_radial, const double *__restrict__ scale,
                int64_t *__restrict__ plain) {
    __shared__ int64_t warp_tot[kScanThreads / 32];
    const int c = n_radial - 1;
    const int side = 2 * c + 1, cols = side + 1;
    const int key = blockIdx.y, a = blockIdx.x;  // a in [0, side]: a == side writes the zero row 0
    int64_t *S = plain + (size_t)key * cols * cols;
    if (a == side) {
        for (int b = threadIdx.x; b < cols; b += blockDim.x) S[b] = 0;
        return;
    }
    const double *prof = radial + (size_t)key * n_radial;
    const double sc = scale[key];
    const int per = (side + kScanThreads - 1) / kScanThreads;  // 8 for side = 1999
    const int b0 = threadIdx.x * per;
    int64_t local[16];
    int64_t run = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (i < per) {
            int b = b0 + i;
            int64_t q = 0;
            if (b < side) q = __double2ll_rn(__dmul_rn(table_sample(prof, n_radial, a - c, b - c), sc));
            run += q;
            local[i] = run;
        }
    }
    // exclusive scan of the per-thread totals
    int64_t incl = run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int64_t base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    base += incl - run;
    int64_t *row = S + (size_t)(a + 1) * cols;
    if (threadIdx.x == 0) row[0] = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (i < per) {
            int b = b0 + i;
            if (b < side) row[b + 1] = base + local[i];
        }
    }
}

// Column pass: P[a][b] += P[a-1][b] down each column; threads cover columns (coalesced).
__global__ void sat_cols_kernel(int cols, int n_keys, int64_t *__restrict__ plain) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    int key = blockIdx.y;
    if (b >= cols || key >= n_keys) return;
    int64_t *S = plain + (size_t)key * cols * cols + b;
    int64_t run = 0;
    // software-pipelined: loads of later rows do not depend on the running sum
    for (int a = 1; a < cols; a += 4) {
        int64_t v0 = S[(size_t)a * cols];
        int64_t v1 = (a + 1 < cols) ? S[(size_t)(a + 1) * cols] : 0;
        int64_t v2 = (a + 2 < cols) ? S[(size_t)(a + 2) * cols] : 0;
        int64_t v3 = (a + 3 < cols) ? S[(size_t)(a + 3) * cols] : 0;
        run += v0; S[(size_t)a * cols] = run;
        if (a + 1 < cols) { run += v1; S[(size_t)(a + 1) * cols] = run; }
        if (a + 2 < cols) { run += v2; S[(size_t)(a + 2) * cols] = run; }
        if (a + 3 < cols) { run += v3; S[(size_t)(a + 3) * cols] = run; }
    }
}

// Re-tile the plain table into the block layout (see SatLayout): one CTA per (phase block, key);
// writes are contiguous, reads gather every M-th sample (the scratch table is L2 resident).
//
// The optional second output is the "box table": entry (r, c) of block (pr, pc) is the box sum
// between slots (r-1, r) x (c-1, c) of that block (slot -1 = the zero sample before the table),
// converted to fp64 -- i.e. exactly the value (double)box that a pixel of a footprint with
// edge phases (pr, pc) multiplies by its weight.  A footprint whose edges are evenly spaced
// reads one dense rectangle of one box block and never touches the SAT itself.
__global__ void __launch_bounds__(256)
sat_blocks_kernel(const int64_t *__restrict__ plain, int64_t *__restrict__ sat, void *__restrict__ box,
                  int box_type, SatLayout L) {
    const int cols = L.side + 1;
    const int block = blockIdx.x, key = blockIdx.y;
    const int pr = block / L.modulus, pc = block - pr * L.modulus;
    const int64_t *P = plain + (size_t)key * cols * cols;
    const size_t at = (size_t)key * L.table_entries() + (size_t)block * L.block_entries();
    const int n = L.slots * L.slots;
    auto value = [&](int ir, int ic) -> int64_t {       // S at slot (ir, ic); slot -1 is zero
        if (ir < 0 || ic < 0) return 0;
        const int a = min(ir * L.modulus + pr, L.side), b = min(ic * L.modulus + pc, L.side);
        return P[(size_t)a * cols + b];
    };
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int ir = e / L.slots, ic = e - ir * L.slots;
        const int64_t here = value(ir, ic);
        sat[at + e] = here;
        if (box) {
            const int64_t sum = (here - value(ir - 1, ic)) - (value(ir, ic - 1) - value(ir - 1, ic - 1));
            if (box_type == SCB_F32) static_cast<float *>(box)[at + e] = __ll2float_rn(sum);
            else static_cast<double *>(box)[at + e] = __ll2double_rn(sum);
        }
    }
}

extern "C" int64_t scb_psf_sat_table_entries(int n_radial, int sat_modulus) {
    if (n_radial < 2) return 0;
    return scb_sat_layout(n_radial, sat_modulus).table_entries();
}

extern "C" int scb_psf_sat_slots(int n_radial, int sat_modulus) {
    if (n_radial < 2) return 0;
    return scb_sat_layout(n_radial, sat_modulus).slots;
}

extern "C" size_t scb_psf_sat_workspace_bytes(int n_radial, int n_keys) {
    if (n_radial < 2 || n_keys < 1) return 0;
    const size_t side = 2 * (size_t)(n_radial - 1) + 1;
    const size_t batch = n_keys < kBuildBatch ? n_keys : kBuildBatch;
    return ((size_t)n_keys * side + (size_t)n_keys) * sizeof(double) + 256 + batch * (side + 1) * (side + 1) * sizeof(int64_t);
}

extern "C" int scb_psf_sat_build(const double *d_radial, int n_radial, int n_keys, int sat_modulus,
                                 int64_t *d_sat, void *d_box, int box_type, double *d_inv_scale, void *d_workspace,
                                 size_t workspace_bytes, void *stream) {
    SCB_REQUIRE(!d_box || box_type == SCB_F32 || box_type == SCB_F64, SCB_E_INVALID, "scb_psf_sat_build: box_type=%d",
                box_type);
    SCB_REQUIRE(d_radial && d_sat && d_inv_scale && d_workspace, SCB_E_NULL,
                "scb_psf_sat_build: NULL pointer");
    SCB_REQUIRE(n_radial >= 2 && n_radial <= 2048 && n_keys >= 1, SCB_E_INVALID,
                "scb_psf_sat_build: n_radial=%d (2..2048) n_keys=%d", n_radial, n_keys);
    SCB_REQUIRE(workspace_bytes >= scb_psf_sat_workspace_bytes(n_radial, n_keys), SCB_E_WORKSPACE,
                "scb_psf_sat_build: workspace %zu < %zu", workspace_bytes,
                scb_psf_sat_workspace_bytes(n_radial, n_keys));
    SCB_REQUIRE(n_keys <= 65535, SCB_E_INVALID, "scb_psf_sat_build: n_keys=%d > 65535", n_keys);
    cudaStream_t s = (cudaStream_t)stream;
    SCB_REQUIRE(sat_modulus >= 1 && sat_modulus <= 4096, SCB_E_INVALID, "scb_psf_sat_build: sat_modulus=%d", sat_modulus);
    const int side = 2 * (n_radial - 1) + 1, cols = side + 1;
    const SatLayout L = scb_sat_layout(n_radial, sat_modulus);
    SCB_REQUIRE(L.table_entries() < ((long long)1 << 31), SCB_E_UNSUPPORTED,
                "scb_psf_sat_build: table of %lld entries (modulus %d, %d slots) is too large", L.table_entries(),
                L.modulus, L.slots);
    double *rowsum = (double *)d_workspace;
    double *scale = rowsum + (size_t)n_keys * side;
    int64_t *plain = (int64_t *)(((uintptr_t)(scale + n_keys) + 255) & ~(uintptr_t)255);
    table_rowsum_kernel<<<dim3(side, n_keys), 256, 0, s>>>(d_radial, n_radial, rowsum);
    table_scale_kernel<<<n_keys, 256, 0, s>>>(rowsum, side, scale, d_inv_scale);
    for (int first = 0; first < n_keys; first += kBuildBatch) {
        const int nb = n_keys - first < kBuildBatch ? n_keys - first : kBuildBatch;
        sat_rows_kernel<<<dim3(side + 1, nb), kScanThreads, 0, s>>>(d_radial + (size_t)first * n_radial, n_radial,
                                                                   scale + first, plain);
        sat_cols_kernel<<<dim3((cols + 127) / 128, nb), 128, 0, s>>>(cols, nb, plain);
        sat_blocks_kernel<<<dim3(L.modulus * L.modulus, nb), 256, 0, s>>>(
            plain, d_sat + (size_t)first * L.table_entries(),
            d_box ? (void *)((char *)d_box + (size_t)first * L.table_entries() * (box_type == SCB_F32 ? 4 : 8)) : nullptr,
            box_type, L);
    }
    SCB_CUDA_LAUNCH_CHECK("scb_psf_sat_build");
    return 0;
}
