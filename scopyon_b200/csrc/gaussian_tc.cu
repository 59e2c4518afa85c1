// Separable-Gaussian PSF rendering on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Reference: fluorophore.type == 'Gaussian' (/root/reference/src/scopyon/_epifm.py:133-134)
// rendered through overlay_signal_ (:224-282).  For a Gaussian the 2-D table factorises,
//   T[a][b] ~= g(a) g(b),   g(a) = exp(-(a-c)^2 nm^2 / 2 sigma^2) / (sqrt(2 pi) sigma),
// (up to the reference's radial linear interpolation: <= 7.2e-6 of the peak for
// sigma >= 100 nm, measured against the reference table), so a pixel box sum is a product of
// two 1-D sums of the same 1-nm samples, Ex_s(i) * Ey_s(j), and a 128 x 128 screen tile is
// the dense contraction
//   D[i][j] = sum_s  (w_s Ex_s(i)) * Ey_s(j)          (M = N = 128, K = spots binned to the tile).
// One CTA owns a tile: CUDA cores build the operand panels in shared memory (prefix table of g
// staged by a TMA bulk copy), one thread issues tcgen05.mma.kind::tf32 with the accumulator in
// TMEM, and each fp32 operand is split into two tf32 terms (hi + lo; three MMAs per K step:
// hi*hi + hi*lo + lo*hi) so the product keeps ~21 mantissa bits.  Double buffering through
// mbarriers lets the panel build of chunk c+1 overlap the MMAs of chunk c.
//
// Accuracy target of this path: 1e-5 of the image maximum (north_star); the SAT path
// (render.cu) is exact and remains the default for every PSF including the Gaussian.
#include "binning.cuh"

namespace {

constexpr int TM = 128;                 // tile edge = UMMA M = UMMA N
constexpr int KC = 16;                  // spots per chunk = two K = 8 steps
constexpr int kThreadsTC = 512;           // 16 warps: one per chunk spot in step A
constexpr int kMaxFoot = 64;            // footprint rows / columns per spot the panels can hold
constexpr int kSortCapTC = 2048;        // spots per tile ordered in shared memory per segment
constexpr int kFootPitch = kMaxFoot + 8; // row pitch of the partial-sum arrays: 72 floats -> spots 0..3 hit disjoint banks
constexpr int kPanelBytes = KC * TM * 4;            // one operand panel (8 KB)
// K-major operand panels without swizzle (the canonical layout of cute::UMMA::Layout_K_INTER):
// a core matrix is 8 rows (pixels) x 16 bytes (4 spots) = 128 contiguous bytes; the second
// half of a K = 8 step lies kLBO further, the next 8 rows kSBO further, the next 8 spots
// kKGroup further.  (MN-major tf32 operands return zeros on this part -- see
// tools/probes/umma_probe.cu, which also verified this layout element by element.)
constexpr uint32_t kLBO = 128;
constexpr uint32_t kSBO = 256;
constexpr uint32_t kKGroup = (TM / 8) * kSBO;       // 4096 bytes per 8 spots

struct TcSmem {
    double G[2048];                                  // prefix sums of g on the 1-nm grid (TMA staged)
    float ex_hi[KC][kFootPitch], ex_lo[KC][kFootPitch];  // w_s * Ex_s split into tf32 hi + lo
    float ey_hi[KC][kFootPitch], ey_lo[KC][kFootPitch];
    int r0[KC], nr[KC], c0[KC], nc[KC];              // footprint rectangle of each chunk spot inside the tile
    int ids[kSortCapTC];
    alignas(128) unsigned char panels[2][4][kPanelBytes];   // [buffer][A_hi, A_lo, B_hi, B_lo]
    alignas(8) unsigned long long mbar_mma[2];
    alignas(8) unsigned long long mbar_tma;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_load(void *dst, const void *src, uint32_t bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// UMMA shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor):
// start address >> 4 in bits [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4
// in [32,46), descriptor version 1 in [46,48).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bit 4), A = B = TF32 (2 << 7,
// 2 << 10), both K-major (bits 15, 16 clear), N >> 3 in [17,23), M >> 4 in [24,29).
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TM >> 3) << 17) |
                                ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(kInstrDesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(void *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr)
        : "memory");
}

__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

template <typename OutT>
__global__ void __launch_bounds__(kThreadsTC)
gaussian_tc_kernel(Geo g, const SpotRec *__restrict__ spots, const uint32_t *__restrict__ edges, int edge_cap,
                   const int *__restrict__ tile_start, const int *__restrict__ pair_spot,
                   const double *__restrict__ prefix, int prefix_len, OutT *__restrict__ out, int accumulate) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem &sm = *reinterpret_cast<TcSmem *>(smem_raw);

    const int tile = blockIdx.x;
    const int ti = tile / g.ntj, tj = tile - ti * g.ntj;
    const int row0 = ti * TM, col0 = tj * TM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seg_begin = tile_start[tile * g.stripes], seg_end = tile_start[(tile + 1) * g.stripes];
    const int n_spots = seg_end - seg_begin;

    if (n_spots == 0) {   // nothing lands on this tile
        if (!accumulate)
            for (int p = tid; p < TM * TM; p += kThreadsTC) {
                const int i = row0 + p / TM, j = col0 + p % TM;
                if (i < g.n_w && j < g.n_h) out[(size_t)i * g.n_h + j] = (OutT)0;
            }
        return;
    }

    // ---- one-time setup: barriers, TMEM columns for the 128 x 128 fp32 accumulator, prefix table
    if (tid == 0) {
        mbar_init(&sm.mbar_mma[0], 1);
        mbar_init(&sm.mbar_mma[1], 1);
        mbar_init(&sm.mbar_tma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "r"((uint32_t)TM));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm.tmem_base;
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)prefix_len * 8u;
        mbar_expect_tx(&sm.mbar_tma, bytes);
        tma_bulk_load(sm.G, prefix, bytes, &sm.mbar_tma);   // TMA-staged PSF lookup table
    }

    int chunk_index = 0;   // chunks issued so far (drives buffer choice and barrier phases)
    for (int seg = seg_begin; seg < seg_end; seg += kSortCapTC) {
        const int n_seg = min(kSortCapTC, seg_end - seg);
        // ---- order the segment by spot index (bitonic sort) for a reproducible K order
        int pow2 = 1;
        while (pow2 < n_seg) pow2 <<= 1;
        __syncthreads();
        for (int t = tid; t < pow2; t += kThreadsTC) sm.ids[t] = t < n_seg ? pair_spot[seg + t] : 0x7fffffff;
        __syncthreads();
        for (int k = 2; k <= pow2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < pow2; t += kThreadsTC) {
                    const int partner = t ^ j;
                    if (partner > t) {
                        const int a = sm.ids[t], b = sm.ids[partner];
                        const bool up = (t & k) == 0;
                        if ((a > b) == up) { sm.ids[t] = b; sm.ids[partner] = a; }
                    }
                }
                __syncthreads();
            }
        if (seg == seg_begin) mbar_wait(&sm.mbar_tma, 0);   // prefix table has landed

        for (int base = 0; base < n_seg; base += KC, ++chunk_index) {
            const int buf = chunk_index & 1;
            // the MMAs that last read this buffer (chunk_index - 2) must have completed
            if (chunk_index >= 2) mbar_wait(&sm.mbar_mma[buf], ((chunk_index >> 1) - 1) & 1);

            // ---- step A: 1-D partial sums of every chunk spot (warp w: spot w)
            {
                const int s = warp;
                int r0 = 0, nr = 0, c0 = 0, nc = 0;
                if (base + s < n_seg) {
                    const int sid = sm.ids[base + s];
                    const SpotRec rec = spots[sid];
                    const int r_lo = max(rec.imin, row0), r_hi = min(rec.imax, row0 + TM);
                    const int c_lo = max(rec.jmin, col0), c_hi = min(rec.jmax, col0 + TM);
                    r0 = r_lo - row0; nr = r_hi - r_lo; c0 = c_lo - col0; nc = c_hi - c_lo;
                    const uint32_t *e = edges + (size_t)sid * 2 * edge_cap;
                    for (int k = lane; k < nr; k += 32) {
                        const int idx = r_lo - rec.imin + k;
                        const float a = (float)(rec.w * (sm.G[e[idx + 1]] - sm.G[e[idx]]));
                        const float hi = to_tf32(a);
                        sm.ex_hi[s][k] = hi;
                        sm.ex_lo[s][k] = a - hi;
                    }
                    for (int k = lane; k < nc; k += 32) {
                        const int idx = edge_cap + c_lo - rec.jmin + k;
                        const float b = (float)(sm.G[e[idx + 1]] - sm.G[e[idx]]);
                        const float hi = to_tf32(b);
                        sm.ey_hi[s][k] = hi;
                        sm.ey_lo[s][k] = b - hi;
                    }
                }
                if (lane == 0) { sm.r0[s] = r0; sm.nr[s] = nr; sm.c0[s] = c0; sm.nc[s] = nc; }
            }
            __syncthreads();

            // ---- step B: operand panels.  A lane owns one tile row of one 4-spot group and writes
            // that row of the core matrix (4 spots x 4 B) with one 16-byte store; a warp task covers
            // 32 consecutive rows (four core matrices, 512 contiguous bytes apart by kSBO).
            for (int task = warp; task < 32; task += kThreadsTC / 32) {
                const bool is_b = task >= 16;
                const int kg = (task >> 3) & 1, half = (task >> 2) & 1, rq = task & 3;
                const int i = rq * 32 + lane, s0 = kg * 8 + half * 4;
                float hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int s = s0 + q;
                    const int k = i - (is_b ? sm.c0[s] : sm.r0[s]);
                    const bool inside = (unsigned)k < (unsigned)(is_b ? sm.nc[s] : sm.nr[s]);
                    const int kk = inside ? k : 0;
                    const float h = is_b ? sm.ey_hi[s][kk] : sm.ex_hi[s][kk];
                    const float l = is_b ? sm.ey_lo[s][kk] : sm.ex_lo[s][kk];
                    hi[q] = inside ? h : 0.0f;
                    lo[q] = inside ? l : 0.0f;
                }
                const int off = kg * (int)kKGroup + (i >> 3) * (int)kSBO + half * (int)kLBO + (i & 7) * 16;
                *reinterpret_cast<float4 *>(&sm.panels[buf][is_b ? 2 : 0][off]) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(&sm.panels[buf][is_b ? 3 : 1][off]) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();

            // ---- MMA issue: one thread, three tf32 products per K step
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_u32(sm.panels[buf][0]), a_lo = smem_u32(sm.panels[buf][1]);
                const uint32_t b_hi = smem_u32(sm.panels[buf][2]), b_lo = smem_u32(sm.panels[buf][3]);
#pragma unroll
                for (int kg = 0; kg < KC / 8; ++kg) {
                    const uint32_t o = (uint32_t)kg * kKGroup;
                    umma_tf32(tmem, umma_desc(a_hi + o), umma_desc(b_hi + o), (chunk_index | kg) != 0);
                    umma_tf32(tmem, umma_desc(a_hi + o), umma_desc(b_lo + o), 1u);
                    umma_tf32(tmem, umma_desc(a_lo + o), umma_desc(b_hi + o), 1u);
                }
                umma_commit(&sm.mbar_mma[buf]);
            }
        }
    }

    // ---- wait for the last MMAs of both buffers, then drain the accumulator
    {
        const int last = chunk_index - 1;
        mbar_wait(&sm.mbar_mma[last & 1], (last >> 1) & 1);
        if (last >= 1) mbar_wait(&sm.mbar_mma[(last - 1) & 1], ((last - 1) >> 1) & 1);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {   // warp w reads TMEM lanes 32w .. 32w+31 = tile rows
        const int i = row0 + tid;
#pragma unroll 1
        for (int cb = 0; cb < TM / 32; ++cb) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32), v);
            if (i < g.n_w) {
                OutT *dst = out + (size_t)i * g.n_h + col0 + cb * 32;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (col0 + cb * 32 + c < g.n_h) {
                        const float val = __uint_as_float(v[c]);
                        dst[c] = accumulate ? (OutT)((double)dst[c] + (double)val) : (OutT)val;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TM));
}

}  // namespace

extern "C" size_t scb_gaussian_tc_workspace_bytes(const scb_geometry *geom, int64_t n_spots) {
    if (check_geometry(geom) != 0 || n_spots < 0) return 0;
    Geo g = make_geo(geom, TM, TM);
    return carve(g, n_spots, nullptr).bytes;
}

extern "C" int scb_render_gaussian_tc(const scb_geometry *geom, int64_t n_spots, const double *d_x,
                                      const double *d_y, const double *d_weight, const double *d_prefix,
                                      void *d_out, int out_type, int accumulate, void *d_workspace,
                                      size_t workspace_bytes, int32_t *d_errors, void *stream) {
    int rc = check_geometry(geom);
    if (rc) return rc;
    SCB_REQUIRE(n_spots >= 0 && n_spots < (int64_t)1 << 31, SCB_E_INVALID, "n_spots=%lld", (long long)n_spots);
    SCB_REQUIRE(d_out && d_workspace && d_errors && d_prefix, SCB_E_NULL, "scb_render_gaussian_tc: NULL pointer");
    SCB_REQUIRE(n_spots == 0 || (d_x && d_y && d_weight), SCB_E_NULL, "scb_render_gaussian_tc: NULL spot pointer");
    SCB_REQUIRE(out_type == SCB_F32 || out_type == SCB_F64, SCB_E_INVALID, "out_type=%d", out_type);
    Geo g = make_geo(geom, TM, TM);
    g.modulus = 1;                       // plain sample indices (special_edges == 0): no SAT in this path
    SCB_REQUIRE(g.side + 1 <= 2048, SCB_E_UNSUPPORTED, "scb_render_gaussian_tc: table side %d > 2047", g.side);
    SCB_REQUIRE(ceil(g.sw / g.pl) + 2 <= kMaxFoot, SCB_E_UNSUPPORTED,
                "scb_render_gaussian_tc: footprint of %g pixels exceeds %d (pixel_length too small)",
                ceil(g.sw / g.pl) + 2, kMaxFoot);
    Workspace w = carve(g, n_spots, d_workspace);
    SCB_REQUIRE(workspace_bytes >= w.bytes, SCB_E_WORKSPACE, "scb_render_gaussian_tc: workspace %zu < %zu",
                workspace_bytes, w.bytes);
    cudaStream_t s = (cudaStream_t)stream;
    const int n_tiles = g.nti * g.ntj;
    SCB_CUDA(cudaMemsetAsync(w.tile_count, 0, (size_t)((char *)w.tile_start - (char *)w.tile_count), s));
    if (n_spots > 0) {
        // depth plays no role for the Gaussian (depth-independent PSF): x doubles as a dummy depth
        spot_prepare_kernel<SCB_PREPARE_CTAS, false><<<scb_grid_for(n_spots, 256), 256, 0, s>>>(
            g, n_spots, 1, nullptr, d_x, d_y, d_weight, nullptr, nullptr, w.spots, w.tile_count, nullptr, d_errors,
            nullptr, 0, nullptr, nullptr, nullptr);
        dim3 egrid, eblock;
        edges_launch_shape(w.edge_cap, n_spots, egrid, eblock);
        spot_edges_kernel<<<egrid, eblock, 0, s>>>(g, n_spots, w.spots, w.edges, w.edge_cap);
    }
    tile_scan_kernel<false><<<kScanCtas, 1024, 0, s>>>(n_tiles, g.stripes, w.tile_count, w.tile_start);
    if (n_spots > 0) {
        tile_fill_kernel<<<scb_grid_for(n_spots, 256), 256, 0, s>>>(g, n_spots, w.spots, w.tile_start, w.tile_cursor,
                                                                   w.pair_spot);
    }
    const size_t smem = sizeof(TcSmem) + 1024;
    if (out_type == SCB_F32) {
        SCB_CUDA(cudaFuncSetAttribute(gaussian_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gaussian_tc_kernel<float><<<n_tiles, kThreadsTC, smem, s>>>(g, w.spots, w.edges, w.edge_cap, w.tile_start,
                                                                    w.pair_spot, d_prefix, g.side + 1, (float *)d_out,
                                                                    accumulate);
    } else {
        SCB_CUDA(cudaFuncSetAttribute(gaussian_tc_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gaussian_tc_kernel<double><<<n_tiles, kThreadsTC, smem, s>>>(g, w.spots, w.edges, w.edge_cap, w.tile_start,
                                                                     w.pair_spot, d_prefix, g.side + 1, (double *)d_out,
                                                                     accumulate);
    }
    SCB_CUDA_LAUNCH_CHECK("scb_render_gaussian_tc");
    return 0;
}
