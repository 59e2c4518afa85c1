// Fused detector pass: background + QE, shot noise, EM gain, readout noise, ADC.
//
// Reference: _EPIFMSimulator.__detector_output, CMOS / EMCCD / CCD and
// __get_analog_to_digital_converter_counts (/root/reference/src/scopyon/_epifm.py:
// 1430-1484, 329-433, 926-963).  The reference loops over pixels in Python (Poisson per
// pixel; for EMCCD it builds a 3 000-50 000-entry pmf per pixel and calls rng.choice).
// Here one thread streams four pixels: one 16-byte load, two Philox4x32-10 blocks
// (shot + readout), one 16-byte store.  Draws are keyed (seed; frame, pixel quad) so a
// frame's noise does not depend on which GPU or launch produced it.
//
// Samplers (statistical parity, see tests/test_detector_gpu.py):
//   Poisson(E)   E < 12: inversion by sequential search;  E >= 12: PTRS (Hoermann 1993)
//   EMCCD        n ~ Poisson(E);  S = rint(g * Gamma(n, 1)), redrawn while outside the
//                reference's support [g*int(E-sigma)+, g*int(E+sigma)), sigma = 5 sqrt(E) + 10
//                (the reference pmf is exactly this Poisson->Gamma mixture evaluated at
//                integer S and truncated, _epifm.py:365-391)
//   readout      N(0, readout_noise) (EMCCD, CCD) or Walker-alias draw from the CMOS
//                read-noise table (CMOS)
#include "scb_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxAlias = 1024;

// Per-pixel overflow stream for rejection loops: Philox keyed (seed; pixel, frame), tag
// EXTRA, block counter in the low bits of the tag word.
struct PixelRng {
    uint64_t seed, pixel;
    uint32_t frame_lo, frame_hi, block;
    Philox4 buf;
    int used;
    __device__ PixelRng(uint64_t s, uint64_t p, uint64_t frame)
        : seed(s), pixel(p), frame_lo((uint32_t)frame), frame_hi((uint32_t)(frame >> 32)), block(0), used(4) {}
    __device__ uint32_t next() {
        if (used == 4) {
            buf = philox4x32_10((uint32_t)pixel, (uint32_t)(pixel >> 32), frame_lo,
                                (SCB_TAG_EXTRA ^ frame_hi) + block, (uint32_t)seed, (uint32_t)(seed >> 32));
            ++block;
            used = 0;
        }
        uint32_t r = used == 0 ? buf.x : used == 1 ? buf.y : used == 2 ? buf.z : buf.w;
        ++used;
        return r;
    }
};

// Poisson by inversion (sequential search), lambda < 12.
__device__ __forceinline__ float poisson_small(float lambda, uint32_t r) {
    const float u = u01_open_low(r);
    float p = __expf(-lambda);
    float s = p;
    int k = 0;
    while (u > s && k < 128) {
        ++k;
        p *= __fdividef(lambda, (float)k);
        s += p;
    }
    return (float)k;
}

// PTRS, W. Hoermann, "The transformed rejection method for generating Poisson random
// variables" (1993); the same algorithm numpy uses for lam >= 10.
__device__ __noinline__ double poisson_ptrs(double lambda, PixelRng &rng) {
    const double slam = sqrt(lambda);
    const double loglam = log(lambda);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int trial = 0; trial < 1000; ++trial) {
        const double U = u01_open_low_53(rng.next(), rng.next()) - 0.5;
        const double V = u01_open_low_53(rng.next(), rng.next());
        const double us = 0.5 - fabs(U);
        const double k = floor((2.0 * a / us + b) * U + lambda + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lambda + k * loglam - lgamma(k + 1.0)) return k;
    }
    return floor(lambda);
}

__device__ __forceinline__ double poisson_any(double lambda, uint32_t r, PixelRng &rng) {
    if (!(lambda > 0.0)) return 0.0;      // E <= 0 (or NaN): no signal, _epifm.py:395-396
    if (lambda < 12.0) return (double)poisson_small((float)lambda, r);
    return poisson_ptrs(lambda, rng);
}

// Gamma(shape = n integer >= 1, scale = 1).
__device__ __noinline__ double gamma_int(double n, PixelRng &rng) {
    if (n < 6.0) {
        float prod = 1.0f;
        for (int i = 0; i < (int)n; ++i) prod *= u01_open_low(rng.next());
        return -(double)logf(prod);
    }
    // Marsaglia & Tsang (2000)
    const double d = n - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (int trial = 0; trial < 1000; ++trial) {
        float x0, x1;
        box_muller(rng.next(), rng.next(), x0, x1);
        const double x = (double)x0;
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = u01_open_low_53(rng.next(), rng.next());
        if (u < 1.0 - 0.0331 * (x * x) * (x * x)) return d * v;
        if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v;
    }
    return d;
}

// EMCCD signal, _epifm.py:372-400: truncated Poisson->Gamma mixture on integer S.
__device__ __noinline__ double emccd_signal(double E, double gain, uint32_t r, PixelRng &rng) {
    if (!(E > 0.0)) return 0.0;
    const double sigma = sqrt(E) * 5.0 + 10.0;
    const double s_min = fmax(0.0, gain * trunc(E - sigma));
    const double s_max = gain * trunc(E + sigma);
    for (int trial = 0; trial < 64; ++trial) {
        const double n = poisson_any(E, trial == 0 ? r : rng.next(), rng);
        const double S = (n > 0.0) ? rint(gain * gamma_int(n, rng)) : 0.0;
        if (S >= s_min && S < s_max) return S;
    }
    return fmin(fmax(rint(gain * E), s_min), s_max - 1.0);
}

template <typename T>
__device__ __forceinline__ void load4(const T *p, T v[4]) {
    if constexpr (sizeof(T) == 4) {
        float4 q = __ldcs(reinterpret_cast<const float4 *>(p));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
        double2 a = __ldcs(reinterpret_cast<const double2 *>(p));
        double2 b = __ldcs(reinterpret_cast<const double2 *>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
}

template <typename T>
__device__ __forceinline__ void store4(T *p, const T v[4]) {
    if constexpr (sizeof(T) == 4) {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    } else {
        __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1]));
        __stcs(reinterpret_cast<double2 *>(p) + 1, make_double2(v[2], v[3]));
    }
}

struct DetArgs {
    uint64_t seed, frame;
    scb_detector det;
    int64_t n_pix;
    int32_t n_h;
    int n_alias;
    double adc_max, pow2bit;
    const void *photons, *offset;
    const scb_alias_entry *alias;
    void *adc, *expectation;
    const void *in_signal, *in_noise;
    void *out_signal, *out_noise;
};

// ADC, _epifm.py:1472-1484 with gain = fullwell / (2^bit - offset), _epifm.py:958.
__device__ __forceinline__ double adc_convert(double pe, double offset, const DetArgs &a) {
    if (pe > a.det.fullwell) pe = a.det.fullwell;
    const double gain = __ddiv_rn(a.det.fullwell, __dsub_rn(a.pow2bit, offset));
    double v = __dadd_rn(__ddiv_rn(pe, gain), offset);
    if (v > a.adc_max) v = a.adc_max;
    if (v < 0.0) v = 0.0;
    return v;
}
__device__ __forceinline__ float adc_convert(float pe, float offset, const DetArgs &a) {
    pe = fminf(pe, (float)a.det.fullwell);
    const float inv_gain = __fdividef((float)a.pow2bit - offset, (float)a.det.fullwell);
    float v = fmaf(pe, inv_gain, offset);
    return fminf(fmaxf(v, 0.0f), (float)a.adc_max);
}

template <typename T, int DET>
__global__ void __launch_bounds__(kThreads)
detector_kernel(DetArgs a) {
    __shared__ scb_alias_entry s_alias[DET == SCB_DET_CMOS ? kMaxAlias : 1];
    if (DET == SCB_DET_CMOS) {
        for (int i = threadIdx.x; i < a.n_alias; i += kThreads) s_alias[i] = a.alias[i];
        __syncthreads();
    }
    const T *photons = (const T *)a.photons;
    const T *offset = (const T *)a.offset;
    T *adc = (T *)a.adc;
    const int64_t n_quads = (a.n_pix + 3) >> 2;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const uint32_t f_lo = (uint32_t)a.frame, f_hi = (uint32_t)(a.frame >> 32);
    const T qe = (T)a.det.qe;
    const T bg = a.det.background_on ? (T)a.det.background : (T)0;

    for (int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x; q < n_quads; q += (int64_t)gridDim.x * kThreads) {
        const int64_t p0 = q << 2;
        const bool full = p0 + 4 <= a.n_pix;
        T ph[4], off[4], sig[4], noi[4], out[4], ex[4];
        if (full) {
            load4(photons + p0, ph);
        } else {
            for (int i = 0; i < 4; ++i) ph[i] = (p0 + i < a.n_pix) ? photons[p0 + i] : (T)0;
        }
        // ADC offset: scalar, per column (axis-1 index), or per pixel  (_epifm.py:941-952)
        if (a.det.fpn_type == SCB_FPN_NONE) {
            for (int i = 0; i < 4; ++i) off[i] = (T)a.det.adc_offset;
        } else if (a.det.fpn_type == SCB_FPN_PIXEL) {
            if (full) load4(offset + p0, off);
            else for (int i = 0; i < 4; ++i) off[i] = (p0 + i < a.n_pix) ? offset[p0 + i] : (T)0;
        } else {
            const int j0 = (int)(p0 % a.n_h);
            if ((a.n_h & 3) == 0 && full) {   // the quad stays inside one image row
                for (int i = 0; i < 4; ++i) off[i] = offset[j0 + i];
            } else {
                for (int i = 0; i < 4; ++i) off[i] = offset[(j0 + i) % a.n_h];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) ex[i] = qe * (ph[i] + bg);    // _epifm.py:1438-1441

        // ---- shot noise (+ EM gain)
        if (a.in_signal) {
            const T *in = (const T *)a.in_signal;
            for (int i = 0; i < 4; ++i) sig[i] = (p0 + i < a.n_pix) ? in[p0 + i] : (T)0;
        } else {
            const Philox4 r = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), f_lo, SCB_TAG_SHOT ^ f_hi, k0, k1);
            const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float lam = (float)ex[i];
                // fast path: small expectation; an EMCCD pixel with zero photoelectrons
                // stays zero (S = 0 lies inside the reference's support whenever E < 12)
                const float n_small = (lam > 0.0f && lam < 12.0f) ? poisson_small(lam, rw[i]) : 0.0f;
                if (lam < 12.0f && (DET != SCB_DET_EMCCD || n_small == 0.0f)) {
                    sig[i] = (T)n_small;
                } else {
                    PixelRng rng(a.seed, (uint64_t)(p0 + i), a.frame);
                    if (DET == SCB_DET_EMCCD) sig[i] = (T)emccd_signal((double)ex[i], a.det.emgain, rw[i], rng);
                    else sig[i] = (T)poisson_any((double)ex[i], rw[i], rng);
                }
            }
        }
        // ---- readout noise
        if (a.in_noise) {
            const T *in = (const T *)a.in_noise;
            for (int i = 0; i < 4; ++i) noi[i] = (p0 + i < a.n_pix) ? in[p0 + i] : (T)0;
        } else {
            const Philox4 r = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), f_lo, SCB_TAG_READ ^ f_hi, k0, k1);
            if (DET == SCB_DET_CMOS) {
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint64_t prod = (uint64_t)rw[i] * (uint32_t)a.n_alias;
                    const scb_alias_entry e = s_alias[(uint32_t)(prod >> 32)];
                    const float frac = (float)(uint32_t)prod * 2.3283064365386963e-10f;
                    noi[i] = (T)(frac < e.threshold ? e.value : e.alias_value);
                }
            } else if (a.det.readout_noise > 0.0) {   // _epifm.py:360-362
                float n0, n1, n2, n3;
                box_muller(r.x, r.y, n0, n1);
                box_muller(r.z, r.w, n2, n3);
                const T rn = (T)a.det.readout_noise;
                noi[0] = rn * (T)n0; noi[1] = rn * (T)n1; noi[2] = rn * (T)n2; noi[3] = rn * (T)n3;
            } else {
                noi[0] = noi[1] = noi[2] = noi[3] = (T)0;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = adc_convert(sig[i] + noi[i], off[i], a);   // _epifm.py:1464-1469

        if (full) {
            store4(adc + p0, out);
            if (a.expectation) store4((T *)a.expectation + p0, ex);
            if (a.out_signal) store4((T *)a.out_signal + p0, sig);
            if (a.out_noise) store4((T *)a.out_noise + p0, noi);
        } else {
            for (int i = 0; i < 4 && p0 + i < a.n_pix; ++i) {
                adc[p0 + i] = out[i];
                if (a.expectation) ((T *)a.expectation)[p0 + i] = ex[i];
                if (a.out_signal) ((T *)a.out_signal)[p0 + i] = sig[i];
                if (a.out_noise) ((T *)a.out_noise)[p0 + i] = noi[i];
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
adc_offsets_kernel(uint64_t seed, int64_t n, double adc0, double fpn_count, T *__restrict__ offset) {
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    Philox4 r = philox_at(seed, (uint64_t)i, 0u, SCB_TAG_FPN);
    float n0, n1;
    box_muller(r.x, r.y, n0, n1);
    offset[i] = (T)rint(adc0 + fpn_count * (double)n0);   // numpy.rint(rng.normal(ADC0, count)), _epifm.py:946,950-952
}

template <typename T, int DET>
void launch_detector(const DetArgs &a, cudaStream_t s) {
    const int64_t n_quads = (a.n_pix + 3) >> 2;
    int64_t blocks = (n_quads + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)SCB_SM_COUNT * 8;   // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    detector_kernel<T, DET><<<(unsigned)blocks, kThreads, 0, s>>>(a);
}

}  // namespace

extern "C" int scb_adc_offsets(uint64_t seed, int64_t n, double adc0, double fpn_count, void *d_offset,
                               int elem_type, void *stream) {
    SCB_REQUIRE(d_offset != nullptr, SCB_E_NULL, "scb_adc_offsets: d_offset is NULL");
    SCB_REQUIRE(n >= 0, SCB_E_INVALID, "scb_adc_offsets: n=%lld", (long long)n);
    SCB_REQUIRE(elem_type == SCB_F32 || elem_type == SCB_F64, SCB_E_INVALID, "elem_type=%d", elem_type);
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (elem_type == SCB_F32)
        adc_offsets_kernel<float><<<scb_grid_for(n, kThreads), kThreads, 0, s>>>(seed, n, adc0, fpn_count, (float *)d_offset);
    else
        adc_offsets_kernel<double><<<scb_grid_for(n, kThreads), kThreads, 0, s>>>(seed, n, adc0, fpn_count, (double *)d_offset);
    SCB_CUDA_LAUNCH_CHECK("scb_adc_offsets");
    return 0;
}

extern "C" int scb_detector_adc(uint64_t seed, uint64_t frame, const scb_detector *det, int32_t n_w, int32_t n_h,
                                int elem_type, const void *d_photons, const void *d_offset,
                                const scb_alias_entry *d_cmos_alias, int n_alias, void *d_adc,
                                void *d_expectation, const void *d_in_signal, const void *d_in_noise,
                                void *d_out_signal, void *d_out_noise, void *stream) {
    SCB_REQUIRE(det && d_photons && d_adc, SCB_E_NULL, "scb_detector_adc: NULL pointer");
    SCB_REQUIRE(n_w > 0 && n_h > 0, SCB_E_INVALID, "scb_detector_adc: image %d x %d", n_w, n_h);
    SCB_REQUIRE(elem_type == SCB_F32 || elem_type == SCB_F64, SCB_E_INVALID, "elem_type=%d", elem_type);
    SCB_REQUIRE(det->type == SCB_DET_CMOS || det->type == SCB_DET_EMCCD || det->type == SCB_DET_CCD, SCB_E_INVALID,
                "Unknown detector type was given [%d]. Use either one of 'CMOS', 'CCD' or 'EMCCD'.", det->type);
    SCB_REQUIRE(det->fpn_type == SCB_FPN_NONE || det->fpn_type == SCB_FPN_PIXEL || det->fpn_type == SCB_FPN_COLUMN,
                SCB_E_INVALID, "FPN type [%d] is invalid ['pixel', 'column' or 'none']", det->fpn_type);
    SCB_REQUIRE(det->fpn_type == SCB_FPN_NONE || d_offset, SCB_E_NULL, "scb_detector_adc: offset map required");
    SCB_REQUIRE(det->bit >= 1 && det->bit <= 32 && det->fullwell > 0, SCB_E_INVALID, "bit=%d fullwell=%g", det->bit,
                det->fullwell);
    const bool need_alias = det->type == SCB_DET_CMOS && d_in_noise == nullptr;
    SCB_REQUIRE(!need_alias || (d_cmos_alias && n_alias >= 1 && n_alias <= kMaxAlias), SCB_E_INVALID,
                "scb_detector_adc: CMOS needs an alias table of 1..%d entries (got %d)", kMaxAlias, n_alias);
    const size_t align = elem_type == SCB_F32 ? 16 : 16;
    SCB_REQUIRE(((uintptr_t)d_photons % align) == 0 && ((uintptr_t)d_adc % align) == 0, SCB_E_INVALID,
                "scb_detector_adc: image buffers must be 16-byte aligned");
    DetArgs a;
    a.seed = seed; a.frame = frame; a.det = *det;
    a.n_pix = (int64_t)n_w * n_h; a.n_h = n_h;
    a.n_alias = need_alias ? n_alias : 0;
    a.pow2bit = ldexp(1.0, det->bit);
    a.adc_max = a.pow2bit - 1.0;
    a.photons = d_photons; a.offset = d_offset; a.alias = d_cmos_alias;
    a.adc = d_adc; a.expectation = d_expectation;
    a.in_signal = d_in_signal; a.in_noise = d_in_noise;
    a.out_signal = d_out_signal; a.out_noise = d_out_noise;
    cudaStream_t s = (cudaStream_t)stream;
    if (elem_type == SCB_F32) {
        if (det->type == SCB_DET_CMOS) launch_detector<float, SCB_DET_CMOS>(a, s);
        else if (det->type == SCB_DET_EMCCD) launch_detector<float, SCB_DET_EMCCD>(a, s);
        else launch_detector<float, SCB_DET_CCD>(a, s);
    } else {
        if (det->type == SCB_DET_CMOS) launch_detector<double, SCB_DET_CMOS>(a, s);
        else if (det->type == SCB_DET_EMCCD) launch_detector<double, SCB_DET_EMCCD>(a, s);
        else launch_detector<double, SCB_DET_CCD>(a, s);
    }
    SCB_CUDA_LAUNCH_CHECK("scb_detector_adc");
    return 0;
}
