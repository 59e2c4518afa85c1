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
//   Poisson(E)   E < 12: inversion by sequential search;  E >= 12: PTRS (Hoermann 1993), its first trial in
//                fp32 where the pixel is streamed (poisson_quick), the rest in the second pass
//   EMCCD        n ~ Poisson(E);  S = rint(g * Gamma(n, 1)), redrawn while outside the
//                reference's support [g*int(E-sigma)+, g*int(E+sigma)), sigma = 5 sqrt(E) + 10
//                (the reference pmf is exactly this Poisson->Gamma mixture evaluated at
//                integer S and truncated, _epifm.py:365-391)
//   readout      N(0, readout_noise) (EMCCD, CCD) or Walker-alias draw from the CMOS
//                read-noise table (CMOS)
#include "scb_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxAlias = 256;
constexpr int kFastCtasPerSm = 4;     // streaming kernel: 62 registers x 256 threads


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

// Poisson by inversion for lambda < 12, on one 32-bit random word.  The word (converted
// with round-toward-zero, so it stays below 2^32 and lambda = 0 can never count) is compared
// against the cumulative probabilities scaled by 2^32.  The first four terms are unrolled and
// branch free (poisson_head); a draw that passes all four continues in poisson_tail, which the
// streaming kernel enters once per pixel quad.  Explicit _rn intrinsics keep the compiler
// from contracting the recurrence differently in different kernels: every kernel draws the
// same count from the same word.
//
// The running sum is fp32: it carries ~2^-21 of relative error (ex2.approx, one rounding per term) and
// moves in steps of 256-512 near 2^32, so a sum scaled by exactly 2^32 can stall a few hundred BELOW the
// largest words -- every count test then passes and the search would run to its bound.  The scale is
// therefore 2^kPoissonScaleLog2 = 2^32 * (1 + 2.6e-6), the next fp32 exponent after 32: the converged sum
// lies ~11 000 above 2^32, beyond any error the recurrence can accumulate (< 6 500), so every word is
// passed while the terms are still far above the rounding step.  The price is the cumulative
// distribution scaled by 1 + 2.6e-6 -- the accuracy an fp32 inversion has anyway.
constexpr float kSmallLambda = 12.0f;
constexpr float kPoissonScaleLog2 = 32.000003814697266f;     // nextafterf(32, 64)

__device__ __forceinline__ float ex2_ftz(float x) {    // argument >= 14 here: no denormal handling needed
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct PoissonRun {   // state of the sequential search after the unrolled head
    float p, s;
};

__device__ __forceinline__ float poisson_head(float lambda, float rf, PoissonRun &run) {
    float p = ex2_ftz(fmaf(lambda, -1.4426950408889634f, kPoissonScaleLog2));            // 2^32 * exp(-lambda)
    float s = p;
    float k = rf >= s ? 1.0f : 0.0f;
    p = __fmul_rn(p, lambda);                              s = __fadd_rn(s, p); k += rf >= s ? 1.0f : 0.0f;
    p = __fmul_rn(p, __fmul_rn(lambda, 0.5f));             s = __fadd_rn(s, p); k += rf >= s ? 1.0f : 0.0f;
    p = __fmul_rn(p, __fmul_rn(lambda, 1.0f / 3.0f));      s = __fadd_rn(s, p); k += rf >= s ? 1.0f : 0.0f;
    run.p = p;
    run.s = s;
    return k;                                              // 0..4; 4 = not found yet
}

// Two pixels per instruction: Blackwell's packed fp32 arithmetic (add/mul/fma .f32x2 on a
// 64-bit register pair).  Round-to-nearest, no flush: bit-identical to the scalar _rn forms.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// poisson_head for two pixels (same operations, same results).
__device__ __forceinline__ void poisson_head2(float la, float lb, float rfa, float rfb, float &ka, float &kb,
                                              PoissonRun &ra, PoissonRun &rb) {
    const f32x2 lam = pack2(la, lb);
    f32x2 p = pack2(ex2_ftz(fmaf(la, -1.4426950408889634f, kPoissonScaleLog2)), ex2_ftz(fmaf(lb, -1.4426950408889634f, kPoissonScaleLog2)));
    f32x2 s = p;
    float sa, sb;
    unpack2(s, sa, sb);
    f32x2 k = pack2(rfa >= sa ? 1.0f : 0.0f, rfb >= sb ? 1.0f : 0.0f);
    p = mul2(p, lam);
    s = add2(s, p); unpack2(s, sa, sb);
    k = add2(k, pack2(rfa >= sa ? 1.0f : 0.0f, rfb >= sb ? 1.0f : 0.0f));
    p = mul2(p, mul2(lam, pack2(0.5f, 0.5f)));
    s = add2(s, p); unpack2(s, sa, sb);
    k = add2(k, pack2(rfa >= sa ? 1.0f : 0.0f, rfb >= sb ? 1.0f : 0.0f));
    p = mul2(p, mul2(lam, pack2(1.0f / 3.0f, 1.0f / 3.0f)));
    s = add2(s, p); unpack2(s, sa, sb);
    k = add2(k, pack2(rfa >= sa ? 1.0f : 0.0f, rfb >= sb ? 1.0f : 0.0f));
    unpack2(k, ka, kb);
    unpack2(p, ra.p, rb.p);
    ra.s = sa;
    rb.s = sb;
}

// With the scale above the sum passes every word long before its terms reach the rounding step; should it
// stop moving all the same (a lost term lies past the mode, later ones only shrink) the search ends at
// that count and never at the loop bound.
__device__ __noinline__ float poisson_tail(float lambda, float rf, float p, float s) {
    int k = 4;
#pragma unroll 1
    while (k < 160) {
        p = __fmul_rn(p, __fdividef(lambda, (float)k));
        const float next = __fadd_rn(s, p);
        if (rf < next || next == s) break;
        s = next;
        ++k;
    }
    return (float)k;
}

__device__ __forceinline__ float poisson_small(float lambda, uint32_t r) {
    const float rf = __uint2float_rz(r);                                    // [0, 2^32)
    PoissonRun run;
    float k = poisson_head(lambda, rf, run);
    if (k == 4.0f) k = poisson_tail(lambda, rf, run.p, run.s);
    return k;
}

// CMOS read noise: Walker alias draw on one 32-bit word.  r * n = (entry index, 32-bit
// fraction); the fraction is compared as an integer with the entry's threshold * 2^32.
__device__ __forceinline__ uint32_t alias_threshold_bits(float threshold) {
    return threshold >= 1.0f ? 0xffffffffu : __float2uint_rz(threshold * 4294967296.0f);
}
struct AliasSlot {    // shared-memory copy of scb_alias_entry with the threshold as 32-bit integer
    float value, alias_value;
    uint32_t threshold_bits, pad;
};
__device__ __forceinline__ float alias_draw(uint32_t r, uint32_t n_alias, const AliasSlot *slots) {
    uint32_t idx, frac;
    mulhilo32(r, n_alias, idx, frac);
    const AliasSlot e = slots[idx];
    return frac < e.threshold_bits ? e.value : e.alias_value;
}
__device__ __forceinline__ float alias_draw(uint32_t r, uint32_t n_alias, const scb_alias_entry *entries) {
    uint32_t idx, frac;
    mulhilo32(r, n_alias, idx, frac);
    const scb_alias_entry e = entries[idx];
    return frac < alias_threshold_bits(e.threshold) ? e.value : e.alias_value;
}
__device__ __forceinline__ void load_alias(AliasSlot *slots, const scb_alias_entry *entries, int n, int n_threads) {
    for (int i = threadIdx.x; i < n; i += n_threads) {
        const scb_alias_entry e = entries[i];
        AliasSlot s = {e.value, e.alias_value, alias_threshold_bits(e.threshold), 0u};
        slots[i] = s;
    }
    __syncthreads();
}

// PTRS, W. Hoermann, "The transformed rejection method for generating Poisson random
// variables" (1993); the same algorithm numpy uses for lam >= 10.
__device__ __noinline__ double poisson_ptrs(double lambda, PixelRng &rng) {
    const double slam = sqrt(lambda);
    double loglam = 0.0;                   // log(lambda), only needed when the quick acceptance fails
    bool have_log = false;
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
        if (!have_log) { loglam = log(lambda); have_log = true; }
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lambda + k * loglam - lgamma(k + 1.0)) return k;
    }
    return floor(lambda);
}

__device__ __forceinline__ double poisson_any(double lambda, uint32_t r, PixelRng &rng) {
    if (!(lambda > 0.0)) return 0.0;      // E <= 0 (or NaN): no signal, _epifm.py:395-396
    if (lambda < (double)kSmallLambda) return (double)poisson_small((float)lambda, r);
    return poisson_ptrs(lambda, rng);
}

// ---- bright pixels (12 <= E < kQuickLambda) without the second pass --------------------------------
// The FIRST trial of PTRS is made in fp32 on the pixel's shot word where it is drawn: U from the word's top 24 bits
// (as the inversion sampler uses them), the top 8 bits of V from its low 8.  The squeeze `us >= 0.07 && V <= vr`
// accepts 86-92 % of the trials, and it can be decided from V's top bits alone whenever the whole interval they leave
// lies below vr (and, for E < 256, from a conservative fp32 estimate of the exact acceptance bound when the squeeze
// does not hold): then the count is final and the pixel never reaches the worklist.  Otherwise the second pass
// CONTINUES that trial -- the same U and V, their remaining bits drawn from the pixel's overflow stream, the full
// acceptance test in fp64 -- and goes on with fresh trials if it fails, so the sampler as a whole is PTRS with
// nothing skipped or drawn twice.  Every operation is an explicit round-to-nearest intrinsic: the streaming kernel
// and the generic kernel must reach the same decision and the same count.
constexpr float kQuickLambda = 2048.0f;     // above it fp32 cannot place floor() reliably: second pass as before
constexpr float kQuickFullTest = 256.0f;    // below it the exact acceptance test is estimated in fp32 as well
__device__ __forceinline__ float __logf_rn(float x) { return logf(x); }   // (named like the pinned arithmetic around it)
struct PtrsShape {
    float b, a2, vr;        // b, 2 a, vr of Hoermann's algorithm for this lambda
};
__device__ __forceinline__ PtrsShape ptrs_shape(float lambda) {
    PtrsShape s;
    s.b = __fmaf_rn(2.53f, __fsqrt_rn(lambda), 0.931f);
    s.a2 = __fmul_rn(2.0f, __fmaf_rn(0.02483f, s.b, -0.059f));
    s.vr = __fsub_rn(0.9277f, __fdiv_rn(3.6224f, __fsub_rn(s.b, 2.0f)));
    return s;
}
__device__ __forceinline__ bool poisson_quick(float lambda, uint32_t r, float &k) {
    if (!(lambda >= kSmallLambda && lambda < kQuickLambda)) return false;
    const PtrsShape s = ptrs_shape(lambda);
    const float U = __fmaf_rn((float)(r >> 8) + 0.5f, 5.9604644775390625e-08f, -0.5f);     // centre of the 2^-24 cell
    const float us = __fsub_rn(0.5f, fabsf(U));
    const float v_top = (float)((r & 0xffu) + 1u) * 0.00390625f;                         // upper end of V's 2^-8 cell
    k = floorf(__fadd_rn(__fmaf_rn(__fadd_rn(__fdiv_rn(s.a2, us), s.b), U, lambda), 0.43f));
    if (us >= 0.07f && v_top <= s.vr) return true;
    // Not in the squeeze.  For moderate lambda the exact test V <= T(k, us) is decided here as well when V's whole
    // cell lies below an fp32 estimate of T taken 0.4 % low (the exponent below is good to ~1e-3 for lambda < 256);
    // a cell that straddles T, a trial that fails and everything brighter go to the second pass, which evaluates
    // the same trial in fp64.
    if (!(lambda < kQuickFullTest) || k < 0.0f || us < 0.013f) return false;
    const float a = __fmul_rn(0.5f, s.a2);
    const float invalpha = __fadd_rn(1.1239f, __fdiv_rn(1.1328f, __fsub_rn(s.b, 3.4f)));
    const float lhs = __fsub_rn(__logf_rn(invalpha), __logf_rn(__fadd_rn(__fdiv_rn(a, __fmul_rn(us, us)), s.b)));
    const float rhs = __fsub_rn(__fmaf_rn(k, __logf_rn(lambda), -lambda), lgammaf(__fadd_rn(k, 1.0f)));
    // accept iff log V + lhs <= rhs, i.e. V <= exp(rhs - lhs)
    const float t_low = __fmul_rn(expf(__fsub_rn(rhs, lhs)), 0.996f);
    return v_top <= t_low;
}
// Second pass for a pixel poisson_quick() declined: trial 1 on the same (U, V) refined by the overflow stream.
__device__ __noinline__ double poisson_after_quick(double lambda, uint32_t r, PixelRng &rng) {
    const double slam = sqrt(lambda);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    const double loglam = log(lambda);
    for (int trial = 0; trial < 1000; ++trial) {
        double U, V;
        if (trial == 0) {       // the cell of the first trial, position inside it from 32 + 32 fresh bits
            U = ((double)(r >> 8) + (double)u01_open_low(rng.next())) * 5.9604644775390625e-08 - 0.5;
            V = ((double)(r & 0xffu) + (double)u01_open_low(rng.next())) * 0.00390625;
        } else {
            U = u01_open_low_53(rng.next(), rng.next()) - 0.5;
            V = u01_open_low_53(rng.next(), rng.next());
        }
        const double us = 0.5 - fabs(U);
        const double k = floor((2.0 * a / us + b) * U + lambda + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lambda + k * loglam - lgamma(k + 1.0)) return k;
    }
    return floor(lambda);
}
// the count of a CCD / CMOS pixel in the second pass
__device__ __forceinline__ double poisson_second_pass(double lambda, uint32_t r, PixelRng &rng) {
    if (lambda >= (double)kSmallLambda && lambda < (double)kQuickLambda) return poisson_after_quick(lambda, r, rng);
    return poisson_any(lambda, r, rng);
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

struct DetArgs {
    PhiloxKeys keys;           // round keys of `seed`, filled on the host: constant-bank operands in the kernels
    uint64_t seed, frame;
    scb_detector det;
    int64_t n_pix;
    int32_t n_h;
    int n_alias;
    double adc_max, pow2bit;
    float inv_fullwell, inv_nh;
    // fp32 copies for the streaming kernel (no per-iteration double -> float conversions)
    float qe_f, qe_bg_f, fullwell_f, pow2bit_f, adc_max_f, readout_f, adc_offset_f;
    int rounds;                // Philox rounds of the per-quad shot / readout streams (10; see detector_rounds())
    const void *photons, *offset;
    const scb_alias_entry *alias;
    void *adc, *expectation;
    const void *in_signal, *in_noise;
    void *out_signal, *out_noise;
    uint32_t *slow_list;       // pixels that left the fast path (bright, or EMCCD with electrons)
    uint32_t *slow_count;
    // a block of frames in one launch (blockIdx.y = frame): images and worklists of consecutive
    // frames lie frame_bytes / slow_words apart; 0 for a single frame
    size_t frame_bytes, slow_words;
};

// Arguments of frame blockIdx.y of a multi-frame launch.
__device__ __forceinline__ DetArgs frame_args(const DetArgs &a) {
    DetArgs f = a;
    const size_t y = blockIdx.y;
    f.frame = a.frame + y;
    f.photons = (const char *)a.photons + y * a.frame_bytes;
    f.adc = (char *)a.adc + y * a.frame_bytes;
    f.slow_list = a.slow_list + y * a.slow_words;
    f.slow_count = a.slow_count + y * a.slow_words;
    return f;
}

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
    const float inv_gain = ((float)a.pow2bit - offset) * a.inv_fullwell;
    return fminf(fmaxf(fmaf(pe, inv_gain, offset), 0.0f), (float)a.adc_max);
}

// column (axis-1 index) of flat pixel p0 without a 64-bit division
__device__ __forceinline__ int column_of(int64_t p0, const DetArgs &a) {
    if (p0 >= ((int64_t)1 << 24)) return (int)(p0 % a.n_h);
    const int row = (int)__fmul_rz((float)(int)p0, a.inv_nh);
    int j = (int)p0 - row * a.n_h;
    if (j < 0) j += a.n_h;
    if (j >= a.n_h) j -= a.n_h;
    return j;
}

template <typename T>
struct PixelOut {
    T adc, ex, sig, noi;
};

// One pixel: expectation -> shot noise (+EM gain) -> readout noise -> ADC.
// Scalars only (no local arrays), so everything stays in registers.
template <typename T, int DET>
__device__ __forceinline__ PixelOut<T> detect_pixel(const DetArgs &a, const AliasSlot *s_alias, int64_t pix,
                                                    bool valid, T photons, T offset, uint32_t r_shot,
                                                    uint32_t r_read, float normal, T qe, T bg, float qe_bg) {
    PixelOut<T> o;
    o.ex = qe * (photons + bg);                                           // _epifm.py:1438-1441
    if (a.in_signal) {
        o.sig = valid ? ((const T *)a.in_signal)[pix] : (T)0;
    } else {
        // the Poisson mean as the streaming kernel forms it (one fma), so both draw the same count
        const float lam = sizeof(T) == 4 ? fmaf((float)qe, (float)photons, qe_bg) : (float)o.ex;
        // fast path: small expectation; an EMCCD pixel with zero photoelectrons stays zero
        // (S = 0 lies inside the reference's support whenever E < 12)
        const float n_small = poisson_small(fminf(lam, kSmallLambda), r_shot);
        o.sig = (T)n_small;
        float n_quick;
        if (DET != SCB_DET_EMCCD && poisson_quick(lam, r_shot, n_quick))
            o.sig = (T)n_quick;                                         // bright, first PTRS trial accepted in place
        else if (valid && (!(lam < kSmallLambda) || (DET == SCB_DET_EMCCD && n_small != 0.0f)))
            a.slow_list[atomicAdd(a.slow_count, 1u)] = (uint32_t)pix;   // finished by detector_slow_kernel
    }
    if (a.in_noise) {
        o.noi = valid ? ((const T *)a.in_noise)[pix] : (T)0;
    } else if (DET == SCB_DET_CMOS) {
        o.noi = (T)alias_draw(r_read, (uint32_t)a.n_alias, s_alias);
    } else {
        o.noi = (T)a.det.readout_noise * (T)normal;                       // _epifm.py:360-362
    }
    o.adc = adc_convert(o.sig + o.noi, offset, a);                        // _epifm.py:1464-1469
    return o;
}

template <typename T>
__device__ __forceinline__ void store_quad(void *base, int64_t p0, int64_t n_pix, bool full, T v0, T v1, T v2, T v3) {
    T *p = (T *)base + p0;
    if (full) {
        if constexpr (sizeof(T) == 4) {
            __stcs(reinterpret_cast<float4 *>(p), make_float4(v0, v1, v2, v3));
        } else {
            __stcs(reinterpret_cast<double2 *>(p), make_double2(v0, v1));
            __stcs(reinterpret_cast<double2 *>(p) + 1, make_double2(v2, v3));
        }
    } else {
        if (p0 + 0 < n_pix) p[0] = v0;
        if (p0 + 1 < n_pix) p[1] = v1;
        if (p0 + 2 < n_pix) p[2] = v2;
        if (p0 + 3 < n_pix) p[3] = v3;
    }
}

template <typename T>
__device__ __forceinline__ void load_quad(const void *base, int64_t p0, int64_t n_pix, bool full, bool streaming,
                                          T &v0, T &v1, T &v2, T &v3) {
    const T *p = (const T *)base + p0;
    if (full) {
        if constexpr (sizeof(T) == 4) {
            const float4 q = streaming ? __ldcs(reinterpret_cast<const float4 *>(p))
                                       : __ldg(reinterpret_cast<const float4 *>(p));
            v0 = q.x; v1 = q.y; v2 = q.z; v3 = q.w;
        } else {
            const double2 lo = streaming ? __ldcs(reinterpret_cast<const double2 *>(p))
                                         : __ldg(reinterpret_cast<const double2 *>(p));
            const double2 hi = streaming ? __ldcs(reinterpret_cast<const double2 *>(p) + 1)
                                         : __ldg(reinterpret_cast<const double2 *>(p) + 1);
            v0 = lo.x; v1 = lo.y; v2 = hi.x; v3 = hi.y;
        }
    } else {
        v0 = (p0 + 0 < n_pix) ? p[0] : (T)0;
        v1 = (p0 + 1 < n_pix) ? p[1] : (T)0;
        v2 = (p0 + 2 < n_pix) ? p[2] : (T)0;
        v3 = (p0 + 3 < n_pix) ? p[3] : (T)0;
    }
}

template <typename T, int DET>
__global__ void __launch_bounds__(kThreads, 4)
detector_kernel(const __grid_constant__ DetArgs a) {
    __shared__ AliasSlot s_alias[DET == SCB_DET_CMOS ? kMaxAlias : 1];
    if (DET == SCB_DET_CMOS) load_alias(s_alias, a.alias, a.n_alias, kThreads);
    const int64_t n_quads = (a.n_pix + 3) >> 2;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const uint32_t f_lo = (uint32_t)a.frame, f_hi = (uint32_t)(a.frame >> 32);
    const T qe = (T)a.det.qe;
    const T bg = a.det.background_on ? (T)a.det.background : (T)0;
    const float qe_bg = a.det.background_on ? (float)(a.det.qe * a.det.background) : 0.0f;
    const bool row_aligned = (a.n_h & 3) == 0;
    const bool gaussian_readout = DET != SCB_DET_CMOS && a.in_noise == nullptr && a.det.readout_noise > 0.0;

    // software pipeline: the next quad's photons are in flight while this one is processed
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    T nx0 = (T)0, nx1 = (T)0, nx2 = (T)0, nx3 = (T)0;
    if (q < n_quads) load_quad<T>(a.photons, q << 2, a.n_pix, (q << 2) + 4 <= a.n_pix, true, nx0, nx1, nx2, nx3);
    for (; q < n_quads; q += stride) {
        const int64_t p0 = q << 2;
        const bool full = p0 + 4 <= a.n_pix;
        const T ph0 = nx0, ph1 = nx1, ph2 = nx2, ph3 = nx3;
        if (q + stride < n_quads) {
            const int64_t pn = (q + stride) << 2;
            load_quad<T>(a.photons, pn, a.n_pix, pn + 4 <= a.n_pix, true, nx0, nx1, nx2, nx3);
        }
        // ADC offset: scalar, per column (axis-1 index), or per pixel  (_epifm.py:941-952)
        T of0, of1, of2, of3;
        if (a.det.fpn_type == SCB_FPN_NONE) {
            of0 = of1 = of2 = of3 = (T)a.det.adc_offset;
        } else if (a.det.fpn_type == SCB_FPN_PIXEL) {
            load_quad<T>(a.offset, p0, a.n_pix, full, true, of0, of1, of2, of3);
        } else {
            const int j0 = column_of(p0, a);
            if (row_aligned) {                 // the quad stays inside one image row
                load_quad<T>(a.offset, j0, a.n_h, true, false, of0, of1, of2, of3);
            } else {
                const T *o = (const T *)a.offset;
                of0 = o[j0];
                of1 = o[j0 + 1 < a.n_h ? j0 + 1 : j0 + 1 - a.n_h];
                of2 = o[j0 + 2 < a.n_h ? j0 + 2 : j0 + 2 - a.n_h];
                of3 = o[j0 + 3 < a.n_h ? j0 + 3 : j0 + 3 - a.n_h];
            }
        }
        Philox4 rs = {0u, 0u, 0u, 0u}, rr = {0u, 0u, 0u, 0u};
        if (a.in_signal == nullptr)
            rs = philox4x32_n((uint32_t)q, (uint32_t)(q >> 32), f_lo, SCB_TAG_SHOT ^ f_hi, k0, k1, a.rounds);
        if (a.in_noise == nullptr && (DET == SCB_DET_CMOS || gaussian_readout))
            rr = philox4x32_n((uint32_t)q, (uint32_t)(q >> 32), f_lo, SCB_TAG_READ ^ f_hi, k0, k1, a.rounds);
        float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
        if (gaussian_readout) {
            box_muller(rr.x, rr.y, n0, n1);
            box_muller(rr.z, rr.w, n2, n3);
        }
        const PixelOut<T> o0 = detect_pixel<T, DET>(a, s_alias, p0 + 0, p0 + 0 < a.n_pix, ph0, of0, rs.x, rr.x, n0, qe, bg, qe_bg);
        const PixelOut<T> o1 = detect_pixel<T, DET>(a, s_alias, p0 + 1, p0 + 1 < a.n_pix, ph1, of1, rs.y, rr.y, n1, qe, bg, qe_bg);
        const PixelOut<T> o2 = detect_pixel<T, DET>(a, s_alias, p0 + 2, p0 + 2 < a.n_pix, ph2, of2, rs.z, rr.z, n2, qe, bg, qe_bg);
        const PixelOut<T> o3 = detect_pixel<T, DET>(a, s_alias, p0 + 3, p0 + 3 < a.n_pix, ph3, of3, rs.w, rr.w, n3, qe, bg, qe_bg);

        store_quad<T>(a.adc, p0, a.n_pix, full, o0.adc, o1.adc, o2.adc, o3.adc);
        if (a.expectation) store_quad<T>(a.expectation, p0, a.n_pix, full, o0.ex, o1.ex, o2.ex, o3.ex);
        if (a.out_signal) store_quad<T>(a.out_signal, p0, a.n_pix, full, o0.sig, o1.sig, o2.sig, o3.sig);
        if (a.out_noise) store_quad<T>(a.out_noise, p0, a.n_pix, full, o0.noi, o1.noi, o2.noi, o3.noi);
    }
}

// Streaming kernel specialised for the production configuration: fp32 frames, no injected
// draws, no taps, n_pix a multiple of 4 and < 2^31.  Same arithmetic and the same random
// streams as detector_kernel<float, DET> (tests/test_gpu_detector.py compares them bit for
// bit), but 32-bit indexing, no per-pixel validity or tap branches, the column index
// advanced incrementally, one slow-list test and one Poisson-tail test per pixel quad, the
// Philox round keys read straight from the constant bank, and the Poisson recurrence and
// the ADC on packed fp32 pairs.  What bounds it (ncu, profiles/): the 40 IMAD.WIDE of the
// two Philox4x32-10 blocks issue at a quarter rate on the heavy FMA pipe
// (tools/probes/imad_probe.cu), 160 of the ~250 issue cycles a warp spends per quad.
template <int DET, int FPN, int ROUNDS>
__global__ void __launch_bounds__(kThreads, kFastCtasPerSm)
detector_fast_kernel(const __grid_constant__ DetArgs launch) {
    const DetArgs a = frame_args(launch);
    __shared__ AliasSlot s_alias[DET == SCB_DET_CMOS ? kMaxAlias : 1];
    if (DET == SCB_DET_CMOS) load_alias(s_alias, a.alias, a.n_alias, kThreads);
    const uint32_t n_quads = (uint32_t)(a.n_pix >> 2);
    const uint32_t f_lo = (uint32_t)a.frame, t_shot = SCB_TAG_SHOT ^ (uint32_t)(a.frame >> 32),
                   t_read = SCB_TAG_READ ^ (uint32_t)(a.frame >> 32);
    const float qe = launch.qe_f, qe_bg = launch.qe_bg_f;
    const float fullwell = launch.fullwell_f, pow2bit = launch.pow2bit_f, adc_max = launch.adc_max_f;
    const float inv_fullwell = launch.inv_fullwell;
    const float rn = launch.readout_f;
    const float4 *photons = reinterpret_cast<const float4 *>(a.photons);
    float4 *adc = reinterpret_cast<float4 *>(a.adc);
    const float *offset = (const float *)a.offset;
    const uint32_t n_alias = (uint32_t)a.n_alias;
    const uint32_t n_h = (uint32_t)a.n_h;

    const uint32_t stride = gridDim.x * kThreads;
    uint32_t q = blockIdx.x * kThreads + threadIdx.x;
    // column of the quad's first pixel, advanced by (4*stride) mod n_h per iteration
    uint32_t j0 = 0, j_step = 0;
    if (FPN == SCB_FPN_COLUMN) {
        j0 = (uint32_t)(((uint64_t)q << 2) % n_h);
        j_step = (uint32_t)(((uint64_t)stride << 2) % n_h);
    }
    // The loop is warp uniform (a warp leaves it together: lanes past the end carry zero photons and store
    // nothing), so the slow-pixel list can be appended to with warp-wide votes.
    const uint32_t lane = threadIdx.x & 31u;
    float4 next = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < n_quads) next = __ldcs(photons + q);
    for (; q - lane < n_quads; q += stride) {
        const bool valid = q < n_quads;
        const float4 ph = next;
        next = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q + stride < n_quads) next = __ldcs(photons + q + stride);
        float4 off;
        if (FPN == SCB_FPN_NONE) {
            off.x = off.y = off.z = off.w = launch.adc_offset_f;
        } else if (FPN == SCB_FPN_PIXEL) {
            off = valid ? __ldcs(reinterpret_cast<const float4 *>(offset) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            off = __ldg(reinterpret_cast<const float4 *>(offset + j0));
            j0 += j_step;
            if (j0 >= n_h) j0 -= n_h;
        }
        const Philox4 rs = philox4x32_r<ROUNDS>(q, 0u, f_lo, t_shot, launch.keys);
        Philox4 rr = {0u, 0u, 0u, 0u};
        if (DET == SCB_DET_CMOS || rn > 0.0f) rr = philox4x32_r<ROUNDS>(q, 0u, f_lo, t_read, launch.keys);
        // shot noise: two packed branch-free heads, one test for the rare tails
        const float l0 = fmaf(qe, ph.x, qe_bg), l1 = fmaf(qe, ph.y, qe_bg),
                    l2 = fmaf(qe, ph.z, qe_bg), l3 = fmaf(qe, ph.w, qe_bg);
        const float c0 = fminf(l0, kSmallLambda), c1 = fminf(l1, kSmallLambda),
                    c2 = fminf(l2, kSmallLambda), c3 = fminf(l3, kSmallLambda);
        const float f0 = __uint2float_rz(rs.x), f1 = __uint2float_rz(rs.y),
                    f2 = __uint2float_rz(rs.z), f3 = __uint2float_rz(rs.w);
        PoissonRun r0, r1, r2, r3;
        float k0, k1, k2, k3;
        poisson_head2(c0, c1, f0, f1, k0, k1, r0, r1);
        poisson_head2(c2, c3, f2, f3, k2, k3, r2, r3);
        if (fmaxf(fmaxf(k0, k1), fmaxf(k2, k3)) == 4.0f) {
            if (k0 == 4.0f) k0 = poisson_tail(c0, f0, r0.p, r0.s);
            if (k1 == 4.0f) k1 = poisson_tail(c1, f1, r1.p, r1.s);
            if (k2 == 4.0f) k2 = poisson_tail(c2, f2, r2.p, r2.s);
            if (k3 == 4.0f) k3 = poisson_tail(c3, f3, r3.p, r3.s);
        }
        // pixels for the general samplers (bright, NaN, or EMCCD with electrons): one vote per quad; a warp
        // that holds any claims its list space with a single atomic (a frame of bright pixels would otherwise
        // serialise half a million atomics on one counter)
        bool slow = valid && !(l0 < kSmallLambda && l1 < kSmallLambda && l2 < kSmallLambda && l3 < kSmallLambda);
        if (DET == SCB_DET_EMCCD) slow = slow || (valid && (k0 + k1) + (k2 + k3) != 0.0f);
        if (__any_sync(0xffffffffu, slow)) {
            bool s0 = slow && (!(l0 < kSmallLambda) || (DET == SCB_DET_EMCCD && k0 != 0.0f));
            bool s1 = slow && (!(l1 < kSmallLambda) || (DET == SCB_DET_EMCCD && k1 != 0.0f));
            bool s2 = slow && (!(l2 < kSmallLambda) || (DET == SCB_DET_EMCCD && k2 != 0.0f));
            bool s3 = slow && (!(l3 < kSmallLambda) || (DET == SCB_DET_EMCCD && k3 != 0.0f));
            if (DET != SCB_DET_EMCCD) {
                // bright pixels: the first PTRS trial here; only the ones it leaves undecided go to the second pass
                float kq;
                if (s0 && poisson_quick(l0, rs.x, kq)) { k0 = kq; s0 = false; }
                if (s1 && poisson_quick(l1, rs.y, kq)) { k1 = kq; s1 = false; }
                if (s2 && poisson_quick(l2, rs.z, kq)) { k2 = kq; s2 = false; }
                if (s3 && poisson_quick(l3, rs.w, kq)) { k3 = kq; s3 = false; }
            }
            const uint32_t b0 = __ballot_sync(0xffffffffu, s0), b1 = __ballot_sync(0xffffffffu, s1),
                           b2 = __ballot_sync(0xffffffffu, s2), b3 = __ballot_sync(0xffffffffu, s3);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(a.slow_count, (uint32_t)(__popc(b0) + __popc(b1) + __popc(b2) + __popc(b3)));
            base = __shfl_sync(0xffffffffu, base, 0);
            const uint32_t below = (1u << lane) - 1u;
            uint32_t pos = base + __popc(b0 & below) + __popc(b1 & below) + __popc(b2 & below) + __popc(b3 & below);
            if (s0) a.slow_list[pos++] = (q << 2) + 0;
            if (s1) a.slow_list[pos++] = (q << 2) + 1;
            if (s2) a.slow_list[pos++] = (q << 2) + 2;
            if (s3) a.slow_list[pos++] = (q << 2) + 3;
        }
        float n0, n1, n2, n3;
        if (DET == SCB_DET_CMOS) {
            n0 = alias_draw(rr.x, n_alias, s_alias);
            n1 = alias_draw(rr.y, n_alias, s_alias);
            n2 = alias_draw(rr.z, n_alias, s_alias);
            n3 = alias_draw(rr.w, n_alias, s_alias);
        } else if (rn > 0.0f) {
            box_muller(rr.x, rr.y, n0, n1);
            box_muller(rr.z, rr.w, n2, n3);
            n0 *= rn; n1 *= rn; n2 *= rn; n3 *= rn;
        } else {
            n0 = n1 = n2 = n3 = rn * 0.0f;
        }
        // full-well clip, gain, offset, clip to the ADC range: two pixels per instruction
        auto convert2 = [&](float sa, float sb, float na, float nb, float oa, float ob, float &ua, float &ub) {
            float pa, pb;
            unpack2(add2(pack2(sa, sb), pack2(na, nb)), pa, pb);
            const f32x2 off2 = pack2(oa, ob);
            const f32x2 inv_gain = mul2(sub2(pack2(pow2bit, pow2bit), off2), pack2(inv_fullwell, inv_fullwell));
            unpack2(fma2(pack2(fminf(pa, fullwell), fminf(pb, fullwell)), inv_gain, off2), ua, ub);
            ua = fminf(fmaxf(ua, 0.0f), adc_max);
            ub = fminf(fmaxf(ub, 0.0f), adc_max);
        };
        float4 out;
        convert2(k0, k1, n0, n1, off.x, off.y, out.x, out.y);
        convert2(k2, k3, n2, n3, off.z, off.w, out.z, out.w);
        if (valid) __stcs(adc + q, out);
    }
}

// Second pass over the (usually short) list of pixels that need the general samplers:
// PTRS Poisson for E >= 12, Poisson -> Gamma multiplication register for EMCCD.  Draws are
// keyed by pixel and frame exactly as in the streaming kernel, so the readout noise added
// here is the one the streaming kernel drew.
template <typename T, int DET>
__global__ void __launch_bounds__(kThreads)
detector_slow_kernel(const __grid_constant__ DetArgs launch) {
    const DetArgs a = frame_args(launch);
    const uint32_t n = *a.slow_count;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const uint32_t f_lo = (uint32_t)a.frame, f_hi = (uint32_t)(a.frame >> 32);
    for (uint32_t t = blockIdx.x * kThreads + threadIdx.x; t < n; t += gridDim.x * kThreads) {
        const int64_t pix = a.slow_list[t];
        const int64_t q = pix >> 2;
        const int w = (int)(pix & 3);
        const T qe = (T)a.det.qe;
        const T bg = a.det.background_on ? (T)a.det.background : (T)0;
        const T photon = ((const T *)a.photons)[pix];
        const T ex = sizeof(T) == 4
            ? (T)fmaf((float)qe, (float)photon, a.det.background_on ? (float)(a.det.qe * a.det.background) : 0.0f)
            : qe * (photon + bg);
        T offset = (T)a.det.adc_offset;
        if (a.det.fpn_type == SCB_FPN_PIXEL) offset = ((const T *)a.offset)[pix];
        else if (a.det.fpn_type == SCB_FPN_COLUMN) offset = ((const T *)a.offset)[pix % a.n_h];
        const Philox4 rs = philox4x32_n((uint32_t)q, (uint32_t)(q >> 32), f_lo, SCB_TAG_SHOT ^ f_hi, k0, k1, a.rounds);
        const uint32_t r_shot = w == 0 ? rs.x : w == 1 ? rs.y : w == 2 ? rs.z : rs.w;
        PixelRng rng(a.seed, (uint64_t)pix, a.frame);
        double sig;
        if (DET == SCB_DET_EMCCD) sig = emccd_signal((double)ex, a.det.emgain, r_shot, rng);
        else sig = poisson_second_pass((double)ex, r_shot, rng);
        T noi = (T)0;
        if (a.in_noise) {
            noi = ((const T *)a.in_noise)[pix];
        } else {
            const Philox4 rr = philox4x32_n((uint32_t)q, (uint32_t)(q >> 32), f_lo, SCB_TAG_READ ^ f_hi, k0, k1, a.rounds);
            if (DET == SCB_DET_CMOS) {
                const uint32_t r_read = w == 0 ? rr.x : w == 1 ? rr.y : w == 2 ? rr.z : rr.w;
                noi = (T)alias_draw(r_read, (uint32_t)a.n_alias, a.alias);
            } else if (a.det.readout_noise > 0.0) {
                float n0, n1;
                if (w < 2) box_muller(rr.x, rr.y, n0, n1);
                else box_muller(rr.z, rr.w, n0, n1);
                noi = (T)a.det.readout_noise * (T)((w & 1) ? n1 : n0);
            }
        }
        ((T *)a.adc)[pix] = adc_convert((T)sig + noi, offset, a);
        if (a.out_signal) ((T *)a.out_signal)[pix] = (T)sig;
    }
}

// test hook: the inversion sampler on given (lambda, word) pairs
__global__ void __launch_bounds__(kThreads)
poisson_inversion_probe_kernel(int64_t n, const float *__restrict__ lambda, const uint32_t *__restrict__ word,
                               float *__restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i < n) count[i] = poisson_small(fminf(lambda[i], kSmallLambda), word[i]);
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

bool fast_path_ok(const DetArgs &a, size_t elem_bytes) {
    return elem_bytes == 4 && !a.in_signal && !a.in_noise && !a.out_signal && !a.out_noise &&
           !a.expectation && (a.n_pix & 3) == 0 && a.n_pix < ((int64_t)1 << 31) &&
           (a.det.fpn_type != SCB_FPN_COLUMN || (a.n_h & 3) == 0);
}

// Philox rounds of the detector's per-quad streams.  10 (the Random123 / cuRAND default) unless the
// environment asks for a measured variant: SCB_DETECTOR_ROUNDS=7 is the smallest Crush-resistant count
// (all kernels follow, so the statistical tests can be run under it); 1..6 only time the streaming
// kernel with a cheaper generator -- the ceiling a free generator would give -- and are not samplers.
int detector_rounds() {
    static const int rounds = [] {
        const char *env = getenv("SCB_DETECTOR_ROUNDS");
        const int r = env ? atoi(env) : 10;
        return r >= 1 && r <= 10 ? r : 10;
    }();
    return rounds;
}

template <int DET, int ROUNDS>
void launch_fast_rounds(const DetArgs &a, const dim3 grid, cudaStream_t s) {
    if (a.det.fpn_type == SCB_FPN_NONE) detector_fast_kernel<DET, SCB_FPN_NONE, ROUNDS><<<grid, kThreads, 0, s>>>(a);
    else if (a.det.fpn_type == SCB_FPN_PIXEL) detector_fast_kernel<DET, SCB_FPN_PIXEL, ROUNDS><<<grid, kThreads, 0, s>>>(a);
    else detector_fast_kernel<DET, SCB_FPN_COLUMN, ROUNDS><<<grid, kThreads, 0, s>>>(a);
}

template <int DET>
void launch_fast(const DetArgs &a, int n_frames, cudaStream_t s) {
    // two waves of the resident CTAs per frame; blockIdx.y = frame of a block
    const int64_t n_quads = a.n_pix >> 2;
    int64_t blocks = (n_quads + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)SCB_SM_COUNT * kFastCtasPerSm * 2;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const dim3 grid((unsigned)blocks, (unsigned)n_frames);
    if (a.rounds == 10) launch_fast_rounds<DET, 10>(a, grid, s);
    else if (a.rounds == 7) launch_fast_rounds<DET, 7>(a, grid, s);
    else launch_fast_rounds<DET, 1>(a, grid, s);          // timing ceiling only
}

template <typename T, int DET>
void launch_detector(const DetArgs &a, int n_frames, cudaStream_t s) {
    const int64_t n_quads = (a.n_pix + 3) >> 2;
    int64_t blocks = (n_quads + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)SCB_SM_COUNT * 4 * 2;   // two waves of 4 resident CTAs (256 threads) per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (n_frames > 1)      // one counter per frame, a.slow_words apart
        cudaMemset2DAsync(a.slow_count, a.slow_words * sizeof(uint32_t), 0, sizeof(uint32_t), (size_t)n_frames, s);
    else
        cudaMemsetAsync(a.slow_count, 0, sizeof(uint32_t), s);
    const dim3 grid((unsigned)blocks, (unsigned)n_frames);
    if (fast_path_ok(a, sizeof(T))) {
        launch_fast<DET>(a, n_frames, s);
    } else {
        detector_kernel<T, DET><<<(unsigned)blocks, kThreads, 0, s>>>(a);     // single frame only
    }
    if (a.in_signal == nullptr)
        detector_slow_kernel<T, DET><<<dim3(SCB_SM_COUNT * 2, (unsigned)n_frames), kThreads, 0, s>>>(a);
}

}  // namespace

extern "C" int scb_test_poisson_inversion(int64_t n, const float *d_lambda, const uint32_t *d_word, float *d_count,
                                          void *stream) {
    SCB_REQUIRE(n >= 0 && (n == 0 || (d_lambda && d_word && d_count)), SCB_E_NULL, "scb_test_poisson_inversion: NULL pointer");
    if (n == 0) return 0;
    poisson_inversion_probe_kernel<<<scb_grid_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(n, d_lambda, d_word,
                                                                                                  d_count);
    SCB_CUDA_LAUNCH_CHECK("scb_test_poisson_inversion");
    return 0;
}

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

extern "C" size_t scb_detector_workspace_bytes(int32_t n_w, int32_t n_h) {
    if (n_w <= 0 || n_h <= 0) return 0;
    return 256 + (size_t)n_w * n_h * sizeof(uint32_t);   // counter + worst-case pixel list
}

static int detector_adc(uint64_t seed, uint64_t frame, int n_frames, const scb_detector *det, int32_t n_w, int32_t n_h,
                        int elem_type, const void *d_photons, const void *d_offset,
                        const scb_alias_entry *d_cmos_alias, int n_alias, void *d_adc,
                        void *d_expectation, const void *d_in_signal, const void *d_in_noise,
                        void *d_out_signal, void *d_out_noise, void *d_workspace,
                        size_t workspace_bytes, void *stream) {
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
    SCB_REQUIRE((int64_t)n_w * n_h < ((int64_t)1 << 32), SCB_E_INVALID, "scb_detector_adc: image too large");
    const size_t frame_work = scb_detector_workspace_bytes(n_w, n_h);
    SCB_REQUIRE(n_frames >= 1 && n_frames <= 65535, SCB_E_INVALID, "scb_detector_adc: n_frames=%d", n_frames);
    SCB_REQUIRE(d_workspace && workspace_bytes >= frame_work * (size_t)n_frames, SCB_E_WORKSPACE,
                "scb_detector_adc: workspace %zu < %zu", workspace_bytes, frame_work * (size_t)n_frames);
    DetArgs a;
    a.frame_bytes = n_frames > 1 ? (size_t)n_w * n_h * (elem_type == SCB_F32 ? 4 : 8) : 0;
    a.slow_words = n_frames > 1 ? frame_work / sizeof(uint32_t) : 0;
    a.slow_count = (uint32_t *)d_workspace;
    a.slow_list = (uint32_t *)((char *)d_workspace + 256);
    a.seed = seed; a.frame = frame; a.det = *det;
    a.keys = philox_round_keys((uint32_t)seed, (uint32_t)(seed >> 32));
    a.rounds = detector_rounds();
    a.n_pix = (int64_t)n_w * n_h; a.n_h = n_h;
    a.n_alias = need_alias ? n_alias : 0;
    a.pow2bit = ldexp(1.0, det->bit);
    a.inv_fullwell = (float)(1.0 / det->fullwell);
    a.inv_nh = (float)(1.0 / (double)n_h);
    a.qe_f = (float)det->qe;
    a.qe_bg_f = det->background_on ? (float)(det->qe * det->background) : 0.0f;
    a.fullwell_f = (float)det->fullwell;
    a.pow2bit_f = (float)a.pow2bit;
    a.readout_f = (float)det->readout_noise;
    a.adc_offset_f = (float)det->adc_offset;
    a.adc_max = a.pow2bit - 1.0;
    a.adc_max_f = (float)a.adc_max;
    a.photons = d_photons; a.offset = d_offset; a.alias = d_cmos_alias;
    a.adc = d_adc; a.expectation = d_expectation;
    a.in_signal = d_in_signal; a.in_noise = d_in_noise;
    a.out_signal = d_out_signal; a.out_noise = d_out_noise;
    cudaStream_t s = (cudaStream_t)stream;
    SCB_REQUIRE(n_frames == 1 || fast_path_ok(a, elem_type == SCB_F32 ? 4 : 8), SCB_E_UNSUPPORTED,
                "scb_detector_adc_frames: a block of frames needs fp32 images of a multiple of 4 pixels and no taps");
    if (elem_type == SCB_F32) {
        if (det->type == SCB_DET_CMOS) launch_detector<float, SCB_DET_CMOS>(a, n_frames, s);
        else if (det->type == SCB_DET_EMCCD) launch_detector<float, SCB_DET_EMCCD>(a, n_frames, s);
        else launch_detector<float, SCB_DET_CCD>(a, n_frames, s);
    } else {
        if (det->type == SCB_DET_CMOS) launch_detector<double, SCB_DET_CMOS>(a, n_frames, s);
        else if (det->type == SCB_DET_EMCCD) launch_detector<double, SCB_DET_EMCCD>(a, n_frames, s);
        else launch_detector<double, SCB_DET_CCD>(a, n_frames, s);
    }
    SCB_CUDA_LAUNCH_CHECK("scb_detector_adc");
    return 0;
}

extern "C" int scb_detector_adc(uint64_t seed, uint64_t frame, const scb_detector *det, int32_t n_w, int32_t n_h,
                                int elem_type, const void *d_photons, const void *d_offset,
                                const scb_alias_entry *d_cmos_alias, int n_alias, void *d_adc,
                                void *d_expectation, const void *d_in_signal, const void *d_in_noise,
                                void *d_out_signal, void *d_out_noise, void *d_workspace,
                                size_t workspace_bytes, void *stream) {
    return detector_adc(seed, frame, 1, det, n_w, n_h, elem_type, d_photons, d_offset, d_cmos_alias, n_alias, d_adc,
                        d_expectation, d_in_signal, d_in_noise, d_out_signal, d_out_noise, d_workspace,
                        workspace_bytes, stream);
}

// Frames first_frame .. first_frame + n_frames - 1 of a movie in one pair of launches: images
// [n_frames][n_w][n_h] in and out (fp32, no taps), scb_detector_workspace_bytes() of scratch per frame.
extern "C" int scb_detector_adc_frames(uint64_t seed, uint64_t first_frame, int n_frames, const scb_detector *det,
                                       int32_t n_w, int32_t n_h, int elem_type, const void *d_photons,
                                       const void *d_offset, const scb_alias_entry *d_cmos_alias, int n_alias,
                                       void *d_adc, void *d_workspace, size_t workspace_bytes, void *stream) {
    return detector_adc(seed, first_frame, n_frames, det, n_w, n_h, elem_type, d_photons, d_offset, d_cmos_alias,
                        n_alias, d_adc, nullptr, nullptr, nullptr, nullptr, nullptr, d_workspace, workspace_bytes,
                        stream);
}
