// Shared device helpers: error plumbing, Philox4x32-10, uniform/normal transforms.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/scopyon_b200.h"

// Multiprocessors of the current device (B200: 148 = 2 dies x 74), queried once per device
// (psf.cu); grids of persistent / capped kernels are sized from it.
int scb_sm_count();
#define SCB_SM_COUNT scb_sm_count()

void scb_set_error(const char *fmt, ...);

// particles.cu: emission / bleaching over coordinate arrays (stride 1) or particle rows (stride 5)
int scb_emit_bleach_strided(uint64_t seed, int64_t n, const double *d_depth, const double *d_x, const double *d_y,
                            const double *d_p_state, const int32_t *d_mol_slot, const int64_t *d_mol_id,
                            const double *d_mol_id_column, int64_t stride, double unit_time, double focal_depth,
                            const scb_photophysics *phys, double *d_budget, double *d_weight,
                            double *d_true_data, void *stream);

#define SCB_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            scb_set_error(__VA_ARGS__);         \
            return (code);                      \
        }                                       \
    } while (0)

#define SCB_CUDA_LAUNCH_CHECK(what)                                          \
    do {                                                                     \
        cudaError_t e__ = cudaGetLastError();                                \
        if (e__ != cudaSuccess) {                                            \
            scb_set_error("%s: %s", what, cudaGetErrorString(e__));          \
            return (int)e__;                                                 \
        }                                                                    \
    } while (0)

#define SCB_CUDA(call)                                                       \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) {                                            \
            scb_set_error("%s: %s", #call, cudaGetErrorString(e__));         \
            return (int)e__;                                                 \
        }                                                                    \
    } while (0)

// Random-stream tags: the 4th counter word.  One tag per consumer keeps every
// (seed, entity, step) stream independent of how the work is sharded.
enum : uint32_t {
    SCB_TAG_DIFFUSE = 0x44494646u,  // 'DIFF'
    SCB_TAG_PLACE   = 0x504c4143u,  // 'PLAC'
    SCB_TAG_BUDGET  = 0x42554447u,  // 'BUDG'
    SCB_TAG_TRANS   = 0x5452414eu,  // 'TRAN'
    SCB_TAG_FPN     = 0x46504e20u,  // 'FPN '
    SCB_TAG_SHOT    = 0x53484f54u,  // 'SHOT' first draw of a pixel quad
    SCB_TAG_READ    = 0x52454144u,  // 'READ' readout-noise draw of a pixel quad
    SCB_TAG_EXTRA   = 0x58545241u,  // 'XTRA' per-pixel overflow stream (rejection loops)
};

struct Philox4 {
    uint32_t x, y, z, w;
};

// 32 x 32 -> (hi, lo): one IMAD.WIDE on the device (inline PTX keeps nvcc from splitting
// the 64-bit product into extra carry adds).
__host__ __device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1, %0}, p;\n\t}" : "=r"(hi), "=r"(lo) : "r"(a), "r"(b));
#else
    const uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}

// Philox4x32-10 (Salmon et al., SC'11).  counter = (c0,c1,c2,c3), key = (k0,k1).
struct PhiloxKeys {   // the ten round keys, hoisted out of per-element loops
    uint32_t a[10], b[10];
};
__host__ __device__ __forceinline__ PhiloxKeys philox_round_keys(uint32_t k0, uint32_t k1) {
    PhiloxKeys k;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        k.a[r] = k0 + (uint32_t)r * 0x9E3779B9u;
        k.b[r] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
    return k;
}
// ROUNDS = 10 is the generator every stream of this library uses.  7 is the smallest round count for which
// Random123 reports Philox4x32 Crush-resistant; it exists for the detector's measured variants only
// (detector.cu), as does any smaller count (timing ceilings, never a sampler).
template <int ROUNDS>
__host__ __device__ __forceinline__ Philox4 philox4x32_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         const PhiloxKeys &k) {
    static_assert(ROUNDS >= 1 && ROUNDS <= 10, "round keys are precomputed for ten rounds");
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(M0, c0, hi0, lo0);
        mulhilo32(M1, c2, hi1, lo1);
        c0 = hi1 ^ c1 ^ k.a[r];
        c2 = hi0 ^ c3 ^ k.b[r];
        c1 = lo1;
        c3 = lo0;
    }
    Philox4 out = {c0, c1, c2, c3};
    return out;
}
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          const PhiloxKeys &k) {
    return philox4x32_r<10>(c0, c1, c2, c3, k);
}
// round count chosen at run time (generic kernels, where the generator is not the bottleneck)
__host__ __device__ __forceinline__ Philox4 philox4x32_n(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1, int rounds) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    for (int r = 0; r < rounds; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(M0, c0, hi0, lo0);
        mulhilo32(M1, c2, hi1, lo1);
        c0 = hi1 ^ c1 ^ (k0 + (uint32_t)r * 0x9E3779B9u);
        c2 = hi0 ^ c3 ^ (k1 + (uint32_t)r * 0xBB67AE85u);
        c1 = lo1;
        c3 = lo0;
    }
    Philox4 out = {c0, c1, c2, c3};
    return out;
}
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox4x32_10(c0, c1, c2, c3, philox_round_keys(k0, k1));
}

__host__ __device__ __forceinline__ Philox4 philox_at(uint64_t seed, uint64_t entity, uint32_t step,
                                                      uint32_t tag) {
    return philox4x32_10((uint32_t)entity, (uint32_t)(entity >> 32), step, tag, (uint32_t)seed,
                         (uint32_t)(seed >> 32));
}

// (0,1] and [0,1) uniforms from 32 random bits; both avoid the value that would break
// the transform they feed (log(0), or an index equal to the table size).
__device__ __forceinline__ float u01_open_low(uint32_t r) {  // (0, 1]
    return ((float)(r >> 8) + 1.0f) * 5.9604644775390625e-08f;  // 2^-24
}
__device__ __forceinline__ float u01_half_open(uint32_t r) {  // [0, 1)
    return (float)(r >> 8) * 5.9604644775390625e-08f;
}
__host__ __device__ __forceinline__ double u01_open_low_53(uint32_t hi, uint32_t lo) {  // (0,1]
    uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;  // 53 bits
    return ((double)v + 1.0) * 1.1102230246251565e-16;  // 2^-53
}

// Box-Muller in fp32 on two 32-bit words: two independent N(0,1).  The radius uses all
// 32 bits of r0 (tail out to 6.66 sigma); sincospif keeps the angle exact at octants.
__device__ __forceinline__ void box_muller(uint32_t r0, uint32_t r1, float &n0, float &n1) {
    float u = ((float)r0 + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]; small values keep full precision
    float radius = sqrtf(-2.0f * logf(u));
    float s, c;
    sincospif((float)r1 * 4.656612873077393e-10f, &s, &c);  // angle = 2 pi r1 / 2^32
    n0 = radius * c;
    n1 = radius * s;
}

// SAT block layout shared by the builder (psf.cu) and the renderers.  With M = sat_modulus
// (the pixel pitch in table samples) the pixel edges of one footprint are M samples apart on
// both axes, i.e. they share one residue ("phase") mod M per axis.  Entry S[a][b] therefore
// lives in block (a % M, b % M) at slot (a / M, b / M):
//     index(pr, pc, ir, ic) = ((pr * M + pc) * B + ir) * B + ic,    value = S[min(ir*M+pr, side)][min(ic*M+pc, side)]
// where B slots per phase leave at least one slot past the last sample: those slots repeat
// the last row / column, so the clamped closing edge of a footprint is simply the next slot.
// All corners a footprint needs are then one dense (rows+1) x (cols+1) rectangle of one
// B x B block -- contiguous B*8-byte rows that a TMA bulk copy moves as they are.  B is a
// multiple of 16 (128-byte lines) unless that would waste more than a quarter of the table,
struct SatLayout {
    int modulus, slots, side;
    __host__ __device__ long long block_entries() const { return (long long)slots * slots; }
    __host__ __device__ long long table_entries() const { return (long long)modulus * modulus * slots * slots; }
};

static inline SatLayout scb_sat_layout(int n_radial, int modulus) {
    SatLayout L;
    L.modulus = modulus < 1 ? 1 : modulus;
    L.side = 2 * (n_radial - 1) + 1;
    const int used = (L.side + 1 + L.modulus - 1) / L.modulus;
    // ... else a multiple of 4: rows of fp32 box values stay multiples of 16 bytes, which every copy engine needs
    const int b16 = (used + 1 + 15) & ~15, b4 = (used + 1 + 3) & ~3;
    L.slots = (b16 * 4 <= (used + 1) * 5) ? b16 : b4;
    return L;
}

static inline unsigned int scb_grid_for(int64_t n, int block, int per_thread = 1) {
    int64_t work = (n + (int64_t)block * per_thread - 1) / ((int64_t)block * per_thread);
    if (work < 1) work = 1;
    return (unsigned int)work;
}
