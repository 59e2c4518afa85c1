// Frame-stack post-processing on the device: 8-bit scaling of finished frames.
//
// Reference: Image.__as_8bit (/root/reference/src/scopyon/image.py:98-123, "same as
// scipy.misc.bytescale") and Video.save (:240-277), which scales every frame of a movie with one
// common (cmin, cmax) = extrema over all frames.  A movie leaves the GPU as uint8 -- a quarter of the
// fp32 bytes -- when only the 8-bit pictures are wanted.  The arithmetic is the reference's, in
// fp64: ((data - cmin) * scale + low).clip(low, high) + 0.5 -> uint8 (truncation).
#include "scb_common.cuh"

namespace {

// doubles <-> unsigned keys with the same order (for atomicMin / atomicMax)
__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

template <typename T>
__global__ void __launch_bounds__(256)
minmax_kernel(const T *__restrict__ data, int64_t n, unsigned long long *__restrict__ keys) {
    __shared__ double s_lo[8], s_hi[8];
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)data[i];
        lo = fmin(lo, v);        // NaNs are skipped, like a frame without them; numpy would propagate them
        hi = fmax(hi, v);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
        atomicMin(&keys[0], order_key(lo));
        atomicMax(&keys[1], order_key(hi));
    }
}

__global__ void minmax_init_kernel(unsigned long long *keys) {
    keys[0] = 0xffffffffffffffffull;
    keys[1] = 0ull;
}

__global__ void minmax_finish_kernel(const unsigned long long *keys, double *out) {
    out[0] = key_value(keys[0]);
    out[1] = key_value(keys[1]);
}

// limits[0..1] = (cmin, cmax) on the device (e.g. from scb_frames_minmax), or the host values.
template <typename T>
__global__ void __launch_bounds__(256)
to_8bit_kernel(const T *__restrict__ data, int64_t n, const double *__restrict__ limits, double cmin_host,
               double cmax_host, double low, double high, uint8_t *__restrict__ out) {
    const double cmin = limits ? limits[0] : cmin_host, cmax = limits ? limits[1] : cmax_host;
    const double cscale = __dsub_rn(cmax, cmin);
    const bool flat = cscale == 0.0;                                    // image.py:117-118
    const double scale = flat ? 0.0 : __ddiv_rn(__dsub_rn(high, low), cscale);
    // four pixels per thread: one 32-bit store
    const int64_t quads = (n + 3) >> 2;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = (q << 2) + k;
            uint32_t byte = 0;
            if (i < n) {
                if (flat) {
                    byte = (uint32_t)(uint8_t)low;                      // numpy.ones(uint8) * low
                } else {
                    double v = __dadd_rn(__dmul_rn(__dsub_rn((double)data[i], cmin), scale), low);   // image.py:120
                    v = fmin(fmax(v, low), high);                        // .clip(low, high)
                    byte = (uint32_t)(uint8_t)(long long)__dadd_rn(v, 0.5);   // (+ 0.5).astype(uint8): truncation
                }
            }
            packed |= byte << (8 * k);
        }
        if ((q << 2) + 3 < n) {
            reinterpret_cast<uint32_t *>(out)[q] = packed;
        } else {
            for (int k = 0; k < 4 && (q << 2) + k < n; ++k) out[(q << 2) + k] = (uint8_t)(packed >> (8 * k));
        }
    }
}

// 16-bit camera counts: uint16(rint(clip(v, 0, 65535))), NaN -> 0.  Eight pixels per thread: one 16-byte store.
template <typename T>
__global__ void __launch_bounds__(256)
to_u16_kernel(const T *__restrict__ data, int64_t n, uint16_t *__restrict__ out) {
    const int64_t octs = (n + 7) >> 3;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < octs; q += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int64_t i = (q << 3) + k;
            uint32_t v = 0;
            if (i < n) {
                const double x = (double)data[i];
                v = (uint32_t)__double2int_rn(fmin(fmax(x, 0.0), 65535.0));     // fmax(NaN, 0) = 0
            }
            w[k >> 1] |= v << (16 * (k & 1));
        }
        if ((q << 3) + 7 < n) {
            reinterpret_cast<uint4 *>(out)[q] = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int k = 0; k < 8 && (q << 3) + k < n; ++k) out[(q << 3) + k] = (uint16_t)(w[k >> 1] >> (16 * (k & 1)));
        }
    }
}

}  // namespace

extern "C" int scb_frames_to_u16(const void *d_frames, int64_t n, int elem_type, uint16_t *d_out, void *stream) {
    SCB_REQUIRE(d_frames && d_out, SCB_E_NULL, "scb_frames_to_u16: NULL pointer");
    SCB_REQUIRE(n > 0, SCB_E_INVALID, "scb_frames_to_u16: n=%lld", (long long)n);
    SCB_REQUIRE(elem_type == SCB_F32 || elem_type == SCB_F64, SCB_E_INVALID, "elem_type=%d", elem_type);
    SCB_REQUIRE(((uintptr_t)d_out & 15) == 0, SCB_E_INVALID, "scb_frames_to_u16: output must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned int want = scb_grid_for((n + 7) >> 3, 256, 2);
    const unsigned int cap = (unsigned int)SCB_SM_COUNT * 16;
    const unsigned int grid = want < cap ? want : cap;
    if (elem_type == SCB_F32) to_u16_kernel<float><<<grid, 256, 0, s>>>((const float *)d_frames, n, d_out);
    else to_u16_kernel<double><<<grid, 256, 0, s>>>((const double *)d_frames, n, d_out);
    SCB_CUDA_LAUNCH_CHECK("scb_frames_to_u16");
    return 0;
}

extern "C" int scb_frames_minmax(const void *d_frames, int64_t n, int elem_type, double *d_minmax, void *d_workspace,
                                 void *stream) {
    SCB_REQUIRE(d_frames && d_minmax && d_workspace, SCB_E_NULL, "scb_frames_minmax: NULL pointer");
    SCB_REQUIRE(n > 0, SCB_E_INVALID, "scb_frames_minmax: n=%lld", (long long)n);
    SCB_REQUIRE(elem_type == SCB_F32 || elem_type == SCB_F64, SCB_E_INVALID, "elem_type=%d", elem_type);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *keys = (unsigned long long *)d_workspace;       // 16 bytes
    minmax_init_kernel<<<1, 1, 0, s>>>(keys);
    const unsigned int grid = scb_grid_for(n, 256, 16) < (unsigned int)SCB_SM_COUNT * 8 ? scb_grid_for(n, 256, 16) : (unsigned int)SCB_SM_COUNT * 8;
    if (elem_type == SCB_F32) minmax_kernel<float><<<grid, 256, 0, s>>>((const float *)d_frames, n, keys);
    else minmax_kernel<double><<<grid, 256, 0, s>>>((const double *)d_frames, n, keys);
    minmax_finish_kernel<<<1, 1, 0, s>>>(keys, d_minmax);
    SCB_CUDA_LAUNCH_CHECK("scb_frames_minmax");
    return 0;
}

extern "C" int scb_frames_to_8bit(const void *d_frames, int64_t n, int elem_type, const double *d_limits,
                                  double cmin, double cmax, double low, double high, uint8_t *d_out, void *stream) {
    SCB_REQUIRE(d_frames && d_out, SCB_E_NULL, "scb_frames_to_8bit: NULL pointer");
    SCB_REQUIRE(n > 0, SCB_E_INVALID, "scb_frames_to_8bit: n=%lld", (long long)n);
    SCB_REQUIRE(elem_type == SCB_F32 || elem_type == SCB_F64, SCB_E_INVALID, "elem_type=%d", elem_type);
    SCB_REQUIRE(((uintptr_t)d_out & 3) == 0, SCB_E_INVALID, "scb_frames_to_8bit: output must be 4-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned int want = scb_grid_for((n + 3) >> 2, 256, 4);
    const unsigned int grid = want < (unsigned int)SCB_SM_COUNT * 16 ? want : (unsigned int)SCB_SM_COUNT * 16;
    if (elem_type == SCB_F32)
        to_8bit_kernel<float><<<grid, 256, 0, s>>>((const float *)d_frames, n, d_limits, cmin, cmax, low, high, d_out);
    else
        to_8bit_kernel<double><<<grid, 256, 0, s>>>((const double *)d_frames, n, d_limits, cmin, cmax, low, high, d_out);
    SCB_CUDA_LAUNCH_CHECK("scb_frames_to_8bit");
    return 0;
}
