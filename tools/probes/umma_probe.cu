// Probe of the tcgen05.mma.kind::tf32 operand layout used by gaussian_tc.cu:
// MN-major, no swizzle, core matrix = 8 (K) x 16 B (4 MN elements); D[m][n] = sum_k A[k][m] B[k][n].
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int TM = 128;
constexpr uint32_t kLBO = 8 * TM * 4, kSBO = 128;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__global__ void probe(float *out, int swap, uint32_t idesc, int layout) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *A = (float *)smem;            // 8 KB: 16 spots x 128
    float *B = (float *)(smem + 8192);
    unsigned long long *bar = (unsigned long long *)(smem + 16384);
    uint32_t *tmem_base = (uint32_t *)(smem + 16384 + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    // A[k][m] = (k + 1) * 1000 + m ; B[k][n] = (k == n % 8) ? 1 : 0  -> D[m][n] = A[n%8][m]
    for (int idx = tid; idx < 16 * TM; idx += blockDim.x) {
        int k = idx / TM, m = idx % TM;
        int off = layout == 0 ? (k / 8) * kLBO + (m / 4) * kSBO + (k % 8) * 16 + (m % 4) * 4      // MN-major cores
                              : (k / 8) * 4096 + (m / 8) * 256 + ((k % 8) / 4) * 128 + (m % 8) * 16 + (k % 4) * 4;   // K-major cores
        *(float *)((char *)A + off) = (k < 8) ? (float)(k * 128 + m) : 0.f;
        *(float *)((char *)B + off) = (k < 8 && k == (m % 8)) ? 1.f : 0.f;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_base;
    if (tid == 0) {
        uint32_t lbo = layout == 0 ? (swap ? kSBO : kLBO) : (swap ? 256u : 128u), sbo = layout == 0 ? (swap ? kLBO : kSBO) : (swap ? 128u : 256u);
        uint64_t da = umma_desc(smem_u32(A), lbo, sbo), db = umma_desc(smem_u32(B), lbo, sbo);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        for (int cb = 0; cb < 4; ++cb) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
                "tcgen05.wait::ld.sync.aligned;"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32) : "memory");
            for (int c = 0; c < 32; ++c) out[tid * TM + cb * 32 + c] = __uint_as_float(v[c]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
}
int main() {
    float *d, *h = new float[TM * TM];
    cudaMalloc(&d, TM * TM * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    const uint32_t base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TM >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    for (int mn = 0; mn >= 0; --mn)
        for (int swap = 0; swap < 2; ++swap) {
            uint32_t idesc = base | (mn ? ((1u << 15) | (1u << 16)) : 0u);
            cudaMemset(d, 0xff, TM * TM * 4);
            probe<<<1, 128, 32768>>>(d, swap, idesc, 1);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, TM * TM * 4, cudaMemcpyDeviceToHost);
            int ok = 0, nz = 0;
            for (int m = 0; m < TM; ++m)
                for (int n = 0; n < TM; ++n) {
                    float want = (float)((n % 8) * 128 + m);
                    ok += (h[m * TM + n] == want);
                    nz += (h[m * TM + n] != 0.f);
                }
            printf("mn_major=%d swap_lbo_sbo=%d err=%s exact=%d/%d nonzero=%d  D[0][0..3]=%g %g %g %g D[5][1]=%g D[77][9]=%g\n", mn, swap,
                   cudaGetErrorString(e), ok, TM * TM, nz, h[0], h[1], h[2], h[3], h[5 * TM + 1], h[77 * TM + 9]);
        }
    return 0;
}
