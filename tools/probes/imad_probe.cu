// Throughput probe: 32x32->64 products as IMAD.WIDE.U32 vs IMAD.HI.U32 + IMAD (lo) on sm_100a.
// Each thread runs 8 independent chains; prints warp-instructions per clock per SM for the multiplies.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t *out, int iters, uint32_t m) {
    uint32_t x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 8 + i + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t hi, lo;
                if (MODE == 0) {
                    asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1, %0}, p;\n\t}" : "=r"(hi), "=r"(lo) : "r"(x[i]), "r"(m));
                } else if (MODE == 1) {
                    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(x[i]), "r"(m));
                    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(x[i]), "r"(m));
                } else if (MODE == 2) {
                    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(x[i]), "r"(m));
                    lo = x[i];
                } else {
                    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(x[i]), "r"(m));
                    hi = 0x9E3779B9u;
                }
                x[i] = hi ^ lo ^ 0x5bd1e995u;
            }
        }
    }
    uint32_t acc = 0;
    for (int i = 0; i < 8; ++i) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, int mults_per_step) {
    const int blocks = 148 * 8, iters = 2000;
    uint32_t *out;
    cudaMalloc(&out, blocks * 256 * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<MODE><<<blocks, 256>>>(out, 10, 0xD2511F53u);
    cudaEventRecord(a);
    probe<MODE><<<blocks, 256>>>(out, iters, 0xD2511F53u);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double steps = (double)blocks * 8 /*warps*/ * iters * 64;   // warp-level chain steps
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("%-28s %8.3f ms  %.3f chain-steps/clk/SM  (%d mult instr per step -> %.3f mult warp-instr/clk/SM)\n", name, ms,
           steps / cycles / 148, mults_per_step, steps * mults_per_step / cycles / 148);
    cudaFree(out);
}

int main() {
    run<0>("IMAD.WIDE.U32", 1);
    run<1>("IMAD.HI.U32 + IMAD(lo)", 2);
    run<2>("IMAD.HI.U32 only", 1);
    run<3>("IMAD(lo) only", 1);
    return 0;
}
