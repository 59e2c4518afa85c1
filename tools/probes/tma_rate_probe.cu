// Probe: how many small TMA bulk copies (cp.async.bulk, S bytes from random 4 KB-aligned blocks of a 16 GB buffer)
// can one SM retire per cycle?  Each warp keeps STAGES-1 copies in flight in its own ring and does nothing else.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_rate_probe tma_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t sa(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <int STAGES>
__global__ void __launch_bounds__(256) rate(const char *__restrict__ buf, uint64_t n_blocks, int bytes, int per_warp, unsigned *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *ring = smem + warp * (STAGES * 1024 + 64);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(ring + STAGES * 1024);
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned acc = 0;
    auto issue = [&](int i) {
        if (lane == 0) {
            const int s = i % STAGES;
            const char *src = buf + (mix(gw * 1000003ull + i) % n_blocks) * 4096 + ((i * 7) & 3) * 1024;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(&bars[s])), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa(ring + s * 1024)),
                         "l"(src), "r"(bytes), "r"(sa(&bars[s])) : "memory");
        }
    };
    for (int i = 0; i < STAGES - 1 && i < per_warp; ++i) issue(i);
    for (int i = 0; i < per_warp; ++i) {
        __syncwarp();
        if (i + STAGES - 1 < per_warp) issue(i + STAGES - 1);
        const int s = i % STAGES;
        const unsigned parity = (i / STAGES) & 1;
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(sa(&bars[s])), "r"(parity) : "memory");
        acc += reinterpret_cast<const unsigned *>(ring + s * 1024)[lane];
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int STAGES>
void run(const char *buf, size_t total, int bytes, int warps_per_cta, int ctas_per_sm, unsigned *sink) {
    const uint64_t n_blocks = total / 4096;
    const int per_warp = 2000;
    const int blocks = 148 * ctas_per_sm;
    const size_t smem = (size_t)warps_per_cta * (STAGES * 1024 + 64);
    cudaFuncSetAttribute(rate<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    rate<STAGES><<<blocks, warps_per_cta * 32, smem>>>(buf, n_blocks, bytes, per_warp, sink);
    cudaEventRecord(a);
    rate<STAGES><<<blocks, warps_per_cta * 32, smem>>>(buf, n_blocks, bytes, per_warp, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double copies = (double)blocks * warps_per_cta * per_warp;
    printf("S=%4d B stages %d warps/SM %2d: %.3f ms, %.0f GB/s, %.1f cycles per copy per SM  (%s)\n", bytes, STAGES,
           warps_per_cta * ctas_per_sm, ms, copies * bytes / ms / 1e6, ms * 1e-3 * 1.965e9 / (copies / 148), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t total = (size_t)16 << 30;
    char *buf; unsigned *sink;
    cudaMalloc(&buf, total); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, total);
    for (int bytes : {128, 256, 512, 640, 1024}) {
        run<4>(buf, total, bytes, 8, 3, sink);
        run<4>(buf, total, bytes, 8, 4, sink);
        run<8>(buf, total, bytes, 8, 3, sink);
    }
    return 0;
}
