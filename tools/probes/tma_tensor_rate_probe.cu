// Probe: retire rate of tensor-map TMA copies (cp.async.bulk.tensor.3d) per SM as a function of the box height R:
// boxes of R rows x 36 columns out of random blocks of a [blocks][32][32] float tensor (16 GB), starting at a random
// row in [-(R-1), 31] (rows outside the block are zero filled, as in the render).  Each warp keeps STAGES-1 copies in
// flight and reads one word per lane of every box.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a ... -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t sa(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <int STAGES, int R>
__global__ void __launch_bounds__(256) rate(const __grid_constant__ CUtensorMap map, uint32_t n_blocks, int per_warp, unsigned *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int kStage = R * 36 * 4;
    constexpr int kWarp = (STAGES * kStage + 64 + 127) & ~127;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *ring = smem + warp * kWarp;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(ring + STAGES * kStage);
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
            const uint64_t h = mix(gw * 1000003ull + i);
            const int block = (int)(h % n_blocks), row = (int)((h >> 40) % (31 + R)) - (R - 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(&bars[s])), "r"(kStage) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(sa(ring + s * kStage)), "l"(&map), "r"(0), "r"(row), "r"(block), "r"(sa(&bars[s])) : "memory");
        }
    };
    for (int i = 0; i < STAGES - 1 && i < per_warp; ++i) issue(i);
    for (int i = 0; i < per_warp; ++i) {
        __syncwarp();
        if (i + STAGES - 1 < per_warp) issue(i + STAGES - 1);
        const int s = i % STAGES;
        const unsigned parity = (i / STAGES) & 1;
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(sa(&bars[s])), "r"(parity) : "memory");
        acc += reinterpret_cast<const unsigned *>(ring + s * kStage)[lane];
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int STAGES, int R>
void run(void *buf, uint32_t n_blocks, int warps_per_cta, int ctas_per_sm, unsigned *sink) {
    CUtensorMap map;
    const cuuint64_t dims[3] = {32, 32, n_blocks};
    const cuuint64_t strides[2] = {128, 4096};
    const cuuint32_t box[3] = {36, (cuuint32_t)R, 1}, elem[3] = {1, 1, 1};
    CUresult rc = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, strides, box, elem,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); return; }
    const int per_warp = 1500;
    const int blocks = 148 * ctas_per_sm;
    const size_t smem = (size_t)warps_per_cta * ((STAGES * R * 144 + 64 + 127) & ~127);
    cudaFuncSetAttribute(rate<STAGES, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    rate<STAGES, R><<<blocks, warps_per_cta * 32, smem>>>(map, n_blocks, per_warp, sink);
    cudaEventRecord(a);
    rate<STAGES, R><<<blocks, warps_per_cta * 32, smem>>>(map, n_blocks, per_warp, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double copies = (double)blocks * warps_per_cta * per_warp;
    // rows actually inside the block: start uniform in [-(R-1), 31] -> mean overlap
    double mean_rows = 0;
    for (int r = -(R - 1); r <= 31; ++r) { int lo = r < 0 ? 0 : r, hi = r + R > 32 ? 32 : r + R; mean_rows += hi - lo; }
    mean_rows /= (31 + R);
    printf("R=%2d stages %d warps/SM %2d (smem %3zu KB/SM): %.3f ms, %.1f cycles per copy per SM, ~%.0f GB/s of table rows  (%s)\n", R, STAGES,
           warps_per_cta * ctas_per_sm, smem * ctas_per_sm / 1024, ms, ms * 1e-3 * 1.965e9 / (copies / 148), copies * mean_rows * 128 / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cuInit(0);
    const size_t total = (size_t)16 << 30;
    void *buf; unsigned *sink;
    cudaMalloc(&buf, total); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, total);
    const uint32_t n_blocks = (uint32_t)(total / 4096);
    run<4, 8>(buf, n_blocks, 8, 3, sink);
    run<4, 8>(buf, n_blocks, 4, 4, sink);
    run<4, 16>(buf, n_blocks, 8, 2, sink);
    run<4, 16>(buf, n_blocks, 4, 4, sink);
    run<4, 32>(buf, n_blocks, 4, 2, sink);
    run<4, 32>(buf, n_blocks, 2, 4, sink);
    run<4, 32>(buf, n_blocks, 1, 6, sink);
    run<8, 32>(buf, n_blocks, 1, 4, sink);
    run<4, 64>(buf, n_blocks, 1, 4, sink);
    run<4, 64>(buf, n_blocks, 2, 2, sink);
    return 0;
}
