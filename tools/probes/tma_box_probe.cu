// Probe: one tensor-map TMA copy (cp.async.bulk.tensor.3d) of an 8 x 36 box out of a [blocks][32][32] float tensor,
// starting at a negative row / a column that overhangs the block: are the out-of-range elements zero filled, and
// does the mbarrier see the full box size?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_box_probe tma_box_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int col, int row, int block, float *out) {
    __shared__ __align__(128) float stage[8 * 36];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar), dst = (unsigned)__cvta_generic_to_shared(stage);
    for (int i = threadIdx.x; i < 288; i += 32) stage[i] = -1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(1152) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(dst), "l"(&map), "r"(col), "r"(row), "r"(block), "r"(bar_a) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar_a) : "memory");
    for (int i = threadIdx.x; i < 288; i += 32) out[i] = stage[i];
}

int main() {
    const int blocks = 7, slots = 32;
    std::vector<float> h((size_t)blocks * slots * slots);
    for (int b = 0; b < blocks; ++b)
        for (int r = 0; r < slots; ++r)
            for (int c = 0; c < slots; ++c) h[((size_t)b * slots + r) * slots + c] = b * 10000 + r * 100 + c + 1;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&out, 1152);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)slots, (cuuint64_t)slots, 1000000};
    const cuuint64_t strides[2] = {slots * 4ull, slots * slots * 4ull};
    const cuuint32_t box[3] = {36, 8, 1}, elem[3] = {1, 1, 1};
    cuInit(0);
    CUresult rc = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, elem,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)rc);
    const int cases[][3] = {{0, 0, 0}, {0, -3, 2}, {4, 29, 6}, {0, -7, 3}, {28, 5, 1}, {1, 0, 0}};
    for (auto &c : cases) {
        probe<<<1, 32>>>(map, c[0], c[1], c[2], out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("col %d row %d block %d: %s\n", c[0], c[1], c[2], cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        float r[288];
        cudaMemcpy(r, out, 1152, cudaMemcpyDeviceToHost);
        for (int k = 0; k < 8; ++k) printf("  row %d: %g %g ... %g %g | %g %g %g %g\n", k, r[k * 36], r[k * 36 + 1], r[k * 36 + 30], r[k * 36 + 31], r[k * 36 + 32], r[k * 36 + 33], r[k * 36 + 34], r[k * 36 + 35]);
    }
    return 0;
}
