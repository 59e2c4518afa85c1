// Probe: HBM read bandwidth for RANDOM aligned chunks of S bytes out of a 16 GB buffer (the access pattern of the
// box-table render: one unit = ~1 KB of a 4 KB block chosen by the spot's depth and sub-pixel phase).
// Each warp reads whole chunks with 16-byte loads, 4 chunks in flight.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <int S>
__global__ void __launch_bounds__(256) gather(const uint4 *__restrict__ buf, uint64_t n_chunks, int per_warp, unsigned *sink) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned acc = 0;
    constexpr int kPieces = S / 512;           // 16-byte loads per lane per chunk
    for (int i = 0; i < per_warp; i += 4) {
        uint4 v[4][kPieces];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint64_t chunk = mix(warp * 1000003ull + i + c) % n_chunks;
            const uint4 *p = buf + chunk * (S / 16) + lane;
#pragma unroll
            for (int k = 0; k < kPieces; ++k) v[c][k] = __ldg(p + k * 32);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int k = 0; k < kPieces; ++k) acc += v[c][k].x ^ v[c][k].y ^ v[c][k].z ^ v[c][k].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int S>
void run(const uint4 *buf, size_t bytes, unsigned *sink) {
    const uint64_t n_chunks = bytes / S;
    const int per_warp = (512 * 1024 / S + 3) / 4 * 4;       // ~0.5 MB per warp
    const int blocks = 148 * 8 * 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather<S><<<blocks, 256>>>(buf, n_chunks, per_warp, sink);
    cudaEventRecord(a);
    gather<S><<<blocks, 256>>>(buf, n_chunks, per_warp, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double total = (double)blocks * 8 * per_warp * S;
    printf("chunk %5d B: %.1f GB in %.3f ms = %.0f GB/s  (%s)\n", S, total / 1e9, ms, total / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t bytes = (size_t)16 << 30;
    uint4 *buf; unsigned *sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    run<512>(buf, bytes, sink);
    run<1024>(buf, bytes, sink);
    run<2048>(buf, bytes, sink);
    run<4096>(buf, bytes, sink);
    run<8192>(buf, bytes, sink);
    return 0;
}
