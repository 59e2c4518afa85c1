// Host-side probe: how fast can T threads widen a 2048x2048 fp32 frame to float64
// (the dtype scopyon's Image carries), with and without non-temporal stores?
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>

static void widen_plain(const float *s, double *d, size_t n) {
    for (size_t i = 0; i < n; ++i) d[i] = (double)s[i];
}
__attribute__((target("avx2"))) static void widen_stream(const float *s, double *d, size_t n) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        __m256 v = _mm256_loadu_ps(s + i);
        _mm256_stream_pd(d + i, _mm256_cvtps_pd(_mm256_castps256_ps128(v)));
        _mm256_stream_pd(d + i + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1)));
    }
    for (; i < n; ++i) d[i] = (double)s[i];
    _mm_sfence();
}

int main() {
    const size_t n = 2048 * 2048;
    printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
    float *src = (float *)aligned_alloc(64, n * 4);
    const int ring = 8;                         // rotate destinations: no cache reuse between frames
    std::vector<double *> dst(ring);
    for (auto &p : dst) { p = (double *)aligned_alloc(64, n * 8); memset(p, 0, n * 8); }
    for (size_t i = 0; i < n; ++i) src[i] = (float)i;
    for (int mode = 0; mode < 2; ++mode)
        for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
            if (T > (int)std::thread::hardware_concurrency()) break;
            const int reps = 40;
            auto t0 = std::chrono::steady_clock::now();
            for (int r = 0; r < reps; ++r) {
                std::vector<std::thread> th;
                double *d = dst[r % ring];
                for (int t = 0; t < T; ++t) {
                    const size_t a = n * t / T / 8 * 8, b = (t + 1 == T) ? n : n * (t + 1) / T / 8 * 8;
                    th.emplace_back([=] { (mode ? widen_stream : widen_plain)(src + a, d + a, b - a); });
                }
                for (auto &x : th) x.join();
            }
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
            printf("%s threads %2d: %.3f ms per frame (%.1f GB/s written)\n", mode ? "stream" : "plain ", T, ms, n * 8 / ms / 1e6);
        }
    return 0;
}
