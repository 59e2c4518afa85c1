// Host side of the end-to-end path: frames leave the device as float32 (half the PCIe bytes
// of the float64 planes scopyon's Image carries, image.py:12-36) into pinned staging memory and
// are widened to float64 by a small pool of host threads while the following frames are in
// flight.  (double)float is exact, so the arrays handed to the caller are bit for bit the
// ones a device-side widening would produce.
//
// One dispatcher thread takes jobs in FIFO order: it waits for the job's CUDA event (the
// download), splits the frame over the workers, and publishes the ticket when they are done.
// Stores bypass the cache (the destination is 2 x the L2-sized source and is read later by
// the caller, not by these threads).  No GPU work happens here.
#include <cuda_runtime_api.h>
#include <emmintrin.h>
#include <immintrin.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <pthread.h>
#include <sched.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/scopyon_b200.h"

void scb_set_error(const char *fmt, ...);

namespace {

__attribute__((target("avx2"))) void widen_avx2(const float *s, double *d, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(d + i) & 31u)) { d[i] = (double)s[i]; ++i; }
    for (; i + 8 <= n; i += 8) {
        const __m256 v = _mm256_loadu_ps(s + i);
        _mm256_stream_pd(d + i, _mm256_cvtps_pd(_mm256_castps256_ps128(v)));
        _mm256_stream_pd(d + i + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1)));
    }
    for (; i < n; ++i) d[i] = (double)s[i];
    _mm_sfence();
}

void widen_sse2(const float *s, double *d, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(d + i) & 15u)) { d[i] = (double)s[i]; ++i; }
    for (; i + 4 <= n; i += 4) {
        const __m128 v = _mm_loadu_ps(s + i);
        _mm_stream_pd(d + i, _mm_cvtps_pd(v));
        _mm_stream_pd(d + i + 2, _mm_cvtps_pd(_mm_movehl_ps(v, v)));
    }
    for (; i < n; ++i) d[i] = (double)s[i];
    _mm_sfence();
}

void widen_range(const float *s, double *d, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) widen_avx2(s, d, n);
    else widen_sse2(s, d, n);
}

struct Job {
    const float *src;
    double *dst;
    size_t n;
    cudaEvent_t event;     // may be null: the source is ready
    int device;
    int64_t ticket;
};

class WidenPool {
  public:
    ~WidenPool() { stop(); }

    int threads() {
        std::lock_guard<std::mutex> lock(mu_);
        return n_workers_;
    }

    // Takes effect when the queue is idle; returns the previous setting.
    int set_threads(int n) {
        std::lock_guard<std::mutex> control(control_);     // no submit() while the pool is rebuilt
        std::unique_lock<std::mutex> lock(mu_);
        const int old = n_workers_;
        if (n >= 1 && n <= 64 && n != n_workers_) {
            done_cv_.wait(lock, [&] { return done_ticket_ == next_ticket_ - 1; });
            lock.unlock();
            stop();
            lock.lock();
            n_workers_ = n;
        }
        return old;
    }

    int64_t submit(const float *src, double *dst, size_t n, cudaEvent_t event, int device) {
        std::lock_guard<std::mutex> control(control_);
        std::lock_guard<std::mutex> lock(mu_);
        if (!running_) start_locked();
        Job job = {src, dst, n, event, device, next_ticket_++};
        queue_.push_back(job);
        queue_cv_.notify_one();
        return job.ticket;
    }

    // 0, or the CUDA error the download of THIS ticket ended with (reported once, then forgotten:
    // one failed download does not poison later frames, engines or devices)
    int wait(int64_t ticket) {
        std::unique_lock<std::mutex> lock(mu_);
        if (ticket < 1 || ticket >= next_ticket_) return -1;
        done_cv_.wait(lock, [&] { return done_ticket_ >= ticket; });
        const auto it = errors_.find(ticket);
        if (it == errors_.end()) return 0;
        const int err = (int)it->second;
        errors_.erase(it);
        return err;
    }

    // Worker w runs on cpus[w % n] from the next start of the pool (n = 0: no pinning).
    void set_affinity(const int *cpus, int n) {
        std::lock_guard<std::mutex> control(control_);
        std::unique_lock<std::mutex> lock(mu_);
        done_cv_.wait(lock, [&] { return done_ticket_ == next_ticket_ - 1; });
        lock.unlock();
        stop();
        lock.lock();
        cpus_.assign(cpus, cpus + (n > 0 ? n : 0));
    }

  private:
    void start_locked() {
        quit_ = false;
        running_ = true;
        for (int w = 0; w < n_workers_; ++w) {
            workers_.emplace_back([this, w, g = generation_] { worker(w, g); });
            if (!cpus_.empty()) {
                cpu_set_t set;
                CPU_ZERO(&set);
                CPU_SET(cpus_[w % cpus_.size()], &set);
                pthread_setaffinity_np(workers_.back().native_handle(), sizeof(set), &set);    // best effort
            }
        }
        dispatcher_ = std::thread([this] { dispatch(); });
    }

    void stop() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            if (!running_) return;
            quit_ = true;
            queue_cv_.notify_all();
            work_cv_.notify_all();
        }
        dispatcher_.join();
        for (auto &t : workers_) t.join();
        workers_.clear();
        std::lock_guard<std::mutex> lock(mu_);
        running_ = false;
    }

    void dispatch() {
        int device = -1;
        for (;;) {
            Job job;
            {
                std::unique_lock<std::mutex> lock(mu_);
                queue_cv_.wait(lock, [&] { return quit_ || !queue_.empty(); });
                if (queue_.empty()) return;
                job = queue_.front();
                queue_.pop_front();
            }
            cudaError_t err = cudaSuccess;
            if (job.event) {
                if (job.device != device) { cudaSetDevice(job.device); device = job.device; }
                err = cudaEventSynchronize(job.event);
            }
            if (err == cudaSuccess && job.n) {
                std::unique_lock<std::mutex> lock(mu_);
                current_ = job;
                parts_left_ = n_workers_;
                ++generation_;
                work_cv_.notify_all();
                parts_cv_.wait(lock, [&] { return parts_left_ == 0; });
            }
            {
                std::lock_guard<std::mutex> lock(mu_);
                if (err != cudaSuccess) {
                    if (errors_.size() >= 1024) errors_.clear();    // never awaited: do not grow without bound
                    errors_[job.ticket] = err;
                }
                done_ticket_ = job.ticket;
                done_cv_.notify_all();
            }
        }
    }

    void worker(int w, uint64_t seen) {
        for (;;) {
            Job job;
            int n_workers;
            {
                std::unique_lock<std::mutex> lock(mu_);
                work_cv_.wait(lock, [&] { return quit_ || generation_ != seen; });
                if (generation_ == seen) return;
                seen = generation_;
                job = current_;
                n_workers = n_workers_;
            }
            // 64-element (256 / 512 byte) granules keep every thread on whole cache lines
            const size_t granules = (job.n + 63) / 64;
            const size_t a = granules * w / n_workers * 64, b = granules * (w + 1) / n_workers * 64;
            const size_t lo = a < job.n ? a : job.n, hi = b < job.n ? b : job.n;
            if (hi > lo) widen_range(job.src + lo, job.dst + lo, hi - lo);
            {
                std::lock_guard<std::mutex> lock(mu_);
                if (--parts_left_ == 0) parts_cv_.notify_one();
            }
        }
    }

    std::mutex control_;     // serialises submit() and set_threads()
    std::mutex mu_;
    std::condition_variable queue_cv_, work_cv_, parts_cv_, done_cv_;
    std::deque<Job> queue_;
    std::vector<std::thread> workers_;
    std::thread dispatcher_;
    Job current_ = {};
    uint64_t generation_ = 0;
    int parts_left_ = 0;
    int n_workers_ = 0;
    bool running_ = false, quit_ = false;
    int64_t next_ticket_ = 1, done_ticket_ = 0;
    std::unordered_map<int64_t, cudaError_t> errors_;     // failed downloads by ticket, until waited for
    std::vector<int> cpus_;                                // worker affinity (empty: none)

  public:
    WidenPool() {
        const unsigned hw = std::thread::hardware_concurrency();
        // measured on the B200 boxes (16 hardware threads): 6 workers stream ~70 GB/s of float64, more only
        // contend with the DMA, the dispatcher and the caller's thread
        n_workers_ = hw >= 12 ? 6 : hw >= 4 ? (int)hw / 2 : 1;
    }
};

WidenPool &pool() {
    static WidenPool *p = new WidenPool();   // never destroyed: no thread joins during process exit
    return *p;
}

}  // namespace

extern "C" int scb_host_widen_threads(int n_threads) { return pool().set_threads(n_threads); }

extern "C" int scb_host_widen_affinity(const int *cpus, int n_cpus) {
    if (n_cpus < 0 || n_cpus > 1024 || (n_cpus > 0 && !cpus)) {
        scb_set_error("scb_host_widen_affinity: cpus=%p n_cpus=%d", (const void *)cpus, n_cpus);
        return SCB_E_INVALID;
    }
    for (int i = 0; i < n_cpus; ++i)
        if (cpus[i] < 0 || cpus[i] >= CPU_SETSIZE) {
            scb_set_error("scb_host_widen_affinity: cpu %d out of range", cpus[i]);
            return SCB_E_INVALID;
        }
    pool().set_affinity(cpus, n_cpus);
    return 0;
}

// Host memory bandwidth with the access patterns of the end-to-end path, on `n_threads` threads
// over private buffers (first touched by the thread that uses them): best of `repeats` passes.
//   mode 0  copy, streaming stores: read b + write b per byte of source  (STREAM "copy")
//   mode 1  float32 -> float64 widening as scb_host_widen_* does it: read b + write 2 b
// Returns bytes moved per second (reads + writes) in *bytes_per_s.
extern "C" int scb_host_bandwidth(int mode, int64_t bytes_per_thread, int n_threads, int repeats, double *bytes_per_s) {
    if ((mode != 0 && mode != 1) || bytes_per_thread < 4096 || n_threads < 1 || n_threads > 256 || repeats < 1 ||
        !bytes_per_s) {
        scb_set_error("scb_host_bandwidth: mode=%d bytes=%lld threads=%d repeats=%d", mode, (long long)bytes_per_thread,
                      n_threads, repeats);
        return SCB_E_INVALID;
    }
    const size_t n = (size_t)bytes_per_thread / 4;              // float32 elements of source per thread
    const size_t dst_bytes = mode == 1 ? n * 8 : n * 4;
    std::vector<float *> src(n_threads, nullptr);
    std::vector<void *> dst(n_threads, nullptr);
    std::atomic<int> ready{0}, failed{0};
    std::atomic<int> go{0};
    std::vector<double> seconds((size_t)n_threads * repeats, 0.0);
    std::vector<std::thread> threads;
    for (int t = 0; t < n_threads; ++t) {
        threads.emplace_back([&, t] {
            if (posix_memalign((void **)&src[t], 64, n * 4) || posix_memalign(&dst[t], 64, dst_bytes)) {
                failed.fetch_add(1);
            } else {
                for (size_t i = 0; i < n; ++i) src[t][i] = (float)(i & 1023);
                memset(dst[t], 0, dst_bytes);
            }
            ready.fetch_add(1);
            for (int r = 0; r < repeats; ++r) {
                while (go.load(std::memory_order_acquire) <= r) std::this_thread::yield();
                if (failed.load()) continue;
                const auto t0 = std::chrono::steady_clock::now();
                if (mode == 1) {
                    widen_range(src[t], (double *)dst[t], n);
                } else {
                    // the same streaming-store path, float for float
                    const float *s = src[t];
                    float *d = (float *)dst[t];
                    for (size_t i = 0; i + 4 <= n; i += 4) _mm_stream_ps(d + i, _mm_load_ps(s + i));
                    _mm_sfence();
                }
                seconds[(size_t)t * repeats + r] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            }
        });
    }
    while (ready.load() < n_threads) std::this_thread::yield();
    double best = 0.0;
    for (int r = 0; r < repeats; ++r) {
        const auto t0 = std::chrono::steady_clock::now();
        go.store(r + 1, std::memory_order_release);
        // a pass ends when the slowest thread has finished: wait by joining on the last repeat, by polling before
        for (;;) {
            bool all = true;
            for (int t = 0; t < n_threads; ++t) all = all && (seconds[(size_t)t * repeats + r] > 0.0 || failed.load());
            if (all) break;
            std::this_thread::yield();
        }
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double moved = (double)n_threads * ((double)n * 4 + (double)dst_bytes);
        if (wall > 0.0 && moved / wall > best) best = moved / wall;
    }
    for (auto &th : threads) th.join();
    for (int t = 0; t < n_threads; ++t) { free(src[t]); free(dst[t]); }
    if (failed.load()) {
        scb_set_error("scb_host_bandwidth: out of memory");
        return SCB_E_INVALID;
    }
    *bytes_per_s = best;
    return 0;
}

extern "C" int64_t scb_host_widen_start(const float *h_src, double *h_dst, int64_t n, void *cuda_event, int device) {
    if (!h_src || !h_dst || n < 0) {
        scb_set_error("scb_host_widen_start: src=%p dst=%p n=%lld", (const void *)h_src, (void *)h_dst, (long long)n);
        return -1;
    }
    return pool().submit(h_src, h_dst, (size_t)n, (cudaEvent_t)cuda_event, device);
}

extern "C" int scb_host_widen_wait(int64_t ticket) {
    const int err = pool().wait(ticket);
    if (err == -1) {
        scb_set_error("scb_host_widen_wait: unknown ticket %lld", (long long)ticket);
        return SCB_E_INVALID;
    }
    if (err != 0) scb_set_error("scb_host_widen_wait: download failed: %s", cudaGetErrorString((cudaError_t)err));
    return err;
}
