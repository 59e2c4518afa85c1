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
#include <thread>
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

    // 0, or the CUDA error the download of that (or an earlier) ticket ended with
    int wait(int64_t ticket) {
        std::unique_lock<std::mutex> lock(mu_);
        if (ticket < 1 || ticket >= next_ticket_) return -1;
        done_cv_.wait(lock, [&] { return done_ticket_ >= ticket; });
        return (int)error_;
    }

  private:
    void start_locked() {
        quit_ = false;
        running_ = true;
        for (int w = 0; w < n_workers_; ++w) workers_.emplace_back([this, w, g = generation_] { worker(w, g); });
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
                if (err != cudaSuccess) error_ = err;
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
    cudaError_t error_ = cudaSuccess;

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
