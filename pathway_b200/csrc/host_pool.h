// Small persistent worker pool for the host side of the index: staging copies (a single core moves ~10 GB/s; the 6.3 MB
// f64 query batch of a 1024 x 768 search would otherwise cost ~0.9 ms of a 12 ms step), the usability test, and the
// hash-table phases of delta batches.  Process-wide, created on first use, shared by all handles.
//
// A delta batch runs several short parallel regions back to back (tens of microseconds each), so the cost of STARTING a
// region matters as much as its body: part hand-out and completion are lock-free atomics, and a worker that finished a
// region keeps polling for the next one for ~400 us before it blocks on the condition variable (waking a sleeping
// thread costs 30-60 us, more than many of the regions last).  Plain C++: tests/native/host_pool_test.cpp.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define PWKNN_CPU_RELAX() _mm_pause()
#else
#define PWKNN_CPU_RELAX() std::this_thread::yield()
#endif

namespace pwknn {

class HostPool {
public:
    static HostPool& get() { static HostPool p; return p; }
    explicit HostPool(int threads_total = 0) {
        int n;
        if (threads_total > 0) {
            n = threads_total - 1;
        } else {
            const unsigned hc = std::thread::hardware_concurrency();
            // caller + workers: 4 copy streams saturate a small host's memory system, a 64+-thread box (the 8-GPU servers:
            // 8 ranks x 8 threads) feeds the staging ring faster with 8
            n = hc >= 64 ? 7 : (hc >= 8 ? 3 : (hc >= 4 ? 1 : 0));
            if (const char* e = std::getenv("PWKNN_HOST_THREADS")) n = std::max(0, std::atoi(e) - 1);
        }
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { worker(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> g(mu_); stop_.store(true, std::memory_order_release); }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    HostPool(const HostPool&) = delete;
    HostPool& operator=(const HostPool&) = delete;

    // runs fn(part, nparts) for part in [0, nparts) on the workers + the caller; returns when all are done.
    // Must not be called from inside a region (one region at a time).
    void parallel(int nparts, const std::function<void(int, int)>& fn) {
        if (nparts <= 1 || workers_.empty()) { for (int i = 0; i < nparts; ++i) fn(i, nparts); return; }
        std::lock_guard<std::mutex> api(api_mu_);
        fn_ = &fn;
        pending_.store(nparts, std::memory_order_relaxed);
        // part count and next part index live in ONE atomic word: a worker that arrives late (still leaving the previous
        // region) either sees the old word (index >= old count: nothing to do) or the new one -- never an old index
        // against a new count
        ticket_.store(static_cast<uint64_t>(static_cast<uint32_t>(nparts)) << 32, std::memory_order_release);  // publishes fn_
        epoch_.fetch_add(1, std::memory_order_release);
        if (sleepers_.load(std::memory_order_acquire) > 0) { std::lock_guard<std::mutex> g(mu_); cv_.notify_all(); }
        run_parts();
        while (pending_.load(std::memory_order_acquire) != 0) PWKNN_CPU_RELAX();
        if (failed_.load(std::memory_order_acquire)) {  // an exception on a worker thread is re-thrown on the caller's
            std::exception_ptr e;
            { std::lock_guard<std::mutex> g(mu_); e = err_; err_ = nullptr; }
            failed_.store(false, std::memory_order_release);
            if (e) std::rethrow_exception(e);
        }
    }
    int threads() const { return static_cast<int>(workers_.size()) + 1; }

private:
    void run_parts() {
        for (;;) {
            const uint64_t v = ticket_.fetch_add(1, std::memory_order_acq_rel);
            const uint32_t part = static_cast<uint32_t>(v), nparts = static_cast<uint32_t>(v >> 32);
            if (part >= nparts) return;
            try {
                (*fn_)(static_cast<int>(part), static_cast<int>(nparts));
            } catch (...) {  // never let an exception escape a worker thread (std::terminate): hand the first one over
                std::lock_guard<std::mutex> g(mu_);
                if (!err_) err_ = std::current_exception();
                failed_.store(true, std::memory_order_release);
            }
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            // poll for the next region for a while, then sleep
            const auto t0 = std::chrono::steady_clock::now();
            bool got = false;
            for (int spin = 0;; ++spin) {
                if (stop_.load(std::memory_order_acquire)) return;
                if (epoch_.load(std::memory_order_acquire) != seen) { got = true; break; }
                PWKNN_CPU_RELAX();
                if ((spin & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(400)) break;
            }
            if (!got) {
                std::unique_lock<std::mutex> g(mu_);
                sleepers_.fetch_add(1, std::memory_order_acq_rel);
                cv_.wait(g, [&] { return stop_.load(std::memory_order_acquire) || epoch_.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1, std::memory_order_acq_rel);
                if (stop_.load(std::memory_order_acquire)) return;
            }
            seen = epoch_.load(std::memory_order_acquire);
            run_parts();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, api_mu_;
    std::condition_variable cv_;
    const std::function<void(int, int)>* fn_ = nullptr;
    std::atomic<uint64_t> ticket_{0};  // (part count << 32) | next part index
    std::atomic<int> pending_{0}, sleepers_{0};
    std::atomic<uint64_t> epoch_{0};
    std::atomic<bool> stop_{false}, failed_{false};
    std::exception_ptr err_;
};

}  // namespace pwknn
