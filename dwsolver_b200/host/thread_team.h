// thread_team.h — a small persistent team of worker threads for loops that last a few microseconds.
//
// The restricted master's simplex iteration is a chain of dense L x L operations (FTRAN, rank-1 update of the
// working inverse) of ~10 us each at L = 256: far too short for a thread pool that sleeps between jobs, but the
// 512 KB inverse does not fit one core's L1/L2 either, so a single thread streams it from the outer cache on
// every iteration.  A team whose members each own a fixed slice of rows keeps every slice resident in its
// owner's cache.  Workers spin (with pause) while jobs keep coming, yield after a short while and go to sleep
// on a condition variable when the master is idle, so an idle master costs no CPU.
//
// run(f) executes f(member, members) on every member (the caller is member 0) and returns when all are done.
// No job may call run() itself.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define DWH_PAUSE() _mm_pause()
#else
#define DWH_PAUSE() std::this_thread::yield()
#endif

namespace dwh {

class ThreadTeam {
public:
    explicit ThreadTeam(int members) : n_(members < 1 ? 1 : members)
    {
        for (int t = 1; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
    }
    ~ThreadTeam()
    {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_.store(true, std::memory_order_release);
            seq_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    ThreadTeam(const ThreadTeam &) = delete;
    ThreadTeam &operator=(const ThreadTeam &) = delete;
    int size() const { return n_; }

    template <class F> void run(F &&f)
    {
        if (n_ == 1) { f(0, 1); return; }
        fn_ = [](void *ctx, int t, int n) { (*static_cast<typename std::remove_reference<F>::type *>(ctx))(t, n); };
        ctx_ = const_cast<void *>(static_cast<const void *>(&f));
        done_.store(0, std::memory_order_relaxed);
        seq_.fetch_add(1, std::memory_order_seq_cst);       // (seq_cst pair with the sleeper's count-then-check: no lost wake-up)
        if (sleepers_.load(std::memory_order_seq_cst) > 0) {
            std::lock_guard<std::mutex> g(mu_);
            cv_.notify_all();
        }
        f(0, n_);
        int spins = 0;
        while (done_.load(std::memory_order_acquire) != n_ - 1) {
            if (++spins < 4096) DWH_PAUSE(); else std::this_thread::yield();
        }
    }

private:
    void loop(int t)
    {
        unsigned long long seen = 0;
        for (;;) {
            // wait for the next job: spin, then yield, then sleep
            int spins = 0;
            auto idle_since = std::chrono::steady_clock::time_point();
            for (;;) {
                if (seq_.load(std::memory_order_acquire) != seen) break;
                ++spins;
                if (spins < 2048) { DWH_PAUSE(); continue; }
                if (spins == 2048) idle_since = std::chrono::steady_clock::now();
                std::this_thread::yield();
                if ((spins & 63) == 0 && std::chrono::steady_clock::now() - idle_since > std::chrono::milliseconds(2)) {
                    std::unique_lock<std::mutex> g(mu_);
                    sleepers_.fetch_add(1, std::memory_order_seq_cst);
                    cv_.wait(g, [&] { return seq_.load(std::memory_order_seq_cst) != seen; });
                    sleepers_.fetch_sub(1, std::memory_order_acq_rel);
                    break;
                }
            }
            if (stop_.load(std::memory_order_acquire)) return;
            seen = seq_.load(std::memory_order_acquire);
            fn_(ctx_, t, n_);
            done_.fetch_add(1, std::memory_order_release);
        }
    }

    const int n_;
    std::vector<std::thread> workers_;
    std::atomic<unsigned long long> seq_{0};
    std::atomic<int> done_{0};
    std::atomic<int> sleepers_{0};
    std::atomic<bool> stop_{false};
    void (*fn_)(void *, int, int) = nullptr;
    void *ctx_ = nullptr;
    std::mutex mu_;
    std::condition_variable cv_;
};

}  // namespace dwh
