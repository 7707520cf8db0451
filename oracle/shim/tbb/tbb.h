/* ORACLE / TEST INFRASTRUCTURE. Minimal std::thread-backed stand-in for the
 * subset of Intel TBB the reference host code uses (parallel_for,
 * parallel_reduce, blocked_range, enumerable_thread_specific, task_group;
 * uses listed in SURVEY.md 2a). A persistent worker pool executes chunks; the
 * calling thread participates, so nested/concurrent submissions (TileScheduler
 * runs several render threads, tile_scheduler.h:69-78) cannot deadlock. */
#ifndef ORACLE_TBB_SHIM_H
#define ORACLE_TBB_SHIM_H
#include <atomic>
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

namespace tbb {

template <typename It>
class blocked_range {
public:
    blocked_range(It b, It e, size_t grain = 1) : b_(b), e_(e), grain_(grain) {}
    It begin() const { return b_; }
    It end() const { return e_; }
    size_t size() const { return size_t(e_ - b_); }
    bool empty() const { return !(b_ < e_); }
    size_t grainsize() const { return grain_; }
private:
    It b_, e_; size_t grain_;
};

namespace shim_detail {

struct Job {
    std::function<void(size_t)> run_chunk;
    size_t chunks;
    std::atomic<size_t> next{0};
    std::atomic<size_t> done{0};
};

class Pool {
public:
    static Pool& get() { static Pool p; return p; }
    int threads() const { return nthreads_; }

    void run(size_t chunks, const std::function<void(size_t)>& fn) {
        if (chunks == 0) return;
        if (chunks == 1 || nthreads_ <= 1) { for (size_t i = 0; i < chunks; ++i) fn(i); return; }
        auto job = std::make_shared<Job>();
        job->run_chunk = fn; job->chunks = chunks;
        { std::lock_guard<std::mutex> l(m_); jobs_.push_back(job); }
        cv_.notify_all();
        work_on(*job);
        // wait for chunks picked up by workers
        while (job->done.load(std::memory_order_acquire) < chunks) std::this_thread::yield();
        std::lock_guard<std::mutex> l(m_);
        auto it = std::find(jobs_.begin(), jobs_.end(), job);
        if (it != jobs_.end()) jobs_.erase(it);
    }
private:
    Pool() {
        const char* e = getenv("IMBA_ORACLE_THREADS");
        nthreads_ = e ? atoi(e) : (int)std::thread::hardware_concurrency();
        if (nthreads_ < 1) nthreads_ = 1;
        for (int i = 0; i < nthreads_ - 1; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    static void work_on(Job& j) {
        for (;;) {
            size_t i = j.next.fetch_add(1, std::memory_order_relaxed);
            if (i >= j.chunks) break;
            j.run_chunk(i);
            j.done.fetch_add(1, std::memory_order_release);
        }
    }
    void loop() {
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] {
                    if (stop_) return true;
                    for (auto& j : jobs_) if (j->next.load(std::memory_order_relaxed) < j->chunks) return true;
                    return false;
                });
                if (stop_) return;
                for (auto& j : jobs_) if (j->next.load(std::memory_order_relaxed) < j->chunks) { job = j; break; }
            }
            if (job) work_on(*job);
        }
    }
    int nthreads_;
    bool stop_ = false;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<Job>> jobs_;
    std::vector<std::thread> workers_;
};

inline size_t pick_chunks(size_t n) {
    size_t t = (size_t)Pool::get().threads();
    size_t c = t * 8;
    if (c > n) c = n;
    if (n < 256) c = 1;       // tiny ranges: serial
    return c ? c : 1;
}

} // namespace shim_detail

template <typename It, typename Body>
void parallel_for(const blocked_range<It>& r, const Body& body) {
    if (r.empty()) return;
    const size_t n = r.size();
    const size_t chunks = shim_detail::pick_chunks(n);
    const It b = r.begin();
    shim_detail::Pool::get().run(chunks, [&](size_t c) {
        size_t lo = n * c / chunks, hi = n * (c + 1) / chunks;
        if (lo < hi) body(blocked_range<It>(b + lo, b + hi));
    });
}

template <typename It, typename T, typename Body, typename Join>
T parallel_reduce(const blocked_range<It>& r, const T& identity, const Body& body, const Join& join) {
    if (r.empty()) return identity;
    const size_t n = r.size();
    const size_t chunks = shim_detail::pick_chunks(n);
    std::vector<T> partial(chunks, identity);
    const It b = r.begin();
    shim_detail::Pool::get().run(chunks, [&](size_t c) {
        size_t lo = n * c / chunks, hi = n * (c + 1) / chunks;
        if (lo < hi) partial[c] = body(blocked_range<It>(b + lo, b + hi), identity);
    });
    T acc = identity;
    for (size_t c = 0; c < chunks; ++c) acc = join(acc, partial[c]);
    return acc;
}

template <typename T> using cache_aligned_allocator = std::allocator<T>;
enum ets_key_usage_type { ets_key_per_instance, ets_no_key };

/// Per-thread instance, destroyed when the owning thread exits (thread_local map keyed by `this`).
template <typename T, typename A = std::allocator<T>, ets_key_usage_type K = ets_no_key>
class enumerable_thread_specific {
public:
    T& local() {
        thread_local std::unordered_map<const void*, std::unique_ptr<T>> tls;
        auto& p = tls[this];
        if (!p) p.reset(new T());
        return *p;
    }
};

class task_group {
public:
    ~task_group() { wait(); }
    template <typename F> void run(F f) {
        pending_.fetch_add(1);
        std::thread([this, f]() mutable { f(); pending_.fetch_sub(1); }).detach();
    }
    void wait() { while (pending_.load() > 0) std::this_thread::yield(); }
private:
    std::atomic<int> pending_{0};
};

template <typename It> void parallel_sort(It b, It e) { std::sort(b, e); }
template <typename It, typename C> void parallel_sort(It b, It e, C c) { std::sort(b, e, c); }

} // namespace tbb
#endif
