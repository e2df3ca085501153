// stand-in for tbb::parallel_for over a blocked_range.  Default (pinning): SamplingIntegrator::render's tile loop runs
// serially as ONE task, so the sampler is cloned once and its sequence runs through all blocks in spiral order -- the
// oracle's reference-seeding mode mirrors that.  With MSK_REF_TBB_THREADS=n > 1 (the timing arm of bench.py only) the range
// is cut into grain-sized tasks handed to n std::threads, as TBB hands them to its workers: every task clones the sampler
// again (integrator.cpp:57), the image then depends on scheduling exactly as the reference's does.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdlib>
#include <thread>
#include <vector>
namespace tbb {
template <typename T> struct blocked_range {
    T b, e;
    size_t grain;
    blocked_range(T b_, T e_, size_t g = 1) : b(b_), e(e_), grain(g ? g : 1) {}
    T begin() const { return b; }
    T end() const { return e; }
};
template <typename R, typename F> void parallel_for(const R &r, const F &f) {
    const char *env = std::getenv("MSK_REF_TBB_THREADS");
    const int nthreads = env ? std::atoi(env) : 1;
    if (nthreads <= 1) { f(r); return; }
    std::atomic<size_t> next((size_t) r.begin());
    const size_t end = (size_t) r.end(), grain = r.grain;
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                const size_t b = next.fetch_add(grain);
                if (b >= end) return;
                f(R(b, b + grain < end ? b + grain : end, grain));
            }
        });
    for (std::thread &t : pool) t.join();
}
} // namespace tbb
