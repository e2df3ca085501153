// stand-in: the pinned build runs SamplingIntegrator::render's tile loop serially, as ONE task (so the sampler is cloned
// once and its sequence runs through all blocks in spiral order -- the oracle's reference-seeding mode mirrors that)
#pragma once
#include <cstddef>
namespace tbb {
template <typename T> struct blocked_range {
    T b, e;
    blocked_range(T b_, T e_, size_t = 1) : b(b_), e(e_) {}
    T begin() const { return b; }
    T end() const { return e; }
};
template <typename R, typename F> void parallel_for(const R &r, const F &f) { f(r); }
} // namespace tbb
