// stand-in: imageblock.h keeps a tbb::spin_mutex for BlockGenerator::next_block (a real lock: the timing arm of bench.py runs
// the tile loop on several threads, see parallel_for.h)
#pragma once
#include <atomic>
namespace tbb {
class spin_mutex {
public:
    void lock() { while (m_flag.test_and_set(std::memory_order_acquire)) {} }
    void unlock() { m_flag.clear(std::memory_order_release); }
    bool try_lock() { return !m_flag.test_and_set(std::memory_order_acquire); }
    class scoped_lock {
    public:
        explicit scoped_lock(spin_mutex &m) : m_mutex(m) { m_mutex.lock(); }
        ~scoped_lock() { m_mutex.unlock(); }
    private:
        spin_mutex &m_mutex;
    };
private:
    std::atomic_flag m_flag = ATOMIC_FLAG_INIT;
};
} // namespace tbb
