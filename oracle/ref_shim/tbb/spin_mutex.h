// stand-in: imageblock.h keeps a tbb::spin_mutex for BlockGenerator::next_block; the pinned build is single-threaded
#pragma once
namespace tbb { class spin_mutex { public: void lock() {} void unlock() {} bool try_lock() { return true; } class scoped_lock { public: explicit scoped_lock(spin_mutex &) {} }; }; }
