// Number of kernels this library has launched (bench.py reports it as gpu_launches).
#pragma once
#include <atomic>
#include <cstdint>
namespace lm {
extern std::atomic<uint64_t> g_kernel_launches;
inline void count_launch(uint64_t n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace lm
