// Process-wide count of the kernels this library has launched (every `<<<>>>` site calls
// note_launch()).  bench.py reads it through mscnn_kernel_launch_count() to report `gpu_launches`
// from a count, not an estimate.
#pragma once
#include <atomic>

namespace mscnn {
extern std::atomic<unsigned long long> g_kernel_launches;
inline void note_launch() { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace mscnn
