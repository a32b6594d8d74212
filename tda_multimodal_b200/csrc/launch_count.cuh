// Per-thread count of kernels this library launched (bench.py reports it as gpu_launches).
#pragma once
#include <cstdint>
namespace tda {
int64_t& launch_counter();  // defined in capi.cu
inline void count_launch(int k = 1) { launch_counter() += k; }
}  // namespace tda
