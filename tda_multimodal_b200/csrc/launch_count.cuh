// Per-thread count of kernels this library launched (bench.py reports it as gpu_launches) and the optional
// stage timer (CUDA events recorded on the launching stream around each stage; bench.py reads them for the roofline).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
namespace tda {
int64_t& launch_counter();  // defined in capi.cu
inline void count_launch(int k = 1) { launch_counter() += k; }
long long option(const char* name);  // tuning options set through tda_set_option (capi.cu)

// stage ids (keep in sync with include/tda_b200.h TDA_STAGE_*)
enum Stage : int {
  STAGE_PDIST_PREP = 0, STAGE_PDIST_GEMM, STAGE_KNN_SMOOTH, STAGE_FUZZY, STAGE_SPECTRAL, STAGE_SGD, STAGE_RIPS_PDIST,
  STAGE_RIPS_SORT, STAGE_RIPS_H0, STAGE_RIPS_APPARENT, STAGE_RIPS_REDUCE, STAGE_COUNT
};
bool stage_timing_on();
void stage_begin(int stage, cudaStream_t s);
void stage_end(int stage, cudaStream_t s);
inline void stage_begin_if(int stage, cudaStream_t s) { if (stage_timing_on()) stage_begin(stage, s); }
inline void stage_end_if(int stage, cudaStream_t s) { if (stage_timing_on()) stage_end(stage, s); }
struct StageScope {
  int stage; cudaStream_t s; bool on;
  StageScope(int st, cudaStream_t str) : stage(st), s(str), on(stage_timing_on()) { if (on) stage_begin(stage, s); }
  ~StageScope() { if (on) stage_end(stage, s); }
};
}  // namespace tda
