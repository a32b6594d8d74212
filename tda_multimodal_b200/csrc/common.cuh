// Shared helpers for the tda_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>

namespace tda {

// ---- error reporting: every extern "C" entry returns 0 or a negative code; text via tda_last_error()
enum : int {
  TDA_OK = 0,
  TDA_ERR_INVALID = -1,    // bad argument
  TDA_ERR_CUDA = -2,       // CUDA runtime error
  TDA_ERR_WORKSPACE = -3,  // workspace too small
  TDA_ERR_CAPACITY = -4,   // an output buffer / internal pool overflowed (result incomplete)
  TDA_ERR_UNSUPPORTED = -5
};

char* tls_error_buffer();  // defined in capi.cu
inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define TDA_CUDA_CHECK(expr)                                                                 \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return tda::set_error(tda::TDA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define TDA_LAUNCH_CHECK() TDA_CUDA_CHECK(cudaGetLastError())

// ---- workspace carving (256-byte aligned bump allocator over caller memory)
struct Carver {
  char* base;
  size_t off;
  size_t cap;
  __host__ Carver(void* b, size_t c) : base((char*)b), off(0), cap(c) {}
  template <typename T>
  __host__ T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* p = base ? (T*)(base + off) : nullptr;
    off += bytes;
    return p;
  }
  __host__ bool ok() const { return off <= cap; }
};

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers
__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// edge (i>j) <-> combinatorial index i(i-1)/2 + j
__host__ __device__ __forceinline__ int64_t edge_index(int i, int j) {
  return (int64_t)i * (i - 1) / 2 + j;
}
__device__ __forceinline__ void edge_vertices(int64_t idx, int& i, int& j) {
  int ii = (int)((1.0 + sqrt(1.0 + 8.0 * (double)idx)) * 0.5);
  while ((int64_t)ii * (ii - 1) / 2 > idx) --ii;
  while ((int64_t)(ii + 1) * ii / 2 <= idx) ++ii;
  i = ii;
  j = (int)(idx - (int64_t)ii * (ii - 1) / 2);
}

}  // namespace tda
