// Microbenchmark 2: (a) scattered RED.XOR rate per CTA vs block size; (b) cost of a single-lane RED / ST / ATOM-with-return issued
// back to back by one thread (what a sequential "resolver" lane pays per global update).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return (uint32_t)x; }
__global__ void scatter(uint32_t* bits, uint32_t words, int iters, long long* cyc) {
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t r = mix(((uint64_t)blockIdx.x << 40) + ((uint64_t)it << 20) + (j << 12) + threadIdx.x);
      atomicXor(&bits[r & (words - 1)], 1u << (r >> 27));
    }
  }
  __threadfence(); __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}
__global__ void single(uint32_t* bits, uint32_t words, int iters, int mode, long long* cyc, uint32_t* sink) {
  if (threadIdx.x != 0) return;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r = mix(it * 7919ull + 13);
    uint32_t* p = &bits[r & (words - 1)];
    if (mode == 0) atomicXor(p, 1u << (r >> 27));            // RED
    else if (mode == 1) *p = r;                               // plain store
    else if (mode == 2) acc += atomicXor(p, 1u << (r >> 27)); // ATOM with return (dependent)
    else acc += __ldcg(p);                                    // dependent load
  }
  long long t1 = clock64();
  cyc[0] = t1 - t0; *sink = acc;
}
int main() {
  const uint32_t words = 1u << 27;  // 512 MB
  uint32_t* bits; long long* cyc; uint32_t* sink;
  cudaMalloc(&bits, (size_t)words * 4); cudaMemset(bits, 0, (size_t)words * 4);
  cudaMalloc(&cyc, 1024 * 8); cudaMalloc(&sink, 4);
  for (int bs : {256, 512, 1024}) for (int ctas : {1, 32}) {
    const int iters = 1000;
    scatter<<<ctas, bs>>>(bits, words, 10, cyc); cudaDeviceSynchronize();
    scatter<<<ctas, bs>>>(bits, words, iters, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("scatter block %4d ctas %2d: %.2f cycles/toggle/CTA\n", bs, ctas, (double)h / ((double)iters * 8 * bs));
  }
  const char* names[] = {"RED.XOR (no return)", "plain store", "ATOM.XOR with return", "dependent ld.cg"};
  for (int mode = 0; mode < 4; ++mode) {
    single<<<1, 32>>>(bits, words, 100, mode, cyc, sink); cudaDeviceSynchronize();
    single<<<1, 32>>>(bits, words, 2000, mode, cyc, sink); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("single lane %-22s: %.1f cycles each\n", names[mode], (double)h / 2000);
  }
  return 0;
}
