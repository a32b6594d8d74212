// Microbenchmark: rate of scattered 32-bit RED.XOR (fire-and-forget atomics) into a large bitset, per CTA of 256 threads.
// Informs the Rips reducer design (coboundary toggles).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 red_throughput.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return (uint32_t)x; }
__global__ void k(uint32_t* bits, uint64_t words_per_cta, int iters, int with_smem_flag, long long* cyc) {
  __shared__ uint32_t s1[8192];
  uint32_t* b = bits + (size_t)blockIdx.x * words_per_cta;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s1[i] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 8
    for (int j = 0; j < 8; ++j) {
      uint64_t r = (uint64_t)mix(((uint64_t)blockIdx.x << 40) + ((uint64_t)it << 16) + (j << 10) + threadIdx.x) * 2654435761ull;
      uint64_t bit = r % (words_per_cta * 32);
      atomicXor(&b[bit >> 5], 1u << (bit & 31));
      if (with_smem_flag) { uint32_t pg = (uint32_t)(bit >> 14); uint32_t m = 1u << (pg & 31); if (!(s1[pg >> 5] & m)) atomicOr(&s1[pg >> 5], m); }
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}
int main() {
  const uint64_t words = (1ull << 32) / 32;  // 512 MB per CTA window
  int ctas_list[] = {1, 32, 148};
  uint32_t* bits; long long* cyc;
  cudaMalloc(&bits, 148ull * words * 4); cudaMemset(bits, 0, 148ull * words * 4);
  cudaMalloc(&cyc, 148 * 8);
  for (uint64_t wpc : {words, words / 64}) for (int flag : {0, 1}) for (int ctas : ctas_list) {
    const int iters = 2000;
    k<<<ctas, 256>>>(bits, wpc, 10, flag, cyc); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<ctas, 256>>>(bits, wpc, iters, flag, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
    double toggles = (double)iters * 8 * 256;
    printf("window %4llu MB flag %d ctas %3d: %.3f ms  %.2f cycles/toggle/CTA  %.2f Gtoggle/s total\n", (unsigned long long)(wpc * 4 >> 20), flag, ctas, ms,
           (double)h[0] / toggles, toggles * ctas / ms / 1e6);
  }
  return 0;
}
