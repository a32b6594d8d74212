// landmarks.cu -- furthest-point ("greedy permutation") landmark selection, ripser.py's `n_perm` semantics.
//
// Replaces ripser.py's getGreedyPerm as reached from ripser(X, n_perm=...) (not used by the reference's three scripts, which call
// ripser(Y, maxdim=1) on at most a few hundred points, but the way BASELINE.json config C5 subsamples a 100k-point cloud to 10k
// landmarks): start at point 0; repeatedly take the point furthest from the chosen ones (lowest index on ties), lambda = its
// distance; ds = min(ds, distances to the new landmark).
//
// One thread-block cluster does the whole selection in ONE launch: the points are split over the CTAs, each CTA keeps the running
// minimum distances of its points in shared memory (or in the workspace when they do not fit), finds its furthest point, and the
// candidates of all CTAs are exchanged through distributed shared memory -- one cluster barrier per landmark instead of ~5 kernel
// launches per landmark.
#include "common.cuh"
#include "launch_count.cuh"
#include "../../include/tda_b200.h"
#include <cstring>

namespace tda {
namespace landmarks {

constexpr int kThreads = 1024;
constexpr int kMaxCluster = 8;
constexpr int kMaxDim = 16;

__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
template <typename T>
__device__ __forceinline__ T* cl_map(T* p, uint32_t rank) {
  uint64_t out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((uint64_t)p), "r"(rank));
  return reinterpret_cast<T*>(out);
}
// (distance, index) candidates: larger distance wins, lower index on ties; packed so that an integer max does it
__device__ __forceinline__ unsigned long long pack(float d, int i) { return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i); }

// grid = C CTAs (one cluster).  X: points [n,d] (is_matrix = 0; euclidean, float64 accumulation like ripser.py's front end) or a
// distance matrix [n,n] (is_matrix = 1).  ds: [n] scratch when ds_in_smem = 0.
__global__ void __launch_bounds__(kThreads, 1) greedy_perm_kernel(const float* __restrict__ X, int n, int d, int n_perm, int is_matrix,
                                                                  int* __restrict__ idx_out, float* __restrict__ lambda_out, float* __restrict__ ds_g,
                                                                  int ds_in_smem, int per_cta) {
  extern __shared__ float s_ds[];
  __shared__ unsigned long long s_cand[2][kMaxCluster];
  __shared__ unsigned long long s_warp[kThreads / 32];
  const uint32_t C = cl_size(), cr = cl_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = min(n, (int)cr * per_cta), i1 = min(n, i0 + per_cta);
  float* ds = ds_in_smem ? s_ds : ds_g + i0;
  for (int i = i0 + tid; i < i1; i += kThreads) ds[i - i0] = INFINITY;
  __syncthreads();
  cl_sync();
  int cur = 0;
  for (int it = 0; it < n_perm; ++it) {
    // distances to the current landmark; running minimum; this CTA's furthest point
    double pc[kMaxDim];
    if (!is_matrix)
      for (int a = 0; a < d; ++a) pc[a] = (double)__ldg(&X[(size_t)cur * d + a]);
    unsigned long long best = 0ull;
    for (int i = i0 + tid; i < i1; i += kThreads) {
      float dist;
      if (is_matrix) dist = __ldg(&X[(size_t)cur * n + i]);
      else {
        double acc = 0.0;
        for (int a = 0; a < d; ++a) { const double t = (double)__ldg(&X[(size_t)i * d + a]) - pc[a]; acc += t * t; }
        dist = (float)sqrt(acc);
      }
      const float m = fminf(ds[i - i0], dist);
      ds[i - i0] = m;
      const unsigned long long c = pack(m, i);
      if (c > best) best = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o); if (t > best) best = t; }
    if (lane == 0) s_warp[warp] = best;
    __syncthreads();
    if (warp == 0) {
      unsigned long long b2 = lane < kThreads / 32 ? s_warp[lane] : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, b2, o); if (t > b2) b2 = t; }
      if (lane < (int)C) *(cl_map(&s_cand[it & 1][0], (uint32_t)lane) + cr) = b2;   // lane r writes this CTA's candidate into CTA r's table
    }
    cl_sync();
    unsigned long long g = 0ull;
    for (uint32_t r = 0; r < C; ++r) { const unsigned long long t = s_cand[it & 1][r]; if (t > g) g = t; }
    const float lam = __uint_as_float((uint32_t)(g >> 32));
    const int nxt = (int)(0xffffffffu - (uint32_t)(g & 0xffffffffu));
    // lambda_it = the furthest remaining distance after landmark `it` joined: what ripser.py stores (the last one is r_cover)
    if (cr == 0 && tid == 0) { idx_out[it] = cur; lambda_out[it] = lam; }
    cur = nxt;
    __syncthreads();
  }
  cl_sync();
}

}  // namespace landmarks
}  // namespace tda

using namespace tda;
using namespace tda::landmarks;

extern "C" size_t tda_greedy_perm_workspace_bytes(int n) { return n > 0 ? sizeof(float) * (size_t)n + 256 : 0; }

extern "C" int tda_greedy_perm(const float* X, int n, int d, int n_perm, int is_matrix, int32_t* idx_out, float* lambda_out, void* ws,
                               size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!X || !idx_out || !lambda_out || n <= 0 || n_perm <= 0 || n_perm > n) return set_error(TDA_ERR_INVALID, "tda_greedy_perm: bad arguments");
  if (!is_matrix && (d < 1 || d > kMaxDim)) return set_error(TDA_ERR_UNSUPPORTED, "tda_greedy_perm: d=%d (points: 1..%d; pass a distance matrix otherwise)", d, kMaxDim);
  int C = n >= 65536 ? 8 : n >= 16384 ? 4 : n >= 4096 ? 2 : 1;
  const int per = (n + C - 1) / C;
  const size_t smem = sizeof(float) * (size_t)per;
  const int in_smem = smem <= (size_t)200 * 1024;
  if (!in_smem && (!ws || ws_bytes < sizeof(float) * (size_t)n)) return set_error(TDA_ERR_WORKSPACE, "tda_greedy_perm: workspace too small");
  const size_t dyn = in_smem ? smem : 0;
  TDA_CUDA_CHECK(cudaFuncSetAttribute(greedy_perm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(dyn > 48 * 1024 ? dyn : 48 * 1024)));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)C, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TDA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, greedy_perm_kernel, X, n, d, n_perm, is_matrix, (int*)idx_out, lambda_out, (float*)ws, in_smem, per));
  count_launch();
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}
