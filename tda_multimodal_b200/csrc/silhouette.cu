// silhouette.cu -- mean silhouette coefficient of labelled 3-D clouds from their distance matrices.
//
// Replaces sklearn.metrics.silhouette_score(point_cloud_low_dim, labels) (debug_tda_pipeline.py:117-118,
// analyze_adversarial_tda.py:108-111): per point i, a = mean distance to the other points of its own label,
// b = smallest mean distance to the points of another label, s = (b - a) / max(a, b) (0 for a singleton label);
// the score is the mean of s.  The distance matrix is the one the Rips stage already built (tda_pdist_lowdim).
// One warp per point streams its row once (coalesced), per-label sums live in shared memory.
#include "common.cuh"
#include "launch_count.cuh"
#include "../../include/tda_b200.h"

namespace tda {
namespace silhouette {

constexpr int kWarps = 8;
constexpr int kMaxLabels = 64;

__global__ void __launch_bounds__(kWarps * 32) silhouette_kernel(const float* __restrict__ dm, const int* __restrict__ labels, int n, int n_labels,
                                                                  double* __restrict__ acc_out) {
  __shared__ float s_sum[kWarps][kMaxLabels];
  __shared__ int s_cnt[kMaxLabels];
  __shared__ double s_part[kWarps];
  const int p = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarps + warp;
  const int* lab = labels + (size_t)p * n;
  for (int l = threadIdx.x; l < n_labels; l += blockDim.x) s_cnt[l] = 0;
  for (int l = lane; l < n_labels; l += 32) s_sum[warp][l] = 0.f;
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) atomicAdd(&s_cnt[lab[j]], 1);
  __syncthreads();
  double s = 0.0;
  if (i < n) {
    const float* row = dm + ((size_t)p * n + i) * n;
    for (int j = lane; j < n; j += 32) atomicAdd(&s_sum[warp][lab[j]], row[j]);
    __syncwarp();
    const int own = lab[i];
    float b = INFINITY;
    for (int l = lane; l < n_labels; l += 32)
      if (l != own && s_cnt[l] > 0) b = fminf(b, s_sum[warp][l] / (float)s_cnt[l]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
    if (lane == 0 && s_cnt[own] > 1 && isfinite(b)) {
      const float a = s_sum[warp][own] / (float)(s_cnt[own] - 1);
      const float m = fmaxf(a, b);
      s = m > 0.f ? (double)((b - a) / m) : 0.0;
    }
  }
  if (lane == 0) s_part[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kWarps; ++w) t += s_part[w];
    atomicAdd(&acc_out[p], t);
  }
}

__global__ void finish_kernel(const double* __restrict__ acc, int n, int batch, float* __restrict__ score) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < batch) score[p] = (float)(acc[p] / (double)n);
}

}  // namespace silhouette
}  // namespace tda

using namespace tda;

extern "C" int tda_silhouette(const float* dm, const int32_t* labels, int n, int batch, int n_labels, float* score, void* ws, size_t ws_bytes,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!dm || !labels || !score || !ws || n <= 1 || batch <= 0) return set_error(TDA_ERR_INVALID, "tda_silhouette: bad arguments");
  if (n_labels < 2 || n_labels > silhouette::kMaxLabels)
    return set_error(TDA_ERR_UNSUPPORTED, "tda_silhouette: n_labels=%d (supported: 2..%d)", n_labels, silhouette::kMaxLabels);
  if (ws_bytes < sizeof(double) * (size_t)batch) return set_error(TDA_ERR_WORKSPACE, "tda_silhouette: workspace too small");
  double* acc = (double*)ws;
  TDA_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double) * batch, stream));
  dim3 grid((n + silhouette::kWarps - 1) / silhouette::kWarps, batch);
  silhouette::silhouette_kernel<<<grid, silhouette::kWarps * 32, 0, stream>>>(dm, labels, n, n_labels, acc);
  silhouette::finish_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(acc, n, batch, score);
  count_launch(2);
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}
