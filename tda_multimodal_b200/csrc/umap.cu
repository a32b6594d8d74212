// umap.cu -- the UMAP stages between the distance matrix and the 3-D embedding, batched over clouds.
//
// Replaces, inside umap.UMAP(...).fit_transform / .fit / .transform (debug_tda_pipeline.py:96-104,
// analyze_tda_over_layers.py:38-44,69,72, analyze_adversarial_tda.py:85-93), umap-learn's
//   fast_knn_indices + smooth_knn_dist      -> knn_smooth_kernel   (one warp per row: top-k + sigma/rho bisection)
//   compute_membership_strengths + fuzzy union (P + P^T - P o P^T) + make_epochs_per_sample -> fuzzy_kernel
//   optimize_layout_euclidean               -> sgd_epoch_kernel    (edge-parallel, on-device negative sampling)
//   noisy_scale_coords / min-max rescale, init_transform          -> small helpers
// SURVEY.md Appendix A is the specification these kernels follow; oracle/umap_oracle.py restates it on the CPU.
#include "common.cuh"
#include "launch_count.cuh"
#include "../../include/tda_b200.h"
#include <cmath>
#include <cstdlib>

namespace tda {
namespace umap {

constexpr double kSmoothKTolerance = 1e-5;
constexpr float kMinKDistScale = 1e-3f;

__device__ __forceinline__ bool pair_less(float d1, int i1, float d2, int i2) { return d1 < d2 || (d1 == d2 && i1 < i2); }

// ------------------------------------------------------------------------------------------------
// exact kNN (self included, as umap-learn's precomputed-metric path does) fused with smooth_knn_dist.
// The running top-k of a row is a sorted list striped over a warp (rank r -> lane r%32, slot r/32); ties go to the
// smaller index (numpy's stable argsort), so the result does not depend on the order candidates are offered in.
template <int KPL>
struct TopK {
  float dv[KPL];
  int iv[KPL];
  float thr_d;   // current k-th smallest (d, index): only candidates below it can enter
  int thr_i;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < KPL; ++s) { dv[s] = INFINITY; iv[s] = 0x7fffffff; }
    thr_d = INFINITY;
    thr_i = 0x7fffffff;
  }
  // every lane offers one candidate (d, j); `ok` false = nothing to offer
  __device__ __forceinline__ void offer(float d, int j, bool ok, int lane, int last_lane, int last_slot) {
    unsigned cand = __ballot_sync(0xffffffffu, ok && pair_less(d, j, thr_d, thr_i));
    while (cand) {
      const int src = __ffs(cand) - 1;
      cand &= cand - 1;
      const float cd = __shfl_sync(0xffffffffu, d, src);
      const int cj = __shfl_sync(0xffffffffu, j, src);
      if (!pair_less(cd, cj, thr_d, thr_i)) continue;  // threshold moved since the ballot
      // insertion position = number of list entries smaller than the candidate
      int pos = 0;
#pragma unroll
      for (int s = 0; s < KPL; ++s) pos += __popc(__ballot_sync(0xffffffffu, pair_less(dv[s], iv[s], cd, cj)));
#pragma unroll
      for (int s = KPL - 1; s >= 0; --s) {
        const int r = s * 32 + lane;
        float upd = __shfl_up_sync(0xffffffffu, dv[s], 1);
        int upi = __shfl_up_sync(0xffffffffu, iv[s], 1);
        if (s > 0) {
          const float wd = __shfl_sync(0xffffffffu, dv[s - 1], 31);
          const int wi = __shfl_sync(0xffffffffu, iv[s - 1], 31);
          if (lane == 0) { upd = wd; upi = wi; }
        }
        if (r == pos) { dv[s] = cd; iv[s] = cj; }
        else if (r > pos) { dv[s] = upd; iv[s] = upi; }
      }
      float td = dv[0];
      int ti = iv[0];
#pragma unroll
      for (int s = 1; s < KPL; ++s)
        if (s == last_slot) { td = dv[s]; ti = iv[s]; }
      thr_d = __shfl_sync(0xffffffffu, td, last_lane);
      thr_i = __shfl_sync(0xffffffffu, ti, last_lane);
    }
  }
};

// Streams ROWS rows of D through their top-k lists.  Every lane keeps kKnnVecLoads 16-byte loads per row in flight
// (2 KB per row per warp and iteration; streaming, evict-first) and the warp looks at a batch only when some lane holds a
// value not above the row's current threshold -- after the first few hundred columns almost never.
constexpr int kKnnVecLoads = 4;
constexpr int kKnnScalarLoads = 8;
template <int KPL, int ROWS>
__device__ __forceinline__ void knn_stream_rows(const float* const (&drow)[ROWS], const bool (&live)[ROWS], int m, int k, int lane,
                                                TopK<KPL> (&tk)[ROWS]) {
  const int last_lane = (k - 1) & 31, last_slot = (k - 1) >> 5;
  bool vec = (m & 3) == 0;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) vec = vec && ((reinterpret_cast<uintptr_t>(drow[r]) & 15) == 0);
  if (vec) {
    const int m4 = m >> 2;
    for (int q0 = 0; q0 < m4; q0 += 32 * kKnnVecLoads) {
      float4 v[ROWS][kKnnVecLoads];
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int u = 0; u < kKnnVecLoads; ++u) {
          const int q = q0 + u * 32 + lane;
          v[r][u] = (live[r] && q < m4) ? __ldcs(reinterpret_cast<const float4*>(drow[r]) + q) : make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
        }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        float lmin = INFINITY;
#pragma unroll
        for (int u = 0; u < kKnnVecLoads; ++u) lmin = fminf(lmin, fminf(fminf(v[r][u].x, v[r][u].y), fminf(v[r][u].z, v[r][u].w)));
        if (live[r] && __any_sync(0xffffffffu, lmin <= tk[r].thr_d)) {
#pragma unroll
          for (int u = 0; u < kKnnVecLoads; ++u) {
            const int q = q0 + u * 32 + lane;
            const bool ok = q < m4;
            tk[r].offer(v[r][u].x, 4 * q + 0, ok, lane, last_lane, last_slot);
            tk[r].offer(v[r][u].y, 4 * q + 1, ok, lane, last_lane, last_slot);
            tk[r].offer(v[r][u].z, 4 * q + 2, ok, lane, last_lane, last_slot);
            tk[r].offer(v[r][u].w, 4 * q + 3, ok, lane, last_lane, last_slot);
          }
        }
      }
    }
  } else {
    for (int j0 = 0; j0 < m; j0 += 32 * kKnnScalarLoads) {
      float v[ROWS][kKnnScalarLoads];
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int u = 0; u < kKnnScalarLoads; ++u) {
          const int j = j0 + u * 32 + lane;
          v[r][u] = (live[r] && j < m) ? __ldcs(drow[r] + j) : INFINITY;
        }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        float lmin = INFINITY;
#pragma unroll
        for (int u = 0; u < kKnnScalarLoads; ++u) lmin = fminf(lmin, v[r][u]);
        if (live[r] && __any_sync(0xffffffffu, lmin <= tk[r].thr_d)) {
#pragma unroll
          for (int u = 0; u < kKnnScalarLoads; ++u) {
            const int j = j0 + u * 32 + lane;
            tk[r].offer(v[r][u], j, j < m, lane, last_lane, last_slot);
          }
        }
      }
    }
  }
}

// ---- k <= 16: the C3 / C5 case ----------------------------------------------------------------------------------
// A CTA of 4 warps owns 32 rows.  SELECTION (warp per row, 8 rows one after the other): the row is read once, in chunks of
// 2048 columns held in registers (16 x 16-byte streaming loads per lane in flight).  The k-th smallest of the 32 lane minima
// of a chunk bounds the row's k-th smallest from above, so only the few columns not above that bound (about 25 of 2000)
// are candidates; they are pushed into a small shared-memory buffer and merged into the sorted best-k list (one 64-bit key
// (distance, column) per lane) by a bitonic sort across the warp -- about a fifth of the instructions of an insertion per
// candidate.  A row whose buffer overflows (ties, infinite distances) is redone with the insertion list.
// BISECTION (all 128 threads, 4 threads per row): every step is evaluated in fp32 first (MUFU ex2); fp32 decides a step
// only when it is at least kScreenBand away from the target (its error is < 1e-5: 15 terms of |x| e^-|x| * 2^-22), a row
// that comes closer is parked, and the parked rows finish with the numba kernel's fp64 sum from exactly that state, so
// lo/mid/hi follow the fp64 sequence.
constexpr float kScreenBand = 1e-4f;
constexpr int kBlkWarps = 4, kBlkRows = 32, kBlkRowsPerWarp = kBlkRows / kBlkWarps;
constexpr int kBlkCand = 64;                        // candidate buffer per warp
constexpr unsigned long long kKeyMax = ~0ull;

__device__ __forceinline__ unsigned long long knn_key(float d, int j) {
  d += 0.0f;                                        // -0 -> +0: equal distances must tie (the smaller column wins)
  const uint32_t b = __float_as_uint(d);
  const uint32_t o = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)o << 32) | (uint32_t)j;
}
__device__ __forceinline__ float knn_key_dist(unsigned long long key) {
  const uint32_t o = (uint32_t)(key >> 32);
  return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}
template <typename T>
__device__ __forceinline__ T warp_sort_asc(T v, int lane) {   // bitonic sort of one value per lane, ascending with the lane
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      const T o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool keep_min = ((lane & j) == 0) == ((lane & k2) == 0);
      const T mn = v < o ? v : o, mx = v < o ? o : v;
      v = keep_min ? mn : mx;
    }
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_merge_bitonic(T v, int lane) {   // v bitonic across the lanes -> ascending
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const T o = __shfl_xor_sync(0xffffffffu, v, j);
    const T mn = v < o ? v : o, mx = v < o ? o : v;
    v = ((lane & j) == 0) ? mn : mx;
  }
  return v;
}

template <int kBlkLoads>                            // 16-byte loads per lane and chunk: a chunk is 128 * kBlkLoads columns
__global__ void __launch_bounds__(kBlkWarps * 32) knn_smooth_block_kernel(const float* __restrict__ D, int n, int m, int k, float local_connectivity,
                                                                            float bandwidth, int n_iter, int* __restrict__ knn_idx,
                                                                            float* __restrict__ knn_dist, float* __restrict__ sigma,
                                                                            float* __restrict__ rho, double* __restrict__ dist_sum) {
  constexpr int kBlkChunk = 32 * 4 * kBlkLoads;
  __shared__ unsigned long long s_cand[kBlkWarps][kBlkCand];
  __shared__ uint32_t s_cnt[kBlkWarps];
  __shared__ float s_kd[kBlkRows][16];
  __shared__ float s_rho[kBlkRows];
  __shared__ double s_rsum[kBlkRows];
  const int p = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_base = blockIdx.x * kBlkRows;
  const int index = (int)floorf(local_connectivity);
  const float interp = local_connectivity - (float)index;

  for (int ri = 0; ri < kBlkRowsPerWarp; ++ri) {
    const int rl = warp * kBlkRowsPerWarp + ri;
    const int row = row_base + rl;
    if (row >= n) break;                               // (warp uniform)
    const float* drow = D + ((size_t)p * n + row) * m;
    const bool vec = ((m & 3) == 0) && ((reinterpret_cast<uintptr_t>(drow) & 15) == 0);
    unsigned long long best = kKeyMax;                 // lane r: r-th smallest (distance, column) so far; kKeyMax beyond k
    float tau = INFINITY;
    bool overflow = false;
    if (lane == 0) s_cnt[warp] = 0;
    __syncwarp();
    for (int c0 = 0; c0 < m && !overflow; c0 += kBlkChunk) {
      float4 v[kBlkLoads];
      const float kPad = __int_as_float(0x7fc00000);   // columns past the end are NaN: fminf skips them and "<= tau" is false
      if (vec) {
#pragma unroll
        for (int u = 0; u < kBlkLoads; ++u) {
          const int e = c0 + (u * 32 + lane) * 4;
          v[u] = e < m ? __ldcs(reinterpret_cast<const float4*>(drow + e)) : make_float4(kPad, kPad, kPad, kPad);
        }
      } else {
#pragma unroll
        for (int u = 0; u < kBlkLoads; ++u) {
          const int e = c0 + (u * 32 + lane) * 4;
          v[u].x = e + 0 < m ? __ldcs(drow + e + 0) : kPad;
          v[u].y = e + 1 < m ? __ldcs(drow + e + 1) : kPad;
          v[u].z = e + 2 < m ? __ldcs(drow + e + 2) : kPad;
          v[u].w = e + 3 < m ? __ldcs(drow + e + 3) : kPad;
        }
      }
      float lmin = INFINITY;
#pragma unroll
      for (int u = 0; u < kBlkLoads; ++u) lmin = fminf(lmin, fminf(fminf(v[u].x, v[u].y), fminf(v[u].z, v[u].w)));
      const float sorted_min = warp_sort_asc<float>(lmin, lane);
      tau = fminf(tau, __shfl_sync(0xffffffffu, sorted_min, k - 1));
      // candidates: everything not above the bound (ties included; the merge orders them by column)
#pragma unroll
      for (int u = 0; u < kBlkLoads; ++u) {
        const int e = c0 + (u * 32 + lane) * 4;
        const float dd[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (dd[c] <= tau) {
            const uint32_t pos = atomicAdd(&s_cnt[warp], 1u);
            if (pos < (uint32_t)kBlkCand) s_cand[warp][pos] = knn_key(dd[c], e + c);
          }
        }
      }
      __syncwarp();
      const uint32_t cnt = s_cnt[warp];
      if (cnt > (uint32_t)kBlkCand) { overflow = true; break; }
      const bool last = c0 + kBlkChunk >= m;
      if (cnt >= 32u || last) {
        unsigned long long a = (uint32_t)lane < cnt ? s_cand[warp][lane] : kKeyMax;
        a = warp_sort_asc<unsigned long long>(a, lane);
        if (cnt > 32u) {
          unsigned long long b = (uint32_t)(32 + lane) < cnt ? s_cand[warp][32 + lane] : kKeyMax;
          b = warp_sort_asc<unsigned long long>(b, lane);
          const unsigned long long br = __shfl_sync(0xffffffffu, b, 31 - lane);
          a = warp_merge_bitonic<unsigned long long>(a < br ? a : br, lane);   // the 32 smallest of a and b
        }
        // the 16 smallest candidates (lanes 0..15, ascending) against the best list reversed on lanes 16..31: bitonic
        const unsigned long long rev = __shfl_sync(0xffffffffu, best, 31 - lane);
        const unsigned long long w = warp_merge_bitonic<unsigned long long>(lane < 16 ? a : rev, lane);
        best = lane < k ? w : kKeyMax;
        const unsigned long long kth = __shfl_sync(0xffffffffu, best, k - 1);
        if (kth != kKeyMax) tau = fminf(tau, knn_key_dist(kth));
        __syncwarp();
        if (lane == 0) s_cnt[warp] = 0;
        __syncwarp();
      }
    }
    float d;
    int ix;
    if (overflow) {
      // many equal / infinite distances: the insertion list handles any input
      const float* drows[1] = {drow};
      const bool lives[1] = {true};
      TopK<1> tks[1];
      tks[0].init();
      knn_stream_rows<1, 1>(drows, lives, m, k, lane, tks);
      d = tks[0].dv[0];
      ix = tks[0].iv[0];
    } else {
      d = lane < k ? knn_key_dist(best) : INFINITY;
      ix = (int)(uint32_t)best;
    }
    const bool in_list = lane < k;
    if (in_list) {
      knn_idx[((size_t)p * n + row) * k + lane] = isinf(d) ? -1 : ix;
      knn_dist[((size_t)p * n + row) * k + lane] = d;
    }
    const double rsum = warp_sum_f64(in_list ? (double)d : 0.0);
    // rho: distance to the local_connectivity-th nearest neighbour at positive distance (interpolated)
    const int zeros = __popc(__ballot_sync(0xffffffffu, in_list && !(d > 0.f)));
    const int nnz = k - zeros;
    const float f_prev = __shfl_sync(0xffffffffu, d, min(max(zeros + index - 1, 0), 31));
    const float f_next = __shfl_sync(0xffffffffu, d, min(zeros + index, 31));
    const float f_first = __shfl_sync(0xffffffffu, d, min(zeros, 31));
    const float f_last = __shfl_sync(0xffffffffu, d, k - 1);
    float rho_i = 0.f;
    if ((float)nnz >= local_connectivity) {
      if (index > 0) {
        rho_i = f_prev;
        if (interp > (float)kSmoothKTolerance) rho_i += interp * (f_next - f_prev);
      } else {
        rho_i = interp * f_first;
      }
    } else if (nnz > 0) {
      rho_i = f_last;  // max of the positive entries = last entry of the sorted list
    }
    if (lane < 16) s_kd[rl][lane] = d;                 // (lanes >= k hold +inf)
    if (lane == 0) { s_rho[rl] = rho_i; s_rsum[rl] = rsum; }
  }
  __syncthreads();

  // ---- sigma: umap-learn's bisection, 4 threads per row (terms q+1, q+5, q+9, q+13 of the row's list)
  const int rl = threadIdx.x >> 2, q = threadIdx.x & 3;
  const int row = row_base + rl;
  const bool row_ok = row < n;
  const float rho_i = row_ok ? s_rho[rl] : 0.f;
  float dd[4];
  bool term[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int j = 1 + q + 4 * t;
    term[t] = row_ok && j < k;
    dd[t] = term[t] ? s_kd[rl][j] - rho_i : 0.f;       // float32 subtraction, as in the numba kernel
  }
  const double target = log2((double)k) * (double)bandwidth;
  const float targetf = (float)target;
  double lo = 0.0, hi = INFINITY, mid = 1.0;
  int it = 0;                                          // bisection steps this row has taken
  int state = row_ok ? 0 : 2;                          // 0: fp32 screening, 1: parked (needs fp64), 2: done
  for (int step = 0; step < n_iter; ++step) {
    if (__all_sync(0xffffffffu, state != 0)) break;
    const float midf = (float)mid;
    float e32 = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (term[t]) e32 += dd[t] > 0.f ? __expf(-__fdividef(dd[t], midf)) : 1.f;
    e32 += __shfl_xor_sync(0xffffffffu, e32, 1);
    e32 += __shfl_xor_sync(0xffffffffu, e32, 2);
    if (state == 0) {
      if (!(fabsf(e32 - targetf) >= kScreenBand)) {    // too close for fp32 (NaN counts as close): redo this step in fp64
        state = 1;
      } else {
        if (e32 > targetf) { hi = mid; mid = (lo + hi) / 2.0; }
        else { lo = mid; if (isinf(hi)) mid *= 2.0; else mid = (lo + hi) / 2.0; }
        if (++it >= n_iter) state = 2;
      }
    }
  }
  for (;;) {
    if (__all_sync(0xffffffffu, state == 2)) break;
    double e64 = 0.0;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (term[t]) e64 += dd[t] > 0.f ? exp(-((double)dd[t] / mid)) : 1.0;
    e64 += __shfl_xor_sync(0xffffffffu, e64, 1);
    e64 += __shfl_xor_sync(0xffffffffu, e64, 2);
    if (state != 2) {
      if (fabs(e64 - target) < kSmoothKTolerance) {
        state = 2;
      } else {
        if (e64 > target) { hi = mid; mid = (lo + hi) / 2.0; }
        else { lo = mid; if (isinf(hi)) mid *= 2.0; else mid = (lo + hi) / 2.0; }
        if (++it >= n_iter) state = 2;
      }
    }
  }
  double rs = 0.0;
  if (q == 0 && row_ok) {
    rs = s_rsum[rl];
    float sg = (float)mid;
    if (rho_i > 0.f) {
      const float mean_i = (float)(rs / k);
      if (sg < kMinKDistScale * mean_i) sg = kMinKDistScale * mean_i;
    }
    sigma[(size_t)p * n + row] = sg;
    rho[(size_t)p * n + row] = rho_i;
  }
  rs = warp_sum_f64(rs);
  if (lane == 0 && rs != 0.0) atomicAdd(&dist_sum[p], rs);
}

// general k (up to 256): one warp per row, list of KPL entries per lane.
template <int KPL>
__global__ void __launch_bounds__(256) knn_smooth_kernel(const float* __restrict__ D, int n, int m, int k, float local_connectivity,
                                                         float bandwidth, int n_iter, int* __restrict__ knn_idx,
                                                         float* __restrict__ knn_dist, float* __restrict__ sigma, float* __restrict__ rho,
                                                         double* __restrict__ dist_sum) {
  const int p = blockIdx.y;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* drows[1] = {D + ((size_t)p * n + row) * m};
  const bool lives[1] = {true};
  TopK<KPL> tks[1];
  tks[0].init();
  knn_stream_rows<KPL, 1>(drows, lives, m, k, lane, tks);
  float (&dv)[KPL] = tks[0].dv;
  int (&iv)[KPL] = tks[0].iv;
  // ---- write the neighbour lists (index -1 where the neighbour is at infinite distance: "disconnected")
  int* oi = knn_idx + ((size_t)p * n + row) * k;
  float* od = knn_dist + ((size_t)p * n + row) * k;
  double rsum = 0.0;
#pragma unroll
  for (int s = 0; s < KPL; ++s) {
    const int r = s * 32 + lane;
    if (r < k) {
      oi[r] = isinf(dv[s]) ? -1 : iv[s];
      od[r] = dv[s];
      rsum += (double)dv[s];
    }
  }
  rsum = warp_sum_f64(rsum);
  // ---- smooth_knn_dist for this row
  // rho: distance to the local_connectivity-th nearest neighbour at positive distance (interpolated)
  int zeros = 0;
#pragma unroll
  for (int s = 0; s < KPL; ++s) zeros += __popc(__ballot_sync(0xffffffffu, (s * 32 + lane) < k && !(dv[s] > 0.f)));
  const int nnz = k - zeros;
  auto fetch = [&](int r) -> float {  // list entry of rank r (warp uniform argument)
    float v = 0.f;
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
      const float t = __shfl_sync(0xffffffffu, dv[s], r & 31);
      if ((r >> 5) == s) v = t;
    }
    return v;
  };
  float rho_i = 0.f;
  if ((float)nnz >= local_connectivity) {
    const int index = (int)floorf(local_connectivity);
    const float interp = local_connectivity - (float)index;
    if (index > 0) {
      rho_i = fetch(zeros + index - 1);
      if (interp > (float)kSmoothKTolerance) rho_i += interp * (fetch(zeros + index) - fetch(zeros + index - 1));
    } else {
      rho_i = interp * fetch(zeros);
    }
  } else if (nnz > 0) {
    rho_i = fetch(k - 1);  // max of the positive entries = last entry of the sorted list
  }
  const double target = log2((double)k) * (double)bandwidth;
  double lo = 0.0, hi = INFINITY, mid = 1.0;
  for (int it = 0; it < n_iter; ++it) {
    double psum = 0.0;
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
      const int r = s * 32 + lane;
      if (r >= 1 && r < k) {
        const float dd = dv[s] - rho_i;  // float32 subtraction, as in the numba kernel
        psum += dd > 0.f ? exp(-((double)dd / mid)) : 1.0;
      }
    }
    psum = warp_sum_f64(psum);
    if (fabs(psum - target) < kSmoothKTolerance) break;
    if (psum > target) { hi = mid; mid = (lo + hi) / 2.0; }
    else { lo = mid; if (isinf(hi)) mid *= 2.0; else mid = (lo + hi) / 2.0; }
  }
  if (lane == 0) {
    float sg = (float)mid;
    if (rho_i > 0.f) {
      const float mean_i = (float)(rsum / k);
      if (sg < kMinKDistScale * mean_i) sg = kMinKDistScale * mean_i;
    }
    sigma[(size_t)p * n + row] = sg;
    rho[(size_t)p * n + row] = rho_i;
    atomicAdd(&dist_sum[p], rsum);
  }
}

// rows whose rho is 0 are floored with the mean over the whole [n,k] distance table
__global__ void sigma_floor_kernel(int n, int k, const float* __restrict__ rho, float* __restrict__ sigma, const double* __restrict__ dist_sum) {
  const int p = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(rho[(size_t)p * n + i] > 0.f)) {
    const float mean_all = (float)(dist_sum[p] / ((double)n * k));
    float& s = sigma[(size_t)p * n + i];
    if (s < kMinKDistScale * mean_all) s = kMinKDistScale * mean_all;
  }
}

// ------------------------------------------------------------------------------------------------
// membership strengths + fuzzy union, written to a fixed slot table: slot (i,t,0) = directed entry i->j,
// slot (i,t,1) = the transposed entry j->i when i is not in j's own list.  weight 0 = empty slot.
__device__ __forceinline__ float membership(float d, float rho_i, float sigma_i, bool is_self, bool bipartite) {
  if (!bipartite && is_self) return 0.f;
  if (d - rho_i <= 0.f || sigma_i == 0.f) return 1.f;
  return expf(-((d - rho_i) / sigma_i));
}
__global__ void fuzzy_kernel(const int* __restrict__ knn_idx, const float* __restrict__ knn_dist, const float* __restrict__ sigma,
                             const float* __restrict__ rho, int n, int k, float mix, int* __restrict__ head, int* __restrict__ tail,
                             float* __restrict__ weight, unsigned int* __restrict__ max_w_bits) {
  const int p = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  float wmax = 0.f;
  if (e < n * k) {
    const int i = e / k;
    const size_t base = (size_t)p * n;
    const int* idx = knn_idx + base * k;
    const float* dist = knn_dist + base * k;
    const int j = idx[e];
    const size_t s0 = ((size_t)p * n * k + e) * 2;
    float w = 0.f;
    bool found = false;
    if (j >= 0 && j != i) {
      const float vij = membership(dist[e], rho[base + i], sigma[base + i], false, false);
      float vji = 0.f;
      for (int t = 0; t < k; ++t)
        if (idx[(size_t)j * k + t] == i) {
          vji = membership(dist[(size_t)j * k + t], rho[base + j], sigma[base + j], false, false);
          found = true;
          break;
        }
      const float prod = vij * vji;
      w = mix * (vij + vji - prod) + (1.f - mix) * prod;
    }
    head[s0] = i; tail[s0] = j < 0 ? i : j; weight[s0] = w;
    head[s0 + 1] = j < 0 ? i : j; tail[s0 + 1] = i; weight[s0 + 1] = found ? 0.f : w;
    wmax = w;
  }
  wmax = warp_max_f32(wmax);
  if ((threadIdx.x & 31) == 0 && wmax > 0.f) atomicMax(&max_w_bits[p], __float_as_uint(wmax));
}
// epochs_per_sample = max_w / w; entries with w < max_w / n_epochs are pruned (eps = -1)
__global__ void epochs_kernel(const float* __restrict__ weight, int slots, int n_epochs, const unsigned int* __restrict__ max_w_bits,
                              float* __restrict__ eps) {
  const int p = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= slots) return;
  const float mw = __uint_as_float(max_w_bits[p]);
  const float w = weight[(size_t)p * slots + e];
  eps[(size_t)p * slots + e] = (w > 0.f && w >= mw / (float)n_epochs) ? mw / w : -1.f;
}

// ------------------------------------------------------------------------------------------------
// SGD: one launch per epoch, one thread per slot of every cloud.  Stateless schedule: an edge with period
// eps fires at the epochs ceil(q*eps), q = 1,2,...; the negative-sample budget follows umap-learn's
// epoch_of_next_negative_sample recurrence in closed form.  Updates use float atomics (every update is kept,
// order is free -- the reference's serial loop is one admissible order).
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}
__device__ __forceinline__ float clip4(float v) { return fminf(fmaxf(v, -4.f), 4.f); }

template <int DIM>
__global__ void __launch_bounds__(256) sgd_epoch_kernel(float* __restrict__ Yh, const float* __restrict__ Yt_in, const int* __restrict__ head,
                                                        const int* __restrict__ tail, const float* __restrict__ eps_arr, int slots, int n_head,
                                                        int n_tail, int epoch, float a, float b, float gamma, float alpha, float nsr,
                                                        int move_other, uint64_t seed) {
  const int p = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= slots) return;
  const float eps = eps_arr[(size_t)p * slots + e];
  if (!(eps > 0.f)) return;
  const int q = (int)floorf((float)epoch / eps);
  if (q < 1 || q <= (int)floorf((float)(epoch - 1) / eps)) return;
  const int j = head[(size_t)p * slots + e], kk = tail[(size_t)p * slots + e];
  float* yh = Yh + ((size_t)p * n_head + j) * DIM;
  float* Yt = move_other ? Yh : const_cast<float*>(Yt_in);
  float* yt = Yt + ((size_t)p * n_tail + kk) * DIM;
  float cur[DIM], oth[DIM], delta[DIM];
  float d2 = 0.f;
#pragma unroll
  for (int d = 0; d < DIM; ++d) { cur[d] = yh[d]; oth[d] = yt[d]; delta[d] = 0.f; const float t = cur[d] - oth[d]; d2 += t * t; }
  float g = 0.f;
  if (d2 > 0.f) {
    const float pw = __powf(d2, b - 1.f);
    g = (-2.f * a * b * pw) / (a * pw * d2 + 1.f);
  }
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const float gd = clip4(g * (cur[d] - oth[d])) * alpha;
    cur[d] += gd; delta[d] += gd;
    if (move_other) atomicAdd(&yt[d], -gd);
  }
  // negatives owed since the previous firing
  const float epsn = eps / nsr;
  int tot = (int)floorf((float)epoch / epsn) - 1;
  if (q > 1) {
    const int prev = (int)ceilf((float)(q - 1) * eps);
    tot -= (int)floorf((float)prev / epsn) - 1;
  }
  for (int s = 0; s < tot; ++s) {
    const uint32_t r = mix32(seed ^ ((uint64_t)e << 20) ^ ((uint64_t)epoch << 4) ^ (uint64_t)s ^ ((uint64_t)s << 40));
    const int kn = (int)(r % (uint32_t)n_tail);
    const float* yn = Yt + ((size_t)p * n_tail + kn) * DIM;
    float dn = 0.f;
    float on[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { on[d] = yn[d]; const float t = cur[d] - on[d]; dn += t * t; }
    float gn = 0.f;
    if (dn > 0.f) gn = (2.f * gamma * b) / ((0.001f + dn) * (a * __powf(dn, b) + 1.f));
    else if (move_other && j == kn) continue;
    if (gn > 0.f) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const float gd = clip4(gn * (cur[d] - on[d])) * alpha;
        cur[d] += gd; delta[d] += gd;
      }
    }
  }
#pragma unroll
  for (int d = 0; d < DIM; ++d) atomicAdd(&yh[d], delta[d]);
}

// 3-D embeddings (the reference's n_components=3) padded to float4: one 16-byte load per point and ONE vector float atomic
// (red.global.add.v4.f32, sm_90+) per moved endpoint instead of three scalar ones -- the per-epoch kernel is bound by the
// atomic throughput of L2.  Same schedule, RNG keys and update rule as sgd_epoch_kernel<3>.
__global__ void pack4_kernel(const float* __restrict__ Y, float4* __restrict__ Y4, size_t npts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npts) Y4[i] = make_float4(Y[3 * i], Y[3 * i + 1], Y[3 * i + 2], 0.f);
}
__global__ void unpack4_kernel(const float4* __restrict__ Y4, float* __restrict__ Y, size_t npts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npts) { const float4 v = Y4[i]; Y[3 * i] = v.x; Y[3 * i + 1] = v.y; Y[3 * i + 2] = v.z; }
}
// Only ~1/6 of the slots fire in an epoch, so a thread-per-slot kernel runs its update code with ~8 of 32 lanes active
// (ncu: 7.8 threads per instruction, issue bound).  Here every warp first tests kSgdSlotsPerLane x 32 consecutive slots and
// queues the ones that fire (ballot compaction into shared memory), then works through the queue with full warps.
constexpr int kSgdSlotsPerLane = 8;
constexpr int kSgdWarps = 8;
constexpr int kSgdNegBatch = 6;   // negative_sample_rate 5 owes 4..6 samples per firing
__global__ void __launch_bounds__(kSgdWarps * 32) sgd_epoch_kernel_v4(float4* __restrict__ Yh, float4* __restrict__ Yt, const int* __restrict__ head,
                                                                      const int* __restrict__ tail, const float* __restrict__ eps_arr, int slots,
                                                                      int n_head, int n_tail, int epoch, float a, float b, float gamma, float alpha,
                                                                      float nsr, int move_other, uint64_t seed) {
  __shared__ int s_queue[kSgdWarps][kSgdSlotsPerLane * 32];
  const int p = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int base = (blockIdx.x * kSgdWarps + warp) * (kSgdSlotsPerLane * 32);
  if (base >= slots) return;
  int* queue = s_queue[warp];
  int count = 0;
#pragma unroll
  for (int i = 0; i < kSgdSlotsPerLane; ++i) {
    const int e = base + i * 32 + lane;
    bool fire = false;
    if (e < slots) {
      const float eps = eps_arr[(size_t)p * slots + e];
      if (eps > 0.f) {
        const int q = (int)floorf((float)epoch / eps);
        fire = !(q < 1 || q <= (int)floorf((float)(epoch - 1) / eps));
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, fire);
    if (fire) queue[count + __popc(bal & ((1u << lane) - 1))] = e;
    count += __popc(bal);
  }
  __syncwarp();
  for (int qi = lane; qi < count; qi += 32) {
    const int e = queue[qi];
    const float eps = eps_arr[(size_t)p * slots + e];
    const int q = (int)floorf((float)epoch / eps);
    const int j = head[(size_t)p * slots + e], kk = tail[(size_t)p * slots + e];
    // negatives owed since the previous firing
    const float epsn = eps / nsr;
    int tot = (int)floorf((float)epoch / epsn) - 1;
    if (q > 1) {
      const int prev = (int)ceilf((float)(q - 1) * eps);
      tot -= (int)floorf((float)prev / epsn) - 1;
    }
    // The negative samples' addresses depend on (slot, epoch, s) only: the first kSgdNegBatch of them are requested here,
    // together with the two endpoints, instead of one L2 round trip per sample inside the dependent update chain.
    float4 n4[kSgdNegBatch];
#pragma unroll
    for (int u = 0; u < kSgdNegBatch; ++u) {
      if (u < tot) {
        const uint32_t r = mix32(seed ^ ((uint64_t)e << 20) ^ ((uint64_t)epoch << 4) ^ (uint64_t)u ^ ((uint64_t)u << 40));
        n4[u] = __ldcg(Yt + (size_t)p * n_tail + (int)(r % (uint32_t)n_tail));
      }
    }
    float4* yh = Yh + (size_t)p * n_head + j;
    float4* yt = Yt + (size_t)p * n_tail + kk;
    const float4 c4 = __ldcg(yh), o4 = __ldcg(yt);
    float cur[3] = {c4.x, c4.y, c4.z};
    const float oth[3] = {o4.x, o4.y, o4.z};
    float delta[3] = {0.f, 0.f, 0.f}, dt[3];
    float d2 = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) { const float t = cur[d] - oth[d]; d2 += t * t; }
    float g = 0.f;
    if (d2 > 0.f) {
      const float pw = __powf(d2, b - 1.f);
      g = (-2.f * a * b * pw) / (a * pw * d2 + 1.f);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float gd = clip4(g * (cur[d] - oth[d])) * alpha;
      cur[d] += gd; delta[d] += gd; dt[d] = -gd;
    }
    if (move_other) atomicAdd(yt, make_float4(dt[0], dt[1], dt[2], 0.f));
    auto repel = [&](const float4& nn) {   // (a negative at zero distance gives no update, whichever vertex it is)
      const float on[3] = {nn.x, nn.y, nn.z};
      float dn = 0.f;
#pragma unroll
      for (int d = 0; d < 3; ++d) { const float t = cur[d] - on[d]; dn += t * t; }
      if (dn > 0.f) {
        const float gn = (2.f * gamma * b) / ((0.001f + dn) * (a * __powf(dn, b) + 1.f));
        if (gn > 0.f) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const float gd = clip4(gn * (cur[d] - on[d])) * alpha;
            cur[d] += gd; delta[d] += gd;
          }
        }
      }
    };
#pragma unroll
    for (int u = 0; u < kSgdNegBatch; ++u)
      if (u < tot) repel(n4[u]);
    for (int sidx = kSgdNegBatch; sidx < tot; ++sidx) {   // (only with a larger negative_sample_rate)
      const uint32_t r = mix32(seed ^ ((uint64_t)e << 20) ^ ((uint64_t)epoch << 4) ^ (uint64_t)sidx ^ ((uint64_t)sidx << 40));
      repel(__ldcg(Yt + (size_t)p * n_tail + (int)(r % (uint32_t)n_tail)));
    }
    atomicAdd(yh, make_float4(delta[0], delta[1], delta[2], 0.f));
  }
}

// ------------------------------------------------------------------------------------------------
// Deterministic SGD for fit (the default for clouds that fit in shared memory): no atomics, all epochs in ONE launch.
//
// The fuzzy graph is symmetric with bit-identical weights in both directions (fuzzy_kernel: vij + vji - vij*vji), so the
// directed entries i->j and j->i have the same period eps and fire in the same epochs.  With all gradients of an epoch
// evaluated on the positions of the epoch's start (what a thread-per-slot kernel does up to scheduling order), vertex i gets
// from an edge {i,j} that fires:   g (its own entry, as head)  +  the negative-sample steps of that entry  +  g again (the
// twin entry j->i moves its tail by -g(j,i) = +g(i,j)).  Each vertex therefore needs only ITS OWN adjacency list (CSR by
// head, sgd_adj_kernel), computes its whole displacement itself and writes its new position: nothing is shared, the sum over
// a vertex's fired entries is taken in list order, the result is bit-reproducible for a given seed (the reference passes
// random_state=42 for exactly that; debug_tda_pipeline.py:100).
// A cloud is owned by a thread-block CLUSTER: every CTA keeps the whole embedding (double buffered, float4) and its slice of
// the adjacency in shared memory, moves its own vertices and stores the new positions into every CTA's next buffer through
// distributed shared memory; one cluster barrier per epoch.
constexpr int kAdjThreads = 1024;
constexpr int kAdjSplit = 4;      // CTAs per cloud: each counts the whole table (the offsets are global), then fills and sorts its share of the vertices
constexpr int kAdjUnroll = 4;     // slots per thread in flight (the passes are chains of dependent loads: bound by their latency)
__global__ void __launch_bounds__(kAdjThreads) sgd_adj_kernel(const int* __restrict__ head, const int* __restrict__ tail, const float* __restrict__ eps_arr,
                                                              int slots, int n, int* __restrict__ adj_off, uint2* __restrict__ adj_ent,
                                                              int* __restrict__ adj_maxdeg) {
  extern __shared__ int s_adj[];   // cnt[n], off[n + 1]
  int* cnt = s_adj;
  int* off = s_adj + n;
  __shared__ int s_part[kAdjThreads / 32];
  __shared__ int s_maxdeg;
  const int p = blockIdx.x / kAdjSplit, part = blockIdx.x % kAdjSplit, tid = threadIdx.x;
  const int vlo = (int)(((long long)n * part) / kAdjSplit), vhi = (int)(((long long)n * (part + 1)) / kAdjSplit);   // this CTA's vertices
  const int* H = head + (size_t)p * slots;
  const int* Tl = tail + (size_t)p * slots;
  const float* EP = eps_arr + (size_t)p * slots;
  uint2* ent = adj_ent + (size_t)p * slots;
  for (int i = tid; i < n; i += kAdjThreads) cnt[i] = 0;
  if (tid == 0) s_maxdeg = 0;
  __syncthreads();
  for (int e0 = tid; e0 < slots; e0 += kAdjUnroll * kAdjThreads) {
    float ep[kAdjUnroll];
    int h[kAdjUnroll];
#pragma unroll
    for (int u = 0; u < kAdjUnroll; ++u) {
      const int e = e0 + u * kAdjThreads;
      ep[u] = e < slots ? EP[e] : 0.f;
      h[u] = e < slots ? H[e] : 0;
    }
#pragma unroll
    for (int u = 0; u < kAdjUnroll; ++u)
      if (ep[u] > 0.f) atomicAdd(&cnt[h[u]], 1);
  }
  __syncthreads();
  // exclusive scan of cnt -> off (each thread a contiguous chunk)
  const int per = (n + kAdjThreads - 1) / kAdjThreads;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  int sum = 0, mx = 0;
  for (int i = lo; i < hi; ++i) { sum += cnt[i]; mx = max(mx, cnt[i]); }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
  if ((tid & 31) == 31) s_part[tid >> 5] = incl;
  if (mx) atomicMax(&s_maxdeg, mx);
  __syncthreads();
  int base = incl - sum;
  for (int w = 0; w < (tid >> 5); ++w) base += s_part[w];
  for (int i = lo; i < hi; ++i) { off[i] = base; base += cnt[i]; }
  if (hi == n && lo < n) off[n] = base;
  if (n == 0 && tid == 0) off[0] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += kAdjThreads) cnt[i] = off[i];   // fill cursors
  __syncthreads();
  for (int e0 = tid; e0 < slots; e0 += kAdjUnroll * kAdjThreads) {
    float ep[kAdjUnroll];
    int h[kAdjUnroll], t[kAdjUnroll];
#pragma unroll
    for (int u = 0; u < kAdjUnroll; ++u) {
      const int e = e0 + u * kAdjThreads;
      ep[u] = e < slots ? EP[e] : 0.f;
      h[u] = e < slots ? H[e] : 0;
      t[u] = e < slots ? Tl[e] : 0;
    }
#pragma unroll
    for (int u = 0; u < kAdjUnroll; ++u)
      if (ep[u] > 0.f && h[u] >= vlo && h[u] < vhi) {
        const int pos = atomicAdd(&cnt[h[u]], 1);
        ent[pos] = make_uint2(__float_as_uint(ep[u]), (uint32_t)t[u] | ((uint32_t)h[u] << 16));   // n <= 8192: 16 bits each
      }
  }
  __threadfence_block();
  __syncthreads();
  // the fill order depends on scheduling: sort every list by neighbour id (a vertex has each neighbour once)
  for (int v = vlo + tid; v < vhi; v += kAdjThreads) {
    const int a0 = off[v], a1 = off[v + 1];
    for (int i = a0 + 1; i < a1; ++i) {
      const uint2 x = ent[i];
      int k = i - 1;
      while (k >= a0 && (ent[k].y & 0xffffu) > (x.y & 0xffffu)) { ent[k + 1] = ent[k]; --k; }
      ent[k + 1] = x;
    }
  }
  for (int i = vlo + tid; i < vhi; i += kAdjThreads) adj_off[(size_t)p * (n + 1) + i] = off[i];
  if (part == kAdjSplit - 1 && tid == 0) adj_off[(size_t)p * (n + 1) + n] = off[n];
  if (tid == 0) adj_maxdeg[p] = s_maxdeg;   // (the same value from every CTA of the cloud)
}

constexpr int kClThreads = 1024;
constexpr int kClWarps = kClThreads / 32;
constexpr int kClQueue = 256;     // fired entries a warp queues before it works through them (a tile of 32 vertices usually fires ~150)
constexpr int kClMaxCluster = 8;
constexpr int kClTile = 16;       // most vertices per warp task (option sgd_tile: more, smaller tasks shorten a warp's chain per epoch)

// MUFU wrappers with flush-to-zero: without .ftz every lg2 / ex2 / rcp carries three extra instructions that rescale denormal
// operands (9 of ~45 per negative sample).  For normal operands the results are the same bits; a squared distance or a schedule
// period is never denormal.
__device__ __forceinline__ float lg2_ftz(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_ftz(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_ftz(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float pow_ftz(float x, float y) { return ex2_ftz(y * lg2_ftz(x)); }   // == __powf for normal x
__device__ __forceinline__ float div_ftz(float x, float y) { return x * rcp_ftz(y); }            // == __fdividef for normal y

struct SgdForce {
  float a, b, gamma, nsr;
  float m2ab, g2b, bm1;   // -2ab, 2*gamma*b, b-1 (set by the host next to a, b, gamma)
  // attractive step of entry (v -> t) on the epoch's positions; returns the step taken by v.  (Approximate division and
  // exp2/log2 intrinsics: the update is a stochastic gradient step; parity is statistical -- trustworthiness, downstream diagrams.)
  __device__ __forceinline__ float3 attract(const float4& yv, const float4& yt, float alpha) const {
    const float dx = yv.x - yt.x, dy = yv.y - yt.y, dz = yv.z - yt.z;
    const float d2 = dx * dx + dy * dy + dz * dz;
    float g = 0.f;
    if (d2 >= 1.17549435e-38f) {   // (a denormal d2 would be flushed by lg2: treated like coincident points)
      const float pw = pow_ftz(d2, bm1);
      g = div_ftz(m2ab * pw, fmaf(a * pw, d2, 1.f));
    }
    return make_float3(clip4(g * dx) * alpha, clip4(g * dy) * alpha, clip4(g * dz) * alpha);
  }
  __device__ __forceinline__ void repel(float3& cur, const float4& yn, float alpha) const {
    const float dx = cur.x - yn.x, dy = cur.y - yn.y, dz = cur.z - yn.z;
    const float dn = dx * dx + dy * dy + dz * dz;
    // a negative at zero distance gives no update, whichever vertex it is: dx = dy = dz = 0 there, and gn stays finite
    // (g2b / 0.001), so the three steps are exact zeros without a branch
    const float gn = div_ftz(g2b, (0.001f + dn) * fmaf(a, pow_ftz(dn, b), 1.f));   // > 0 for gamma > 0: always applied
    cur.x = fmaf(clip4(gn * dx), alpha, cur.x); cur.y = fmaf(clip4(gn * dy), alpha, cur.y); cur.z = fmaf(clip4(gn * dz), alpha, cur.z);
  }
  // does an entry of period eps fire in `epoch` (it fires when floor(epoch / eps) steps up; approximate division: the schedule is
  // this kernel's own definition, evaluated the same way everywhere) ...
  __device__ __forceinline__ bool fires(float eps, int epoch, int& q) const {
    const float r = rcp_ftz(eps);
    q = (int)floorf((float)epoch * r);
    return q >= 1 && q > (int)floorf((float)(epoch - 1) * r);
  }
  // ... and how many negative samples it owes since its previous firing
  __device__ __forceinline__ int negatives(float eps, int epoch, int q) const {
    const float rn = rcp_ftz(div_ftz(eps, nsr));
    int tot = (int)floorf((float)epoch * rn) - 1;
    if (q > 1) {
      const int prev = (int)ceilf((float)(q - 1) * eps);
      tot -= (int)floorf((float)prev * rn) - 1;
    }
    return tot;
  }
};
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(const void* smem_ptr, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void dsmem_store_f4(uint32_t addr, const float4& v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// grid = batch * C CTAs, cluster (C,1,1); dynamic shared memory: Yb[2][n] float4, soff[nown + 1], sent[own entries] (optional),
// per warp: queue[kClQueue] + acc[32] float4
__global__ void __launch_bounds__(kClThreads, 1) sgd_cluster_kernel(float* __restrict__ Y, const int* __restrict__ adj_off, const uint2* __restrict__ adj_ent,
                                                                    int slots, int n, int n_epochs, SgdForce F, float alpha0, uint64_t seed,
                                                                    int max_own_ent, int tile) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ int s_next_tile[2];
  const uint32_t C = cluster_nctarank(), cr = cluster_ctarank();
  const int p = blockIdx.x / (int)C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_next_tile[0] = 0; s_next_tile[1] = 0; }
  // the vertices are dealt to the CTAs in whole tiles [tile*j, tile*j + tile): a tile's fired entries then fall into the same
  // 32-lane batches for every cluster size, so the embedding does not depend on how many CTAs share a cloud
  const int ntile = (n + tile - 1) / tile;
  const int v0 = min(n, tile * (int)(((long long)ntile * cr) / C)), v1 = min(n, tile * (int)(((long long)ntile * (cr + 1)) / C));
  const int nown = v1 - v0;
  const int nown_max = ((ntile + (int)C - 1) / (int)C) * tile + 1;
  float4* Yb = reinterpret_cast<float4*>(s_raw);
  float4* acc_all = Yb + 2 * (size_t)n;
  uint32_t* queue_all = reinterpret_cast<uint32_t*>(acc_all + kClWarps * kClTile);
  int* soff = reinterpret_cast<int*>(queue_all + kClWarps * kClQueue);
  uint2* sent = reinterpret_cast<uint2*>(soff + ((nown_max + 2) & ~1));
  const int* goff = adj_off + (size_t)p * (n + 1);
  const uint2* gent = adj_ent + (size_t)p * slots;
  float* Yg = Y + (size_t)p * n * 3;
  for (int i = tid; i < n; i += kClThreads) Yb[i] = make_float4(Yg[3 * i], Yg[3 * i + 1], Yg[3 * i + 2], 0.f);
  for (int i = tid; i <= nown; i += kClThreads) soff[i] = goff[v0 + i];
  __syncthreads();
  const int e0 = soff[0], ne = soff[nown] - e0;
  const bool ent_in_smem = ne <= max_own_ent;
  if (ent_in_smem)
    for (int i = tid; i < ne; i += kClThreads) sent[i] = gent[e0 + i];
  const uint2* ent = ent_in_smem ? sent : gent + e0;   // entry i of this CTA's slice
  float4* acc = acc_all + warp * kClTile;
  uint32_t* queue = queue_all + warp * kClQueue;
  uint32_t remote[kClMaxCluster];
#pragma unroll
  for (int r = 0; r < kClMaxCluster; ++r) remote[r] = r < (int)C ? dsmem_addr(Yb, (uint32_t)r) : 0u;
  cluster_sync_all();   // every CTA of the cluster has its buffers up before anybody stores into them

  for (int ep = 0; ep < n_epochs; ++ep) {
    const float alpha = ep == 0 ? alpha0 : alpha0 * (1.f - (float)(ep - 1) / (float)n_epochs);
    const float4* src = Yb + (size_t)(ep & 1) * n;
    const uint32_t dst_off = (uint32_t)(((ep + 1) & 1) * n) * 16u;
    const uint32_t key_ep = hash32((uint32_t)seed ^ (uint32_t)(seed >> 32) ^ hash32(0x27d4eb2fu + (uint32_t)ep));
    // tiles of `tile` consecutive owned vertices, handed to the warps one by one (a counter per epoch parity in shared memory): a
    // tile needs 1..4 rounds of 32 fired entries, and the epoch barrier waits for the slowest warp.  Which warp takes a tile does
    // not matter for the result: a tile reads the epoch's old buffer and writes only its own vertices.
    if (tid == 0) s_next_tile[(ep + 1) & 1] = 0;   // (last used in epoch ep - 1, which every warp has left)
    for (;;) {
      int tix = 0;
      if (lane == 0) tix = atomicAdd(&s_next_tile[ep & 1], 1);
      tix = __shfl_sync(0xffffffffu, tix, 0);
      const int tv = tix * tile;
      if (tv >= nown) break;
      const int cnt = min(tile, nown - tv);
      const int ebase = soff[tv] - e0;
      if (lane < tile) acc[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
      int qn = 0;
      // phase 2 (called whenever the queue may overflow, and at the end): full warps over the queued (= fired) entries; the
      // displacement of each, summed per vertex in queue (= list) order
      auto drain = [&]() {
        __syncwarp();
        for (int q0 = 0; q0 < qn; q0 += 32) {
          const int qi = q0 + lane;
          float3 dl = make_float3(0.f, 0.f, 0.f);
          int vi = -1 - lane;   // distinct dummy keys for idle lanes
          if (qi < qn) {
            const int i = ebase + (int)queue[qi];
            const uint2 en = ent[i];
            const float eps = __uint_as_float(en.x);
            const int v = (int)(en.y >> 16);
            vi = v - (v0 + tv);
            int q;
            F.fires(eps, ep, q);
            const int tot = F.negatives(eps, ep, q);
            const float4 yv = src[v], yt = src[en.y & 0xffffu];
            const float3 g1 = F.attract(yv, yt, alpha);
            float3 cur = make_float3(yv.x + g1.x, yv.y + g1.y, yv.z + g1.z);
            uint32_t r = hash32(key_ep + (uint32_t)(e0 + i) * 0x9E3779B1u);   // stream of this (entry, epoch): one LCG step per sample
            for (int sidx = 0; sidx < tot; ++sidx) {
              r = r * 0x2c9277b5u + 0xac564b05u;
              F.repel(cur, src[__umulhi(r, (uint32_t)n)], alpha);
            }
            dl = make_float3((cur.x - yv.x) + g1.x, (cur.y - yv.y) + g1.y, (cur.z - yv.z) + g1.z);
          }
          // segmented inclusive scan over the lanes (segments = runs of equal vi), the last lane of a run owns its sum
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const float tx = __shfl_up_sync(0xffffffffu, dl.x, o), ty = __shfl_up_sync(0xffffffffu, dl.y, o), tz = __shfl_up_sync(0xffffffffu, dl.z, o);
            const int tvi = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o && tvi == vi) { dl.x += tx; dl.y += ty; dl.z += tz; }
          }
          const int nvi = __shfl_down_sync(0xffffffffu, vi, 1);
          if (qi < qn && (lane == 31 || nvi != vi)) {
            float4 t = acc[vi];
            t.x += dl.x; t.y += dl.y; t.z += dl.z;
            acc[vi] = t;
          }
          __syncwarp();
        }
        qn = 0;
      };
      // phase 1: which entries of the tile fire (lanes over the tile's contiguous entry range: the queue stays in list order)
      const int eend = soff[tv + cnt] - e0;
      for (int i0 = ebase; i0 < eend; i0 += 32) {
        if (qn + 32 > kClQueue) drain();
        const int i = i0 + lane;
        bool fire = false;
        int q;
        if (i < eend) fire = F.fires(__uint_as_float(ent[i].x), ep, q);
        const unsigned bal = __ballot_sync(0xffffffffu, fire);
        if (fire) queue[qn + __popc(bal & ((1u << lane) - 1))] = (uint32_t)(i - ebase);
        qn += __popc(bal);
      }
      drain();
      // phase 3: new positions of the tile's vertices into the next buffer of every CTA of the cluster
      if (lane < cnt) {
        const int v = v0 + tv + lane;
        const float4 y = src[v], d = acc[lane];
        const float4 yn = make_float4(y.x + d.x, y.y + d.y, y.z + d.z, 0.f);
#pragma unroll
        for (int r = 0; r < kClMaxCluster; ++r)
          if (r < (int)C) dsmem_store_f4(remote[r] + dst_off + (uint32_t)v * 16u, yn);
      }
      __syncwarp();
    }
    cluster_sync_all();
  }
  const float4* fin = Yb + (size_t)(n_epochs & 1) * n;
  for (int i = tid; i < nown; i += kClThreads) {
    const float4 y = fin[v0 + i];
    Yg[3 * (v0 + i)] = y.x; Yg[3 * (v0 + i) + 1] = y.y; Yg[3 * (v0 + i) + 2] = y.z;
  }
}

// transform (move_other = 0): a query point only ever reads the fixed training embedding, so it is optimised on its own -- one
// warp per query point, lanes over its k entries, all epochs in registers; no atomics, deterministic.
__global__ void __launch_bounds__(256) sgd_transform_kernel(float* __restrict__ Y, const float* __restrict__ Yt, const int* __restrict__ tail,
                                                            const float* __restrict__ eps_arr, int k, int n_query, int n_train, int n_epochs,
                                                            SgdForce F, float alpha0, uint64_t seed) {
  const int p = blockIdx.y;
  const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= n_query) return;
  const float* T3 = Yt + (size_t)p * n_train * 3;
  float* yq = Y + ((size_t)p * n_query + v) * 3;
  float4 cur = make_float4(yq[0], yq[1], yq[2], 0.f);
  const size_t sbase = (size_t)p * n_query * k + (size_t)v * k;
  for (int ep = 0; ep < n_epochs; ++ep) {
    const float alpha = ep == 0 ? alpha0 : alpha0 * (1.f - (float)(ep - 1) / (float)n_epochs);
    float3 dl = make_float3(0.f, 0.f, 0.f);
    const uint32_t key_ep = hash32((uint32_t)seed ^ (uint32_t)(seed >> 32) ^ hash32(0x27d4eb2fu + (uint32_t)ep));
    for (int t = lane; t < k; t += 32) {
      const float eps = eps_arr[sbase + t];
      int q;
      if (eps > 0.f && F.fires(eps, ep, q)) {
        const int tot = F.negatives(eps, ep, q);
        const int j = tail[sbase + t];
        const float4 yt = make_float4(__ldg(&T3[3 * j]), __ldg(&T3[3 * j + 1]), __ldg(&T3[3 * j + 2]), 0.f);
        const float3 g1 = F.attract(cur, yt, alpha);
        float3 c = make_float3(cur.x + g1.x, cur.y + g1.y, cur.z + g1.z);
        uint32_t r = hash32(key_ep + (uint32_t)(sbase + t) * 0x9E3779B1u);
        for (int sidx = 0; sidx < tot; ++sidx) {
          r = r * 0x2c9277b5u + 0xac564b05u;
          const int kn = (int)__umulhi(r, (uint32_t)n_train);
          F.repel(c, make_float4(__ldg(&T3[3 * kn]), __ldg(&T3[3 * kn + 1]), __ldg(&T3[3 * kn + 2]), 0.f), alpha);
        }
        dl.x += c.x - cur.x; dl.y += c.y - cur.y; dl.z += c.z - cur.z;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dl.x += __shfl_xor_sync(0xffffffffu, dl.x, o); dl.y += __shfl_xor_sync(0xffffffffu, dl.y, o); dl.z += __shfl_xor_sync(0xffffffffu, dl.z, o);
    }
    cur.x += dl.x; cur.y += dl.y; cur.z += dl.z;
  }
  if (lane == 0) { yq[0] = cur.x; yq[1] = cur.y; yq[2] = cur.z; }
}

// ------------------------------------------------------------------------------------------------
// initialisation helpers
__device__ __forceinline__ float u01(uint64_t key) { return ((float)(mix32(key) >> 8) + 0.5f) * (1.f / 16777216.f); }

__global__ void init_random_kernel(float* __restrict__ Y, int total, int per_cloud, float lo, float hi, uint64_t seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (the draw depends on the index inside the cloud, not on the cloud's place in the batch)
  if (i < total) Y[i] = lo + (hi - lo) * u01(seed * 0x9E3779B97F4A7C15ull + (uint64_t)(i % per_cloud));
}

// Y <- 10 * minmax( Y * (10 / max|Y|) + N(0, noise) ) per cloud and axis  (umap-learn's noisy_scale_coords
// followed by the [0,10] rescale of simplicial_set_embedding).  One CTA per cloud.
__global__ void __launch_bounds__(1024) rescale_kernel(float* __restrict__ Yg, int n, int dim, float noise, uint64_t seed) {
  __shared__ float s_red[32];
  __shared__ float s_val[2 * 8 + 1];
  const int p = blockIdx.x;
  float* Y = Yg + (size_t)p * n * dim;
  const int tid = threadIdx.x, nt = blockDim.x;
  auto block_max = [&](float v) -> float {
    v = warp_max_f32(v);
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    float m = -INFINITY;
    for (int w = 0; w < (nt >> 5); ++w) m = fmaxf(m, s_red[w]);
    __syncthreads();
    return m;
  };
  float am = 0.f;
  for (int i = tid; i < n * dim; i += nt) am = fmaxf(am, fabsf(Y[i]));
  am = block_max(am);
  const float expansion = am > 0.f ? 10.f / am : 1.f;
  for (int i = tid; i < n * dim; i += nt) {
    const uint64_t key = seed * 0xD6E8FEB86659FD93ull + (uint64_t)i * 2;   // (no cloud index: every cloud of a batch draws what it would draw alone)
    const float u1 = u01(key), u2 = u01(key + 1);
    const float g = sqrtf(-2.f * logf(u1)) * cosf(6.28318530718f * u2);
    Y[i] = Y[i] * expansion + noise * g;
  }
  __syncthreads();
  for (int d = 0; d < dim; ++d) {
    float mx = -INFINITY, mn = INFINITY;
    for (int i = tid; i < n; i += nt) { const float v = Y[(size_t)i * dim + d]; mx = fmaxf(mx, v); mn = fminf(mn, v); }
    mx = block_max(mx);
    mn = -block_max(-mn);
    if (tid == 0) { s_val[2 * d] = mn; s_val[2 * d + 1] = mx; }
  }
  __syncthreads();
  for (int i = tid; i < n * dim; i += nt) {
    const int d = i % dim;
    const float mn = s_val[2 * d], mx = s_val[2 * d + 1];
    Y[i] = mx > mn ? 10.f * (Y[i] - mn) / (mx - mn) : 0.f;
  }
}

// init_transform: new point = sum_t w_t * train_embedding[idx_t], w = l1-normalised membership strengths
__global__ void transform_init_kernel(const int* __restrict__ knn_idx, const float* __restrict__ knn_dist, const float* __restrict__ sigma,
                                      const float* __restrict__ rho, const float* __restrict__ train, int nq, int ntrain, int k, int dim,
                                      float* __restrict__ Y, int* __restrict__ head, int* __restrict__ tail, float* __restrict__ weight,
                                      unsigned int* __restrict__ max_w_bits) {
  const int p = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const size_t base = ((size_t)p * nq + i) * k;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float wsum = 0.f, wmax = 0.f;
  for (int t = 0; t < k; ++t) {
    const int j = knn_idx[base + t];
    float w = 0.f;
    if (j >= 0) w = membership(knn_dist[base + t], rho[(size_t)p * nq + i], sigma[(size_t)p * nq + i], false, true);
    head[base + t] = i; tail[base + t] = j < 0 ? 0 : j; weight[base + t] = w;
    wsum += w; wmax = fmaxf(wmax, w);
    if (j >= 0)
      for (int d = 0; d < dim; ++d) acc[d] += w * train[((size_t)p * ntrain + j) * dim + d];
  }
  for (int d = 0; d < dim; ++d) Y[((size_t)p * nq + i) * dim + d] = wsum > 0.f ? acc[d] / wsum : 0.f;
  if (wmax > 0.f) atomicMax(&max_w_bits[p], __float_as_uint(wmax));
}

}  // namespace umap
}  // namespace tda

using namespace tda;
using namespace tda::umap;

extern "C" int tda_knn_smooth(const float* D, int n, int m, int batch, int k, float local_connectivity, float bandwidth, int n_iter,
                              int32_t* knn_idx, float* knn_dist, float* sigma, float* rho, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!D || !knn_idx || !knn_dist || !sigma || !rho || !ws || n <= 0 || m <= 0 || batch <= 0) return set_error(TDA_ERR_INVALID, "tda_knn_smooth: bad arguments");
  if (k < 1 || k > m) return set_error(TDA_ERR_INVALID, "tda_knn_smooth: k=%d out of range (1..%d)", k, m);
  if (k > 256) return set_error(TDA_ERR_UNSUPPORTED, "tda_knn_smooth: k=%d > 256", k);
  if (ws_bytes < sizeof(double) * (size_t)batch) return set_error(TDA_ERR_WORKSPACE, "tda_knn_smooth: workspace too small");
  double* dist_sum = (double*)ws;
  StageScope st(STAGE_KNN_SMOOTH, stream);
  TDA_CUDA_CHECK(cudaMemsetAsync(dist_sum, 0, sizeof(double) * batch, stream));
  dim3 grid((n + 7) / 8, batch);
  const int kpl = (k + 31) / 32;
#define TDA_KNN_LAUNCH(KPL) knn_smooth_kernel<KPL><<<grid, 256, 0, stream>>>(D, n, m, k, local_connectivity, bandwidth, n_iter, knn_idx, knn_dist, sigma, rho, dist_sum)
  if (k <= 16) {
    dim3 gp((n + kBlkRows - 1) / kBlkRows, batch);  // 4 warps, 32 rows per CTA
    const int loads = (int)option("knn_loads");
    if (loads == 8) knn_smooth_block_kernel<8><<<gp, kBlkWarps * 32, 0, stream>>>(D, n, m, k, local_connectivity, bandwidth, n_iter, knn_idx, knn_dist, sigma, rho, dist_sum);
    else knn_smooth_block_kernel<16><<<gp, kBlkWarps * 32, 0, stream>>>(D, n, m, k, local_connectivity, bandwidth, n_iter, knn_idx, knn_dist, sigma, rho, dist_sum);
  } else if (kpl == 1) TDA_KNN_LAUNCH(1);
  else if (kpl == 2) TDA_KNN_LAUNCH(2);
  else if (kpl <= 4) TDA_KNN_LAUNCH(4);
  else TDA_KNN_LAUNCH(8);
#undef TDA_KNN_LAUNCH
  dim3 g2((n + 255) / 256, batch);
  sigma_floor_kernel<<<g2, 256, 0, stream>>>(n, k, rho, sigma, dist_sum);
  count_launch(2);
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}

// ---- exact kNN + sigma/rho of the rows [row_begin, row_end) of ONE cloud against all of its points, without the n x n matrix:
// the tensor-core distance GEMM runs on blocks of `row_block` rows and every block goes straight into the top-k kernel (the block
// lives in the workspace and is overwritten by the next one).  This is the per-rank body of the row-sharded kNN of config C5
// (SURVEY.md section 8b `tda_knn_fused`, section 8e).
namespace tda { namespace umap {
__global__ void zero_self_kernel(float* __restrict__ D, int rows, int m, int col0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows && col0 + i < m) D[(size_t)i * m + col0 + i] = 0.f;   // the point itself: exactly zero (sklearn zeroes the diagonal)
}
} }
extern "C" size_t tda_knn_fused_workspace_bytes(int n, int d, int row_block) {
  if (n <= 0 || d <= 0 || row_block <= 0) return 0;
  const int rb = row_block < n ? row_block : n;
  return ((tda_pdist_workspace_bytes(rb, n, d, 1, 0) + 255) & ~(size_t)255) + ((sizeof(float) * (size_t)rb * (size_t)n + 255) & ~(size_t)255) + 1024;
}
extern "C" int tda_knn_fused(const float* X, int n, int d, int row_begin, int row_end, int k, int metric, float disconnect,
                             float local_connectivity, int32_t* knn_idx, float* knn_dist, float* sigma, float* rho, int row_block, void* ws,
                             size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!X || !knn_idx || !knn_dist || !sigma || !rho || !ws || n <= 0 || d <= 0 || row_begin < 0 || row_end > n || row_begin >= row_end || row_block <= 0)
    return set_error(TDA_ERR_INVALID, "tda_knn_fused: bad arguments");
  if (k < 1 || k > n) return set_error(TDA_ERR_INVALID, "tda_knn_fused: k=%d out of range (1..%d)", k, n);
  const int rb = row_block < n ? row_block : n;
  const size_t need = tda_knn_fused_workspace_bytes(n, d, rb);
  if (ws_bytes < need) return set_error(TDA_ERR_WORKSPACE, "tda_knn_fused: workspace %zu < required %zu", ws_bytes, need);
  if ((((uintptr_t)ws) & 255) != 0) return set_error(TDA_ERR_INVALID, "tda_knn_fused: workspace must be 256-byte aligned");
  const size_t pd_bytes = (tda_pdist_workspace_bytes(rb, n, d, 1, 0) + 255) & ~(size_t)255;
  float* Dblk = (float*)((char*)ws + pd_bytes);
  double* ksum = (double*)((char*)Dblk + ((sizeof(float) * (size_t)rb * (size_t)n + 255) & ~(size_t)255));
  for (int b0 = row_begin; b0 < row_end; b0 += rb) {
    const int rows = (row_end - b0) < rb ? (row_end - b0) : rb;
    // Y = X + something would be "the same set" for tda_pdist; a row block against all points is the general (two-operand) case
    int rc = tda_pdist(X + (size_t)b0 * d, X, rows, n, d, 1, metric, disconnect, Dblk, ws, pd_bytes, stream_);
    if (rc != TDA_OK) return rc;
    tda::umap::zero_self_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(Dblk, rows, n, b0);
    count_launch();
    const size_t off = (size_t)(b0 - row_begin);
    rc = tda_knn_smooth(Dblk, rows, n, 1, k, local_connectivity, 1.0f, 64, knn_idx + off * k, knn_dist + off * k, sigma + off, rho + off, ksum, 64,
                        stream_);
    if (rc != TDA_OK) return rc;
  }
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}

extern "C" int tda_fuzzy_graph(const int32_t* knn_idx, const float* knn_dist, const float* sigma, const float* rho, int n, int k, int batch,
                               float mix_ratio, int n_epochs, int32_t* head, int32_t* tail, float* weight, float* eps, float* max_weight,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!knn_idx || !knn_dist || !sigma || !rho || !head || !tail || !weight || !eps || !max_weight || n <= 0 || k <= 0 || batch <= 0)
    return set_error(TDA_ERR_INVALID, "tda_fuzzy_graph: bad arguments");
  StageScope st(STAGE_FUZZY, stream);
  TDA_CUDA_CHECK(cudaMemsetAsync(max_weight, 0, sizeof(float) * batch, stream));
  dim3 g((n * k + 255) / 256, batch);
  fuzzy_kernel<<<g, 256, 0, stream>>>(knn_idx, knn_dist, sigma, rho, n, k, mix_ratio, head, tail, weight, (unsigned int*)max_weight);
  dim3 g2((2 * n * k + 255) / 256, batch);
  epochs_kernel<<<g2, 256, 0, stream>>>(weight, 2 * n * k, n_epochs, (const unsigned int*)max_weight, eps);
  count_launch(2);
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}

// workspace of tda_umap_sgd: the float4-padded embeddings of the per-epoch kernels, or the per-vertex adjacency (CSR by head:
// offsets, (eps, neighbour) entries, largest degree) of the deterministic cluster kernel -- whichever is larger
static size_t sgd_adj_bytes(int slots, int n, int batch) {
  return (((size_t)batch * (n + 1) * sizeof(int) + 255) & ~(size_t)255) + (((size_t)batch * slots * sizeof(uint2) + 255) & ~(size_t)255) +
         (((size_t)batch * sizeof(int) + 255) & ~(size_t)255);
}
extern "C" size_t tda_umap_sgd_workspace_bytes(int slots, int n_head, int n_tail, int dim, int batch, int move_other) {
  if (slots <= 0 || n_head <= 0 || batch <= 0) return 0;
  size_t b = sizeof(float4) * (size_t)batch * ((size_t)n_head + (move_other ? 0 : (size_t)n_tail));
  if (dim == 3 && move_other) b = b > sgd_adj_bytes(slots, n_head, batch) ? b : sgd_adj_bytes(slots, n_head, batch);
  return b + 256;
}

extern "C" int tda_umap_sgd(float* Y, const float* Y_other, const int32_t* head, const int32_t* tail, const float* eps, int slots, int n_head,
                            int n_tail, int dim, int batch, int n_epochs, float a, float b, float gamma, float alpha0,
                            float negative_sample_rate, int move_other, uint64_t seed, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!Y || !head || !tail || !eps || slots <= 0 || n_head <= 0 || n_tail <= 0 || batch <= 0 || n_epochs < 0)
    return set_error(TDA_ERR_INVALID, "tda_umap_sgd: bad arguments");
  if (!move_other && !Y_other) return set_error(TDA_ERR_INVALID, "tda_umap_sgd: Y_other required when move_other=0");
  if (dim < 1 || dim > 4) return set_error(TDA_ERR_UNSUPPORTED, "tda_umap_sgd: n_components=%d (supported: 1..4)", dim);
  dim3 g((slots + 255) / 256, batch);
  StageScope st(STAGE_SGD, stream);
  const int sgd_mode = (int)option("sgd_mode");
  SgdForce F;
  F.a = a; F.b = b; F.gamma = gamma; F.nsr = negative_sample_rate;
  F.m2ab = -2.f * a * b; F.g2b = 2.f * gamma * b; F.bm1 = b - 1.f;
  if (n_epochs == 0) return TDA_OK;
  // ---- deterministic paths (sgd_mode 0): cluster kernel for fit, warp-per-point kernel for transform
  if (sgd_mode == 0 && dim == 3 && !move_other && slots % n_head == 0 && batch <= 65535) {
    dim3 gt((n_head + 7) / 8, batch);
    sgd_transform_kernel<<<gt, 256, 0, stream>>>(Y, Y_other, tail, eps, slots / n_head, n_head, n_tail, n_epochs, F, alpha0, seed);
    count_launch();
    TDA_LAUNCH_CHECK();
    return TDA_OK;
  }
  if (sgd_mode == 0 && dim == 3 && move_other && n_head == n_tail && n_head <= 8192 && ws && ws_bytes >= sgd_adj_bytes(slots, n_head, batch) &&
      (((uintptr_t)ws) & 255) == 0) {
    const int n = n_head;
    int C = (int)option("sgd_cluster");
    if (C != 1 && C != 2 && C != 4 && C != 8) C = batch <= 4 ? 8 : 4;   // auto: few clouds -> more SMs per cloud (the kernel is latency bound)
    int tile = (int)option("sgd_tile");
    if (tile < 1 || tile > kClTile) tile = kClTile;
    const int nown_max = (((n + tile - 1) / tile + C - 1) / C) * tile + 1;
    const size_t base = sizeof(float4) * 2 * (size_t)n + sizeof(float4) * kClWarps * kClTile + sizeof(uint32_t) * kClWarps * kClQueue +
                        sizeof(int) * (size_t)((nown_max + 2) & ~1);
    const size_t smem_max = (size_t)224 * 1024;
    if (base + 1024 <= smem_max) {
      Carver c(ws, ws_bytes);
      int* adj_off = c.take<int>((size_t)batch * (n + 1));
      uint2* adj_ent = c.take<uint2>((size_t)batch * slots);
      int* adj_maxdeg = c.take<int>(batch);
      const size_t adj_smem = sizeof(int) * (size_t)(2 * n + 1);
      TDA_CUDA_CHECK(cudaFuncSetAttribute(sgd_adj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adj_smem));
      sgd_adj_kernel<<<batch * kAdjSplit, kAdjThreads, adj_smem, stream>>>(head, tail, eps, slots, n, adj_off, adj_ent, adj_maxdeg);
      // the CTA's slice of the adjacency goes to shared memory when it fits into what the embedding buffers leave
      size_t cap_ent = (smem_max - base) / sizeof(uint2);
      const size_t want_ent = (size_t)slots / C + (size_t)slots / (2 * C) + 64;   // 1.5x the mean slice (the kernel checks its actual size)
      if (cap_ent > want_ent) cap_ent = want_ent;
      const size_t dyn = base + cap_ent * sizeof(uint2);
      TDA_CUDA_CHECK(cudaFuncSetAttribute(sgd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)(batch * C), 1, 1);
      cfg.blockDim = dim3(kClThreads, 1, 1);
      cfg.dynamicSmemBytes = dyn;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      TDA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sgd_cluster_kernel, Y, (const int*)adj_off, (const uint2*)adj_ent, slots, n, n_epochs, F, alpha0, seed,
                                        (int)cap_ent, tile));
      count_launch(2);
      TDA_LAUNCH_CHECK();
      return TDA_OK;
    }
  }
  // ---- per-epoch kernels with float atomics (large clouds, other dimensions, sgd_mode 3): not reproducible run to run
  const size_t np_h = (size_t)batch * n_head, np_t = move_other ? 0 : (size_t)batch * n_tail;
  if (dim == 3 && ws && ws_bytes >= sizeof(float4) * (np_h + np_t) && (((uintptr_t)ws) & 15) == 0) {
    float4* Yh4 = (float4*)ws;
    float4* Yt4 = move_other ? Yh4 : Yh4 + np_h;
    pack4_kernel<<<(unsigned)((np_h + 255) / 256), 256, 0, stream>>>(Y, Yh4, np_h);
    if (!move_other) pack4_kernel<<<(unsigned)((np_t + 255) / 256), 256, 0, stream>>>(Y_other, Yt4, np_t);
    for (int ep = 0; ep < n_epochs; ++ep) {
      const float alpha = ep == 0 ? alpha0 : alpha0 * (1.f - (float)(ep - 1) / (float)n_epochs);
      const int per_block = kSgdWarps * kSgdSlotsPerLane * 32;
      dim3 g4((slots + per_block - 1) / per_block, batch);
      sgd_epoch_kernel_v4<<<g4, kSgdWarps * 32, 0, stream>>>(Yh4, Yt4, head, tail, eps, slots, n_head, n_tail, ep, a, b, gamma, alpha,
                                                             negative_sample_rate, move_other, seed);
    }
    unpack4_kernel<<<(unsigned)((np_h + 255) / 256), 256, 0, stream>>>(Yh4, Y, np_h);
    count_launch(n_epochs + 2 + (move_other ? 0 : 1));
    TDA_LAUNCH_CHECK();
    return TDA_OK;
  }
  for (int ep = 0; ep < n_epochs; ++ep) {
    const float alpha = ep == 0 ? alpha0 : alpha0 * (1.f - (float)(ep - 1) / (float)n_epochs);
#define TDA_SGD_LAUNCH(DIM) sgd_epoch_kernel<DIM><<<g, 256, 0, stream>>>(Y, Y_other, head, tail, eps, slots, n_head, n_tail, ep, a, b, gamma, alpha, negative_sample_rate, move_other, seed)
    switch (dim) {
      case 1: TDA_SGD_LAUNCH(1); break;
      case 2: TDA_SGD_LAUNCH(2); break;
      case 3: TDA_SGD_LAUNCH(3); break;
      default: TDA_SGD_LAUNCH(4); break;
    }
#undef TDA_SGD_LAUNCH
  }
  count_launch(n_epochs);
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}

extern "C" int tda_umap_init_random(float* Y, int n, int dim, int batch, float lo, float hi, uint64_t seed, void* stream_) {
  if (!Y || n <= 0 || dim <= 0 || batch <= 0) return set_error(TDA_ERR_INVALID, "tda_umap_init_random: bad arguments");
  const int total = n * dim * batch;
  init_random_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(Y, total, n * dim, lo, hi, seed);
  count_launch();
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}

extern "C" int tda_umap_rescale(float* Y, int n, int dim, int batch, float noise, uint64_t seed, void* stream_) {
  if (!Y || n <= 0 || dim <= 0 || dim > 8 || batch <= 0) return set_error(TDA_ERR_INVALID, "tda_umap_rescale: bad arguments");
  rescale_kernel<<<batch, 1024, 0, (cudaStream_t)stream_>>>(Y, n, dim, noise, seed);
  count_launch();
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}

extern "C" int tda_umap_transform_init(const int32_t* knn_idx, const float* knn_dist, const float* sigma, const float* rho,
                                       const float* train_embedding, int n_query, int n_train, int k, int dim, int batch, int n_epochs,
                                       float* Y, int32_t* head, int32_t* tail, float* weight, float* eps, float* max_weight, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!knn_idx || !knn_dist || !sigma || !rho || !train_embedding || !Y || !head || !tail || !weight || !eps || !max_weight || dim > 8)
    return set_error(TDA_ERR_INVALID, "tda_umap_transform_init: bad arguments");
  TDA_CUDA_CHECK(cudaMemsetAsync(max_weight, 0, sizeof(float) * batch, stream));
  dim3 g((n_query + 127) / 128, batch);
  transform_init_kernel<<<g, 128, 0, stream>>>(knn_idx, knn_dist, sigma, rho, train_embedding, n_query, n_train, k, dim, Y, head, tail, weight,
                                               (unsigned int*)max_weight);
  dim3 g2((n_query * k + 255) / 256, batch);
  epochs_kernel<<<g2, 256, 0, stream>>>(weight, n_query * k, n_epochs, (const unsigned int*)max_weight, eps);
  count_launch(2);
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}
