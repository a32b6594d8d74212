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
    const uint32_t r = mix32(seed ^ ((uint64_t)p << 52) ^ ((uint64_t)e << 20) ^ ((uint64_t)epoch << 4) ^ (uint64_t)s ^ ((uint64_t)s << 40));
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
        const uint32_t r = mix32(seed ^ ((uint64_t)p << 52) ^ ((uint64_t)e << 20) ^ ((uint64_t)epoch << 4) ^ (uint64_t)u ^ ((uint64_t)u << 40));
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
      const uint32_t r = mix32(seed ^ ((uint64_t)p << 52) ^ ((uint64_t)e << 20) ^ ((uint64_t)epoch << 4) ^ (uint64_t)sidx ^ ((uint64_t)sidx << 40));
      repel(__ldcg(Yt + (size_t)p * n_tail + (int)(r % (uint32_t)n_tail)));
    }
    atomicAdd(yh, make_float4(delta[0], delta[1], delta[2], 0.f));
  }
}

// EXPERIMENT (off unless TDA_SGD_AGG=1; written after the round's GPU minutes were spent: compiled, never run): the per-epoch
// kernel with the updates of the slot block's own vertex summed in the warp.  The slot table is row-major -- the 2k slots of
// source vertex i are consecutive, k with head i and k with tail i (fuzzy_kernel) -- so every fired slot of a block moves
// vertex i; a segmented shuffle reduction over the (queue-ordered, hence contiguous) lanes of the same block leaves one vector
// RED per block for vertex i plus one per fired edge for the far endpoint: ~1.1 instead of 2 atomics per fired edge, on a
// stage that is bound by the L2 atomic rate.  Same schedule, RNG keys and update rule; fit (move_other) only.
__global__ void __launch_bounds__(kSgdWarps * 32) sgd_epoch_kernel_v4_agg(float4* __restrict__ Y, const int* __restrict__ head,
                                                                          const int* __restrict__ tail, const float* __restrict__ eps_arr, int slots,
                                                                          int n, int twok, int epoch, float a, float b, float gamma, float alpha,
                                                                          float nsr, uint64_t seed) {
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
  float4* Yp = Y + (size_t)p * n;
  for (int q0 = 0; q0 < count; q0 += 32) {   // (warp uniform: every lane takes part in the shuffles)
    const int qi = q0 + lane;
    const bool act = qi < count;
    int own = -1 - lane;                     // vertex of the slot block (distinct negative keys for idle lanes)
    float own_upd[3] = {0.f, 0.f, 0.f};
    if (act) {
      const int e = queue[qi];
      const float eps = eps_arr[(size_t)p * slots + e];
      const int q = (int)floorf((float)epoch / eps);
      const int j = head[(size_t)p * slots + e], kk = tail[(size_t)p * slots + e];
      const float epsn = eps / nsr;
      int tot = (int)floorf((float)epoch / epsn) - 1;
      if (q > 1) {
        const int prev = (int)ceilf((float)(q - 1) * eps);
        tot -= (int)floorf((float)prev / epsn) - 1;
      }
      float4 n4[kSgdNegBatch];
#pragma unroll
      for (int u = 0; u < kSgdNegBatch; ++u) {
        if (u < tot) {
          const uint32_t r = mix32(seed ^ ((uint64_t)p << 52) ^ ((uint64_t)e << 20) ^ ((uint64_t)epoch << 4) ^ (uint64_t)u ^ ((uint64_t)u << 40));
          n4[u] = __ldcg(Yp + (int)(r % (uint32_t)n));
        }
      }
      const float4 c4 = __ldcg(Yp + j), o4 = __ldcg(Yp + kk);
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
      auto repel = [&](const float4& nn) {
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
      for (int sidx = kSgdNegBatch; sidx < tot; ++sidx) {
        const uint32_t r = mix32(seed ^ ((uint64_t)p << 52) ^ ((uint64_t)e << 20) ^ ((uint64_t)epoch << 4) ^ (uint64_t)sidx ^ ((uint64_t)sidx << 40));
        repel(__ldcg(Yp + (int)(r % (uint32_t)n)));
      }
      // the block's own vertex collects in the warp, the far endpoint is updated directly
      const int blk = e / twok;
      if (j == blk) {
        own = blk;
#pragma unroll
        for (int d = 0; d < 3; ++d) own_upd[d] = delta[d];
        atomicAdd(Yp + kk, make_float4(dt[0], dt[1], dt[2], 0.f));
      } else if (kk == blk) {
        own = blk;
#pragma unroll
        for (int d = 0; d < 3; ++d) own_upd[d] = dt[d];
        atomicAdd(Yp + j, make_float4(delta[0], delta[1], delta[2], 0.f));
      } else {   // (not produced by fuzzy_kernel; handled like the plain kernel)
        atomicAdd(Yp + kk, make_float4(dt[0], dt[1], dt[2], 0.f));
        atomicAdd(Yp + j, make_float4(delta[0], delta[1], delta[2], 0.f));
      }
    }
    // segmented inclusive sum over equal keys (contiguous lanes), last lane of a segment writes
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int ko = __shfl_up_sync(0xffffffffu, own, off);
      const float v0 = __shfl_up_sync(0xffffffffu, own_upd[0], off);
      const float v1 = __shfl_up_sync(0xffffffffu, own_upd[1], off);
      const float v2 = __shfl_up_sync(0xffffffffu, own_upd[2], off);
      if (lane >= off && ko == own) { own_upd[0] += v0; own_upd[1] += v1; own_upd[2] += v2; }
    }
    const int knext = __shfl_down_sync(0xffffffffu, own, 1);
    if (own >= 0 && (lane == 31 || knext != own)) atomicAdd(Yp + own, make_float4(own_upd[0], own_upd[1], own_upd[2], 0.f));
  }
}

// EXPERIMENT (off unless TDA_SGD_CLOUD=1): one CTA per cloud, the whole optimisation in ONE launch -- the embedding (16 B
// per point) lives in shared memory, updates are shared-memory float atomics, epochs are separated by __syncthreads(), and
// nothing but the read-only slot table (eps, head, tail) leaves the SM.  Same schedule, RNG keys and update rule as the
// per-epoch kernel (tests/test_umap_gpu.py: trustworthiness 0.99315 vs 0.99313 on 8 clouds).  Measured on the C3 sweep: 64 ms
// per 16 clouds x 500 epochs against 15.5 ms for the per-epoch kernel -- a cloud fires ~21 k edges per epoch at ~300
// instructions each (six __powf), which is ~50 us of issue time on ONE SM, while the per-epoch kernel spreads the same work
// over all 148 SMs and pays L2 atomics instead (shared-memory float add is a CAS loop, ATOMS.CAST.SPIN, on sm_100a).
// The design that would win is a cluster of 8 CTAs per cloud with the embedding in distributed shared memory.
constexpr int kSgdCloudThreads = 1024;
constexpr int kSgdCloudWarps = kSgdCloudThreads / 32;
constexpr int kSgdCloudMinBatch = 8;
constexpr size_t kSgdCloudMaxBytes = 160 * 1024;
__global__ void __launch_bounds__(kSgdCloudThreads, 1) sgd_cloud_kernel(float4* __restrict__ Y4, const int* __restrict__ head, const int* __restrict__ tail,
                                                                        const float* __restrict__ eps_arr, int slots, int n, int n_epochs, float a,
                                                                        float b, float gamma, float alpha0, float nsr, uint64_t seed) {
  extern __shared__ float4 s_y[];                       // [n] the cloud's embedding
  __shared__ int s_queue[kSgdCloudWarps][kSgdSlotsPerLane * 32];
  const int p = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* Yg = Y4 + (size_t)p * n;
  for (int i = threadIdx.x; i < n; i += kSgdCloudThreads) s_y[i] = Yg[i];
  __syncthreads();
  const float* ep_p = eps_arr + (size_t)p * slots;
  const int* hd_p = head + (size_t)p * slots;
  const int* tl_p = tail + (size_t)p * slots;
  int* queue = s_queue[warp];
  constexpr int kChunk = kSgdSlotsPerLane * 32;
  for (int epoch = 0; epoch < n_epochs; ++epoch) {
    const float alpha = epoch == 0 ? alpha0 : alpha0 * (1.f - (float)(epoch - 1) / (float)n_epochs);
    for (int base = warp * kChunk; base < slots; base += kSgdCloudWarps * kChunk) {
      int count = 0;
#pragma unroll
      for (int i = 0; i < kSgdSlotsPerLane; ++i) {
        const int e = base + i * 32 + lane;
        bool fire = false;
        if (e < slots) {
          const float eps = __ldg(&ep_p[e]);
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
        const float eps = __ldg(&ep_p[e]);
        const int q = (int)floorf((float)epoch / eps);
        const int j = __ldg(&hd_p[e]), kk = __ldg(&tl_p[e]);
        const float epsn = eps / nsr;
        int tot = (int)floorf((float)epoch / epsn) - 1;
        if (q > 1) {
          const int prev = (int)ceilf((float)(q - 1) * eps);
          tot -= (int)floorf((float)prev / epsn) - 1;
        }
        const float4 c4 = s_y[j], o4 = s_y[kk];
        float cur[3] = {c4.x, c4.y, c4.z};
        const float oth[3] = {o4.x, o4.y, o4.z};
        float delta[3] = {0.f, 0.f, 0.f};
        float d2 = 0.f;
#pragma unroll
        for (int d = 0; d < 3; ++d) { const float t = cur[d] - oth[d]; d2 += t * t; }
        float g = 0.f;
        if (d2 > 0.f) {
          const float pw = __powf(d2, b - 1.f);
          g = (-2.f * a * b * pw) / (a * pw * d2 + 1.f);
        }
        float* yt = reinterpret_cast<float*>(&s_y[kk]);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float gd = clip4(g * (cur[d] - oth[d])) * alpha;
          cur[d] += gd; delta[d] += gd;
          atomicAdd(&yt[d], -gd);
        }
        for (int sidx = 0; sidx < tot; ++sidx) {
          const uint32_t r = mix32(seed ^ ((uint64_t)p << 52) ^ ((uint64_t)e << 20) ^ ((uint64_t)epoch << 4) ^ (uint64_t)sidx ^ ((uint64_t)sidx << 40));
          const float4 nn = s_y[(int)(r % (uint32_t)n)];
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
        }
        float* yh = reinterpret_cast<float*>(&s_y[j]);
#pragma unroll
        for (int d = 0; d < 3; ++d) atomicAdd(&yh[d], delta[d]);
      }
      __syncwarp();   // the queue is reused by the next chunk
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += kSgdCloudThreads) Yg[i] = s_y[i];
}

// ------------------------------------------------------------------------------------------------
// initialisation helpers
__device__ __forceinline__ float u01(uint64_t key) { return ((float)(mix32(key) >> 8) + 0.5f) * (1.f / 16777216.f); }

__global__ void init_random_kernel(float* __restrict__ Y, int total, float lo, float hi, uint64_t seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) Y[i] = lo + (hi - lo) * u01(seed * 0x9E3779B97F4A7C15ull + (uint64_t)i);
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
    const uint64_t key = seed * 0xD6E8FEB86659FD93ull + ((uint64_t)p << 40) + (uint64_t)i * 2;
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
  const size_t np_h = (size_t)batch * n_head, np_t = move_other ? 0 : (size_t)batch * n_tail;
  if (dim == 3 && ws && ws_bytes >= sizeof(float4) * (np_h + np_t) && (((uintptr_t)ws) & 15) == 0 && n_epochs > 0) {
    float4* Yh4 = (float4*)ws;
    float4* Yt4 = move_other ? Yh4 : Yh4 + np_h;
    pack4_kernel<<<(unsigned)((np_h + 255) / 256), 256, 0, stream>>>(Y, Yh4, np_h);
    // TDA_SGD_CLOUD=1: one CTA per cloud with the embedding in shared memory, all epochs in one launch (see sgd_cloud_kernel)
    const int sgd_mode = (int)option("sgd_mode");
    const bool cloud_mode = sgd_mode == 1;
    const size_t cloud_bytes = sizeof(float4) * (size_t)n_head;
    if (cloud_mode && move_other && n_head == n_tail && batch >= kSgdCloudMinBatch && cloud_bytes <= kSgdCloudMaxBytes) {
      TDA_CUDA_CHECK(cudaFuncSetAttribute(sgd_cloud_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cloud_bytes));
      sgd_cloud_kernel<<<batch, kSgdCloudThreads, cloud_bytes, stream>>>(Yh4, head, tail, eps, slots, n_head, n_epochs, a, b, gamma, alpha0,
                                                                         negative_sample_rate, seed);
      unpack4_kernel<<<(unsigned)((np_h + 255) / 256), 256, 0, stream>>>(Yh4, Y, np_h);
      count_launch(3);
      TDA_LAUNCH_CHECK();
      return TDA_OK;
    }
    if (!move_other) pack4_kernel<<<(unsigned)((np_t + 255) / 256), 256, 0, stream>>>(Y_other, Yt4, np_t);
    const bool agg_mode = sgd_mode == 2 && move_other && n_head == n_tail && slots % n_head == 0 && (slots / n_head) % 2 == 0;
    for (int ep = 0; ep < n_epochs; ++ep) {
      const float alpha = ep == 0 ? alpha0 : alpha0 * (1.f - (float)(ep - 1) / (float)n_epochs);
      const int per_block = kSgdWarps * kSgdSlotsPerLane * 32;
      dim3 g4((slots + per_block - 1) / per_block, batch);
      if (agg_mode) {
        sgd_epoch_kernel_v4_agg<<<g4, kSgdWarps * 32, 0, stream>>>(Yh4, head, tail, eps, slots, n_head, slots / n_head, ep, a, b, gamma, alpha,
                                                                   negative_sample_rate, seed);
        continue;
      }
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
  init_random_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(Y, total, lo, hi, seed);
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
