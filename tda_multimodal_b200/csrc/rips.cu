// rips.cu -- Vietoris-Rips persistence (H0, H1) as a GPU pipeline, batched over independent clouds.
//
// Replaces ripser(X, maxdim=1)['dgms'] (reference: debug_tda_pipeline.py:109-110,
// analyze_tda_over_layers.py:76, analyze_adversarial_tda.py:100-101).  Design (DESIGN.md section "Rips"):
//
//   1. edge filtration sort   : (float32 distance bits, edge index desc) radix sort -> every edge gets an
//                               integer RANK; all later stages are pure integer work on the rank matrix.
//   2. H0                     : Boruvka MST on the rank matrix (one CTA per cloud).
//   3. apparent pairs         : one warp per edge finds its "apex" (largest vertex in the edge's lune);
//                               every non-MST edge with an apex is a zero-persistence apparent pair whose
//                               coboundary column never needs reducing.
//   4. residual reduction     : only edges with an EMPTY lune (relative-neighbourhood-graph edges that are
//                               not in the MST) are reduced, by implicit persistent cohomology over Z/2.
//                               Default: rips_sweep2.cuh (apparent-pair additions as a forward substitution by
//                               rank, windows verified against the cocycle condition); older reducers kept for
//                               comparison: the row sweep with a resolver warp / per-chunk verify (Sweeper) and
//                               the key bitset (Reducer<1>, any n; Reducer<2> is the H2 reducer).  Pivots owned by
//                               reduced columns are found through a hash map.
//
// A triangle is keyed by (rank of its longest edge, opposite vertex): key = M*n + (n-1-w).  Ascending key
// order refines (diameter asc) and puts faces before cofaces, so it is a valid simplex-wise filtration; the
// persistence diagram (as a multiset of (birth,death) values) does not depend on how ties are broken.
#include "common.cuh"
#include "launch_count.cuh"
#include "../../include/tda_b200.h"
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>
#include <vector>
#include <cstdlib>
#include <cstring>
#include <cmath>

namespace tda {
namespace rips {

constexpr int kReduceThreads = 256;
constexpr int kRankDiag = 0x7fffffff;

// stats slots
enum { ST_COLUMNS = 0, ST_APPARENT, ST_REDUCED, ST_ADDITIONS, ST_PUSHES, ST_POPS, ST_EXTENSIONS, ST_MAXV,
       ST_CYC_EXTRACT, ST_CYC_OWNER, ST_CYC_GEN, ST_CYC_BADD, ST_CYC_EXT, ST_CYC_FINAL, ST_BADD_EDGES, ST_EXT_EDGES,
       ST_S2_ROUNDS, ST_S2_LATE, ST_S2_PM, ST_S2_DENSE, ST_S2_SPURIOUS, ST_S2_UNDONE, ST_SPARE0, ST_SPARE1, ST_N };
static_assert(ST_N == TDA_RIPS_STATS, "stats slots");

// ------------------------------------------------------------------------------------------------
// low-dimensional euclidean distance matrix (ripser.py front end)
__global__ void pdist_lowdim_kernel(const float* __restrict__ pts, int n, int d, float* __restrict__ dm) {
  int p = blockIdx.z;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= n || j >= n) return;
  const float* a = pts + ((size_t)p * n + i) * d;
  const float* b = pts + ((size_t)p * n + j) * d;
  double acc = 0.0;
  for (int c = 0; c < d; ++c) {
    double df = (double)a[c] - (double)b[c];
    acc += df * df;  // same evaluation order as the oracle (plain left-to-right sum, separate mul and add)
  }
  dm[((size_t)p * n + i) * n + j] = sqrtf((float)acc);
}

// ------------------------------------------------------------------------------------------------
// enclosing radius: thresh[p] = min_i max_j dm[p][i][j]  (float bits are monotone for d >= 0)
__global__ void enclosing_init_kernel(uint32_t* thresh_bits, int batch, float user_thresh) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < batch) thresh_bits[p] = isinf(user_thresh) ? 0x7f800000u : __float_as_uint(user_thresh);
}
__global__ void enclosing_kernel(const float* __restrict__ dm, int n, uint32_t* thresh_bits) {
  int p = blockIdx.y;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* row = dm + ((size_t)p * n + warp) * n;
  float m = 0.f;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, row[j]);
  m = warp_max_f32(m);
  if (lane == 0) atomicMin(&thresh_bits[p], __float_as_uint(m));
}

// ------------------------------------------------------------------------------------------------
// edge keys: composite 64-bit key (problem << 32 | distance bits), payload = edge index; laid out with
// the edge index DESCENDING so that a stable sort yields (distance asc, index desc) = ripser's order.
__global__ void edge_keys_kernel(const float* __restrict__ dm, int n, int64_t E, const uint32_t* __restrict__ thresh_bits,
                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int* __restrict__ T) {
  int p = blockIdx.y;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int valid = 0;
  if (t < E) {
    int64_t idx = E - 1 - t;
    int i, j;
    edge_vertices(idx, i, j);
    uint32_t bits = __float_as_uint(dm[((size_t)p * n + i) * n + j]);
    valid = bits <= thresh_bits[p];
    keys[(size_t)p * E + t] = ((uint64_t)p << 32) | (valid ? bits : 0xffffffffu);
    vals[(size_t)p * E + t] = (uint32_t)idx;
  }
  int c = __syncthreads_count(valid);
  if (threadIdx.x == 0 && c) atomicAdd(&T[p], c);
}

// rank matrix + per-rank tables
__global__ void rank_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int n, int64_t E,
                                    int* __restrict__ rank, uint32_t* __restrict__ ends, float* __restrict__ sdist) {
  int p = blockIdx.y;
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= E) return;
  uint32_t idx = vals[(size_t)p * E + s];
  int i, j;
  edge_vertices((int64_t)idx, i, j);
  int* R = rank + (size_t)p * n * n;
  R[(size_t)i * n + j] = (int)s;
  R[(size_t)j * n + i] = (int)s;
  ends[(size_t)p * E + s] = ((uint32_t)i << 16) | (uint32_t)j;
  sdist[(size_t)p * E + s] = __uint_as_float((uint32_t)(keys[(size_t)p * E + s] & 0xffffffffu));
  if (s < n) R[(size_t)s * n + s] = kRankDiag;
}
__global__ void rank_diag_kernel(int n, int* __restrict__ rank) {  // n == 1 or E < n corner cases
  int p = blockIdx.y;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) rank[(size_t)p * n * n + (size_t)s * n + s] = kRankDiag;
}

// ------------------------------------------------------------------------------------------------
// Subsets of a cloud whose edges are already in filtration order (bootstrap resamples, config C4): a subset given by ASCENDING
// parent indices keeps the order of its edges -- equal lengths are the same float bits (the distance of two points does not
// depend on the cloud they sit in) and the tie-break, the edge index C(i,2)+j, is monotone under an order-preserving relabelling
// -- so the rank of a subset edge is the number of subset edges before it in the parent's order: a flag + prefix count over the
// parent's sorted edge list replaces pairwise distances, key generation and the radix sort of every resample.
constexpr int kSubThreads = 256;
constexpr int kSubPerThread = 32;
constexpr int kSubChunk = kSubThreads * kSubPerThread;   // parent edges per CTA
__global__ void sorted_edges_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t E,
                                    uint32_t* __restrict__ ends, float* __restrict__ sdist) {
  const int p = blockIdx.y;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= E) return;
  int i, j;
  edge_vertices((int64_t)vals[(size_t)p * E + s], i, j);
  ends[(size_t)p * E + s] = ((uint32_t)i << 16) | (uint32_t)j;
  sdist[(size_t)p * E + s] = __uint_as_float((uint32_t)(keys[(size_t)p * E + s] & 0xffffffffu));
}
// enclosing radius of a subset from the parent's distance matrix: one warp per (subset, point)
__global__ void subset_enclosing_kernel(const float* __restrict__ dm, int n_parent, const int32_t* __restrict__ idx, int m,
                                        uint32_t* __restrict__ thresh_bits) {
  const int p = blockIdx.y;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= m) return;
  const int32_t* I = idx + (size_t)p * m;
  const float* row = dm + (size_t)I[warp] * n_parent;
  float mx = 0.f;
  for (int j = lane; j < m; j += 32) mx = fmaxf(mx, __ldg(&row[I[j]]));
  mx = warp_max_f32(mx);
  if (lane == 0) atomicMin(&thresh_bits[p], __float_as_uint(mx));
}
// parent vertex -> index inside the subset (0xffff: not in it), in shared memory: two look-ups per parent edge, random over the
// table -- from a global table every warp load touched up to 32 cache lines and the L1 throughput bounded both passes
__device__ __forceinline__ void subset_build_map(uint16_t* smap, const int32_t* __restrict__ I, int m, int n_parent) {
  uint32_t* w = reinterpret_cast<uint32_t*>(smap);
  for (int i = threadIdx.x; i < (n_parent + 1) / 2; i += blockDim.x) w[i] = 0xffffffffu;
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const int v = I[i];
    if (v >= 0 && v < n_parent) smap[v] = (uint16_t)i;
  }
  __syncthreads();
}
// subset edges per chunk of the parent's list, and the number of subset edges within the threshold
__global__ void __launch_bounds__(kSubThreads) subset_count_kernel(const uint32_t* __restrict__ pends, const float* __restrict__ psdist, int64_t Ep,
                                                                   const int32_t* __restrict__ idx, int m, int n_parent,
                                                                   const uint32_t* __restrict__ thresh_bits, int* __restrict__ chunk_cnt,
                                                                   int nchunks, int* __restrict__ T) {
  extern __shared__ uint16_t s_submap[];
  __shared__ int s_cnt, s_valid;
  const int p = blockIdx.y, c = blockIdx.x, tid = threadIdx.x;
  const uint32_t tb = thresh_bits[p];
  if (tid == 0) { s_cnt = 0; s_valid = 0; }
  subset_build_map(s_submap, idx + (size_t)p * m, m, n_parent);
  int cnt = 0, valid = 0;
  const int64_t base = (int64_t)c * kSubChunk;
#pragma unroll 2
  for (int k = 0; k < kSubPerThread / 4; ++k) {
    const int64_t r0 = base + ((int64_t)k * kSubThreads + tid) * 4;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r0 + u < Ep) {
        const uint32_t e = __ldg(&pends[r0 + u]);
        if (s_submap[e >> 16] != 0xffffu && s_submap[e & 0xffffu] != 0xffffu) {
          ++cnt;
          valid += __float_as_uint(__ldg(&psdist[r0 + u])) <= tb ? 1 : 0;
        }
      }
    }
  }
  cnt = warp_sum_i32(cnt); valid = warp_sum_i32(valid);
  if ((tid & 31) == 0) { if (cnt) atomicAdd(&s_cnt, cnt); if (valid) atomicAdd(&s_valid, valid); }
  __syncthreads();
  if (tid == 0) {
    chunk_cnt[(size_t)p * nchunks + c] = s_cnt;
    if (s_valid) atomicAdd(&T[p], s_valid);
  }
}
// exclusive scan of the chunk counts of one subset (one CTA per subset, in place)
__global__ void __launch_bounds__(512) subset_scan_kernel(int* __restrict__ chunk_cnt, int nchunks) {
  __shared__ int s_part[16];
  const int p = blockIdx.x, tid = threadIdx.x;
  int* a = chunk_cnt + (size_t)p * nchunks;
  const int per = (nchunks + 511) / 512;
  const int lo = min(tid * per, nchunks), hi = min(lo + per, nchunks);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += a[i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
  if ((tid & 31) == 31) s_part[tid >> 5] = incl;
  __syncthreads();
  int base = incl - sum;
  for (int w = 0; w < (tid >> 5); ++w) base += s_part[w];
  for (int i = lo; i < hi; ++i) { const int t = a[i]; a[i] = base; base += t; }
}
// ranks, end points and lengths of the subset edges (the same tables rank_scatter_kernel writes from a sorted key list)
__global__ void __launch_bounds__(kSubThreads) subset_scatter_kernel(const uint32_t* __restrict__ pends, const float* __restrict__ psdist, int64_t Ep,
                                                                     const int32_t* __restrict__ idx, int n_parent,
                                                                     const int* __restrict__ chunk_off, int nchunks, int m, int64_t E,
                                                                     int* __restrict__ rank, uint32_t* __restrict__ ends, float* __restrict__ sdist) {
  extern __shared__ uint16_t s_submap[];
  __shared__ int s_warp[2][kSubThreads / 32];
  const int p = blockIdx.y, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  subset_build_map(s_submap, idx + (size_t)p * m, m, n_parent);
  int* R = rank + (size_t)p * m * m;
  int run = chunk_off[(size_t)p * nchunks + c];
  const int64_t base = (int64_t)c * kSubChunk;
  // four CONSECUTIVE parent edges per thread and trip (positions ascend with the thread, then inside the thread): one warp scan and
  // one barrier per 1024 edges instead of per 256 -- at one edge per thread the prefix bookkeeping was 125 instructions per
  // 32 edges and the kernel was bound by instruction issue, not by its 3.6 GB of traffic (profiles/r02n_c4_raw.csv)
  for (int k = 0; k < kSubPerThread / 4; ++k) {
    const int64_t r0 = base + ((int64_t)k * kSubThreads + tid) * 4;
    uint32_t vi[4], vj[4];
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vi[u] = 0xffffu; vj[u] = 0xffffu;
      if (r0 + u < Ep) {
        const uint32_t e = __ldg(&pends[r0 + u]);
        vi[u] = s_submap[e >> 16];
        vj[u] = s_submap[e & 0xffffu];
      }
      cnt += (vi[u] != 0xffffu && vj[u] != 0xffffu) ? 1 : 0;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[k & 1][warp] = incl;   // (double buffered: one barrier per trip)
    __syncthreads();
    int off = run + incl - cnt, tot = 0;
#pragma unroll
    for (int w = 0; w < kSubThreads / 32; ++w) { const int t = s_warp[k & 1][w]; if (w < warp) off += t; tot += t; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (vi[u] != 0xffffu && vj[u] != 0xffffu) {
        const int64_t sr = off++;
        if (sr < E) {   // (always, for distinct indices inside the parent)
          R[(size_t)vi[u] * m + vj[u]] = (int)sr;
          R[(size_t)vj[u] * m + vi[u]] = (int)sr;
          ends[(size_t)p * E + sr] = (vi[u] << 16) | vj[u];   // i > j: the relabelling preserves the order
          sdist[(size_t)p * E + sr] = __ldg(&psdist[r0 + u]);
        }
      }
    }
    run += tot;
  }
}

// ------------------------------------------------------------------------------------------------
// H0: Boruvka MST on the rank matrix.  Ranks are distinct, so the minimum spanning forest is unique and
// equals the set of merging edges of ripser's union-find sweep.  Per round: a grid-wide scan kernel (one
// warp per matrix row: smallest rank leaving the row's component -> atomicMin per component) and a
// one-CTA-per-cloud merge kernel (hook, break 2-cycles, pointer jumping, relabel).
constexpr uint32_t kNoEdge = 0xffffffffu;
__global__ void boruvka_init_kernel(int n, uint32_t* __restrict__ comp, uint32_t* __restrict__ cbest, int* __restrict__ done,
                                    int* __restrict__ mstcount) {
  int p = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { comp[(size_t)p * n + i] = i; cbest[(size_t)p * n + i] = kNoEdge; }
  if (i == 0) { done[p] = 0; mstcount[p] = 0; }
}
__global__ void __launch_bounds__(256) boruvka_scan_kernel(const int* __restrict__ rank, const int* __restrict__ Tarr, int n,
                                                           const uint32_t* __restrict__ comp_g, uint32_t* __restrict__ cbest_g,
                                                           const int* __restrict__ done) {
  const int p = blockIdx.y;
  if (done[p]) return;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const int T = Tarr[p];
  const uint32_t* comp = comp_g + (size_t)p * n;
  const int* row = rank + ((size_t)p * n + i) * n;
  const uint32_t ci = comp[i];
  uint32_t best = kNoEdge;
  for (int j = lane; j < n; j += 32) {
    int r = row[j];
    if (r < T && comp[j] != ci) best = min(best, (uint32_t)r);
  }
  best = warp_min_u32(best);
  if (lane == 0 && best != kNoEdge) atomicMin(&cbest_g[(size_t)p * n + ci], best);
}
__global__ void __launch_bounds__(1024) boruvka_merge_kernel(const uint32_t* __restrict__ ends, int n, int64_t E, uint32_t* __restrict__ comp_g,
                                                             uint32_t* __restrict__ parent_g, uint32_t* __restrict__ cbest_g,
                                                             uint8_t* __restrict__ mst, int* __restrict__ done,
                                                             int* __restrict__ mstlist_g, int mst_stride, int* __restrict__ mstcount) {
  const int p = blockIdx.x;
  if (done[p]) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  uint32_t* comp = comp_g + (size_t)p * n;
  uint32_t* parent = parent_g + (size_t)p * n;
  uint32_t* cbest = cbest_g + (size_t)p * n;
  const uint32_t* EN = ends + (size_t)p * E;
  uint8_t* M = mst + (size_t)p * E;
  int merged = 0;
  for (int c = tid; c < n; c += nt) {
    uint32_t r = cbest[c];
    uint32_t par = c;
    if (r != kNoEdge) {  // only component roots ever receive a candidate
      uint32_t e = EN[r];
      uint32_t u = e >> 16, v = e & 0xffffu;
      par = comp[u] == (uint32_t)c ? comp[v] : comp[u];
      M[r] = 1;  // both sides may pick the same edge: same value written twice
      // ... but it enters the MST list once: two components that picked each other did so through this edge; the smaller id lists it
      if (!(cbest[par] == r && par < (uint32_t)c)) {
        const int pos = atomicAdd(&mstcount[p], 1);
        if (pos < n) mstlist_g[(size_t)p * mst_stride + pos] = (int)r;
      }
      merged = 1;
    }
    parent[c] = par;
  }
  if (!__syncthreads_or(merged)) {
    if (tid == 0) done[p] = 1;
    return;
  }
  // two components that chose each other did so through the same edge: the smaller id becomes the root
  for (int c = tid; c < n; c += nt) {
    uint32_t q = parent[c];
    if (q != (uint32_t)c && parent[q] == (uint32_t)c && (uint32_t)c < q) parent[c] = c;
  }
  __syncthreads();
  for (int it = 0; it < 32; ++it) {  // pointer jumping
    int changed = 0;
    for (int c = tid; c < n; c += nt) {
      uint32_t q = parent[c], g = parent[q];
      if (g != q) { parent[c] = g; changed = 1; }
    }
    if (!__syncthreads_or(changed)) break;
  }
  for (int i = tid; i < n; i += nt) { comp[i] = parent[comp[i]]; cbest[i] = kNoEdge; }
}

// The same forest from the SORTED EDGE LIST, one CTA per cloud, one launch (default for n <= 16384): the list is taken in chunks of
// kBorChunk ranks; on a chunk, Boruvka rounds (lowest rank of the chunk leaving each component -> hook -> pointer jumping, all in
// shared memory) repeat until no edge of the chunk joins two components, then the next chunk follows.  Earlier chunks hold no
// joining edge any more, so the lowest joining edge of a component inside the current chunk is its lowest joining edge overall:
// every hook is an edge of the (unique) minimum spanning forest.  For a point cloud nearly all of the forest lies in the first
// chunk; the few long edges between clusters are found by one cheap pass per later chunk (Kruskal by chunks; the chunks double
// in length while they are clean).  Replaces ~25 launches
// that each re-read the whole rank matrix: H0 of 256 x 1000 points 8-13 ms -> well under 1 ms, and one launch instead of 25 in
// the sweep, where every small launch queues behind the other groups' kernels.
constexpr int kBorChunk = 32768;
__global__ void __launch_bounds__(1024) boruvka_chunked_kernel(const uint32_t* __restrict__ ends, const int* __restrict__ Tarr, int n, int64_t E,
                                                               uint32_t* __restrict__ comp_g, uint8_t* __restrict__ mst,
                                                               int* __restrict__ mstlist_g, int mst_stride, int* __restrict__ mstcount) {
  extern __shared__ uint32_t s_bor[];   // comp[n], cbest[n], parent[n]
  uint32_t* comp = s_bor;
  uint32_t* cbest = comp + n;
  uint32_t* parent = cbest + n;
  __shared__ int s_count;
  const int p = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const uint32_t* EN = ends + (size_t)p * E;
  uint8_t* M = mst + (size_t)p * E;
  int* list = mstlist_g + (size_t)p * mst_stride;
  const int T = Tarr[p];
  for (int i = tid; i < n; i += nt) { comp[i] = (uint32_t)i; cbest[i] = kNoEdge; }
  if (tid == 0) s_count = 0;
  __syncthreads();
  int c0 = 0, len = kBorChunk;
  while (c0 < T && s_count < n - 1) {     // (s_count only changes between barriers: every thread sees the same value here)
    const int c1 = (int)min((long long)T, (long long)c0 + len);
    int any = 0;
    for (int e0 = c0 + tid; e0 < c1; e0 += nt * 8) {   // eight edges per thread in flight: the pass is bound by the latency of its loads
      uint32_t en[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int e = e0 + u * nt; en[u] = e < c1 ? __ldg(&EN[e]) : 0u; }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = e0 + u * nt;
        if (e < c1) {
          const uint32_t cu = comp[en[u] >> 16], cv = comp[en[u] & 0xffffu];
          if (cu != cv) { atomicMin(&cbest[cu], (uint32_t)e); atomicMin(&cbest[cv], (uint32_t)e); any = 1; }
        }
      }
    }
    if (!__syncthreads_or(any)) {   // no joining edge left in this chunk: on to the next one, twice as long (few components are
      c0 = c1;                      // left once a chunk is clean, and what joins them may lie far up the list)
      if (len < (1 << 20)) len <<= 1;
      continue;
    }
    for (int c = tid; c < n; c += nt) {
      const uint32_t r = cbest[c];
      uint32_t par = (uint32_t)c;
      if (r != kNoEdge) {   // only component roots ever receive a candidate
        const uint32_t e = __ldg(&EN[r]);
        const uint32_t u = e >> 16, v = e & 0xffffu;
        par = comp[u] == (uint32_t)c ? comp[v] : comp[u];
        M[r] = 1;           // both sides may pick the same edge: same value written twice ...
        if (!(cbest[par] == r && par < (uint32_t)c)) {   // ... but it enters the list once (the smaller id of a mutual pick lists it)
          const int pos = atomicAdd(&s_count, 1);
          if (pos < n) list[pos] = (int)r;
        }
      }
      parent[c] = par;
    }
    __syncthreads();
    for (int c = tid; c < n; c += nt) {   // a mutual pick is a 2-cycle: the smaller id becomes the root
      const uint32_t q = parent[c];
      if (q != (uint32_t)c && parent[q] == (uint32_t)c && (uint32_t)c < q) parent[c] = (uint32_t)c;
    }
    __syncthreads();
    for (int it = 0; it < 32; ++it) {     // pointer jumping
      int changed = 0;
      for (int c = tid; c < n; c += nt) {
        const uint32_t q = parent[c], g = parent[q];
        if (g != q) { parent[c] = g; changed = 1; }
      }
      if (!__syncthreads_or(changed)) break;
    }
    for (int i = tid; i < n; i += nt) { comp[i] = parent[comp[i]]; cbest[i] = kNoEdge; }
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) comp_g[(size_t)p * n + i] = comp[i];
  if (tid == 0) mstcount[p] = s_count;
}

// sort the MST ranks (collected by the merge kernel), emit the H0 rows.  One CTA per cloud; the list is sorted in shared
// memory when it fits (n <= 8192), else in place in global memory (the per-problem list is padded to a power of two).
__global__ void __launch_bounds__(1024) h0_emit_kernel(const uint32_t* __restrict__ ends, const float* __restrict__ sdist,
                                                       const int* __restrict__ Tarr, int n, int64_t E, const int* __restrict__ mstcount,
                                                       const uint32_t* __restrict__ comp_g, float* __restrict__ h0_pairs,
                                                       int64_t* __restrict__ h0_simplex, int32_t* __restrict__ counts, int* __restrict__ mstlist_g,
                                                       int mst_stride, int use_smem) {
  extern __shared__ int s_list[];
  __shared__ int s_zero, s_rows;
  __shared__ int s_wcnt[32];
  const int p = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const uint32_t* EN = ends + (size_t)p * E;
  const uint32_t* comp = comp_g + (size_t)p * n;
  const int T = Tarr[p];
  int* glist = mstlist_g + (size_t)p * mst_stride;
  const int nm = min(mstcount[p], n - 1);
  int np2 = 1;
  while (np2 < nm) np2 <<= 1;
  int* list = use_smem ? s_list : glist;
  if (tid == 0) { s_zero = 0; s_rows = 0; }
  if (use_smem)
    for (int i = tid; i < nm; i += nt) s_list[i] = glist[i];
  // bitonic sort of the list padded to a power of two (the per-problem list holds next_pow2(n) entries)
  for (int i = nm + tid; i < np2; i += nt) list[i] = 0x7fffffff;
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += nt) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int a = list[i], b = list[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { list[i] = b; list[ixj] = a; }
        }
      }
      if (!use_smem) __threadfence_block();
      __syncthreads();
    }
  // rows: (0, d) for d != 0 ascending, then one (0, inf) per remaining component.  Zero-length merging
  // edges have the smallest ranks, so they are a prefix of the sorted list.
  float* out = h0_pairs + (size_t)p * n * 2;
  int64_t* outs = h0_simplex ? h0_simplex + (size_t)p * n * 2 : nullptr;
  int zloc = 0;
  for (int k = tid; k < nm; k += nt)
    if (sdist[(size_t)p * E + list[k]] == 0.f) ++zloc;
  if (zloc) atomicAdd(&s_zero, zloc);
  __syncthreads();
  const int z = s_zero;
  for (int k = z + tid; k < nm; k += nt) {
    const int r = list[k];
    const int row = k - z;
    out[2 * row] = 0.f; out[2 * row + 1] = sdist[(size_t)p * E + r];
    if (outs) {
      const uint32_t e = EN[r];
      outs[2 * row] = -1; outs[2 * row + 1] = edge_index((int)(e >> 16), (int)(e & 0xffffu));
    }
  }
  // one essential class per component, reported at its root vertex, in vertex order (ordered block compaction)
  int rows = nm - z;
  for (int base = 0; base < n; base += nt) {
    const int i = base + tid;
    const bool isroot = i < n && comp[i] == (uint32_t)i;
    const unsigned bal = __ballot_sync(0xffffffffu, isroot);
    if ((tid & 31) == 0) s_wcnt[tid >> 5] = __popc(bal);
    __syncthreads();
    int off = 0, total = 0;
    for (int w = 0; w < (nt >> 5); ++w) {
      const int cw = s_wcnt[w];
      if (w < (tid >> 5)) off += cw;
      total += cw;
    }
    if (isroot) {
      const int row = rows + off + __popc(bal & ((1u << (tid & 31)) - 1));
      if (row < n) {
        out[2 * row] = 0.f; out[2 * row + 1] = INFINITY;
        if (outs) { outs[2 * row] = i; outs[2 * row + 1] = -1; }
      }
    }
    rows += total;
    __syncthreads();
  }
  if (tid == 0) {
    counts[p * 4 + 0] = min(rows, n);
    counts[p * 4 + 2] = T;
  }
}

// ------------------------------------------------------------------------------------------------
// apparent pairs: one warp per edge rank r < T.  apex[r] = largest vertex v with rank(a,v) < r and
// rank(b,v) < r (the first cofacet of the edge in filtration order has the edge as its longest edge),
// -1 if the lune is empty, -2 for MST edges (negative edges are not columns).
__global__ void apparent_kernel(const int* __restrict__ rank, const uint32_t* __restrict__ ends, const uint8_t* __restrict__ mst,
                                const int* __restrict__ Tarr, int n, int64_t E, int* __restrict__ apex, uint2* __restrict__ ea,
                                int* __restrict__ blist, int* __restrict__ bcount, int cap1, unsigned long long* __restrict__ stats) {
  const int p = blockIdx.y;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int T = Tarr[p];
  if (r >= T) return;
  int* A = apex + (size_t)p * E;
  const uint32_t e = ends[(size_t)p * E + r];
  if (mst[(size_t)p * E + r]) {
    if (lane == 0) { A[r] = -2; ea[(size_t)p * E + r] = make_uint2(e, (uint32_t)-2); }
    return;
  }
  const int* ra = rank + (size_t)p * n * n + (size_t)(e >> 16) * n;
  const int* rb = rank + (size_t)p * n * n + (size_t)(e & 0xffffu) * n;
  int found = -1;
  for (int base = ((n - 1) | 31); base >= 31; base -= 32) {  // chunks from the top, lanes descending
    int v = base - lane;
    bool hit = v < n && ra[v] < (int)r && rb[v] < (int)r;
    unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) { found = base - (__ffs(m) - 1); break; }
  }
  if (lane == 0) {
    A[r] = found;
    ea[(size_t)p * E + r] = make_uint2(e, (uint32_t)found);
    if (found < 0) {
      int pos = atomicAdd(&bcount[p], 1);
      if (pos < cap1) blist[(size_t)p * cap1 + pos] = (int)r;
    }
  }
}

// The same by matrix ROW (default for n <= 12288): one CTA per vertex a keeps row a of the rank matrix in shared memory and
// handles every edge (a, b), b < a, one warp per edge.  The a-side of the lune test is then a shared-memory read, the b-side rows
// are fetched only for the lanes (and only in the chunks) that pass the a-side test -- short edges, whose lune test fails in
// almost every chunk, no longer stream their two rank rows from L2.  MST edges always have an empty lune (the longest edge of a
// triangle is not in the MST), so the MST flag is read only for the edges without an apex.
__global__ void __launch_bounds__(256) apparent_rows_kernel(const int* __restrict__ rank, const uint8_t* __restrict__ mst,
                                                            const int* __restrict__ Tarr, int n, int64_t E, uint2* __restrict__ ea,
                                                            int* __restrict__ blist, int* __restrict__ bcount, int cap1) {
  extern __shared__ int s_rowa[];
  const int p = blockIdx.y, a = blockIdx.x;
  const int T = Tarr[p];
  const int* R = rank + (size_t)p * n * n;
  for (int v = threadIdx.x; v < n; v += blockDim.x) s_rowa[v] = R[(size_t)a * n + v];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint2* EAo = ea + (size_t)p * E;
  for (int b = warp; b < a; b += nwarps) {
    const int r = s_rowa[b];
    if (r >= T) continue;                    // beyond the threshold: not in the filtration
    const int* rb = R + (size_t)b * n;
    int found = -1;
    for (int base = ((n - 1) | 31); base >= 31; base -= 32) {   // chunks from the top, lanes descending
      const int v = base - lane;
      const bool ha = v < n && s_rowa[v] < r;
      if (!__any_sync(0xffffffffu, ha)) continue;
      const bool hit = ha && __ldg(&rb[v]) < r;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) { found = base - (__ffs(m) - 1); break; }
    }
    if (lane == 0) {
      if (found < 0 && mst[(size_t)p * E + r]) found = -2;
      EAo[r] = make_uint2(((uint32_t)a << 16) | (uint32_t)b, (uint32_t)found);
      if (found == -1) {
        const int pos = atomicAdd(&bcount[p], 1);
        if (pos < cap1) blist[(size_t)p * cap1 + pos] = r;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// residual reduction
//
// Working column of one edge b with an empty lune = set of triangle keys with odd multiplicity in
// delta(V), V = edges added so far.  One CTA owns one cloud at a time and walks its columns in ripser's
// order.  The column lives in a BITSET over the key space (global memory, one window of `wbits` bits per
// resident CTA): adding the coboundary of an edge is n-2 fire-and-forget atomic XORs (Z/2 cancellation is
// free, keys never move), the next pivot is the next set bit.  A one-bit-per-page summary of the bitset sits
// in shared memory (page = 8192 bits = 1 KB), so the scan skips empty regions and the clean-up after a
// column touches only dirty pages.  The 2^18 keys just ahead of the scan position (the "near" range) are held in
// shared memory instead: the chain of pivots is followed there without a global fence or a global read per step,
// and the range is refilled from the global bitset (and zeroed there) when the scan runs off its end.  If the key
// space exceeds the window (H2 of more than ~300 points: E * n^2 tetrahedron keys), a key beyond the window is appended to the
// FAR BUCKET of the window it belongs to (fixed region per window, counter in shared memory); when the scan runs off the
// window it jumps to the next non-empty bucket and toggles that bucket's keys into the (all-zero) bitset, so every key is
// written once and read once however many windows a column crosses.  Only if a bucket overflows does the column fall back to
// re-enumerating delta(V) for each later window (the first version of this reducer did that for every slide:
// C2 at n = 2000 crossed up to 1.8 k windows per column and did not finish).
// -DTDA_BOUNDS_CHECK: every indexed access of the reducer is checked; the first violations are printed (device printf) and
// the access is skipped, so a bad index can be located without compute-sanitizer
#ifdef TDA_BOUNDS_CHECK
#define TDA_BC(cond, what, a, b)                                                                                              \
  ((cond) ? true : (printf("TDA_BOUNDS %s: %lld %lld (block %d thread %d line %d)\n", what, (long long)(a), (long long)(b), \
                           (int)blockIdx.x, (int)threadIdx.x, __LINE__), false))
#else
#define TDA_BC(cond, what, a, b) true
#endif
constexpr int kPageShift = 13;                 // 8192 bits per page
constexpr int kPageWords = 1 << (kPageShift - 5);  // 256 words = kReduceThreads
static_assert(kPageWords == kReduceThreads, "one thread per word of a page");
constexpr int kNearShift = 18;                 // near range: 2^18 bits = 32 pages = 32 KB of shared memory
constexpr int kNearWords = 1 << (kNearShift - 5);
constexpr uint64_t kNearBits = 1ull << kNearShift;
constexpr uint32_t kScanRows = 8;              // rows of kReduceThreads near words examined per barrier
constexpr uint32_t kMaxFarBuckets = 8192;      // fill counters of the far buckets live in shared memory (32 KB at most)

struct ReduceParams {
  const int* rank; const uint32_t* ends; const float* sdist; const int* T; const uint2* ea;  // ea[r] = (endpoints, apex)
  int* blist; const int* bcount;
  int n; int64_t E; int batch; int cap1;
  float* h1_pairs; int64_t* h1_simplex; int32_t* counts;
  // per-CTA scratch
  uint32_t* bits; uint64_t wbits;           // [grid, wbits/32]  (all zero between columns)   (bitset reducer)
  uint32_t* xmat; int xw;                   // [grid, n, xw]  V as a symmetric bit matrix    (sweep reducer)
  uint32_t* pmat;                           // [grid, n, xw]  adjacency bit matrix of the edges below the cursor
  uint32_t* vbits; int64_t vwords;          // [grid, vwords]
  uint32_t* vlist; int64_t vcap;            // [grid, 2, vcap]
  // per-problem
  uint64_t* hkeys; int* hvals; int hcap;    // [batch, hcap]
  uint32_t* vpool; int64_t vpool_cap;       // [batch, vpool_cap]
  int64_t* vstart; int* vlen;               // [batch, cap1]
  int* work_counter; unsigned long long* stats;  // [batch, ST_N]
  const short* apex4;                       // (H2) [batch, E*n] per triangle key: the vertex of its apparent cofacet, or -1
  // far buckets (key space larger than one window): keys beyond the window wait in the bucket of their window instead of
  // being re-enumerated from V when the window gets there.  [grid, far_cap] keys; nbk = most windows a column can see
  uint64_t* far; uint64_t far_cap; uint32_t nbk;
  int verify_mode;                          // sweep reducer: substitute-then-verify instead of the sequential resolver (opt-in)
  // sweep2 reducer (rips_sweep2.cuh)
  const uint2* par;                         // [batch, E] parents of the apparent edges (rank | apparent << 31)
  uint2* s2_pend; uint2* s2_heavy;          // per-cluster spill lists: [grid, 2, wmax + 64], [grid, wmax + 64]
  uint32_t* s2_rec;                         // per-cluster column records: [grid, s2_nrec, kWcRec]
  int* s2_lists;                            // per-cluster work lists: [grid, 3 * s2_nrec + cap1] (two round lists, cluster-engine list, final record per column)
  int s2_nrec;                              // records per cloud (every reduction of a column writes a new one)
  int s2_warp_engine;                       // 1: short columns are reduced by single warps first (stage A + commit loop)
  int s2_wc_max_rows;                       // rows a warp sweeps before it hands its column to the cluster engine
  int s2_debug;                             // option rips_debug: one line of counters per cloud (device printf)
  int s2_w0, s2_wsparse, s2_wmax, s2_dense_min, s2_dense_div;
};

struct ReduceSmem {
  alignas(16) uint32_t nearw[kNearWords];
  uint32_t bal[2][kReduceThreads / 32];
  uint32_t val[2][kReduceThreads / 32];
  uint32_t vcount, vcount2, vsel;
  int abort_flag, problem;
  int far_overflow;        // a bucket of this column ran out of room: its later windows are re-enumerated from V instead
  unsigned long long toggles;
};

// DIM = 1: columns are edges (by rank), rows triangles, key = rank(longest edge) * n + (n-1-opposite vertex)          (H1)
// DIM = 2: columns are triangles (by that key), rows tetrahedra, key = rank(longest edge) * n^2 + (n-1-p) * n + (n-1-q), p > q
//          the two vertices off the longest edge.  Ascending key order refines the diameter, faces precede cofaces: a valid
//          simplex-wise filtration, so the (birth, death) values of the diagram do not depend on the tie-break.        (H2)
template <int DIM>
struct Reducer {
  static constexpr uint64_t kEmpty = ~0ull;
  const ReduceParams& P;
  ReduceSmem& S;
  uint32_t* s1;            // page summary (dynamic shared memory), s1words words
  const int tid, lane, warp;
  const int* R; const uint32_t* EN; const uint2* EA; int T; int n;
  uint32_t* bits; uint32_t* vbits; uint32_t* vl0;
  uint64_t* hkeys; int* hvals;
  uint64_t wbits; uint32_t npages, s1words;
  uint64_t wbase;          // first key of the window
  uint64_t nbase;          // first relative position of the near range (page aligned)
  uint32_t par;            // parity of the bal/val double buffer
  uint64_t* far; uint32_t* fcnt;   // far buckets of this CTA, their fill counters (dynamic shared memory, after s1)
  uint64_t* foff;                  // start of every bucket's region (dynamic shared memory): regions grow with the window index,
                                   // because a cofacet's longest edge is the max of several ranks (density ~ x^2 over the key space)
  uint64_t wbase0;                 // first window of the current column
  uint32_t cur_b, ncol_b;          // current window / number of windows of the current column
  bool far_on;
  unsigned long long my_toggles;
  unsigned long long n_refills, n_passes;   // diagnostics

  __device__ __forceinline__ Reducer(const ReduceParams& p, ReduceSmem& s, uint32_t* s1_)
      : P(p), S(s), s1(s1_), tid(threadIdx.x), lane(threadIdx.x & 31), warp(threadIdx.x >> 5) {
    n = P.n;
    wbits = P.wbits;
    npages = (uint32_t)(wbits >> kPageShift);
    s1words = (npages + 31) >> 5;
    bits = P.bits + (size_t)blockIdx.x * (wbits >> 5);
    vbits = P.vbits + (size_t)blockIdx.x * P.vwords;
    vl0 = P.vlist + (size_t)blockIdx.x * 2 * P.vcap;
    par = 0;
    my_toggles = 0;
    n_refills = n_passes = 0;
    far = P.far ? P.far + (size_t)blockIdx.x * P.far_cap : nullptr;
    fcnt = s1 + s1words;
    foff = reinterpret_cast<uint64_t*>(s1 + ((s1words + P.nbk + 1) & ~1u));
    wbase0 = 0; cur_b = 0; ncol_b = 1; far_on = false;
  }

  // a key beyond the current window: park it in the bucket of its window
  __device__ __forceinline__ void far_append(uint64_t key) {
    const uint32_t b = (uint32_t)((key - wbase0) / wbits);
    if (!TDA_BC(b > cur_b && b < ncol_b, "far_append bucket/ncol_b", b, ncol_b)) return;
    const uint32_t slot = atomicAdd(&fcnt[b], 1u);
    const uint64_t o0 = foff[b];
    if ((uint64_t)slot < foff[b + 1] - o0) far[o0 + slot] = key;
    else S.far_overflow = 1;
  }

  // smallest thread index whose word is non-zero (and that word), or -1; one barrier per call
  __device__ __forceinline__ int block_first_nonzero(uint32_t w, uint32_t& wout) {
    const unsigned b = __ballot_sync(0xffffffffu, w != 0);
    const uint32_t wfirst = __shfl_sync(0xffffffffu, w, b ? __ffs(b) - 1 : 0);
    if (lane == 0) {
      S.bal[par][warp] = b;
      S.val[par][warp] = wfirst;
    }
    __syncthreads();
    int first = -1;
#pragma unroll
    for (int i = kReduceThreads / 32 - 1; i >= 0; --i)
      if (S.bal[par][i]) { first = i * 32 + __ffs(S.bal[par][i]) - 1; wout = S.val[par][i]; }
    par ^= 1;
    return first;
  }

  // block-wide minimum; one barrier per call
  __device__ __forceinline__ uint32_t block_min(uint32_t v) {
    v = __reduce_min_sync(0xffffffffu, v);
    if (lane == 0) S.bal[par][warp] = v;
    __syncthreads();
    uint32_t m = S.bal[par][0];
#pragma unroll
    for (int i = 1; i < kReduceThreads / 32; ++i) m = min(m, S.bal[par][i]);
    par ^= 1;
    return m;
  }

  // next set bit of the window at relative position >= pos; false when the window holds none
  __device__ __forceinline__ bool scan(uint64_t& pos) {
    for (;;) {
      // near range (shared memory): the word under the cursor first (uniform address, no barrier), then every
      // thread walks its own column of the word array (conflict free) and the block takes the minimum index
      while (pos < nbase + kNearBits) {
        const uint32_t wi = (uint32_t)((pos - nbase) >> 5);
        const uint32_t w0 = S.nearw[wi] & (0xffffffffu << (pos & 31));
        if (w0) { pos = nbase + ((uint64_t)wi << 5) + (uint64_t)(__ffs(w0) - 1); return true; }
        const uint32_t row0 = (wi + 1) / kReduceThreads;
        uint32_t cand = 0xffffffffu;
        for (uint32_t r = row0; r < row0 + kScanRows && r < (uint32_t)(kNearWords / kReduceThreads); ++r) {
          const uint32_t idx = r * kReduceThreads + tid;
          if (cand == 0xffffffffu && idx > wi && S.nearw[idx] != 0) cand = idx;
        }
        cand = block_min(cand);
        ++n_passes;
        if (cand != 0xffffffffu) {
          const uint32_t w = S.nearw[cand];
          pos = nbase + ((uint64_t)cand << 5) + (uint64_t)(__ffs(w) - 1);
          return true;
        }
        pos = nbase + ((uint64_t)min((row0 + kScanRows) * kReduceThreads, (uint32_t)kNearWords) << 5);
      }
      // near range exhausted: move it to the next dirty page of the global bitset
      ++n_refills;
      __threadfence();  // every thread's XORs are performed before anybody reads the bitset
      __syncthreads();
      uint32_t page = (uint32_t)((nbase + kNearBits) >> kPageShift);
      if (page >= npages) return false;
      uint32_t found_page = 0xffffffffu;
      for (uint32_t base = page >> 5; base < s1words; base += kReduceThreads) {
        const uint32_t idx = base + tid;
        uint32_t w = idx < s1words ? s1[idx] : 0u;
        if (idx == (page >> 5)) w &= 0xffffffffu << (page & 31);
        uint32_t wv = 0;
        const int f = block_first_nonzero(w, wv);
        if (f >= 0) { found_page = (base + f) * 32 + (__ffs(wv) - 1); break; }
      }
      if (found_page == 0xffffffffu || found_page >= npages) { nbase = (uint64_t)npages << kPageShift; return false; }
      nbase = (uint64_t)found_page << kPageShift;
      pos = nbase;
      // move [nbase, nbase + 2^18) from the global bitset to shared memory: 8 x 16-byte loads per thread, all in flight
      const uint64_t wwords = wbits >> 5;
      const uint64_t g0 = nbase >> 5;  // multiple of 256 words: 16-byte aligned
      uint4 buf[kNearWords / 4 / kReduceThreads];
#pragma unroll
      for (int j = 0; j < kNearWords / 4 / kReduceThreads; ++j) {
        const uint64_t g = g0 + ((uint64_t)j * kReduceThreads + tid) * 4;
        buf[j] = g < wwords ? __ldcg(reinterpret_cast<const uint4*>(bits + g)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < kNearWords / 4 / kReduceThreads; ++j) {
        const uint32_t idx4 = (uint32_t)j * kReduceThreads + tid;
        const uint64_t g = g0 + (uint64_t)idx4 * 4;
        reinterpret_cast<uint4*>(S.nearw)[idx4] = buf[j];
        if (buf[j].x | buf[j].y | buf[j].z | buf[j].w) *reinterpret_cast<uint4*>(bits + g) = make_uint4(0, 0, 0, 0);
      }
      if (tid < (1 << (kNearShift - kPageShift))) {
        const uint32_t pg = found_page + tid;
        if (pg < npages) atomicAnd(&s1[pg >> 5], ~(1u << (pg & 31)));
      }
      __syncthreads();
    }
  }

  __device__ __forceinline__ void toggle_key(uint64_t key) {
    const uint64_t rel = key - wbase;
    ++my_toggles;
    if (rel - nbase < kNearBits) {  // (rel >= nbase always: keys below the scan position are never generated)
      atomicXor(&S.nearw[(rel - nbase) >> 5], 1u << (rel & 31));
      return;
    }
    if (!TDA_BC(rel < wbits, "toggle_key rel/wbits", rel, wbits)) return;
    atomicXor(&bits[rel >> 5], 1u << (rel & 31));
    const uint32_t page = (uint32_t)(rel >> kPageShift);
    const uint32_t m = 1u << (page & 31);
    if (!(s1[page >> 5] & m)) atomicOr(&s1[page >> 5], m);
  }
  // cofacets of edge `re` with key in [lo, wbase + wbits): vertices strided over `nthr` threads
  __device__ __forceinline__ void gen(int re, uint64_t lo, int t0, int nthr) {
    if constexpr (DIM == 1) gen(re, __ldg(&EN[re]), lo, t0, nthr);
    else gen_tri((uint32_t)re, lo, t0, nthr);
  }
  // cofacets (tetrahedra) of the triangle with key k3 = M * n + (n-1-w)
  __device__ __forceinline__ void gen_tri(uint32_t k3, uint64_t lo, int t0, int nthr) {
    const int M = (int)(k3 / (uint32_t)n);
    const int w = n - 1 - (int)(k3 - (uint32_t)M * (uint32_t)n);
    if (!TDA_BC(M >= 0 && M < T, "gen_tri M/T", M, T)) return;
    const uint32_t e = __ldg(&EN[M]);
    const int x = (int)(e >> 16), y = (int)(e & 0xffffu);
    const int* rowx = R + (size_t)x * n;
    const int* rowy = R + (size_t)y * n;
    const int* roww = R + (size_t)w * n;
    const uint64_t hi = wbase + wbits;
    const uint64_t n2 = (uint64_t)n * (uint64_t)n;
    // the three rank rows are read once each and never again (L2 latency per element): keep kTriLoads vertices per thread in
    // flight -- one warp per triangle walks n/32 vertices, which cost 25 k cycles per triangle with one load group at a time
    constexpr int kTriLoads = 8;
    for (int v0 = t0; v0 < n; v0 += nthr * kTriLoads) {
      int rxs[kTriLoads], rys[kTriLoads], rws[kTriLoads];
#pragma unroll
      for (int u = 0; u < kTriLoads; ++u) {
        const int v = v0 + u * nthr;
        const bool in = v < n;
        rxs[u] = in ? __ldg(&rowx[v]) : kRankDiag;   // (v in {x,y,w}: kRankDiag as well -> skipped below)
        rys[u] = in ? __ldg(&rowy[v]) : kRankDiag;
        rws[u] = in ? __ldg(&roww[v]) : kRankDiag;
      }
#pragma unroll
      for (int u = 0; u < kTriLoads; ++u) {
        const int v = v0 + u * nthr;
        const int rx = rxs[u], ry = rys[u], rw = rws[u];
        const int M4 = max(max(M, rx), max(ry, rw));
        if (M4 >= T) continue;
        int p2, q2;   // the two vertices off the longest edge
        if (M4 == M) { p2 = w; q2 = v; }
        else if (M4 == rx) { p2 = y; q2 = w; }
        else if (M4 == ry) { p2 = x; q2 = w; }
        else { p2 = x; q2 = y; }
        const int hi_v = max(p2, q2), lo_v = min(p2, q2);
        const uint64_t key = (uint64_t)M4 * n2 + (uint64_t)(n - 1 - hi_v) * (uint64_t)n + (uint64_t)(n - 1 - lo_v);
        if (key >= lo) {
          if (key < hi) toggle_key(key);
          else if (far_on) far_append(key);
        }
      }
    }
  }
  __device__ __forceinline__ void gen(int re, uint32_t e, uint64_t lo, int t0, int nthr) {
    const int a = (int)(e >> 16), b = (int)(e & 0xffffu);
    const int* rowa = R + (size_t)a * n;
    const int* rowb = R + (size_t)b * n;
    const uint64_t hi = wbase + wbits;
    for (int v0 = t0; v0 < n; v0 += nthr * 4) {
      int ra[4], rb[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int v = v0 + it * nthr;
        ra[it] = v < n ? __ldg(&rowa[v]) : kRankDiag;
        rb[it] = v < n ? __ldg(&rowb[v]) : kRankDiag;
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int v = v0 + it * nthr;
        const int M = max(re, max(ra[it], rb[it]));
        if (M < T) {
          const int opp = (M == re) ? v : (M == ra[it] ? b : a);
          const uint64_t key = (uint64_t)M * (uint64_t)n + (uint64_t)(n - 1 - opp);
          if (key >= lo) {
            if (key < hi) toggle_key(key);
            else if (far_on) far_append(key);
          }
        }
      }
    }
  }
  // the next scan reads shared memory only (the global XORs are fenced when the near range is refilled)
  // The fence is required: the toggles are fire-and-forget RED operations, and without it a batch of two clouds (two CTAs
  // reducing at the same time) faulted in about one run of five (an illegal address a few columns later); bar.sync alone
  // does not wait for them.  160 repetitions are clean with it (scripts/h2_flaky.py).
  __device__ __forceinline__ void publish() { __threadfence(); __syncthreads(); }

  // ---- V (the reduction column as a set of edges, by rank)
  __device__ __forceinline__ uint32_t* vlist(uint32_t sel) const { return vl0 + (size_t)sel * P.vcap; }
  __device__ __forceinline__ void v_toggle(int re) {  // any single thread; fire-and-forget (no atomic round trip)
    if (!TDA_BC((int64_t)((uint32_t)re >> 5) < P.vwords, "v_toggle word/vwords", (uint32_t)re >> 5, P.vwords)) return;
    atomicXor(&vbits[(uint32_t)re >> 5], 1u << (re & 31));
    const uint32_t pos = atomicAdd(&S.vcount, 1u);  // the list may hold an edge several times: v_compact keeps it once iff its bit is set
    if (pos < (uint32_t)P.vcap) vlist(S.vsel)[pos] = (uint32_t)re;
    else S.abort_flag = TDA_ERR_CAPACITY;
  }
  // compact the list: keep each edge whose bit is set exactly once
  __device__ __forceinline__ void v_compact() {
    __syncthreads();
    const uint32_t nin = min(S.vcount, (uint32_t)P.vcap);
    const uint32_t* src = vlist(S.vsel);
    uint32_t* dst = vlist(S.vsel ^ 1);
    if (tid == 0) S.vcount2 = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < nin; i0 += kReduceThreads) {
      const uint32_t i = i0 + tid;
      bool keep = false;
      uint32_t e = 0;
      if (i < nin) {
        e = src[i];
        const uint32_t m = 1u << (e & 31);
        keep = (atomicAnd(&vbits[e >> 5], ~m) & m) != 0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      uint32_t bs = 0;
      if (lane == 0 && bal) bs = atomicAdd(&S.vcount2, (uint32_t)__popc(bal));
      bs = __shfl_sync(0xffffffffu, bs, 0);
      if (keep) dst[bs + __popc(bal & ((1u << lane) - 1))] = e;
    }
    __syncthreads();
    const uint32_t nout = S.vcount2;
    for (uint32_t i = tid; i < nout; i += kReduceThreads) {
      const uint32_t e = dst[i];
      atomicOr(&vbits[e >> 5], 1u << (e & 31));
    }
    __syncthreads();
    if (tid == 0) { S.vsel ^= 1; S.vcount = nout; }
    __syncthreads();
  }
  __device__ __forceinline__ void v_clear() {  // after v_compact: clear bits, empty list
    const uint32_t nin = S.vcount;
    const uint32_t* src = vlist(S.vsel);
    for (uint32_t i = tid; i < nin; i += kReduceThreads) {
      const uint32_t e = src[i];
      atomicAnd(&vbits[e >> 5], ~(1u << (e & 31)));
    }
    __syncthreads();
    if (tid == 0) S.vcount = 0;
    __syncthreads();
  }

  // ---- pivot hash map: probed by every thread redundantly (uniform addresses -> one transaction, no barrier)
  __device__ __forceinline__ int hash_find(uint64_t key) const {
    uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & (uint32_t)(P.hcap - 1);
    for (;;) {
      const uint64_t k = hkeys[h];
      if (k == key) return hvals[h];
      if (k == kEmpty) return -1;
      h = (h + 1) & (uint32_t)(P.hcap - 1);
    }
  }
  __device__ __forceinline__ void hash_insert(uint64_t key, int val) {  // thread 0
    uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & (uint32_t)(P.hcap - 1);
    while (hkeys[h] != kEmpty) h = (h + 1) & (uint32_t)(P.hcap - 1);
    hkeys[h] = key;
    hvals[h] = val;
  }

  // sort blist[0..nb) ascending in place (bitonic, global memory)
  __device__ __forceinline__ void sort_blist(int* bl, int nb) {   // (triangle keys of DIM 2 are compared as unsigned)
    int np2 = 1;
    while (np2 < nb) np2 <<= 1;
    for (int i = nb + tid; i < np2; i += kReduceThreads) bl[i] = DIM == 1 ? 0x7fffffff : (int)0xffffffffu;  // cap1 is a power of two >= nb
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < np2; i += kReduceThreads) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const int a = bl[i], b = bl[ixj];
            const bool up = ((i & k) == 0);
            const bool gt = DIM == 1 ? (a > b) : ((uint32_t)a > (uint32_t)b);
            if (gt == up) { bl[i] = b; bl[ixj] = a; }
          }
        }
        __syncthreads();
      }
  }

  // add the edges list[0..cnt) (global memory): one warp per edge
  __device__ __forceinline__ void add_edges(const uint32_t* list, uint32_t cnt, uint64_t lo, bool track_v) {
    for (uint32_t i = warp; i < cnt; i += kReduceThreads / 32) {
      const int re = (int)list[i];
      if (track_v && lane == 0) v_toggle(re);
      gen(re, lo, lane, 32);
    }
  }

  // zero the window again after a column: bits at relative position >= pos may be set
  __device__ __forceinline__ void clean_window(uint64_t pos, uint32_t nv) {
    const uint64_t lo = wbase + pos;
    if ((uint64_t)nv * (uint64_t)n <= 32ull * npages) {
      // few keys: toggle them back (V holds every edge exactly once after v_compact)
      add_edges(vlist(S.vsel), nv, lo, false);
      __threadfence();
    } else {
      // many keys: sweep the dirty pages, one warp per page
      const uint32_t p0 = (uint32_t)(min(nbase + kNearBits, wbits) >> kPageShift);
      for (uint32_t wbase_i = (p0 >> 5); wbase_i < s1words; wbase_i += kReduceThreads / 32) {
        const uint32_t wi = wbase_i + warp;
        uint32_t f = wi < s1words ? s1[wi] : 0u;
        while (f) {
          const uint32_t page = wi * 32 + (__ffs(f) - 1);
          f &= f - 1;
          if (!TDA_BC(page < npages, "clean page/npages", page, npages)) continue;
          uint4* dst = reinterpret_cast<uint4*>(bits + (size_t)page * kPageWords);
          dst[lane] = make_uint4(0, 0, 0, 0);
          dst[lane + 32] = make_uint4(0, 0, 0, 0);
        }
      }
      __threadfence();
    }
    __syncthreads();
    for (uint32_t i = tid; i < s1words; i += kReduceThreads) s1[i] = 0;
    for (uint32_t i = tid; i < (uint32_t)kNearWords; i += kReduceThreads) S.nearw[i] = 0;
    __syncthreads();
  }

  __device__ __forceinline__ void run_problem(int p) {
    R = P.rank + (size_t)p * n * n;
    EN = P.ends + (size_t)p * P.E;
    EA = P.ea + (size_t)p * P.E;
    T = P.T[p];
    hkeys = P.hkeys + (size_t)p * P.hcap;
    hvals = P.hvals + (size_t)p * P.hcap;
    const float* SD = P.sdist + (size_t)p * P.E;
    int* bl = P.blist + (size_t)p * P.cap1;
    const int nb = P.bcount[p];
    unsigned long long* st = P.stats + (size_t)p * ST_N;
    if (nb > P.cap1) {
      if (tid == 0) { P.counts[p * 4 + 3] = TDA_ERR_CAPACITY; P.counts[p * 4 + 1] = 0; }
      return;
    }
    sort_blist(bl, nb);
    for (int i = tid; i < P.hcap; i += kReduceThreads) hkeys[i] = kEmpty;
    for (uint32_t i = tid; i < s1words; i += kReduceThreads) s1[i] = 0;
    for (uint32_t i = tid; i < (uint32_t)kNearWords; i += kReduceThreads) S.nearw[i] = 0;
    if (tid == 0) { S.vcount = 0; S.vsel = 0; S.abort_flag = 0; S.toggles = 0; }
    my_toggles = 0;
    n_refills = n_passes = 0;
    __threadfence();
    __syncthreads();
    const uint64_t n2 = (uint64_t)n * (uint64_t)n;
    const short* A4 = DIM == 2 ? P.apex4 + (size_t)p * (size_t)P.E * (size_t)n : nullptr;
    const uint64_t kmax = DIM == 1 ? (uint64_t)T * (uint64_t)n : (uint64_t)T * n2;  // keys are < kmax
    int nrows = 0;
    int64_t vpool_used = 0;
    unsigned long long additions = 0, slides = 0, maxv = 0, pops = 0;
    long long cyc[6] = {0, 0, 0, 0, 0, 0};
    unsigned long long badd_edges = 0, ext_edges = 0, far_keys = 0;
    long long t0;
    float* out = P.h1_pairs + (size_t)p * P.cap1 * 2;
    int64_t* outs = P.h1_simplex ? P.h1_simplex + (size_t)p * P.cap1 * 2 : nullptr;

    for (int ci = nb - 1; ci >= 0; --ci) {
      const int rbirth = bl[ci];   // DIM 1: rank of the birth edge; DIM 2: key of the birth triangle
      // DIM 1: the lune of rbirth is empty, every cofacet key is >= (rbirth+1)*n.  DIM 2: a cofacet's longest edge is >= the triangle's
      const uint64_t first = DIM == 1 ? (uint64_t)(rbirth + 1) * (uint64_t)n : (uint64_t)((uint32_t)rbirth / (uint32_t)n) * n2;
      wbase = first & ~((1ull << kPageShift) - 1);
      nbase = 0;
      uint64_t pos = first - wbase;
      wbase0 = wbase;
      cur_b = 0;
      ncol_b = (uint32_t)((kmax - wbase0 + wbits - 1) / wbits);
      far_on = far != nullptr && ncol_b > 1 && ncol_b <= P.nbk;
      if (far_on) {
        // region of bucket b: 30 % of the room split evenly, 70 % in proportion to x^3 over the part of the key space the
        // column can still reach (x = rank of the longest edge / T)
        const double x0 = (double)wbase0 / (double)kmax, x03 = x0 * x0 * x0, den = 1.0 - x03;
        for (uint32_t b = tid; b <= ncol_b; b += kReduceThreads) {
          double xb = ((double)wbase0 + (double)b * (double)wbits) / (double)kmax;
          if (xb > 1.0) xb = 1.0;
          const double cubic = den > 1e-9 ? (xb * xb * xb - x03) / den : (double)b / (double)ncol_b;
          const double f = 0.3 * (double)b / (double)ncol_b + 0.7 * cubic;
          foff[b] = b == ncol_b ? P.far_cap : (uint64_t)((double)P.far_cap * (f < 1.0 ? f : 1.0));
        }
        for (uint32_t b = tid; b < ncol_b; b += kReduceThreads) fcnt[b] = 0;
        if (tid == 0) S.far_overflow = 0;
        __syncthreads();
      }
      if (tid == 0) v_toggle(rbirth);
      gen(rbirth, first, tid, kReduceThreads);
      publish();
      bool essential = false;
      uint64_t pivot = 0;
      for (;;) {
        if (S.abort_flag) break;
        t0 = clock64();
        const bool ok = scan(pos);
        cyc[0] += clock64() - t0;
        if (!ok) {
          if (far_on && S.far_overflow) far_on = false;   // (uniform: written before the barriers inside scan())
          if (far_on) {
            // jump to the next window that holds parked keys and toggle them in (the bitset is all zero now)
            t0 = clock64();
            uint32_t cand = 0xffffffffu;
            for (uint32_t b = cur_b + 1 + tid; b < ncol_b; b += kReduceThreads)
              if (fcnt[b]) { cand = b; break; }
            cand = block_min(cand);
            if (cand == 0xffffffffu) { essential = true; break; }
            cur_b = cand;
            wbase = wbase0 + (uint64_t)cand * wbits;
            nbase = 0;
            pos = 0;
            const uint32_t cnt = fcnt[cand];
            const uint64_t* src = far + foff[cand];
            for (uint32_t i = tid; i < cnt; i += kReduceThreads) toggle_key(src[i]);
            publish();
            ++slides;
            far_keys += cnt;
            cyc[4] += clock64() - t0;
            continue;
          }
          if (wbase + wbits >= kmax) { essential = true; break; }
          // slide the window (it is all zero now) and re-enumerate delta(V) for the new range
          t0 = clock64();
          wbase += wbits;
          nbase = 0;
          pos = 0;
          v_compact();
          add_edges(vlist(S.vsel), S.vcount, wbase, false);
          publish();
          ++slides;
          ext_edges += S.vcount;
          cyc[4] += clock64() - t0;
          continue;
        }
        ++pops;
        const uint64_t pk = wbase + pos;
        t0 = clock64();
        int M, w;
        uint2 eaM = make_uint2(0, 0);
        int owner = -2;  // -1 none, -2 apparent, >=0 reduced column id
        if constexpr (DIM == 1) {
          if ((pk >> 32) == 0) { M = (int)((uint32_t)pk / (uint32_t)n); w = n - 1 - (int)((uint32_t)pk - (uint32_t)M * (uint32_t)n); }
          else { M = (int)(pk / (uint64_t)n); w = n - 1 - (int)(pk % (uint64_t)n); }
          eaM = __ldg(&EA[M]);
          if ((int)eaM.y != w) owner = hash_find(pk);
        } else {
          // tetrahedron (M4; p > q): its youngest facet is the triangle (M4, q); the pair is apparent iff that triangle's first
          // cofacet is this tetrahedron, i.e. apex4[(M4, q)] == p.  M holds the facet's key, the column that owns the pivot.
          const uint64_t M4 = pk / n2, rem = pk - M4 * n2;
          const int pv = n - 1 - (int)(rem / (uint64_t)n), qv = n - 1 - (int)(rem % (uint64_t)n);
          M = (int)(M4 * (uint64_t)n + (uint64_t)(n - 1 - qv));
          w = pv;
          if (!TDA_BC((uint64_t)(uint32_t)M < (uint64_t)P.E * (uint64_t)n && M4 < (uint64_t)T, "pivot facet key/M4", (uint32_t)M, M4)) { S.abort_flag = TDA_ERR_INVALID; break; }
          if ((int)A4[(uint32_t)M] != pv) owner = hash_find(pk);
        }
        cyc[1] += clock64() - t0;
        if (owner == -1) { pivot = pk; break; }
        ++additions;
        t0 = clock64();
        if (owner == -2) {
          if (tid == 0) v_toggle(M);
          if constexpr (DIM == 1) gen(M, eaM.x, pk, tid, kReduceThreads);
          else gen_tri((uint32_t)M, pk, tid, kReduceThreads);
          publish();
          cyc[2] += clock64() - t0;
        } else {
          if (!TDA_BC(owner < P.cap1, "owner/cap1", owner, P.cap1)) { S.abort_flag = TDA_ERR_INVALID; break; }
          const int64_t vs = P.vstart[(size_t)p * P.cap1 + owner];
          const int vn = P.vlen[(size_t)p * P.cap1 + owner];
          if (!TDA_BC(vs >= 0 && vn >= 0 && vs + vn <= P.vpool_cap, "vstart+vlen/vpool_cap", vs, vn)) { S.abort_flag = TDA_ERR_INVALID; break; }
          const uint32_t* ov = P.vpool + (size_t)p * P.vpool_cap + vs;
          if (S.vcount + (uint32_t)vn > (uint32_t)P.vcap) v_compact();
          add_edges(ov, (uint32_t)vn, pk, true);
          publish();
          badd_edges += vn;
          cyc[3] += clock64() - t0;
        }
        pos += 1;  // pk itself was toggled off by the addition
      }
      if (S.abort_flag) break;
      t0 = clock64();
      // finalise the column
      far_on = false;   // (clean_window re-enumerates delta(V) inside the last window only)
      v_compact();
      const uint32_t nv = S.vcount;
      if (nv > maxv) maxv = nv;
      if (!essential) {
        // store V (all edges, including the column's own) for later additions
        if (vpool_used + nv > P.vpool_cap) { if (tid == 0) S.abort_flag = TDA_ERR_CAPACITY; __syncthreads(); break; }
        uint32_t* dst = P.vpool + (size_t)p * P.vpool_cap + vpool_used;
        const uint32_t* list = vlist(S.vsel);
        for (uint32_t i = tid; i < nv; i += kReduceThreads) dst[i] = list[i];
        if (tid == 0) {
          P.vstart[(size_t)p * P.cap1 + ci] = vpool_used;
          P.vlen[(size_t)p * P.cap1 + ci] = (int)nv;
          hash_insert(pivot, ci);
        }
        vpool_used += nv;
        clean_window(pos, nv);
      } else {
        for (uint32_t i = tid; i < s1words; i += kReduceThreads) s1[i] = 0;
      }
      const float birth = DIM == 1 ? SD[rbirth] : SD[(uint32_t)rbirth / (uint32_t)n];
      float death = INFINITY;
      int Md = -1, wd = -1;
      if (!essential) {
        if constexpr (DIM == 1) { Md = (int)(pivot / (uint64_t)n); wd = n - 1 - (int)(pivot % (uint64_t)n); }
        else Md = (int)(pivot / n2);
        death = SD[Md];
      }
      if (essential || death > birth) {
        if (tid == 0 && TDA_BC(nrows < P.cap1, "nrows/cap1", nrows, P.cap1)) {
          out[2 * nrows] = birth; out[2 * nrows + 1] = death;
          if (DIM == 1 && outs) {
            const uint32_t e = EN[rbirth];
            outs[2 * nrows] = edge_index((int)(e >> 16), (int)(e & 0xffffu));
            if (essential) outs[2 * nrows + 1] = -1;
            else {
              const uint32_t em = EN[Md];
              int x = (int)(em >> 16), y = (int)(em & 0xffffu), z = wd, t;
              if (x < y) { t = x; x = y; y = t; }
              if (y < z) { t = y; y = z; z = t; }
              if (x < y) { t = x; x = y; y = t; }
              outs[2 * nrows + 1] = (int64_t)x * (x - 1) * (x - 2) / 6 + (int64_t)y * (y - 1) / 2 + z;
            }
          }
        }
        ++nrows;
      }
      v_clear();   // also makes the hash / vpool writes of thread 0 visible block-wide (barriers inside)
      __threadfence();
      cyc[5] += clock64() - t0;
    }
    __syncthreads();
    if (S.abort_flag) {  // leave the scratch clean for the next problem: wipe the whole window
      v_compact();
      v_clear();
      for (uint64_t i = tid; i < (wbits >> 5); i += kReduceThreads) bits[i] = 0;
      for (uint32_t i = tid; i < s1words; i += kReduceThreads) s1[i] = 0;
      for (uint32_t i = tid; i < (uint32_t)kNearWords; i += kReduceThreads) S.nearw[i] = 0;
      __threadfence();
      __syncthreads();
    }
    atomicAdd(&S.toggles, my_toggles);
    __syncthreads();
    if (tid == 0) {
      P.counts[p * 4 + 1] = nrows;
      P.counts[p * 4 + 3] = S.abort_flag;
      st[ST_REDUCED] = (unsigned long long)nb;
      st[ST_ADDITIONS] = additions;
      st[ST_PUSHES] = S.toggles;
      st[ST_POPS] = pops;
      st[ST_EXTENSIONS] = slides + n_refills;
      st[ST_MAXV] = maxv;
      for (int q = 0; q < 6; ++q) st[ST_CYC_EXTRACT + q] = (unsigned long long)cyc[q];
      st[ST_BADD_EDGES] = badd_edges;
      st[ST_EXT_EDGES] = ext_edges + n_passes + far_keys;
    }
    __syncthreads();
  }
};

template <int DIM>
__global__ void __launch_bounds__(kReduceThreads) rips_reduce_kernel(const __grid_constant__ ReduceParams P) {
  __shared__ ReduceSmem S;
  extern __shared__ uint32_t s1_dyn[];
  Reducer<DIM> red(P, S, s1_dyn);
  for (;;) {
    if (threadIdx.x == 0) S.problem = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int p = S.problem;
    __syncthreads();
    if (p >= P.batch) break;
    red.run_problem(p);
  }
}

// ------------------------------------------------------------------------------------------------
// residual reduction, ROW-SWEEP formulation (the default)
//
// The same reduction (ripser's order, same pivots, same V's), organised around the rows of the key space instead
// of around the keys: row M = all triangles whose longest edge has rank M, one bit per opposite vertex.  With
// X = V as a symmetric n x n bit matrix (x_e = X[a][b]), the working column restricted to row M=(c,d) is
//        r(M) = ( x_M * 1  ^  X[c][.]  ^  X[d][.] )  &  lune(M),      lune(M) = { w : rank(c,w) < M and rank(d,w) < M }
// so a row costs a few word-parallel bit operations instead of n scattered atomics per added edge, and rows whose
// endpoints are not incident to V ("untouched") are zero and are skipped by a streaming filter over the edge list.
// The column walks the rows upward from its birth edge; inside a row the lowest key is the highest vertex:
// apparent pivot (w == apex[M]) -> x_M flips, r ^= lune;  pivot owned by a reduced column -> X ^= V_owner and the
// row is recomputed;  unowned pivot -> death.  Per CTA: 16 warps compute 16 heavy rows at a time (lune words by
// coalesced rank-row loads + ballot), warp 0 then resolves them in order from shared memory and patches the later
// rows of the group for every flip, so the dependent chain of pivots costs ~100 cycles per pivot instead of a
// global-memory round trip per step.
constexpr int kSweepThreads = 512;
constexpr int kSweepWarps = kSweepThreads / 32;
constexpr int kGroupRows = 32;                    // heavy rows resolved per group (one lane of the resolver per row)
constexpr int kChunkRows = kSweepThreads;         // rows filtered per pass
constexpr int kDenseThreshold = 64;               // heavy rows in a chunk from which the column switches to the Pm lune path
enum { SW_DONE = 0, SW_RESTART = 1, SW_REDUCED = 2, SW_DEATH = 3 };

struct SweepSmem {
  uint2 chunk_ea[kChunkRows];
  uint32_t heavy[kChunkRows];
  uint32_t wcnt[kSweepWarps];
  uint32_t row_rank[2][kGroupRows], row_c[2][kGroupRows], row_d[2][kGroupRows];
  int row_apex[2][kGroupRows];
  uint32_t dirty[2];   // bit s: row s of the group in this buffer has a non-zero working row
  uint32_t simple[2];  // bit s: r == lune (the whole row is odd): one apparent pivot clears it -- the common case, since V is a
                       //        cocycle below the cursor: x_M ^ x_cw ^ x_dw is the same for every w of the lune
  uint32_t both[2];    // bit s: both endpoints of row s were already touched when the row was formed
  uint32_t conf[2][kGroupRows];  // conf[.][f] bit l (l > f): rows f and l of the group share a vertex (a flip of f patches l)
  uint32_t patched;    // rows of the NEXT group that the flips of the group just resolved have changed
  uint32_t flipmask;   // rows of the group just resolved whose edge joined V
  uint32_t nheavy;
  uint32_t vcount, vcount2, vsel;
  int abort_flag, problem;
  int res_status, res_s, res_owner;
  uint32_t res_w;
  unsigned long long additions, pivots, heavy_rows, restarts, groups;
  // substitute-then-verify mode
  uint32_t sub_done[kChunkRows / 32];   // chunk rows whose x is final for this pass (all but the heavy apparent rows not yet substituted)
  uint16_t fliprow[kChunkRows];         // chunk-local rows whose edge was flipped since the chunk started (for the undo)
  uint32_t nflip;
  uint32_t new_touch;                   // a flip touched a vertex that V did not touch before: filter the chunk again
  int fail_w;                           // highest vertex of the first non-empty row
};

template <int WPL, bool VERIFY = false>   // WPL: words of a row per lane of the resolver (W <= 32 * WPL); VERIFY: substitute-then-verify
struct Sweeper {
  static constexpr uint64_t kEmpty = ~0ull;
  const ReduceParams& P;
  SweepSmem& S;
  uint32_t* touched;   // [W]
  uint32_t* Sr;        // [2][kGroupRows][W]   working rows   } double buffered: the next group is produced (from the X of
  uint32_t* Slm;       // [2][kGroupRows][W]   lune masks     } before this group's flips) while this group is resolved
  const int tid, lane, warp;
  const int* R; const uint32_t* EN; const uint2* EA; int T; int n; int W;
  uint32_t* X; uint32_t* vbits; uint32_t* vl0;
  uint32_t* Pm;          // Pm[c] = { w : rank(c,w) < p_pos }  (exact; advanced chunk by chunk, repositioned at column start)
  uint32_t p_pos; bool p_valid;
  bool p_mode;           // dense mode: Pm is maintained and the lune masks come from it (else from the rank rows)
  uint32_t* vhead;       // [n]  dense mode: per vertex, head of the list of endpoint slots of the current chunk's rows
  uint32_t* vnext;       // [2 * kChunkRows]  slot 2j+side -> next slot of the same vertex (0xffffffff = end)
  uint64_t* hkeys; int* hvals;

  __device__ Sweeper(const ReduceParams& p, SweepSmem& s, uint32_t* dyn)
      : P(p), S(s), tid(threadIdx.x), lane(threadIdx.x & 31), warp(threadIdx.x >> 5) {
    n = P.n;
    W = P.xw;
    touched = dyn;
    Sr = dyn + W;
    Slm = Sr + (size_t)2 * kGroupRows * W;
    vhead = Slm + (size_t)2 * kGroupRows * W;
    vnext = vhead + n;
    X = P.xmat + (size_t)blockIdx.x * (size_t)n * W;
    Pm = P.pmat + (size_t)blockIdx.x * (size_t)n * W;
    p_pos = 0; p_valid = false; p_mode = false;
    vbits = P.vbits + (size_t)blockIdx.x * P.vwords;
    vl0 = P.vlist + (size_t)blockIdx.x * 2 * P.vcap;
  }
  __device__ __forceinline__ uint32_t* vlist(uint32_t sel) const { return vl0 + (size_t)sel * P.vcap; }
  __device__ __forceinline__ bool tbit(uint32_t v) const { return (touched[v >> 5] >> (v & 31)) & 1u; }

  // ---- V bookkeeping (same scheme as the bitset reducer: parity bits + a list that may hold duplicates)
  __device__ __forceinline__ void v_toggle(int re) {
    atomicXor(&vbits[(uint32_t)re >> 5], 1u << (re & 31));
    const uint32_t pos = atomicAdd(&S.vcount, 1u);
    if (pos < (uint32_t)P.vcap) vlist(S.vsel)[pos] = (uint32_t)re;
    else S.abort_flag = TDA_ERR_CAPACITY;
  }
  __device__ __forceinline__ void x_flip(uint32_t c, uint32_t d) {
    atomicXor(&X[(size_t)c * W + (d >> 5)], 1u << (d & 31));
    atomicXor(&X[(size_t)d * W + (c >> 5)], 1u << (c & 31));
  }
  __device__ __forceinline__ void v_compact() {
    __syncthreads();
    const uint32_t nin = min(S.vcount, (uint32_t)P.vcap);
    const uint32_t* src = vlist(S.vsel);
    uint32_t* dst = vlist(S.vsel ^ 1);
    if (tid == 0) S.vcount2 = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < nin; i0 += kSweepThreads) {
      const uint32_t i = i0 + tid;
      bool keep = false;
      uint32_t e = 0;
      if (i < nin) {
        e = src[i];
        const uint32_t m = 1u << (e & 31);
        keep = (atomicAnd(&vbits[e >> 5], ~m) & m) != 0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      uint32_t bs = 0;
      if (lane == 0 && bal) bs = atomicAdd(&S.vcount2, (uint32_t)__popc(bal));
      bs = __shfl_sync(0xffffffffu, bs, 0);
      if (keep) dst[bs + __popc(bal & ((1u << lane) - 1))] = e;
    }
    __syncthreads();
    const uint32_t nout = S.vcount2;
    for (uint32_t i = tid; i < nout; i += kSweepThreads) {
      const uint32_t e = dst[i];
      atomicOr(&vbits[e >> 5], 1u << (e & 31));
    }
    __syncthreads();
    if (tid == 0) { S.vsel ^= 1; S.vcount = nout; }
    __syncthreads();
  }
  // after v_compact: clear the parity bits, the X bits of every edge of V and the touched mask; empty the list
  __device__ __forceinline__ void v_clear() {
    const uint32_t nin = S.vcount;
    const uint32_t* src = vlist(S.vsel);
    for (uint32_t i = tid; i < nin; i += kSweepThreads) {
      const uint32_t e = src[i];
      atomicAnd(&vbits[e >> 5], ~(1u << (e & 31)));
      const uint32_t en = __ldg(&EN[e]);
      x_flip(en >> 16, en & 0xffffu);
    }
    for (int i = tid; i < W; i += kSweepThreads) touched[i] = 0;
    __threadfence();
    __syncthreads();
    if (tid == 0) S.vcount = 0;
    __syncthreads();
  }
  __device__ __forceinline__ int hash_find(uint64_t key) const {
    uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & (uint32_t)(P.hcap - 1);
    for (;;) {
      const uint64_t k = hkeys[h];
      if (k == key) return hvals[h];
      if (k == kEmpty) return -1;
      h = (h + 1) & (uint32_t)(P.hcap - 1);
    }
  }
  __device__ __forceinline__ void hash_insert(uint64_t key, int val) {  // thread 0
    uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & (uint32_t)(P.hcap - 1);
    while (hkeys[h] != kEmpty) h = (h + 1) & (uint32_t)(P.hcap - 1);
    hkeys[h] = key;
    hvals[h] = val;
  }
  __device__ __forceinline__ void sort_blist(int* bl, int nb) {
    int np2 = 1;
    while (np2 < nb) np2 <<= 1;
    for (int i = nb + tid; i < np2; i += kSweepThreads) bl[i] = 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < np2; i += kSweepThreads) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const int a = bl[i], b = bl[ixj];
            const bool up = ((i & k) == 0);
            if ((a > b) == up) { bl[i] = b; bl[ixj] = a; }
          }
        }
        __syncthreads();
      }
  }

  // ---- Pm bookkeeping.  All threads call; ends with the bits performed and a barrier.
  __device__ __forceinline__ void p_set_rows(uint32_t lo, uint32_t hi, bool set) {  // rows [lo, hi)
    for (uint32_t row = lo + tid; row < hi; row += kSweepThreads) {
      const uint32_t en = __ldg(&EA[row]).x;
      const uint32_t c = en >> 16, d = en & 0xffffu;
      if (set) {
        atomicOr(&Pm[(size_t)c * W + (d >> 5)], 1u << (d & 31));
        atomicOr(&Pm[(size_t)d * W + (c >> 5)], 1u << (c & 31));
      } else {
        atomicAnd(&Pm[(size_t)c * W + (d >> 5)], ~(1u << (d & 31)));
        atomicAnd(&Pm[(size_t)d * W + (c >> 5)], ~(1u << (c & 31)));
      }
    }
  }
  __device__ __forceinline__ void p_move(uint32_t target) {
    const uint32_t dist = target > p_pos ? target - p_pos : p_pos - target;
    if (!p_valid || (uint64_t)dist * 32ull > (uint64_t)n * (uint64_t)n) {
      // rebuild from the rank matrix: one warp per vertex row, a word per ballot
      for (int c = warp; c < n; c += kSweepWarps) {
        const int* Rc = R + (size_t)c * n;
        for (int k0 = 0; k0 < W; k0 += 32) {
          uint32_t mine = 0;
          const int kend = min(32, W - k0);
#pragma unroll 8
          for (int kk = 0; kk < kend; ++kk) {
            const int w = (k0 + kk) * 32 + lane;
            const int ra = w < n ? __ldg(&Rc[w]) : kRankDiag;
            const unsigned word = __ballot_sync(0xffffffffu, ra < (int)target);
            if (lane == kk) mine = word;
          }
          if (k0 + lane < W) Pm[(size_t)c * W + k0 + lane] = mine;
        }
      }
      p_valid = true;
    } else if (target > p_pos) {
      p_set_rows(p_pos, target, true);
    } else if (target < p_pos) {
      p_set_rows(target, p_pos, false);
    }
    p_pos = target;
    __threadfence();
    __syncthreads();
  }

  // ---- lune(M=(c,d)) = { w : rank(c,w) < M and rank(d,w) < M } of one heavy row (one warp), two ways.
  // (sparse mode, Pm not maintained) from the two rank rows: coalesced loads, a 32-word block in flight at once, ballot
  __device__ __forceinline__ void lune_from_ranks(uint32_t* lm, uint32_t c, uint32_t d, uint32_t Mrow) {
    const int* Rc = R + (size_t)c * n;
    const int* Rd = R + (size_t)d * n;
    for (int k0 = 0; k0 < W; k0 += 16) {
      int ra[16], rb[16];
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const int w = (k0 + kk) * 32 + lane;
        const bool ok = (k0 + kk) < W && w < n;
        ra[kk] = ok ? __ldg(&Rc[w]) : kRankDiag;
        rb[kk] = ok ? __ldg(&Rd[w]) : kRankDiag;
      }
      uint32_t lmine = 0;
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const unsigned word = __ballot_sync(0xffffffffu, ra[kk] < (int)Mrow && rb[kk] < (int)Mrow);
        if (lane == kk) lmine = word;
      }
      if (lane < 16 && k0 + lane < W) lm[k0 + lane] = lmine;
    }
  }
  // (dense mode) Pm[c] & Pm[d] (the words `pw`, loaded by the caller).  Pm has been advanced to the END of the chunk by the
  // filter, so the chunk's edges that are not below M (the rows after this one) are taken out again: lanes 0 and 1 walk the
  // chunk's per-vertex row lists of c and d (a vertex occurs in ~1.5 rows of a chunk on average).
  __device__ __forceinline__ void lune_from_pm(uint32_t* lm, const uint32_t (&pw)[WPL], uint32_t c, uint32_t d, uint32_t hidx) {
#pragma unroll
    for (int q = 0; q < WPL; ++q)
      if ((lane + 32 * q) < W) lm[lane + 32 * q] = pw[q];
    __syncwarp();
    if (lane < 2) {
      uint32_t slot = vhead[lane == 0 ? c : d];
      while (slot != 0xffffffffu) {
        const uint32_t j = slot >> 1;
        if (j > hidx) {
          const uint32_t e2 = S.chunk_ea[j].x;
          const uint32_t other = (slot & 1u) ? (e2 >> 16) : (e2 & 0xffffu);   // slot side 0 = the row's c endpoint, 1 = its d endpoint
          atomicAnd(&lm[other >> 5], ~(1u << (other & 31)));
        }
        slot = vnext[slot];
      }
    }
  }

  // ---- the heavy rows [i, i + ns) of the chunk at `pos` into buffer `buf`: lune mask, working row
  // r = (x_M ^ X[c] ^ X[d]) & lune and its class.  Rows are dealt round-robin to the warps [wfirst, wfirst + nw).  X is not
  // written while a group is being resolved (flips are deferred), so a group produced during the resolve of the previous
  // one sees the X of before that group's flips; patch_next() applies them afterwards.
  __device__ __forceinline__ void finish_row(int buf, int sl, uint32_t Mrow, uint2 ea, uint32_t xmw, const uint32_t (&xx)[WPL]) {
    const uint32_t c = ea.x >> 16, d = ea.x & 0xffffu;
    const uint32_t* lm = Slm + ((size_t)buf * kGroupRows + sl) * W;
    uint32_t* r = Sr + ((size_t)buf * kGroupRows + sl) * W;
    const uint32_t xm = ((xmw >> (d & 31)) & 1u) ? 0xffffffffu : 0u;
    uint32_t any = 0, diff = 0;
#pragma unroll
    for (int q = 0; q < WPL; ++q)
      if ((lane + 32 * q) < W) {
        const uint32_t l = lm[lane + 32 * q];
        const uint32_t v = (xm ^ xx[q]) & l;
        r[lane + 32 * q] = v;
        any |= v;
        diff |= v ^ l;
      }
    const bool nz = __any_sync(0xffffffffu, any != 0);
    const bool full = !__any_sync(0xffffffffu, diff != 0);
    if (lane == 0) {
      S.row_rank[buf][sl] = Mrow; S.row_c[buf][sl] = c; S.row_d[buf][sl] = d; S.row_apex[buf][sl] = (int)ea.y;
      if (nz) atomicOr(&S.dirty[buf], 1u << sl);
      if (nz && full) atomicOr(&S.simple[buf], 1u << sl);
      if (tbit(c) && tbit(d)) atomicOr(&S.both[buf], 1u << sl);
    }
  }
  __device__ __forceinline__ void produce(int buf, uint32_t pos, uint32_t i, int ns, int wfirst, int nw) {
    constexpr int kMaxMine = 3;   // 32 rows over >= 15 warps
    if (p_mode) {
      // dense mode: every global word of all of this warp's rows is requested before any of them is used
      uint2 ea[kMaxMine];
      uint32_t xmw[kMaxMine], xx[kMaxMine][WPL], pw[kMaxMine][WPL];
#pragma unroll
      for (int t = 0; t < kMaxMine; ++t) {
        const int sl = warp - wfirst + t * nw;
        if (sl < ns) {
          ea[t] = S.chunk_ea[S.heavy[i + sl]];
          const uint32_t c = ea[t].x >> 16, d = ea[t].x & 0xffffu;
          const uint32_t* Xc = X + (size_t)c * W;
          const uint32_t* Xd = X + (size_t)d * W;
          const uint32_t* Pc = Pm + (size_t)c * W;
          const uint32_t* Pd = Pm + (size_t)d * W;
          xmw[t] = __ldcg(&Xc[d >> 5]);
#pragma unroll
          for (int q = 0; q < WPL; ++q) {
            const bool ok = (lane + 32 * q) < W;
            xx[t][q] = ok ? (__ldcg(&Xc[lane + 32 * q]) ^ __ldcg(&Xd[lane + 32 * q])) : 0u;
            pw[t][q] = ok ? (__ldcg(&Pc[lane + 32 * q]) & __ldcg(&Pd[lane + 32 * q])) : 0u;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < kMaxMine; ++t) {
        const int sl = warp - wfirst + t * nw;
        if (sl < ns) {
          const uint32_t hidx = S.heavy[i + sl];
          uint32_t* lm = Slm + ((size_t)buf * kGroupRows + sl) * W;
          lune_from_pm(lm, pw[t], ea[t].x >> 16, ea[t].x & 0xffffu, hidx);
          __syncwarp();
          finish_row(buf, sl, pos + hidx, ea[t], xmw[t], xx[t]);
        }
      }
      return;
    }
    for (int sl = warp - wfirst; sl < ns; sl += nw) {
      const uint32_t hidx = S.heavy[i + sl];
      const uint2 ea = S.chunk_ea[hidx];
      const uint32_t Mrow = pos + hidx;
      const uint32_t c = ea.x >> 16, d = ea.x & 0xffffu;
      const uint32_t* Xc = X + (size_t)c * W;
      const uint32_t* Xd = X + (size_t)d * W;
      uint32_t* lm = Slm + ((size_t)buf * kGroupRows + sl) * W;
      const uint32_t xmw = __ldcg(&Xc[d >> 5]);
      uint32_t xx[WPL];
#pragma unroll
      for (int q = 0; q < WPL; ++q) xx[q] = (lane + 32 * q) < W ? (__ldcg(&Xc[lane + 32 * q]) ^ __ldcg(&Xd[lane + 32 * q])) : 0u;
      lune_from_ranks(lm, c, d, Mrow);
      __syncwarp();
      finish_row(buf, sl, Mrow, ea, xmw, xx);
    }
  }

  // ---- after a group: its flips go out to global memory (X, the parity bits, the V list), one lane of warp 0 per flip ...
  __device__ __forceinline__ void write_flips(int buf) {  // warp 0
    const uint32_t flipmask = S.flipmask;
    if (!flipmask) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&S.vcount, (uint32_t)__popc(flipmask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if ((flipmask >> lane) & 1u) {
      const uint32_t c = S.row_c[buf][lane], d = S.row_d[buf][lane], rk = S.row_rank[buf][lane];
      atomicXor(&X[(size_t)c * W + (d >> 5)], 1u << (d & 31));
      atomicXor(&X[(size_t)d * W + (c >> 5)], 1u << (c & 31));
      atomicXor(&vbits[rk >> 5], 1u << (rk & 31));
      const uint32_t vp = base + __popc(flipmask & ((1u << lane) - 1));
      if (vp < (uint32_t)P.vcap) vlist(S.vsel)[vp] = rk;
      else S.abort_flag = TDA_ERR_CAPACITY;
    }
  }
  // ... and are applied to the rows of the next group (produced from the X of before): a flip of edge (c,d) changes exactly one
  // bit of every row that shares a vertex with it.  All threads: thread t looks at row t%32 and flips t/32, t/32 + 16.
  __device__ __forceinline__ void patch_next(int buf, int ns_next) {
    const uint32_t flipmask = S.flipmask;
    const int l2 = tid & 31;
    if (!flipmask || l2 >= ns_next) return;
    const int nb = buf ^ 1;
    const uint32_t c2 = S.row_c[nb][l2], d2 = S.row_d[nb][l2];
    for (int f = tid >> 5; f < kGroupRows; f += kSweepWarps) {
      if (!((flipmask >> f) & 1u)) continue;
      const uint32_t c = S.row_c[buf][f], d = S.row_d[buf][f];
      int vb = -1;
      if (c2 == c) vb = (int)d; else if (c2 == d) vb = (int)c; else if (d2 == c) vb = (int)d; else if (d2 == d) vb = (int)c;
      if (vb >= 0) {
        const uint32_t m = 1u << (vb & 31);
        const size_t o = ((size_t)nb * kGroupRows + l2) * W + (vb >> 5);
        if (Slm[o] & m) { atomicXor(&Sr[o], m); atomicOr(&S.patched, 1u << l2); }
      }
    }
  }

  // ---- which rows of a produced group share a vertex (all threads; same thread layout as patch_next)
  __device__ __forceinline__ void conflicts(int buf, int ns) {
    const int l2 = tid & 31;
    if (l2 >= ns) return;
    const uint32_t c2 = S.row_c[buf][l2], d2 = S.row_d[buf][l2];
    for (int f = tid >> 5; f < kGroupRows; f += kSweepWarps) {
      if (f >= l2) continue;
      const uint32_t c = S.row_c[buf][f], d = S.row_d[buf][f];
      if (c2 == c || c2 == d || d2 == c || d2 == d) atomicOr(&S.conf[buf][f], 1u << l2);
    }
  }

  // ---- warp 0: walk the dirty rows of the group in order.  The current row lives in registers (WPL words per lane);
  // lane l also watches row l of the group for patches.  Nothing here touches global memory except the pivot hash
  // look-up of a non-apparent pivot; the flips are written out once per group.
  __device__ __forceinline__ void resolve(int buf, int ns) {
    int status = SW_DONE, res_s = 0, res_owner = -1;
    uint32_t res_w = 0;
    uint32_t additions = 0, pivots = 0;
    bool new_touch = false;
    const uint32_t myc = lane < ns ? S.row_c[buf][lane] : 0xffffffffu;
    const uint32_t myd = lane < ns ? S.row_d[buf][lane] : 0xffffffffu;
    const uint32_t myrank = lane < ns ? S.row_rank[buf][lane] : 0u;
    const int myapex = lane < ns ? S.row_apex[buf][lane] : -3;
    const uint32_t* lmbase = Slm + (size_t)buf * kGroupRows * W;
    uint32_t* rbase = Sr + (size_t)buf * kGroupRows * W;
    const uint32_t patched0 = S.patched;   // rows changed by the previous group's flips: re-examine them from shared memory
    uint32_t flipmask = 0;   // rows whose edge joined V in this group (at most once per row: its apparent key is unique)
    uint32_t mask = S.dirty[buf] | patched0;
    uint32_t fast = S.simple[buf] & S.both[buf] & ~patched0;   // rows that take the register-only path (until a patch touches them)
    while (mask && status == SW_DONE) {
      const int s = __ffs(mask) - 1;
      mask &= mask - 1;
      if ((fast >> s) & 1u) {
        // whole row odd, endpoints already touched: the apparent pivot (M, apex) joins the row's edge to V and clears the row
        flipmask |= 1u << s;
        ++pivots;
        ++additions;
        const uint32_t cm = S.conf[buf][s];   // later rows of the group that share a vertex with this edge (rare)
        if (cm) {
          const uint32_t c = __shfl_sync(0xffffffffu, myc, s), d = __shfl_sync(0xffffffffu, myd, s);
          bool patched = false;
          if ((cm >> lane) & 1u) {
            int vb = -1;
            if (myc == c) vb = (int)d; else if (myc == d) vb = (int)c; else if (myd == c) vb = (int)d; else if (myd == d) vb = (int)c;
            if (vb >= 0) {
              const uint32_t m = 1u << (vb & 31);
              if (lmbase[(size_t)lane * W + (vb >> 5)] & m) { rbase[(size_t)lane * W + (vb >> 5)] ^= m; patched = true; }
            }
          }
          const uint32_t pm = __ballot_sync(0xffffffffu, patched);
          mask |= pm;
          fast &= ~pm;
        }
        continue;
      }
      const uint32_t c = __shfl_sync(0xffffffffu, myc, s), d = __shfl_sync(0xffffffffu, myd, s);
      const int apex = __shfl_sync(0xffffffffu, myapex, s);
      const uint32_t cm = S.conf[buf][s];
      uint32_t* r = rbase + (size_t)s * W;
      const uint32_t* lm = lmbase + (size_t)s * W;
      uint32_t rw[WPL];
      uint32_t anyw = 0, diffw = 0;
#pragma unroll
      for (int q = 0; q < WPL; ++q) {
        const bool ok = (lane + 32 * q) < W;
        rw[q] = ok ? r[lane + 32 * q] : 0u;
        anyw |= rw[q];
        diffw |= rw[q] ^ (ok ? lm[lane + 32 * q] : 0u);
      }
      // a row that was patched (or formed before its endpoints were touched) is usually all-or-nothing again: re-test its class
      if (!__any_sync(0xffffffffu, anyw != 0)) continue;
      if (!__any_sync(0xffffffffu, diffw != 0) && tbit(c) && tbit(d)) {
        flipmask |= 1u << s;
        ++pivots;
        ++additions;
        if (cm) {
          bool patched = false;
          if ((cm >> lane) & 1u) {
            int vb = -1;
            if (myc == c) vb = (int)d; else if (myc == d) vb = (int)c; else if (myd == c) vb = (int)d; else if (myd == d) vb = (int)c;
            if (vb >= 0) {
              const uint32_t m = 1u << (vb & 31);
              if (lmbase[(size_t)lane * W + (vb >> 5)] & m) { rbase[(size_t)lane * W + (vb >> 5)] ^= m; patched = true; }
            }
          }
          const uint32_t pm = __ballot_sync(0xffffffffu, patched);
          mask |= pm;
          fast &= ~pm;
        }
        continue;
      }
      for (;;) {
        int best = -1;
#pragma unroll
        for (int q = 0; q < WPL; ++q)
          if (rw[q]) best = (lane + 32 * q) * 32 + 31 - __clz(rw[q]);   // q ascending: the last hit is this lane's highest
        best = __reduce_max_sync(0xffffffffu, best);
        if (best < 0) break;
        ++pivots;
        if (best == apex) {  // apparent pair (row edge, this triangle): the row's edge joins V; x_M flips, so r ^= lune
          const bool nt = !tbit(c) || !tbit(d);
          if (nt) {
            __syncwarp();
            if (lane == 0) {
              touched[c >> 5] |= 1u << (c & 31);
              touched[d >> 5] |= 1u << (d & 31);
            }
            new_touch = true;
          }
          flipmask |= 1u << s;
#pragma unroll
          for (int q = 0; q < WPL; ++q)
            if ((lane + 32 * q) < W) rw[q] ^= lm[lane + 32 * q];
          // later rows of the group that share a vertex with this edge see exactly one bit of X flip
          bool patched = false;
          if ((cm >> lane) & 1u) {
            int vb = -1;
            if (myc == c) vb = (int)d; else if (myc == d) vb = (int)c; else if (myd == c) vb = (int)d; else if (myd == d) vb = (int)c;
            if (vb >= 0) {
              const uint32_t m = 1u << (vb & 31);
              if (lmbase[(size_t)lane * W + (vb >> 5)] & m) { rbase[(size_t)lane * W + (vb >> 5)] ^= m; patched = true; }
            }
          }
          const uint32_t pm = __ballot_sync(0xffffffffu, patched);
          mask |= pm;
          fast &= ~pm;
          ++additions;
          continue;
        }
        const uint32_t Mrow = __shfl_sync(0xffffffffu, myrank, s);
        const uint64_t key = (uint64_t)Mrow * (uint64_t)n + (uint64_t)(n - 1 - best);
        const int owner = hash_find(key);
        res_s = s; res_w = (uint32_t)best; res_owner = owner;
        status = owner < 0 ? SW_DEATH : SW_REDUCED;
        if (owner >= 0) ++additions;
        break;
      }
      if (status == SW_DONE && new_touch) { status = SW_RESTART; res_s = s; }  // rows skipped as untouched may matter now
    }
    __syncwarp();
    if (lane == 0) {
      S.res_status = status; S.res_s = res_s; S.res_owner = res_owner; S.res_w = res_w;
      S.additions += additions; S.pivots += pivots;
      S.dirty[buf] = 0; S.simple[buf] = 0; S.both[buf] = 0; S.patched = 0;
      S.flipmask = flipmask;
    }
    S.conf[buf][lane] = 0;
    {
    }
  }

  __device__ __forceinline__ void run_problem(int p) {
    R = P.rank + (size_t)p * n * n;
    EN = P.ends + (size_t)p * P.E;
    EA = P.ea + (size_t)p * P.E;
    T = P.T[p];
    hkeys = P.hkeys + (size_t)p * P.hcap;
    hvals = P.hvals + (size_t)p * P.hcap;
    const float* SD = P.sdist + (size_t)p * P.E;
    int* bl = P.blist + (size_t)p * P.cap1;
    const int nb = P.bcount[p];
    unsigned long long* st = P.stats + (size_t)p * ST_N;
    if (nb > P.cap1) {
      if (tid == 0) { P.counts[p * 4 + 3] = TDA_ERR_CAPACITY; P.counts[p * 4 + 1] = 0; }
      return;
    }
    p_valid = false;   // Pm belongs to the previous cloud's rank matrix
    sort_blist(bl, nb);
    for (int i = tid; i < P.hcap; i += kSweepThreads) hkeys[i] = kEmpty;
    for (int i = tid; i < W; i += kSweepThreads) touched[i] = 0;
    if (tid == 0) {
      S.vcount = 0; S.vsel = 0; S.abort_flag = 0; S.dirty[0] = S.dirty[1] = 0; S.simple[0] = S.simple[1] = 0; S.both[0] = S.both[1] = 0; S.patched = 0; S.flipmask = 0;
      for (int q = 0; q < kGroupRows; ++q) S.conf[0][q] = S.conf[1][q] = 0;
      S.additions = 0; S.pivots = 0; S.heavy_rows = 0; S.restarts = 0; S.groups = 0;
    }
    __threadfence();
    __syncthreads();
    int nrows = 0;
    int64_t vpool_used = 0;
    unsigned long long maxv = 0, rows_swept = 0, badd_edges = 0;
    long long cyc[6] = {0, 0, 0, 0, 0, 0};
    long long t0;
    float* out = P.h1_pairs + (size_t)p * P.cap1 * 2;
    int64_t* outs = P.h1_simplex ? P.h1_simplex + (size_t)p * P.cap1 * 2 : nullptr;

    for (int ci = nb - 1; ci >= 0; --ci) {
      const int rbirth = bl[ci];
      {  // V = {birth edge}
        const uint32_t en = __ldg(&EN[rbirth]);
        const uint32_t c = en >> 16, d = en & 0xffffu;
        if (tid == 0) {
          x_flip(c, d);
          v_toggle(rbirth);
          touched[c >> 5] |= 1u << (c & 31);
          touched[d >> 5] |= 1u << (d & 31);
        }
        __threadfence();
        __syncthreads();
      }
      bool essential = false;
      uint64_t pivot = 0;
      uint32_t pos = (uint32_t)rbirth + 1;
      uint32_t vchunk_pos = 0xffffffffu;   // (verify mode) start of the chunk the flip list belongs to
      p_mode = false;
      for (;;) {
        if (S.abort_flag) break;
        if (pos >= (uint32_t)T) { essential = true; break; }
        // ---- streaming filter over the next kChunkRows rows
        t0 = clock64();
        if (p_mode) {
          for (int i2 = tid; i2 < n; i2 += kSweepThreads) vhead[i2] = 0xffffffffu;
          __syncthreads();
        }
        const uint32_t row = pos + tid;
        bool heavy = false;
        if (row < (uint32_t)T) {
          const uint2 ea = __ldg(&EA[row]);
          S.chunk_ea[tid] = ea;
          const uint32_t c = ea.x >> 16, d = ea.x & 0xffffu;
          heavy = tbit(c) || tbit(d);
          if (p_mode) {
            if (row >= p_pos) {  // the edge enters Pm (exact state = end of this chunk; the lune producer compensates)
              atomicOr(&Pm[(size_t)c * W + (d >> 5)], 1u << (d & 31));
              atomicOr(&Pm[(size_t)d * W + (c >> 5)], 1u << (c & 31));
            }
            vnext[2 * tid] = atomicExch(&vhead[c], 2u * tid);        // endpoint slots of this row into the per-vertex lists
            vnext[2 * tid + 1] = atomicExch(&vhead[d], 2u * tid + 1);
          }
        }
        if (p_mode) p_pos = max(p_pos, min(pos + (uint32_t)kChunkRows, (uint32_t)T));
        __threadfence();
        const unsigned bal = __ballot_sync(0xffffffffu, heavy);
        if (lane == 0) S.wcnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t off = 0, total = 0;
#pragma unroll
        for (int w2 = 0; w2 < kSweepWarps; ++w2) {
          const uint32_t cw = S.wcnt[w2];
          if (w2 < warp) off += cw;
          total += cw;
        }
        if (heavy) S.heavy[off + __popc(bal & ((1u << lane) - 1))] = (uint32_t)tid;
        __syncthreads();
        rows_swept += min((uint32_t)kChunkRows, (uint32_t)T - pos);
        cyc[0] += clock64() - t0;
        const uint32_t nh = total;
        if (!p_mode && nh >= (uint32_t)kDenseThreshold) {
          // the column has become dense: from here on keep Pm (2 atomics per streamed row) and read the lune masks from
          // it (1 KB per heavy row instead of two rank rows); bring Pm to the end of this chunk first
          t0 = clock64();
          p_move(pos);
          const uint32_t cend = min(pos + (uint32_t)kChunkRows, (uint32_t)T);
          p_set_rows(pos, cend, true);
          p_pos = cend;
          p_mode = true;
          __threadfence();
          __syncthreads();
          cyc[5] += clock64() - t0;
          continue;   // filter this chunk again in dense mode (builds the per-vertex row lists)
        }
        if constexpr (VERIFY) {
          // ---- substitute, then verify (CPU model: oracle/rips_propagate_model.cpp, windowed variant).
          // (A) substitution: for an apparent row M=(c,d) with apex a the reduction ends with x_M = x_(c,a) ^ x_(d,a), whatever
          //     happened before; one thread per heavy row, a row whose parent edge is a row of this chunk waits for it.
          if (pos != vchunk_pos) {
            vchunk_pos = pos;
            if (tid == 0) S.nflip = 0;
          }
          if (tid < kChunkRows / 32) S.sub_done[tid] = 0xffffffffu;
          if (tid == 0) S.new_touch = 0;
          __syncthreads();
          t0 = clock64();
          int apex_j = -3, dep_a = -1, dep_b = -1;
          uint32_t hj = 0, cj = 0, dj = 0;
          bool pending = false;
          if ((uint32_t)tid < nh) {
            hj = S.heavy[tid];
            const uint2 eaj = S.chunk_ea[hj];
            cj = eaj.x >> 16; dj = eaj.x & 0xffffu;
            apex_j = (int)eaj.y;
            if (apex_j >= 0) {
              const int ra = __ldg(&R[(size_t)cj * n + apex_j]), rb = __ldg(&R[(size_t)dj * n + apex_j]);
              if (ra >= (int)pos) dep_a = ra - (int)pos;
              if (rb >= (int)pos) dep_b = rb - (int)pos;
              pending = true;
              atomicAnd(&S.sub_done[hj >> 5], ~(1u << (hj & 31)));
            }
          }
          __syncthreads();
          for (;;) {
            bool can = pending;
            if (can && dep_a >= 0) can = (S.sub_done[dep_a >> 5] >> (dep_a & 31)) & 1u;
            if (can && dep_b >= 0) can = (S.sub_done[dep_b >> 5] >> (dep_b & 31)) & 1u;
            if (can) {
              const uint32_t xa = (__ldcg(&X[(size_t)cj * W + ((uint32_t)apex_j >> 5)]) >> (apex_j & 31)) & 1u;
              const uint32_t xb = (__ldcg(&X[(size_t)dj * W + ((uint32_t)apex_j >> 5)]) >> (apex_j & 31)) & 1u;
              const uint32_t cur = (__ldcg(&X[(size_t)cj * W + (dj >> 5)]) >> (dj & 31)) & 1u;
              if ((xa ^ xb) != cur) {
                x_flip(cj, dj);
                v_toggle((int)(pos + hj));
                const uint32_t oc = atomicOr(&touched[cj >> 5], 1u << (cj & 31));
                const uint32_t od = atomicOr(&touched[dj >> 5], 1u << (dj & 31));
                if (!((oc >> (cj & 31)) & 1u) || !((od >> (dj & 31)) & 1u)) S.new_touch = 1;
                const uint32_t fi = atomicAdd(&S.nflip, 1u);
                if (fi < (uint32_t)kChunkRows) S.fliprow[fi] = (uint16_t)hj;
                else S.abort_flag = TDA_ERR_CAPACITY;   // (a row flips at most once per pass; repeated passes could exceed the list)
              }
              pending = false;
            }
            __threadfence();   // the flips are performed before another thread reads these X words
            __syncthreads();
            if (can) atomicOr(&S.sub_done[hj >> 5], 1u << (hj & 31));
            if (__syncthreads_count(pending) == 0) break;
          }
          cyc[2] += clock64() - t0;
          if (S.new_touch) {   // rows filtered out as untouched may matter now: filter this chunk again (substitution is idempotent)
            if (tid == 0) S.restarts += 1;
            __syncthreads();
            continue;
          }
          // (B) verification: every heavy row, formed from the X that holds all the flips of the chunk, must be empty
          t0 = clock64();
          bool failed = false;
          uint32_t fi0 = 0;
          int fs = 0;
          for (uint32_t i0 = 0; i0 < nh && !failed; i0 += kGroupRows) {
            const int ns = (int)min((uint32_t)kGroupRows, nh - i0);
            produce(0, pos, i0, ns, 0, kSweepWarps);
            __syncthreads();
            const uint32_t dirty = S.dirty[0];
            if (dirty) {
              failed = true; fi0 = i0; fs = __ffs(dirty) - 1;
              if (warp == 0) {   // the first non-empty row: its highest vertex is the next non-apparent pivot
                const uint32_t* r = Sr + (size_t)fs * W;
                int best = -1;
#pragma unroll
                for (int q = 0; q < WPL; ++q) {
                  const uint32_t rwq = (lane + 32 * q) < W ? r[lane + 32 * q] : 0u;
                  if (rwq) best = (lane + 32 * q) * 32 + 31 - __clz(rwq);
                }
                best = __reduce_max_sync(0xffffffffu, best);
                if (lane == 0) S.fail_w = best;
              }
            }
            __syncthreads();
            if (tid == 0) { S.dirty[0] = 0; S.simple[0] = 0; S.both[0] = 0; S.groups += 1; S.heavy_rows += failed ? (uint32_t)(fs + 1) : (uint32_t)ns; }
            __syncthreads();
          }
          cyc[1] += clock64() - t0;
          if (!failed) {
            if (tid == 0) S.additions += S.nflip;
            pos += kChunkRows;
            __syncthreads();
            continue;
          }
          // (C) event at row srow: the flips above it are undone (they were made with an X that the event changes)
          t0 = clock64();
          const uint32_t srow = pos + S.heavy[fi0 + fs];
          const int fw = S.fail_w;
          const uint32_t nfl = min(S.nflip, (uint32_t)kChunkRows);
          uint32_t kept = 0;
          for (uint32_t f = tid; f < nfl; f += kSweepThreads) {
            const uint32_t hf = S.fliprow[f];
            if (pos + hf > srow) {
              const uint2 eaf = S.chunk_ea[hf];
              x_flip(eaf.x >> 16, eaf.x & 0xffffu);
              v_toggle((int)(pos + hf));
            } else {
              ++kept;
            }
          }
          (void)kept;
          __syncthreads();
          if (tid == 0) {
            uint32_t k2 = 0;
            for (uint32_t f = 0; f < nfl; ++f) k2 += (pos + S.fliprow[f] <= srow) ? 1u : 0u;
            S.additions += k2;
            S.pivots += 1;
          }
          __threadfence();
          __syncthreads();
          const uint64_t fkey = (uint64_t)srow * (uint64_t)n + (uint64_t)(n - 1 - fw);
          const int fowner = hash_find(fkey);
          if (fowner < 0) { pivot = fkey; cyc[4] += clock64() - t0; break; }   // death
          {
            const int64_t vs = P.vstart[(size_t)p * P.cap1 + fowner];
            const int vn = P.vlen[(size_t)p * P.cap1 + fowner];
            const uint32_t* ov = P.vpool + (size_t)p * P.vpool_cap + vs;
            if (S.vcount + (uint32_t)vn > (uint32_t)P.vcap) v_compact();
            for (int q = tid; q < vn; q += kSweepThreads) {
              const int re = (int)ov[q];
              const uint32_t en = __ldg(&EN[re]);
              const uint32_t c = en >> 16, d = en & 0xffffu;
              x_flip(c, d);
              v_toggle(re);
              atomicOr(&touched[c >> 5], 1u << (c & 31));
              atomicOr(&touched[d >> 5], 1u << (d & 31));
            }
            badd_edges += vn;
            if (tid == 0) { S.additions += 1; S.restarts += 1; }
            __threadfence();
            __syncthreads();
          }
          pos = srow;   // this row again: the pivot just handled is even now, lower vertices may remain (its substitution is a no-op)
          cyc[3] += clock64() - t0;
          __syncthreads();
          continue;
        }
        uint32_t i = 0;
        bool restart = false, done = false;
        uint32_t newpos = pos + kChunkRows;
        int buf = 0;
        if (nh) {
          t0 = clock64();
          produce(0, pos, 0, (int)min((uint32_t)kGroupRows, nh), 0, kSweepWarps);
          __syncthreads();
          conflicts(0, (int)min((uint32_t)kGroupRows, nh));
          __syncthreads();
          cyc[1] += clock64() - t0;
        }
        while (i < nh) {
          const int ns = (int)min((uint32_t)kGroupRows, nh - i);
          const int ns_next = (int)min((uint32_t)kGroupRows, nh - i - ns);
          t0 = clock64();
          if (warp == 0) resolve(buf, ns);
          else if (ns_next > 0) produce(buf ^ 1, pos, i + ns, ns_next, 1, kSweepWarps - 1);
          __syncthreads();
          cyc[2] += clock64() - t0;
          t0 = clock64();
          if (warp == 0) write_flips(buf);
          if (S.res_status == SW_DONE && ns_next > 0) { patch_next(buf, ns_next); conflicts(buf ^ 1, ns_next); }
          __threadfence();   // the flips are performed before anybody reads X again
          __syncthreads();
          cyc[4] += clock64() - t0;
          const int status = S.res_status;
          const int s = S.res_s;
          if (tid == 0) { S.groups += 1; S.heavy_rows += (status == SW_DONE) ? ns : s + 1; }
          if (status == SW_DONE) { i += ns; buf ^= 1; continue; }
          if (tid == 0) { S.dirty[buf ^ 1] = 0; S.simple[buf ^ 1] = 0; S.both[buf ^ 1] = 0; S.patched = 0; }  // the group produced ahead is dropped
          if (tid < kGroupRows) S.conf[buf ^ 1][tid] = 0;
          const uint32_t srow = pos + S.heavy[i + s];
          if (status == SW_RESTART) { newpos = srow + 1; restart = true; break; }
          if (status == SW_REDUCED) {
            t0 = clock64();
            const int owner = S.res_owner;
            const int64_t vs = P.vstart[(size_t)p * P.cap1 + owner];
            const int vn = P.vlen[(size_t)p * P.cap1 + owner];
            const uint32_t* ov = P.vpool + (size_t)p * P.vpool_cap + vs;
            if (S.vcount + (uint32_t)vn > (uint32_t)P.vcap) v_compact();
            for (int q = tid; q < vn; q += kSweepThreads) {
              const int re = (int)ov[q];
              const uint32_t en = __ldg(&EN[re]);
              const uint32_t c = en >> 16, d = en & 0xffffu;
              x_flip(c, d);
              v_toggle(re);
              atomicOr(&touched[c >> 5], 1u << (c & 31));
              atomicOr(&touched[d >> 5], 1u << (d & 31));
            }
            badd_edges += vn;
            __threadfence();
            __syncthreads();
            newpos = srow;  // recompute this row: the pivot just handled is now even, lower vertices may remain
            restart = true;
            cyc[3] += clock64() - t0;
            break;
          }
          // SW_DEATH
          pivot = (uint64_t)srow * (uint64_t)n + (uint64_t)(n - 1 - S.res_w);
          done = true;
          break;
        }
        if (done) break;
        if (restart && tid == 0) S.restarts += 1;
        pos = newpos;
        __syncthreads();
      }
      if (S.abort_flag) break;
      t0 = clock64();
      // finalise the column
      v_compact();
      const uint32_t nv = S.vcount;
      if (nv > maxv) maxv = nv;
      if (!essential) {
        if (vpool_used + nv > P.vpool_cap) { if (tid == 0) S.abort_flag = TDA_ERR_CAPACITY; __syncthreads(); break; }
        uint32_t* dst = P.vpool + (size_t)p * P.vpool_cap + vpool_used;
        const uint32_t* list = vlist(S.vsel);
        for (uint32_t i = tid; i < nv; i += kSweepThreads) dst[i] = list[i];
        if (tid == 0) {
          P.vstart[(size_t)p * P.cap1 + ci] = vpool_used;
          P.vlen[(size_t)p * P.cap1 + ci] = (int)nv;
          hash_insert(pivot, ci);
        }
        vpool_used += nv;
      }
      const float birth = SD[rbirth];
      float death = INFINITY;
      int Md = -1, wd = -1;
      if (!essential) { Md = (int)(pivot / (uint64_t)n); wd = n - 1 - (int)(pivot % (uint64_t)n); death = SD[Md]; }
      if (essential || death > birth) {
        if (tid == 0) {
          out[2 * nrows] = birth; out[2 * nrows + 1] = death;
          if (outs) {
            const uint32_t e = EN[rbirth];
            outs[2 * nrows] = edge_index((int)(e >> 16), (int)(e & 0xffffu));
            if (essential) outs[2 * nrows + 1] = -1;
            else {
              const uint32_t em = EN[Md];
              int x = (int)(em >> 16), y = (int)(em & 0xffffu), z = wd, t;
              if (x < y) { t = x; x = y; y = t; }
              if (y < z) { t = y; y = z; z = t; }
              if (x < y) { t = x; x = y; y = t; }
              outs[2 * nrows + 1] = (int64_t)x * (x - 1) * (x - 2) / 6 + (int64_t)y * (y - 1) / 2 + z;
            }
          }
        }
        ++nrows;
      }
      v_clear();
      cyc[5] += clock64() - t0;
    }
    __syncthreads();
    if (S.abort_flag) {  // leave the scratch clean for the next problem
      v_compact();
      v_clear();
      for (size_t i = tid; i < (size_t)n * W; i += kSweepThreads) X[i] = 0;
      __threadfence();
      __syncthreads();
    }
    if (tid == 0) {
      P.counts[p * 4 + 1] = nrows;
      P.counts[p * 4 + 3] = S.abort_flag;
      st[ST_REDUCED] = (unsigned long long)nb;
      st[ST_ADDITIONS] = S.additions;
      st[ST_PUSHES] = rows_swept;      // rows streamed through the filter
      st[ST_POPS] = S.pivots;
      st[ST_EXTENSIONS] = S.restarts;
      st[ST_MAXV] = maxv;
      for (int q = 0; q < 6; ++q) st[ST_CYC_EXTRACT + q] = (unsigned long long)cyc[q];
      st[ST_BADD_EDGES] = badd_edges;
      st[ST_EXT_EDGES] = S.heavy_rows;
      st[ST_CYC_OWNER] = S.groups;   // (diagnostic) number of row groups
    }
    __syncthreads();
  }
};

template <int WPL, bool VERIFY = false>
__global__ void __launch_bounds__(kSweepThreads, 1) rips_sweep_kernel(const __grid_constant__ ReduceParams P) {
  __shared__ SweepSmem S;
  extern __shared__ uint32_t sweep_dyn[];
  Sweeper<WPL, VERIFY> sw(P, S, sweep_dyn);
  for (;;) {
    if (threadIdx.x == 0) S.problem = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int p = S.problem;
    __syncthreads();
    if (p >= P.batch) break;
    sw.run_problem(p);
  }
}

#include "rips_sweep2.cuh"

// ------------------------------------------------------------------------------------------------
// H2 pre-pass (runs after the H1 stage on the same rank matrix)
//  h2_clear_kernel    : bitset of the triangles that are H1 death simplices (apparent pairs (r, apex[r]) and the pivots of the
//                       reduced H1 columns): they are not H2 columns (clearing)
//  h2_apparent_kernel : one warp per edge M walks the triangles (M, w), w in lune(M): apex4 = the largest vertex v > w of the lune
//                       with rank(w,v) < M, i.e. the oldest cofacet of the triangle has the same diameter and the triangle is
//                       its youngest facet -- an apparent (zero-persistence) pair; triangles that are neither cleared nor
//                       apparent are the residual H2 columns.
__global__ void h2_clear_kernel(const uint2* __restrict__ ea, const int* __restrict__ Tarr, const uint64_t* __restrict__ hkeys, int hcap,
                                int n, int64_t E, uint32_t* __restrict__ cbits, int64_t cwords) {
  const int p = blockIdx.y;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t* cb = cbits + (size_t)p * cwords;
  if (t < Tarr[p]) {
    const int apex = (int)ea[(size_t)p * E + t].y;   // -2 for MST edges, -1 for an empty lune
    if (apex >= 0) {
      const uint64_t k = (uint64_t)t * (uint64_t)n + (uint64_t)(n - 1 - apex);
      if (TDA_BC(apex < n, "h2_clear apex/n", apex, n)) atomicOr(&cb[k >> 5], 1u << (k & 31));
    }
  }
  if (t < hcap) {
    const uint64_t k = hkeys[(size_t)p * hcap + t];
    if (k != ~0ull && TDA_BC((int64_t)(k >> 5) < cwords, "h2_clear pivot key/cwords", k, cwords)) atomicOr(&cb[k >> 5], 1u << (k & 31));
  }
}

constexpr int kH2MaxWords = 64;   // n <= 2048 (two lune words per lane; triangle keys E*n < 2^32)
__global__ void __launch_bounds__(256) h2_apparent_kernel(const int* __restrict__ rank, const uint32_t* __restrict__ ends, const int* __restrict__ Tarr,
                                                          int n, int64_t E, const uint32_t* __restrict__ cbits, int64_t cwords,
                                                          short* __restrict__ apex4, int* __restrict__ blist2, int* __restrict__ bcount2, int cap2) {
  const int p = blockIdx.y;
  const int64_t M = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (M >= Tarr[p]) return;
  const int W = (n + 31) >> 5;
  const int* R = rank + (size_t)p * n * n;
  const uint32_t e = ends[(size_t)p * E + M];
  const int x = (int)(e >> 16), y = (int)(e & 0xffffu);
  const int* Rx = R + (size_t)x * n;
  const int* Ry = R + (size_t)y * n;
  uint32_t lunew0 = 0, lunew1 = 0;   // lane k holds words k and k + 32 of the lune mask
  for (int k = 0; k < W; ++k) {
    const int v = 32 * k + lane;
    const bool in = v < n && Rx[v] < (int)M && Ry[v] < (int)M;
    const unsigned b = __ballot_sync(0xffffffffu, in);
    if (lane == (k & 31)) { if (k < 32) lunew0 = b; else lunew1 = b; }
  }
  auto lune_word = [&](int k) -> uint32_t {   // warp-uniform k
    const uint32_t a0 = __shfl_sync(0xffffffffu, lunew0, k & 31), a1 = __shfl_sync(0xffffffffu, lunew1, k & 31);
    return k < 32 ? a0 : a1;
  };
  const uint32_t* cb = cbits + (size_t)p * cwords;
  short* A4 = apex4 + (size_t)p * (size_t)E * (size_t)n;
  for (int k = W - 1; k >= 0; --k) {
    uint32_t word = lune_word(k);
    while (word) {
      const int bit = 31 - __clz(word);
      word &= ~(1u << bit);
      const int w = 32 * k + bit;
      const int* Rw = R + (size_t)w * n;
      int found = -1;
      for (int k2 = W - 1; k2 >= (w >> 5); --k2) {
        const uint32_t lw = lune_word(k2);
        const int v = 32 * k2 + lane;
        const bool ok = ((lw >> lane) & 1u) && v > w && Rw[v] < (int)M;
        const unsigned b = __ballot_sync(0xffffffffu, ok);
        if (b) { found = 32 * k2 + 31 - __clz(b); break; }
      }
      if (lane == 0) {
        const uint64_t k3 = (uint64_t)M * (uint64_t)n + (uint64_t)(n - 1 - w);
        A4[k3] = (short)found;
        const bool cleared = (cb[k3 >> 5] >> (k3 & 31)) & 1u;
        if (!cleared && found < 0) {
          const int pos = atomicAdd(&bcount2[p], 1);
          if (pos < cap2) blist2[(size_t)p * cap2 + pos] = (int)k3;
        }
      }
    }
  }
}

__global__ void finalize_stats_kernel(const int* __restrict__ T, const int* __restrict__ bcount, int n, int batch,
                                      unsigned long long* __restrict__ stats, const int32_t* __restrict__ counts) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= batch) return;
  long long nm = 0;
  // number of finite H0 rows + zero-length merges is not tracked; columns = T - (#MST edges) is
  // reported approximately as T - (n - components); components = #inf rows is unknown here, use counts
  (void)counts;
  (void)nm;
  stats[(size_t)p * ST_N + ST_COLUMNS] = (unsigned long long)max(T[p] - (n - 1), 0);
  stats[(size_t)p * ST_N + ST_APPARENT] = (unsigned long long)max(T[p] - (n - 1) - bcount[p], 0);
}

// ------------------------------------------------------------------------------------------------
struct Layout {
  uint64_t *keys_a, *keys_b;
  uint32_t *vals_a, *vals_b;
  void* cub_tmp; size_t cub_bytes;
  int* rank; uint32_t* ends; float* sdist; uint32_t* thresh_bits; int* T;
  uint8_t* mst; int* mstlist; int* mstcount; int* apex; uint2* ea; int* blist; int* bcount;
  uint32_t *comp, *parent, *cbest; int* done;
  uint32_t* vbits; int64_t vwords; uint32_t* vlist; int64_t vcap;
  uint64_t* hkeys; int* hvals; int hcap;
  uint32_t* vpool; int64_t vpool_cap; int64_t* vstart; int* vlen;
  uint32_t* bits; uint64_t wbits;   // per-CTA key windows (bitset reducer)
  uint32_t* xmat; int xw;           // per-CTA V bit matrix (sweep reducer)
  uint32_t* pmat;                   // per-CTA "edges below the cursor" bit matrix (sweep reducer)
  bool sweep;                       // one of the row-sweep reducers (X / Pm bit matrices) rather than the key bitset
  int reducer;                      // 0 sweep2, 1 sweep (resolver), 2 sweep (verify), 3 bitset
  uint2* par;                       // parents of the apparent edges (sweep2)
  uint2* s2_pend; uint2* s2_heavy; uint32_t* s2_rec; int* s2_lists; int s2_nrec; int s2_wmax;
  int* work_counter; unsigned long long* stats;
  int grid; size_t total;
};

static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

static int s2_wmax_option() {
  long long w = option("rips_wmax");
  if (w < 64) w = 64;
  if (w > kS2MaxWindow) w = kS2MaxWindow;
  return (int)(w & ~31ll);
}

static Layout make_layout(void* ws, int n, int batch, int maxdim, int cap1, size_t pool_bytes, int sm_count) {
  Layout L;
  memset(&L, 0, sizeof(L));
  L.reducer = (int)option("rips_reducer");
  if (L.reducer < 0 || L.reducer > 3) L.reducer = 0;
  if ((L.reducer == 1 || L.reducer == 2) && n > 8192) L.reducer = 0;   // the resolver keeps <= 8 words of a row per lane
  L.sweep = L.reducer != 3;
  const int64_t E = (int64_t)n * (n - 1) / 2;
  const int64_t BE = (int64_t)batch * E;
  Carver c(ws, ~size_t(0));
  L.keys_a = c.take<uint64_t>(BE);
  L.keys_b = c.take<uint64_t>(BE);
  L.vals_a = c.take<uint32_t>(BE);
  L.vals_b = c.take<uint32_t>(BE);
  L.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, L.cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, BE > 0 ? BE : 1, 0, 64, (cudaStream_t)0);
  L.cub_tmp = c.take<char>(L.cub_bytes + 256);
  L.rank = c.take<int>((int64_t)batch * n * n);
  L.ends = c.take<uint32_t>(BE);
  L.sdist = c.take<float>(BE);
  L.thresh_bits = c.take<uint32_t>(batch);
  L.T = c.take<int>(batch);
  L.mst = c.take<uint8_t>(BE);
  L.mstlist = c.take<int>((int64_t)batch * next_pow2(n));
  L.mstcount = c.take<int>(batch);
  L.comp = c.take<uint32_t>((int64_t)batch * n);
  L.parent = c.take<uint32_t>((int64_t)batch * n);
  L.cbest = c.take<uint32_t>((int64_t)batch * n);
  L.done = c.take<int>(batch);
  L.stats = c.take<unsigned long long>((int64_t)batch * ST_N);
  L.work_counter = c.take<int>(1);
  if (maxdim >= 1) {
    L.apex = c.take<int>(BE);
    L.ea = c.take<uint2>(BE);
    L.blist = c.take<int>((int64_t)batch * cap1);
    L.bcount = c.take<int>(batch);
    L.grid = batch < 2 * sm_count ? batch : 2 * sm_count;
    if (L.grid < 1) L.grid = 1;
    L.vwords = (E + 31) / 32 + 1;
    L.vbits = c.take<uint32_t>((int64_t)L.grid * L.vwords);
    L.vcap = E + 1024;
    L.vlist = c.take<uint32_t>((int64_t)L.grid * 2 * L.vcap);
    L.hcap = next_pow2(2 * cap1);
    L.hkeys = c.take<uint64_t>((int64_t)batch * L.hcap);
    L.hvals = c.take<int>((int64_t)batch * L.hcap);
    L.vstart = c.take<int64_t>((int64_t)batch * cap1);
    L.vlen = c.take<int>((int64_t)batch * cap1);
    L.xw = ((n + 31) / 32 + 1) & ~1;   // even: rows of X / Pm are 8-byte aligned
    if (L.reducer == 0) {
      L.par = c.take<uint2>(BE);
      L.s2_wmax = s2_wmax_option();
      L.s2_pend = c.take<uint2>((size_t)L.grid * 2 * (size_t)(L.s2_wmax + 64));
      L.s2_heavy = c.take<uint2>((size_t)L.grid * (size_t)(L.s2_wmax + 64));
      L.s2_nrec = 4 * cap1 + 64;
      L.s2_rec = c.take<uint32_t>((size_t)L.grid * (size_t)L.s2_nrec * kWcRec);
      L.s2_lists = c.take<int>((size_t)L.grid * (size_t)(3 * L.s2_nrec + cap1));
    }
    if (L.sweep) {
      L.xmat = c.take<uint32_t>((size_t)L.grid * (size_t)n * L.xw);
      L.pmat = c.take<uint32_t>((size_t)L.grid * (size_t)n * L.xw);
      L.vpool_cap = (int64_t)(pool_bytes / (size_t)batch / sizeof(uint32_t));
      if (L.vpool_cap < 4 * (int64_t)cap1) L.vpool_cap = 4 * (int64_t)cap1;
      L.vpool = c.take<uint32_t>((int64_t)batch * L.vpool_cap);
    } else {
    // key windows: the whole key space E*n when it fits in 2^32 bits and in the pool, else a sliding window
    const uint64_t page = 1ull << kPageShift;
    uint64_t want = ((uint64_t)E * (uint64_t)n + 32 * page - 1) / (32 * page) * (32 * page);
    if (want > (1ull << 32)) want = 1ull << 32;
    uint64_t afford = (uint64_t)(pool_bytes / 2 / (size_t)L.grid) * 8 / (32 * page) * (32 * page);
    if (afford < 32 * page) afford = 32 * page;
    L.wbits = want < afford ? want : afford;
    L.bits = c.take<uint32_t>((size_t)L.grid * (L.wbits >> 5));
    // reduction columns (V) of finished columns: whatever the pool leaves, at least 4 * cap1 entries per cloud
    const size_t bits_bytes = (size_t)L.grid * (L.wbits >> 3);
    L.vpool_cap = pool_bytes > bits_bytes ? (int64_t)((pool_bytes - bits_bytes) / (size_t)batch / sizeof(uint32_t)) : 0;
    if (L.vpool_cap < 4 * (int64_t)cap1) L.vpool_cap = 4 * (int64_t)cap1;
    L.vpool = c.take<uint32_t>((int64_t)batch * L.vpool_cap);
    }
  }
  L.total = c.off;
  return L;
}

static int sm_count_cached() {
  static int v = 0;
  if (!v) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
  }
  return v;
}

}  // namespace rips
}  // namespace tda

using namespace tda;
using namespace tda::rips;

extern "C" int tda_pdist_lowdim(const float* pts, int n, int d, int batch, float* dm, void* stream) {
  if (!pts || !dm || n <= 0 || d <= 0 || d > 64 || batch <= 0) return set_error(TDA_ERR_INVALID, "tda_pdist_lowdim: bad arguments");
  dim3 block(32, 8), grid((n + 31) / 32, (n + 7) / 8, batch);
  StageScope st(STAGE_RIPS_PDIST, (cudaStream_t)stream);
  pdist_lowdim_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(pts, n, d, dm);
  count_launch();
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}

extern "C" size_t tda_rips_workspace_bytes(int n, int batch, int maxdim, int cap1, size_t pool_bytes) {
  if (n <= 0 || batch <= 0 || cap1 <= 0) return 0;
  Layout L = make_layout(nullptr, n, batch, maxdim, next_pow2(cap1), pool_bytes, 148);
  return L.total + 4096;
}

// the clouds of a call as SUBSETS of one parent cloud whose edges are already sorted (tda_rips_sort_edges): see subset_*_kernel
struct SubsetInput { const uint32_t* ends; const float* sdist; const float* dm; int n_parent; const int32_t* idx; };

static int rips_enqueue(const float* dm, int n, int batch, int maxdim, float thresh, float* h0_pairs, int64_t* h0_simplex,
                        float* h1_pairs, int64_t* h1_simplex, int cap1, int32_t* counts, float* thresh_out, void* ws,
                        size_t ws_bytes, size_t pool_bytes, void* stream_, const SubsetInput* sub = nullptr) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if ((!dm && !sub) || !h0_pairs || !counts || !ws || n <= 0 || batch <= 0) return set_error(TDA_ERR_INVALID, "tda_rips: bad arguments");
  if (maxdim < 0) return set_error(TDA_ERR_INVALID, "tda_rips: maxdim < 0");
  if (maxdim > 1) return set_error(TDA_ERR_UNSUPPORTED, "tda_rips: maxdim=%d not implemented (H0/H1 only)", maxdim);
  if (n > 65535) return set_error(TDA_ERR_UNSUPPORTED, "tda_rips: n=%d > 65535", n);
  if (maxdim >= 1 && (!h1_pairs || cap1 <= 0)) return set_error(TDA_ERR_INVALID, "tda_rips: h1 buffers missing");
  if (batch > 65535) return set_error(TDA_ERR_INVALID, "tda_rips: batch > 65535 (chunk the call)");
  const int cap1p = next_pow2(cap1 > 0 ? cap1 : 1);
  if (cap1p != cap1 && maxdim >= 1) return set_error(TDA_ERR_INVALID, "tda_rips: cap1 must be a power of two");
  const int sms = sm_count_cached();
  Layout L = make_layout(ws, n, batch, maxdim, cap1p, pool_bytes, sms);
  if (L.total > ws_bytes) return set_error(TDA_ERR_WORKSPACE, "tda_rips: workspace %zu < required %zu", ws_bytes, L.total);
  const int64_t E = (int64_t)n * (n - 1) / 2;
  const int64_t BE = (int64_t)batch * E;

  stage_begin_if(STAGE_RIPS_SORT, stream);
  TDA_CUDA_CHECK(cudaMemsetAsync(L.T, 0, sizeof(int) * batch, stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.stats, 0, sizeof(unsigned long long) * batch * ST_N, stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.work_counter, 0, sizeof(int), stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 4 * batch, stream));
  enclosing_init_kernel<<<(batch + 255) / 256, 256, 0, stream>>>(L.thresh_bits, batch, thresh);
  count_launch();
  if (isinf(thresh)) {
    dim3 g((n * 32 + 255) / 256, batch);
    if (sub) subset_enclosing_kernel<<<g, 256, 0, stream>>>(sub->dm, sub->n_parent, sub->idx, n, L.thresh_bits);
    else enclosing_kernel<<<g, 256, 0, stream>>>(dm, n, L.thresh_bits);
    count_launch();
  }
  TDA_LAUNCH_CHECK();
  if (thresh_out) TDA_CUDA_CHECK(cudaMemcpyAsync(thresh_out, L.thresh_bits, sizeof(float) * batch, cudaMemcpyDeviceToDevice, stream));
  if (sub && E > 0) {
    // flag + prefix count over the parent's sorted edges; the chunk counts live in the (unused) sort buffers
    const int64_t Ep = (int64_t)sub->n_parent * (sub->n_parent - 1) / 2;
    const int64_t nchunks64 = (Ep + kSubChunk - 1) / kSubChunk;
    if (nchunks64 > 65535 * 32ll) return set_error(TDA_ERR_UNSUPPORTED, "tda_rips_subsets: parent of %d points is too large", sub->n_parent);
    const int nchunks = (int)nchunks64;
    int* chunk_cnt = reinterpret_cast<int*>(L.keys_a);
    if (sizeof(int) * (size_t)batch * nchunks > sizeof(uint64_t) * (size_t)BE)
      return set_error(TDA_ERR_WORKSPACE, "tda_rips_subsets: subsets of %d points are too small for a parent of %d (use tda_rips)", n, sub->n_parent);
    const size_t map_bytes = sizeof(uint32_t) * (size_t)((sub->n_parent + 1) / 2);
    if (map_bytes > 48 * 1024) {
      TDA_CUDA_CHECK(cudaFuncSetAttribute(subset_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)map_bytes));
      TDA_CUDA_CHECK(cudaFuncSetAttribute(subset_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)map_bytes));
    }
    dim3 gc((unsigned)nchunks, batch);
    subset_count_kernel<<<gc, kSubThreads, map_bytes, stream>>>(sub->ends, sub->sdist, Ep, sub->idx, n, sub->n_parent, L.thresh_bits, chunk_cnt, nchunks, L.T);
    subset_scan_kernel<<<batch, 512, 0, stream>>>(chunk_cnt, nchunks);
    subset_scatter_kernel<<<gc, kSubThreads, map_bytes, stream>>>(sub->ends, sub->sdist, Ep, sub->idx, sub->n_parent, chunk_cnt, nchunks, n, E, L.rank, L.ends, L.sdist);
    count_launch(3);
    TDA_LAUNCH_CHECK();
  } else if (E > 0) {
    dim3 g((unsigned)((E + 255) / 256), batch);
    edge_keys_kernel<<<g, 256, 0, stream>>>(dm, n, E, L.thresh_bits, L.keys_a, L.vals_a, L.T);
    count_launch();
    TDA_LAUNCH_CHECK();
    int pbits = 0;
    while ((1 << pbits) < batch) ++pbits;
    size_t tmp = L.cub_bytes;
    TDA_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(L.cub_tmp, tmp, (const uint64_t*)L.keys_a, L.keys_b, (const uint32_t*)L.vals_a, L.vals_b,
                                                   BE, 0, 32 + pbits, stream));
    count_launch(4 + (pbits + 7) / 8 + 2);
    rank_scatter_kernel<<<g, 256, 0, stream>>>(L.keys_b, L.vals_b, n, E, L.rank, L.ends, L.sdist);
    count_launch();
    TDA_LAUNCH_CHECK();
  }
  {
    dim3 g((n + 255) / 256, batch);
    rank_diag_kernel<<<g, 256, 0, stream>>>(n, L.rank);
    count_launch();
  }
  stage_end_if(STAGE_RIPS_SORT, stream);
  TDA_CUDA_CHECK(cudaMemsetAsync(L.mst, 0, (size_t)(BE > 0 ? BE : 1), stream));
  {
    StageScope st(STAGE_RIPS_H0, stream);
    if (n <= 16384 && option("rips_h0_chunked") != 0) {   // one launch: Boruvka / Kruskal by chunks of the sorted edge list
      const size_t dyn = sizeof(uint32_t) * 3 * (size_t)n;
      if (dyn > 48 * 1024) TDA_CUDA_CHECK(cudaFuncSetAttribute(boruvka_chunked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      boruvka_chunked_kernel<<<batch, 1024, dyn, stream>>>(L.ends, L.T, n, E, L.comp, L.mst, L.mstlist, next_pow2(n), L.mstcount);
      count_launch();
    } else {
      dim3 gi((n + 255) / 256, batch);
      boruvka_init_kernel<<<gi, 256, 0, stream>>>(n, L.comp, L.cbest, L.done, L.mstcount);
      count_launch();
      int rounds = 1;
      while ((1 << rounds) < n) ++rounds;
      ++rounds;  // one extra round detects "no merge" and is a no-op otherwise
      dim3 gs((n + 7) / 8, batch);
      for (int r = 0; r < rounds && E > 0; ++r) {
        boruvka_scan_kernel<<<gs, 256, 0, stream>>>(L.rank, L.T, n, L.comp, L.cbest, L.done);
        boruvka_merge_kernel<<<batch, 1024, 0, stream>>>(L.ends, n, E, L.comp, L.parent, L.cbest, L.mst, L.done, L.mstlist, next_pow2(n), L.mstcount);
        count_launch(2);
      }
    }
    {
      int np2 = 1;
      while (np2 < n) np2 <<= 1;
      const int use_smem = np2 <= 8192;
      h0_emit_kernel<<<batch, 1024, use_smem ? sizeof(int) * (size_t)np2 : 0, stream>>>(L.ends, L.sdist, L.T, n, E, L.mstcount, L.comp, h0_pairs,
                                                                                       h0_simplex, counts, L.mstlist, np2, use_smem);
    }
    count_launch();
    TDA_LAUNCH_CHECK();
  }
  if (maxdim >= 1 && E > 0) {
    TDA_CUDA_CHECK(cudaMemsetAsync(L.bcount, 0, sizeof(int) * batch, stream));
    if (L.sweep) TDA_CUDA_CHECK(cudaMemsetAsync(L.xmat, 0, sizeof(uint32_t) * (size_t)L.grid * (size_t)n * L.xw, stream));
    else TDA_CUDA_CHECK(cudaMemsetAsync(L.bits, 0, (size_t)L.grid * (L.wbits >> 3), stream));
    TDA_CUDA_CHECK(cudaMemsetAsync(L.vbits, 0, sizeof(uint32_t) * (size_t)L.grid * L.vwords, stream));
    dim3 g((unsigned)((E * 32 + 255) / 256), batch);
    {
      StageScope st(STAGE_RIPS_APPARENT, stream);
      if (n <= 12288 && option("rips_apparent_rows") != 0) {
        const size_t dyn = sizeof(int) * (size_t)n;
        if (dyn > 48 * 1024) TDA_CUDA_CHECK(cudaFuncSetAttribute(apparent_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        apparent_rows_kernel<<<dim3((unsigned)n, (unsigned)batch), 256, dyn, stream>>>(L.rank, L.mst, L.T, n, E, L.ea, L.blist, L.bcount, cap1p);
      } else {
        apparent_kernel<<<g, 256, 0, stream>>>(L.rank, L.ends, L.mst, L.T, n, E, L.apex, L.ea, L.blist, L.bcount, cap1p, L.stats);
      }
    }
    count_launch();
    TDA_LAUNCH_CHECK();
    ReduceParams P;
    P.rank = L.rank; P.ends = L.ends; P.sdist = L.sdist; P.T = L.T; P.ea = L.ea; P.blist = L.blist; P.bcount = L.bcount;
    P.n = n; P.E = E; P.batch = batch; P.cap1 = cap1p;
    P.h1_pairs = h1_pairs; P.h1_simplex = h1_simplex; P.counts = counts;
    P.vbits = L.vbits; P.vwords = L.vwords; P.vlist = L.vlist; P.vcap = L.vcap;
    P.hkeys = L.hkeys; P.hvals = L.hvals; P.hcap = L.hcap;
    P.bits = L.bits; P.wbits = L.wbits; P.xmat = L.xmat; P.xw = L.xw; P.pmat = L.pmat;
    P.vpool = L.vpool; P.vpool_cap = L.vpool_cap; P.vstart = L.vstart; P.vlen = L.vlen;
    P.work_counter = L.work_counter; P.stats = L.stats;
    P.apex4 = nullptr; P.far = nullptr; P.far_cap = 0; P.nbk = 0;
    P.verify_mode = L.reducer == 2 ? 1 : 0;
    P.par = L.par; P.s2_pend = L.s2_pend; P.s2_heavy = L.s2_heavy; P.s2_rec = L.s2_rec; P.s2_lists = L.s2_lists; P.s2_nrec = L.s2_nrec;
    P.s2_warp_engine = option("rips_warp_engine") != 0 ? 1 : 0;
    P.s2_debug = (int)option("rips_debug");
    P.s2_wc_max_rows = (int)option("rips_wc_max_rows");
    if (P.s2_wc_max_rows < 1024) P.s2_wc_max_rows = 1024;
    P.s2_wmax = L.s2_wmax;
    P.s2_w0 = (int)option("rips_w0");
    if (P.s2_w0 < 32) P.s2_w0 = 32;
    if (P.s2_w0 > P.s2_wmax) P.s2_w0 = P.s2_wmax;
    P.s2_wsparse = (int)option("rips_wsparse");
    if (P.s2_wsparse < P.s2_w0) P.s2_wsparse = P.s2_w0;
    if (P.s2_wsparse > P.s2_wmax) P.s2_wsparse = P.s2_wmax;
    P.s2_dense_min = (int)option("rips_dense_min");
    if (P.s2_dense_min < 1) P.s2_dense_min = 1;
    P.s2_dense_div = (int)option("rips_dense_div");
    if (P.s2_dense_div < 1) P.s2_dense_div = 1;
    if (L.reducer == 0) {
      StageScope st(STAGE_RIPS_APPARENT, stream);
      dim3 gp((unsigned)((E + 255) / 256), batch);
      parents_kernel<<<gp, 256, 0, stream>>>(L.rank, L.ea, L.T, n, E, L.par);
      count_launch();
      TDA_LAUNCH_CHECK();
    }
    {
      StageScope st(STAGE_RIPS_REDUCE, stream);
      if (L.reducer == 0) {
        // one thread-block cluster per cloud at a time: cluster size from the option, shrunk for big batches (more clouds than
        // clusters fit on the machine: rather one cloud per SM) 
        int C = (int)option("rips_cluster");
        if (C != 1 && C != 2 && C != 4 && C != 8) C = batch <= 4 ? 8 : 4;   // auto: few clouds -> more SMs per cloud
        while (C > 1 && (long long)batch * C > 2ll * sms) C >>= 1;
        const size_t dyn = Sweeper2::dyn_bytes(L.xw, L.s2_wmax);
        if (dyn > (size_t)200 * 1024) return set_error(TDA_ERR_UNSUPPORTED, "tda_rips: sweep2 needs %zu bytes of shared memory (n=%d, rips_wmax=%d)", dyn, n, L.s2_wmax);
        TDA_CUDA_CHECK(cudaFuncSetAttribute(rips_sweep2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        // (tried: CTAs of 256 threads, two per SM, for batches with more clouds than SMs -- 256 bootstrap resamples: 23.4 ms against
        // 19.4 ms for one 512-thread CTA per SM, profiles/r02n_c4_one_batch.log; not kept)
        int nclusters = sms / C;
        if (nclusters > batch) nclusters = batch;
        if (nclusters > L.grid) nclusters = L.grid;
        if (nclusters < 1) nclusters = 1;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(nclusters * C), 1, 1);
        cfg.blockDim = dim3(kS2Threads, 1, 1);
        cfg.dynamicSmemBytes = dyn;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        TDA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rips_sweep2_kernel, P));
      } else if (L.sweep) {
        // TDA_SWEEP_EXCLUSIVE=1: ask for all of the SM's shared memory, so that no CTA of a kernel running on another stream
        // (UMAP SGD of the next chunk ...) shares the SM -- and the issue slots -- with the latency-bound resolver warp
        const bool exclusive = option("sweep_exclusive") != 0;
        size_t dyn = sizeof(uint32_t) * ((size_t)L.xw * (1 + 4 * kGroupRows) + (size_t)n + 2 * kChunkRows);
#define TDA_SWEEP_LAUNCH_K(KERNEL)                                                                                            \
  do {                                                                                                                        \
    if (exclusive) {                                                                                                          \
      cudaFuncAttributes fa;                                                                                                  \
      TDA_CUDA_CHECK(cudaFuncGetAttributes(&fa, KERNEL));                                                                     \
      const size_t room = (size_t)227 * 1024 - fa.sharedSizeBytes;                                                            \
      if (dyn < room) dyn = room;                                                                                             \
    }                                                                                                                         \
    TDA_CUDA_CHECK(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));                      \
    KERNEL<<<L.grid, kSweepThreads, dyn, stream>>>(P);                                                                        \
  } while (0)
#define TDA_SWEEP_LAUNCH(WPL)                                                                                                 \
  do {                                                                                                                        \
    if (P.verify_mode) TDA_SWEEP_LAUNCH_K((rips_sweep_kernel<WPL, true>));                                                    \
    else TDA_SWEEP_LAUNCH_K((rips_sweep_kernel<WPL, false>));                                                                 \
  } while (0)
        if (L.xw <= 32) TDA_SWEEP_LAUNCH(1);
        else if (L.xw <= 64) TDA_SWEEP_LAUNCH(2);
        else if (L.xw <= 128) TDA_SWEEP_LAUNCH(4);
        else if (L.xw <= 256) TDA_SWEEP_LAUNCH(8);
        else return set_error(TDA_ERR_UNSUPPORTED, "tda_rips: internal: sweep reducer selected for n=%d", n);
#undef TDA_SWEEP_LAUNCH
#undef TDA_SWEEP_LAUNCH_K
      } else {
        const size_t s1_bytes = (size_t)(((L.wbits >> kPageShift) + 31) / 32) * sizeof(uint32_t);
        TDA_CUDA_CHECK(cudaFuncSetAttribute(rips_reduce_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1_bytes));
        rips_reduce_kernel<1><<<L.grid, kReduceThreads, s1_bytes, stream>>>(P);
      }
    }
    count_launch();
    TDA_LAUNCH_CHECK();
    finalize_stats_kernel<<<(batch + 255) / 256, 256, 0, stream>>>(L.T, L.bcount, n, batch, L.stats, counts);
    count_launch();
  }
  return TDA_OK;
}

extern "C" int tda_rips_launch(const float* dm, int n, int batch, int maxdim, float thresh, float* h0_pairs, int64_t* h0_simplex,
                               float* h1_pairs, int64_t* h1_simplex, int cap1, int32_t* counts, float* thresh_out, void* ws,
                               size_t ws_bytes, size_t pool_bytes, void* stream_) {
  return rips_enqueue(dm, n, batch, maxdim, thresh, h0_pairs, h0_simplex, h1_pairs, h1_simplex, cap1, counts, thresh_out, ws, ws_bytes,
                      pool_bytes, stream_);
}

extern "C" int tda_rips(const float* dm, int n, int batch, int maxdim, float thresh, float* h0_pairs, int64_t* h0_simplex,
                        float* h1_pairs, int64_t* h1_simplex, int cap1, int32_t* counts, float* thresh_out, void* ws,
                        size_t ws_bytes, size_t pool_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int rc = rips_enqueue(dm, n, batch, maxdim, thresh, h0_pairs, h0_simplex, h1_pairs, h1_simplex, cap1, counts, thresh_out, ws,
                              ws_bytes, pool_bytes, stream_);
  if (rc != TDA_OK) return rc;
  // overflow status must be known to the caller
  TDA_CUDA_CHECK(cudaStreamSynchronize(stream));
  {
    static thread_local int32_t* hc = nullptr;
    static thread_local int hc_cap = 0;
    if (hc_cap < batch) { delete[] hc; hc = new int32_t[(size_t)batch * 4]; hc_cap = batch; }
    TDA_CUDA_CHECK(cudaMemcpy(hc, counts, sizeof(int32_t) * 4 * batch, cudaMemcpyDeviceToHost));
    for (int p = 0; p < batch; ++p)
      if (hc[p * 4 + 3] != 0)
        return set_error(TDA_ERR_CAPACITY, "tda_rips: problem %d overflowed (cap1=%d or column pool %zu bytes); retry with larger sizes", p, cap1, pool_bytes);
  }
  return TDA_OK;
}

// ---- the edges of `batch` clouds in filtration order (ALL of them: no threshold), for tda_rips_subsets_launch
namespace {
struct SortLayout { uint64_t *keys_a, *keys_b; uint32_t *vals_a, *vals_b; void* cub_tmp; size_t cub_bytes; uint32_t* thresh_bits; int* T; size_t total; };
SortLayout make_sort_layout(void* ws, int n, int batch) {
  SortLayout L;
  const int64_t BE = (int64_t)batch * ((int64_t)n * (n - 1) / 2);
  Carver c(ws, ~size_t(0));
  L.keys_a = c.take<uint64_t>(BE); L.keys_b = c.take<uint64_t>(BE);
  L.vals_a = c.take<uint32_t>(BE); L.vals_b = c.take<uint32_t>(BE);
  L.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, L.cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, BE > 0 ? BE : 1, 0, 64, (cudaStream_t)0);
  L.cub_tmp = c.take<char>(L.cub_bytes + 256);
  L.thresh_bits = c.take<uint32_t>(batch);
  L.T = c.take<int>(batch);
  L.total = c.off;
  return L;
}
}  // namespace

extern "C" size_t tda_rips_sort_edges_workspace_bytes(int n, int batch) {
  if (n <= 0 || batch <= 0) return 0;
  return make_sort_layout(nullptr, n, batch).total + 4096;
}

extern "C" int tda_rips_sort_edges(const float* dm, int n, int batch, uint32_t* ends_out, float* sdist_out, void* ws, size_t ws_bytes,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!dm || !ends_out || !sdist_out || !ws || n <= 1 || batch <= 0) return set_error(TDA_ERR_INVALID, "tda_rips_sort_edges: bad arguments");
  if (n > 65535 || batch > 65535) return set_error(TDA_ERR_UNSUPPORTED, "tda_rips_sort_edges: n=%d batch=%d (at most 65535 each)", n, batch);
  SortLayout L = make_sort_layout(ws, n, batch);
  if (L.total > ws_bytes) return set_error(TDA_ERR_WORKSPACE, "tda_rips_sort_edges: workspace %zu < required %zu", ws_bytes, L.total);
  const int64_t E = (int64_t)n * (n - 1) / 2, BE = (int64_t)batch * E;
  StageScope st(STAGE_RIPS_SORT, stream);
  TDA_CUDA_CHECK(cudaMemsetAsync(L.T, 0, sizeof(int) * batch, stream));
  enclosing_init_kernel<<<(batch + 255) / 256, 256, 0, stream>>>(L.thresh_bits, batch, INFINITY);   // every finite length is in
  dim3 g((unsigned)((E + 255) / 256), batch);
  edge_keys_kernel<<<g, 256, 0, stream>>>(dm, n, E, L.thresh_bits, L.keys_a, L.vals_a, L.T);
  int pbits = 0;
  while ((1 << pbits) < batch) ++pbits;
  size_t tmp = L.cub_bytes;
  TDA_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(L.cub_tmp, tmp, (const uint64_t*)L.keys_a, L.keys_b, (const uint32_t*)L.vals_a, L.vals_b, BE, 0,
                                                 32 + pbits, stream));
  sorted_edges_kernel<<<g, 256, 0, stream>>>(L.keys_b, L.vals_b, E, ends_out, sdist_out);
  count_launch(3 + 4 + (pbits + 7) / 8 + 2);
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}

extern "C" int tda_rips_subsets_launch(const uint32_t* parent_ends, const float* parent_sdist, const float* parent_dm, int n_parent,
                                       const int32_t* subset_idx, int m, int batch, int maxdim, float thresh, float* h0_pairs,
                                       int64_t* h0_simplex, float* h1_pairs, int64_t* h1_simplex, int cap1, int32_t* counts, float* thresh_out,
                                       void* ws, size_t ws_bytes, size_t pool_bytes, void* stream_) {
  if (!parent_ends || !parent_sdist || !subset_idx || n_parent < 2 || n_parent > 65535 || m < 1 || m > n_parent)
    return set_error(TDA_ERR_INVALID, "tda_rips_subsets_launch: bad arguments");
  if (isinf(thresh) && !parent_dm) return set_error(TDA_ERR_INVALID, "tda_rips_subsets_launch: the enclosing radius needs the parent's distance matrix");
  SubsetInput sub{parent_ends, parent_sdist, parent_dm, n_parent, subset_idx};
  return rips_enqueue(nullptr, m, batch, maxdim, thresh, h0_pairs, h0_simplex, h1_pairs, h1_simplex, cap1, counts, thresh_out, ws, ws_bytes,
                      pool_bytes, stream_, &sub);
}

extern "C" int tda_rips_stats(const void* ws, int n, int batch, int maxdim, int cap1, size_t pool_bytes, int64_t* stats_host) {
  if (!ws || !stats_host) return set_error(TDA_ERR_INVALID, "tda_rips_stats: bad arguments");
  Layout L = make_layout((void*)ws, n, batch, maxdim, next_pow2(cap1), pool_bytes, sm_count_cached());
  TDA_CUDA_CHECK(cudaMemcpy(stats_host, L.stats, sizeof(int64_t) * ST_N * batch, cudaMemcpyDeviceToHost));
  return TDA_OK;
}

// ------------------------------------------------------------------------------------------------
// H2 on top of a finished H1 run
namespace tda {
namespace rips {
struct Layout2 {
  uint32_t* cbits; int64_t cwords; short* apex4; int* blist2; int* bcount2;
  uint64_t* hkeys; int* hvals; int hcap; int64_t* vstart; int* vlen; uint32_t* vpool; int64_t vpool_cap;
  uint32_t* bits; uint64_t wbits; uint32_t* vbits; int64_t vwords; uint32_t* vlist; int64_t vcap;
  int* work_counter; unsigned long long* stats;
  uint64_t* far; uint64_t far_cap; uint32_t nbk;
  int grid; size_t total;
};
static Layout2 make_layout2(void* ws, int n, int batch, int cap2, size_t pool_bytes, size_t far_bytes, int sm_count) {
  Layout2 L;
  memset(&L, 0, sizeof(L));
  const int64_t E = (int64_t)n * (n - 1) / 2;
  const int64_t K3 = E * n;   // triangle key space
  Carver c(ws, ~size_t(0));
  L.cwords = (K3 + 31) / 32 + 1;
  L.cbits = c.take<uint32_t>((size_t)batch * L.cwords);
  L.apex4 = c.take<short>((size_t)batch * K3);
  L.blist2 = c.take<int>((size_t)batch * cap2);
  L.bcount2 = c.take<int>(batch);
  L.hcap = next_pow2(2 * cap2);
  L.hkeys = c.take<uint64_t>((size_t)batch * L.hcap);
  L.hvals = c.take<int>((size_t)batch * L.hcap);
  L.vstart = c.take<int64_t>((size_t)batch * cap2);
  L.vlen = c.take<int>((size_t)batch * cap2);
  L.stats = c.take<unsigned long long>((size_t)batch * ST_N);
  L.work_counter = c.take<int>(1);
  L.grid = batch < 2 * sm_count ? batch : 2 * sm_count;
  if (L.grid < 1) L.grid = 1;
  L.vwords = L.cwords;
  L.vbits = c.take<uint32_t>((size_t)L.grid * L.vwords);
  L.vcap = (K3 < (int64_t)(16 << 20) ? K3 : (int64_t)(16 << 20)) + 1024;
  L.vlist = c.take<uint32_t>((size_t)L.grid * 2 * L.vcap);
  const uint64_t page = 1ull << kPageShift;
  const double want_d = (double)E * (double)n * (double)n;
  uint64_t want = want_d >= 4294967296.0 ? (1ull << 32) : (((uint64_t)want_d + 32 * page - 1) / (32 * page) * (32 * page));
  if (want > (1ull << 32)) want = 1ull << 32;
  uint64_t afford = (uint64_t)(pool_bytes / 2 / (size_t)L.grid) * 8 / (32 * page) * (32 * page);
  if (afford < 32 * page) afford = 32 * page;
  L.wbits = want < afford ? want : afford;
  L.bits = c.take<uint32_t>((size_t)L.grid * (L.wbits >> 5));
  const size_t bits_bytes = (size_t)L.grid * (L.wbits >> 3);
  L.vpool_cap = pool_bytes > bits_bytes ? (int64_t)((pool_bytes - bits_bytes) / (size_t)batch / sizeof(uint32_t)) : 0;
  if (L.vpool_cap < 4 * (int64_t)cap2) L.vpool_cap = 4 * (int64_t)cap2;
  L.vpool = c.take<uint32_t>((size_t)batch * L.vpool_cap);
  // far buckets: one region per window of the key space and resident CTA (only when the key space exceeds the window)
  L.nbk = want_d > (double)L.wbits ? (uint32_t)(want_d / (double)L.wbits) + 2 : 0;
  L.far_cap = L.nbk ? (uint64_t)(far_bytes / (size_t)L.grid / sizeof(uint64_t)) : 0;
  if (L.nbk > kMaxFarBuckets || L.far_cap < 256ull * L.nbk) { L.nbk = 0; L.far_cap = 0; }   // too many windows / no room: re-enumerate
  L.far = L.far_cap ? c.take<uint64_t>((size_t)L.grid * L.far_cap) : nullptr;
  L.total = c.off;
  return L;
}
}  // namespace rips
}  // namespace tda

extern "C" size_t tda_rips_h2_workspace_bytes(int n, int batch, int cap2, size_t pool_bytes, size_t far_bytes) {
  if (n <= 0 || batch <= 0 || cap2 <= 0) return 0;
  return make_layout2(nullptr, n, batch, next_pow2(cap2), pool_bytes, far_bytes, 148).total + 4096;
}

extern "C" int tda_rips_h2(const void* ws1, int n, int batch, int cap1, size_t pool_bytes1, float* h2_pairs, int cap2, int32_t* counts2,
                           void* ws2, size_t ws2_bytes, size_t pool_bytes2, size_t far_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!ws1 || !h2_pairs || !counts2 || !ws2 || n <= 0 || batch <= 0 || cap2 <= 0) return set_error(TDA_ERR_INVALID, "tda_rips_h2: bad arguments");
  if (n > 32 * kH2MaxWords) return set_error(TDA_ERR_UNSUPPORTED, "tda_rips_h2: n=%d > %d (H2 is implemented for small clouds)", n, 32 * kH2MaxWords);
  if (next_pow2(cap2) != cap2) return set_error(TDA_ERR_INVALID, "tda_rips_h2: cap2 must be a power of two");
  if (batch > 65535) return set_error(TDA_ERR_INVALID, "tda_rips_h2: batch > 65535");
  const int sms = sm_count_cached();
  Layout L1 = make_layout((void*)ws1, n, batch, 1, next_pow2(cap1), pool_bytes1, sms);
  Layout2 L = make_layout2(ws2, n, batch, cap2, pool_bytes2, far_bytes, sms);
  if (L.total > ws2_bytes) return set_error(TDA_ERR_WORKSPACE, "tda_rips_h2: workspace %zu < required %zu", ws2_bytes, L.total);
  const int64_t E = (int64_t)n * (n - 1) / 2;
  if (E == 0) { TDA_CUDA_CHECK(cudaMemsetAsync(counts2, 0, sizeof(int32_t) * 4 * batch, stream)); return TDA_OK; }
  const int64_t K3 = E * n;
  TDA_CUDA_CHECK(cudaMemsetAsync(counts2, 0, sizeof(int32_t) * 4 * batch, stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.cbits, 0, sizeof(uint32_t) * (size_t)batch * L.cwords, stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.apex4, 0xff, sizeof(short) * (size_t)batch * K3, stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.bcount2, 0, sizeof(int) * batch, stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.work_counter, 0, sizeof(int), stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.stats, 0, sizeof(unsigned long long) * batch * ST_N, stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.vbits, 0, sizeof(uint32_t) * (size_t)L.grid * L.vwords, stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.bits, 0, (size_t)L.grid * (L.wbits >> 3), stream));
  {
    const int64_t work = E > L1.hcap ? E : L1.hcap;
    dim3 g((unsigned)((work + 255) / 256), batch);
    const bool debug_sync = option("debug_sync") != 0;   // locate a faulting kernel: synchronise after each one
    if (debug_sync) TDA_CUDA_CHECK(cudaStreamSynchronize(stream));
    h2_clear_kernel<<<g, 256, 0, stream>>>(L1.ea, L1.T, L1.hkeys, L1.hcap, n, E, L.cbits, L.cwords);
    if (debug_sync) TDA_CUDA_CHECK(cudaStreamSynchronize(stream));
    dim3 g2((unsigned)((E * 32 + 255) / 256), batch);
    h2_apparent_kernel<<<g2, 256, 0, stream>>>(L1.rank, L1.ends, L1.T, n, E, L.cbits, L.cwords, L.apex4, L.blist2, L.bcount2, cap2);
    if (debug_sync) TDA_CUDA_CHECK(cudaStreamSynchronize(stream));
    count_launch(2);
    TDA_LAUNCH_CHECK();
  }
  ReduceParams P;
  memset(&P, 0, sizeof(P));
  P.rank = L1.rank; P.ends = L1.ends; P.sdist = L1.sdist; P.T = L1.T; P.ea = L1.ea;
  P.blist = L.blist2; P.bcount = L.bcount2;
  P.n = n; P.E = E; P.batch = batch; P.cap1 = cap2;
  P.h1_pairs = h2_pairs; P.h1_simplex = nullptr; P.counts = counts2;
  P.bits = L.bits; P.wbits = L.wbits;
  P.vbits = L.vbits; P.vwords = L.vwords; P.vlist = L.vlist; P.vcap = L.vcap;
  P.hkeys = L.hkeys; P.hvals = L.hvals; P.hcap = L.hcap;
  P.vpool = L.vpool; P.vpool_cap = L.vpool_cap; P.vstart = L.vstart; P.vlen = L.vlen;
  P.work_counter = L.work_counter; P.stats = L.stats;
  P.apex4 = L.apex4;
  P.far = L.far; P.far_cap = L.far_cap; P.nbk = L.nbk;
  {
    // page summary, then (far buckets) nbk fill counters and nbk + 1 region offsets
    const size_t s1_bytes = ((size_t)(((L.wbits >> kPageShift) + 31) / 32) + (size_t)L.nbk + 2) * sizeof(uint32_t) + ((size_t)L.nbk + 1) * sizeof(uint64_t);
    TDA_CUDA_CHECK(cudaFuncSetAttribute(rips_reduce_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1_bytes));
    rips_reduce_kernel<2><<<L.grid, kReduceThreads, s1_bytes, stream>>>(P);
    count_launch();
    TDA_LAUNCH_CHECK();
  }
  TDA_CUDA_CHECK(cudaStreamSynchronize(stream));
  {
    std::vector<int32_t> hc((size_t)batch * 4);
    TDA_CUDA_CHECK(cudaMemcpy(hc.data(), counts2, sizeof(int32_t) * 4 * batch, cudaMemcpyDeviceToHost));
    if (option("h2_stats")) {   // device counters of the H2 reduction, one line per problem (diagnostics)
      std::vector<unsigned long long> hs((size_t)batch * ST_N);
      std::vector<int> hb(batch);
      TDA_CUDA_CHECK(cudaMemcpy(hs.data(), L.stats, sizeof(unsigned long long) * ST_N * batch, cudaMemcpyDeviceToHost));
      TDA_CUDA_CHECK(cudaMemcpy(hb.data(), L.bcount2, sizeof(int) * batch, cudaMemcpyDeviceToHost));
      for (int p = 0; p < batch; ++p) {
        const unsigned long long* q = hs.data() + (size_t)p * ST_N;
        fprintf(stderr, "tda_rips_h2 stats p=%d n=%d: residual columns %d (cap2 %d), additions %llu, toggles %llu, pivots %llu, slides+refills %llu, max|V| %llu, "
                        "Mcycles scan %.1f owner %.1f apparent-add %.1f column-add %.1f slide %.1f finalise %.1f, edges via columns %llu, far keys+passes %llu, "
                        "windows %u far_cap %llu wbits %llu\n",
                p, n, hb[p], cap2, q[ST_ADDITIONS], q[ST_PUSHES], q[ST_POPS], q[ST_EXTENSIONS], q[ST_MAXV], q[ST_CYC_EXTRACT] / 1e6, q[ST_CYC_OWNER] / 1e6,
                q[ST_CYC_GEN] / 1e6, q[ST_CYC_BADD] / 1e6, q[ST_CYC_EXT] / 1e6, q[ST_CYC_FINAL] / 1e6, q[ST_BADD_EDGES], q[ST_EXT_EDGES], L.nbk,
                (unsigned long long)L.far_cap, (unsigned long long)L.wbits);
      }
    }
    for (int p = 0; p < batch; ++p)
      if (hc[p * 4 + 3] != 0)
        return set_error(TDA_ERR_CAPACITY, "tda_rips_h2: problem %d overflowed (cap2=%d or pool %zu bytes); retry with larger sizes", p, cap2, pool_bytes2);
  }
  return TDA_OK;
}
