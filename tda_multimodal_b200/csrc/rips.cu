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
//                               not in the MST) are reduced, by implicit persistent cohomology over Z/2:
//                               the working column is a monotone radix heap of triangle keys in a chunked
//                               global-memory pool, pivots owned by apparent pairs are recognised in O(1)
//                               from the apex table, pivots owned by reduced columns through a hash map.
//
// A triangle is keyed by (rank of its longest edge, opposite vertex): key = M*n + (n-1-w).  Ascending key
// order refines (diameter asc) and puts faces before cofaces, so it is a valid simplex-wise filtration; the
// persistence diagram (as a multiset of (birth,death) values) does not depend on how ties are broken.
#include "common.cuh"
#include "launch_count.cuh"
#include "../../include/tda_b200.h"
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>
#include <cmath>

namespace tda {
namespace rips {

constexpr int kReduceThreads = 256;
constexpr int kChunk = 512;            // keys per heap chunk
constexpr uint32_t kNil = 0xffffffffu;
constexpr int kGenItems = 8;           // cofacets generated per thread per sub-batch (256*8 = 2048 vertices)
constexpr int kMaxNew = kReduceThreads * kGenItems / kChunk + 2;
constexpr int kRankDiag = 0x7fffffff;

// stats slots
enum { ST_COLUMNS = 0, ST_APPARENT, ST_REDUCED, ST_ADDITIONS, ST_PUSHES, ST_POPS, ST_EXTENSIONS, ST_MAXV,
       ST_CYC_EXTRACT, ST_CYC_OWNER, ST_CYC_GEN, ST_CYC_BADD, ST_CYC_EXT, ST_CYC_FINAL, ST_BADD_EDGES, ST_EXT_EDGES, ST_N };

// ------------------------------------------------------------------------------------------------
// low-dimensional euclidean distance matrix (ripser.py front end)
__global__ void pdist_lowdim_kernel(const float* __restrict__ pts, int n, int d, float* __restrict__ dm) {
  extern __shared__ float sp[];  // [n_tile_i + n_tile_j][d] not needed: d is tiny, read directly
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
// H0: Boruvka MST on the rank matrix.  Ranks are distinct, so the minimum spanning forest is unique and
// equals the set of merging edges of ripser's union-find sweep.  Per round: a grid-wide scan kernel (one
// warp per matrix row: smallest rank leaving the row's component -> atomicMin per component) and a
// one-CTA-per-cloud merge kernel (hook, break 2-cycles, pointer jumping, relabel).
constexpr uint32_t kNoEdge = 0xffffffffu;
__global__ void boruvka_init_kernel(int n, uint32_t* __restrict__ comp, uint32_t* __restrict__ cbest, int* __restrict__ done) {
  int p = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { comp[(size_t)p * n + i] = i; cbest[(size_t)p * n + i] = kNoEdge; }
  if (i == 0) done[p] = 0;
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
                                                             uint8_t* __restrict__ mst, int* __restrict__ done) {
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

// gather the MST ranks, sort them, emit the H0 rows.  One CTA per cloud.
__global__ void __launch_bounds__(1024) h0_emit_kernel(const uint32_t* __restrict__ ends, const float* __restrict__ sdist,
                                                       const int* __restrict__ Tarr, int n, int64_t E, const uint8_t* __restrict__ mst,
                                                       const uint32_t* __restrict__ comp_g, float* __restrict__ h0_pairs,
                                                       int64_t* __restrict__ h0_simplex, int32_t* __restrict__ counts, int* __restrict__ mstlist_g) {
  __shared__ int s_zero, s_nmst;
  const int p = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const uint32_t* EN = ends + (size_t)p * E;
  const uint32_t* comp = comp_g + (size_t)p * n;
  const int T = Tarr[p];
  const uint8_t* M = mst + (size_t)p * E;
  int* mstlist = mstlist_g + (size_t)p * n;
  if (tid == 0) { s_nmst = 0; s_zero = 0; }
  __syncthreads();
  for (int64_t r = tid; r < T; r += nt)
    if (M[r]) {
      int pos = atomicAdd(&s_nmst, 1);
      if (pos < n) mstlist[pos] = (int)r;
    }
  __syncthreads();
  const int nm = min(s_nmst, n - 1);
  int np2 = 1;
  while (np2 < nm) np2 <<= 1;
  if (np2 <= n) {
    for (int i = nm + tid; i < np2; i += nt) mstlist[i] = 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < np2; i += nt) {
          int ixj = i ^ j;
          if (ixj > i) {
            int a = mstlist[i], b = mstlist[ixj];
            bool up = ((i & k) == 0);
            if ((a > b) == up) { mstlist[i] = b; mstlist[ixj] = a; }
          }
        }
        __syncthreads();
      }
  } else {  // np2 > n only for tiny n: serial insertion sort
    if (tid == 0)
      for (int i = 1; i < nm; ++i) {
        int v = mstlist[i], k = i - 1;
        while (k >= 0 && mstlist[k] > v) { mstlist[k + 1] = mstlist[k]; --k; }
        mstlist[k + 1] = v;
      }
    __syncthreads();
  }
  // rows: (0, d) for d != 0 ascending, then one (0, inf) per remaining component.  Zero-length merging
  // edges have the smallest ranks, so they are a prefix of the sorted list.
  float* out = h0_pairs + (size_t)p * n * 2;
  int64_t* outs = h0_simplex ? h0_simplex + (size_t)p * n * 2 : nullptr;
  for (int k = tid; k < nm; k += nt)
    if (sdist[(size_t)p * E + mstlist[k]] == 0.f) atomicAdd(&s_zero, 1);
  __syncthreads();
  const int z = s_zero;
  for (int k = z + tid; k < nm; k += nt) {
    int r = mstlist[k];
    int row = k - z;
    out[2 * row] = 0.f; out[2 * row + 1] = sdist[(size_t)p * E + r];
    if (outs) {
      uint32_t e = EN[r];
      outs[2 * row] = -1; outs[2 * row + 1] = edge_index((int)(e >> 16), (int)(e & 0xffffu));
    }
  }
  if (tid == 0) {
    int rows = nm - z;
    for (int i = 0; i < n && rows < n; ++i)
      if (comp[i] == (uint32_t)i) {  // one essential class per component, reported at its root vertex
        out[2 * rows] = 0.f; out[2 * rows + 1] = INFINITY;
        if (outs) { outs[2 * rows] = i; outs[2 * rows + 1] = -1; }
        ++rows;
      }
    counts[p * 4 + 0] = rows;
    counts[p * 4 + 2] = T;
  }
}

// ------------------------------------------------------------------------------------------------
// apparent pairs: one warp per edge rank r < T.  apex[r] = largest vertex v with rank(a,v) < r and
// rank(b,v) < r (the first cofacet of the edge in filtration order has the edge as its longest edge),
// -1 if the lune is empty, -2 for MST edges (negative edges are not columns).
__global__ void apparent_kernel(const int* __restrict__ rank, const uint32_t* __restrict__ ends, const uint8_t* __restrict__ mst,
                                const int* __restrict__ Tarr, int n, int64_t E, int* __restrict__ apex, int* __restrict__ blist,
                                int* __restrict__ bcount, int cap1, unsigned long long* __restrict__ stats) {
  const int p = blockIdx.y;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int T = Tarr[p];
  if (r >= T) return;
  int* A = apex + (size_t)p * E;
  if (mst[(size_t)p * E + r]) {
    if (lane == 0) A[r] = -2;
    return;
  }
  const uint32_t e = ends[(size_t)p * E + r];
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
    if (found < 0) {
      int pos = atomicAdd(&bcount[p], 1);
      if (pos < cap1) blist[(size_t)p * cap1 + pos] = (int)r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// residual reduction
//
// Working column of one edge b with an empty lune = set of triangle keys with odd multiplicity in
// delta(V), V = edges added so far.  Three levels, all owned by one CTA:
//   F      : open-addressing hash SET in shared memory holding the keys in [base, fmax]; inserting a key
//            that is already present removes it (Z/2 cancellation is free); pop-min = block-wide scan.
//   heap   : monotone radix heap for the keys > fmax: bucket q>=1 holds keys whose highest bit differing
//            from `base` is q-1, as linked lists of 512-key chunks in a global pool (per-CTA free list +
//            global bump).  When F runs dry the first non-empty bucket [L,U] is streamed into F
//            (base=L, fmax=U); when F overflows its upper half is spilled back (fmax lowered).
// Keys above the horizon H are not stored at all; when everything <= H is consumed the horizon is
// extended and delta(V) re-enumerated for the new window.
template <typename K> struct KeyTraits;
template <> struct KeyTraits<uint32_t> {
  static constexpr int NB = 33;
  static __device__ __forceinline__ int bucket(uint32_t x) { return 32 - __clz(x); }
  static __device__ __forceinline__ uint32_t maxv() { return 0xffffffffu; }
  static __device__ __forceinline__ uint32_t low_mask(int b) { return b >= 32 ? 0xffffffffu : ((1u << b) - 1u); }
  static __device__ __forceinline__ uint32_t cas(uint32_t* a, uint32_t c, uint32_t v) { return atomicCAS(a, c, v); }
};
template <> struct KeyTraits<uint64_t> {
  static constexpr int NB = 65;
  static __device__ __forceinline__ int bucket(uint64_t x) { return 64 - __clzll((long long)x); }
  static __device__ __forceinline__ uint64_t maxv() { return ~0ull; }
  static __device__ __forceinline__ uint64_t low_mask(int b) { return b >= 64 ? ~0ull : ((1ull << b) - 1ull); }
  static __device__ __forceinline__ uint64_t cas(uint64_t* a, uint64_t c, uint64_t v) {
    return (uint64_t)atomicCAS((unsigned long long*)a, (unsigned long long)c, (unsigned long long)v);
  }
};

constexpr int kFCap = 4096;        // slots of the shared-memory front set
constexpr int kFLoad = 1024;       // largest bucket loaded into F in one go / live keys kept after a spill
constexpr int kFUsedMax = 1536;    // used slots (live + tombstones) allowed before a 2048-key batch
constexpr int kFreeCache = 48;

struct ReduceParams {
  const int* rank; const uint32_t* ends; const float* sdist; const int* T; const int* apex;
  int* blist; const int* bcount;
  int n; int64_t E; int batch; int cap1;
  float* h1_pairs; int64_t* h1_simplex; int32_t* counts;
  // per-CTA scratch
  uint32_t* vbits; int64_t vwords;          // [grid, vwords]
  uint32_t* vlist; int64_t vcap;            // [grid, 2, vcap]
  // per-problem
  void* hkeys; int* hvals; int hcap;        // [batch, hcap]
  uint32_t* vpool; int64_t vpool_cap;       // [batch, vpool_cap]
  int64_t* vstart; int* vlen;               // [batch, cap1]
  // heap pool (shared)
  void* pool_keys; uint32_t* pool_next; uint32_t pool_chunks; unsigned int* pool_top;
  int* work_counter; unsigned long long* stats;  // [batch, ST_N]
};

template <typename K>
struct ReduceSmem {
  K F[kFCap];
  K base, fmax, red[kReduceThreads / 32], red2[kReduceThreads / 32];
  uint32_t head[KeyTraits<K>::NB], tail[KeyTraits<K>::NB], fill[KeyTraits<K>::NB], count[KeyTraits<K>::NB];
  uint32_t bcnt[KeyTraits<K>::NB], oldtail[KeyTraits<K>::NB], oldfill[KeyTraits<K>::NB];
  uint32_t newchunk[KeyTraits<K>::NB][kMaxNew];
  uint32_t fcache[kFreeCache];
  uint32_t nzmask[3];
  uint32_t fcache_n, freehead, vcount, vcount2, vsel, f_used;
  int f_live;
  int bc_int, abort_flag, problem;
  unsigned long long pushes, pops;
};

template <typename K>
struct Reducer {
  using TR = KeyTraits<K>;
  static constexpr int NB = TR::NB;
  static constexpr int kFItems = kFCap / kReduceThreads;
  const ReduceParams& P;
  ReduceSmem<K>& S;
  const int tid;
  const int* R; const uint32_t* EN; const int* A; int T; int n;
  K* pool; uint32_t* vbits; uint32_t* vl[2];
  K* hkeys; int* hvals;

  static __device__ __forceinline__ K empty_key() { return TR::maxv(); }
  static __device__ __forceinline__ K tomb_key() { return TR::maxv() - 1; }

  __device__ Reducer(const ReduceParams& p, ReduceSmem<K>& s) : P(p), S(s), tid(threadIdx.x) {
    pool = (K*)P.pool_keys;
    n = P.n;
    vbits = P.vbits + (size_t)blockIdx.x * P.vwords;
    vl[0] = P.vlist + (size_t)blockIdx.x * 2 * P.vcap;
    vl[1] = vl[0] + P.vcap;
  }

  // ---- chunk pool (thread 0 only)
  __device__ uint32_t alloc_chunk() {
    if (S.fcache_n) return S.fcache[--S.fcache_n];
    uint32_t c = S.freehead;
    if (c != kNil) { S.freehead = P.pool_next[c]; return c; }
    c = atomicAdd(P.pool_top, 1u);
    if (c >= P.pool_chunks) { S.abort_flag = TDA_ERR_CAPACITY; return P.pool_chunks; /* trash chunk */ }
    return c;
  }
  __device__ void free_chunk(uint32_t c) {
    if (c >= P.pool_chunks) return;
    if (S.fcache_n < (uint32_t)kFreeCache) { S.fcache[S.fcache_n++] = c; return; }
    P.pool_next[c] = S.freehead;
    S.freehead = c;
  }

  __device__ void heap_reset() {
    if (tid < NB) { S.head[tid] = kNil; S.tail[tid] = kNil; S.fill[tid] = kChunk; S.count[tid] = 0; }
  }
  // return every chunk of every bucket to the free list (O(1) per bucket: splice)
  __device__ void heap_release() {
    __syncthreads();
    if (tid == 0)
      for (int b = 0; b < NB; ++b)
        if (S.head[b] != kNil && S.head[b] < P.pool_chunks) {
          uint32_t t = S.tail[b];
          if (t < P.pool_chunks) { P.pool_next[t] = S.freehead; S.freehead = S.head[b]; }
        }
    __syncthreads();
    heap_reset();
    __syncthreads();
  }

  // ---- F: shared-memory front set with toggle semantics
  __device__ void f_clear() {
#pragma unroll
    for (int it = 0; it < kFItems; ++it) S.F[tid + it * kReduceThreads] = empty_key();
    if (tid == 0) { S.f_used = 0; S.f_live = 0; }
    __syncthreads();
  }
  __device__ __forceinline__ void f_toggle(K key) {
    uint32_t h = (uint32_t)(((uint64_t)key * 0x9E3779B97F4A7C15ull) >> 40) & (kFCap - 1);
    for (;;) {
      K cur = S.F[h];
      if (cur == key) {
        if (TR::cas(&S.F[h], key, tomb_key()) == key) { atomicSub(&S.f_live, 1); return; }
        cur = S.F[h];  // someone else removed it first: keep probing
      }
      if (cur == empty_key()) {
        K prev = TR::cas(&S.F[h], empty_key(), key);
        if (prev == empty_key()) { atomicAdd(&S.f_used, 1u); atomicAdd(&S.f_live, 1); return; }
        if (prev == key) continue;  // the same key landed here concurrently: retry this slot (will remove it)
      }
      h = (h + 1) & (kFCap - 1);
    }
  }
  __device__ K block_min(K v) {
    v = sizeof(K) == 4 ? (K)warp_min_u32((uint32_t)v) : (K)warp_min_u64((uint64_t)v);
    if ((tid & 31) == 0) S.red[tid >> 5] = v;
    __syncthreads();
    K m = S.red[0];
#pragma unroll
    for (int w = 1; w < kReduceThreads / 32; ++w) m = S.red[w] < m ? S.red[w] : m;
    __syncthreads();
    return m;
  }
  // smallest live key of F, removed; false if F holds no live key
  __device__ bool f_popmin(K& out) {
    if (S.f_live <= 0) return false;
    K m = tomb_key();
    int at = -1;
#pragma unroll
    for (int it = 0; it < kFItems; ++it) {
      int i = tid + it * kReduceThreads;
      K k = S.F[i];
      if (k < m) { m = k; at = i; }
    }
    const K g = block_min(m);
    if (g >= tomb_key()) return false;
    if (m == g && at >= 0) { S.F[at] = tomb_key(); S.f_live -= 1; }  // exactly one slot holds a live key
    __syncthreads();
    out = g;
    return true;
  }

  // ---- push a batch of keys held in registers to the heap (all threads call).
  // force_bucket: -1 => by radix relative to S.base (keys > fmax >= base)
  template <int ITEMS>
  __device__ void push_batch(const K (&keys)[ITEMS], const bool (&valid)[ITEMS], int force_bucket) {
    int any = 0;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) any |= valid[it] ? 1 : 0;
    if (tid < NB) S.bcnt[tid] = 0;
    if (!__syncthreads_or(any)) return;
    int bk[ITEMS];
    uint32_t off[ITEMS];
    const K base = S.base;
    const unsigned lane_lt = (1u << (tid & 31)) - 1u;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      // warp-aggregated slot reservation: one shared-memory atomic per distinct bucket per warp
      bk[it] = valid[it] ? (force_bucket >= 0 ? force_bucket : TR::bucket(keys[it] ^ base)) : -1;
      unsigned act = __ballot_sync(0xffffffffu, valid[it]);
      if (act) {
        unsigned peers = __match_any_sync(0xffffffffu, bk[it]);
        if (valid[it]) {
          int leader = __ffs(peers) - 1;
          uint32_t b0 = 0;
          if ((tid & 31) == leader) b0 = atomicAdd(&S.bcnt[bk[it]], (uint32_t)__popc(peers));
          b0 = __shfl_sync(peers, b0, leader);
          off[it] = b0 + __popc(peers & lane_lt);
        }
      }
    }
    __syncthreads();
    if (tid < NB) {
      if (S.bcnt[tid]) {
        S.oldtail[tid] = S.tail[tid];
        S.oldfill[tid] = S.fill[tid];
        atomicOr(&S.nzmask[tid >> 5], 1u << (tid & 31));
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long added = 0;
      for (int wq = 0; wq < 3; ++wq) {
        unsigned msk = S.nzmask[wq];
        S.nzmask[wq] = 0;
        while (msk) {
          int b = wq * 32 + __ffs(msk) - 1;
          msk &= msk - 1;
          uint32_t c = S.bcnt[b];
          added += c;
          uint32_t total = S.fill[b] + c;
          uint32_t nnew = total > (uint32_t)kChunk ? (total - kChunk + kChunk - 1) / kChunk : 0;
          uint32_t prev = S.tail[b];
          for (uint32_t j = 0; j < nnew; ++j) {
            uint32_t ch = alloc_chunk();
            S.newchunk[b][j] = ch;
            if (prev != kNil && prev < P.pool_chunks) P.pool_next[prev] = ch; else if (prev == kNil) S.head[b] = ch;
            prev = ch;
          }
          if (nnew) { S.tail[b] = prev; S.fill[b] = total - kChunk * nnew; } else S.fill[b] = total;
          S.count[b] += c;
        }
      }
      S.pushes += added;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it)
      if (valid[it]) {
        int b = bk[it];
        uint32_t pos = S.oldfill[b] + off[it];
        uint32_t ch;
        if (pos < (uint32_t)kChunk) ch = S.oldtail[b];
        else { pos -= kChunk; ch = S.newchunk[b][pos / kChunk]; pos %= kChunk; }
        pool[(size_t)ch * kChunk + pos] = keys[it];
      }
    __syncthreads();
  }

  // ---- keep F within its load limits: compact tombstones, spill the upper half to the heap when too many live keys
  __device__ void f_maintain(uint32_t max_used) {
    if (S.f_used <= max_used) return;
    K keep[kFItems];
    bool live[kFItems];
    for (;;) {
#pragma unroll
      for (int it = 0; it < kFItems; ++it) {
        keep[it] = S.F[tid + it * kReduceThreads];
        live[it] = keep[it] < tomb_key();
      }
      const int nlive = S.f_live;
      __syncthreads();
      K mid = TR::maxv();
      if (nlive > kFLoad) {  // spill keys above the midpoint of the live range
        K mn = TR::maxv(), mx = 0;
#pragma unroll
        for (int it = 0; it < kFItems; ++it)
          if (live[it]) { mn = keep[it] < mn ? keep[it] : mn; mx = keep[it] > mx ? keep[it] : mx; }
        mn = block_min(mn);
        mx = ~block_min((K)~mx);
        mid = mn + (mx - mn) / 2;
      }
      f_clear();
      bool spill[kFItems];
#pragma unroll
      for (int it = 0; it < kFItems; ++it) {
        spill[it] = live[it] && keep[it] > mid;
        if (live[it] && !spill[it]) f_toggle(keep[it]);
      }
      if (nlive > kFLoad) {
        if (tid == 0) S.fmax = mid;
        push_batch<kFItems>(keep, spill, -1);
      }
      __syncthreads();
      if (S.f_live <= kFLoad) break;
    }
  }

  // ---- refill F from the radix heap; false when nothing is stored at all
  __device__ bool refill() {
    for (;;) {
      if (S.f_live > 0) return true;
      int b = 0;
      for (int q = 1; q < NB; ++q)
        if (S.count[q]) { b = q; break; }
      if (!b) return false;
      f_clear();
      const uint32_t hb = S.head[b], tb = S.tail[b], fb = S.fill[b], cb = S.count[b];
      if (tid == 0) {  // detach the list; F now covers the bucket's whole key range [L, U]
        S.head[b] = kNil; S.tail[b] = kNil; S.fill[b] = kChunk; S.count[b] = 0;
        S.pops += cb;
        const K old = S.base;
        S.fmax = old | TR::low_mask(b);
        S.base = (old & ~TR::low_mask(b)) | ((K)1 << (b - 1));
      }
      __syncthreads();
      for (uint32_t c = hb; c != kNil;) {
        const uint32_t cnt = (c == tb) ? fb : (uint32_t)kChunk;
        const uint32_t nxt = (c == tb) ? kNil : P.pool_next[c];
        const K fmax = S.fmax;  // may have been lowered by a spill while streaming a long list
        K keys[kChunk / kReduceThreads];
        bool valid[kChunk / kReduceThreads];
#pragma unroll
        for (int it = 0; it < kChunk / kReduceThreads; ++it) {
          uint32_t i = tid + it * kReduceThreads;
          valid[it] = false;
          if (i < cnt) {
            K k = pool[(size_t)c * kChunk + i];
            if (k <= fmax) f_toggle(k); else { keys[it] = k; valid[it] = true; }
          }
        }
        __syncthreads();
        if (tid == 0) free_chunk(c);
        push_batch<kChunk / kReduceThreads>(keys, valid, -1);
        f_maintain(kFCap - 2 * kChunk);
        c = nxt;
      }
      __syncthreads();
    }
  }

  // next pivot: smallest key with odd multiplicity; false when the stored part of the column is empty
  __device__ bool extract(K& out) {
    if (!refill()) return false;
    return f_popmin(out);
  }

  // ---- cofacets of edge `re` with key in (lo, hi] -> F / heap
  __device__ void gen_push(int re, K lo, K hi) {
    const uint32_t e = EN[re];
    const int a = (int)(e >> 16), b = (int)(e & 0xffffu);
    const int* rowa = R + (size_t)a * n;
    const int* rowb = R + (size_t)b * n;
    for (int v0 = 0; v0 < n; v0 += kReduceThreads * kGenItems) {
      f_maintain(kFUsedMax);
      K keys[kGenItems];
      bool valid[kGenItems];
      int ra[kGenItems], rb[kGenItems];
#pragma unroll
      for (int it = 0; it < kGenItems; ++it) {
        int v = v0 + it * kReduceThreads + tid;
        ra[it] = v < n ? rowa[v] : kRankDiag;
        rb[it] = v < n ? rowb[v] : kRankDiag;
      }
      const K fmax = S.fmax;
#pragma unroll
      for (int it = 0; it < kGenItems; ++it) {
        int v = v0 + it * kReduceThreads + tid;
        int M = max(re, max(ra[it], rb[it]));
        valid[it] = false;
        if (M < T) {
          int opp = (M == re) ? v : (M == ra[it] ? b : a);
          K key = (K)M * (K)n + (K)(n - 1 - opp);
          if (key > lo && key <= hi) {
            keys[it] = key;
            if (key <= fmax) f_toggle(key); else valid[it] = true;
          }
        }
      }
      push_batch<kGenItems>(keys, valid, -1);
    }
  }

  // ---- V (the reduction column as a set of edges, by rank)
  __device__ void v_toggle_single(int re) {  // thread 0
    uint32_t w = (uint32_t)re >> 5, m = 1u << (re & 31);
    uint32_t old = atomicXor(&vbits[w], m);
    if (!(old & m)) {
      uint32_t pos = S.vcount;
      if (pos < (uint32_t)P.vcap) { vl[S.vsel][pos] = (uint32_t)re; S.vcount = pos + 1; }
      else S.abort_flag = TDA_ERR_CAPACITY;
    }
  }
  // compact the list: keep each edge whose bit is set exactly once
  __device__ void v_compact() {
    __syncthreads();
    const uint32_t nin = S.vcount;
    const uint32_t* src = vl[S.vsel];
    uint32_t* dst = vl[S.vsel ^ 1];
    if (tid == 0) S.vcount2 = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < nin; i0 += kReduceThreads) {
      uint32_t i = i0 + tid;
      bool keep = false;
      uint32_t e = 0;
      if (i < nin) {
        e = src[i];
        uint32_t m = 1u << (e & 31);
        keep = (atomicAnd(&vbits[e >> 5], ~m) & m) != 0;
      }
      unsigned bal = __ballot_sync(0xffffffffu, keep);
      uint32_t bs = 0;
      if ((tid & 31) == 0 && bal) bs = atomicAdd(&S.vcount2, (uint32_t)__popc(bal));
      bs = __shfl_sync(0xffffffffu, bs, 0);
      if (keep) dst[bs + __popc(bal & ((1u << (tid & 31)) - 1))] = e;
    }
    __syncthreads();
    const uint32_t nout = S.vcount2;
    for (uint32_t i = tid; i < nout; i += kReduceThreads) {
      uint32_t e = dst[i];
      atomicOr(&vbits[e >> 5], 1u << (e & 31));
    }
    __syncthreads();
    if (tid == 0) { S.vsel ^= 1; S.vcount = nout; }
    __syncthreads();
  }
  __device__ void v_clear() {  // after v_compact: clear bits, empty list
    const uint32_t nin = S.vcount;
    const uint32_t* src = vl[S.vsel];
    for (uint32_t i = tid; i < nin; i += kReduceThreads) {
      uint32_t e = src[i];
      atomicAnd(&vbits[e >> 5], ~(1u << (e & 31)));
    }
    __syncthreads();
    if (tid == 0) S.vcount = 0;
    __syncthreads();
  }

  // ---- pivot hash map (thread 0)
  __device__ int hash_find(K key) {
    uint32_t h = (uint32_t)((uint64_t)key * 0x9E3779B97F4A7C15ull >> 32) & (uint32_t)(P.hcap - 1);
    for (;;) {
      K k = hkeys[h];
      if (k == key) return hvals[h];
      if (k == TR::maxv()) return -1;
      h = (h + 1) & (uint32_t)(P.hcap - 1);
    }
  }
  __device__ void hash_insert(K key, int val) {
    uint32_t h = (uint32_t)((uint64_t)key * 0x9E3779B97F4A7C15ull >> 32) & (uint32_t)(P.hcap - 1);
    while (hkeys[h] != TR::maxv()) h = (h + 1) & (uint32_t)(P.hcap - 1);
    hkeys[h] = key;
    hvals[h] = val;
  }

  // sort blist[0..nb) ascending in place (bitonic, global memory)
  __device__ void sort_blist(int* bl, int nb) {
    int np2 = 1;
    while (np2 < nb) np2 <<= 1;
    for (int i = nb + tid; i < np2; i += kReduceThreads) bl[i] = 0x7fffffff;  // cap1 is a power of two >= nb
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < np2; i += kReduceThreads) {
          int ixj = i ^ j;
          if (ixj > i) {
            int a = bl[i], b = bl[ixj];
            bool up = ((i & k) == 0);
            if ((a > b) == up) { bl[i] = b; bl[ixj] = a; }
          }
        }
        __syncthreads();
      }
  }

  __device__ void run_problem(int p) {
    R = P.rank + (size_t)p * n * n;
    EN = P.ends + (size_t)p * P.E;
    A = P.apex + (size_t)p * P.E;
    T = P.T[p];
    hkeys = (K*)P.hkeys + (size_t)p * P.hcap;
    hvals = P.hvals + (size_t)p * P.hcap;
    const float* SD = P.sdist + (size_t)p * P.E;
    int* bl = P.blist + (size_t)p * P.cap1;
    int nb = P.bcount[p];
    unsigned long long* st = P.stats + (size_t)p * ST_N;
    if (nb > P.cap1) {
      if (tid == 0) { P.counts[p * 4 + 3] = TDA_ERR_CAPACITY; P.counts[p * 4 + 1] = 0; }
      return;
    }
    sort_blist(bl, nb);
    for (int i = tid; i < P.hcap; i += kReduceThreads) hkeys[i] = TR::maxv();
    if (tid == 0) { S.pushes = 0; S.pops = 0; S.vcount = 0; S.vsel = 0; S.abort_flag = 0; }
    heap_reset();
    __syncthreads();
    const K kmax = (K)T * (K)n - 1;
    const uint64_t span0 = (uint64_t)max(T / 64, 256) * (uint64_t)n;
    int nrows = 0;
    int64_t vpool_used = 0;
    unsigned long long additions = 0, extensions = 0, maxv = 0;
    long long cyc[6] = {0, 0, 0, 0, 0, 0};
    unsigned long long badd_edges = 0, ext_edges = 0;
    long long t0;
    float* out = P.h1_pairs + (size_t)p * P.cap1 * 2;
    int64_t* outs = P.h1_simplex ? P.h1_simplex + (size_t)p * P.cap1 * 2 : nullptr;

    for (int ci = nb - 1; ci >= 0; --ci) {
      const int rbirth = bl[ci];
      const K start = (K)(rbirth + 1) * (K)n - 1;
      K H = ((uint64_t)(kmax - start) > span0) ? (K)(start + (K)span0) : kmax;
      uint64_t span = span0;
      f_clear();
      if (tid == 0) { S.base = start; S.fmax = start; v_toggle_single(rbirth); }
      __syncthreads();
      gen_push(rbirth, start, H);
      bool essential = false;
      K pivot = 0;
      for (;;) {
        if (S.abort_flag) break;
        K pk;
        t0 = clock64();
        bool ok = extract(pk);
        cyc[0] += clock64() - t0;
        if (!ok) {
          if (H >= kmax) { essential = true; break; }
          // extend the horizon: re-enumerate V for keys in (H, H2]
          t0 = clock64();
          span = span * 4;
          K H2 = ((uint64_t)(kmax - H) > span) ? (K)(H + (K)span) : kmax;
          v_compact();
          const uint32_t nv = S.vcount;
          const uint32_t* list = vl[S.vsel];
          for (uint32_t i = 0; i < nv; ++i) gen_push((int)list[i], H, H2);
          H = H2;
          ++extensions;
          ext_edges += nv;
          cyc[4] += clock64() - t0;
          continue;
        }
        // owner of the pivot
        t0 = clock64();
        const int M = (int)(pk / (K)n);
        const int w = n - 1 - (int)(pk % (K)n);
        if (tid == 0) {
          int kind = -1;  // -1 none, -2 apparent, >=0 reduced column id
          if (A[M] == w) kind = -2; else kind = hash_find(pk);
          S.bc_int = kind;
        }
        __syncthreads();
        const int owner = S.bc_int;
        __syncthreads();
        cyc[1] += clock64() - t0;
        if (owner == -1) { pivot = pk; break; }
        ++additions;
        t0 = clock64();
        if (owner == -2) {
          if (tid == 0) v_toggle_single(M);
          __syncthreads();
          gen_push(M, pk, H);
          cyc[2] += clock64() - t0;
        } else {
          const int64_t vs = P.vstart[(size_t)p * P.cap1 + owner];
          const int vn = P.vlen[(size_t)p * P.cap1 + owner];
          const uint32_t* ov = P.vpool + (size_t)p * P.vpool_cap + vs;
          if (S.vcount + (uint32_t)vn > (uint32_t)P.vcap) v_compact();
          for (int i = 0; i < vn; ++i) {
            int re = (int)ov[i];
            if (tid == 0) v_toggle_single(re);
            gen_push(re, pk, H);
          }
          badd_edges += vn;
          cyc[3] += clock64() - t0;
        }
      }
      if (S.abort_flag) break;
      t0 = clock64();
      // finalise the column
      v_compact();
      const uint32_t nv = S.vcount;
      if (nv > maxv) maxv = nv;
      if (!essential) {
        // store V (all edges, including the column's own) for later additions
        if (vpool_used + nv > P.vpool_cap) { if (tid == 0) S.abort_flag = TDA_ERR_CAPACITY; __syncthreads(); break; }
        uint32_t* dst = P.vpool + (size_t)p * P.vpool_cap + vpool_used;
        const uint32_t* list = vl[S.vsel];
        for (uint32_t i = tid; i < nv; i += kReduceThreads) dst[i] = list[i];
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
      if (!essential) { Md = (int)(pivot / (K)n); wd = n - 1 - (int)(pivot % (K)n); death = SD[Md]; }
      if (essential || death > birth) {
        if (tid == 0) {
          out[2 * nrows] = birth; out[2 * nrows + 1] = death;
          if (outs) {
            uint32_t e = EN[rbirth];
            outs[2 * nrows] = edge_index((int)(e >> 16), (int)(e & 0xffffu));
            if (essential) outs[2 * nrows + 1] = -1;
            else {
              uint32_t em = EN[Md];
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
      heap_release();
      cyc[5] += clock64() - t0;
    }
    __syncthreads();
    if (S.abort_flag) {  // leave the scratch clean for the next problem
      v_compact();
      v_clear();
      heap_release();
    }
    if (tid == 0) {
      P.counts[p * 4 + 1] = nrows;
      P.counts[p * 4 + 3] = S.abort_flag;
      st[ST_REDUCED] = (unsigned long long)nb;
      st[ST_ADDITIONS] = additions;
      st[ST_PUSHES] = S.pushes;
      st[ST_POPS] = S.pops;
      st[ST_EXTENSIONS] = extensions;
      st[ST_MAXV] = maxv;
      for (int q = 0; q < 6; ++q) st[ST_CYC_EXTRACT + q] = (unsigned long long)cyc[q];
      st[ST_BADD_EDGES] = badd_edges;
      st[ST_EXT_EDGES] = ext_edges;
    }
    __syncthreads();
  }
};

template <typename K>
__global__ void __launch_bounds__(kReduceThreads, 1) reduce_kernel(ReduceParams P) {
  __shared__ ReduceSmem<K> S;
  Reducer<K> red(P, S);
  if (threadIdx.x == 0) { S.freehead = kNil; S.fcache_n = 0; S.nzmask[0] = S.nzmask[1] = S.nzmask[2] = 0; }
  __syncthreads();
  for (;;) {
    if (threadIdx.x == 0) S.problem = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int p = S.problem;
    __syncthreads();
    if (p >= P.batch) break;
    red.run_problem(p);
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
  uint8_t* mst; int* mstlist; int* apex; int* blist; int* bcount;
  uint32_t *comp, *parent, *cbest; int* done;
  uint32_t* vbits; int64_t vwords; uint32_t* vlist; int64_t vcap;
  void* hkeys; int* hvals; int hcap;
  uint32_t* vpool; int64_t vpool_cap; int64_t* vstart; int* vlen;
  void* pool_keys; uint32_t* pool_next; uint32_t pool_chunks; unsigned int* pool_top;
  int* work_counter; unsigned long long* stats;
  int grid; size_t total;
  bool wide;  // 64-bit triangle keys
};

static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

static Layout make_layout(void* ws, int n, int batch, int maxdim, int cap1, size_t pool_bytes, int sm_count) {
  Layout L;
  memset(&L, 0, sizeof(L));
  const int64_t E = (int64_t)n * (n - 1) / 2;
  const int64_t BE = (int64_t)batch * E;
  Carver c(ws, ~size_t(0));
  L.wide = ((double)E * (double)n >= 4294967295.0);
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
  L.mstlist = c.take<int>((int64_t)batch * n);
  L.comp = c.take<uint32_t>((int64_t)batch * n);
  L.parent = c.take<uint32_t>((int64_t)batch * n);
  L.cbest = c.take<uint32_t>((int64_t)batch * n);
  L.done = c.take<int>(batch);
  L.stats = c.take<unsigned long long>((int64_t)batch * ST_N);
  L.work_counter = c.take<int>(1);
  if (maxdim >= 1) {
    L.apex = c.take<int>(BE);
    L.blist = c.take<int>((int64_t)batch * cap1);
    L.bcount = c.take<int>(batch);
    L.grid = batch < 2 * sm_count ? batch : 2 * sm_count;
    if (L.grid < 1) L.grid = 1;
    L.vwords = (E + 31) / 32 + 1;
    L.vbits = c.take<uint32_t>((int64_t)L.grid * L.vwords);
    L.vcap = E + 1024;
    L.vlist = c.take<uint32_t>((int64_t)L.grid * 2 * L.vcap);
    L.hcap = next_pow2(2 * cap1);
    L.hkeys = L.wide ? (void*)c.take<uint64_t>((int64_t)batch * L.hcap) : (void*)c.take<uint32_t>((int64_t)batch * L.hcap);
    L.hvals = c.take<int>((int64_t)batch * L.hcap);
    L.vpool_cap = (int64_t)(pool_bytes / 8 / (size_t)batch / sizeof(uint32_t));
    if (L.vpool_cap < 4 * (int64_t)cap1) L.vpool_cap = 4 * (int64_t)cap1;
    L.vpool = c.take<uint32_t>((int64_t)batch * L.vpool_cap);
    L.vstart = c.take<int64_t>((int64_t)batch * cap1);
    L.vlen = c.take<int>((int64_t)batch * cap1);
    const size_t ksz = L.wide ? 8 : 4;
    size_t chunks = pool_bytes / (kChunk * ksz);
    if (chunks < (size_t)L.grid * 80) chunks = (size_t)L.grid * 80;
    if (chunks > 0x7ffffff0u) chunks = 0x7ffffff0u;
    L.pool_chunks = (uint32_t)chunks;
    L.pool_keys = c.take<char>((chunks + 1) * kChunk * ksz);
    L.pool_next = c.take<uint32_t>(chunks + 1);
    L.pool_top = c.take<unsigned int>(1);
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

extern "C" int tda_rips(const float* dm, int n, int batch, int maxdim, float thresh, float* h0_pairs, int64_t* h0_simplex,
                        float* h1_pairs, int64_t* h1_simplex, int cap1, int32_t* counts, float* thresh_out, void* ws,
                        size_t ws_bytes, size_t pool_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!dm || !h0_pairs || !counts || !ws || n <= 0 || batch <= 0) return set_error(TDA_ERR_INVALID, "tda_rips: bad arguments");
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

  TDA_CUDA_CHECK(cudaMemsetAsync(L.T, 0, sizeof(int) * batch, stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.stats, 0, sizeof(unsigned long long) * batch * ST_N, stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.work_counter, 0, sizeof(int), stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 4 * batch, stream));
  enclosing_init_kernel<<<(batch + 255) / 256, 256, 0, stream>>>(L.thresh_bits, batch, thresh);
  count_launch();
  if (isinf(thresh)) {
    dim3 g((n * 32 + 255) / 256, batch);
    enclosing_kernel<<<g, 256, 0, stream>>>(dm, n, L.thresh_bits);
    count_launch();
  }
  TDA_LAUNCH_CHECK();
  if (thresh_out) TDA_CUDA_CHECK(cudaMemcpyAsync(thresh_out, L.thresh_bits, sizeof(float) * batch, cudaMemcpyDeviceToDevice, stream));
  if (E > 0) {
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
  TDA_CUDA_CHECK(cudaMemsetAsync(L.mst, 0, (size_t)(BE > 0 ? BE : 1), stream));
  {
    dim3 gi((n + 255) / 256, batch);
    boruvka_init_kernel<<<gi, 256, 0, stream>>>(n, L.comp, L.cbest, L.done);
    count_launch();
    int rounds = 1;
    while ((1 << rounds) < n) ++rounds;
    ++rounds;  // one extra round detects "no merge" and is a no-op otherwise
    dim3 gs((n + 7) / 8, batch);
    for (int r = 0; r < rounds && E > 0; ++r) {
      boruvka_scan_kernel<<<gs, 256, 0, stream>>>(L.rank, L.T, n, L.comp, L.cbest, L.done);
      boruvka_merge_kernel<<<batch, 1024, 0, stream>>>(L.ends, n, E, L.comp, L.parent, L.cbest, L.mst, L.done);
      count_launch(2);
    }
    h0_emit_kernel<<<batch, 1024, 0, stream>>>(L.ends, L.sdist, L.T, n, E, L.mst, L.comp, h0_pairs, h0_simplex, counts, L.mstlist);
    count_launch();
    TDA_LAUNCH_CHECK();
  }
  if (maxdim >= 1 && E > 0) {
    TDA_CUDA_CHECK(cudaMemsetAsync(L.bcount, 0, sizeof(int) * batch, stream));
    TDA_CUDA_CHECK(cudaMemsetAsync(L.pool_top, 0, sizeof(unsigned int), stream));
    TDA_CUDA_CHECK(cudaMemsetAsync(L.vbits, 0, sizeof(uint32_t) * (size_t)L.grid * L.vwords, stream));
    dim3 g((unsigned)((E * 32 + 255) / 256), batch);
    apparent_kernel<<<g, 256, 0, stream>>>(L.rank, L.ends, L.mst, L.T, n, E, L.apex, L.blist, L.bcount, cap1p, L.stats);
    count_launch();
    TDA_LAUNCH_CHECK();
    ReduceParams P;
    P.rank = L.rank; P.ends = L.ends; P.sdist = L.sdist; P.T = L.T; P.apex = L.apex; P.blist = L.blist; P.bcount = L.bcount;
    P.n = n; P.E = E; P.batch = batch; P.cap1 = cap1p;
    P.h1_pairs = h1_pairs; P.h1_simplex = h1_simplex; P.counts = counts;
    P.vbits = L.vbits; P.vwords = L.vwords; P.vlist = L.vlist; P.vcap = L.vcap;
    P.hkeys = L.hkeys; P.hvals = L.hvals; P.hcap = L.hcap;
    P.vpool = L.vpool; P.vpool_cap = L.vpool_cap; P.vstart = L.vstart; P.vlen = L.vlen;
    P.pool_keys = L.pool_keys; P.pool_next = L.pool_next; P.pool_chunks = L.pool_chunks; P.pool_top = L.pool_top;
    P.work_counter = L.work_counter; P.stats = L.stats;
    if (L.wide) reduce_kernel<uint64_t><<<L.grid, kReduceThreads, 0, stream>>>(P);
    else reduce_kernel<uint32_t><<<L.grid, kReduceThreads, 0, stream>>>(P);
    count_launch();
    TDA_LAUNCH_CHECK();
    finalize_stats_kernel<<<(batch + 255) / 256, 256, 0, stream>>>(L.T, L.bcount, n, batch, L.stats, counts);
    count_launch();
  }
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

extern "C" int tda_rips_stats(const void* ws, int n, int batch, int maxdim, int cap1, size_t pool_bytes, int64_t* stats_host) {
  if (!ws || !stats_host) return set_error(TDA_ERR_INVALID, "tda_rips_stats: bad arguments");
  Layout L = make_layout((void*)ws, n, batch, maxdim, next_pow2(cap1), pool_bytes, sm_count_cached());
  TDA_CUDA_CHECK(cudaMemcpy(stats_host, L.stats, sizeof(int64_t) * ST_N * batch, cudaMemcpyDeviceToHost));
  return TDA_OK;
}
