// rips_sweep2.cuh -- residual H1 reduction, "substitute by rank, verify by window" (the default reducer; included by rips.cu).
//
// Same reduction as ripser's (same column order, same pivots), organised so that nothing in it is a chain of dependent pivots
// (CPU model, checked pair for pair against the ripser restatement: oracle/rips_propagate_model.cpp, rips_model_h1_modes):
//
//  * x_e = 1 iff edge e (by rank) is in the reduction column V.  When the sweep passes row M of an apparent edge M=(c,d) with
//    apex a, the reduction ends with  x_M = x_(c,a) ^ x_(d,a)  whatever it was before: the apparent-pair additions are a forward
//    SUBSTITUTION over a static graph (two parent edges per apparent edge, `par`, written by parents_kernel).  A window of rows
//    [pos, hi) is substituted in rounds: round 1 takes every row whose parents are final (below the window, or rows that the
//    substitution does not change), the later rounds run in shared memory on the few rows that wait for a row of the window.
//  * verification: V must be a cocycle of the complex below the cursor, i.e. every row (x_M ^ X[c] ^ X[d]) & lune(M) must be
//    empty (X = V as a symmetric bit matrix).  Only rows with a touched endpoint can fail ("heavy" rows).  The candidate bits
//    of a row are (x_M ^ X[c] ^ X[d]); a candidate w is a true failure iff rank(c,w) < M and rank(d,w) < M (two probes).
//      sparse mode : few candidates per row; they are probed directly.
//      dense mode  : (>= 1/8 of a window heavy) the candidates are first masked with Pend[c] & Pend[d], Pend = adjacency of ALL
//                    edges below the window's end, kept as a bit matrix that only moves forward inside a column.  If the window
//                    leaves V a cocycle of the complex at its end, every row passes; a surviving bit (M, w) belongs to the
//                    triangle {c,d,w}, whose own row lies in the window, so probing the survivors finds the first true failure.
//    The smallest true failing key of the window is the event.
//  * event (M*, w*): the flips above M* are undone; unowned pivot -> death of the column; pivot owned by a reduced column j ->
//    V ^= V_j and the sweep resumes at row M* (whose substitution is then a no-op).
//  * windows grow (w0, doubling) while they are clean and shrink back after an event (sparse mode); in dense mode the window
//    keeps its end after an event.
//
// One thread-block CLUSTER per cloud at a time (dynamic work counter): every pass over the rows / heavy rows of a window is dealt
// to the warps of all CTAs of the cluster (the passes are bound by L2 / HBM latency: more warps = more loads in flight).  The
// window's bookkeeping (x words, done bits, pending and heavy lists, counters) lives in the shared memory of CTA 0 and is reached
// by the other CTAs through distributed shared memory; the touched-vertex bitmaps are replicated.  X, Pm, the x bits by rank and
// the V list are global (L2-resident).  All per-window passes are warp-per-32-rows with coalesced 8-byte loads of the row tables;
// global x bits are rewritten one word per 32 rows without atomics.

constexpr int kS2Threads = 512;
constexpr int kS2Warps = kS2Threads / 32;
constexpr int kS2MaxCluster = 8;
constexpr int kS2ListSmem = 2048;      // entries of the pending / heavy lists kept in shared memory (the rest spills to global scratch)
constexpr int kS2MaxRounds = 4096;     // substitution rounds per window (depth of the apparent graph is ~30): beyond -> internal error
constexpr int kS2MaxWindow = 65472;    // rows of a window (16-bit local indices in the pending entries)
constexpr int kS2Batch = 4;            // heavy rows a warp verifies at once
constexpr int kS2Unroll = 4;           // 32-row groups a warp keeps in flight in the first substitution round
constexpr int kS2ProbeMax = 256;       // candidate bits of a row up to which they are probed one by one (more: the whole lune at once)
enum { TDA_ERR_INTERNAL_S2 = -6 };
// warp engine (short, sparse columns: one warp per column, V in the lanes)
constexpr int kWcMaxV = 32;            // edges of V a warp holds (one per lane); more -> the column goes to the cluster engine
constexpr int kWcRec = 40;             // words of a column record: status, pos, key lo, key hi, nv (0xffffffff: V in the pool), column,
                                       // pool length, pool start, then up to 32 edge ranks.  Records are written once and never changed.
constexpr uint32_t kWcMaxRows = 1u << 18;   // rows a warp sweeps before it hands the column over
constexpr uint32_t kWcMaxHeavy = 2048;      // heavy rows a warp handles before it hands the column over
enum { WC_NONE = 0, WC_DEATH = 1, WC_ESSENTIAL = 2, WC_BIG = 3 };

struct Sweep2Smem {
  uint32_t vcount, vcount2, vsel;
  int abort_flag, problem;
  uint32_t npend[3];
  uint32_t nheavy, nfail, newtouch, nundone;
  uint32_t fail_key;           // smallest failing key of the window: (row - base_row) * n + (n - 1 - vertex), 32 bits (< 65536 * n)
  // column bookkeeping shared by the warp engine (stage A / commit loop) and the cluster engine
  int nrows;                   // H1 rows written so far
  int next_col;                // next entry of the round's work list to hand to a warp
  uint32_t nrec;               // records allocated
  uint32_t nact[2];            // entries of the two work lists (this round / next round)
  uint32_t nbig, nbig_done;    // columns waiting for the cluster engine / taken by it
  int ev_rec;                  // cluster engine: record of the column that owns the pivot of the event (-1: nobody -> death)
  int ev_slot;                 //                 its slot in the pivot table
  long long vpool_used;
  unsigned long long maxv, badd;
  unsigned long long wc_cols, wc_resumed, wc_big, wc_rows, wc_heavy, wc_scans;   // warp engine counters
  unsigned long long st[16];
};
enum { S2_WINDOWS = 0, S2_ROUNDS, S2_HEAVY, S2_FLIPS, S2_UNDONE, S2_EVENTS, S2_SPURIOUS, S2_SUBST, S2_LATE, S2_PM, S2_DENSE, S2_EXACT, S2_DEATHS };

__device__ __forceinline__ uint32_t s2_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t s2_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t s2_clusterid() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
// generic pointer to the same shared-memory object in CTA `rank` of the cluster
template <typename T>
__device__ __forceinline__ T* s2_map(T* p, uint32_t rank) {
  uint64_t out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((uint64_t)p), "r"(rank));
  return reinterpret_cast<T*>(out);
}

struct Sweeper2 {
  static constexpr uint64_t kEmpty = ~0ull;
  static constexpr unsigned kFull = 0xffffffffu;
  const ReduceParams& P;
  Sweep2Smem& S;               // CTA 0's control block (remote for the other CTAs)
  uint32_t *touched, *tnew;    // this CTA's copies of the touched-vertex bitmaps
  uint32_t *xs, *xo, *done;    // CTA 0's window bitmaps
  uint2 *pend_s, *heavy_s;     // CTA 0's lists: [2][kS2ListSmem], [kS2ListSmem]
  const int tid, lane, warp;
  const uint32_t crank, csize;
  const int gtid, gwarp, nthreads, nwarps;   // cluster-wide thread / warp ids and counts
  const int slot;                           // scratch slot of this cluster
  const int* R; const uint32_t* EN; const uint2* EA; const uint2* PAR; int T; int n; int W;
  uint32_t *X, *Pm, *vbits, *vl0;
  uint2 *pend_g, *heavy_g;
  uint32_t* wcd;               // this warp's vertex bitmap (warp engine), W words of this CTA's shared memory
  uint32_t* recs;              // column records of this cluster's cloud: [s2_nrec][kWcRec] words (global)
  int* act;                    // work lists of the rounds: [2][s2_nrec] records to reduce further (global)
  int* biglist;                // [s2_nrec] records of columns for the cluster engine
  int* cur_rec;                // [cap1] final (latest) record of every column
  int act_next;                // which work list the displaced columns go to
  uint32_t p_pos; bool p_valid;
  uint64_t* hkeys; int* hvals;
  long long cyc_sync; unsigned long long n_sync;   // (diagnostics) cycles this thread spent in cluster barriers, and their number

  __device__ Sweeper2(const ReduceParams& p, Sweep2Smem& s_local, uint32_t* dyn)
      : P(p), S(*s2_map(&s_local, 0)), tid(threadIdx.x), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), crank(s2_ctarank()), csize(s2_nctarank()),
        gtid((int)(s2_ctarank() * kS2Threads + threadIdx.x)), gwarp((int)(s2_ctarank() * kS2Warps + (threadIdx.x >> 5))),
        nthreads((int)(s2_nctarank() * kS2Threads)), nwarps((int)(s2_nctarank() * kS2Warps)), slot((int)s2_clusterid()) {
    n = P.n;
    W = P.xw;
    const int nw = P.s2_wmax / 32 + 4;
    touched = dyn;
    tnew = touched + W;
    uint32_t* dyn0 = s2_map(dyn, 0);
    xs = dyn0 + 2 * W;
    xo = xs + nw;
    done = xo + nw;
    pend_s = reinterpret_cast<uint2*>(done + nw + ((2 * W + 3 * nw) & 1));   // 8-byte aligned
    heavy_s = pend_s + 2 * kS2ListSmem;
    {
      uint2* pend_local = reinterpret_cast<uint2*>(dyn + 2 * W + 3 * nw + ((2 * W + 3 * nw) & 1));
      wcd = reinterpret_cast<uint32_t*>(pend_local + 3 * kS2ListSmem) + (size_t)warp * W;
    }
    X = P.xmat + (size_t)slot * (size_t)n * W;
    Pm = P.pmat + (size_t)slot * (size_t)n * W;
    p_pos = 0; p_valid = false;
    cyc_sync = 0; n_sync = 0;
    vbits = P.vbits + (size_t)slot * P.vwords;
    vl0 = P.vlist + (size_t)slot * 2 * P.vcap;
    pend_g = P.s2_pend + (size_t)slot * 2 * (size_t)(P.s2_wmax + 64);
    heavy_g = P.s2_heavy + (size_t)slot * (size_t)(P.s2_wmax + 64);
    recs = P.s2_rec + (size_t)slot * (size_t)P.s2_nrec * kWcRec;
    act = P.s2_lists + (size_t)slot * (size_t)(3 * P.s2_nrec + P.cap1);
    biglist = act + 2 * (size_t)P.s2_nrec;
    cur_rec = biglist + P.s2_nrec;
    act_next = 1;
  }
  static __host__ __device__ size_t dyn_bytes(int W, int wmax) {
    const int nw = wmax / 32 + 4;
    return sizeof(uint32_t) * (size_t)(2 * W + 3 * nw + 2) + sizeof(uint2) * (size_t)(3 * kS2ListSmem) + sizeof(uint32_t) * (size_t)kS2Warps * W;
  }
  // barrier over the cluster (release / acquire: shared-memory and global writes of every CTA are visible afterwards)
  __device__ __forceinline__ void csync() {
    const long long t = clock64();
    if (csize > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
    cyc_sync += clock64() - t;
    ++n_sync;
  }
  __device__ __forceinline__ uint32_t* vlist(uint32_t sel) const { return vl0 + (size_t)sel * P.vcap; }
  __device__ __forceinline__ bool tbit(uint32_t v) const { return (touched[v >> 5] >> (v & 31)) & 1u; }
  __device__ __forceinline__ bool tnewbit(uint32_t v) const { return (tnew[v >> 5] >> (v & 31)) & 1u; }
  __device__ __forceinline__ uint32_t xg(uint32_t e) const { return (__ldcg(&vbits[e >> 5]) >> (e & 31)) & 1u; }
  __device__ __forceinline__ uint2* pend_ref(int buf, uint32_t i) const {
    return i < (uint32_t)kS2ListSmem ? &pend_s[(size_t)buf * kS2ListSmem + i] : &pend_g[(size_t)buf * (P.s2_wmax + 64) + (i - kS2ListSmem)];
  }
  __device__ __forceinline__ uint2* heavy_ref(uint32_t i) const { return i < (uint32_t)kS2ListSmem ? &heavy_s[i] : &heavy_g[i - kS2ListSmem]; }
  __device__ __forceinline__ void fail(int code) { S.abort_flag = code; }

  __device__ __forceinline__ void x_flip(uint32_t c, uint32_t d) {
    atomicXor(&X[(size_t)c * W + (d >> 5)], 1u << (d & 31));
    atomicXor(&X[(size_t)d * W + (c >> 5)], 1u << (c & 31));
  }
  // vertex v becomes touched: in every CTA's bitmap (the new-in-this-window bitmap tnew sits W words behind touched)
  __device__ __forceinline__ void touch(uint32_t v) {
    const uint32_t m = 1u << (v & 31);
    if (!(touched[v >> 5] & m)) {
      const uint32_t old = atomicOr(&touched[v >> 5], m);
      if (!(old & m)) {
        for (uint32_t r = 0; r < csize; ++r) {
          uint32_t* tr = s2_map(touched, r);
          if (r != crank) atomicOr(&tr[v >> 5], m);
          atomicOr(&tr[W + (v >> 5)], m);
        }
        S.newtouch = 1;
      }
    }
  }
  // edge e=(c,d) toggles in V: x bit, X, touched, V list (any thread; duplicates in the list are fine, v_compact drops them)
  __device__ __forceinline__ void toggle_edge(uint32_t e, uint32_t c, uint32_t d) {
    atomicXor(&vbits[e >> 5], 1u << (e & 31));
    x_flip(c, d);
    touch(c); touch(d);
    const uint32_t pos = atomicAdd(&S.vcount, 1u);
    if (pos < (uint32_t)P.vcap) vlist(S.vsel)[pos] = e;
    else fail(TDA_ERR_CAPACITY);
  }
  // V list -> the distinct edges with x = 1 (x bits are exact; the list may hold an edge several times or with x = 0)
  __device__ __forceinline__ void v_compact() {
    csync();
    const uint32_t nin = min(S.vcount, (uint32_t)P.vcap);
    const uint32_t sel = S.vsel;
    const uint32_t* src = vlist(sel);
    uint32_t* dst = vlist(sel ^ 1);
    csync();
    if (gtid == 0) S.vcount2 = 0;
    csync();
    for (uint32_t i0 = 0; i0 < nin; i0 += (uint32_t)nthreads) {
      const uint32_t i = i0 + (uint32_t)gtid;
      bool keep = false;
      uint32_t e = 0;
      if (i < nin) {
        e = __ldcg(&src[i]);
        const uint32_t m = 1u << (e & 31);
        keep = (atomicAnd(&vbits[e >> 5], ~m) & m) != 0;
      }
      const unsigned bal = __ballot_sync(kFull, keep);
      uint32_t bs = 0;
      if (lane == 0 && bal) bs = atomicAdd(&S.vcount2, (uint32_t)__popc(bal));
      bs = __shfl_sync(kFull, bs, 0);
      if (keep) __stcg(&dst[bs + __popc(bal & ((1u << lane) - 1))], e);
    }
    __threadfence();
    csync();
    const uint32_t nout = S.vcount2;
    for (uint32_t i = (uint32_t)gtid; i < nout; i += (uint32_t)nthreads) {
      const uint32_t e = __ldcg(&dst[i]);
      atomicOr(&vbits[e >> 5], 1u << (e & 31));
    }
    __threadfence();
    csync();
    if (gtid == 0) { S.vsel = sel ^ 1; S.vcount = nout; }
    csync();
  }
  // after v_compact: x bits, X bits and the touched masks back to zero; the list empty
  __device__ __forceinline__ void v_clear() {
    const uint32_t nin = S.vcount;
    const uint32_t* src = vlist(S.vsel);
    for (uint32_t i = (uint32_t)gtid; i < nin; i += (uint32_t)nthreads) {
      const uint32_t e = __ldcg(&src[i]);
      atomicAnd(&vbits[e >> 5], ~(1u << (e & 31)));
      const uint32_t en = __ldg(&EN[e]);
      x_flip(en >> 16, en & 0xffffu);
    }
    for (int i = tid; i < W; i += kS2Threads) { touched[i] = 0; tnew[i] = 0; }
    __threadfence();
    csync();
    if (gtid == 0) S.vcount = 0;
    csync();
  }
  // ---- pivot table (open addressing): key = death triangle, value = RECORD of the column that currently owns it (-1: nobody).
  // Lock-free reduction (Morozov & Nigmetov): columns are reduced in any order; a column b that arrives at a pivot owned by a
  // column with a larger birth (earlier in ripser's order) adds that column's V and goes on; otherwise it takes the pivot, and
  // the column it displaces is reduced further in the next round.  Every addition is "earlier column into later column", and at
  // the fixpoint all pivots are distinct, so the pairing is THE persistence pairing (same pairs as the sequential order).
  __device__ __forceinline__ int tab_slot(uint64_t key) {   // find or insert (one thread)
    uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & (uint32_t)(P.hcap - 1);
    for (int guard = 0; guard < P.hcap; ++guard) {
      const uint64_t k = __ldcg(&hkeys[h]);
      if (k == key) return (int)h;
      if (k == kEmpty) {
        const unsigned long long old = atomicCAS((unsigned long long*)&hkeys[h], (unsigned long long)kEmpty, (unsigned long long)key);
        if (old == kEmpty || old == key) return (int)h;
      }
      h = (h + 1) & (uint32_t)(P.hcap - 1);
    }
    fail(TDA_ERR_CAPACITY);
    return 0;
  }
  __device__ __forceinline__ uint32_t* rec_ptr(int rid) const { return recs + (size_t)rid * kWcRec; }
  __device__ __forceinline__ int rec_col(int rid) const { return rid < 0 ? -1 : (int)__ldcg(&rec_ptr(rid)[5]); }
  __device__ __forceinline__ int rec_alloc() {   // one thread
    const uint32_t r = atomicAdd(&S.nrec, 1u);
    if (r >= (uint32_t)P.s2_nrec) { fail(TDA_ERR_CAPACITY); return (int)P.s2_nrec - 1; }
    return (int)r;
  }
  __device__ __forceinline__ void act_push(int which, int rid) {   // one thread: record `rid` is reduced further in the next round
    const uint32_t k = atomicAdd(&S.nact[which], 1u);
    if (k < (uint32_t)P.s2_nrec) __stcg(&act[(size_t)which * P.s2_nrec + k], rid);
    else fail(TDA_ERR_CAPACITY);
  }
  // CTA 0 sorts the birth list (global memory, bitonic)
  __device__ __forceinline__ void sort_blist(int* bl, int nb) {
    if (crank == 0) {
      int np2 = 1;
      while (np2 < nb) np2 <<= 1;
      for (int i = nb + tid; i < np2; i += kS2Threads) bl[i] = 0x7fffffff;
      __syncthreads();
      for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < np2; i += kS2Threads) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const int a = bl[i], b = bl[ixj];
              const bool up = ((i & k) == 0);
              if ((a > b) == up) { bl[i] = b; bl[ixj] = a; }
            }
          }
          __syncthreads();
        }
      __threadfence();
    }
  }

  // ---- Pm = adjacency bit matrix of the edges with rank < p_pos.  All threads of the cluster; ends with the bits performed and a barrier.
  __device__ __forceinline__ void p_set_rows(uint32_t lo, uint32_t hi, bool set) {
    for (uint32_t row0 = lo + (uint32_t)gtid; row0 < hi; row0 += 4u * (uint32_t)nthreads) {
      uint32_t en[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const uint32_t row = row0 + (uint32_t)(u * nthreads); en[u] = row < hi ? __ldg(&EN[row]) : 0xffffffffu; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (en[u] == 0xffffffffu) continue;
        const uint32_t c = en[u] >> 16, d = en[u] & 0xffffu;
        if (set) {
          atomicOr(&Pm[(size_t)c * W + (d >> 5)], 1u << (d & 31));
          atomicOr(&Pm[(size_t)d * W + (c >> 5)], 1u << (c & 31));
        } else {
          atomicAnd(&Pm[(size_t)c * W + (d >> 5)], ~(1u << (d & 31)));
          atomicAnd(&Pm[(size_t)d * W + (c >> 5)], ~(1u << (c & 31)));
        }
      }
    }
  }
  __device__ __forceinline__ void p_move(uint32_t target) {
    const uint32_t dist = target > p_pos ? target - p_pos : p_pos - target;
    if (!p_valid || (uint64_t)dist * 32ull > (uint64_t)n * (uint64_t)n) {
      // rebuild from the rank matrix: one warp per vertex row, a word per ballot
      for (int c = gwarp; c < n; c += nwarps) {
        const int* Rc = R + (size_t)c * n;
        for (int k0 = 0; k0 < W; k0 += 32) {
          uint32_t mine = 0;
          const int kend = min(32, W - k0);
#pragma unroll 8
          for (int kk = 0; kk < kend; ++kk) {
            const int w = (k0 + kk) * 32 + lane;
            const int ra = w < n ? __ldg(&Rc[w]) : kRankDiag;
            const unsigned word = __ballot_sync(kFull, ra < (int)target);
            if (lane == kk) mine = word;
          }
          if (k0 + lane < W) __stcg(&Pm[(size_t)c * W + k0 + lane], mine);
        }
      }
      p_valid = true;
      if (gtid == 0) S.st[S2_PM] += (unsigned long long)n * n / 32;
    } else if (target > p_pos) {
      p_set_rows(p_pos, target, true);
      if (gtid == 0) S.st[S2_PM] += dist;
    } else if (target < p_pos) {
      p_set_rows(target, p_pos, false);
      if (gtid == 0) S.st[S2_PM] += dist;
    }
    p_pos = target;
    __threadfence();
    csync();
  }

  // ---- one row with its exact lune { w : rank(c,w) < M and rank(d,w) < M } (one warp): the highest failing vertex, or -1
  __device__ __forceinline__ int exact_row(uint32_t M, uint32_t c, uint32_t d, uint32_t xm) const {
    const int* Rc = R + (size_t)c * n;
    const int* Rd = R + (size_t)d * n;
    const uint32_t* Xc = X + (size_t)c * W;
    const uint32_t* Xd = X + (size_t)d * W;
    int best = -1;
    for (int k0 = 0; k0 < W; k0 += 32) {   // 32 words = 1024 vertices per block; lane kk ends up with the lune word k0 + kk
      const int kend = min(32, W - k0);
      const int k = k0 + lane;
      const uint32_t xx = k < W ? (__ldcg(&Xc[k]) ^ __ldcg(&Xd[k])) : 0u;
      uint32_t lmine = 0;
      for (int kk0 = 0; kk0 < kend; kk0 += 16) {
        int ra[16], rb[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int w = (k0 + kk0 + u) * 32 + lane;
          const bool ok = (kk0 + u) < kend && w < n;
          ra[u] = ok ? __ldg(&Rc[w]) : kRankDiag;
          rb[u] = ok ? __ldg(&Rd[w]) : kRankDiag;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const unsigned word = __ballot_sync(kFull, ra[u] < (int)M && rb[u] < (int)M);
          if (lane == kk0 + u) lmine = word;
        }
      }
      const uint32_t v = (xm ^ xx) & lmine;
      if (v) best = k * 32 + 31 - __clz(v);
    }
    return __reduce_max_sync(kFull, best);
  }

  // ---- appends one entry per flagged lane to a list (warp-aggregated).  `cnt` lives in CTA 0's shared memory
  template <typename Ref>
  __device__ __forceinline__ void warp_append(bool flag, uint2 ent, uint32_t* cnt, Ref ref) {
    const unsigned bal = __ballot_sync(kFull, flag);
    if (!bal) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(cnt, (uint32_t)__popc(bal));
    base = __shfl_sync(kFull, base, 0);
    if (flag) *ref(base + __popc(bal & ((1u << lane) - 1))) = ent;
  }

  // =================================================================================================================
  // Warp engine: one warp reduces one column whose V stays small (<= 32 edges, one per lane: rank `vr`, endpoints `ven`).
  // The rows are taken strictly in order, 32 at a time; a row matters only if one of its endpoints is an endpoint of an edge of V
  // ("heavy").  For a heavy apparent row M=(c,d): substitution x_M = [pa in V] ^ [pb in V] (toggle M in V if it differs), then
  // the row is verified at once: candidates w with x_cw ^ x_dw ^ x_M = 1 come from the edges of V at c and d (x_M = 0: probed
  // one by one) or are the whole lune minus those (x_M = 1: one scan of the two rank rows).  The first failing (M, w) ends the
  // sweep: a tentative death (stage A, no look-up) or, with `owned` look-ups (commit loop), the owner's V is added and the row
  // verified again.  Columns that outgrow the warp (V, rows, heavy rows) are handed to the cluster engine with their state.
  struct WcState { uint32_t vr, ven; int nv; };
  __device__ __forceinline__ bool wc_in(const WcState& v, uint32_t r) const { return __ballot_sync(kFull, lane < v.nv && v.vr == r) != 0; }
  __device__ __forceinline__ bool wc_toggle(WcState& v, uint32_t r, uint32_t en) const {
    const unsigned hit = __ballot_sync(kFull, lane < v.nv && v.vr == r);
    if (hit) {   // remove: the last edge moves into its lane
      const int l = __ffs(hit) - 1;
      const uint32_t lr = __shfl_sync(kFull, v.vr, v.nv - 1), le = __shfl_sync(kFull, v.ven, v.nv - 1);
      if (lane == l) { v.vr = lr; v.ven = le; }
      --v.nv;
      return true;
    }
    if (v.nv >= kWcMaxV) return false;
    if (lane == v.nv) { v.vr = r; v.ven = en; }
    ++v.nv;
    return true;
  }
  // this lane's edge of V as seen from row (c,d) (not the row's own edge `M`): the vertex w with (c,w) or (d,w) in V, or -1
  __device__ __forceinline__ int wc_candidate(const WcState& v, uint32_t c, uint32_t d, uint32_t M) const {
    if (lane >= v.nv || v.vr == M) return -1;
    const uint32_t a = v.ven >> 16, b = v.ven & 0xffffu;
    if (a == c) return (int)b;
    if (b == c) return (int)a;
    if (a == d) return (int)b;
    if (b == d) return (int)a;
    return -1;
  }
  // highest failing vertex of row M=(c,d) for the column V, or -1
  __device__ __forceinline__ int wc_verify(const WcState& v, uint32_t M, uint32_t c, uint32_t d, bool xm, uint32_t& scans) const {
    const int w = wc_candidate(v, c, d, M);
    if (!xm) {
      // x_M = 0: a vertex fails iff exactly one of (c,w), (d,w) is in V (a vertex seen from both sides cancels) and w is in the lune
      const unsigned same = __match_any_sync(kFull, w >= 0 ? (uint32_t)w : (0x80000000u | (uint32_t)lane));
      const bool odd = w >= 0 && (__popc(same) & 1);
      int best = -1;
      if (odd && __ldg(&R[(size_t)c * n + w]) < (int)M && __ldg(&R[(size_t)d * n + w]) < (int)M) best = w;
      return __reduce_max_sync(kFull, best);
    }
    // x_M = 1: every lune vertex must have exactly one of its two edges in V: D = those vertices (parity by XOR), fail = lune & ~D
    ++scans;
    for (int k = lane; k < W; k += 32) wcd[k] = 0u;
    __syncwarp();
    if (w >= 0) atomicXor(&wcd[w >> 5], 1u << (w & 31));
    __syncwarp();
    const int* Rc = R + (size_t)c * n;
    const int* Rd = R + (size_t)d * n;
    int best = -1;
    for (int k0 = 0; k0 < W; k0 += 32) {
      const int kend = min(32, W - k0);
      uint32_t lmine = 0;
      for (int kk0 = 0; kk0 < kend; kk0 += 16) {
        int ra[16], rb[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int ww = (k0 + kk0 + u) * 32 + lane;
          const bool ok = (kk0 + u) < kend && ww < n;
          ra[u] = ok ? __ldg(&Rc[ww]) : kRankDiag;
          rb[u] = ok ? __ldg(&Rd[ww]) : kRankDiag;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const unsigned word = __ballot_sync(kFull, ra[u] < (int)M && rb[u] < (int)M);
          if (lane == kk0 + u) lmine = word;
        }
      }
      const int k = k0 + lane;
      const uint32_t f = k < W ? (lmine & ~wcd[k]) : 0u;
      if (f) best = k * 32 + 31 - __clz(f);
    }
    __syncwarp();
    return __reduce_max_sync(kFull, best);
  }
  // the sweep.  In: V, first row `pos`.  Out: status, V (final, or the state to hand over), pos / key.
  __device__ __forceinline__ int wc_sweep(WcState& v, uint32_t& pos, unsigned long long& key, int b, int p) {
    uint32_t nheavy = 0, nrows = 0, nscan = 0;
    int status = WC_NONE;
    int myrec = -1;   // the record this sweep publishes (allocated when it first needs one)
    // the table entries of the NEXT 32-row group are requested while this group is examined: almost every group has no heavy
    // row and the sweep just moves on, so the L2 latency of the tables (the whole cost of such a group) overlaps the heavy test
    uint32_t pf_base = 0xffffffffu;
    uint2 pf_ea = make_uint2(0u, 0xffffffffu), pf_par = make_uint2(0u, 0u);
    const uint32_t max_rows = (uint32_t)P.s2_wc_max_rows;
    while (status == WC_NONE) {
      if (pos >= (uint32_t)T) { status = WC_ESSENTIAL; break; }
      if (nrows > max_rows || nheavy > kWcMaxHeavy) { status = WC_BIG; break; }
      const uint32_t base = pos & ~31u;
      const uint32_t row = base + lane;
      const bool inwin = row >= pos && row < (uint32_t)T;
      uint2 ea = make_uint2(0u, 0xffffffffu), par = make_uint2(0u, 0u);
      if (base == pf_base && pos == base) { ea = pf_ea; par = pf_par; }   // (a whole group: what was prefetched is exactly this)
      else if (inwin) { ea = __ldg(&EA[row]); par = __ldg(&PAR[row]); }
      {
        const uint32_t nrow = base + 32u + lane;
        pf_base = base + 32u;
        pf_ea = make_uint2(0u, 0xffffffffu); pf_par = make_uint2(0u, 0u);
        if (nrow < (uint32_t)T) { pf_ea = __ldg(&EA[nrow]); pf_par = __ldg(&PAR[nrow]); }
      }
      nrows += 32;
      const bool app = (int)ea.y >= 0;
      const uint32_t c = ea.x >> 16, d = ea.x & 0xffffu;
      uint32_t from = 0;   // rows of this group below `from` are done
      for (;;) {
        // heavy rows of the group at or above `from` under the current V
        bool h = false;
        for (int e = 0; e < v.nv; ++e) {
          const uint32_t q = __shfl_sync(kFull, v.ven, e);
          const uint32_t a = q >> 16, b = q & 0xffffu;
          h |= (c == a) | (c == b) | (d == a) | (d == b);
        }
        const unsigned hb = __ballot_sync(kFull, h && app && lane >= from);
        if (!hb) break;
        const int sl = __ffs(hb) - 1;
        from = (uint32_t)sl + 1;
        ++nheavy;
        const uint32_t M = base + (uint32_t)sl;
        const uint32_t en = __shfl_sync(kFull, ea.x, sl);
        const uint32_t pa = __shfl_sync(kFull, par.x, sl) & 0x7fffffffu, pb = __shfl_sync(kFull, par.y, sl) & 0x7fffffffu;
        const uint32_t cc = en >> 16, dd = en & 0xffffu;
        // substitution (the parents are below M: everything below M is final)
        const bool want = wc_in(v, pa) != wc_in(v, pb);
        bool cur = wc_in(v, M);
        if (want != cur) {
          if (!wc_toggle(v, M, en)) { pos = M; status = WC_BIG; break; }
          cur = want;
        }
        // verification of the row; an owned pivot adds the owner's V and the row is verified again
        for (;;) {
          const int w = wc_verify(v, M, cc, dd, cur, nscan);
          if (w < 0) break;
          const unsigned long long k = (unsigned long long)M * (unsigned long long)n + (unsigned long long)(n - 1 - w);
          // who owns this pivot?  an earlier column (larger index): add its V and verify the row again; else: take it
          int slot = 0, rid = -1;
          if (lane == 0) { slot = tab_slot(k); rid = *(volatile int*)&hvals[slot]; }
          slot = __shfl_sync(kFull, slot, 0);
          rid = __shfl_sync(kFull, rid, 0);
          bool added = false;
          for (;;) {
            const int c = rec_col(rid);
            if (c > b) {
              const uint32_t* r = rec_ptr(rid);
              const uint32_t nvc = __ldcg(&r[4]);
              const WcState saved = v;
              bool fits = true;
              if (nvc == 0xffffffffu) {   // the owner's V is in the pool
                const uint32_t vn = __ldcg(&r[6]);
                const uint32_t* ov = P.vpool + (size_t)p * P.vpool_cap + __ldcg(&r[7]);
                fits = vn <= 2u * kWcMaxV;
                for (uint32_t q = 0; q < vn && fits; ++q) {
                  const uint32_t re = __ldcg(&ov[q]);
                  fits = wc_toggle(v, re, __ldg(&EN[re]));
                }
              } else {
                const uint32_t mine = lane < (int)nvc ? __ldcg(&r[8 + lane]) : 0u;
                for (uint32_t q = 0; q < nvc && fits; ++q) {
                  const uint32_t re = __shfl_sync(kFull, mine, (int)q);
                  fits = wc_toggle(v, re, __ldg(&EN[re]));
                }
              }
              if (!fits) { v = saved; pos = M; status = WC_BIG; }   // the cluster engine meets the same pivot again and adds it itself
              added = true;
              break;
            }
            // take the pivot: the record first (complete before anybody can see it), then the claim
            if (myrec < 0) { if (lane == 0) myrec = rec_alloc(); myrec = __shfl_sync(kFull, myrec, 0); }
            rec_write(myrec, WC_DEATH, v, M, k, b);
            __threadfence();
            int old = 0;
            if (lane == 0) old = atomicCAS(&hvals[slot], rid, myrec);
            old = __shfl_sync(kFull, old, 0);
            if (old == rid) {
              if (lane == 0) { if (rid >= 0) act_push(act_next, rid); __stcg(&cur_rec[b], myrec); }
              key = k; pos = M; status = WC_DEATH;
              break;
            }
            rid = old;   // somebody else was faster: look again
          }
          if (status != WC_NONE) break;
          if (added) cur = wc_in(v, M);
        }
        if (status != WC_NONE) break;
      }
      if (status == WC_NONE) pos = base + 32u;
    }
    if (status != WC_DEATH) {   // essential: final;  big: the state goes to the cluster engine
      if (myrec < 0) { if (lane == 0) myrec = rec_alloc(); myrec = __shfl_sync(kFull, myrec, 0); }
      rec_write(myrec, status, v, pos, 0ull, b);
      __threadfence();
      if (lane == 0) {
        __stcg(&cur_rec[b], myrec);
        if (status == WC_BIG) {
          const uint32_t kb = atomicAdd(&S.nbig, 1u);
          if (kb < (uint32_t)P.s2_nrec) __stcg(&biglist[kb], myrec); else fail(TDA_ERR_CAPACITY);
        }
      }
    }
    if (lane == 0) {
      atomicAdd(&S.wc_rows, (unsigned long long)nrows); atomicAdd(&S.wc_heavy, (unsigned long long)nheavy); atomicAdd(&S.wc_scans, (unsigned long long)nscan);
    }
    return status;
  }
  __device__ __forceinline__ void rec_write(int rid, int status, const WcState& v, uint32_t pos, unsigned long long key, int col) {
    uint32_t* r = rec_ptr(rid);
    if (lane == 0) {
      __stcg(&r[0], (uint32_t)status); __stcg(&r[1], pos); __stcg(&r[2], (uint32_t)key); __stcg(&r[3], (uint32_t)(key >> 32)); __stcg(&r[4], (uint32_t)v.nv);
      __stcg(&r[5], (uint32_t)col);
    }
    if (lane < v.nv) __stcg(&r[8 + lane], v.vr);
  }
  // state of a displaced column.  Returns false if its V does not fit a warp (it is in the pool and longer than 32 edges).
  __device__ __forceinline__ bool rec_load(int rid, WcState& v, uint32_t& pos, int& col, int p) const {
    const uint32_t* r = rec_ptr(rid);
    pos = __ldcg(&r[1]);
    col = (int)__ldcg(&r[5]);
    const uint32_t nvr = __ldcg(&r[4]);
    if (nvr == 0xffffffffu) {   // V in the pool (the column was reduced by the cluster engine)
      const uint32_t vn = __ldcg(&r[6]);
      if (vn > (uint32_t)kWcMaxV) return false;
      const uint32_t* ov = P.vpool + (size_t)p * P.vpool_cap + __ldcg(&r[7]);
      v.nv = (int)vn;
      v.vr = lane < v.nv ? __ldcg(&ov[lane]) : 0xffffffffu;
    } else {
      v.nv = (int)nvr;
      v.vr = lane < v.nv ? __ldcg(&r[8 + lane]) : 0xffffffffu;
    }
    v.ven = lane < v.nv ? __ldg(&EN[v.vr]) : 0u;
    return true;
  }
  // the H1 rows in ripser's order (column index descending), zero-persistence pairs dropped (as ripser does).  CTA 0.
  __device__ __forceinline__ void emit_all(int p, const int* bl, int nb) {
    const float* SD = P.sdist + (size_t)p * P.E;
    float* out = P.h1_pairs + (size_t)p * P.cap1 * 2;
    int64_t* outs = P.h1_simplex ? P.h1_simplex + (size_t)p * P.cap1 * 2 : nullptr;
    __shared__ int s_wsum[kS2Warps];
    __shared__ int s_base;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int k0 = 0; k0 < nb; k0 += kS2Threads) {
      const int k = k0 + tid;
      bool keep = false, essential = false;
      float birth = 0.f, death = INFINITY;
      int rbirth = 0, Md = -1, wd = -1;
      if (k < nb) {
        const int ci = nb - 1 - k;
        rbirth = __ldcg(&bl[ci]);
        const uint32_t* r = rec_ptr(__ldcg(&cur_rec[ci]));
        essential = __ldcg(&r[0]) == (uint32_t)WC_ESSENTIAL;
        birth = SD[rbirth];
        if (!essential) {
          const uint64_t pivot = (uint64_t)__ldcg(&r[2]) | ((uint64_t)__ldcg(&r[3]) << 32);
          Md = (int)(pivot / (uint64_t)n); wd = n - 1 - (int)(pivot % (uint64_t)n); death = SD[Md];
        }
        keep = essential || death > birth;
      }
      const unsigned bal = __ballot_sync(kFull, keep);
      if (lane == 0) s_wsum[warp] = __popc(bal);
      __syncthreads();
      int off = s_base;
      for (int w2 = 0; w2 < warp; ++w2) off += s_wsum[w2];
      if (keep) {
        const int row = off + __popc(bal & ((1u << lane) - 1));
        out[2 * row] = birth; out[2 * row + 1] = death;
        if (outs) {
          const uint32_t e = EN[rbirth];
          outs[2 * row] = edge_index((int)(e >> 16), (int)(e & 0xffffu));
          if (essential) outs[2 * row + 1] = -1;
          else {
            const uint32_t em = EN[Md];
            int x = (int)(em >> 16), y = (int)(em & 0xffffu), z = wd, t;
            if (x < y) { t = x; x = y; y = t; }
            if (y < z) { t = y; y = z; z = t; }
            if (x < y) { t = x; x = y; y = t; }
            outs[2 * row + 1] = (int64_t)x * (x - 1) * (x - 2) / 6 + (int64_t)y * (y - 1) / 2 + z;
          }
        }
      }
      __syncthreads();
      if (tid == 0) { int tot = 0; for (int w2 = 0; w2 < kS2Warps; ++w2) tot += s_wsum[w2]; s_base += tot; }
      __syncthreads();
    }
    if (tid == 0) S.nrows = s_base;
  }

  // =================================================================================================================
  // Cluster engine: column ci with windows (all threads of the cluster).  The column starts from the state the warp engine left:
  // `nv0` edges in the column's record, first row pos0 (nv0 = 1, the birth edge, pos0 = birth + 1 for a fresh column).
  __device__ __forceinline__ void cluster_column(int p, int rid0, long long* cyc, unsigned long long& badd_edges) {
    const uint32_t w0 = (uint32_t)P.s2_w0, wsparse = (uint32_t)P.s2_wsparse, wmax = (uint32_t)P.s2_wmax;
    const int nb = P.bcount[p];
    long long t0;
    const uint32_t* r0 = rec_ptr(rid0);
    const int ci = (int)__ldcg(&r0[5]);
    const uint32_t pos0 = __ldcg(&r0[1]);
    {  // V = the recorded state (in the record, or in the pool for a column the cluster engine reduced before)
      const uint32_t nvr = __ldcg(&r0[4]);
      const bool pooled = nvr == 0xffffffffu;
      const int nv0 = pooled ? (int)__ldcg(&r0[6]) : (int)nvr;
      const uint32_t* ov = pooled ? P.vpool + (size_t)p * P.vpool_cap + __ldcg(&r0[7]) : r0 + 8;
      bool any = false;
      for (int q = gtid; q < nv0; q += nthreads) {
        const uint32_t e = __ldcg(&ov[q]);
        const uint32_t en = __ldg(&EN[e]);
        toggle_edge(e, en >> 16, en & 0xffffu);
        any = true;
      }
      if (any) __threadfence();
      csync();
    }
    {
      bool essential = false, dense = false, keep_hi = false;
      uint64_t pivot = 0;
      uint32_t pos = pos0, win = w0, hi = 0;
      unsigned long long guard = 0;
      const unsigned long long guard_max = 1024ull + 4ull * ((unsigned long long)T / max(1u, w0) + 1ull) + 64ull * (unsigned long long)nb;
      for (;;) {
        if (S.abort_flag) break;
        if (pos >= (uint32_t)T) { essential = true; break; }
        if (++guard > guard_max) { csync(); if (gtid == 0) fail(TDA_ERR_INTERNAL_S2); csync(); break; }
        if (!keep_hi || hi <= pos) {
          hi = (pos + win + 31u) & ~31u;
          if (hi > (uint32_t)T) hi = (uint32_t)T;
        }
        keep_hi = false;
        if (S.vcount + (hi - pos) + 64u > (uint32_t)P.vcap) {
          v_compact();
          if (S.vcount + (hi - pos) + 64u > (uint32_t)P.vcap) { csync(); if (gtid == 0) fail(TDA_ERR_CAPACITY); csync(); break; }
        }
        const uint32_t g0 = pos >> 5, g1 = (hi + 31u) >> 5;
        const uint32_t base_row = g0 << 5;
        const uint32_t vmark = S.vcount;
        csync();   // everybody has read the control block of the previous window
        if (gtid == 0) {
          S.npend[0] = S.npend[1] = S.npend[2] = 0; S.nheavy = 0; S.nfail = 0;
          S.fail_key = 0xffffffffu; S.newtouch = 0;
          S.st[S2_WINDOWS] += 1; S.st[S2_SUBST] += hi - pos;
        }
        for (int i = tid; i < W; i += kS2Threads) tnew[i] = 0;
        csync();
        t0 = clock64();
        // ---- substitution, round 1: every row of the window, one warp per 32 consecutive ranks, kS2Unroll groups in flight
        for (uint32_t gb = g0 + (uint32_t)gwarp * kS2Unroll; gb < g1; gb += (uint32_t)nwarps * kS2Unroll) {
          uint32_t xold[kS2Unroll];
          uint2 ea[kS2Unroll], par[kS2Unroll];
#pragma unroll
          for (int u = 0; u < kS2Unroll; ++u) {
            const uint32_t g = gb + u;
            const uint32_t row = (g << 5) + lane;
            const bool inwin = g < g1 && row >= pos && row < hi;
            xold[u] = g < g1 ? __ldcg(&vbits[g]) : 0u;
            ea[u] = make_uint2(0u, 0xffffffffu);
            par[u] = make_uint2(0u, 0u);
            if (inwin) { ea[u] = __ldg(&EA[row]); par[u] = __ldg(&PAR[row]); }
          }
          uint32_t base[kS2Unroll];
#pragma unroll
          for (int u = 0; u < kS2Unroll; ++u) {
            base[u] = 0;
            if ((int)ea[u].y >= 0) {
              const uint32_t pa = par[u].x & 0x7fffffffu, pb = par[u].y & 0x7fffffffu;
              const bool depA = (par[u].x >> 31) && pa >= pos, depB = (par[u].y >> 31) && pb >= pos;
              const uint32_t c = ea[u].x >> 16, d = ea[u].x & 0xffffu, a = ea[u].y;
              const bool tc = tbit(c), td = tbit(d), ta = tbit(a);
              // a final parent edge with x = 1 has both endpoints touched (touched = as of the start of the window)
              uint32_t xa = 0, xb = 0;
              if (!depA && ta && tc) xa = __ldcg(&vbits[pa >> 5]) >> (pa & 31);
              if (!depB && ta && td) xb = __ldcg(&vbits[pb >> 5]) >> (pb & 31);
              base[u] = (xa ^ xb) & 1u;
            }
          }
          // the window bitmaps and the two lists live in CTA 0: the list cursors are remote atomics for the other CTAs (~700 cycles
          // each), so the entries of the kS2Unroll groups of a trip are appended with ONE atomic per list
          uint2 pent[kS2Unroll];
          unsigned bpend[kS2Unroll], bheavy[kS2Unroll];
          uint32_t np_trip = 0, nh_trip = 0;
#pragma unroll
          for (int u = 0; u < kS2Unroll; ++u) {
            const uint32_t g = gb + u;
            bpend[u] = 0u; bheavy[u] = 0u;
            pent[u] = make_uint2(0u, 0u);
            if (g >= g1) continue;   // (warp-uniform)
            const uint32_t row = (g << 5) + lane;
            const bool app = (int)ea[u].y >= 0;   // (rows outside the window carry apex = -1)
            const uint32_t pa = par[u].x & 0x7fffffffu, pb = par[u].y & 0x7fffffffu;
            const bool depA = app && (par[u].x >> 31) && pa >= pos;   // the parent is an apparent row of this window: not final yet
            const bool depB = app && (par[u].y >> 31) && pb >= pos;
            const uint32_t c = ea[u].x >> 16, d = ea[u].x & 0xffffu;
            const bool heavy = app && (tbit(c) || tbit(d));
            const bool computed = app && !depA && !depB;
            const unsigned bapp = __ballot_sync(kFull, app);
            const unsigned bone = __ballot_sync(kFull, computed && base[u]);
            const unsigned bdone = __ballot_sync(kFull, computed);
            if (lane == 0) { xo[g - g0] = xold[u]; xs[g - g0] = (xold[u] & ~bapp) | bone; done[g - g0] = bdone; }
            const bool pend = app && (depA || depB);
            pent[u] = make_uint2((row - base_row) | (base[u] << 31) | ((depA ? 1u : 0u) << 30) | ((depB ? 1u : 0u) << 29),
                                 (depA ? (pa - base_row) : 0u) | ((depB ? (pb - base_row) : 0u) << 16));
            bpend[u] = __ballot_sync(kFull, pend);
            bheavy[u] = __ballot_sync(kFull, heavy);
            np_trip += (uint32_t)__popc(bpend[u]);
            nh_trip += (uint32_t)__popc(bheavy[u]);
          }
          if (np_trip | nh_trip) {
            uint32_t pbase = 0, hbase = 0;
            if (lane == 0) {
              if (np_trip) pbase = atomicAdd(&S.npend[0], np_trip);
              if (nh_trip) hbase = atomicAdd(&S.nheavy, nh_trip);
            }
            pbase = __shfl_sync(kFull, pbase, 0);
            hbase = __shfl_sync(kFull, hbase, 0);
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int u = 0; u < kS2Unroll; ++u) {
              if ((bpend[u] >> lane) & 1u) *pend_ref(0, pbase + (uint32_t)__popc(bpend[u] & lt)) = pent[u];
              if ((bheavy[u] >> lane) & 1u) *heavy_ref(hbase + (uint32_t)__popc(bheavy[u] & lt)) = make_uint2(((gb + u) << 5) + lane, ea[u].x);
              pbase += (uint32_t)__popc(bpend[u]);
              hbase += (uint32_t)__popc(bheavy[u]);
            }
          }
        }
        csync();
        cyc[2] += clock64() - t0;
        t0 = clock64();
        // ---- later rounds (CTA 0 alone): rows that wait for rows of the window; shared memory only (+ the spill part of the list)
        if (crank == 0 && S.npend[0] != 0) {
          int r = 0;
          for (;;) {
            const uint32_t np = S.npend[r % 3];
            if (np == 0) break;
            if (r >= kS2MaxRounds) { if (tid == 0) fail(TDA_ERR_INTERNAL_S2); break; }
            if (tid == 0) { S.npend[(r + 2) % 3] = 0; S.st[S2_ROUNDS] += 1; S.st[S2_LATE] += np; }
            const int rb = r & 1, wb = rb ^ 1;
            const uint32_t np_pad = (np + 31u) & ~31u;
            for (uint32_t i = tid; i < np_pad; i += kS2Threads) {
              const bool have = i < np;
              uint2 ent = make_uint2(0u, 0u);
              if (have) ent = *pend_ref(rb, i);
              const uint32_t rl = ent.x & 0x00ffffffu, base = ent.x >> 31;
              const bool hasA = (ent.x >> 30) & 1u, hasB = (ent.x >> 29) & 1u;
              const uint32_t pal = ent.y & 0xffffu, pbl = ent.y >> 16;
              const volatile uint32_t* vdone = done;
              const volatile uint32_t* vxs = xs;
              const bool okA = !hasA || ((vdone[pal >> 5] >> (pal & 31)) & 1u);
              const bool okB = !hasB || ((vdone[pbl >> 5] >> (pbl & 31)) & 1u);
              const bool ready = have && okA && okB;
              if (ready) {
                uint32_t v = base;
                if (hasA) v ^= (vxs[pal >> 5] >> (pal & 31)) & 1u;
                if (hasB) v ^= (vxs[pbl >> 5] >> (pbl & 31)) & 1u;
                if (v) atomicOr(&xs[rl >> 5], 1u << (rl & 31));
                __threadfence_block();
                atomicOr(&done[rl >> 5], 1u << (rl & 31));
              }
              warp_append(have && !ready, ent, &S.npend[(r + 1) % 3], [&](uint32_t j) { return pend_ref(wb, j); });
            }
            __syncthreads();
            if (S.npend[(r + 1) % 3] == np) { if (tid == 0) fail(TDA_ERR_INTERNAL_S2); __syncthreads(); break; }   // no progress: broken order
            ++r;
          }
        }
        csync();
        if (S.abort_flag) break;
        cyc[2] += clock64() - t0;
        t0 = clock64();
        // ---- apply: the rows whose x changed flip in X / the V list; the global x words are rewritten (one owner per word)
        {
          bool flipped = false;
          for (uint32_t g = g0 + (uint32_t)gwarp; g < g1; g += (uint32_t)nwarps) {
            const uint32_t xnew = xs[g - g0];
            const uint32_t diff = xo[g - g0] ^ xnew;
            if (diff) {
              uint32_t vb = 0;
              if (lane == 0) { __stcg(&vbits[g], xnew); vb = atomicAdd(&S.vcount, (uint32_t)__popc(diff)); }
              vb = __shfl_sync(kFull, vb, 0);
              if ((diff >> lane) & 1u) {
                const uint32_t row = (g << 5) + lane;
                const uint32_t en = __ldg(&EA[row]).x;
                const uint32_t c = en >> 16, d = en & 0xffffu;
                x_flip(c, d);
                touch(c); touch(d);
                const uint32_t vp = vb + __popc(diff & ((1u << lane) - 1));
                if (vp < (uint32_t)P.vcap) __stcg(&vlist(S.vsel)[vp], row);
                else fail(TDA_ERR_CAPACITY);
              }
              flipped = true;
            }
          }
          if (flipped) __threadfence();   // the flips are REDs: performed before anybody reads X after the barrier
        }
        csync();
        if (S.abort_flag) break;
        if (gtid == 0) { S.st[S2_FLIPS] += S.vcount - vmark; S.st[S2_UNDONE] += S.nundone; S.nundone = 0; }
        // rows that became heavy through a vertex touched in this window
        if (S.newtouch) {
          for (uint32_t g = g0 + (uint32_t)gwarp; g < g1; g += (uint32_t)nwarps) {
            const uint32_t row = (g << 5) + lane;
            const bool inwin = row >= pos && row < hi;
            uint2 ea = make_uint2(0u, 0xffffffffu);
            if (inwin) ea = __ldg(&EA[row]);
            const uint32_t c = ea.x >> 16, d = ea.x & 0xffffu;
            bool h = false;
            if (inwin && (int)ea.y >= 0) {
              const bool oc = tbit(c) && !tnewbit(c), od = tbit(d) && !tnewbit(d);
              h = !(oc || od) && (tbit(c) || tbit(d));
            }
            warp_append(h, make_uint2(row, ea.x), &S.nheavy, [&](uint32_t i) { return heavy_ref(i); });
          }
          csync();
        }
        cyc[2] += clock64() - t0;
        const uint32_t nh = S.nheavy;
        if (!dense && nh >= max((uint32_t)P.s2_dense_min, (hi - pos) / (uint32_t)P.s2_dense_div)) {
          dense = true;
          if (gtid == 0) S.st[S2_DENSE] += 1;
        }
        if (gtid == 0) S.st[S2_HEAVY] += nh;
        t0 = clock64();
        if (dense) p_move(hi);
        cyc[5] += clock64() - t0;
        t0 = clock64();
        // ---- verification of the heavy rows, kS2Batch rows per warp in flight (see the header)
        for (uint32_t i0 = (uint32_t)gwarp * kS2Batch; i0 < nh; i0 += (uint32_t)nwarps * kS2Batch) {
          uint32_t Mr[kS2Batch], cd[kS2Batch], xm[kS2Batch];
#pragma unroll
          for (int b = 0; b < kS2Batch; ++b) {
            const uint32_t i = i0 + b;
            uint2 ent = make_uint2(0xffffffffu, 0u);
            if (i < nh) ent = *heavy_ref(i);
            Mr[b] = ent.x; cd[b] = ent.y;
          }
#pragma unroll
          for (int b = 0; b < kS2Batch; ++b) {
            const uint32_t rl = Mr[b] - base_row;
            xm[b] = (Mr[b] != 0xffffffffu && ((xs[rl >> 5] >> (rl & 31)) & 1u)) ? 0xffffffffu : 0u;
          }
          for (int k0 = 0; k0 < W; k0 += 64) {   // W is even, rows are 8-byte aligned; the loop is warp-uniform
            const int k = k0 + 2 * lane;
            const bool kin = k < W;
            uint2 cw[kS2Batch];
#pragma unroll
            for (int b = 0; b < kS2Batch; ++b) {
              const uint32_t c = cd[b] >> 16, d = cd[b] & 0xffffu;
              cw[b] = make_uint2(0u, 0u);
              if (kin) {
                const uint2 xc = __ldcg(reinterpret_cast<const uint2*>(X + (size_t)c * W + k));
                const uint2 xd = __ldcg(reinterpret_cast<const uint2*>(X + (size_t)d * W + k));
                cw[b] = make_uint2(xm[b] ^ xc.x ^ xd.x, xm[b] ^ xc.y ^ xd.y);
                if (dense) {
                  const uint2 pc = __ldcg(reinterpret_cast<const uint2*>(Pm + (size_t)c * W + k));
                  const uint2 pd = __ldcg(reinterpret_cast<const uint2*>(Pm + (size_t)d * W + k));
                  cw[b].x &= pc.x & pd.x; cw[b].y &= pc.y & pd.y;
                }
              }
            }
            // which rows of the batch have candidates, and few enough to probe them one by one: all their probes go out together
            int best[kS2Batch];
            bool big[kS2Batch];
            // a later row than the best failure found so far cannot be the event.  The control block lives in CTA 0 (a remote read
            // for the other CTAs: ~700 cycles), so it is read once per batch, by one lane (the test must be warp-uniform), and only
            // if some row of the batch has candidate bits at all
            bool anyc = false;
#pragma unroll
            for (int b = 0; b < kS2Batch; ++b) anyc |= Mr[b] != 0xffffffffu && (cw[b].x | cw[b].y) != 0;
            uint32_t fk_now = 0xffffffffu;
            if (__any_sync(kFull, anyc)) {
              if (lane == 0) { S.nfail = 1; fk_now = *(volatile uint32_t*)&S.fail_key; }
              fk_now = __shfl_sync(kFull, fk_now, 0);
            }
#pragma unroll
            for (int b = 0; b < kS2Batch; ++b) {
              best[b] = -1;
              big[b] = false;
              if (Mr[b] == 0xffffffffu) { cw[b] = make_uint2(0u, 0u); continue; }
              if (!__any_sync(kFull, (cw[b].x | cw[b].y) != 0)) continue;
              if (fk_now != 0xffffffffu && (Mr[b] - base_row) > fk_now / (uint32_t)n) { cw[b] = make_uint2(0u, 0u); continue; }
              const int npop = __reduce_add_sync(kFull, (unsigned)(__popc(cw[b].x) + __popc(cw[b].y)));
              if (npop > kS2ProbeMax) { big[b] = true; cw[b] = make_uint2(0u, 0u); }
            }
            // probe this lane's candidates from the highest vertex down (the 64 bits of the lane's two words as one number); the
            // first true one is this lane's highest.  The loop runs while any row of the batch has an unprobed candidate here;
            // the probes of all rows of the batch go out together.
            unsigned long long cand[kS2Batch];
#pragma unroll
            for (int b = 0; b < kS2Batch; ++b) cand[b] = ((unsigned long long)cw[b].y << 32) | (unsigned long long)cw[b].x;
            for (;;) {
              int w[kS2Batch];
              bool any = false;
#pragma unroll
              for (int b = 0; b < kS2Batch; ++b) {
                w[b] = -1;
                if (cand[b]) {
                  const int bit = 63 - __clzll((long long)cand[b]);
                  cand[b] &= ~(1ull << bit);
                  w[b] = k * 32 + bit;
                  any = true;
                }
              }
              if (!any) break;
              int ra[kS2Batch], rb[kS2Batch];
#pragma unroll
              for (int b = 0; b < kS2Batch; ++b) {
                ra[b] = rb[b] = kRankDiag;
                if (w[b] >= 0 && w[b] < n) {
                  ra[b] = __ldg(&R[(size_t)(cd[b] >> 16) * n + w[b]]);
                  rb[b] = __ldg(&R[(size_t)(cd[b] & 0xffffu) * n + w[b]]);
                }
              }
#pragma unroll
              for (int b = 0; b < kS2Batch; ++b)
                if (w[b] >= 0 && ra[b] < (int)Mr[b] && rb[b] < (int)Mr[b]) {
                  best[b] = w[b];   // candidates are taken in descending order: the first true one is the highest
                  cand[b] = 0;
                }
            }
#pragma unroll
            for (int b = 0; b < kS2Batch; ++b) {
              if (Mr[b] == 0xffffffffu) continue;
              int bb = __reduce_max_sync(kFull, best[b]);
              if (big[b]) {   // (dense candidate set: the whole lune at once, coalesced)
                bb = exact_row(Mr[b], cd[b] >> 16, cd[b] & 0xffffu, xm[b]);
                if (lane == 0) atomicAdd(&S.st[S2_EXACT], 1ull);
              }
              if (bb >= 0 && lane == 0) atomicMin(&S.fail_key, (Mr[b] - base_row) * (uint32_t)n + (uint32_t)(n - 1 - bb));
            }
          }
        }
        csync();
        cyc[3] += clock64() - t0;
        t0 = clock64();
        // ---- decision
        uint32_t evM = 0xffffffffu;
        int evw = -1;
        {
          const uint32_t fk = S.fail_key;
          if (fk != 0xffffffffu) { evM = base_row + fk / (uint32_t)n; evw = n - 1 - (int)(fk % (uint32_t)n); }
          if (gtid == 0 && S.nfail && fk == 0xffffffffu) S.st[S2_SPURIOUS] += 1;   // candidate bits, none of them true: a clean window
        }
        if (evM == 0xffffffffu) {   // clean window
          pos = hi;
          win = min(2u * win, dense ? wmax : wsparse);
          cyc[4] += clock64() - t0;
          continue;
        }
        // ---- event at (evM, evw): undo the flips above the row
        {
          const uint32_t vend = min(S.vcount, (uint32_t)P.vcap);
          const uint32_t* vl = vlist(S.vsel);
          uint32_t und = 0;
          for (uint32_t f = vmark + (uint32_t)gtid; f < vend; f += (uint32_t)nthreads) {
            const uint32_t e = __ldcg(&vl[f]);
            if (e > evM) {
              const uint32_t en = __ldg(&EN[e]);
              atomicXor(&vbits[e >> 5], 1u << (e & 31));
              x_flip(en >> 16, en & 0xffffu);
              ++und;
            }
          }
          if (und) { atomicAdd(&S.nundone, und); __threadfence(); }
        }
        csync();
        const uint64_t fkey = (uint64_t)evM * (uint64_t)n + (uint64_t)(n - 1 - evw);
        if (gtid == 0) {   // who owns the pivot (no other column is being reduced while the cluster engine runs)
          const int slot = tab_slot(fkey);
          const int rid = *(volatile int*)&hvals[slot];
          S.ev_slot = slot;
          S.ev_rec = rec_col(rid) > ci ? rid : -1;
        }
        csync();
        const int orec = S.ev_rec;
        if (orec < 0) { pivot = fkey; if (gtid == 0) S.st[S2_DEATHS] += 1; cyc[4] += clock64() - t0; break; }   // death (the claim follows)
        {
          const uint32_t* orp = rec_ptr(orec);
          const uint32_t nvc = __ldcg(&orp[4]);
          const bool pooled = nvc == 0xffffffffu;
          const int vn = pooled ? (int)__ldcg(&orp[6]) : (int)nvc;
          const uint32_t* ov = pooled ? P.vpool + (size_t)p * P.vpool_cap + __ldcg(&orp[7]) : orp + 8;
          if (S.vcount + (uint32_t)vn + 64u > (uint32_t)P.vcap) v_compact();
          bool any = false;
          for (int q = gtid; q < vn; q += nthreads) {
            const uint32_t re = __ldcg(&ov[q]);
            const uint32_t en = __ldg(&EN[re]);
            toggle_edge(re, en >> 16, en & 0xffffu);
            any = true;
          }
          badd_edges += vn;
          if (gtid == 0) S.st[S2_EVENTS] += 1;
          if (any) __threadfence();
          csync();
        }
        pos = evM;   // this row again: its substitution is a no-op now, the handled vertex is even, lower vertices may remain
        if (dense) keep_hi = true;
        else win = w0;
        cyc[4] += clock64() - t0;
      }
      if (S.abort_flag) return;
      t0 = clock64();
      // ---- finalise the column: V into the pool, a record of it, the pivot claimed (whoever held it is reduced further later)
      v_compact();
      const uint32_t nv = S.vcount;
      const long long used = S.vpool_used;
      if (!essential) {
        if (used + nv > P.vpool_cap) { csync(); if (gtid == 0) fail(TDA_ERR_CAPACITY); csync(); return; }
        uint32_t* dst = P.vpool + (size_t)p * P.vpool_cap + used;
        const uint32_t* list = vlist(S.vsel);
        for (uint32_t i = (uint32_t)gtid; i < nv; i += (uint32_t)nthreads) __stcg(&dst[i], __ldcg(&list[i]));
        __threadfence();
      }
      csync();
      if (gtid == 0) {
        const int rid = rec_alloc();
        uint32_t* r = rec_ptr(rid);
        __stcg(&r[0], (uint32_t)(essential ? WC_ESSENTIAL : WC_DEATH)); __stcg(&r[1], (uint32_t)(pivot / (uint64_t)n));
        __stcg(&r[2], (uint32_t)pivot); __stcg(&r[3], (uint32_t)(pivot >> 32)); __stcg(&r[4], 0xffffffffu); __stcg(&r[5], (uint32_t)ci);
        __stcg(&r[6], nv); __stcg(&r[7], (uint32_t)used);
        __threadfence();
        if (!essential) {
          S.vpool_used = used + nv;
          const int slot = tab_slot(pivot);
          for (;;) {
            const int cur = *(volatile int*)&hvals[slot];
            if (atomicCAS(&hvals[slot], cur, rid) == cur) { if (cur >= 0) act_push(act_next, cur); break; }
          }
        }
        __stcg(&cur_rec[ci], rid);
        if ((unsigned long long)nv > S.maxv) S.maxv = nv;
      }
      v_clear();
      cyc[5] += clock64() - t0;
    }
  }

  __device__ __forceinline__ void run_problem(int p) {
    R = P.rank + (size_t)p * n * n;
    EN = P.ends + (size_t)p * P.E;
    EA = P.ea + (size_t)p * P.E;
    PAR = P.par + (size_t)p * P.E;
    T = P.T[p];
    hkeys = P.hkeys + (size_t)p * P.hcap;
    hvals = P.hvals + (size_t)p * P.hcap;
    int* bl = P.blist + (size_t)p * P.cap1;
    const int nb = P.bcount[p];
    unsigned long long* st = P.stats + (size_t)p * ST_N;
    if (nb > P.cap1) {
      if (gtid == 0) { P.counts[p * 4 + 3] = TDA_ERR_CAPACITY; P.counts[p * 4 + 1] = 0; }
      return;
    }
    p_valid = false;   // Pm belongs to the previous cloud's rank matrix
    sort_blist(bl, nb);
    for (int i = gtid; i < P.hcap; i += nthreads) { __stcg(&hkeys[i], kEmpty); __stcg(&hvals[i], -1); }
    for (int i = tid; i < W; i += kS2Threads) { touched[i] = 0; tnew[i] = 0; }
    if (gtid == 0) {
      S.vcount = 0; S.vsel = 0; S.abort_flag = 0; S.nundone = 0;
      S.nrows = 0; S.vpool_used = 0; S.maxv = 0; S.badd = 0; S.next_col = 0;
      S.nrec = 0; S.nact[0] = S.nact[1] = 0; S.nbig = 0; S.nbig_done = 0; S.ev_rec = -1; S.ev_slot = 0;
      S.wc_cols = S.wc_resumed = S.wc_big = S.wc_rows = S.wc_heavy = S.wc_scans = 0;
      for (int q = 0; q < 16; ++q) S.st[q] = 0;
    }
    __threadfence();
    csync();
    unsigned long long badd_edges = 0;
    long long cyc[6] = {0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    cyc_sync = 0; n_sync = 0;
    const bool use_warp_engine = P.s2_warp_engine != 0;

    // ---- rounds.  Round 0 reduces every column from its birth edge; later rounds reduce the displaced columns further from
    // their records; every column is one warp's work (dynamic list).  Columns that outgrow a warp wait for the cluster engine,
    // which takes them one by one between the rounds; its claims may displace columns again.  Fixpoint: both lists empty.
    if (!use_warp_engine) {   // every column goes to the cluster engine, from its birth edge (column order = ripser's)
      for (int k = gtid; k < nb; k += nthreads) {
        const int ci = nb - 1 - k;
        uint32_t* r = rec_ptr(k);
        const uint32_t rbirth = (uint32_t)__ldcg(&bl[ci]);
        __stcg(&r[0], (uint32_t)WC_BIG); __stcg(&r[1], rbirth + 1u); __stcg(&r[4], 1u); __stcg(&r[5], (uint32_t)ci); __stcg(&r[8], rbirth);
        __stcg(&biglist[k], k);
      }
      if (gtid == 0) { S.nrec = (uint32_t)nb; S.nbig = (uint32_t)nb; }
      __threadfence();
      csync();
    }
    int cur_list = 0;       // work list of this round (round 0 of the warp engine: the nb fresh columns, no list)
    bool fresh = use_warp_engine;
    for (int round = 0;; ++round) {
      if (round > 4 * nb + 64) { csync(); if (gtid == 0) fail(TDA_ERR_INTERNAL_S2); csync(); break; }
      act_next = cur_list ^ 1;
      const uint32_t nwork = fresh ? (uint32_t)nb : S.nact[cur_list];
      // (a) warp rounds
      t0 = clock64();
      if (nwork) {
        for (;;) {
          int k = 0;
          if (lane == 0) k = atomicAdd(&S.next_col, 1);
          k = __shfl_sync(kFull, k, 0);
          if (k >= (int)nwork) break;
          WcState v;
          uint32_t pos;
          int b;
          if (fresh) {
            b = nb - 1 - k;   // ripser's order first: fewer displacements
            const uint32_t rbirth = (uint32_t)__ldcg(&bl[b]);
            v.nv = 1; v.vr = lane == 0 ? rbirth : 0xffffffffu; v.ven = lane == 0 ? __ldg(&EN[rbirth]) : 0u;
            pos = rbirth + 1;
          } else {
            const int rid = __ldcg(&act[(size_t)cur_list * P.s2_nrec + k]);
            if (lane == 0) atomicAdd(&S.wc_resumed, 1ull);
            if (!rec_load(rid, v, pos, b, p)) {   // too long for a warp: the cluster engine goes on with it
              if (lane == 0) {
                const uint32_t kb = atomicAdd(&S.nbig, 1u);
                if (kb < (uint32_t)P.s2_nrec) __stcg(&biglist[kb], rid); else fail(TDA_ERR_CAPACITY);
              }
              continue;
            }
          }
          unsigned long long key = 0;
          wc_sweep(v, pos, key, b, p);
          if (lane == 0) atomicAdd(&S.wc_cols, 1ull);
        }
        __threadfence();
      }
      csync();
      if (gtid == 0) { S.next_col = 0; S.nact[cur_list] = 0; }
      csync();
      cyc[fresh ? 0 : 1] += clock64() - t0;
      fresh = false;
      if (S.abort_flag) break;
      // (b) nothing displaced: the cluster engine takes the next waiting column (largest index first would be ideal; the list
      //     is in arrival order, which follows the column order of round 0 closely)
      if (S.nact[cur_list ^ 1] == 0) {
        if (S.nbig_done >= S.nbig) break;   // fixpoint
        const int rid = __ldcg(&biglist[S.nbig_done]);
        csync();
        if (gtid == 0) { S.nbig_done += 1; S.wc_big += 1; }
        cluster_column(p, rid, cyc, badd_edges);
        csync();
        if (S.abort_flag) break;
      }
      cur_list ^= 1;
    }
    csync();
    if (!S.abort_flag && crank == 0) emit_all(p, bl, nb);
    csync();
    if (S.abort_flag) {  // leave the scratch clean for the next problem
      for (size_t i = (size_t)gtid; i < (size_t)n * W; i += (size_t)nthreads) X[i] = 0;
      for (size_t i = (size_t)gtid; i < (size_t)P.vwords; i += (size_t)nthreads) vbits[i] = 0;
      for (int i = tid; i < W; i += kS2Threads) { touched[i] = 0; tnew[i] = 0; }
      __threadfence();
      csync();
      if (gtid == 0) S.vcount = 0;
    }
    if (gtid == 0 && P.s2_debug)
      printf("sweep2 cloud %d: columns %d to_cluster %llu windows %llu heavy %llu exact_rows %llu events %llu deaths %llu late_rows %llu | Mcyc warp %.2f commit %.2f subst %.2f verify %.2f events %.2f pm/final %.2f | barriers %llu (%.2f Mcyc)\n",
             p, nb, S.wc_big, S.st[S2_WINDOWS], S.st[S2_HEAVY], S.st[S2_EXACT], S.st[S2_EVENTS], S.st[S2_DEATHS], S.st[S2_LATE], cyc[0] / 1e6, cyc[1] / 1e6,
             cyc[2] / 1e6, cyc[3] / 1e6, cyc[4] / 1e6, cyc[5] / 1e6, n_sync, cyc_sync / 1e6);
    if (gtid == 0) {
      P.counts[p * 4 + 1] = S.nrows;
      P.counts[p * 4 + 3] = S.abort_flag;
      st[ST_REDUCED] = (unsigned long long)nb;
      st[ST_ADDITIONS] = S.st[S2_FLIPS] - S.st[S2_UNDONE] + S.st[S2_EVENTS];
      st[ST_PUSHES] = S.st[S2_SUBST] + S.wc_rows;        // rows substituted (windows) + rows swept by warps
      st[ST_POPS] = S.st[S2_EVENTS] + S.st[S2_DEATHS];   // non-apparent pivots of the cluster engine
      st[ST_EXTENSIONS] = S.st[S2_WINDOWS];
      st[ST_MAXV] = S.maxv;
      for (int q = 0; q < 6; ++q) st[ST_CYC_EXTRACT + q] = (unsigned long long)cyc[q];
      st[ST_BADD_EDGES] = badd_edges;
      st[ST_EXT_EDGES] = S.st[S2_HEAVY] + S.wc_heavy;
      st[ST_S2_ROUNDS] = S.st[S2_ROUNDS];
      st[ST_S2_LATE] = S.st[S2_LATE];
      st[ST_S2_PM] = S.st[S2_PM];
      st[ST_S2_DENSE] = S.st[S2_DENSE];
      st[ST_S2_SPURIOUS] = S.wc_resumed;     // columns reduced further after they were displaced from their pivot
      st[ST_S2_UNDONE] = S.wc_big;           // columns handed to the cluster engine
      st[ST_SPARE0] = (unsigned long long)cyc_sync;
      st[ST_SPARE1] = n_sync;
    }
    csync();
  }
};

// grid = clusters * cluster size (launch attribute); one cluster per cloud at a time
__global__ void __launch_bounds__(kS2Threads, 1) rips_sweep2_kernel(const __grid_constant__ ReduceParams P) {
  __shared__ Sweep2Smem S;
  extern __shared__ __align__(16) uint32_t sweep2_dyn[];
  Sweeper2 sw(P, S, sweep2_dyn);
  sw.csync();   // every CTA of the cluster is up before anybody touches its shared memory
  for (;;) {
    if (sw.gtid == 0) sw.S.problem = atomicAdd(P.work_counter, 1);
    sw.csync();
    const int p = sw.S.problem;
    sw.csync();
    if (p >= P.batch) break;
    sw.run_problem(p);
  }
  sw.csync();   // no CTA leaves while another may still reach into its shared memory
}

// parents of the apparent edges: par[M] = (rank(c,apex) | apparent?<<31, rank(d,apex) | apparent?<<31)
__global__ void parents_kernel(const int* __restrict__ rank, const uint2* __restrict__ ea, const int* __restrict__ Tarr, int n, int64_t E,
                               uint2* __restrict__ par) {
  const int p = blockIdx.y;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Tarr[p]) return;
  const uint2 e = ea[(size_t)p * E + r];
  const int a = (int)e.y;
  uint2 o = make_uint2(0u, 0u);
  if (a >= 0) {
    const int* Rp = rank + (size_t)p * n * n;
    const uint32_t pa = (uint32_t)Rp[(size_t)(e.x >> 16) * n + a], pb = (uint32_t)Rp[(size_t)(e.x & 0xffffu) * n + a];
    const uint32_t fa = (int)ea[(size_t)p * E + pa].y >= 0 ? 1u : 0u, fb = (int)ea[(size_t)p * E + pb].y >= 0 ? 1u : 0u;
    o = make_uint2(pa | (fa << 31), pb | (fb << 31));
  }
  par[(size_t)p * E + r] = o;
}
