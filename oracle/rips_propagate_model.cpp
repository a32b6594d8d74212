// oracle/rips_propagate_model.cpp -- TEST INFRASTRUCTURE ONLY: a CPU model of a "propagate, then verify" formulation of the
// residual H1 reduction, written to check the idea against the ripser restatement (rips_oracle.cpp) before it becomes a
// kernel.  Never on the product path.
//
// Setting (as in csrc/rips.cu): edges in ripser's filtration order get ranks 0..T-1; a non-MST edge M=(c,d) whose lune
// {w : rank(c,w) < M, rank(d,w) < M} is not empty is in an apparent pair with the triangle (M, apex(M)), apex = the largest
// lune vertex; the other non-MST edges are the residual columns, reduced in descending rank order.  Triangle keys
// (M, w) ascend with M and, inside a row M, with descending w.
//
// Observation the model rests on: when the reduction of a column reaches row M of an apparent edge, the entry at the apex
// decides whether column M is added, and after that step x_M = x_(c,apex) ^ x_(d,apex) whatever x_M was before (x_e = 1 iff
// edge e is in the reduction column V).  So the tens of thousands of apparent-pair additions of a long column are a forward
// substitution along a STATIC graph (two parents per apparent edge) whose depth is ~30 on the C3 clouds, not a chain of
// dependent pivots.  The model therefore repeats, per column:
//   propagate  x_M = x_pa(M) ^ x_pb(M) for every apparent edge above the cursor (speculative view),
//   verify     the rows above the cursor in order: r(M) = { w in lune(M) : x_M ^ x_(c,w) ^ x_(d,w) = 1 }; the first non-empty
//              row holds the next NON-apparent pivot (M, max r(M)),
//   event      pivot owned by an earlier column j -> V ^= V_j, cursor = M, again;  unowned -> death;  no row -> essential.
// Everything below the cursor is final (the coboundary of V_j vanishes below its pivot).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <unordered_map>
#include <vector>

namespace {
struct Model {
  int n; int64_t T;
  std::vector<int> R;                 // [n*n] rank, INT_MAX above the threshold / on the diagonal
  std::vector<int> ea, eb;            // endpoints by rank (ea > eb)
  std::vector<float> len;             // length by rank
  std::vector<int> apex, pa, pb;      // per rank: apex (-1 none, -2 MST), parents of an apparent edge
};
}  // namespace

namespace {
// edges in filtration order, ranks, MST flags, apex and the two parent edges of every apparent edge
static void build_model(Model& m, const float* dist, int n, std::vector<int64_t>& residual, int64_t& n_app, int& maxdepth) {
  m.n = n;
  const int64_t E = (int64_t)n * (n - 1) / 2;
  std::vector<float> elen(E); std::vector<int> ia(E), ib(E);
  { int64_t q = 0; for (int i = 1; i < n; ++i) for (int j = 0; j < i; ++j) { elen[q] = dist[(size_t)i * n + j]; ia[q] = i; ib[q] = j; ++q; } }
  float thr = INFINITY;
  for (int i = 0; i < n; ++i) { float mx = 0.f; for (int j = 0; j < n; ++j) if (j != i) mx = std::max(mx, dist[(size_t)i * n + j]); thr = std::min(thr, mx); }
  std::vector<int64_t> ord(E); std::iota(ord.begin(), ord.end(), 0);
  std::sort(ord.begin(), ord.end(), [&](int64_t x, int64_t y) { return elen[x] < elen[y] || (elen[x] == elen[y] && x > y); });
  int64_t T = 0; while (T < E && elen[ord[T]] <= thr) ++T;
  m.T = T;
  m.R.assign((size_t)n * n, 0x7fffffff);
  m.ea.resize(T); m.eb.resize(T); m.len.resize(T);
  for (int64_t r = 0; r < T; ++r) { const int a = ia[ord[r]], b = ib[ord[r]]; m.ea[r] = a; m.eb[r] = b; m.len[r] = elen[ord[r]]; m.R[(size_t)a * n + b] = m.R[(size_t)b * n + a] = (int)r; }
  std::vector<int> uf(n); std::iota(uf.begin(), uf.end(), 0);
  auto find = [&](int x) { while (uf[x] != x) { uf[x] = uf[uf[x]]; x = uf[x]; } return x; };
  m.apex.assign(T, -1); m.pa.assign(T, -1); m.pb.assign(T, -1);
  std::vector<int> depth(T, 0); maxdepth = 0; n_app = 0;
  for (int64_t r = 0; r < T; ++r) {
    const int a = m.ea[r], b = m.eb[r];
    const int ra = find(a), rb = find(b);
    if (ra != rb) { uf[ra] = rb; m.apex[r] = -2; continue; }
    const int* Ra = &m.R[(size_t)a * n]; const int* Rb = &m.R[(size_t)b * n];
    int ap = -1; for (int w = n - 1; w >= 0; --w) if (Ra[w] < r && Rb[w] < r) { ap = w; break; }
    m.apex[r] = ap;
    if (ap < 0) { residual.push_back(r); continue; }
    ++n_app; m.pa[r] = Ra[ap]; m.pb[r] = Rb[ap];
    depth[r] = 1 + std::max(depth[m.pa[r]], depth[m.pb[r]]); maxdepth = std::max(maxdepth, depth[r]);
  }
}
}  // namespace

extern "C" {

// dist: n*n float32, symmetric, zero diagonal.  pairs_out: [cap,2] doubles (birth, death; death = inf for essential classes),
// in processing order, zero-persistence pairs dropped (as ripser does).  stats: [10] int64: residual columns, apparent edges,
// events (additions of reduced columns), propagated flips (every pass recomputes the view from the cursor), heavy rows verified,
// largest |V|, propagate passes, depth of the apparent graph, flips an INCREMENTAL propagation would do (edges whose view
// value differs from the previous pass of the same column), rows verified whose own edge is in V.
// Returns the number of pairs (or -1 if cap is too small).
int64_t rips_model_h1(const float* dist, int n, double* pairs_out, int64_t cap, int64_t* stats) {
  Model m;
  std::vector<int64_t> residual; int64_t n_app = 0; int maxdepth = 0;
  build_model(m, dist, n, residual, n_app, maxdepth);
  const int64_t T = m.T;
  const int W = (n + 63) / 64;
  std::vector<uint8_t> xr(T, 0), xs(T, 0);
  std::vector<uint64_t> X((size_t)n * W, 0), lune(W);
  std::vector<uint8_t> touched(n, 0);
  std::unordered_map<int64_t, int> owner;              // pivot key -> index into Vs
  std::vector<std::vector<int>> Vs;
  int64_t n_pairs = 0, events = 0, flips = 0, heavy = 0, maxv = 0, passes = 0, delta_flips = 0, heavy_in_v = 0;
  std::vector<uint8_t> xprev(T, 0);
  std::vector<int64_t> prev_list;
  auto toggle_view = [&](int64_t e) {                  // xs[e] ^= 1 with the vertex bit matrix kept in step
    xs[e] ^= 1; const int a = m.ea[e], b = m.eb[e];
    X[(size_t)a * W + (b >> 6)] ^= 1ull << (b & 63); X[(size_t)b * W + (a >> 6)] ^= 1ull << (a & 63);
    touched[a] = touched[b] = 1;
  };
  for (int64_t ci = (int64_t)residual.size() - 1; ci >= 0; --ci) {
    const int64_t b = residual[ci];
    std::vector<int64_t> real_set;                     // edges with xr = 1 (for the clean-up)
    auto toggle_real = [&](int64_t e) { xr[e] ^= 1; if (xr[e]) real_set.push_back(e); };
    toggle_real(b);
    int64_t M0 = b; int w_last = n;                    // rows <= M0 settled; in row M0 only vertices below w_last remain
    std::vector<int64_t> view_set;                     // edges with xs = 1 at some point (for the clean-up)
    bool essential = false; int64_t pivM = -1; int pivw = -1;
    for (;;) {
      // ---- propagate: the speculative view above the cursor
      for (int64_t e : view_set) if (xs[e]) toggle_view(e);
      view_set.clear();
      std::fill(touched.begin(), touched.end(), 0);
      ++passes;
      for (int64_t e : real_set) if (xr[e] && !xs[e]) { toggle_view(e); view_set.push_back(e); }
      for (int64_t M = M0 + 1; M < T; ++M) {
        if (m.apex[M] < 0) continue;
        const uint8_t want = xs[m.pa[M]] ^ xs[m.pb[M]];
        if (want != xs[M]) { toggle_view(M); view_set.push_back(M); ++flips; }
      }
      {  // what an incremental propagation would have touched: the symmetric difference with the previous pass
        for (int64_t e : view_set) if (xs[e] != xprev[e]) { ++delta_flips; xprev[e] = xs[e]; }
        for (int64_t e : prev_list) if (xs[e] != xprev[e]) { ++delta_flips; xprev[e] = xs[e]; }
        prev_list = view_set;
      }
      // ---- verify rows in order
      pivM = -1;
      for (int64_t M = M0; M < T && pivM < 0; ++M) {
        if (m.apex[M] == -2) continue;                 // MST edge: empty lune
        const int c = m.ea[M], d = m.eb[M];
        if (!touched[c] && !touched[d]) continue;
        ++heavy;
        heavy_in_v += xs[M];
        const int* Rc = &m.R[(size_t)c * n]; const int* Rd = &m.R[(size_t)d * n];
        const int wtop = M == M0 ? w_last - 1 : n - 1;
        for (int w = wtop; w >= 0; --w) {
          if (!(Rc[w] < M && Rd[w] < M)) continue;
          const int bit = (int)xs[M] ^ (int)((X[(size_t)c * W + (w >> 6)] >> (w & 63)) & 1) ^ (int)((X[(size_t)d * W + (w >> 6)] >> (w & 63)) & 1);
          if (bit) { pivM = M; pivw = w; break; }
        }
      }
      if (pivM < 0) { essential = true; break; }
      // rows up to pivM are settled: the view becomes real there
      for (int64_t e : view_set) if (e <= pivM && xs[e] != xr[e]) toggle_real(e);
      const int64_t key = pivM * (int64_t)n + (n - 1 - pivw);
      auto it = owner.find(key);
      if (it == owner.end()) break;                    // death
      ++events;
      for (int e : Vs[it->second]) toggle_real(e);
      M0 = pivM; w_last = pivw;
    }
    // V of this column = the real set
    std::vector<int> V;
    for (int64_t e : real_set) if (xr[e]) { V.push_back((int)e); xr[e] = 0; }
    std::sort(V.begin(), V.end()); V.erase(std::unique(V.begin(), V.end()), V.end());
    maxv = std::max<int64_t>(maxv, (int64_t)V.size());
    for (int64_t e : view_set) if (xs[e]) toggle_view(e);
    for (int64_t e : real_set) if (xs[e]) toggle_view(e);
    for (int64_t e : prev_list) xprev[e] = 0;
    for (int64_t e : view_set) xprev[e] = 0;
    prev_list.clear();
    const float birth = m.len[b];
    if (essential) {
      if (n_pairs >= cap) return -1;
      pairs_out[2 * n_pairs] = birth; pairs_out[2 * n_pairs + 1] = INFINITY; ++n_pairs;
    } else {
      const float death = m.len[pivM];
      owner.emplace(pivM * (int64_t)n + (n - 1 - pivw), (int)Vs.size());
      Vs.push_back(std::move(V));
      if (death > birth) {
        if (n_pairs >= cap) return -1;
        pairs_out[2 * n_pairs] = birth; pairs_out[2 * n_pairs + 1] = death; ++n_pairs;
      }
    }
  }
  if (stats) {
    stats[0] = (int64_t)residual.size(); stats[1] = n_app; stats[2] = events; stats[3] = flips; stats[4] = heavy; stats[5] = maxv;
    stats[6] = passes; stats[7] = maxdepth; stats[8] = delta_flips; stats[9] = heavy_in_v;
  }
  return n_pairs;
}


// The same reduction organised the way the sweep kernel would run it: the rows above the cursor are
// taken in windows of `window` consecutive ranks.  Per window: (A) substitution x_M = x_pa ^ x_pb for the apparent rows whose
// endpoints V touches, flips applied to X at once; (B) every row of the window verified against that X -- a flip of a LATER row
// cannot show up in an earlier row, its edge is outside that row's lune; (C) at the first non-empty row the flips above it are
// undone and the event is handled (owned pivot: V ^= V_j, resume in that row below the pivot vertex; unowned: death).
// stats: [8] residual columns, apparent edges, events, flips, flips undone, heavy rows verified, largest |V|, windows visited.
int64_t rips_model_h1_windowed(const float* dist, int n, int window, double* pairs_out, int64_t cap, int64_t* stats) {
  Model m;
  std::vector<int64_t> residual; int64_t n_app = 0; int maxdepth = 0;
  build_model(m, dist, n, residual, n_app, maxdepth);
  const int64_t T = m.T;
  const int W = (n + 63) / 64;
  if (window < 1) window = 1;
  std::vector<uint8_t> x(T, 0);
  std::vector<uint64_t> X((size_t)n * W, 0);
  std::vector<uint8_t> touched(n, 0);
  std::unordered_map<int64_t, int> owner;
  std::vector<std::vector<int>> Vs;
  int64_t n_pairs = 0, events = 0, flips = 0, undone = 0, heavy = 0, maxv = 0, windows = 0;
  std::vector<int64_t> members;                        // every edge whose x was ever set in this column (clean-up, V)
  auto toggle = [&](int64_t e) {
    x[e] ^= 1; const int a = m.ea[e], b = m.eb[e];
    X[(size_t)a * W + (b >> 6)] ^= 1ull << (b & 63); X[(size_t)b * W + (a >> 6)] ^= 1ull << (a & 63);
    touched[a] = touched[b] = 1;
    if (x[e]) members.push_back(e);
  };
  for (int64_t ci = (int64_t)residual.size() - 1; ci >= 0; --ci) {
    const int64_t b = residual[ci];
    members.clear();
    std::fill(touched.begin(), touched.end(), 0);
    toggle(b);
    int64_t lo = b + 1;          // first row of the next window
    int64_t resume_row = -1; int resume_w = 0;   // after an owned pivot: that row again, below the pivot vertex, no substitution
    bool essential = false; int64_t pivM = -1; int pivw = -1;
    for (;;) {
      const int64_t first = resume_row >= 0 ? resume_row : lo;
      if (first >= T) { essential = true; break; }
      const int64_t hi = std::min<int64_t>(T, first + window);
      ++windows;
      // (A) substitution
      std::vector<int64_t> flipped;
      for (int64_t M = first; M < hi; ++M) {
        if (M == resume_row || m.apex[M] < 0) continue;
        const int c = m.ea[M], d = m.eb[M];
        if (!touched[c] && !touched[d]) continue;
        const uint8_t want = x[m.pa[M]] ^ x[m.pb[M]];
        if (want != x[M]) { toggle(M); flipped.push_back(M); ++flips; }
      }
      // (B) verification against the X that already holds every flip of the window
      pivM = -1;
      for (int64_t M = first; M < hi && pivM < 0; ++M) {
        if (m.apex[M] == -2) continue;
        const int c = m.ea[M], d = m.eb[M];
        if (!touched[c] && !touched[d]) continue;
        ++heavy;
        const int* Rc = &m.R[(size_t)c * n]; const int* Rd = &m.R[(size_t)d * n];
        const int wtop = M == resume_row ? resume_w - 1 : n - 1;
        for (int w = wtop; w >= 0; --w) {
          if (!(Rc[w] < M && Rd[w] < M)) continue;
          const int bit = (int)x[M] ^ (int)((X[(size_t)c * W + (w >> 6)] >> (w & 63)) & 1) ^ (int)((X[(size_t)d * W + (w >> 6)] >> (w & 63)) & 1);
          if (bit) { pivM = M; pivw = w; break; }
        }
      }
      if (pivM < 0) { lo = hi; resume_row = -1; continue; }
      // (C) event: undo the flips above the failing row
      for (int64_t e : flipped) if (e > pivM) { toggle(e); ++undone; }
      const int64_t key = pivM * (int64_t)n + (n - 1 - pivw);
      auto it = owner.find(key);
      if (it == owner.end()) break;                    // death
      ++events;
      for (int e : Vs[it->second]) toggle(e);
      resume_row = pivM; resume_w = pivw; lo = pivM + 1;
    }
    std::vector<int> V;
    for (int64_t e : members) if (x[e]) { V.push_back((int)e); toggle(e); }
    std::sort(V.begin(), V.end()); V.erase(std::unique(V.begin(), V.end()), V.end());
    maxv = std::max<int64_t>(maxv, (int64_t)V.size());
    const float birth = m.len[b];
    if (essential) {
      if (n_pairs >= cap) return -1;
      pairs_out[2 * n_pairs] = birth; pairs_out[2 * n_pairs + 1] = INFINITY; ++n_pairs;
    } else {
      const float death = m.len[pivM];
      owner.emplace(pivM * (int64_t)n + (n - 1 - pivw), (int)Vs.size());
      Vs.push_back(std::move(V));
      if (death > birth) {
        if (n_pairs >= cap) return -1;
        pairs_out[2 * n_pairs] = birth; pairs_out[2 * n_pairs + 1] = death; ++n_pairs;
      }
    }
  }
  if (stats) {
    stats[0] = (int64_t)residual.size(); stats[1] = n_app; stats[2] = events; stats[3] = flips; stats[4] = undone; stats[5] = heavy;
    stats[6] = maxv; stats[7] = windows;
  }
  return n_pairs;
}


// The windowed variant again, but step for step as Sweeper<WPL, VERIFY=true> (csrc/rips.cu) does it, so that the kernel's
// control flow is checked on the CPU too: the chunk's heavy rows are fixed by a filter at the start of a pass; the
// substitution runs in rounds (a row waits until the parent edges that are rows of the same chunk are done; rows that are
// not heavy count as done); a flip that touches a new vertex makes the whole chunk be filtered and substituted again
// (substitution is idempotent); every recorded flip above the failing row is undone, even if a row was recorded twice.
// stats: [8] residual columns, apparent edges, events, flips, flips undone, heavy rows verified, re-filter passes, rounds.
int64_t rips_model_h1_kernel(const float* dist, int n, int chunk, double* pairs_out, int64_t cap, int64_t* stats) {
  Model m;
  std::vector<int64_t> residual; int64_t n_app = 0; int maxdepth = 0;
  build_model(m, dist, n, residual, n_app, maxdepth);
  const int64_t T = m.T;
  const int W = (n + 63) / 64;
  if (chunk < 1) chunk = 1;
  std::vector<uint8_t> x(T, 0);
  std::vector<uint64_t> X((size_t)n * W, 0);
  std::vector<uint8_t> touched(n, 0);
  std::unordered_map<int64_t, int> owner;
  std::vector<std::vector<int>> Vs;
  int64_t n_pairs = 0, events = 0, flips = 0, undone = 0, heavy_rows = 0, refilters = 0, rounds = 0;
  std::vector<int64_t> members;
  auto xbit = [&](int a, int b) { return (int)((X[(size_t)a * W + (b >> 6)] >> (b & 63)) & 1); };
  auto flip = [&](int64_t e) {   // x_flip + v_toggle
    x[e] ^= 1; const int a = m.ea[e], b = m.eb[e];
    X[(size_t)a * W + (b >> 6)] ^= 1ull << (b & 63); X[(size_t)b * W + (a >> 6)] ^= 1ull << (a & 63);
    if (x[e]) members.push_back(e);
  };
  for (int64_t ci = (int64_t)residual.size() - 1; ci >= 0; --ci) {
    const int64_t b = residual[ci];
    members.clear();
    std::fill(touched.begin(), touched.end(), 0);
    flip(b); touched[m.ea[b]] = touched[m.eb[b]] = 1;
    int64_t pos = b + 1, vchunk_pos = -1;
    std::vector<int64_t> fliplist;
    bool essential = false; int64_t pivM = -1; int pivw = -1;
    for (;;) {
      if (pos >= T) { essential = true; break; }
      const int64_t hi = std::min<int64_t>(T, pos + chunk);
      // filter
      std::vector<int64_t> heavy;
      for (int64_t M = pos; M < hi; ++M) if (touched[m.ea[M]] || touched[m.eb[M]]) heavy.push_back(M);
      if (pos != vchunk_pos) { vchunk_pos = pos; fliplist.clear(); }
      // (A) substitution in rounds
      std::vector<uint8_t> done((size_t)(hi - pos), 1), pending(heavy.size(), 0);
      for (size_t h = 0; h < heavy.size(); ++h) if (m.apex[heavy[h]] >= 0) { pending[h] = 1; done[heavy[h] - pos] = 0; }
      bool new_touch = false;
      for (;;) {
        ++rounds;
        std::vector<size_t> can;
        for (size_t h = 0; h < heavy.size(); ++h) {
          if (!pending[h]) continue;
          const int64_t M = heavy[h];
          const int64_t ra = m.pa[M], rb = m.pb[M];
          if (ra >= pos && !done[ra - pos]) continue;
          if (rb >= pos && !done[rb - pos]) continue;
          can.push_back(h);
        }
        for (size_t h : can) {
          const int64_t M = heavy[h];
          const int c = m.ea[M], d = m.eb[M], a = m.apex[M];
          const int want = xbit(c, a) ^ xbit(d, a), cur = xbit(c, d);
          if (want != cur) {
            flip(M); ++flips;
            if (!touched[c] || !touched[d]) new_touch = true;
            touched[c] = touched[d] = 1;
            fliplist.push_back(M);
          }
          pending[h] = 0;
        }
        for (size_t h : can) done[heavy[h] - pos] = 1;
        bool left = false; for (uint8_t q : pending) left |= q != 0;
        if (!left) break;
        if (can.empty()) return -2;   // no progress: the dependency order is broken
      }
      if (new_touch) { ++refilters; continue; }
      // (B) verification of the heavy rows in order
      pivM = -1;
      for (int64_t M : heavy) {
        ++heavy_rows;
        if (m.apex[M] == -2) continue;
        const int c = m.ea[M], d = m.eb[M];
        const int* Rc = &m.R[(size_t)c * n]; const int* Rd = &m.R[(size_t)d * n];
        for (int w = n - 1; w >= 0; --w) {
          if (!(Rc[w] < M && Rd[w] < M)) continue;
          if (x[M] ^ xbit(c, w) ^ xbit(d, w)) { pivM = M; pivw = w; break; }
        }
        if (pivM >= 0) break;
      }
      if (pivM < 0) { pos = hi; continue; }
      // (C) event
      for (int64_t e : fliplist) if (e > pivM) { flip(e); ++undone; }
      const int64_t key = pivM * (int64_t)n + (n - 1 - pivw);
      auto it = owner.find(key);
      if (it == owner.end()) break;                    // death
      ++events;
      for (int e : Vs[it->second]) { flip(e); touched[m.ea[e]] = touched[m.eb[e]] = 1; }
      pos = pivM;
    }
    std::vector<int> V;
    for (int64_t e : members) if (x[e]) { V.push_back((int)e); flip(e); }
    std::sort(V.begin(), V.end()); V.erase(std::unique(V.begin(), V.end()), V.end());
    const float birth = m.len[b];
    if (essential) {
      if (n_pairs >= cap) return -1;
      pairs_out[2 * n_pairs] = birth; pairs_out[2 * n_pairs + 1] = INFINITY; ++n_pairs;
    } else {
      const float death = m.len[pivM];
      owner.emplace(pivM * (int64_t)n + (n - 1 - pivw), (int)Vs.size());
      Vs.push_back(std::move(V));
      if (death > birth) {
        if (n_pairs >= cap) return -1;
        pairs_out[2 * n_pairs] = birth; pairs_out[2 * n_pairs + 1] = death; ++n_pairs;
      }
    }
  }
  if (stats) {
    stats[0] = (int64_t)residual.size(); stats[1] = n_app; stats[2] = events; stats[3] = flips; stats[4] = undone; stats[5] = heavy_rows;
    stats[6] = refilters; stats[7] = rounds;
  }
  return n_pairs;
}


// The formulation the cluster reducer (csrc/rips.cu, ClusterSweeper) implements.  Differences to the kernel-step model above:
//   * windows grow (w0, doubling up to wmax after every clean window, back to w0 after a failure);
//   * the substitution runs over EVERY apparent row of the window (x by rank: x_M = x[pa] ^ x[pb]; no touched filter, no
//     re-filter passes), in rounds: a row waits for a parent that is an apparent row of the same window;
//   * verification uses the SUPERSET mask Pend[c] & Pend[d] (Pend = adjacency of all edges below the window's end) instead of
//     the exact lune.  If the window leaves V a cocycle of the complex at the window's end, every row passes; a failing bit
//     (row M, vertex w) under the superset mask belongs to the triangle {c, d, w} whose own row max(M, rank(c,w), rank(d,w))
//     lies in the window, so the smallest failing row Mf bounds the first true failure from below.  The flips above Mf are
//     undone and row Mf is verified with its exact lune: non-empty -> event at (Mf, highest vertex); empty -> rows <= Mf are
//     settled (spurious stop) and the sweep continues behind it.
// stats: [12] residual columns, apparent edges, events, flips, flips undone, heavy rows verified, windows, substitution rounds,
//             spurious stops, rows substituted, largest window reached, rows of Pm moved (set/cleared)
// percol (optional): [cap_cols, 6] per residual column in processing order: birth rank, last row reached, windows, rounds,
//             heavy rows, events
int64_t rips_model_h1_pend(const float* dist, int n, int w0, int wmax, double* pairs_out, int64_t cap, int64_t* stats,
                           int64_t* percol, int64_t cap_cols) {
  Model m;
  std::vector<int64_t> residual; int64_t n_app = 0; int maxdepth = 0;
  build_model(m, dist, n, residual, n_app, maxdepth);
  const int64_t T = m.T;
  const int W = (n + 63) / 64;
  if (w0 < 1) w0 = 1;
  if (wmax < w0) wmax = w0;
  std::vector<uint8_t> x(T, 0);
  std::vector<uint64_t> X((size_t)n * W, 0), Pm((size_t)n * W, 0);
  int64_t p_pos = 0;   // Pm = adjacency of the edges with rank < p_pos
  std::vector<uint8_t> touched(n, 0);
  std::unordered_map<int64_t, int> owner;
  std::vector<std::vector<int>> Vs;
  int64_t n_pairs = 0, events = 0, flips = 0, undone = 0, heavy_rows = 0, windows = 0, rounds = 0, spurious = 0, subst_rows = 0, maxwin = 0, pm_moved = 0;
  std::vector<int64_t> members;
  auto flip = [&](int64_t e) {
    x[e] ^= 1; const int a = m.ea[e], b = m.eb[e];
    X[(size_t)a * W + (b >> 6)] ^= 1ull << (b & 63); X[(size_t)b * W + (a >> 6)] ^= 1ull << (a & 63);
    touched[a] = touched[b] = 1;
    if (x[e]) members.push_back(e);
  };
  auto p_move = [&](int64_t target) {
    while (p_pos < target) { const int a = m.ea[p_pos], b = m.eb[p_pos]; Pm[(size_t)a * W + (b >> 6)] |= 1ull << (b & 63); Pm[(size_t)b * W + (a >> 6)] |= 1ull << (a & 63); ++p_pos; ++pm_moved; }
    while (p_pos > target) { --p_pos; const int a = m.ea[p_pos], b = m.eb[p_pos]; Pm[(size_t)a * W + (b >> 6)] &= ~(1ull << (b & 63)); Pm[(size_t)b * W + (a >> 6)] &= ~(1ull << (a & 63)); ++pm_moved; }
  };
  int64_t col_no = 0;
  for (int64_t ci = (int64_t)residual.size() - 1; ci >= 0; --ci, ++col_no) {
    const int64_t b = residual[ci];
    members.clear();
    std::fill(touched.begin(), touched.end(), 0);
    flip(b);
    int64_t pos = b + 1;
    int64_t win = w0;
    bool essential = false; int64_t pivM = -1; int pivw = -1;
    int64_t c_windows = 0, c_rounds = 0, c_heavy = 0, c_events = 0;
    for (;;) {
      if (pos >= T) { essential = true; break; }
      const int64_t hi = std::min<int64_t>(T, pos + win);
      ++windows; ++c_windows; maxwin = std::max(maxwin, hi - pos);
      // (A) substitution over every apparent row of the window, in dependency rounds
      std::vector<int64_t> fliplist;
      {
        std::vector<uint8_t> done((size_t)(hi - pos), 0);
        std::vector<int64_t> pending;
        for (int64_t M = pos; M < hi; ++M) { if (m.apex[M] < 0) done[M - pos] = 1; else pending.push_back(M); }
        subst_rows += (int64_t)pending.size();
        while (!pending.empty()) {
          ++rounds; ++c_rounds;
          std::vector<int64_t> can, wait;
          for (int64_t M : pending) {
            const int64_t ra = m.pa[M], rb = m.pb[M];
            if ((ra >= pos && !done[ra - pos]) || (rb >= pos && !done[rb - pos])) wait.push_back(M); else can.push_back(M);
          }
          if (can.empty()) return -2;
          for (int64_t M : can) {
            const uint8_t want = x[m.pa[M]] ^ x[m.pb[M]];
            if (want != x[M]) { flip(M); ++flips; fliplist.push_back(M); }
          }
          for (int64_t M : can) done[M - pos] = 1;
          pending.swap(wait);
        }
      }
      // (B) verification of the heavy apparent rows against the superset mask Pend[c] & Pend[d]
      p_move(hi);
      int64_t Mf = -1;
      for (int64_t M = pos; M < hi; ++M) {
        if (m.apex[M] < 0) continue;
        const int c = m.ea[M], d = m.eb[M];
        if (!touched[c] && !touched[d]) continue;
        ++heavy_rows; ++c_heavy;
        const uint64_t xm = x[M] ? ~0ull : 0ull;
        bool bad = false;
        for (int k = 0; k < W && !bad; ++k)
          bad = ((xm ^ X[(size_t)c * W + k] ^ X[(size_t)d * W + k]) & Pm[(size_t)c * W + k] & Pm[(size_t)d * W + k]) != 0;
        if (bad) { Mf = M; break; }   // (the kernel takes the minimum over all failing rows; in order, the first one is it)
      }
      if (Mf < 0) { pos = hi; win = std::min<int64_t>(2 * win, wmax); continue; }
      // (C) stop at Mf: undo the flips above it, verify the row exactly
      for (int64_t e : fliplist) if (e > Mf) { flip(e); ++undone; }
      win = w0;
      {
        const int c = m.ea[Mf], d = m.eb[Mf];
        const int* Rc = &m.R[(size_t)c * n]; const int* Rd = &m.R[(size_t)d * n];
        pivM = -1;
        for (int w = n - 1; w >= 0; --w) {
          if (!(Rc[w] < Mf && Rd[w] < Mf)) continue;
          const int bit = (int)x[Mf] ^ (int)((X[(size_t)c * W + (w >> 6)] >> (w & 63)) & 1) ^ (int)((X[(size_t)d * W + (w >> 6)] >> (w & 63)) & 1);
          if (bit) { pivM = Mf; pivw = w; break; }
        }
      }
      if (pivM < 0) { ++spurious; pos = Mf + 1; continue; }
      const int64_t key = pivM * (int64_t)n + (n - 1 - pivw);
      auto it = owner.find(key);
      if (it == owner.end()) break;   // death
      ++events; ++c_events;
      for (int e : Vs[it->second]) flip(e);
      pos = pivM;   // this row again (its substitution is a no-op; the handled vertex is even now)
    }
    if (percol && col_no < cap_cols) {
      int64_t* o = percol + col_no * 6;
      o[0] = b; o[1] = essential ? T : pivM; o[2] = c_windows; o[3] = c_rounds; o[4] = c_heavy; o[5] = c_events;
    }
    std::vector<int> V;
    for (int64_t e : members) if (x[e]) { V.push_back((int)e); flip(e); }
    std::sort(V.begin(), V.end()); V.erase(std::unique(V.begin(), V.end()), V.end());
    maxwin = std::max<int64_t>(maxwin, 0);
    const float birth = m.len[b];
    if (essential) {
      if (n_pairs >= cap) return -1;
      pairs_out[2 * n_pairs] = birth; pairs_out[2 * n_pairs + 1] = INFINITY; ++n_pairs;
    } else {
      const float death = m.len[pivM];
      owner.emplace(pivM * (int64_t)n + (n - 1 - pivw), (int)Vs.size());
      Vs.push_back(std::move(V));
      if (death > birth) {
        if (n_pairs >= cap) return -1;
        pairs_out[2 * n_pairs] = birth; pairs_out[2 * n_pairs + 1] = death; ++n_pairs;
      }
    }
  }
  if (stats) {
    stats[0] = (int64_t)residual.size(); stats[1] = n_app; stats[2] = events; stats[3] = flips; stats[4] = undone; stats[5] = heavy_rows;
    stats[6] = windows; stats[7] = rounds; stats[8] = spurious; stats[9] = subst_rows; stats[10] = maxwin; stats[11] = pm_moved;
  }
  return n_pairs;
}


// rips_model_h1_pend with the two modes of the kernel (csrc/rips.cu, Sweeper2):
//   sparse mode (every column starts in it): windows w0, doubling up to wsparse; exact lunes from the rank rows, so the first
//       failing (row, vertex) of the window is the event; after an event the window size falls back to w0;
//   dense mode (entered when a window holds >= max(dense_min, rows / dense_div) heavy rows): Pm is moved to the window's end and
//       the superset mask is used; windows double up to wmax; the failing rows are examined in ascending order with the exact
//       lune until one is a true failure (the others are spurious stops: nothing is undone, nothing re-substituted); after an
//       event the window keeps its END (Pm only ever moves forward inside a column) and restarts at the event's row.
// stats: [14] residual columns, apparent edges, events, flips, flips undone, heavy rows verified, windows, substitution rounds,
//             spurious rows examined, rows substituted, rows handled in rounds >= 2, rows of Pm moved, dense columns, exact row checks
int64_t rips_model_h1_modes(const float* dist, int n, int w0, int wsparse, int wmax, int dense_min, int dense_div, double* pairs_out,
                            int64_t cap, int64_t* stats) {
  Model m;
  std::vector<int64_t> residual; int64_t n_app = 0; int maxdepth = 0;
  build_model(m, dist, n, residual, n_app, maxdepth);
  const int64_t T = m.T;
  const int W = (n + 63) / 64;
  std::vector<uint8_t> x(T, 0);
  std::vector<uint64_t> X((size_t)n * W, 0), Pm((size_t)n * W, 0);
  int64_t p_pos = 0;
  std::vector<uint8_t> touched(n, 0);
  std::unordered_map<int64_t, int> owner;
  std::vector<std::vector<int>> Vs;
  int64_t n_pairs = 0, events = 0, flips = 0, undone = 0, heavy_rows = 0, windows = 0, rounds = 0, spurious = 0, subst_rows = 0, late_rows = 0, pm_moved = 0,
          dense_cols = 0, exact_checks = 0;
  std::vector<int64_t> members;
  auto flip = [&](int64_t e) {
    x[e] ^= 1; const int a = m.ea[e], b = m.eb[e];
    X[(size_t)a * W + (b >> 6)] ^= 1ull << (b & 63); X[(size_t)b * W + (a >> 6)] ^= 1ull << (a & 63);
    touched[a] = touched[b] = 1;
    if (x[e]) members.push_back(e);
  };
  auto p_move = [&](int64_t target) {
    const int64_t dist_rows = target > p_pos ? target - p_pos : p_pos - target;
    if (dist_rows * 32 > (int64_t)n * n) {   // rebuild from the rank matrix (counted as n*n/32 rows)
      std::fill(Pm.begin(), Pm.end(), 0);
      for (int64_t r = 0; r < target; ++r) { const int a = m.ea[r], b = m.eb[r]; Pm[(size_t)a * W + (b >> 6)] |= 1ull << (b & 63); Pm[(size_t)b * W + (a >> 6)] |= 1ull << (a & 63); }
      p_pos = target; pm_moved += (int64_t)n * n / 32;
      return;
    }
    while (p_pos < target) { const int a = m.ea[p_pos], b = m.eb[p_pos]; Pm[(size_t)a * W + (b >> 6)] |= 1ull << (b & 63); Pm[(size_t)b * W + (a >> 6)] |= 1ull << (a & 63); ++p_pos; ++pm_moved; }
    while (p_pos > target) { --p_pos; const int a = m.ea[p_pos], b = m.eb[p_pos]; Pm[(size_t)a * W + (b >> 6)] &= ~(1ull << (b & 63)); Pm[(size_t)b * W + (a >> 6)] &= ~(1ull << (a & 63)); ++pm_moved; }
  };
  // exact check of row M: highest failing vertex, or -1
  auto exact_row = [&](int64_t M) {
    ++exact_checks;
    const int c = m.ea[M], d = m.eb[M];
    const int* Rc = &m.R[(size_t)c * n]; const int* Rd = &m.R[(size_t)d * n];
    for (int w = n - 1; w >= 0; --w) {
      if (!(Rc[w] < M && Rd[w] < M)) continue;
      const int bit = (int)x[M] ^ (int)((X[(size_t)c * W + (w >> 6)] >> (w & 63)) & 1) ^ (int)((X[(size_t)d * W + (w >> 6)] >> (w & 63)) & 1);
      if (bit) return w;
    }
    return -1;
  };
  for (int64_t ci = (int64_t)residual.size() - 1; ci >= 0; --ci) {
    const int64_t b = residual[ci];
    members.clear();
    std::fill(touched.begin(), touched.end(), 0);
    flip(b);
    int64_t pos = b + 1, win = w0, hi = -1;   // hi >= 0: the window's end is kept (dense mode after an event)
    bool dense = false;
    bool essential = false; int64_t pivM = -1; int pivw = -1;
    for (;;) {
      if (pos >= T) { essential = true; break; }
      if (hi < 0 || hi <= pos) hi = std::min<int64_t>(T, pos + win);
      ++windows;
      const size_t mark = members.size();   // (the kernel: position in the V list where this window's flips start)
      std::vector<int64_t> fliplist;
      {
        std::vector<uint8_t> done((size_t)(hi - pos), 0);
        std::vector<int64_t> pending;
        for (int64_t M = pos; M < hi; ++M) if (m.apex[M] >= 0) pending.push_back(M);
        subst_rows += (int64_t)pending.size();
        bool first = true;
        while (!pending.empty()) {
          ++rounds;
          if (!first) late_rows += (int64_t)pending.size();
          first = false;
          std::vector<int64_t> can, wait;
          for (int64_t M : pending) {
            const int64_t ra = m.pa[M], rb = m.pb[M];
            const bool wa = ra >= pos && m.apex[ra] >= 0 && !done[ra - pos], wb = rb >= pos && m.apex[rb] >= 0 && !done[rb - pos];
            if (wa || wb) wait.push_back(M); else can.push_back(M);
          }
          if (can.empty()) return -2;
          for (int64_t M : can) {
            const uint8_t want = x[m.pa[M]] ^ x[m.pb[M]];
            if (want != x[M]) { flip(M); ++flips; fliplist.push_back(M); }
          }
          for (int64_t M : can) done[M - pos] = 1;
          pending.swap(wait);
        }
      }
      (void)mark;
      // heavy rows of the window
      std::vector<int64_t> heavy;
      for (int64_t M = pos; M < hi; ++M) if (m.apex[M] >= 0 && (touched[m.ea[M]] || touched[m.eb[M]])) heavy.push_back(M);
      if (!dense && (int64_t)heavy.size() >= std::max<int64_t>(dense_min, (hi - pos) / dense_div)) { dense = true; ++dense_cols; }
      heavy_rows += (int64_t)heavy.size();
      int64_t evM = -1; int evw = -1;
      if (!dense) {
        for (int64_t M : heavy) { const int w = exact_row(M); --exact_checks; if (w >= 0) { evM = M; evw = w; break; } }
      } else {
        p_move(hi);
        std::vector<int64_t> failing;
        for (int64_t M : heavy) {
          const int c = m.ea[M], d = m.eb[M];
          const uint64_t xm = x[M] ? ~0ull : 0ull;
          bool bad = false;
          for (int k = 0; k < W && !bad; ++k)
            bad = ((xm ^ X[(size_t)c * W + k] ^ X[(size_t)d * W + k]) & Pm[(size_t)c * W + k] & Pm[(size_t)d * W + k]) != 0;
          if (bad) failing.push_back(M);
        }
        for (int64_t M : failing) {   // ascending
          const int w = exact_row(M);
          if (w >= 0) { evM = M; evw = w; break; }
          ++spurious;
        }
        // all spurious is impossible: a failing bit under the superset mask is a true failure of a row of this window
        if (!failing.empty() && evM < 0) return -3;
      }
      if (evM < 0) { pos = hi; hi = -1; win = std::min<int64_t>(2 * win, dense ? wmax : wsparse); continue; }
      for (int64_t e : fliplist) if (e > evM) { flip(e); ++undone; }
      pivM = evM; pivw = evw;
      const int64_t key = pivM * (int64_t)n + (n - 1 - pivw);
      auto it = owner.find(key);
      if (it == owner.end()) break;   // death
      ++events;
      for (int e : Vs[it->second]) flip(e);
      pos = pivM;
      if (!dense) { win = w0; hi = -1; }
    }
    std::vector<int> V;
    for (int64_t e : members) if (x[e]) { V.push_back((int)e); flip(e); }
    std::sort(V.begin(), V.end()); V.erase(std::unique(V.begin(), V.end()), V.end());
    const float birth = m.len[b];
    if (essential) {
      if (n_pairs >= cap) return -1;
      pairs_out[2 * n_pairs] = birth; pairs_out[2 * n_pairs + 1] = INFINITY; ++n_pairs;
    } else {
      const float death = m.len[pivM];
      owner.emplace(pivM * (int64_t)n + (n - 1 - pivw), (int)Vs.size());
      Vs.push_back(std::move(V));
      if (death > birth) {
        if (n_pairs >= cap) return -1;
        pairs_out[2 * n_pairs] = birth; pairs_out[2 * n_pairs + 1] = death; ++n_pairs;
      }
    }
  }
  if (stats) {
    stats[0] = (int64_t)residual.size(); stats[1] = n_app; stats[2] = events; stats[3] = flips; stats[4] = undone; stats[5] = heavy_rows;
    stats[6] = windows; stats[7] = rounds; stats[8] = spurious; stats[9] = subst_rows; stats[10] = late_rows; stats[11] = pm_moved;
    stats[12] = dense_cols; stats[13] = exact_checks;
  }
  return n_pairs;
}

}  // extern "C"
