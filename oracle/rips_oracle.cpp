// oracle/rips_oracle.cpp -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline), never the product path.
//
// CPU restatement of the Vietoris-Rips persistent cohomology algorithm that the reference
// reaches through `from ripser import ripser` (debug_tda_pipeline.py:10,109-110;
// analyze_tda_over_layers.py:5,76; analyze_adversarial_tda.py:12,100-101).  The algorithm
// itself lives in the third-party package `ripser` (ripser.py, UNPINNED in README.md:28; C++
// core = Bauer's Ripser), which is not vendored under /root/reference and not installed in
// this image, so this file restates the published algorithm (U. Bauer, "Ripser: efficient
// computation of Vietoris-Rips persistence barcodes", JACT 2021):
//   * dense float32 distances, Z/2 coefficients, enclosing-radius threshold,
//   * simplices indexed by the combinatorial number system (64-bit),
//   * filtration order = (diameter asc, index desc); columns processed in reverse order,
//   * H0 by union-find over sorted edges, emitting (0,d) for every merging edge with d != 0,
//   * H_q (q>=1) by implicit coboundary-matrix reduction with a binary heap working column,
//     the emergent-pair shortcut, a pivot->column hash map and clearing between dimensions,
//   * pairs emitted in processing order, zero-persistence pairs dropped.
//   * optional (`apparent` = 1; Ripser 1.2, Bauer 2021 section 4.2 "apparent pairs"): a zero-persistence pair (s, c) with c the
//     oldest cofacet of s and s the youngest facet of c is a persistence pair that needs no reduction; such columns are neither
//     assembled nor stored in the pivot map but recognised on the fly.  Same pairs, same rows, same order (tests compare the two
//     modes bit for bit); it only makes the top dimension of config C2 (1.2e9 triangles at n = 2000) fit in memory.
//     In this mode the working coboundary is also WINDOWED by diameter: only cofacets up to a bound B (a number of edge ranks
//     above the column's own diameter) are kept in the heap; every cofacet that is left out is longer than everything kept, so
//     the heap's pivot is the column's pivot; when the kept part cancels completely, B moves up (doubling) and the cofacets
//     between the old and the new bound are enumerated again from the column and its reduction column.  Exact, and the long
//     reductions of C2 (one column: > 1e9 live heap entries otherwise) stay in a few GB.
// Parity is PINNED for this file: tests/test_oracle_golden.py checks it against the 32
// shipped point clouds + summary_stats.json of the reference (tda-output/), see tests/golden/.
//
// Single-threaded on purpose (Ripser is single-threaded) -- it doubles as the CPU baseline (bench.py times the lean mode, the
// faster of the two).  One exception, outside the baseline's path: with maxdim >= 2 the lean mode assembles the columns of the next
// dimension on all host threads (the apparent-pair test of 1.2e9 triangles for config C2); the reduction itself stays serial.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <queue>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

typedef int64_t idx_t;
typedef float val_t;

struct Simplex {
  val_t diam;
  idx_t idx;
};

// "reverse filtration" comparator: greater diameter first, ties -> smaller index first.
struct RevFiltLess {
  bool operator()(const Simplex& a, const Simplex& b) const {
    return (a.diam > b.diam) || (a.diam == b.diam && a.idx < b.idx);
  }
};
// heap comparator: top() = smallest diameter, ties -> largest index (the column pivot).
struct HeapCmp {
  bool operator()(const Simplex& a, const Simplex& b) const {
    return (a.diam > b.diam) || (a.diam == b.diam && a.idx < b.idx);
  }
};
// Binary heap of the working coboundary (std::push_heap / pop_heap on a vector = what std::priority_queue does), with one
// addition: entries cancel in pairs (Z/2), and pop_pivot only cancels them when they reach the top, so a long reduction piles up
// billions of dead pairs (config C2 at n = 2000: tens of GB).  compact() removes the pairs in place once the heap has grown past
// a bound; the multiset of surviving entries, hence every pivot, is unchanged.
size_t g_window_min = 64, g_window_div = 256;   // first window of the lean mode: max(min, edges / div) edge ranks (tests: 1 rank)
size_t g_compact_min = (size_t)1 << 25;   // 32 M entries = 512 MB (tests lower it through rips_oracle_set_compact)
struct Heap {
  std::vector<Simplex> v;
  size_t next_compact = g_compact_min;
  bool empty() const { return v.empty(); }
  const Simplex& top() const { return v.front(); }
  void push(const Simplex& s) {
    v.push_back(s);
    std::push_heap(v.begin(), v.end(), HeapCmp());
    if (v.size() >= next_compact) compact();
  }
  void pop() {
    std::pop_heap(v.begin(), v.end(), HeapCmp());
    v.pop_back();
  }
  void compact() {
    std::sort(v.begin(), v.end(), [](const Simplex& a, const Simplex& b) { return a.idx < b.idx; });
    size_t w = 0;
    for (size_t a = 0; a < v.size();) {
      size_t b = a;
      while (b < v.size() && v[b].idx == v[a].idx) ++b;
      if ((b - a) & 1) v[w++] = v[a];
      a = b;
    }
    v.resize(w);
    std::make_heap(v.begin(), v.end(), HeapCmp());
    next_compact = std::max(g_compact_min, 2 * w);
  }
};

struct Binom {
  std::vector<std::vector<idx_t>> t;  // t[k][n]
  void init(int n, int k) {
    t.assign(k + 1, std::vector<idx_t>(n + 1, 0));
    for (int i = 0; i <= n; ++i) {
      t[0][i] = 1;
      for (int j = 1; j <= std::min(i, k); ++j) t[j][i] = (j == i) ? 1 : t[j - 1][i - 1] + t[j][i - 1];
    }
  }
  idx_t operator()(int n, int k) const { return (k > n || n < 0) ? 0 : t[k][n]; }
};

struct Stats {
  int64_t columns[3], emergent[3], reduced[3], additions[3], pops[3], max_v[3], cofacets[3];
  // dependency structure of the reduced (non-emergent) columns: cost of a column = its pivot steps (additions + 1); a column
  // depends on every reduced column it adds.  total = sum of costs, critical = heaviest dependency chain, depth = its length.
  int64_t dep_total[3], dep_critical[3], dep_depth[3];
};

struct Rips {
  int n;
  int maxdim;
  val_t thresh;
  std::vector<val_t> dist;  // full n*n
  std::vector<val_t> distT;  // its transpose: the cofacet enumerator reads d(v, s) for consecutive v along a ROW of distT (one
                            // cache line per 16 cofacets instead of one miss per cofacet); same values, same semantics
  Binom C;
  std::vector<std::vector<double>> dgm;       // per dim: flat (birth, death)
  std::vector<std::vector<idx_t>> pair_simplex;  // per dim: flat (birth idx, death idx or -1)
  int64_t num_edges;
  Stats st;
  bool apparent = false;  // Ripser 1.2's zero-apparent-pair shortcut + windowed working coboundary (see the header)
  std::vector<val_t> edge_diams;  // ascending lengths of the edges <= thresh (the window bounds of the lean mode)

  val_t d(int i, int j) const { return dist[(size_t)i * n + j]; }

  int max_vertex(idx_t idx, int k, int top) const {
    // largest v <= top with C(v,k) <= idx
    int lo = k - 1, hi = top;  // C(k-1,k)=0 <= idx always
    while (lo < hi) {
      int mid = lo + (hi - lo + 1) / 2;
      if (C(mid, k) <= idx) lo = mid; else hi = mid - 1;
    }
    return lo;
  }
  void vertices(idx_t idx, int dim, int* out) const {
    int top = n - 1;
    for (int k = dim + 1; k > 0; --k) {
      int v = max_vertex(idx, k, top);
      out[dim + 1 - k] = v;  // descending order
      idx -= C(v, k);
      top = v - 1;
    }
  }
  val_t diameter(idx_t idx, int dim) const {
    int vs[8];
    vertices(idx, dim, vs);
    val_t m = 0;
    for (int a = 0; a <= dim; ++a)
      for (int b = a + 1; b <= dim; ++b) m = std::max(m, d(vs[a], vs[b]));
    return m;
  }

  // cofacet enumeration in decreasing order of the added vertex (= decreasing cofacet index)
  struct Cofacets {
    const Rips& r;
    idx_t below, above;
    int v, k, dim;
    int vs[8];
    const val_t* row[8];  // row[a][v] == d(v, vs[a])
    val_t sdiam;
    Cofacets(const Rips& r_, Simplex s, int dim_) : r(r_), below(s.idx), above(0), v(r_.n - 1), k(dim_ + 1), dim(dim_), sdiam(s.diam) {
      r.vertices(s.idx, dim, vs);
      for (int a = 0; a <= dim; ++a) row[a] = r.distT.data() + (size_t)vs[a] * r.n;
    }
    bool has_next(bool all = true) {
      if (!all) return v >= k && r.C(v, k) > below;  // only vertices above every simplex vertex
      while (v != -1 && r.C(v, k) <= below) {
        below -= r.C(v, k);
        above += r.C(v, k + 1);
        --v; --k;
      }
      return v != -1;
    }
    Simplex next() {
      val_t cd = sdiam;
      for (int a = 0; a <= dim; ++a) cd = std::max(cd, row[a][v]);
      idx_t ci = above + r.C(v, k + 1) + below;
      --v;
      return Simplex{cd, ci};
    }
  };

  // oldest cofacet of s with the diameter of s (the enumerator runs in decreasing index = increasing age among equal diameters)
  Simplex zero_pivot_cofacet(const Simplex& s, int dim) const {
    Cofacets cf(*this, s, dim);
    while (cf.has_next()) {
      Simplex c = cf.next();
      if (c.diam == s.diam) return c;
    }
    return Simplex{0, -1};
  }
  // youngest facet of c with the diameter of c: dropping a larger vertex gives a smaller index, so the first hit is the youngest
  Simplex zero_pivot_facet(const Simplex& c, int dimc) const {
    int vs[8];
    vertices(c.idx, dimc, vs);
    for (int r = 0; r <= dimc; ++r) {
      val_t m = 0;
      idx_t fi = 0;
      int k = dimc;
      for (int a = 0; a <= dimc; ++a) {
        if (a == r) continue;
        fi += C(vs[a], k);
        --k;
        for (int b = a + 1; b <= dimc; ++b)
          if (b != r) m = std::max(m, d(vs[a], vs[b]));
      }
      if (m == c.diam) return Simplex{m, fi};
    }
    return Simplex{0, -1};
  }
  // the facet e of c with (e, c) a zero apparent pair, or idx -1
  Simplex zero_apparent_facet(const Simplex& c, int dimc) const {
    Simplex f = zero_pivot_facet(c, dimc);
    if (f.idx == -1) return f;
    Simplex cc = zero_pivot_cofacet(f, dimc - 1);
    return (cc.idx == c.idx) ? f : Simplex{0, -1};
  }
  // the cofacet c of s with (s, c) a zero apparent pair, or idx -1
  Simplex zero_apparent_cofacet(const Simplex& s, int dim) const {
    Simplex c = zero_pivot_cofacet(s, dim);
    if (c.idx == -1) return c;
    Simplex f = zero_pivot_facet(c, dim + 1);
    return (f.idx == s.idx) ? c : Simplex{0, -1};
  }
  // column s (dimension dim >= 1) takes no part in the reduction: it is the birth or the death of a zero apparent pair.  The facet
  // side is asked for dim >= 2 only: the pairs of edges with vertices are dimension 0's (union-find), as in Ripser.
  bool in_zero_apparent_pair(const Simplex& s, int dim) const {
    if (zero_apparent_cofacet(s, dim).idx != -1) return true;
    return dim >= 2 && zero_apparent_facet(s, dim).idx != -1;
  }

  static Simplex pop_pivot(Heap& h, int64_t& pops) {
    if (h.empty()) return Simplex{0, -1};
    Simplex p = h.top(); h.pop(); ++pops;
    while (!h.empty() && h.top().idx == p.idx) {
      h.pop(); ++pops;
      if (h.empty()) return Simplex{0, -1};
      p = h.top(); h.pop(); ++pops;
    }
    return p;
  }
  static Simplex get_pivot(Heap& h, int64_t& pops) {
    Simplex p = pop_pivot(h, pops);
    if (p.idx != -1) h.push(p);
    return p;
  }

  void run() {
    C.init(n, maxdim + 2);
    dgm.assign(maxdim + 1, {});
    pair_simplex.assign(maxdim + 1, {});
    std::memset(&st, 0, sizeof(st));
    const double INF = std::numeric_limits<double>::infinity();
    if (!(thresh < std::numeric_limits<val_t>::infinity())) {
      // enclosing radius: min_i max_j d(i,j)
      val_t enc = std::numeric_limits<val_t>::infinity();
      for (int i = 0; i < n; ++i) {
        val_t r = 0;
        for (int j = 0; j < n; ++j) r = std::max(r, d(i, j));
        enc = std::min(enc, r);
      }
      thresh = (n > 0) ? enc : 0;
    }
    // ---- dimension 0
    std::vector<Simplex> edges;
    for (int i = 1; i < n; ++i)
      for (int j = 0; j < i; ++j)
        if (d(i, j) <= thresh) edges.push_back(Simplex{d(i, j), C(i, 2) + j});
    num_edges = (int64_t)edges.size();
    std::sort(edges.begin(), edges.end(), RevFiltLess());  // reverse filtration order
    if (apparent) {
      edge_diams.resize(edges.size());
      for (size_t i = 0; i < edges.size(); ++i) edge_diams[i] = edges[edges.size() - 1 - i].diam;
    }
    std::vector<int> parent(n), rnk(n, 0);
    for (int i = 0; i < n; ++i) parent[i] = i;
    auto find = [&](int x) {
      int y = x, z;
      while ((z = parent[y]) != y) y = z;
      while ((z = parent[x]) != y) { parent[x] = y; x = z; }
      return y;
    };
    std::vector<Simplex> columns;
    for (auto it = edges.rbegin(); it != edges.rend(); ++it) {  // filtration order
      int vs[2];
      vertices(it->idx, 1, vs);
      int u = find(vs[0]), v = find(vs[1]);
      if (u != v) {
        if (it->diam != 0) {
          dgm[0].push_back(0.0); dgm[0].push_back((double)it->diam);
          pair_simplex[0].push_back(-1); pair_simplex[0].push_back(it->idx);
        }
        if (rnk[u] > rnk[v]) parent[v] = u;
        else { parent[u] = v; if (rnk[u] == rnk[v]) ++rnk[v]; }
      } else if (maxdim >= 1) {
        columns.push_back(*it);
      }
    }
    std::reverse(columns.begin(), columns.end());
    if (apparent) {
      std::vector<Simplex> kept;
      for (const Simplex& e : columns)
        if (zero_apparent_cofacet(e, 1).idx == -1) kept.push_back(e);
      columns.swap(kept);
    }
    for (int i = 0; i < n; ++i)
      if (find(i) == i) {
        dgm[0].push_back(0.0); dgm[0].push_back(INF);
        pair_simplex[0].push_back(i); pair_simplex[0].push_back(-1);
      }
    // ---- higher dimensions
    std::vector<Simplex> simplices = edges;  // all dim-1 simplices <= thresh (any order)
    for (int dim = 1; dim <= maxdim; ++dim) {
      std::unordered_map<idx_t, int64_t> pivot_col;
      pivot_col.reserve(columns.size());
      reduce(columns, pivot_col, dim);
      if (dim < maxdim) {
        std::vector<Simplex> next_simplices, next_columns;
        const bool keep_simplices = dim + 1 < maxdim;   // the list is only read to assemble the dimension after the next
        if (!apparent) {
          for (const Simplex& s : simplices) {
            Cofacets cf(*this, s, dim);
            while (cf.has_next(false)) {
              Simplex c = cf.next();
              if (c.diam <= thresh) {
                if (keep_simplices) next_simplices.push_back(c);
                if (pivot_col.find(c.idx) == pivot_col.end()) next_columns.push_back(c);
              }
            }
          }
        } else {
          // the apparent-pair test dominates (one cofacet scan + one facet scan per simplex): threads over slices of the list;
          // the order of the result is fixed by the sort below
          unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
          std::vector<std::vector<Simplex>> part_s(nt), part_c(nt);
          std::vector<std::thread> pool;
          for (unsigned t = 0; t < nt; ++t)
            pool.emplace_back([&, t]() {
              for (size_t i = t; i < simplices.size(); i += nt) {
                Cofacets cf(*this, simplices[i], dim);
                while (cf.has_next(false)) {
                  Simplex c = cf.next();
                  if (c.diam <= thresh) {
                    if (keep_simplices) part_s[t].push_back(c);
                    if (pivot_col.find(c.idx) == pivot_col.end() && !in_zero_apparent_pair(c, dim + 1)) part_c[t].push_back(c);
                  }
                }
              }
            });
          for (auto& th : pool) th.join();
          for (unsigned t = 0; t < nt; ++t) {
            next_simplices.insert(next_simplices.end(), part_s[t].begin(), part_s[t].end());
            next_columns.insert(next_columns.end(), part_c[t].begin(), part_c[t].end());
          }
        }
        std::sort(next_columns.begin(), next_columns.end(), RevFiltLess());
        simplices.swap(next_simplices);
        columns.swap(next_columns);
      }
    }
  }

  static void cancel_pairs(std::vector<idx_t>& v) {   // sorted, entries that occur an even number of times removed
    std::sort(v.begin(), v.end());
    size_t w = 0;
    for (size_t a = 0; a < v.size();) {
      size_t b = a;
      while (b < v.size() && v[b] == v[a]) ++b;
      if ((b - a) & 1) v[w++] = v[a];
      a = b;
    }
    v.resize(w);
  }

  void reduce(const std::vector<Simplex>& columns, std::unordered_map<idx_t, int64_t>& pivot_col, int dim) {
    const double INF = std::numeric_limits<double>::infinity();
    std::vector<std::vector<idx_t>> V(columns.size());  // reduction columns (excluding the column itself)
    st.columns[dim] = (int64_t)columns.size();
    std::vector<Simplex> buf;
    std::vector<int64_t> chain(columns.size(), 0), depth(columns.size(), 0);   // 0 = emergent column (no work to wait for)
    for (size_t j = 0; j < columns.size(); ++j) {
      const Simplex col = columns[j];
      int64_t steps = 1, dep_chain = 0, dep_depth = 0;
      Heap work;           // working coboundary
      std::vector<idx_t> vcol;  // working reduction column entries (with multiplicity)
      size_t vcol_compact = g_compact_min;
      // lean mode: the heap holds the cofacets with diameter <= bound only (bound = +inf: everything up to thresh)
      const val_t VINF = std::numeric_limits<val_t>::infinity();
      val_t bound = VINF;
      size_t bound_rank = 0, bound_step = 0;
      if (apparent && !edge_diams.empty()) {
        bound_step = std::max<size_t>(g_window_min, edge_diams.size() / g_window_div);
        bound_rank = (size_t)(std::upper_bound(edge_diams.begin(), edge_diams.end(), col.diam) - edge_diams.begin()) + bound_step;
        bound = bound_rank < edge_diams.size() ? edge_diams[bound_rank] : VINF;
      }
      Simplex pivot{0, -1};
      bool emergent = false;
      {  // init coboundary + emergent-pair check
        buf.clear();
        bool check = true;
        Cofacets cf(*this, col, dim);
        while (cf.has_next()) {
          Simplex c = cf.next();
          ++st.cofacets[dim];
          if (c.diam <= thresh) {
            buf.push_back(c);
            if (check && c.diam == col.diam) {
              if (pivot_col.find(c.idx) == pivot_col.end() && (!apparent || zero_apparent_facet(c, dim + 1).idx == -1)) { pivot = c; emergent = true; break; }
              check = false;
            }
          }
        }
        if (!emergent) {
          for (const Simplex& c : buf)
            if (c.diam <= bound) work.push(c);
          pivot = get_pivot(work, st.pops[dim]);
        }
      }
      if (emergent) ++st.emergent[dim]; else ++st.reduced[dim];
      for (;;) {
        if (pivot.idx == -1 && bound < VINF) {
          // everything up to the bound cancelled: move the bound up and enumerate the cofacets in (old bound, new bound] of the
          // column and of its reduction column again
          const val_t lo = bound;
          bound_step *= 2;
          bound_rank += bound_step;
          bound = bound_rank < edge_diams.size() ? edge_diams[bound_rank] : VINF;
          cancel_pairs(vcol);
          auto refill = [&](const Simplex& s) {
            Cofacets cf(*this, s, dim);
            while (cf.has_next()) {
              Simplex c = cf.next();
              ++st.cofacets[dim];
              if (c.diam <= thresh && c.diam > lo && c.diam <= bound) work.push(c);
            }
          };
          refill(col);
          for (idx_t sidx : vcol) refill(Simplex{diameter(sidx, dim), sidx});
          pivot = get_pivot(work, st.pops[dim]);
          continue;
        }
        if (pivot.idx == -1) {
          dgm[dim].push_back((double)col.diam); dgm[dim].push_back(INF);
          pair_simplex[dim].push_back(col.idx); pair_simplex[dim].push_back(-1);
          break;
        }
        auto it = pivot_col.find(pivot.idx);
        if (it != pivot_col.end()) {
          size_t a = (size_t)it->second;
          ++st.additions[dim];
          ++steps;
          if (chain[a] > dep_chain) dep_chain = chain[a];
          if (depth[a] > dep_depth) dep_depth = depth[a];
          // add column a: its own coboundary plus the coboundaries of its reduction column
          auto add_simplex = [&](idx_t sidx) {
            vcol.push_back(sidx);
            if (vcol.size() >= vcol_compact) {   // same idea for the working reduction column: cancel pairs early
              cancel_pairs(vcol);
              vcol_compact = std::max(g_compact_min, 2 * vcol.size());
            }
            Simplex s{diameter(sidx, dim), sidx};
            Cofacets cf(*this, s, dim);
            while (cf.has_next()) {
              Simplex c = cf.next();
              ++st.cofacets[dim];
              if (c.diam <= thresh && c.diam <= bound) work.push(c);
            }
          };
          add_simplex(columns[a].idx);
          for (idx_t s : V[a]) add_simplex(s);
          pivot = get_pivot(work, st.pops[dim]);
          continue;
        }
        if (apparent) {
          Simplex e = zero_apparent_facet(pivot, dim + 1);
          if (e.idx != -1) {   // the pivot belongs to an apparent column e (younger than col, empty reduction column): add it
            ++st.additions[dim];
            ++steps;
            vcol.push_back(e.idx);
            Cofacets cf(*this, e, dim);
            while (cf.has_next()) {
              Simplex c = cf.next();
              ++st.cofacets[dim];
              if (c.diam <= thresh && c.diam <= bound) work.push(c);
            }
            pivot = get_pivot(work, st.pops[dim]);
            continue;
          }
        }
        if (pivot.diam > col.diam) {
          dgm[dim].push_back((double)col.diam); dgm[dim].push_back((double)pivot.diam);
          pair_simplex[dim].push_back(col.idx); pair_simplex[dim].push_back(pivot.idx);
        }
        pivot_col.emplace(pivot.idx, (int64_t)j);
        // store the reduction column mod 2
        cancel_pairs(vcol);
        V[j].assign(vcol.begin(), vcol.end());
        st.max_v[dim] = std::max<int64_t>(st.max_v[dim], (int64_t)V[j].size());
        if (!emergent) {
          chain[j] = steps + dep_chain;
          depth[j] = 1 + dep_depth;
          st.dep_total[dim] += steps;
          st.dep_critical[dim] = std::max(st.dep_critical[dim], chain[j]);
          st.dep_depth[dim] = std::max(st.dep_depth[dim], depth[j]);
        }
        break;
      }
    }
  }
};

}  // namespace

extern "C" {

// dist: full n*n float32 matrix (row-major, symmetric, zero diagonal). thresh = +inf -> enclosing radius.
void* rips_oracle_run2(const float* dist, int n, int maxdim, float thresh, int apparent);
void* rips_oracle_run(const float* dist, int n, int maxdim, float thresh) { return rips_oracle_run2(dist, n, maxdim, thresh, 0); }
void* rips_oracle_run2(const float* dist, int n, int maxdim, float thresh, int apparent) {
  Rips* r = new Rips();
  r->n = n; r->maxdim = maxdim; r->thresh = thresh; r->apparent = apparent != 0;
  r->dist.assign(dist, dist + (size_t)n * n);
  r->distT.resize((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) r->distT[(size_t)j * n + i] = dist[(size_t)i * n + j];
  r->run();
  return r;
}
int64_t rips_oracle_count(void* h, int dim) { return (int64_t)((Rips*)h)->dgm[dim].size() / 2; }
void rips_oracle_get(void* h, int dim, double* pairs, int64_t* simplices) {
  Rips* r = (Rips*)h;
  if (pairs) std::memcpy(pairs, r->dgm[dim].data(), r->dgm[dim].size() * sizeof(double));
  if (simplices) std::memcpy(simplices, r->pair_simplex[dim].data(), r->pair_simplex[dim].size() * sizeof(int64_t));
}
int64_t rips_oracle_num_edges(void* h) { return ((Rips*)h)->num_edges; }
float rips_oracle_thresh(void* h) { return ((Rips*)h)->thresh; }
// stats layout per dim: columns, emergent, reduced, additions, pops, max_v, cofacets
void rips_oracle_stats(void* h, int dim, int64_t* out) {
  Rips* r = (Rips*)h;
  out[0] = r->st.columns[dim]; out[1] = r->st.emergent[dim]; out[2] = r->st.reduced[dim];
  out[3] = r->st.additions[dim]; out[4] = r->st.pops[dim]; out[5] = r->st.max_v[dim]; out[6] = r->st.cofacets[dim];
}
// dependency statistics of the reduced columns of one dimension: total pivot steps, heaviest dependency chain, its length
void rips_oracle_dep_stats(void* h, int dim, int64_t* out) {
  Rips* r = (Rips*)h;
  out[0] = r->st.dep_total[dim]; out[1] = r->st.dep_critical[dim]; out[2] = r->st.dep_depth[dim];
}
void rips_oracle_free(void* h) { delete (Rips*)h; }
// entries a working column may hold before its cancelling pairs are removed (default 2^25); tests use a few dozen
void rips_oracle_set_window(int64_t min_ranks, int64_t div) {
  g_window_min = (size_t)std::max<int64_t>(1, min_ranks);
  g_window_div = (size_t)std::max<int64_t>(1, div);
}
void rips_oracle_set_compact(int64_t min_entries) { g_compact_min = (size_t)std::max<int64_t>(4, min_entries); }

}  // extern "C"
