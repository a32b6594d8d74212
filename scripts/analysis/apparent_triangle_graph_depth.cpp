// depth of the apparent-TRIANGLE dependency graph (H2): x_t = x_(M,v) ^ x_(c,w,v) ^ x_(d,w,v), v = apex4(t)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>
#include <numeric>
#include <cstdint>
int main(int argc, char** argv) {
  int n = atoi(argv[2]); int dim = atoi(argv[3]);
  std::vector<float> P((size_t)n * dim);
  FILE* f = fopen(argv[1], "rb"); if (fread(P.data(), 4, (size_t)n * dim, f) != (size_t)n * dim) return 1; fclose(f);
  int64_t E = (int64_t)n * (n - 1) / 2;
  std::vector<float> len(E); std::vector<int> ea(E), eb(E);
  int64_t q = 0;
  for (int i = 1; i < n; ++i) for (int j = 0; j < i; ++j) { double s = 0; for (int d = 0; d < dim; ++d) { double t = (double)P[(size_t)i*dim+d] - P[(size_t)j*dim+d]; s += t*t; } len[q] = sqrtf((float)s); ea[q] = i; eb[q] = j; ++q; }
  std::vector<int64_t> ord(E); std::iota(ord.begin(), ord.end(), 0);
  std::sort(ord.begin(), ord.end(), [&](int64_t x, int64_t y) { return len[x] < len[y] || (len[x] == len[y] && x > y); });
  std::vector<int> R((size_t)n * n, 0x7fffffff);
  for (int64_t r = 0; r < E; ++r) { int a = ea[ord[r]], b = eb[ord[r]]; R[(size_t)a*n+b] = R[(size_t)b*n+a] = (int)r; }
  float thr = 1e30f; for (int i = 0; i < n; ++i) { float mx = 0; for (int j = 0; j < n; ++j) if (j != i) { int a = std::max(i,j), b = std::min(i,j); mx = std::max(mx, len[(int64_t)a*(a-1)/2+b]); } thr = std::min(thr, mx); }
  int64_t T = 0; while (T < E && len[ord[T]] <= thr) ++T;
  // triangle (M, w): key M*n + (n-1-w).  depth stored per (M,w)
  std::vector<uint16_t> depth((size_t)T * n, 0);
  auto tri_key = [&](int a, int b, int c2, int64_t& M, int& w) {  // triangle {a,b,c2}: longest edge rank M, opposite vertex w
    int rab = R[(size_t)a*n+b], rac = R[(size_t)a*n+c2], rbc = R[(size_t)b*n+c2];
    if (rab >= rac && rab >= rbc) { M = rab; w = c2; } else if (rac >= rab && rac >= rbc) { M = rac; w = b; } else { M = rbc; w = a; }
  };
  int64_t ntri = 0, napp = 0; int maxd = 0; std::vector<int64_t> hist(4096, 0);
  std::vector<int> lune;
  for (int64_t M = 0; M < T; ++M) {
    int c = ea[ord[M]], d = eb[ord[M]];
    const int* Rc = &R[(size_t)c*n]; const int* Rd = &R[(size_t)d*n];
    lune.clear(); for (int w = 0; w < n; ++w) if (Rc[w] < M && Rd[w] < M) lune.push_back(w);
    // process w descending (ascending key) so that (M, v) with v > w is already known
    for (int li = (int)lune.size() - 1; li >= 0; --li) {
      int w = lune[li]; ++ntri;
      const int* Rw = &R[(size_t)w*n];
      int v = -1; for (int lj = (int)lune.size() - 1; lj > li; --lj) if (Rw[lune[lj]] < M) { v = lune[lj]; break; }
      if (v < 0) continue;
      ++napp;
      int dd = depth[(size_t)M * n + v];
      int64_t M2; int w2; tri_key(c, w, v, M2, w2); dd = std::max<int>(dd, depth[(size_t)M2 * n + w2]);
      tri_key(d, w, v, M2, w2); dd = std::max<int>(dd, depth[(size_t)M2 * n + w2]);
      ++dd; depth[(size_t)M * n + w] = (uint16_t)dd; maxd = std::max(maxd, dd); if (dd < 4096) ++hist[dd];
    }
  }
  printf("n=%d T=%lld triangles=%lld apparent=%lld (%.4f) max depth=%d\n", n, (long long)T, (long long)ntri, (long long)napp, (double)napp / ntri, maxd);
  int64_t acc = 0; for (int d = 1; d <= maxd && d < 4096; ++d) { acc += hist[d]; if ((d & (d - 1)) == 0 || d == maxd) printf("  depth<=%d: %.4f\n", d, (double)acc / napp); }
  return 0;
}
