// depth of the apparent-edge dependency DAG: x_M = x_(c,w) ^ x_(d,w), w = apex(M)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>
#include <numeric>
#include <cstdint>
int main(int argc, char** argv) {
  int n = 2000; std::vector<float> P(n * 3);
  FILE* f = fopen(argv[1], "rb"); if (fread(P.data(), 4, n * 3, f) != (size_t)n * 3) return 1; fclose(f);
  int64_t E = (int64_t)n * (n - 1) / 2;
  std::vector<float> len(E); std::vector<int> ea(E), eb(E);
  int64_t q = 0;
  for (int i = 1; i < n; ++i) for (int j = 0; j < i; ++j) { double s = 0; for (int d = 0; d < 3; ++d) { double t = (double)P[i*3+d] - P[j*3+d]; s += t*t; } len[q] = sqrtf((float)s); ea[q] = i; eb[q] = j; ++q; }
  std::vector<int64_t> ord(E); std::iota(ord.begin(), ord.end(), 0);
  std::sort(ord.begin(), ord.end(), [&](int64_t x, int64_t y) { return len[x] < len[y] || (len[x] == len[y] && x > y); });
  std::vector<int> R((size_t)n * n, 0x7fffffff);
  for (int64_t r = 0; r < E; ++r) { int a = ea[ord[r]], b = eb[ord[r]]; R[(size_t)a*n+b] = R[(size_t)b*n+a] = (int)r; }
  // enclosing radius threshold
  float thr = 1e30f; for (int i = 0; i < n; ++i) { float mx = 0; for (int j = 0; j < n; ++j) if (j != i) { int a = std::max(i,j), b = std::min(i,j); mx = std::max(mx, len[(int64_t)a*(a-1)/2+b]); } thr = std::min(thr, mx); }
  int64_t T = 0; while (T < E && len[ord[T]] <= thr) ++T;
  std::vector<int> uf(n); std::iota(uf.begin(), uf.end(), 0);
  auto find = [&](int x) { while (uf[x] != x) { uf[x] = uf[uf[x]]; x = uf[x]; } return x; };
  std::vector<int> depth(T, 0); std::vector<char> kind(T, 0);  // 0 mst, 1 apparent, 2 residual
  int64_t napp = 0, nres = 0; int maxd = 0; std::vector<int64_t> hist(64, 0);
  for (int64_t r = 0; r < T; ++r) {
    int a = ea[ord[r]], b = eb[ord[r]];
    int ra = find(a), rb = find(b);
    if (ra != rb) { uf[ra] = rb; kind[r] = 0; continue; }
    const int* Ra = &R[(size_t)a*n]; const int* Rb = &R[(size_t)b*n];
    int apex = -1; for (int w = n - 1; w >= 0; --w) if (Ra[w] < r && Rb[w] < r) { apex = w; break; }
    if (apex < 0) { kind[r] = 2; ++nres; continue; }
    kind[r] = 1; ++napp;
    int d = 1 + std::max(depth[Ra[apex]], depth[Rb[apex]]);
    depth[r] = d; maxd = std::max(maxd, d);
  }
  // distribution of depth
  std::vector<int64_t> cnt(maxd + 1, 0); for (int64_t r = 0; r < T; ++r) if (kind[r] == 1) ++cnt[depth[r]];
  printf("n=%d T=%lld apparent=%lld residual=%lld max depth=%d\n", n, (long long)T, (long long)napp, (long long)nres, maxd);
  int64_t acc = 0; for (int d = 1; d <= maxd; ++d) { acc += cnt[d]; if (d == 1 || d == 2 || d == 4 || d == 8 || d == 16 || d == 32 || d == 64 || d == 128 || d == 256 || d == 512 || d == maxd) printf("  depth<=%d: %.4f of apparent edges\n", d, (double)acc / napp); }
  return 0;
}
