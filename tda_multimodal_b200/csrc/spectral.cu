// spectral.cu -- UMAP's spectral initialisation on the device.
//
// Replaces umap-learn's spectral_layout / multi_component_layout (scipy connected_components + ARPACK eigsh of
// the symmetric normalised Laplacian L = I - D^-1/2 W D^-1/2, k = dim+1 smallest eigenpairs, first dropped)
// as reached from umap.UMAP(init='spectral') (debug_tda_pipeline.py:96-104 and the other call sites).
//
//   components_kernel : connected components of the pruned fuzzy graph (min-label propagation + pointer
//                       jumping), component ids numbered by smallest member vertex like scipy does, degrees.
//   lanczos_kernel    : one CTA per (cloud, component): Lanczos with full (CGS2) reorthogonalisation on
//                       A = D^-1/2 W D^-1/2 restricted to the component, the trivial eigenvector D^1/2 1
//                       deflated; the small tridiagonal problem is solved by Sturm-sequence multisection +
//                       inverse iteration (fp64); Ritz vectors of the `dim` largest eigenvalues of A
//                       (= smallest non-trivial of L) are written as unit vectors, like ARPACK returns them.
// Signs / rotations inside degenerate eigenspaces are arbitrary in ARPACK too; parity is judged on the
// embedding's trustworthiness and the downstream diagrams (BASELINE.json north_star).
#include "common.cuh"
#include "launch_count.cuh"
#include "../../include/tda_b200.h"
#include <cmath>
#include <cstring>

namespace tda {
namespace spectral {

constexpr int kMaxKrylov = 96;
constexpr int kLanczosThreads = 512;
constexpr int kMaxDim = 4;
constexpr double kFixScale = 1099511627776.0;   // 2^40

// one CTA per cloud
__global__ void __launch_bounds__(1024) components_kernel(const int* __restrict__ head_g, const int* __restrict__ tail_g,
                                                          const float* __restrict__ weight_g, const float* __restrict__ eps_g, int slots, int n,
                                                          int* __restrict__ label_g, unsigned long long* __restrict__ degfix_g, int* __restrict__ comp_g,
                                                          float* __restrict__ deg_g, int* __restrict__ ncomp_g, int* __restrict__ csize_g) {
  __shared__ int s_scan[1024];
  __shared__ int s_carry;
  const int p = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int* head = head_g + (size_t)p * slots;
  const int* tail = tail_g + (size_t)p * slots;
  const float* weight = weight_g + (size_t)p * slots;
  const float* eps = eps_g + (size_t)p * slots;
  int* label = label_g + (size_t)p * n;
  int* comp = comp_g + (size_t)p * n;
  float* deg = deg_g + (size_t)p * n;
  int* csize = csize_g + (size_t)p * n;
  // degrees are summed in 2^-40 fixed point: integer atomics are associative, so the sum does not depend on the order the slots
  // arrive in (float atomics made the whole initialisation irreproducible from run to run)
  unsigned long long* degfix = degfix_g + (size_t)p * n;
  for (int i = tid; i < n; i += nt) { label[i] = i; degfix[i] = 0ull; csize[i] = 0; }
  __syncthreads();
  for (int e = tid; e < slots; e += nt)
    if (eps[e] > 0.f) atomicAdd(&degfix[head[e]], (unsigned long long)((double)weight[e] * kFixScale));
  __syncthreads();
  for (int i = tid; i < n; i += nt) deg[i] = (float)((double)degfix[i] * (1.0 / kFixScale));
  for (int round = 0; round < 4 * 1024; ++round) {
    int changed = 0;
    for (int e = tid; e < slots; e += nt)
      if (eps[e] > 0.f) {
        const int u = head[e], v = tail[e];
        const int lu = label[u], lv = label[v];
        if (lu < lv) { atomicMin(&label[v], lu); changed = 1; }
        else if (lv < lu) { atomicMin(&label[u], lv); changed = 1; }
      }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {  // pointer jumping
      int l = label[i];
      while (label[l] != l) l = label[l];
      label[i] = l;
    }
    if (!__syncthreads_or(changed)) break;
  }
  // component ids in order of the smallest member (roots are exactly the vertices with label[i] == i)
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += nt) {
    const int i = base + tid;
    const int isroot = (i < n && label[i] == i) ? 1 : 0;
    s_scan[tid] = isroot;
    __syncthreads();
    for (int off = 1; off < nt; off <<= 1) {
      int v = tid >= off ? s_scan[tid - off] : 0;
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    if (isroot) comp[i] = s_carry + s_scan[tid] - 1;
    __syncthreads();
    if (tid == nt - 1) s_carry += s_scan[tid];
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) {
    const int c = comp[label[i]];
    if (label[i] != i) comp[i] = c;
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) atomicAdd(&csize[comp[i]], 1);
  if (tid == 0) ncomp_g[p] = s_carry;
}

__device__ __forceinline__ uint32_t hash32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}

// eigenpairs of the meff x meff tridiagonal T (alpha diagonal, beta off-diagonal), the `dim` largest: Sturm-sequence multisection
// per eigenvalue (one warp each, 33 sections per round), inverse iteration (one thread each: the tridiagonal LU with partial
// pivoting is factored ONCE in shared memory, then three solves), Gram-Schmidt between the Ritz coefficient vectors.
// All threads of the CTA call; returns nev.  (Deterministic: CTAs that call it with equal inputs get equal outputs.)
// Measured before this version (scripts/lanczos_phase_cycles.py): 1.0 M of the 2.85 M cycles of a component -- the LU was
// re-factored in every sweep on arrays in local memory.
__device__ __forceinline__ int tridiag_eigs(int meff, int dim, const double* s_alpha, const double* s_beta, double* s_lam,
                                            double (*s_vec)[kMaxKrylov], int tid, int lane, int warp) {
  __shared__ double s_lu[kMaxDim][4][kMaxKrylov];      // dl (multipliers), dd (diagonal), du, du2 of the factored T - lam I
  __shared__ unsigned char s_piv[kMaxDim][kMaxKrylov];
  const int nev = min(dim, meff);
  if (warp < nev) {
    // Gershgorin bounds
    double lo = 1e300, hi = -1e300;
    for (int i = 0; i < meff; ++i) {
      const double r = (i > 0 ? fabs(s_beta[i - 1]) : 0.0) + (i < meff - 1 ? fabs(s_beta[i]) : 0.0);
      lo = fmin(lo, s_alpha[i] - r); hi = fmax(hi, s_alpha[i] + r);
    }
    const int want = meff - 1 - warp;  // index (ascending) of the eigenvalue this warp looks for
    for (int round = 0; round < 11; ++round) {   // the bracket shrinks 33x per round: 33^11 > 2^53 of the Gershgorin width
      const double x = lo + (hi - lo) * (double)(lane + 1) / 33.0;
      int cnt = 0;  // number of eigenvalues < x (Sturm sequence)
      double d = 1.0;
      for (int i = 0; i < meff; ++i) {
        const double b2 = i > 0 ? s_beta[i - 1] * s_beta[i - 1] : 0.0;
        d = s_alpha[i] - x - (i > 0 ? b2 / d : 0.0);
        if (fabs(d) < 1e-300) d = -1e-300;
        if (d < 0.0) ++cnt;
      }
      // eigenvalue `want` lies in (x_l, x_{l+1}] where l = last lane with cnt <= want
      const unsigned below = __ballot_sync(0xffffffffu, cnt <= want);
      const int nb = __popc(below);  // lanes 0..nb-1 have cnt <= want (monotone)
      const double step = (hi - lo) / 33.0;
      const double nlo = lo + step * nb, nhi = lo + step * (nb + 1);
      lo = nlo; hi = nhi;
    }
    if (lane == 0) s_lam[warp] = 0.5 * (lo + hi);
  }
  __syncthreads();
  if (tid < nev) {
    // inverse iteration on (T - lam I) with a tiny shift; tridiagonal LU with partial pivoting, factored once, 3 solves
    const double lam = s_lam[tid] + 1e-9 * (1.0 + fabs(s_lam[tid])) * (tid + 1);
    double* y = s_vec[tid];
    double* dl = s_lu[tid][0];
    double* dd = s_lu[tid][1];
    double* du = s_lu[tid][2];
    double* du2 = s_lu[tid][3];
    unsigned char* piv = s_piv[tid];
    for (int i = 0; i < meff; ++i) {
      y[i] = 1.0 + 0.01 * ((i * 37 + tid * 11) % 17);
      dd[i] = s_alpha[i] - lam;
      du[i] = i < meff - 1 ? s_beta[i] : 0.0;
      dl[i] = i < meff - 1 ? s_beta[i] : 0.0;
      du2[i] = 0.0;
    }
    for (int i = 0; i < meff - 1; ++i) {
      if (fabs(dd[i]) >= fabs(dl[i])) {
        piv[i] = 0;
        if (dd[i] == 0.0) dd[i] = 1e-300;
        const double f = dl[i] / dd[i];
        dl[i] = f;
        dd[i + 1] -= f * du[i];
      } else {
        piv[i] = 1;
        const double f = dd[i] / dl[i];
        dd[i] = dl[i];
        dl[i] = f;
        const double t = du[i];
        du[i] = dd[i + 1];
        dd[i + 1] = t - f * du[i];
        du2[i] = du[i + 1];
        du[i + 1] = -f * du2[i];
      }
    }
    if (dd[meff - 1] == 0.0) dd[meff - 1] = 1e-300;
    for (int sweep = 0; sweep < 3; ++sweep) {
      for (int i = 0; i < meff - 1; ++i) {
        if (piv[i]) { const double t = y[i]; y[i] = y[i + 1]; y[i + 1] = t - dl[i] * y[i]; }
        else y[i + 1] -= dl[i] * y[i];
      }
      y[meff - 1] /= dd[meff - 1];
      if (meff > 1) y[meff - 2] = (y[meff - 2] - du[meff - 2] * y[meff - 1]) / dd[meff - 2];
      for (int i = meff - 3; i >= 0; --i) y[i] = (y[i] - du[i] * y[i + 1] - du2[i] * y[i + 2]) / dd[i];
      double nn = 0.0;
      for (int i = 0; i < meff; ++i) nn += y[i] * y[i];
      nn = 1.0 / sqrt(nn);
      for (int i = 0; i < meff; ++i) y[i] *= nn;
    }
  }
  __syncthreads();
  if (tid == 0) {  // Gram-Schmidt between the few Ritz coefficient vectors (clustered eigenvalues)
    for (int a = 0; a < nev; ++a) {
      for (int b = 0; b < a; ++b) {
        double dot = 0.0;
        for (int i = 0; i < meff; ++i) dot += s_vec[a][i] * s_vec[b][i];
        for (int i = 0; i < meff; ++i) s_vec[a][i] -= dot * s_vec[b][i];
      }
      double nn = 0.0;
      for (int i = 0; i < meff; ++i) nn += s_vec[a][i] * s_vec[a][i];
      nn = nn > 0.0 ? 1.0 / sqrt(nn) : 0.0;
      for (int i = 0; i < meff; ++i) s_vec[a][i] *= nn;
    }
  }
  __syncthreads();
  return nev;
}

struct LanczosParams {
  const int* head; const int* tail; const float* weight; const float* eps; int slots; int n; int dim;
  const int* comp; const float* deg; const int* ncomp; const int* csize;
  float* Q;        // [batch, maxcomp, kMaxKrylov + 2, n]
  float* out;      // [batch, n, dim] eigenvectors (rows of other components untouched)
  float* evals;    // [batch, maxcomp, kMaxDim]
  int4* entries;   // [batch, slots]  (head, tail, weight / sqrt(deg_h deg_t), -) of the live slots, one segment per component
  int* ent_count;  // [batch]         segment allocation cursor
  unsigned long long* wfix;   // [batch, maxcomp, n]  sparse matrix-vector product accumulated in 2^-40 fixed point (order independent)
  int maxcomp; int min_size; uint64_t seed;
  int skip_connected;   // 1: clouds with one component are left to the cluster kernel
};

__device__ __forceinline__ void lanczos_body(const LanczosParams& P);
__global__ void __launch_bounds__(kLanczosThreads) lanczos_kernel(LanczosParams P) { lanczos_body(P); }
// the same with the minimum component size applied only to clouds with several components (a connected cloud is always laid out)
__global__ void __launch_bounds__(kLanczosThreads) lanczos_kernel_ms(LanczosParams P, int min_size_multi) {
  if (P.ncomp[blockIdx.y] > 1) P.min_size = min_size_multi;
  lanczos_body(P);
}
__device__ __forceinline__ void lanczos_body(const LanczosParams& P) {
  __shared__ double s_alpha[kMaxKrylov], s_beta[kMaxKrylov];
  __shared__ float s_coef[kMaxKrylov + 2];
  __shared__ float s_red[kLanczosThreads / 32];
  __shared__ double s_lam[kMaxDim];
  __shared__ double s_vec[kMaxDim][kMaxKrylov];
  __shared__ int s_m, s_base, s_cursor;
  const int p = blockIdx.y, c = blockIdx.x;
  if (c >= P.ncomp[p] || (P.skip_connected && P.ncomp[p] == 1)) return;
  const int n = P.n, dim = P.dim;
  const int nc = P.csize[(size_t)p * n + c];
  if (nc < P.min_size) return;  // tiny components are placed at random by the host
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kLanczosThreads / 32;
  const int* head = P.head + (size_t)p * P.slots;
  const int* tail = P.tail + (size_t)p * P.slots;
  const float* weight = P.weight + (size_t)p * P.slots;
  const float* eps = P.eps + (size_t)p * P.slots;
  const int* comp = P.comp + (size_t)p * n;
  const float* deg = P.deg + (size_t)p * n;
  float* Q = P.Q + ((size_t)p * P.maxcomp + c) * (size_t)(kMaxKrylov + 2) * n;
  float* u1 = Q;            // trivial eigenvector, normalised
  float* q0 = Q + n;        // Lanczos vectors q_0 ..
  unsigned long long* wfix = P.wfix + ((size_t)p * P.maxcomp + c) * (size_t)n;
  const int m = min(kMaxKrylov, nc - 1);

  auto block_sum = [&](float v) -> float {
    v = warp_sum_f32(v);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += s_red[w];
    __syncthreads();
    return t;
  };
  // u1 = D^1/2 1 restricted to the component, q_0 = random, orthogonal to u1
  float acc = 0.f;
  for (int i = tid; i < n; i += kLanczosThreads) {
    const float v = comp[i] == c ? sqrtf(deg[i]) : 0.f;
    u1[i] = v;
    acc += v * v;
  }
  float nrm = sqrtf(block_sum(acc));
  acc = 0.f;
  float acc2 = 0.f;
  for (int i = tid; i < n; i += kLanczosThreads) {
    const float u = u1[i] / nrm;
    u1[i] = u;
    const float r = comp[i] == c ? ((float)(hash32(P.seed + (uint64_t)i) >> 8) * (1.f / 8388608.f) - 1.f) : 0.f;
    q0[i] = r;
    acc += r * u;
  }
  const float d0 = block_sum(acc);
  for (int i = tid; i < n; i += kLanczosThreads) { const float r = q0[i] - d0 * u1[i]; q0[i] = r; acc2 += r * r; }
  nrm = sqrtf(block_sum(acc2));
  for (int i = tid; i < n; i += kLanczosThreads) q0[i] /= nrm;
  __syncthreads();

  // the component's live slots, compacted once with their matrix entry: the 96 sparse matrix-vector products then stream 16 bytes
  // per entry with independent loads instead of chasing eps -> head/tail -> deg per slot
  int4* ent = P.entries + (size_t)p * P.slots;
  int my_cnt = 0;
  for (int e = tid; e < P.slots; e += kLanczosThreads)
    if (eps[e] > 0.f && comp[head[e]] == c) ++my_cnt;
  const int cnt = (int)(block_sum((float)my_cnt) + 0.5f);   // (exact: counts are far below 2^24)
  if (tid == 0) { s_base = atomicAdd(&P.ent_count[p], cnt); s_cursor = 0; }
  __syncthreads();
  ent += s_base;
  for (int e0 = 0; e0 < P.slots; e0 += kLanczosThreads) {
    const int e = e0 + tid;
    int h = 0, t = 0;
    bool ok = false;
    if (e < P.slots && eps[e] > 0.f) { h = head[e]; t = tail[e]; ok = comp[h] == c; }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    int wbase = 0;
    if (lane == 0 && bal) wbase = atomicAdd(&s_cursor, __popc(bal));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (ok) ent[wbase + __popc(bal & ((1u << lane) - 1))] = make_int4(h, t, __float_as_int(weight[e] * rsqrtf(deg[h] * deg[t])), 0);
  }
  __syncthreads();

  int meff = m;
  for (int j = 0; j < m; ++j) {
    const float* qj = q0 + (size_t)j * n;
    float* w = q0 + (size_t)(j + 1) * n;  // becomes q_{j+1}
    for (int i = tid; i < n; i += kLanczosThreads) wfix[i] = 0ull;
    __syncthreads();
#pragma unroll 4
    for (int k = tid; k < cnt; k += kLanczosThreads) {
      const int4 E = ent[k];
      const long long v = __double2ll_rn((double)(__int_as_float(E.z) * qj[E.y]) * kFixScale);
      atomicAdd(&wfix[E.x], (unsigned long long)v);   // two's complement: signed sums wrap correctly
    }
    __syncthreads();
    for (int i = tid; i < n; i += kLanczosThreads) w[i] = (float)((double)(long long)wfix[i] * (1.0 / kFixScale));
    __syncthreads();
    // classical Gram-Schmidt, twice, against u1 and q_0..q_j; the first pass's coefficient on q_j is alpha_j
    for (int pass = 0; pass < 2; ++pass) {
      for (int v = warp; v <= j + 1; v += nwarps) {  // v = 0: u1, v = 1..j+1: q_{v-1}
        const float* qv = Q + (size_t)v * n;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // four independent chains: the loop is bound by the L2 latency of Q
        int i = lane;
        for (; i + 96 < n; i += 128) {
          s0 += w[i] * qv[i]; s1 += w[i + 32] * qv[i + 32]; s2 += w[i + 64] * qv[i + 64]; s3 += w[i + 96] * qv[i + 96];
        }
        for (; i < n; i += 32) s0 += w[i] * qv[i];
        float s = warp_sum_f32((s0 + s1) + (s2 + s3));
        if (lane == 0) s_coef[v] = s;
      }
      __syncthreads();
      if (pass == 0 && tid == 0) s_alpha[j] = (double)s_coef[j + 1];
      else if (pass == 1 && tid == 0) s_alpha[j] += (double)s_coef[j + 1];
      for (int i0 = tid; i0 < n; i0 += 4 * kLanczosThreads) {   // four elements per thread and four vectors per trip in flight
        const int i1 = i0 + kLanczosThreads, i2 = i0 + 2 * kLanczosThreads, i3 = i0 + 3 * kLanczosThreads;
        float v0 = w[i0], v1 = i1 < n ? w[i1] : 0.f, v2 = i2 < n ? w[i2] : 0.f, v3 = i3 < n ? w[i3] : 0.f;
#pragma unroll 4
        for (int t = 0; t <= j + 1; ++t) {
          const float c = s_coef[t];
          const float* q = Q + (size_t)t * n;
          v0 -= c * q[i0];
          if (i1 < n) v1 -= c * q[i1];
          if (i2 < n) v2 -= c * q[i2];
          if (i3 < n) v3 -= c * q[i3];
        }
        w[i0] = v0;
        if (i1 < n) w[i1] = v1;
        if (i2 < n) w[i2] = v2;
        if (i3 < n) w[i3] = v3;
      }
      __syncthreads();
    }
    float a2 = 0.f;
    for (int i = tid; i < n; i += kLanczosThreads) a2 += w[i] * w[i];
    const float beta = sqrtf(block_sum(a2));
    if (tid == 0) s_beta[j] = (double)beta;
    if (beta < 1e-6f || j == m - 1) { meff = j + 1; break; }
    for (int i = tid; i < n; i += kLanczosThreads) w[i] /= beta;
    __syncthreads();
  }
  __syncthreads();
  // ---- eigenpairs of the tridiagonal matrix
  const int nev = tridiag_eigs(meff, dim, s_alpha, s_beta, s_lam, s_vec, tid, lane, warp);
  if (tid == 0) {
    for (int a = 0; a < kMaxDim; ++a) P.evals[((size_t)p * P.maxcomp + c) * kMaxDim + a] = a < nev ? (float)s_lam[a] : 0.f;
    if (dim < kMaxDim) {   // last slot: the largest Ritz residual estimate |beta_m * s_m| of the returned pairs (what ARPACK iterates on)
      double rmax = 0.0;
      for (int a = 0; a < nev; ++a) rmax = fmax(rmax, fabs(s_beta[meff - 1] * s_vec[a][meff - 1]));
      P.evals[((size_t)p * P.maxcomp + c) * kMaxDim + kMaxDim - 1] = (float)rmax;
    }
    s_m = meff;
  }
  __syncthreads();
  // Ritz vectors
  float* out = P.out + (size_t)p * n * dim;
  for (int i = tid; i < n; i += kLanczosThreads) {
    if (comp[i] != c) continue;
    for (int a = 0; a < dim; ++a) {
      float v = 0.f;
      if (a < nev)
        for (int t = 0; t < meff; ++t) v += (float)s_vec[a][t] * q0[(size_t)t * n + i];
      else  // fewer Ritz pairs than dimensions (tiny component): fill with small hash noise
        v = ((float)(hash32(P.seed * 31 + (uint64_t)i * 7 + a) >> 8) * (1.f / 8388608.f) - 1.f) * 1e-3f;
      out[(size_t)i * dim + a] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Lanczos for CONNECTED graphs on a thread-block cluster (the usual case: one component = the whole cloud).  The vertex rows are
// split over the C CTAs of a cluster; each CTA keeps its rows of ALL Krylov vectors and its rows of the matrix (CSR, sorted by
// column: the products are summed in a fixed order, so the result is reproducible) in shared memory, plus the whole current
// Lanczos vector.  Scalar products are reduced through distributed shared memory: every CTA writes its partial sums into every
// CTA's table and, after one cluster barrier, adds them up in rank order (same bits everywhere).  Four cluster barriers per
// Lanczos step instead of ~100 L2 round trips per thread.
constexpr int kLcThreads = 512;
constexpr int kLcMaxCluster = 8;
__device__ __forceinline__ uint32_t lc_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t lc_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void lc_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
template <typename T>
__device__ __forceinline__ T* lc_map(T* p, uint32_t rank) {
  uint64_t out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((uint64_t)p), "r"(rank));
  return reinterpret_cast<T*>(out);
}
struct LanczosClusterParams {
  const int* head; const int* tail; const float* weight; const float* eps; int slots; int n; int dim;
  const float* deg; float* out; float* evals; uint64_t seed;
  int rows_per; int ent_cap;   // rows of a CTA (ceil(n / C)), entries of its CSR slice that fit in shared memory
  const int* ncomp;            // optional [batch] with comp [batch,n], csize [batch,n]: component blockIdx.y of every cloud is laid out
  const int* comp; const int* csize; int min_size_multi;   // (components smaller than min_size_multi of a multi-component cloud are skipped)
  int2* gent;                  // [batch, Gy, C, slots] global room for a CSR slice that does not fit in shared memory (rare: hubs)
  int maxcomp;                 // clouds with more components are skipped; stride of evals
  int debug;                   // option spectral_debug: cloud 0 prints the cycles of the phases of its first component
};
// dynamic shared memory: Qloc[(kMaxKrylov + 2) * rows_per] | qfull[n] | w[rows_per] | part[3][C][kMaxKrylov + 2] | coef[kMaxKrylov + 2]
//                        | roff[rows_per + 1] | rcnt[rows_per] | ent[ent_cap] (int2: column, value bits)
__device__ __forceinline__ void lc_component(const LanczosClusterParams& P, const int p, const int c, const int n) {
  extern __shared__ __align__(16) unsigned char lc_raw[];
  __shared__ double s_alpha[kMaxKrylov], s_beta[kMaxKrylov];
  __shared__ double s_lam[kMaxDim];
  __shared__ double s_vec[kMaxDim][kMaxKrylov];
  __shared__ float s_red[kLcThreads / 32];
  __shared__ int s_cnt;
  const uint32_t C = lc_size(), cr = lc_rank();
  const int nfull = P.n, dim = P.dim, RP = P.rows_per;   // n = vertices of the component = size of the problem
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kLcThreads / 32;
  const int rp = (n + (int)C - 1) / (int)C;   // rows per CTA of THIS component (<= RP, the allocation)
  const int r0 = min(n, (int)cr * rp), r1 = min(n, r0 + rp), nr = r1 - r0;
  constexpr int NV = kMaxKrylov + 2;
  float* Qloc = reinterpret_cast<float*>(lc_raw);
  float* qfull = Qloc + (size_t)NV * RP;
  float* w = qfull + nfull;
  float* part = w + RP;                       // [3][kLcMaxCluster][NV]
  float* coef = part + 3 * kLcMaxCluster * NV;
  int* roff = reinterpret_cast<int*>(coef + NV);
  int* rcnt = roff + RP + 1;
  int* lidx = rcnt + RP;                      // [nfull] vertex -> index inside the component (-1 outside)
  int* glob = lidx + nfull;                   // [RP] this CTA's rows -> vertex
  int2* ent = reinterpret_cast<int2*>((reinterpret_cast<uintptr_t>(glob + RP) + 7) & ~(uintptr_t)7);
  const int* head = P.head + (size_t)p * P.slots;
  const int* tail = P.tail + (size_t)p * P.slots;
  const float* weight = P.weight + (size_t)p * P.slots;
  const float* eps = P.eps + (size_t)p * P.slots;
  const float* deg = P.deg + (size_t)p * nfull;
  const int m = min(kMaxKrylov, n - 1);
  // vertex -> index inside the component (order preserving), and the vertices of this CTA's rows
  if (P.ncomp) {
    const int* comp = P.comp + (size_t)p * nfull;
    __shared__ int s_carry;
    __shared__ int s_wsum[kLcThreads / 32];
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nfull; b0 += kLcThreads) {
      const int i = b0 + tid;
      const bool in = i < nfull && comp[i] == c;
      const unsigned bal = __ballot_sync(0xffffffffu, in);
      if (lane == 0) s_wsum[warp] = __popc(bal);
      __syncthreads();
      int off = s_carry;
      for (int q = 0; q < warp; ++q) off += s_wsum[q];
      if (i < nfull) lidx[i] = in ? off + __popc(bal & ((1u << lane) - 1)) : -1;
      __syncthreads();
      if (tid == 0) { int t = 0; for (int q = 0; q < nwarps; ++q) t += s_wsum[q]; s_carry += t; }
      __syncthreads();
    }
  } else {
    for (int i = tid; i < nfull; i += kLcThreads) lidx[i] = i;
    __syncthreads();
  }
  for (int i = tid; i < nfull; i += kLcThreads) { const int l = lidx[i]; if (l >= r0 && l < r1) glob[l - r0] = i; }
  __syncthreads();

  auto block_sum = [&](float v) -> float {
    v = warp_sum_f32(v);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int q = 0; q < nwarps; ++q) t += s_red[q];
    __syncthreads();
    return t;
  };
  // cluster-wide sums of `cnt` values held in coef[0..cnt) as this CTA's partial sums: into coef[0..cnt), same bits in every CTA
  int red_phase = 0;
  auto cluster_reduce = [&](int cnt) {
    float* mine = part + (size_t)red_phase * kLcMaxCluster * NV;
    __syncthreads();
    for (uint32_t r = 0; r < C; ++r) {
      float* dst = lc_map(mine, r) + (size_t)cr * NV;
      for (int v = tid; v < cnt; v += kLcThreads) dst[v] = coef[v];
    }
    lc_sync();
    for (int v = tid; v < cnt; v += kLcThreads) {
      float t = 0.f;
      for (uint32_t r = 0; r < C; ++r) t += mine[(size_t)r * NV + v];
      coef[v] = t;
    }
    __syncthreads();
    red_phase = (red_phase + 1) % 3;
  };
  // every CTA's rows of a vector (local array `src`, nr entries) into every CTA's qfull
  auto broadcast_rows = [&](const float* src) {
    for (uint32_t r = 0; r < C; ++r) {
      float* dst = lc_map(qfull, r) + r0;
      for (int i = tid; i < nr; i += kLcThreads) dst[i] = src[i];
    }
    lc_sync();
  };

  // ---- this CTA's rows of A = D^-1/2 W D^-1/2 as CSR, columns ascending
  for (int i = tid; i <= RP; i += kLcThreads) { roff[i] = 0; if (i < RP) rcnt[i] = 0; }
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  // (both passes over the slot table are chains of dependent loads: four slots per thread in flight)
  for (int e0 = tid; e0 < P.slots; e0 += 4 * kLcThreads) {
    float ep4[4];
    int hg4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kLcThreads;
      ep4[u] = e < P.slots ? eps[e] : 0.f;
      hg4[u] = e < P.slots ? head[e] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ep4[u] > 0.f) { const int h = lidx[hg4[u]]; if (h >= r0 && h < r1) atomicAdd(&rcnt[h - r0], 1); }
  }
  __syncthreads();
  if (tid == 0) { int acc = 0; for (int i = 0; i < nr; ++i) { roff[i] = acc; acc += rcnt[i]; } roff[nr] = acc; s_cnt = acc; }
  __syncthreads();
  const int nent = s_cnt;
  if (nent > P.ent_cap) ent = P.gent + (((size_t)p * gridDim.y + blockIdx.y) * C + cr) * (size_t)P.slots;   // the slice lives in global memory instead
  for (int i = tid; i < nr; i += kLcThreads) rcnt[i] = roff[i];
  __syncthreads();
  {
    for (int e0 = tid; e0 < P.slots; e0 += 4 * kLcThreads) {
      float ep4[4];
      int hg4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * kLcThreads;
        ep4[u] = e < P.slots ? eps[e] : 0.f;
        hg4[u] = e < P.slots ? head[e] : 0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ep4[u] > 0.f) {
          const int e = e0 + u * kLcThreads;
          const int hg = hg4[u], h = lidx[hg];
          if (h >= r0 && h < r1) {
            const int tg = tail[e];
            const int pos = atomicAdd(&rcnt[h - r0], 1);
            ent[pos] = make_int2(lidx[tg], __float_as_int(weight[e] * rsqrtf(deg[hg] * deg[tg])));
          }
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < nr; i += kLcThreads) {   // the fill order depends on scheduling: sort every row by column
      const int a0 = roff[i], a1 = roff[i + 1];
      for (int q = a0 + 1; q < a1; ++q) {
        const int2 x = ent[q];
        int k = q - 1;
        while (k >= a0 && ent[k].x > x.x) { ent[k + 1] = ent[k]; --k; }
        ent[k + 1] = x;
      }
    }
  }
  __syncthreads();
  lc_sync();   // every CTA of the cluster is up (its buffers may be written from now on)

  // ---- u1 = D^1/2 1 normalised (vector 0), q_0 = random, orthogonal to u1, normalised (vector 1)
  float* u1 = Qloc;
  float* q0 = Qloc + RP;
  {
    float acc = 0.f;
    for (int i = tid; i < nr; i += kLcThreads) { const float v = sqrtf(deg[glob[i]]); u1[i] = v; acc += v * v; }
    const float s = block_sum(acc);
    if (tid == 0) coef[0] = s;
    cluster_reduce(1);
    const float nrm = sqrtf(coef[0]);
    __syncthreads();
    acc = 0.f;
    for (int i = tid; i < nr; i += kLcThreads) {
      const float u = u1[i] / nrm;
      u1[i] = u;
      const float r = (float)(hash32(P.seed + (uint64_t)glob[i]) >> 8) * (1.f / 8388608.f) - 1.f;
      q0[i] = r;
      acc += r * u;
    }
    const float s2 = block_sum(acc);
    if (tid == 0) coef[0] = s2;
    cluster_reduce(1);
    const float d0 = coef[0];
    __syncthreads();
    acc = 0.f;
    for (int i = tid; i < nr; i += kLcThreads) { const float r = q0[i] - d0 * u1[i]; q0[i] = r; acc += r * r; }
    const float s3 = block_sum(acc);
    if (tid == 0) coef[0] = s3;
    cluster_reduce(1);
    const float nrm2 = sqrtf(coef[0]);
    __syncthreads();
    for (int i = tid; i < nr; i += kLcThreads) q0[i] /= nrm2;
    __syncthreads();
  }
  int meff = m;
  long long cyc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tq = clock64();
  const long long t_setup = tq;
#define TDA_LC_T(k) do { if (P.debug) { const long long tn_ = clock64(); cyc[k] += tn_ - tq; tq = tn_; } } while (0)
  for (int j = 0; j < m; ++j) {
    const float* qj = Qloc + (size_t)(j + 1) * RP;
    float* wn = Qloc + (size_t)(j + 2) * RP;   // becomes q_{j+1}
    broadcast_rows(qj);
    TDA_LC_T(0);
    // w = A q_j on this CTA's rows (one thread per row, entries in column order)
    for (int i = tid; i < nr; i += kLcThreads) {
      float acc = 0.f;
      for (int q = roff[i]; q < roff[i + 1]; ++q) { const int2 E = ent[q]; acc += __int_as_float(E.y) * qfull[E.x]; }
      w[i] = acc;
    }
    __syncthreads();
    TDA_LC_T(1);
    // classical Gram-Schmidt, twice, against u1 and q_0..q_j; the coefficient on q_j is alpha_j
    for (int pass = 0; pass < 2; ++pass) {
      // four vectors per warp and trip: they share the loads of w and their four butterflies interleave (each sum is taken in
      // the same order as a one-vector-at-a-time loop: same bits)
      for (int vb = warp * 4; vb <= j + 1; vb += nwarps * 4) {
        const int nv = min(4, j + 2 - vb);
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = lane; i < nr; i += 32) {
          const float wi = w[i];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (u < nv) s4[u] += wi * Qloc[(size_t)(vb + u) * RP + i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int u = 0; u < 4; ++u) s4[u] += __shfl_xor_sync(0xffffffffu, s4[u], o);
        }
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (u < nv) coef[vb + u] = s4[u];
        }
      }
      TDA_LC_T(2);
      cluster_reduce(j + 2);
      TDA_LC_T(3);
      if (tid == 0) { if (pass == 0) s_alpha[j] = (double)coef[j + 1]; else s_alpha[j] += (double)coef[j + 1]; }
      for (int i = tid; i < nr; i += kLcThreads) {
        float v0 = w[i];
        for (int t = 0; t <= j + 1; ++t) v0 -= coef[t] * Qloc[(size_t)t * RP + i];
        w[i] = v0;
      }
      __syncthreads();
      TDA_LC_T(4);
    }
    float a2 = 0.f;
    for (int i = tid; i < nr; i += kLcThreads) a2 += w[i] * w[i];
    const float sb = block_sum(a2);
    if (tid == 0) coef[0] = sb;
    cluster_reduce(1);
    const float beta = sqrtf(coef[0]);
    __syncthreads();
    if (tid == 0) s_beta[j] = (double)beta;
    if (beta < 1e-6f || j == m - 1) { meff = j + 1; break; }   // (same decision in every CTA: same bits)
    for (int i = tid; i < nr; i += kLcThreads) wn[i] = w[i] / beta;
    __syncthreads();
    TDA_LC_T(5);
  }
  __syncthreads();
  const int nev = tridiag_eigs(meff, dim, s_alpha, s_beta, s_lam, s_vec, tid, lane, warp);
  TDA_LC_T(6);
  if (cr == 0 && tid == 0 && P.evals) {
    for (int a = 0; a < kMaxDim; ++a) P.evals[((size_t)p * P.maxcomp + c) * kMaxDim + a] = a < nev ? (float)s_lam[a] : 0.f;
    if (dim < kMaxDim) {   // last slot: the largest Ritz residual estimate |beta_m * s_m| of the returned pairs
      double rmax = 0.0;
      for (int a = 0; a < nev; ++a) rmax = fmax(rmax, fabs(s_beta[meff - 1] * s_vec[a][meff - 1]));
      P.evals[((size_t)p * P.maxcomp + c) * kMaxDim + kMaxDim - 1] = (float)rmax;
    }
  }
  float* out = P.out + (size_t)p * nfull * dim;
  for (int i = tid; i < nr; i += kLcThreads) {
    for (int a = 0; a < dim; ++a) {
      float v = 0.f;
      if (a < nev)
        for (int t = 0; t < meff; ++t) v += (float)s_vec[a][t] * Qloc[(size_t)(t + 1) * RP + i];
      else
        v = ((float)(hash32(P.seed * 31 + (uint64_t)glob[i] * 7 + a) >> 8) * (1.f / 8388608.f) - 1.f) * 1e-3f;
      out[(size_t)glob[i] * dim + a] = v;
    }
  }
  TDA_LC_T(7);
  if (P.debug && p == 0 && blockIdx.y == 0 && tid == 0)
    printf("lanczos cloud 0 comp %d cta %u: n=%d nr=%d meff=%d nent=%d | cycles: broadcast %lld spmv %lld dots %lld reduce %lld update %lld norm %lld tridiag %lld ritz %lld | since loop start %lld\n",
           c, cr, n, nr, meff, nent, cyc[0], cyc[1], cyc[2], cyc[3], cyc[4], cyc[5], cyc[6], cyc[7], clock64() - t_setup);
#undef TDA_LC_T
  lc_sync();   // nobody leaves (or starts the next component) while its shared memory may still be written
}
// grid: (batch * C, Gy) CTAs in clusters of C.  Cluster (p, y) lays out the components y, y + Gy, ... of cloud p one after the other
// (every CTA of a cluster takes the same decisions: they depend on p and c only).
__global__ void __launch_bounds__(kLcThreads, 1) lanczos_cluster_kernel(LanczosClusterParams P) {
  const int p = blockIdx.x / (int)lc_size();
  if (!P.ncomp) {             // the caller knows the graphs are connected
    if (blockIdx.y == 0) lc_component(P, p, 0, P.n);
    return;
  }
  const int ncp = P.ncomp[p];
  if (ncp > P.maxcomp) return;   // left to the caller (status 1)
  for (int c = blockIdx.y; c < ncp; c += gridDim.y) {
    const int nc = P.csize[(size_t)p * P.n + c];
    if (ncp > 1 && nc < P.min_size_multi) continue;   // too small for a spectral layout: random points (multi_component_kernel)
    lc_component(P, p, c, nc);
  }
}

// ------------------------------------------------------------------------------------------------
// component_layout (umap-learn spectral.py) for clouds with MORE than 2*dim components, on the device: the meta positions are the
// spectral embedding (sklearn SpectralEmbedding, affinity = exp(-d^2), normalised Laplacian, first eigenvector dropped, vectors
// divided by sqrt(degree), deterministic sign flip) of the component centroids in DATA space, divided by their largest entry.
constexpr int kMetaMaxComp = 32;     // clouds with more components go back to the caller (status 1)
constexpr int kCentThreads = 128;
// centroids: one thread per data column, the points in order (a fixed summation order: reproducible).  grid (ceil(d/128), batch)
__global__ void __launch_bounds__(kCentThreads) centroid_kernel(const float* __restrict__ X_g, int n, int d, const int* __restrict__ comp_g,
                                                                const int* __restrict__ ncomp_g, const int* __restrict__ csize_g, int dim, int maxcomp,
                                                                float* __restrict__ cent_g) {
  __shared__ float s_acc[kMetaMaxComp][kCentThreads];
  const int p = blockIdx.y, tid = threadIdx.x;
  const int nc = ncomp_g[p];
  if (nc <= 2 * dim || nc > maxcomp) return;
  const int col = blockIdx.x * kCentThreads + tid;
  const int* comp = comp_g + (size_t)p * n;
  const float* X = X_g + (size_t)p * n * d;
  for (int c = 0; c < nc; ++c) s_acc[c][tid] = 0.f;
  if (col < d) {
    int i = 0;
    for (; i + 4 <= n; i += 4) {   // four rows in flight
      const float x0 = __ldg(&X[(size_t)i * d + col]), x1 = __ldg(&X[(size_t)(i + 1) * d + col]);
      const float x2 = __ldg(&X[(size_t)(i + 2) * d + col]), x3 = __ldg(&X[(size_t)(i + 3) * d + col]);
      s_acc[comp[i]][tid] += x0; s_acc[comp[i + 1]][tid] += x1; s_acc[comp[i + 2]][tid] += x2; s_acc[comp[i + 3]][tid] += x3;
    }
    for (; i < n; ++i) s_acc[comp[i]][tid] += __ldg(&X[(size_t)i * d + col]);
    for (int c = 0; c < nc; ++c)
      cent_g[((size_t)p * kMetaMaxComp + c) * d + col] = s_acc[c][tid] / (float)max(1, csize_g[(size_t)p * n + c]);
  }
}
// meta positions: one CTA (256 threads) per cloud.  metric: 0 sqeuclidean, 1 euclidean, 2 cosine (pdist.cu's numbering).
__global__ void __launch_bounds__(256) meta_layout_kernel(const float* __restrict__ cent_g, int d, const int* __restrict__ ncomp_g, int dim, int maxcomp,
                                                          int metric, float* __restrict__ meta_g) {
  __shared__ double s_A[kMetaMaxComp][kMetaMaxComp + 1];   // distance -> affinity -> Laplacian -> (diagonalised)
  __shared__ double s_V[kMetaMaxComp][kMetaMaxComp + 1];   // eigenvectors (columns)
  __shared__ double s_isd[kMetaMaxComp];
  __shared__ int s_order[kMetaMaxComp];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int nc = ncomp_g[p];
  if (nc <= 2 * dim || nc > maxcomp) return;
  const float* cent = cent_g + (size_t)p * kMetaMaxComp * d;
  // Gram matrix of the centroids in fp64: one warp per pair
  for (int pr = warp; pr < nc * (nc + 1) / 2; pr += nwarps) {
    int a = 0, rem = pr;
    while (rem > a) { rem -= a + 1; ++a; }   // pr = a(a+1)/2 + b, b <= a
    const int b = rem;
    double acc = 0.0;
    for (int j = lane; j < d; j += 32) acc += (double)cent[(size_t)a * d + j] * (double)cent[(size_t)b * d + j];
    acc = warp_sum_f64(acc);
    if (lane == 0) { s_V[a][b] = acc; s_V[b][a] = acc; }
  }
  __syncthreads();
  for (int e = tid; e < nc * nc; e += blockDim.x) {
    const int a = e / nc, b = e % nc;
    const double gaa = s_V[a][a], gbb = s_V[b][b], gab = s_V[a][b];
    double dist;
    if (metric == 2) {   // sklearn cosine_distances: 1 - normalised dot, clipped to [0, 2], exact zero diagonal
      const double den = sqrt(gaa) * sqrt(gbb);
      dist = den > 0.0 ? 1.0 - gab / den : 1.0;
      dist = fmin(fmax(dist, 0.0), 2.0);
    } else {
      const double d2 = fmax(gaa + gbb - 2.0 * gab, 0.0);
      dist = metric == 0 ? d2 : sqrt(d2);
    }
    s_A[a][b] = a == b ? 0.0 : exp(-dist * dist);
  }
  __syncthreads();
  if (tid < nc) {
    double deg = 0.0;
    for (int b = 0; b < nc; ++b) deg += s_A[tid][b];
    s_isd[tid] = 1.0 / sqrt(fmax(deg, 1e-300));
  }
  __syncthreads();
  for (int e = tid; e < nc * nc; e += blockDim.x) {
    const int a = e / nc, b = e % nc;
    s_A[a][b] = (a == b ? 1.0 : 0.0) - s_isd[a] * s_A[a][b] * s_isd[b];
    s_V[a][b] = a == b ? 1.0 : 0.0;
  }
  __syncthreads();
  if (warp == 0) {   // cyclic Jacobi, lane k owns row / column entry k of the two rotated lines
    for (int sweep = 0; sweep < 60; ++sweep) {
      double off = 0.0;
      for (int a = lane; a < nc; a += 32)
        for (int b = 0; b < nc; ++b) if (b != a) off += s_A[a][b] * s_A[a][b];
      off = warp_sum_f64(off);
      off = __shfl_sync(0xffffffffu, off, 0);
      if (off < 1e-28) break;
      for (int a = 0; a < nc - 1; ++a)
        for (int b = a + 1; b < nc; ++b) {
          const double apq = s_A[a][b];
          if (fabs(apq) < 1e-300) continue;   // (uniform: every lane reads the same entry)
          const double theta = (s_A[b][b] - s_A[a][a]) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
          __syncwarp();
          if (lane < nc) {   // columns a, b
            const int k = lane;
            const double xa = s_A[k][a], xb = s_A[k][b];
            s_A[k][a] = cs * xa - sn * xb; s_A[k][b] = sn * xa + cs * xb;
            const double va = s_V[k][a], vb = s_V[k][b];
            s_V[k][a] = cs * va - sn * vb; s_V[k][b] = sn * va + cs * vb;
          }
          __syncwarp();
          if (lane < nc) {   // rows a, b
            const int k = lane;
            const double xa = s_A[a][k], xb = s_A[b][k];
            s_A[a][k] = cs * xa - sn * xb; s_A[b][k] = sn * xa + cs * xb;
          }
          __syncwarp();
        }
    }
    if (lane == 0) {   // eigenvalues ascending (ties: lower index first)
      for (int a = 0; a < nc; ++a) s_order[a] = a;
      for (int a = 1; a < nc; ++a) {
        const int x = s_order[a];
        int k = a - 1;
        while (k >= 0 && s_A[s_order[k]][s_order[k]] > s_A[x][x]) { s_order[k + 1] = s_order[k]; --k; }
        s_order[k + 1] = x;
      }
    }
  }
  __syncthreads();
  // embedding column q = eigenvector q+1 / sqrt(degree), sign: the entry of largest magnitude positive; all divided by the largest entry
  __shared__ double s_sign[kMaxDim], s_max;
  if (tid < dim) {
    double best = 0.0;
    if (tid + 1 < nc) {
      const int col = s_order[tid + 1];
      for (int a = 0; a < nc; ++a) { const double v = s_V[a][col] * s_isd[a]; if (fabs(v) > fabs(best)) best = v; }
    }
    s_sign[tid] = best < 0.0 ? -1.0 : 1.0;
  }
  __syncthreads();
  if (tid == 0) {
    double mx = 0.0;
    for (int q = 0; q < dim && q + 1 < nc; ++q)
      for (int a = 0; a < nc; ++a) mx = fmax(mx, s_sign[q] * s_V[a][s_order[q + 1]] * s_isd[a]);
    s_max = mx > 0.0 ? mx : 1.0;
  }
  __syncthreads();
  for (int e = tid; e < nc * dim; e += blockDim.x) {
    const int a = e / dim, q = e % dim;
    const double v = q + 1 < nc ? s_sign[q] * s_V[a][s_order[q + 1]] * s_isd[a] / s_max : 0.0;
    meta_g[((size_t)p * kMetaMaxComp + a) * kMaxDim + q] = (float)v;
  }
}

// ------------------------------------------------------------------------------------------------
// multi_component_layout (umap-learn spectral.py) on the device: the components are placed around their meta positions (+-e_k for
// up to 2*dim components, else the component_layout above, read from meta_g), each scaled to half the distance to the nearest
// other meta position; components too small for a spectral layout get uniform random points of that range.  One CTA per cloud.
// status[p]: 0 done, 1 = more than `maxcomp` components (or more than 2*dim and no meta layout: meta_g == nullptr): the caller's job.
__global__ void __launch_bounds__(256) multi_component_kernel(const int* __restrict__ comp_g, const int* __restrict__ ncomp_g, const int* __restrict__ csize_g,
                                                              int n, int dim, int maxcomp, int min_size, uint64_t seed, const float* __restrict__ meta_g,
                                                              float* __restrict__ Y_g, int* __restrict__ status) {
  __shared__ float s_meta[kMetaMaxComp][kMaxDim];
  __shared__ float s_range[kMetaMaxComp];
  __shared__ unsigned int s_amax[kMetaMaxComp];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int nc = ncomp_g[p];
  const bool todo = nc <= 2 * dim ? nc <= maxcomp : (meta_g != nullptr && nc <= maxcomp && nc <= kMetaMaxComp);
  if (tid == 0) status[p] = todo ? 0 : 1;
  if (nc == 1 || !todo) return;
  const int* comp = comp_g + (size_t)p * n;
  const int* csize = csize_g + (size_t)p * n;
  float* Y = Y_g + (size_t)p * n * dim;
  if (tid < nc) {
    if (nc <= 2 * dim) {
      const int k = (nc + 1) / 2;
      for (int a = 0; a < dim; ++a) s_meta[tid][a] = 0.f;
      if (tid < k) s_meta[tid][tid] = 1.f; else s_meta[tid][tid - k] = -1.f;
    } else {
      for (int a = 0; a < dim; ++a) s_meta[tid][a] = meta_g[((size_t)p * kMetaMaxComp + tid) * kMaxDim + a];
    }
    s_amax[tid] = 0u;
  }
  __syncthreads();
  if (tid < nc) {
    float best = INFINITY;
    for (int o = 0; o < nc; ++o) {
      float d2 = 0.f;
      for (int a = 0; a < dim; ++a) { const float t = s_meta[tid][a] - s_meta[o][a]; d2 += t * t; }
      const float d = sqrtf(d2);
      if (d > 0.f && d < best) best = d;
    }
    s_range[tid] = isfinite(best) ? best * 0.5f : 1.f;
  }
  for (int i = tid; i < n; i += blockDim.x) {
    float m = 0.f;
    for (int a = 0; a < dim; ++a) m = fmaxf(m, fabsf(Y[(size_t)i * dim + a]));
    atomicMax(&s_amax[comp[i]], __float_as_uint(m));
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    const int c = comp[i];
    const float r = s_range[c];
    if (csize[c] < min_size) {
      for (int a = 0; a < dim; ++a) {
        const float u = (float)(hash32(seed * 0x9E3779B97F4A7C15ull + ((uint64_t)i << 8) + (uint64_t)a + 7919ull) >> 8) * (1.f / 8388608.f) - 1.f;
        Y[(size_t)i * dim + a] = u * r + s_meta[c][a];
      }
    } else {
      const float sc = r / fmaxf(__uint_as_float(s_amax[c]), 1e-30f);
      for (int a = 0; a < dim; ++a) Y[(size_t)i * dim + a] = Y[(size_t)i * dim + a] * sc + s_meta[c][a];
    }
  }
}

struct Layout {
  int *label, *comp, *ncomp, *csize;
  float *deg, *Q, *evals, *cent, *meta;
  int4* entries; int* ent_count; unsigned long long* wfix; int2* lc_ent;
  size_t total;
};
constexpr int kLcGridY = 8;   // clusters per cloud of the per-component launch (components c, c + 8, ... one after the other)
// n_lanczos: components per cloud the one-CTA Lanczos kernel needs room for (0: the cluster kernel does every component);
// n_cluster: cluster slots per cloud with spill room for a CSR slice;  d > 0: room for the component_layout of clouds with more
// than 2*dim components (centroids in data space)
static Layout make_layout(void* ws, int n, int batch, int maxcomp, int slots, int n_lanczos, int n_cluster, int d) {
  Layout L;
  Carver c(ws, ~size_t(0));
  const size_t sl = slots > 0 ? (size_t)slots : 0;
  L.label = c.take<int>((size_t)batch * n);
  L.comp = c.take<int>((size_t)batch * n);
  L.csize = c.take<int>((size_t)batch * n);
  L.ncomp = c.take<int>(batch);
  L.deg = c.take<float>((size_t)batch * n);
  L.evals = c.take<float>((size_t)batch * (maxcomp > 0 ? maxcomp : 1) * kMaxDim);
  L.Q = c.take<float>((size_t)batch * n_lanczos * (kMaxKrylov + 2) * n);
  L.entries = c.take<int4>(n_lanczos > 0 ? (size_t)batch * sl : 0);
  L.ent_count = c.take<int>(batch);
  L.wfix = c.take<unsigned long long>((size_t)batch * n_lanczos * n);
  L.lc_ent = c.take<int2>((size_t)batch * n_cluster * kLcMaxCluster * sl);
  L.cent = c.take<float>(d > 0 ? (size_t)batch * kMetaMaxComp * d : 0);
  L.meta = c.take<float>(d > 0 ? (size_t)batch * kMetaMaxComp * kMaxDim : 0);
  L.total = c.off;
  return L;
}

}  // namespace spectral
}  // namespace tda

using namespace tda;
using namespace tda::spectral;

// shape of a lanczos_cluster_kernel launch for clouds of n vertices; false if the option is off or the rows do not fit in shared memory
struct ClusterShape { int C, RP, cap_ent; size_t dyn; };
static bool cluster_shape(int n, int slots, ClusterShape& S) {
  if (option("spectral_cluster") == 0 || n < 64) return false;
  int C = (int)option("spectral_cluster");
  if (C != 2 && C != 4 && C != 8) C = 8;
  const int RP = (n + C - 1) / C;
  constexpr int NV = kMaxKrylov + 2;
  const size_t fixed = sizeof(float) * ((size_t)NV * RP + (size_t)n + (size_t)RP + 3 * (size_t)kLcMaxCluster * NV + NV) +
                       sizeof(int) * ((size_t)(RP + 1) + RP + (size_t)n + RP + 4);
  const size_t smem_max = (size_t)206 * 1024;   // (the kernel has ~18 KB of static shared memory: Lanczos scalars + the tridiagonal LU)
  const size_t want_ent = (size_t)slots / C + (size_t)slots / (2 * C) + 64;   // 1.5x the mean slice
  if (fixed + 8 * 1024 >= smem_max) return false;
  size_t cap_ent = (smem_max - fixed) / sizeof(int2);
  if (cap_ent > want_ent) cap_ent = want_ent;
  S.C = C; S.RP = RP; S.cap_ent = (int)cap_ent; S.dyn = fixed + cap_ent * sizeof(int2);
  return true;
}
// launches lanczos_cluster_kernel for the clouds whose graph is connected (ncomp == nullptr: all of them, one component each), or
// for every component of every cloud with up to `maxcomp` components (grid.y = min(maxcomp, kLcGridY) clusters per cloud).
// Returns TDA_OK (launched), 1 (not applicable: option off / rows do not fit in shared memory), or an error code.
static int launch_cluster_lanczos(const int32_t* head, const int32_t* tail, const float* weight, const float* eps, int slots, int n, int dim, int batch,
                                  const float* degree, const int* ncomp, const int* comp, const int* csize, int maxcomp, int min_size_multi,
                                  uint64_t seed, float* Y, float* evals, int2* gent, cudaStream_t stream) {
  ClusterShape S;
  if (!cluster_shape(n, slots, S)) return 1;
  const int C = S.C;
  LanczosClusterParams Q;
  Q.head = head; Q.tail = tail; Q.weight = weight; Q.eps = eps; Q.slots = slots; Q.n = n; Q.dim = dim;
  Q.deg = degree; Q.out = Y; Q.evals = evals; Q.seed = seed; Q.rows_per = S.RP; Q.ent_cap = S.cap_ent;
  Q.gent = gent; Q.ncomp = ncomp; Q.comp = comp; Q.csize = csize; Q.min_size_multi = min_size_multi; Q.maxcomp = ncomp ? maxcomp : 1;
  Q.debug = (int)option("spectral_debug");
  TDA_CUDA_CHECK(cudaFuncSetAttribute(lanczos_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.dyn));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(batch * C), (unsigned)(ncomp ? (maxcomp < kLcGridY ? maxcomp : kLcGridY) : 1), 1);
  cfg.blockDim = dim3(kLcThreads, 1, 1);
  cfg.dynamicSmemBytes = S.dyn;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  StageScope st(STAGE_SPECTRAL, stream);
  TDA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, lanczos_cluster_kernel, Q));
  count_launch();
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}

extern "C" size_t tda_spectral_workspace_bytes(int n, int batch, int maxcomp, int slots) {
  if (n <= 0 || batch <= 0 || maxcomp < 0 || slots < 0) return 0;
  return make_layout(nullptr, n, batch, maxcomp, slots, maxcomp, maxcomp == 1 ? 1 : 0, 0).total + 1024;
}

extern "C" int tda_graph_components(const int32_t* head, const int32_t* tail, const float* weight, const float* eps, int slots, int n, int batch,
                                    int32_t* comp, int32_t* ncomp, int32_t* comp_size, float* degree, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!head || !tail || !weight || !eps || !comp || !ncomp || !comp_size || !degree || !ws || n <= 0 || batch <= 0)
    return set_error(TDA_ERR_INVALID, "tda_graph_components: bad arguments");
  if (ws_bytes < 12 * (size_t)batch * n) return set_error(TDA_ERR_WORKSPACE, "tda_graph_components: workspace too small (12 * batch * n bytes)");
  if ((((uintptr_t)ws) & 7) != 0) return set_error(TDA_ERR_INVALID, "tda_graph_components: workspace must be 8-byte aligned");
  StageScope st(STAGE_SPECTRAL, stream);
  unsigned long long* degfix = (unsigned long long*)ws;
  components_kernel<<<batch, 1024, 0, stream>>>(head, tail, weight, eps, slots, n, (int*)(degfix + (size_t)batch * n), degfix, comp, degree, ncomp, comp_size);
  count_launch();
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}

extern "C" int tda_spectral_embed(const int32_t* head, const int32_t* tail, const float* weight, const float* eps, int slots, int n, int dim,
                                  int batch, const int32_t* comp, const int32_t* ncomp, const int32_t* comp_size, const float* degree,
                                  int maxcomp, int min_size, uint64_t seed, float* Y, float* evals, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!head || !tail || !weight || !eps || !comp || !ncomp || !comp_size || !degree || !Y || !ws || n <= 0 || batch <= 0 || maxcomp <= 0)
    return set_error(TDA_ERR_INVALID, "tda_spectral_embed: bad arguments");
  if (dim < 1 || dim > kMaxDim) return set_error(TDA_ERR_UNSUPPORTED, "tda_spectral_embed: dim=%d (supported 1..%d)", dim, kMaxDim);
  Layout L = make_layout(ws, n, batch, maxcomp, slots, maxcomp, maxcomp == 1 ? 1 : 0, 0);
  if (L.total > ws_bytes) return set_error(TDA_ERR_WORKSPACE, "tda_spectral_embed: workspace %zu < required %zu", ws_bytes, L.total);
  // connected graphs (maxcomp == 1, one component holding every vertex): the cluster kernel, if the rows fit in shared memory.
  // (The caller passes maxcomp = 1 either because it knows, or speculatively -- then it checks ncomp afterwards.)
  if (maxcomp == 1) {
    const int rc = launch_cluster_lanczos(head, tail, weight, eps, slots, n, dim, batch, degree, nullptr, nullptr, nullptr, 1, 0, seed, Y,
                                          evals ? evals : L.evals, L.lc_ent, stream);
    if (rc != 1) return rc;   // launched (TDA_OK) or failed; 1 = not applicable: the per-component kernel below
  }
  TDA_CUDA_CHECK(cudaMemsetAsync(L.ent_count, 0, sizeof(int) * batch, stream));
  LanczosParams P;
  P.head = head; P.tail = tail; P.weight = weight; P.eps = eps; P.slots = slots; P.n = n; P.dim = dim;
  P.comp = comp; P.deg = degree; P.ncomp = ncomp; P.csize = comp_size;
  P.Q = L.Q; P.entries = L.entries; P.ent_count = L.ent_count; P.wfix = L.wfix; P.out = Y; P.evals = evals ? evals : L.evals; P.maxcomp = maxcomp; P.min_size = min_size; P.seed = seed; P.skip_connected = 0;
  dim3 grid(maxcomp, batch);
  StageScope st(STAGE_SPECTRAL, stream);
  lanczos_kernel<<<grid, kLanczosThreads, 0, stream>>>(P);
  count_launch();
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}

// ---- the whole spectral initialisation without a host round trip
static Layout init_layout(void* ws, int n, int batch, int maxcomp, int slots, int d) {
  ClusterShape S;
  const bool cl = cluster_shape(n, slots, S);
  return make_layout(ws, n, batch, maxcomp, slots, cl ? 0 : maxcomp, cl ? (maxcomp < kLcGridY ? maxcomp : kLcGridY) : 0, d);
}
extern "C" size_t tda_spectral_init_workspace_bytes(int n, int batch, int maxcomp, int slots, int d) {
  if (n <= 0 || batch <= 0 || maxcomp <= 0 || slots < 0 || d < 0) return 0;
  if (maxcomp > kMetaMaxComp) maxcomp = kMetaMaxComp;
  return init_layout(nullptr, n, batch, maxcomp, slots, d).total + 12 * (size_t)batch * n + 4096;
}
extern "C" int tda_spectral_init(const int32_t* head, const int32_t* tail, const float* weight, const float* eps, int slots, int n, int dim,
                                 int batch, int maxcomp, uint64_t seed, const float* X, int d, int metric, float* Y, int32_t* ncomp_out,
                                 int32_t* status_out, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!head || !tail || !weight || !eps || !Y || !ncomp_out || !status_out || !ws || n <= 0 || batch <= 0 || maxcomp <= 0)
    return set_error(TDA_ERR_INVALID, "tda_spectral_init: bad arguments");
  if (dim < 1 || dim > kMaxDim) return set_error(TDA_ERR_UNSUPPORTED, "tda_spectral_init: dim=%d (supported 1..%d)", dim, kMaxDim);
  if (X && (d <= 0 || metric < 0 || metric > 2)) return set_error(TDA_ERR_INVALID, "tda_spectral_init: X given with d=%d metric=%d", d, metric);
  if (!X) d = 0;
  if (maxcomp > kMetaMaxComp) maxcomp = kMetaMaxComp;
  if (!X && maxcomp > 2 * dim) maxcomp = 2 * dim;   // no data: no component_layout
  Layout L = init_layout(ws, n, batch, maxcomp, slots, d);
  const size_t need = L.total + 12 * (size_t)batch * n + 256;
  if (need > ws_bytes) return set_error(TDA_ERR_WORKSPACE, "tda_spectral_init: workspace %zu < required %zu", ws_bytes, need);
  unsigned long long* degfix = (unsigned long long*)((char*)ws + ((L.total + 255) & ~(size_t)255));
  {
    StageScope st(STAGE_SPECTRAL, stream);
    components_kernel<<<batch, 1024, 0, stream>>>(head, tail, weight, eps, slots, n, L.label, degfix, L.comp, L.deg, ncomp_out, L.csize);
    count_launch();
    TDA_LAUNCH_CHECK();
  }
  TDA_CUDA_CHECK(cudaMemsetAsync(Y, 0, sizeof(float) * (size_t)batch * n * dim, stream));
  // every component (up to maxcomp per cloud) by a thread-block cluster; if the rows do not fit in shared memory: one CTA each
  const int min_size = 2 * dim > dim + 2 ? 2 * dim : dim + 2;
  const int rc = launch_cluster_lanczos(head, tail, weight, eps, slots, n, dim, batch, L.deg, ncomp_out, L.comp, L.csize, maxcomp, min_size, seed, Y,
                                        L.evals, L.lc_ent, stream);
  if (rc < 0) return rc;
  {
    StageScope st(STAGE_SPECTRAL, stream);
    if (rc != TDA_OK) {
      TDA_CUDA_CHECK(cudaMemsetAsync(L.ent_count, 0, sizeof(int) * batch, stream));
      LanczosParams P;
      P.head = head; P.tail = tail; P.weight = weight; P.eps = eps; P.slots = slots; P.n = n; P.dim = dim;
      P.comp = L.comp; P.deg = L.deg; P.ncomp = ncomp_out; P.csize = L.csize;
      P.Q = L.Q; P.entries = L.entries; P.ent_count = L.ent_count; P.wfix = L.wfix; P.out = Y; P.evals = L.evals; P.maxcomp = maxcomp;
      P.min_size = 1; P.seed = seed; P.skip_connected = 0;
      dim3 grid(maxcomp, batch);
      lanczos_kernel_ms<<<grid, kLanczosThreads, 0, stream>>>(P, min_size);
      count_launch();
    }
    if (X && maxcomp > 2 * dim) {   // component_layout for the clouds with more than 2*dim components (the kernels leave at once otherwise)
      centroid_kernel<<<dim3((unsigned)((d + kCentThreads - 1) / kCentThreads), (unsigned)batch), kCentThreads, 0, stream>>>(
          X, n, d, L.comp, ncomp_out, L.csize, dim, maxcomp, L.cent);
      count_launch();
      meta_layout_kernel<<<batch, 256, 0, stream>>>(L.cent, d, ncomp_out, dim, maxcomp, metric, L.meta);
      count_launch();
    }
    multi_component_kernel<<<batch, 256, 0, stream>>>(L.comp, ncomp_out, L.csize, n, dim, maxcomp, min_size, seed, X ? L.meta : nullptr, Y, status_out);
    count_launch();
    TDA_LAUNCH_CHECK();
  }
  return TDA_OK;
}
