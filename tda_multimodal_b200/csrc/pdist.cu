// pdist.cu -- pairwise distance matrix of high-dimensional activations as a tcgen05/TMEM GEMM fed by TMA.
//
// Replaces sklearn.metrics.pairwise_distances as it is reached from umap-learn's small-data path
// (metric='cosine'; debug_tda_pipeline.py:96-104, analyze_tda_over_layers.py:38-44,69,72,
// analyze_adversarial_tda.py:85-93), from ripser.py on raw 4096-d clouds (metric='euclidean'),
// torch.cdist in metrics.py:143 and the Gram matrix of metrics.py:368.
//
// Precision: the reference computes in float32 (sgemm).  bf16/tf32 single-pass MMA cannot meet the 1e-5
// parity bound, so every operand is split x = hi + lo with hi = tf32(x), lo = tf32(x - hi) and the kernel
// accumulates lo*hi + hi*lo + hi*hi (3xTF32) from the four operand tiles of every k-block.  The tensor core's fp32 accumulate truncates, which biases a 12288-term sum by ~3e-5
// (measured), so K is cut into chunks of 128 elements: each chunk is accumulated in TMEM on its own (small
// partial sums, few truncating adds) and the epilogue warps add the chunks in registers with round-to-nearest.
// Useful FLOPs = 2*N*M*D, issued FLOPs = 3x that.
//
// Kernel structure (one persistent CTA per SM, 192 threads):
//   warp 0      : TMA producer   (cp.async.bulk.tensor.2d, 128B swizzle, 3-stage mbarrier ring of 64 KB stages:
//                                 A_hi, A_lo, B_hi, B_lo of one k-block, loaded once for the three products)
//   warp 1      : MMA issuer     (tcgen05.mma.cta_group::1.kind::tf32, M=128 N=128 K=8; owns TMEM alloc)
//   warps 2..5  : epilogue       (tcgen05.ld 32x32b.x32 -> distance formula -> global), double-buffered
//                                 accumulators (2 x 128 TMEM columns) so it overlaps the next tile's K loop
#include "common.cuh"
#include "launch_count.cuh"
#include "../../include/tda_b200.h"
#include <cuda.h>
#include <cmath>

namespace tda {
namespace pdist {

constexpr int BM = 128, BN = 128, BK = 32;       // tile; BK floats = 128 bytes = one swizzle row
constexpr int kStages = 3;
constexpr int kTileABytes = BM * BK * 4;         // 16 KB
constexpr int kTileBBytes = BN * BK * 4;         // 16 KB
constexpr int kStageBytes = 2 * kTileABytes + 2 * kTileBBytes;  // 64 KB: A_hi, A_lo, B_hi, B_lo of one k-block
constexpr int kThreads = 192;
constexpr int kTmemCols = 256;                   // 2 accumulator stages x 128 columns
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 28)) __trap();
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major operand tile written by TMA with 128-byte swizzle
// (rows of 128 B, 8-row groups 1024 B apart), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address
  d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                     // version
  d |= (uint64_t)2 << 61;                     // SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct GemmParams {
  int n, m, batch;         // rows of A / rows of B per problem, number of problems
  int kblocks;             // ceil(dp / BK)
  int kchunk;              // k-blocks per accumulation chunk (per pass)
  int tiles_m, tiles_n;
  int metric;              // TDA_METRIC_*
  int symmetric;           // A and B are the same set (zero the diagonal)
  int tri;                 // symmetric and square tiling: only the tiles on and above the diagonal are computed; the epilogue also
                           // stores the transposed tile (the distance of (j,i) IS the value computed for (i,j): D is exactly symmetric)
  int tiles_per_problem;
  float disconnect;        // distances >= disconnect become +inf (umap-learn's disconnection_distance); +inf = off
  const float* norm_a;     // [batch*n] squared norms (euclidean metrics)
  const float* norm_b;     // [batch*m]
  float* out;              // [batch, n, m]
};

// tile t of the launch -> problem b, tile row ti, tile column tj
__device__ __forceinline__ void tile_coords(const GemmParams& P, int t, int& b, int& ti, int& tj) {
  b = t / P.tiles_per_problem;
  int r = t % P.tiles_per_problem;
  if (!P.tri) { ti = r / P.tiles_n; tj = r % P.tiles_n; return; }
  ti = 0;
  while (r >= P.tiles_n - ti) { r -= P.tiles_n - ti; ++ti; }   // row ti of the upper triangle holds tiles_n - ti tiles
  tj = ti + r;
}

__global__ void __launch_bounds__(kThreads, 1)
pdist_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const GemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // 128B swizzle needs 1024B alignment
  uint64_t* bars = (uint64_t*)(smem + kStages * kStageBytes);
  uint64_t* full = bars;                       // [kStages]  TMA -> MMA
  uint64_t* empty = bars + kStages;            // [kStages]  MMA -> TMA
  uint64_t* tfull = bars + 2 * kStages;        // [2]        MMA -> epilogue
  uint64_t* tempty = bars + 2 * kStages + 2;   // [2]        epilogue -> MMA
  uint32_t* tmem_base_slot = (uint32_t*)(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = P.tiles_per_problem * P.batch;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int b, ti, tj;
        tile_coords(P, t, b, ti, tj);
        const int row_a = b * P.n + ti * BM;
        const int row_b = b * P.m + tj * BN;
        // every k-block brings its four operand tiles once; the three products of the split are formed from them
        // (the earlier version re-loaded a tile pair per product: 1.5x the L2 -> shared-memory traffic, which is what
        // bounds this kernel)
        for (int kb = 0; kb < P.kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * kStageBytes;
          mbar_expect_tx(&full[stage], kStageBytes);
          tma_load_2d(&map_a_hi, &full[stage], st, kb * BK, row_a);
          tma_load_2d(&map_a_lo, &full[stage], st + kTileABytes, kb * BK, row_a);
          tma_load_2d(&map_b_hi, &full[stage], st + 2 * kTileABytes, kb * BK, row_b);
          tma_load_2d(&map_b_lo, &full[stage], st + 2 * kTileABytes + kTileBBytes, kb * BK, row_b);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc_iter = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int k0 = 0; k0 < P.kblocks; k0 += P.kchunk, ++acc_iter) {
        const int as = acc_iter & 1;
        const uint32_t aphase = (acc_iter >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);  // epilogue has drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        const int total_kb = min(k0 + P.kchunk, P.kblocks) - k0;
        for (int it = 0; it < total_kb; ++it) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          if (lane == 0) {
            const uint32_t s0 = smem_u32(smem + stage * kStageBytes);
            const uint64_t a_hi = make_smem_desc(s0), a_lo = make_smem_desc(s0 + kTileABytes);
            const uint64_t b_hi = make_smem_desc(s0 + 2 * kTileABytes), b_lo = make_smem_desc(s0 + 2 * kTileABytes + kTileBBytes);
            // lo*hi, hi*lo, hi*hi (small terms first); advance 8 floats = 32 bytes inside the swizzle row: +2 in the >>4 address field
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) tcgen05_mma_tf32(tmem_d, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), kIdesc, (it | k) != 0);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) tcgen05_mma_tf32(tmem_d, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), kIdesc, 1);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) tcgen05_mma_tf32(tmem_d, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), kIdesc, 1);
            tcgen05_commit(&empty[stage]);                       // frees the smem stage once these MMAs retire
            if (it == total_kb - 1) tcgen05_commit(&tfull[as]);  // chunk complete -> epilogue
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue warps (TMEM lane quarter = warp % 4) =====
    const int q = warp & 3;
    uint32_t acc_iter = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int b, ti, tj;
      tile_coords(P, t, b, ti, tj);
      const int i = ti * BM + q * 32 + lane;   // row inside the problem
      const int j0 = tj * BN;
      const bool mirror = P.tri && tj > ti;    // off-diagonal tile of a symmetric problem: its transpose is stored as well
      float acc[BN];
#pragma unroll
      for (int u = 0; u < BN; ++u) acc[u] = 0.f;
      for (int k0 = 0; k0 < P.kblocks; k0 += P.kchunk, ++acc_iter) {
        const int as = acc_iter & 1;
        const uint32_t aphase = (acc_iter >> 1) & 1;
        mbar_wait(&tfull[as], aphase);
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c), v);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u) acc[c + u] += __uint_as_float(v[u]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
      }
      if (i < P.n) {
        const float na = P.norm_a ? P.norm_a[(size_t)b * P.n + i] : 0.f;
        float* orow = P.out + ((size_t)b * P.n + i) * P.m;
#pragma unroll
        for (int u = 0; u < BN; u += 4) {
          float o[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int j = j0 + u + w;
            const float dot = acc[u + w];
            float d;
            if (P.metric == TDA_METRIC_COSINE) {
              d = fminf(fmaxf(1.f - dot, 0.f), 2.f);
            } else if (P.metric == TDA_METRIC_DOT) {
              d = dot;
            } else {
              const float nb = (j < P.m) ? P.norm_b[(size_t)b * P.m + j] : 0.f;
              d = fmaxf(na + nb - 2.f * dot, 0.f);
              if (P.metric == TDA_METRIC_EUCLIDEAN) d = sqrtf(d);
            }
            if (P.metric != TDA_METRIC_DOT) {
              if (P.symmetric && j == i) d = 0.f;
              if (d >= P.disconnect) d = INFINITY;
            }
            o[w] = d;
          }
          const int j = j0 + u;
          if (j + 3 < P.m && (P.m & 3) == 0) {
            *reinterpret_cast<float4*>(orow + j) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
#pragma unroll
            for (int w = 0; w < 4; ++w)
              if (j + w < P.m) orow[j + w] = o[w];
          }
          if (mirror) {   // D[j][i]: the 32 lanes of a warp hold 32 consecutive i -> one 128-byte store per column
#pragma unroll
            for (int w = 0; w < 4; ++w)
              if (j + w < P.n) P.out[((size_t)b * P.n + (j + w)) * P.m + i] = o[w];
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// operand preparation: optional column centring (euclidean) / row normalisation (cosine), then the
// tf32 hi/lo split, written zero-padded to a leading dimension that is a multiple of BK.
__global__ void col_mean_kernel(const float* __restrict__ X, int n, int d, float* __restrict__ mean) {
  const int p = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const float* x = X + (size_t)p * n * d;
  double acc = 0.0;
  for (int r = 0; r < n; ++r) acc += (double)x[(size_t)r * d + c];
  mean[(size_t)p * d + c] = (float)(acc / n);
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// one CTA (128 threads) per row
__global__ void __launch_bounds__(128) prep_kernel(const float* __restrict__ X, int rows_per_problem, int d, int dp, int metric,
                                                   const float* __restrict__ mean, float* __restrict__ hi, float* __restrict__ lo,
                                                   float* __restrict__ norm) {
  __shared__ double red[4];
  const size_t row = blockIdx.x;
  const int p = (int)(row / rows_per_problem);
  const float* x = X + row * d;
  const float* mu = mean ? mean + (size_t)p * d : nullptr;
  double acc = 0.0;
  for (int c = threadIdx.x; c < d; c += 128) {
    float v = x[c] - (mu ? mu[c] : 0.f);
    acc += (double)v * (double)v;
  }
  acc = warp_sum_f64(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  const double nrm2 = red[0] + red[1] + red[2] + red[3];
  float scale = 1.f;
  if (metric == TDA_METRIC_COSINE) {
    const float nr = (float)sqrt(nrm2);
    scale = nr > 0.f ? 1.f / nr : 1.f;  // sklearn.preprocessing.normalize leaves all-zero rows untouched
  }
  if (threadIdx.x == 0 && norm) norm[row] = (float)nrm2;
  for (int c = threadIdx.x; c < dp; c += 128) {
    float v = 0.f;
    if (c < d) v = (x[c] - (mu ? mu[c] : 0.f)) * scale;
    const float h = to_tf32(v);
    hi[row * dp + c] = h;
    lo[row * dp + c] = to_tf32(v - h);
  }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
static int make_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t dp, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(TDA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {dp, rows};
  cuuint64_t strides[1] = {dp * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(TDA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return TDA_OK;
}

struct Layout {
  float *a_hi, *a_lo, *b_hi, *b_lo, *norm_a, *norm_b, *mean;
  int dp;
  size_t total;
};
static Layout make_layout(void* ws, int n, int m, int d, int batch, bool symmetric) {
  Layout L;
  memset(&L, 0, sizeof(L));
  Carver c(ws, ~size_t(0));
  L.dp = (d + BK - 1) / BK * BK;
  // one extra tile of rows so that a TMA box starting inside the last problem never leaves the allocation
  const size_t rows_a = (size_t)batch * n + BM, rows_b = (size_t)batch * m + BN;
  L.a_hi = c.take<float>(rows_a * L.dp);
  L.a_lo = c.take<float>(rows_a * L.dp);
  if (!symmetric) {
    L.b_hi = c.take<float>(rows_b * L.dp);
    L.b_lo = c.take<float>(rows_b * L.dp);
  } else { L.b_hi = L.a_hi; L.b_lo = L.a_lo; }
  L.norm_a = c.take<float>((size_t)batch * n);
  L.norm_b = symmetric ? L.norm_a : c.take<float>((size_t)batch * m);
  L.mean = c.take<float>((size_t)batch * d);
  L.total = c.off;
  return L;
}

}  // namespace pdist
}  // namespace tda

using namespace tda;
using namespace tda::pdist;

extern "C" size_t tda_pdist_workspace_bytes(int n, int m, int d, int batch, int symmetric) {
  if (n <= 0 || m <= 0 || d <= 0 || batch <= 0) return 0;
  return make_layout(nullptr, n, m, d, batch, symmetric != 0).total + 1024;
}

extern "C" int tda_pdist(const float* X, const float* Y, int n, int m, int d, int batch, int metric, float disconnect, float* D,
                         void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!X || !D || !ws || n <= 0 || d <= 0 || batch <= 0) return set_error(TDA_ERR_INVALID, "tda_pdist: bad arguments");
  if (metric < TDA_METRIC_SQEUCLIDEAN || metric > TDA_METRIC_DOT) return set_error(TDA_ERR_INVALID, "tda_pdist: unknown metric %d", metric);
  const bool symmetric = (Y == nullptr) || (Y == X && m == n);   // (a row block of X against all of X shares X's base pointer)
  if (symmetric) m = n;
  if (m <= 0) return set_error(TDA_ERR_INVALID, "tda_pdist: bad m");
  if ((int64_t)batch * n + BM >= (1ll << 31) || (int64_t)batch * m + BN >= (1ll << 31)) return set_error(TDA_ERR_UNSUPPORTED, "tda_pdist: too many rows");
  Layout L = make_layout(ws, n, m, d, batch, symmetric);
  if (L.total > ws_bytes) return set_error(TDA_ERR_WORKSPACE, "tda_pdist: workspace %zu < required %zu", ws_bytes, L.total);
  const bool euclid = (metric == TDA_METRIC_SQEUCLIDEAN || metric == TDA_METRIC_EUCLIDEAN);
  const float* mean = nullptr;
  stage_begin_if(STAGE_PDIST_PREP, stream);
  if (euclid && symmetric) {  // centring is distance preserving and removes the ||x||^2 cancellation of raw activations
    dim3 g((d + 127) / 128, batch);
    col_mean_kernel<<<g, 128, 0, stream>>>(X, n, d, L.mean);
    count_launch();
    mean = L.mean;
  }
  // padding rows beyond the last problem must be finite (they are multiplied, then masked)
  TDA_CUDA_CHECK(cudaMemsetAsync(L.a_hi + (size_t)batch * n * L.dp, 0, sizeof(float) * BM * L.dp, stream));
  TDA_CUDA_CHECK(cudaMemsetAsync(L.a_lo + (size_t)batch * n * L.dp, 0, sizeof(float) * BM * L.dp, stream));
  prep_kernel<<<(unsigned)((size_t)batch * n), 128, 0, stream>>>(X, n, d, L.dp, metric, mean, L.a_hi, L.a_lo, L.norm_a);
  count_launch();
  if (!symmetric) {
    TDA_CUDA_CHECK(cudaMemsetAsync(L.b_hi + (size_t)batch * m * L.dp, 0, sizeof(float) * BN * L.dp, stream));
    TDA_CUDA_CHECK(cudaMemsetAsync(L.b_lo + (size_t)batch * m * L.dp, 0, sizeof(float) * BN * L.dp, stream));
    prep_kernel<<<(unsigned)((size_t)batch * m), 128, 0, stream>>>(Y, m, d, L.dp, metric, nullptr, L.b_hi, L.b_lo, L.norm_b);
    count_launch();
  }
  stage_end_if(STAGE_PDIST_PREP, stream);
  TDA_LAUNCH_CHECK();
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, L.a_hi, (uint64_t)batch * n + BM, L.dp, BM))) return rc;
  if ((rc = make_map(&ma_lo, L.a_lo, (uint64_t)batch * n + BM, L.dp, BM))) return rc;
  if ((rc = make_map(&mb_hi, L.b_hi, (uint64_t)batch * m + BN, L.dp, BN))) return rc;
  if ((rc = make_map(&mb_lo, L.b_lo, (uint64_t)batch * m + BN, L.dp, BN))) return rc;
  GemmParams P;
  P.n = n; P.m = m; P.batch = batch; P.kblocks = L.dp / BK; P.kchunk = 4;
  P.tiles_m = (n + BM - 1) / BM; P.tiles_n = (m + BN - 1) / BN;
  P.metric = metric; P.symmetric = symmetric ? 1 : 0; P.disconnect = disconnect;
  P.tri = (symmetric && P.tiles_m == P.tiles_n) ? 1 : 0;
  P.tiles_per_problem = P.tri ? P.tiles_n * (P.tiles_n + 1) / 2 : P.tiles_m * P.tiles_n;
  P.norm_a = L.norm_a; P.norm_b = L.norm_b; P.out = D;
  int sms = 0, dev = 0;
  TDA_CUDA_CHECK(cudaGetDevice(&dev));
  TDA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TDA_CUDA_CHECK(cudaFuncSetAttribute(pdist_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  const long long tiles = (long long)P.tiles_per_problem * batch;
  const int grid = (int)(tiles < sms ? tiles : sms);
  {
    StageScope st(STAGE_PDIST_GEMM, stream);
    pdist_gemm_kernel<<<grid, kThreads, kSmemBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, P);
  }
  count_launch();
  TDA_LAUNCH_CHECK();
  return TDA_OK;
}
