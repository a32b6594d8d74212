/* tda_b200.h -- C ABI of libtda_b200.so: hand-written sm_100a kernels for the per-layer TDA hot path of
 * Princeton-Applied-Geometry-Topology/tda-multimodal (pairwise distances -> UMAP stages -> Vietoris-Rips).
 *
 * The reference has no FFI of its own; its seam is three Python imports (`import umap`,
 * `from ripser import ripser`, `from persim import plot_diagrams`; debug_tda_pipeline.py:9-11).  Each
 * entry point below names the reference call (file:line) / upstream stage it replaces.  INTEGRATION.md
 * shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`;
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it unless stated otherwise;
 *  - all problems of one call have the same shape; `batch` = number of independent problems
 *    (layers x bootstrap resamples), laid out contiguously [batch, ...];
 *  - return value 0 = success, negative = error (tda_last_error() gives the text; thread local);
 *  - the library keeps no global state besides the error string; scratch memory is the caller's
 *    workspace `ws` (size from the matching *_workspace_bytes call).
 */
#ifndef TDA_B200_H
#define TDA_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TDA_METRIC_SQEUCLIDEAN 0
#define TDA_METRIC_EUCLIDEAN 1
#define TDA_METRIC_COSINE 2
#define TDA_METRIC_DOT 3 /* plain Gram matrix X Y^T (metrics.py:368) */

int tda_version(void);
const char* tda_last_error(void);
/* number of kernels launched by this library in the calling thread since the last reset (bench.py gpu_launches) */
int64_t tda_launch_count(void);
void tda_launch_count_reset(void);

/* ---- Rips persistence -------------------------------------------------------------------------
 * Replaces ripser(X, maxdim=1)['dgms'] (debug_tda_pipeline.py:109-110; analyze_tda_over_layers.py:76;
 * analyze_adversarial_tda.py:100-101).
 *
 * tda_pdist_lowdim: ripser.py's `pairwise_distances(X, metric='euclidean')` for a low-dimensional cloud
 *   (the 3-D UMAP output): squared distance by exact differences in float64, rounded to float32, float32
 *   sqrt.  pts [batch,n,d] float32, dm [batch,n,n] float32.  d <= 64.
 */
int tda_pdist_lowdim(const float* pts, int n, int d, int batch, float* dm, void* stream);

/* tda_rips: Vietoris-Rips persistence (Z/2) of `batch` dense distance matrices, H0 and H1.
 *   dm      [batch,n,n] float32 symmetric, zero diagonal
 *   thresh  +inf => per-problem enclosing radius min_i max_j dm (ripser's default)
 *   maxdim  0 or 1 (2 is not implemented in this round: returns TDA_ERR_UNSUPPORTED)
 *   h0_pairs [batch,n,2] float32: rows (0,death) ascending, then one (0,+inf) per component
 *   h0_simplex [batch,n,2] int64 or NULL: (birth vertex or -1, death edge index i(i-1)/2+j or -1)
 *   h1_pairs [batch,cap1,2] float32: (birth,death) in ripser's emission order (birth descending)
 *   h1_simplex [batch,cap1,2] int64 or NULL: (birth edge index, death triangle index C(a,3)+C(b,2)+c or -1)
 *   counts  [batch,4] int32: n_h0 rows, n_h1 rows, num_edges (<= thresh), status (0 ok, else TDA_ERR_*)
 *   thresh_out [batch] float32 or NULL: threshold actually used
 * Returns TDA_ERR_CAPACITY (after synchronising `stream`) if any problem overflowed cap1 or the
 * internal column pool; sizes come from tda_rips_workspace_bytes(n, batch, maxdim, cap1, pool_bytes).
 * This call synchronises `stream` before returning (it needs the overflow status).
 */
size_t tda_rips_workspace_bytes(int n, int batch, int maxdim, int cap1, size_t pool_bytes);
int tda_rips(const float* dm, int n, int batch, int maxdim, float thresh,
             float* h0_pairs, int64_t* h0_simplex, float* h1_pairs, int64_t* h1_simplex, int cap1,
             int32_t* counts, float* thresh_out, void* ws, size_t ws_bytes, size_t pool_bytes, void* stream);
/* device statistics of the last tda_rips call on this workspace: [batch,16] int64:
 * columns(non-MST edges<=thresh), apparent, reduced, additions, pushes, pops, horizon_extensions, max_V,
 * SM cycles in extract / owner lookup / apparent add / reduced-column add / horizon extension / finalise,
 * edges re-enumerated by reduced-column adds, edges re-enumerated by horizon extensions */
int tda_rips_stats(const void* ws, int n, int batch, int maxdim, int cap1, size_t pool_bytes, int64_t* stats_host);

#ifdef __cplusplus
}
#endif
#endif
