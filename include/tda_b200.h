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
 *  - the library keeps no global state besides the thread-local error string / launch counter / stage timer and the
 *    process-wide tuning options of tda_set_option (it reads no environment variables); scratch memory is the caller's
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
/* Tuning options (process-wide integers; defaults in parentheses).  Unknown names return TDA_ERR_INVALID / -1.
 *   rips_reducer (0): residual H1 reducer -- 0 "sweep2" substitute by rank + verify by window, 1 row sweep with a sequential
 *                     resolver warp, 2 row sweep substitute-then-verify per 512-row chunk, 3 key bitset (any n)
 *   rips_w0 (1024), rips_wsparse (8192), rips_wmax (32768), rips_dense_min (64), rips_dense_div (8): sweep2 window schedule
 *   rips_cluster (0 = auto: 8 for up to 4 clouds per launch, else 4): CTAs of the thread-block cluster that reduces one cloud (sweep2)
 *   rips_apparent_rows (1): apparent pairs by matrix row (rank row of a vertex in shared memory); 0: one warp per edge in rank order
 *   rips_h0_chunked (1): H0 in one launch from the sorted edge list (Boruvka / Kruskal by chunks, n <= 16384); 0: Boruvka rounds on the rank matrix
 *   rips_wc_max_rows (262144): sweep2, rows a single warp sweeps before it hands its column to the cluster engine
 *   rips_warp_engine (1): sweep2 reduces every column by a single warp first (speculatively, committed in ripser's order); 0: windows only
 *   sgd_mode (0): 0 deterministic SGD (thread-block cluster per cloud for fit, warp per point for transform; bit-reproducible
 *                 for a given seed), 3 per-epoch kernels with float atomics (used anyway for n > 8192 or n_components != 3)
 *   sgd_cluster (0 = auto, like rips_cluster): CTAs per cloud of the deterministic fit kernel, sgd_tile (16): vertices per warp task;  spectral_cluster (8): CTAs per cloud of the Lanczos kernel
 *   for connected graphs (0: always the one-CTA-per-component kernel);  sweep_exclusive (0), knn_loads (8), debug_sync (0),
 *   h2_stats (0) */
int tda_set_option(const char* name, long long value);
long long tda_get_option(const char* name);

/* Optional stage timer (off by default): when enabled, every entry point brackets its stages with CUDA events on
 * the stream it launches on; tda_stage_timing_read synchronises those events and returns the accumulated
 * milliseconds and call counts per stage (arrays of TDA_STAGE_COUNT entries); returns TDA_STAGE_COUNT.
 * bench.py uses it to time the dominant kernel inside the timed region (roofline.achieved). */
#define TDA_STAGE_PDIST_PREP 0
#define TDA_STAGE_PDIST_GEMM 1
#define TDA_STAGE_KNN_SMOOTH 2
#define TDA_STAGE_FUZZY 3
#define TDA_STAGE_SPECTRAL 4
#define TDA_STAGE_SGD 5
#define TDA_STAGE_RIPS_PDIST 6
#define TDA_STAGE_RIPS_SORT 7
#define TDA_STAGE_RIPS_H0 8
#define TDA_STAGE_RIPS_APPARENT 9
#define TDA_STAGE_RIPS_REDUCE 10
#define TDA_STAGE_COUNT 11
void tda_stage_timing_enable(int on);
void tda_stage_timing_reset(void);
int tda_stage_timing_read(double* ms_out, int64_t* calls_out, int n);
/* every timed stage span since the last reset, in enqueue order: stage id, start and end in ms after the reset (CUDA events on the
 * launching streams, so spans of different streams overlap); returns the number of spans (at most `cap` are written). */
int tda_stage_timeline_read(int* stage_out, double* start_ms_out, double* end_ms_out, int cap);

/* ---- pairwise distances on high-dimensional activations (tcgen05 / TMEM / TMA GEMM, 3xTF32) ------
 * Replaces sklearn.metrics.pairwise_distances inside umap-learn's small-data path (metric='cosine';
 * debug_tda_pipeline.py:96-104, analyze_tda_over_layers.py:69,72, analyze_adversarial_tda.py:85-93), inside
 * ripser.py for raw high-dimensional clouds, torch.cdist of metrics.py:143 and the Gram matrix of
 * metrics.py:368.
 *   X [batch,n,d] float32; Y [batch,m,d] float32 or NULL (=> Y = X, m = n, exact-zero diagonal)
 *   metric TDA_METRIC_*; disconnect: distances >= disconnect become +inf (umap-learn's
 *   disconnection_distance, 2.0 for cosine); pass +inf to disable
 *   D [batch,n,m] float32
 */
size_t tda_pdist_workspace_bytes(int n, int m, int d, int batch, int symmetric);
int tda_pdist(const float* X, const float* Y, int n, int m, int d, int batch, int metric, float disconnect, float* D,
              void* ws, size_t ws_bytes, void* stream);

/* ---- UMAP stages ----------------------------------------------------------------------------------
 * Replace the inside of umap.UMAP(n_neighbors, n_components=3, min_dist=0.1, random_state=42, metric='cosine')
 * .fit_transform / .fit / .transform (debug_tda_pipeline.py:96-104; analyze_tda_over_layers.py:38-44,69,72;
 * analyze_adversarial_tda.py:85-93).  Stage names are umap-learn's.
 *
 * tda_knn_smooth: fast_knn_indices + smooth_knn_dist, fused, one warp per row of D.
 *   D [batch,n,m] float32 (m = n for fit; m = n_train for transform); k <= 256
 *   knn_idx [batch,n,k] int32 ascending by (distance, index), self included, -1 where the distance is +inf
 *   knn_dist [batch,n,k] float32; sigma, rho [batch,n] float32; ws: 8*batch bytes
 */
int tda_knn_smooth(const float* D, int n, int m, int batch, int k, float local_connectivity, float bandwidth, int n_iter,
                   int32_t* knn_idx, float* knn_dist, float* sigma, float* rho, void* ws, size_t ws_bytes, void* stream);
/* tda_fuzzy_graph: compute_membership_strengths + fuzzy union mix*(P+P^T-P.P^T)+(1-mix)*P.P^T + make_epochs_per_sample.
 *   Output is a slot table of 2*n*k directed entries per cloud: head, tail [batch,2nk] int32, weight [batch,2nk]
 *   float32 (0 = empty slot), eps [batch,2nk] float32 = max_w/w, or -1 for empty slots and for entries pruned by
 *   w < max_w/n_epochs; max_weight [batch] float32.
 */
/* tda_knn_fused: exact kNN + sigma/rho of the rows [row_begin, row_end) of ONE cloud X [n,d] against all of its points WITHOUT the
 *   n x n distance matrix (the per-rank body of the row-sharded kNN of a 100k-point cloud, config C5 of
 *   BASELINE.json): tda_pdist on blocks of `row_block` rows (tensor-core GEMM), each block consumed by the top-k kernel and
 *   overwritten by the next.  Outputs are indexed from row_begin: knn_idx / knn_dist [row_end-row_begin, k], sigma / rho
 *   [row_end-row_begin]; column indices are global.  The sigma floor of rows without a positive neighbour distance uses the mean
 *   distance of the row block.  ws: tda_knn_fused_workspace_bytes(n, d, row_block), 256-byte aligned. */
size_t tda_knn_fused_workspace_bytes(int n, int d, int row_block);
int tda_knn_fused(const float* X, int n, int d, int row_begin, int row_end, int k, int metric, float disconnect,
                  float local_connectivity, int32_t* knn_idx, float* knn_dist, float* sigma, float* rho, int row_block, void* ws,
                  size_t ws_bytes, void* stream);
int tda_fuzzy_graph(const int32_t* knn_idx, const float* knn_dist, const float* sigma, const float* rho, int n, int k, int batch,
                    float mix_ratio, int n_epochs, int32_t* head, int32_t* tail, float* weight, float* eps, float* max_weight,
                    void* stream);
/* tda_umap_sgd: optimize_layout_euclidean.  Y [batch,n_head,dim] in/out; Y_other [batch,n_tail,dim] (NULL and
 *   move_other=1 for fit: tail == head embedding); slot table as produced above; dim 1..4.
 *   ws (256-byte aligned, tda_umap_sgd_workspace_bytes): the per-vertex adjacency of the deterministic fit kernel, or the
 *   float4-padded embeddings of the per-epoch kernels.  With dim == 3 and option sgd_mode == 0 the result is bit-reproducible for
 *   a given seed (all gradients of an epoch are evaluated on the positions of the epoch's start, every vertex sums its own
 *   displacement in a fixed order); without ws, for n > 8192 or other dims the per-epoch kernels with float atomics run. */
size_t tda_umap_sgd_workspace_bytes(int slots, int n_head, int n_tail, int dim, int batch, int move_other);
int tda_umap_sgd(float* Y, const float* Y_other, const int32_t* head, const int32_t* tail, const float* eps, int slots, int n_head,
                 int n_tail, int dim, int batch, int n_epochs, float a, float b, float gamma, float alpha0,
                 float negative_sample_rate, int move_other, uint64_t seed, void* ws, size_t ws_bytes, void* stream);
int tda_umap_init_random(float* Y, int n, int dim, int batch, float lo, float hi, uint64_t seed, void* stream);
/* noisy_scale_coords (max|Y| -> 10, + N(0,noise)) followed by the per-axis rescale to [0,10] */
int tda_umap_rescale(float* Y, int n, int dim, int batch, float noise, uint64_t seed, void* stream);
/* transform(): bipartite membership strengths, init_transform (weighted mean of the train embedding) and schedule */
int tda_umap_transform_init(const int32_t* knn_idx, const float* knn_dist, const float* sigma, const float* rho,
                            const float* train_embedding, int n_query, int n_train, int k, int dim, int batch, int n_epochs,
                            float* Y, int32_t* head, int32_t* tail, float* weight, float* eps, float* max_weight, void* stream);

/* spectral initialisation (umap-learn spectral_layout / multi_component_layout):
 *  tda_graph_components: connected components of the pruned graph (entries with eps > 0); comp [batch,n] ids numbered
 *   by smallest member vertex (scipy order), ncomp [batch], comp_size [batch,n] (first ncomp entries), degree [batch,n];
 *   ws: 12*batch*n bytes, 8-byte aligned (degrees are accumulated in 64-bit fixed point: order independent, reproducible).
 *  tda_spectral_embed: for every component with >= min_size vertices, the `dim` non-trivial bottom eigenvectors of the
 *   symmetric normalised Laplacian (Lanczos, full reorthogonalisation), written as unit vectors into the component's
 *   rows of Y [batch,n,dim]; evals [batch,maxcomp,4] or NULL: the `dim` eigenvalues of D^-1/2 W D^-1/2 and, for dim < 4, in the
 *   last slot the largest Ritz residual estimate |beta_m s_m| of the returned pairs after the fixed 96 Lanczos steps (ARPACK, which
 *   umap-learn uses, iterates until that falls below its tolerance; a caller can warn or fall back to a random init on a large value).
 */
size_t tda_spectral_workspace_bytes(int n, int batch, int maxcomp, int slots);
/* tda_spectral_init: components + eigenvectors + multi_component_layout in one call WITHOUT a host round trip.  Every component
 *   (up to min(maxcomp, 32) per cloud) is laid out by a thread-block-cluster Lanczos kernel; clouds with 2..2*dim components get
 *   umap-learn's +-e_k meta positions; clouds with more components get umap-learn's component_layout on the device when the data
 *   is given (X [batch,n,d] float32, metric TDA_METRIC_SQEUCLIDEAN / EUCLIDEAN / COSINE): centroids of the components in data
 *   space, affinity exp(-dist^2), spectral embedding of the normalised Laplacian (Jacobi, fp64).  X may be NULL (d, metric
 *   ignored): such clouds then get status 1.  ncomp_out [batch] = number of components; status_out [batch] = 0 done, 1 = more
 *   components than handled here (the caller lays those clouds out through tda_graph_components / tda_spectral_embed and its
 *   own component_layout).  Y [batch,n,dim] is overwritten. */
size_t tda_spectral_init_workspace_bytes(int n, int batch, int maxcomp, int slots, int d);
int tda_spectral_init(const int32_t* head, const int32_t* tail, const float* weight, const float* eps, int slots, int n, int dim,
                      int batch, int maxcomp, uint64_t seed, const float* X, int d, int metric, float* Y, int32_t* ncomp_out,
                      int32_t* status_out, void* ws, size_t ws_bytes, void* stream);
int tda_graph_components(const int32_t* head, const int32_t* tail, const float* weight, const float* eps, int slots, int n, int batch,
                         int32_t* comp, int32_t* ncomp, int32_t* comp_size, float* degree, void* ws, size_t ws_bytes, void* stream);
int tda_spectral_embed(const int32_t* head, const int32_t* tail, const float* weight, const float* eps, int slots, int n, int dim,
                       int batch, const int32_t* comp, const int32_t* ncomp, const int32_t* comp_size, const float* degree,
                       int maxcomp, int min_size, uint64_t seed, float* Y, float* evals, void* ws, size_t ws_bytes, void* stream);

/* ---- silhouette ------------------------------------------------------------------------------------
 * Replaces sklearn.metrics.silhouette_score(point_cloud_low_dim, labels) (debug_tda_pipeline.py:117-118;
 * analyze_adversarial_tda.py:108-111) on the euclidean distance matrix of the 3-D cloud (tda_pdist_lowdim).
 *   dm [batch,n,n] float32; labels [batch,n] int32 in [0, n_labels), 2 <= n_labels <= 64; score [batch] float32 (mean
 *   silhouette coefficient; points of a singleton label contribute 0, as in scikit-learn); ws: 8*batch bytes. */
int tda_silhouette(const float* dm, const int32_t* labels, int n, int batch, int n_labels, float* score, void* ws, size_t ws_bytes,
                   void* stream);

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
 *   maxdim  0 or 1 (H2 is the separate call tda_rips_h2 on top of a finished maxdim=1 run; maxdim >= 2 here returns
 *           TDA_ERR_UNSUPPORTED)
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
/* tda_rips_launch: the same work, only enqueued on `stream` (no synchronisation): the caller synchronises the stream and
 * reads counts[:,3] (0 = ok, TDA_ERR_CAPACITY = that problem overflowed: call again with larger cap1 / pool) itself.  Lets
 * several batches overlap on different streams (tda_multimodal_b200.pipeline.layer_sweep). */
int tda_rips_launch(const float* dm, int n, int batch, int maxdim, float thresh,
                    float* h0_pairs, int64_t* h0_simplex, float* h1_pairs, int64_t* h1_simplex, int cap1,
                    int32_t* counts, float* thresh_out, void* ws, size_t ws_bytes, size_t pool_bytes, void* stream);
/* Subsets of one cloud (bootstrap resamples: `ripser(Y[idx], maxdim=1)` for many index sets of the same Y; config 4 of
 * BASELINE.json, the resample loop around debug_tda_pipeline.py:109-110).  A subset given by strictly ascending parent indices keeps
 * the order of its edges (equal float lengths, tie-break index monotone under the relabelling), so its filtration ranks are a
 * flag + prefix count over the parent's sorted edge list: no distance matrix, key generation or radix sort per subset.
 * tda_rips_sort_edges: ALL edges of `batch` clouds in ripser's order (length ascending, index descending; no threshold).
 *   dm [batch,n,n] float32; ends_out [batch,E] uint32 = (i << 16 | j), i > j; sdist_out [batch,E] float32; E = n(n-1)/2.
 * tda_rips_subsets_launch: as tda_rips_launch for the `batch` clouds points[subset_idx[b]] (m points each) of ONE parent:
 *   parent_ends / parent_sdist [E_parent] from tda_rips_sort_edges, parent_dm [n_parent,n_parent] (only read for the enclosing
 *   radius: may be NULL when `thresh` is finite), subset_idx [batch,m] int32, strictly ascending per row (not checked).
 *   Outputs, workspace (tda_rips_workspace_bytes(m, batch, ...)) and overflow protocol as tda_rips_launch; the results are
 *   bit-identical to tda_rips_launch on tda_pdist_lowdim(points[subset_idx[b]]). */
size_t tda_rips_sort_edges_workspace_bytes(int n, int batch);
int tda_rips_sort_edges(const float* dm, int n, int batch, uint32_t* ends_out, float* sdist_out, void* ws, size_t ws_bytes, void* stream);
int tda_rips_subsets_launch(const uint32_t* parent_ends, const float* parent_sdist, const float* parent_dm, int n_parent,
                            const int32_t* subset_idx, int m, int batch, int maxdim, float thresh,
                            float* h0_pairs, int64_t* h0_simplex, float* h1_pairs, int64_t* h1_simplex, int cap1,
                            int32_t* counts, float* thresh_out, void* ws, size_t ws_bytes, size_t pool_bytes, void* stream);
/* tda_rips_h2: H2 (optional `maxdim=2` of ripser(X, maxdim)) on top of a FINISHED tda_rips(maxdim=1) call: `ws1` is that
 * call's workspace (with the same n, batch, cap1, pool_bytes1), which still holds the rank matrix and the H1 pivots (clearing).
 * Triangles in an apparent pair with a tetrahedron are skipped in parallel, the residual triangle columns are reduced like the
 * H1 columns (implicit cohomology over Z/2, working column = bitset over tetrahedron keys).  n <= 2048 (triangle keys E*n < 2^32).
 *   h2_pairs [batch,cap2,2] float32 (birth, death), death > birth; counts2 [batch,4] int32: -, n_h2 rows, -, status;
 *   cap2 a power of two; returns TDA_ERR_CAPACITY (after synchronising) if a problem overflowed cap2 or the pool.
 *   far_bytes: room for the far buckets (tetrahedron keys beyond the 2^32-bit window of the working column wait there, one
 *   region per window and resident CTA, until the window reaches them); 0 = none (later windows are re-enumerated instead). */
size_t tda_rips_h2_workspace_bytes(int n, int batch, int cap2, size_t pool_bytes, size_t far_bytes);
int tda_rips_h2(const void* ws1, int n, int batch, int cap1, size_t pool_bytes1, float* h2_pairs, int cap2, int32_t* counts2,
                void* ws2, size_t ws2_bytes, size_t pool_bytes2, size_t far_bytes, void* stream);
/* tda_greedy_perm: furthest-point landmark selection with ripser.py's `n_perm` semantics (ripser(X, n_perm=...): start at point 0,
 * lowest index on ties), one launch of one thread-block cluster.  X = points [n,d] float32 (is_matrix = 0, euclidean, d <= 16) or a
 * distance matrix [n,n] (is_matrix = 1); idx_out [n_perm] int32; lambda_out [n_perm] float32 (lambda_out[n_perm-1] = r_cover);
 * ws: tda_greedy_perm_workspace_bytes(n) (only used when a CTA's share of the minimum distances does not fit in shared memory). */
size_t tda_greedy_perm_workspace_bytes(int n);
int tda_greedy_perm(const float* X, int n, int d, int n_perm, int is_matrix, int32_t* idx_out, float* lambda_out, void* ws, size_t ws_bytes,
                    void* stream);
/* device statistics of the last tda_rips call on this workspace: [batch, TDA_RIPS_STATS] int64.  For the default reducer (sweep2):
 *  0 columns (non-MST edges <= thresh), 1 apparent pairs, 2 reduced columns, 3 column additions (flips kept + reduced columns
 *  added), 4 rows substituted, 5 non-apparent pivots (events + deaths), 6 windows, 7 largest |V|, 8..13 SM cycles of CTA
 *  thread 0 in: the warp stage (every column by one warp), the commit loop, and in the windows of the cluster engine: substitution,
 *  verification, decision + events, Pm moves + column finalisation; 14 edges added through reduced columns, 15 heavy rows verified, 16 substitution rounds after the first,
 *  17 rows handled in those rounds, 18 rows of Pm moved, 19 columns that went dense, 20 columns the commit loop resumed (tentative pivot owned), 21 columns handed to the cluster engine,
 *  22 SM cycles thread 0 spent in cluster barriers (part of 8..13), 23 number of those barriers.
 *  (reducers 1-3 fill 0..15 with their own counters: rows streamed, pivots, restarts, ...) */
#define TDA_RIPS_STATS 24
int tda_rips_stats(const void* ws, int n, int batch, int maxdim, int cap1, size_t pool_bytes, int64_t* stats_host);

#ifdef __cplusplus
}
#endif
#endif
