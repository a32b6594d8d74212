// capi.cu -- library-wide pieces of the C ABI (version, error string, launch counter, stage timer).
#include "common.cuh"
#include "launch_count.cuh"
#include "../../include/tda_b200.h"
#include <vector>

namespace tda {
char* tls_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
int64_t& launch_counter() {
  static thread_local int64_t c = 0;
  return c;
}

// ---- stage timer: per thread, a list of (stage, start event, stop event); events are pooled and reused
struct StageRec { int stage; cudaEvent_t e0, e1; };
struct StageSpan { int stage; double t0, t1; };
struct StageState {
  bool on = false;
  cudaEvent_t ref = nullptr;          // recorded at reset: the origin of the timeline
  std::vector<StageSpan> spans;       // (stage, start, end) in ms since `ref`, in the order the stages were enqueued
  std::vector<StageRec> recs;
  std::vector<cudaEvent_t> pool;
  double ms[STAGE_COUNT] = {0};
  int64_t calls[STAGE_COUNT] = {0};
  int open[STAGE_COUNT];
  StageState() { for (int i = 0; i < STAGE_COUNT; ++i) open[i] = -1; }
};
static StageState& stage_state() {
  static thread_local StageState s;
  return s;
}
bool stage_timing_on() { return stage_state().on; }
static cudaEvent_t take_event(StageState& S) {
  cudaEvent_t e = nullptr;
  if (!S.pool.empty()) { e = S.pool.back(); S.pool.pop_back(); }
  else cudaEventCreate(&e);
  return e;
}
void stage_begin(int stage, cudaStream_t s) {
  StageState& S = stage_state();
  StageRec r{stage, take_event(S), take_event(S)};
  cudaEventRecord(r.e0, s);
  S.open[stage] = (int)S.recs.size();
  S.recs.push_back(r);
}
void stage_end(int stage, cudaStream_t s) {
  StageState& S = stage_state();
  if (S.open[stage] < 0) return;
  cudaEventRecord(S.recs[S.open[stage]].e1, s);
  S.open[stage] = -1;
}
static void stage_collect(StageState& S) {
  for (StageRec& r : S.recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      S.ms[r.stage] += (double)ms;
      S.calls[r.stage] += 1;
      float a = 0.f;
      if (S.ref && cudaEventElapsedTime(&a, S.ref, r.e0) == cudaSuccess && S.spans.size() < 65536) S.spans.push_back({r.stage, (double)a, (double)a + (double)ms});
    }
    S.pool.push_back(r.e0);
    S.pool.push_back(r.e1);
  }
  S.recs.clear();
  (void)cudaGetLastError();
}

// ---- tuning options (process-wide; the library reads no environment variables)
struct OptionDef { const char* name; long long value; };
static OptionDef g_options[] = {
    {"rips_reducer", 0},       // 0 sweep2 (substitute by rank, verify by window; default), 1 row sweep with sequential resolver, 2 row sweep substitute-then-verify, 3 key bitset
    {"rips_w0", 1024},         // sweep2: first window of a column (rows)
    {"rips_wsparse", 8192},    // sweep2: largest window in sparse mode
    {"rips_wmax", 32768},      // sweep2: largest window in dense mode (<= 65472)
    {"rips_dense_min", 64},    // sweep2: a window with >= max(dense_min, rows / dense_div) heavy rows switches the column to dense mode
    {"rips_dense_div", 8},
    {"rips_warp_engine", 1},   // sweep2: short columns are reduced by single warps first, speculatively, and committed in order
    {"rips_wc_max_rows", 262144}, // sweep2: rows a warp sweeps before it hands its column to the cluster engine
    {"rips_cluster", 0},       // sweep2: CTAs per cloud (thread-block cluster: 1, 2, 4 or 8; 0 = auto: 8 for up to 4 clouds per launch, else 4; halved while batch * cluster > 2 * SMs)
    {"rips_h0_chunked", 1},    // H0 in one launch from the sorted edge list (Boruvka / Kruskal by chunks, n <= 16384); 0: Boruvka rounds on the rank matrix
    {"rips_apparent_rows", 1}, // apparent pairs by matrix row (one CTA per vertex, its rank row in shared memory); 0: one warp per edge by rank
    {"sweep_exclusive", 0},    // reducers 1/2: ask for the whole shared memory of the SM
    {"sgd_mode", 0},           // 0 deterministic kernels (cluster per cloud for fit, warp per point for transform), 3 per-epoch kernels with float atomics
    {"spectral_cluster", 8},   // CTAs per cloud of the Lanczos kernel for connected graphs (2, 4, 8; 0: the one-CTA kernel)
    {"sgd_cluster", 0},        // CTAs per cloud of the deterministic fit kernel (1, 2, 4 or 8; 0 = auto: 8 for up to 4 clouds per launch, else 4)
    {"sgd_tile", 16},          // vertices per warp task of that kernel (1..16; tasks are handed out dynamically)
    {"knn_loads", 8},          // 16-byte loads per lane in flight in the k <= 16 kNN kernel (8 or 16)
    {"rips_debug", 0},         // sweep2 prints one line of counters and phase cycles per cloud (diagnostics)
    {"spectral_debug", 0},     // print the phase cycles of the cluster Lanczos kernel for cloud 0 (diagnostics)
    {"debug_sync", 0},         // synchronise after every kernel of tda_rips_h2 (fault location)
    {"h2_stats", 0},           // print the H2 reducer's device counters to stderr
};
long long option(const char* name) {
  for (const OptionDef& o : g_options)
    if (strcmp(o.name, name) == 0) return o.value;
  return 0;
}
}  // namespace tda

extern "C" int tda_set_option(const char* name, long long value) {
  if (!name) return tda::set_error(tda::TDA_ERR_INVALID, "tda_set_option: null name");
  for (tda::OptionDef& o : tda::g_options)
    if (strcmp(o.name, name) == 0) { o.value = value; return tda::TDA_OK; }
  return tda::set_error(tda::TDA_ERR_INVALID, "tda_set_option: unknown option '%s'", name);
}
extern "C" long long tda_get_option(const char* name) {
  if (!name) return -1;
  for (const tda::OptionDef& o : tda::g_options)
    if (strcmp(o.name, name) == 0) return o.value;
  return -1;
}

extern "C" int tda_version(void) { return 200; }
extern "C" const char* tda_last_error(void) { return tda::tls_error_buffer(); }
extern "C" int64_t tda_launch_count(void) { return tda::launch_counter(); }
extern "C" void tda_launch_count_reset(void) { tda::launch_counter() = 0; }

extern "C" void tda_stage_timing_enable(int on) {
  tda::StageState& S = tda::stage_state();
  S.on = on != 0;
  // events are created here, not between the launches of a timed step: cudaEventCreate goes through the kernel driver, and on a
  // shared host such calls were seen to stall the launching thread for tens of milliseconds (profiles/r02_step_variance.txt)
  if (S.on)
    while (S.pool.size() < 4096) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess) break;
      S.pool.push_back(e);
    }
}
extern "C" void tda_stage_timing_reset(void) {
  tda::StageState& S = tda::stage_state();
  tda::stage_collect(S);
  for (int i = 0; i < tda::STAGE_COUNT; ++i) { S.ms[i] = 0; S.calls[i] = 0; }
  S.spans.clear();
  if (!S.ref) cudaEventCreate(&S.ref);
  cudaEventRecord(S.ref, 0);   // legacy default stream: ordered after everything enqueued so far
  (void)cudaGetLastError();
}
extern "C" int tda_stage_timeline_read(int* stage_out, double* start_ms_out, double* end_ms_out, int cap) {
  tda::StageState& S = tda::stage_state();
  tda::stage_collect(S);
  const int k = (int)S.spans.size() < cap ? (int)S.spans.size() : cap;
  for (int i = 0; i < k; ++i) {
    if (stage_out) stage_out[i] = S.spans[i].stage;
    if (start_ms_out) start_ms_out[i] = S.spans[i].t0;
    if (end_ms_out) end_ms_out[i] = S.spans[i].t1;
  }
  return (int)S.spans.size();
}
extern "C" int tda_stage_timing_read(double* ms_out, int64_t* calls_out, int n) {
  tda::StageState& S = tda::stage_state();
  tda::stage_collect(S);
  for (int i = 0; i < n && i < tda::STAGE_COUNT; ++i) {
    if (ms_out) ms_out[i] = S.ms[i];
    if (calls_out) calls_out[i] = S.calls[i];
  }
  return tda::STAGE_COUNT;
}
