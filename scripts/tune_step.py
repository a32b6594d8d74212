"""A/B of the tuning options on the C3 step (32 layers x 2000 x 4096 resident in HBM): the workload is generated once, then every
configuration (chunks of the sweep + library options) runs 2 warm-up and 6 timed steps; prints min / median ms per step.
  python scripts/tune_step.py [name=value,... ...]     e.g.  chunks=2 chunks=4,sgd_cluster=8,sgd_tile=8
Without arguments: the built-in list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import _lib, pipeline, workloads

DEFAULTS = {"sgd_cluster": 0, "sgd_tile": 16, "rips_cluster": 0, "spectral_cluster": 8, "rips_apparent_rows": 1, "rips_wc_max_rows": 262144, "rips_h0_chunked": 1}   # (tail_* keys: options of the last group only)
BUILTIN = ["chunks=2", "chunks=3", "chunks=4", "chunks=2,sgd_tile=8", "chunks=2,sgd_cluster=8,sgd_tile=8", "chunks=4,sgd_cluster=8,sgd_tile=8",
           "chunks=2,rips_cluster=8", "chunks=4,rips_cluster=8", "chunks=4,sgd_tile=8", "chunks=4,sgd_tile=8,rips_cluster=8", "chunks=1"]


def main():
    cfgs = sys.argv[1:] or BUILTIN
    Xh = torch.from_numpy(workloads.c3_layers(n_layers=32)).pin_memory()
    X = Xh.cuda()
    L = _lib.lib()
    e2e = os.environ.get("TUNE_E2E") == "1"     # time the sweep from pinned host memory (H2D inside, results to the host)
    if e2e:
        Xdev = X
        class _HostSweep:
            @staticmethod
            def layer_sweep(_, chunks=None):
                return pipeline.layer_sweep_host(Xh, chunks=chunks)
        sweep = _HostSweep.layer_sweep
    else:
        sweep = pipeline.layer_sweep
    for cfg in cfgs:
        kv = dict(DEFAULTS)
        chunks = 2
        tail = {}
        for item in cfg.split(","):
            k, v = item.split("=")
            if k == "chunks":
                chunks = int(v)
            elif k == "split":
                chunks = [int(x) for x in v.split("-")]
            elif k.startswith("tail_"):
                tail[k[5:]] = int(v)
            else:
                kv[k] = int(v)
        for k, v in kv.items():
            _lib.set_option(k, v)
        pipeline.TAIL_OPTIONS = tail
        for _ in range(2):
            sweep(X, chunks=chunks)
        torch.cuda.synchronize()
        ts, host, mallocs = [], [], []
        import time, gc
        if os.environ.get("TUNE_GC") == "off":
            gc.collect(); gc.disable()
        gcs0 = [g["collections"] for g in gc.get_stats()]
        timing_on = os.environ.get("TUNE_STAGE_TIMING") == "1"    # stage timer enabled inside the timed steps, as bench.py does
        if timing_on:
            L.tda_stage_timing_reset(); L.tda_stage_timing_enable(1)
        for it_ in range(int(os.environ.get("TUNE_STEPS", "6"))):
            if timing_on and it_ % 8 == 0:
                L.tda_stage_timing_reset()
            m0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            sweep(X, chunks=chunks)
            e1.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            host.append(1e3 * (t1 - t0))
            mallocs.append(torch.cuda.memory_stats().get("num_device_alloc", 0) - m0)
        if timing_on:
            L.tda_stage_timing_enable(0)
        if os.environ.get("TUNE_VERBOSE"):
            print("   gpu ms", [round(t, 1) for t in ts], "cudaMallocs", sum(mallocs), "gc collections per generation", [g["collections"] - a for g, a in zip(gc.get_stats(), gcs0)], flush=True)
        if os.environ.get("TUNE_TIMELINE"):   # timelines of the fastest and the slowest of 8 more steps
            L.tda_stage_timing_enable(1)
            runs = []
            for _ in range(8):
                L.tda_stage_timing_reset()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); sweep(X, chunks=chunks); e1.record(); torch.cuda.synchronize()
                runs.append((e0.elapsed_time(e1), _lib.stage_timeline()))
            runs.sort(key=lambda r: r[0])
            for tag, (ms, spans) in (("fastest", runs[0]), ("slowest", runs[-1])):
                t00 = min(s_[1] for s_ in spans)
                print(f"   {tag}: {ms:.1f} ms: " + "  ".join(f"{nm.replace('rips_', 'r_')}[{a - t00:.1f},{b - t00:.1f}]" for nm, a, b in sorted(spans, key=lambda s_: s_[1]) if b - a > 0.4), flush=True)
            L.tda_stage_timing_enable(0)
        L.tda_stage_timing_reset(); L.tda_stage_timing_enable(1)
        sweep(X, chunks=chunks)
        torch.cuda.synchronize()
        st = _lib.stage_times()
        L.tda_stage_timing_enable(0)
        stages = " ".join(f"{k.replace('rips_', 'r_').replace('pdist_', 'p_')}={v[0]:.1f}" for k, v in st.items() if v[0] >= 0.5)
        print(f"{cfg:48s} min {min(ts):7.2f}  med {float(np.median(ts)):7.2f}  max {max(ts):7.2f} ms | {stages}", flush=True)


if __name__ == "__main__":
    main()
