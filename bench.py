#!/usr/bin/env python
"""bench.py -- UMAP + Rips(H0/H1) layers/sec on the C3 workload of BASELINE.json
("32 layers x 2k tokens x 4096-d, UMAP k=15 then Rips H0/H1"), synthetic activations (tda_multimodal_b200/workloads.py).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the CPU implementation of the path, all host cores: the real
                                                           umap-learn + ripser when they are importable, else the oracle port
  --scaling weak (default): 32 layers PER GPU;  --scaling strong: the 32 layers of config C3 split over the ranks
  --workload c3 (default) | c3c4: C3 followed by 256 bootstrap resamples (1000 points, Rips H0/H1) of every layer's cloud
                                  (the north-star target run; a "layer" then includes its resamples)

A step = one pass of the hot path over one batch: every rank runs the layer sweep (pairwise distances -> exact kNN
+ sigma/rho -> fuzzy graph -> spectral init -> SGD to 3-D -> Rips H0/H1) on its own layers and the ranks gather the
diagrams (NCCL) -- the only cross-GPU traffic of the path.
`value` times the sweep with the activations already resident in HBM; `e2e` times the same sweep from pinned HOST
buffers (H2D of the activations and D2H of embeddings + diagrams inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "umap_rips_h0h1_layers_per_sec"
UNIT = "layers/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the stage's dominant kernel, from `ncu --set full` captures of this
# command's kernels at the default shape (committed artefact; written by scripts/ncu_traffic.py from profiles/*.csv)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_ncu_dram_bytes.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c3c4"])
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--points", type=int, default=2000)
    ap.add_argument("--dim", type=int, default=4096)
    ap.add_argument("--neighbors", type=int, default=15)
    ap.add_argument("--resamples", type=int, default=256)
    ap.add_argument("--resample-points", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peaks", action="store_true", help="skip the on-box TF32 matmul peak measurement")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    return ap.parse_args()


def layers_of_rank(a, rank, world):
    """weak: every rank has its own a.layers layers (seeds offset by rank); strong: the a.layers layers of C3 dealt l -> l mod world."""
    if a.scaling == "weak":
        return list(range(a.layers)), 3000 + 1000 * rank
    return list(range(rank, a.layers, world)), 3000


def workload_config(a, world):
    per = a.layers if a.scaling == "weak" else f"{a.layers}/{world}"
    cfg = {"workload": f"C3: {a.layers} layers x {a.points} points x {a.dim}-d synthetic activations {'per GPU' if a.scaling == 'weak' else 'in total'}, "
                       f"UMAP(n_neighbors={a.neighbors}, n_components=3, min_dist=0.1, metric=cosine, 500 epochs) -> ripser(maxdim=1)",
           "layers_per_gpu": per, "points": a.points, "dim": a.dim, "n_neighbors": a.neighbors,
           "parallelism": f"layer-sharded x{world}, NCCL gather of diagrams", "l2": "inputs (1.05 GB per 32 layers) larger than L2"}
    if a.workload == "c3c4":
        cfg["workload"] += f" + C4: {a.resamples} bootstrap resamples of {a.resample_points} points per layer, Rips H0/H1 each"
        cfg["resamples_per_layer"] = a.resamples
    return cfg


# ------------------------------------------------------------------------------------------------------------------
# CPU side: used by cpu_baseline and by --impl reference
def _real_libs():
    """the real libraries first.  Returns (umap module, ripser function) or None."""
    try:
        import umap as _umap
        from ripser import ripser as _ripser
        if "tda_multimodal_b200" in (getattr(_umap, "__file__", "") or "") or "shims" in (getattr(_umap, "__file__", "") or ""):
            return None   # this repo's own shims are not the reference
        return _umap, _ripser
    except Exception:
        return None


def _cpu_layer(args):
    layer, n, d, k, n_layers = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from tda_multimodal_b200 import workloads
    X = workloads.c3_layer(layer, n=n, d=d, n_layers=n_layers)
    real = _real_libs()
    t0 = time.perf_counter()
    if real:
        Y = real[0].UMAP(n_neighbors=k, n_components=3, min_dist=0.1, metric="cosine", random_state=42).fit_transform(X)
        t1 = time.perf_counter()
        real[1](Y, maxdim=1)
    else:
        from oracle import umap_oracle as uo, rips as orips
        Y = uo.UMAPOracle(n_neighbors=k, n_components=3, min_dist=0.1, metric="cosine", random_state=42).fit_transform(X)
        t1 = time.perf_counter()
        orips.ripser(Y, maxdim=1, apparent=True)   # the oracle's fastest mode (2-3x the plain restatement of ripser's loop; same diagrams)
    t2 = time.perf_counter()
    return layer, t1 - t0, t2 - t1


def _cpu_warm():
    """numba JIT + library load, untimed (tiny cloud)."""
    from tda_multimodal_b200 import workloads
    X = workloads.c3_layer(0, n=120, d=64)
    real = _real_libs()
    if real:
        real[1](real[0].UMAP(n_neighbors=10, n_components=3, metric="cosine", random_state=42).fit_transform(X), maxdim=1)
        return
    from oracle import umap_oracle as uo, rips as orips
    Y = uo.UMAPOracle(n_neighbors=10, n_components=3, metric="cosine", random_state=42).fit_transform(X)
    orips.ripser(Y, maxdim=1, apparent=True)


def _cpu_kind():
    return ("reference", "umap-learn + ripser (real libraries)") if _real_libs() else \
        ("port", "oracle/umap_oracle.py (numba) + oracle/rips_oracle.cpp in its fastest mode (apparent-pair shortcut of Ripser 1.2 + diameter-windowed "
                 "working column: 2-3x faster than the plain restatement of ripser.py's loop, same diagrams); umap-learn and ripser are not installed in this image")


def cpu_baseline_serial(a, budget_s):
    """Single core, as the reference runs it (serial `for i in range(32)`, random_state set => serial numba SGD,
    ripser single-threaded): whole layers of the same workload until the time budget is used."""
    _cpu_warm()
    kind, what = _cpu_kind()
    done, t_total, per = 0, 0.0, []
    for layer in range(a.layers):
        _, tu, tr = _cpu_layer((layer, a.points, a.dim, a.neighbors, a.layers))
        done += 1
        t_total += tu + tr
        per.append((round(tu, 2), round(tr, 2)))
        if t_total >= budget_s:
            break
    return {"value": done / t_total, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"layers 0..{done - 1} of the {a.layers}-layer C3 workload, serial on one core ({what}); per-layer (umap_s, rips_s) = {per}"}


def run_reference(a):
    """--impl reference: the CPU implementation on all host cores (one process per layer, layers are independent)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, a.layers))
    kind, what = _cpu_kind()
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs, initializer=_pool_init) as pool:
        for _ in range(max(1, a.warmup)):
            pool.map(_warm_job, range(procs))          # warm-up = JIT / library load on every worker (tiny clouds)
        times = []
        budget_s = float(os.environ.get("TDA_REFERENCE_BUDGET_S", "240"))
        for _ in range(a.steps):
            # bounded sample: the first `procs` layers of the workload, one per core, concurrently (a step costs ~30 s of wall
            # clock: the slowest of those layers); the run is time-boxed so that any --steps ends within a few minutes
            t0 = time.perf_counter()
            pool.map(_cpu_layer, [(l, a.points, a.dim, a.neighbors, a.layers) for l in range(procs)])
            times.append(time.perf_counter() - t0)
            if sum(times) + times[-1] > budget_s:
                break
    total = sum(times)
    done = len(times)
    value = procs * done / total
    sample = (f"per step: layers 0..{procs - 1} of the C3 workload concurrently, one process per core ({what}); warm-up steps run tiny clouds "
              f"(numba JIT only); {done} of {a.steps} requested steps timed (time box {budget_s:.0f} s)")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": done, "warmup": a.warmup,
            "ms_per_step": 1e3 * total / done, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference", "config": workload_config(a, a.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _pool_init():
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)


def _warm_job(_):
    _cpu_warm()
    return 0


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/tda_clocks_{os.getpid()}.csv"
        self.active = False

    def start(self):
        """NVML from a thread of this process (the same counters nvidia-smi prints, without a second process polling the driver
        every 200 ms); falls back to `nvidia-smi -lms 200` when pynvml is missing."""
        self.samples, self.thread, self.stop_flag, self.active = [], None, False, False
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown if hasattr(pynvml, "nvmlClocksEventReasonHwSlowdown") else pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0)),
                    "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0)),
                    "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0))}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self.stop_flag:
                    try:
                        smp = (float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), mx, int(get_reasons(h)))
                        if self.active:          # the thread is started before the warm-up (NVML's first queries are slow and
                            self.samples.append(smp)   # take driver locks); only samples of the timed regions are kept
                    except Exception:
                        pass
                    time.sleep(0.1)
            self.bits = bits
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if getattr(self, "thread", None) is not None:
            self.stop_flag = True
            self.thread.join(2)
            if self.samples:
                reasons = sorted(nm for nm, bit in self.bits.items() if bit and any(s_[2] & bit for s_ in self.samples))
                out.update(sm_mhz=float(np.median([s_[0] for s_ in self.samples])), sm_max_mhz=float(max(s_[1] for s_ in self.samples)),
                           reasons=reasons, samples=len(self.samples), source="NVML (pynvml), 100 ms period, during the timed regions")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm), source="nvidia-smi -lms 200")
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def measure_tf32_peak(torch, dev):
    """Dense TF32 tensor-core peak of THIS box: torch.matmul fp32 with TF32 allowed, 8192^3; burst = best of
    10 single calls, sustained = back to back for ~2 s.  The distance GEMM issues kind::tf32 MMAs: this is its roofline."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        A = torch.randn((n, n), device=dev, dtype=torch.float32)
        B = torch.randn((n, n), device=dev, dtype=torch.float32)
        C = torch.empty((n, n), device=dev, dtype=torch.float32)
        for _ in range(3):
            torch.matmul(A, B, out=C)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(A, B, out=C); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(20, int(2000.0 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(A, B, out=C)
        e1.record(); torch.cuda.synchronize()
        flop = 2.0 * n ** 3
        return {"burst": flop / (best / 1e3) / 1e12, "sustained": flop * reps / (e0.elapsed_time(e1) / 1e3) / 1e12,
                "how": "torch.matmul fp32 (allow_tf32) 8192^3 on this box: best of 10 (burst), back to back for ~2 s (sustained)"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def algorithmic_work(stage, a, L, extra):
    """(bound, algorithmic bytes or flops) of one stage over L layers (DESIGN.md 'Measurement')."""
    n, d, k = a.points, a.dim, a.neighbors
    E = n * (n - 1) // 2
    if stage == "pdist_gemm":
        return "tensor", 2.0 * L * n * n * d
    table = {
        "pdist_prep": L * (4.0 * n * d + 8.0 * n * d),                 # read X, write hi + lo
        "knn_smooth": L * (4.0 * n * n + 8.0 * n * k + 8.0 * n),       # read D, write idx+dist, sigma+rho
        "fuzzy_graph": L * (12.0 * n * k + 8.0 * n + 16.0 * 2 * n * k),
        # Lanczos, m = 96 steps with two Gram-Schmidt passes over j + 2 vectors each (dot + update: 4 vector reads of n floats per
        # vector and pass) + the sparse matrix (8 B per entry, nnz <= 2nk) once per step -- all of it shared-memory traffic
        "spectral_init": L * (16.0 * n * (96 * 97 / 2 + 2 * 96) + 8.0 * 2 * n * k * 96),
        # SGD: per fired edge two endpoints + ~5 negative samples read (16 B each) and one position written: 124 B
        "umap_sgd": 124.0 * extra.get("sgd_fired_per_layer", 0.0) * L,
        "rips_pdist": L * (12.0 * n + 4.0 * n * n),
        "rips_edge_sort": L * (4.0 * n * n + 16.0 * E + 16.0 * E),     # read dm, radix sort key+payload, rank scatter
        # H0 from the sorted edge list (boruvka_chunked_kernel): at most one pass over the 4-byte end points of the E edges, plus a
        # few repeated passes over the chunks that hold forest edges (not counted)
        "rips_h0": L * 4.0 * E,
        "rips_apparent": L * 8.0 * (E - n + 1) * (n - 2),
        # residual reduction: 16 B (endpoints, apex, parents) per row swept / substituted; per heavy row two V rows + two
        # adjacency rows (4 * n/8 B)
        "rips_reduce": L * (16.0 * extra.get("reduce_rows_per_layer", 0.0) + 4.0 * (n / 8.0) * extra.get("reduce_heavy_rows_per_layer", 0.0)),
    }
    return "hbm", table[stage]


def union_ms(spans):
    """total length of the union of [start, end) intervals"""
    tot, cur_a, cur_b = 0.0, None, None
    for a_, b_ in sorted(spans):
        if cur_b is None or a_ > cur_b:
            if cur_b is not None:
                tot += cur_b - cur_a
            cur_a, cur_b = a_, b_
        else:
            cur_b = max(cur_b, b_)
    if cur_b is not None:
        tot += cur_b - cur_a
    return tot


def run_b200(a):
    import torch
    import torch.distributed as dist
    from tda_multimodal_b200 import _lib, pipeline, workloads
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the b200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    my_layers, seed = layers_of_rank(a, rank, world)
    nl = len(my_layers)
    Xh = torch.empty((max(nl, 1), a.points, a.dim), dtype=torch.float32).pin_memory()
    if nl:
        workloads.c3_layers(n_layers=a.layers, n=a.points, d=a.dim, seed=seed, layers=my_layers, out=Xh.numpy())
    Xd = Xh.to(dev)
    layers_per_step = a.layers * world if a.scaling == "weak" else a.layers

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def units_of(out):
        """per-layer diagram sets of one sweep (+ the bootstrap resamples of every layer for c3c4)"""
        units = [r["dgms"] for r in out["results"]]
        if a.workload == "c3c4" and nl:
            Y = out["embedding"]
            Yd = Y if isinstance(Y, torch.Tensor) else torch.from_numpy(Y).to(dev)
            boot = pipeline.bootstrap_rips(Yd, n_resamples=a.resamples, size=a.resample_points, layer_ids=my_layers)
            units += [r["dgms"] for per_layer in boot for r in per_layer]
        return units

    def step_resident():
        out = pipeline.layer_sweep(Xd[:nl], n_neighbors=a.neighbors, return_embedding=True) if nl else {"results": [], "embedding": None}
        return gather(units_of(out))

    def step_e2e():
        out = pipeline.layer_sweep_host(Xh[:nl], device=dev, n_neighbors=a.neighbors) if nl else {"results": [], "embedding": np.zeros((0, a.points, 3), np.float32)}
        return gather(units_of(out)), out

    def gather(units):
        if world == 1:
            return units
        counts, payload = pipeline.pack_diagrams([{"dgms": u} for u in units])
        sizes = torch.tensor([counts.shape[0], payload.shape[0]], device=dev)
        all_sizes = [torch.empty_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes)
        mc, mp = max(int(s[0]) for s in all_sizes), max(int(s[1]) for s in all_sizes)
        cpad = torch.zeros((max(mc, 1), 2), dtype=torch.int32, device=dev)
        cpad[:counts.shape[0]] = torch.from_numpy(counts).to(dev)
        ppad = torch.zeros((max(mp, 1), 2), dtype=torch.float32, device=dev)
        ppad[:payload.shape[0]] = torch.from_numpy(payload).to(dev)
        gc = [torch.empty_like(cpad) for _ in range(world)]
        gp = [torch.empty_like(ppad) for _ in range(world)]
        dist.all_gather(gc, cpad)
        dist.all_gather(gp, ppad)
        if rank != 0:
            return None
        full = []
        for r in range(world):
            nc, npay = int(all_sizes[r][0]), int(all_sizes[r][1])
            full += pipeline.unpack_diagrams(gc[r][:nc].cpu().numpy(), gp[r][:npay].cpu().numpy())
        return full

    each = {}

    def timed(fn, steps, tag):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        evs[0].record()
        for k in range(steps):
            res = fn()
            evs[k + 1].record()
        barrier()
        ms = torch.tensor([evs[0].elapsed_time(evs[-1])], dtype=torch.float64, device=dev)
        each[tag] = [round(evs[k].elapsed_time(evs[k + 1]), 2) for k in range(steps)]
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(a.warmup):
        step_resident()
    # CPython's cyclic collector scans the whole heap of a process with torch loaded in 15-60 ms -- a third of a step -- and a pass
    # that lands in the unpacking of a step's last group delays that step (profiles/r02_step_variance.txt; one 57.9 ms step in the
    # r02m line).  Everything alive after the warm-up is moved to the permanent generation, as a long-running service would do
    # after start-up: later passes only look at what the steps themselves allocate.
    import gc as _gc
    _gc.collect()
    _gc.freeze()
    sampler.active = True
    L.tda_launch_count_reset()
    L.tda_stage_timing_reset()
    L.tda_stage_timing_enable(1)
    ms_res, dgms = timed(step_resident, a.steps, "resident")
    launches = int(L.tda_launch_count())
    stages = _lib.stage_times()
    timeline = _lib.stage_timeline()
    L.tda_stage_timing_enable(0)
    sampler.active = False
    for _ in range(a.warmup):        # the host-input path has its own streams (and allocator pools): warm it up as well, untimed
        step_e2e()
    sampler.active = True
    ms_e2e, (dg2, out2) = timed(step_e2e, a.steps, "e2e")
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        value = layers_per_step * a.steps / (ms_res / 1e3)
        e2e_value = layers_per_step * a.steps / (ms_e2e / 1e3)
        d2h = int(out2["embedding"].nbytes + sum(sum(d.shape[0] * 8 for d in r["dgms"]) for r in out2["results"]) + 16 * nl)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        tf32 = None
        if not a.no_peaks:
            try:
                tf32 = measure_tf32_peak(torch, dev)
            except Exception as ex:
                tf32 = {"error": repr(ex)}
        if tf32 and "sustained" in tf32:
            tf32_peak, tf32_src = tf32["sustained"], "measured on this box by this run (cuBLAS TF32 8192^3, sustained)"
        else:
            tf32_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))) / 2.0
            tf32_src = "derived: measured bf16 dense / 2"
        # ---- device-side counters of one more sweep: the reduction's rows, the SGD's fired edges
        extra = {}
        try:
            from tda_multimodal_b200 import umap_
            nsub = min(nl, 8)
            dm = pipeline.pdist_lowdim(torch.from_numpy(out2["embedding"][:nsub]).to(dev))
            st = pipeline.rips_batch(dm, maxdim=1, want_stats=True)
            rows_key = "rows_substituted" if "rows_substituted" in st[0]["stats"] else "rows_streamed"
            extra["reduce_rows_per_layer"] = float(sum(r["stats"][rows_key] for r in st)) / nsub
            extra["reduce_heavy_rows_per_layer"] = float(sum(r["stats"]["heavy_rows"] for r in st)) / nsub
            extra["rips_stats_per_layer"] = {k: round(sum(r["stats"][k] for r in st) / nsub, 1) for k in st[0]["stats"]
                                             if not k.startswith(("cyc_", "spare", "max_v", "barrier"))}
            extra["rips_reducer"] = _lib.rips_reducer()
            _, state = umap_.umap_fit_batch(Xd[:min(nl, 4)], n_neighbors=a.neighbors, n_components=3, metric="cosine", random_state=42, return_state=True)
            eps = state["eps"]
            n_ep = int(state["n_epochs"])
            fired = torch.where(eps > 0, torch.floor(n_ep / eps.clamp(min=1.0)), torch.zeros_like(eps)).sum().item()
            extra["sgd_fired_per_layer"] = float(fired) / min(nl, 4)     # directed entries fired over all epochs (counted from the schedule)
        except Exception as ex:  # optional evidence
            extra["stats_error"] = repr(ex)
        traffic = {}
        try:
            traffic = json.load(open(TRAFFIC_FILE))
        except Exception:
            pass
        default_shape = (a.points, a.dim, a.neighbors) == (2000, 4096, 15)
        # ---- roofline table: every stage; spans of the two chunk streams overlap, so `wall_ms` (union of the stage's spans) is
        # what the stage occupies on the clock and `sum_ms` what its launches add up to
        n_done = nl * a.steps
        spans_by_stage = {}
        for name, t0_, t1_ in timeline:
            spans_by_stage.setdefault(name, []).append((t0_, t1_))
        table = {}
        for sname, (ms_sum, calls) in stages.items():
            if calls == 0:
                continue
            kind, work = algorithmic_work(sname, a, n_done, extra)
            peak = tf32_peak if kind == "tensor" else hbm_peak
            unit = "TFLOP/s" if kind == "tensor" else "GB/s"
            scale = 1e12 if kind == "tensor" else 1e9
            achieved = work / (ms_sum / 1e3) / scale if ms_sum > 0 else 0.0
            ent = {"bound": "tensor" if kind == "tensor" else "hbm", "sum_ms_per_step": round(ms_sum / a.steps, 3),
                   "wall_ms_per_step": round(union_ms(spans_by_stage.get(sname, [])) / a.steps, 3), "launches_per_step": calls / a.steps,
                   "avg_launch_ms": round(ms_sum / calls, 4), "work_per_step": work / a.steps, "achieved": achieved, "peak": peak, "unit": unit,
                   "frac": achieved / peak if peak else None}
            if sname == "rips_apparent":
                ent["note"] = ("algorithmic bytes = SURVEY 8d's 8(E-n+1)(n-2) (every cofacet of every column); the kernel stops at the first "
                               "cofacet of equal diameter (top-down scan, ~3 of 63 row chunks per edge), so frac > 1 is expected here")
            if sname == "pdist_gemm":
                ent["note"] = "useful FLOPs 2*L*n*n*d of the delivered [n,n] matrices; the kernel computes the tiles on and above the diagonal (53 %) and stores each twice"
            if sname in ("umap_sgd", "spectral_init", "rips_reduce"):
                ent["note"] = "state is shared-memory / L2 resident: the algorithmic bytes do not reach HBM (see traffic); the kernel is bound by barrier latency and instruction issue (profiles/r02b_*)"
            tr = traffic.get(sname) if default_shape else None
            if tr:
                ent["traffic"] = tr.get("dram_bytes_per_launch")
                ent["traffic_source"] = tr.get("source")
            table[sname] = ent
        dom = max(table, key=lambda s_: table[s_]["sum_ms_per_step"])
        D = table[dom]
        tot_ms = sum(v[0] for v in stages.values())
        roofline = {"kernel": dom, "bound": D["bound"], "achieved": D["achieved"], "peak": D["peak"], "unit": D["unit"], "frac": D["frac"],
                    "traffic": D.get("traffic"), "traffic_source": D.get("traffic_source"),
                    "peak_source": tf32_src if D["bound"] == "tensor" else peak_src, "avg_launch_ms": D["avg_launch_ms"],
                    "share_of_device_stage_time": D["sum_ms_per_step"] * a.steps / tot_ms if tot_ms else None,
                    "stages_ms_per_step": {s_: round(v[0] / a.steps, 3) for s_, v in stages.items()},
                    "stages": table,
                    "peaks": {"hbm_gbs": hbm_peak, "hbm_source": peak_src, "tf32_tflops": tf32_peak, "tf32_source": tf32_src, "tf32_measurement": tf32},
                    "note": "stage times are CUDA-event spans on the launching streams; the two chunk streams overlap, so sum_ms adds up to more "
                            "than the step and wall_ms is the union of a stage's spans"}
        # the two kernel-level figures BASELINE.json's metric names next to layers/s, each timed alone (cold L2)
        secondary = {}
        try:
            from tda_multimodal_b200 import umap_
            half = max(1, nl // 2)                      # one chunk of the sweep
            Xc = Xd[:half]
            flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

            def alone(fn, reps=5):
                fn()
                ts = []
                for _ in range(reps):
                    flush.fill_(1)
                    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s0.record(); out = fn(); s1.record(); torch.cuda.synchronize()
                    ts.append(s0.elapsed_time(s1))
                return sorted(ts)[len(ts) // 2], out
            ms_pd, Dm = alone(lambda: umap_.distance_matrix(Xc, metric="cosine"))
            ms_kn, _ = alone(lambda: umap_.knn_smooth(Dm, a.neighbors))
            n_, d_, k_ = a.points, a.dim, a.neighbors
            tn = (n_ + 127) // 128
            tile_frac = (tn * (tn + 1) / 2) / (tn * tn)
            secondary["alone"] = {
                "layers": half,
                "pdist_ms": ms_pd, "pdist_useful_tflops": 2.0 * half * n_ * n_ * d_ / (ms_pd / 1e3) / 1e12,
                # issued MMA work: 3 products (3xTF32) on the tiles actually computed (symmetric: the T(T+1)/2 tiles on and above the diagonal)
                "pdist_tiles_computed_frac": tile_frac,
                "pdist_issued_tflops": 3.0 * tile_frac * 2.0 * half * tn * tn * 128 * 128 * d_ / (ms_pd / 1e3) / 1e12,
                "pdist_issued_frac_of_tf32_peak": 3.0 * tile_frac * 2.0 * half * tn * tn * 128 * 128 * d_ / (ms_pd / 1e3) / 1e12 / tf32_peak,
                "knn_ms": ms_kn, "knn_hbm_gbs": half * (4.0 * n_ * n_ + 8.0 * n_ * k_ + 8.0 * n_) / (ms_kn / 1e3) / 1e9,
                "knn_frac_of_hbm_peak": half * (4.0 * n_ * n_ + 8.0 * n_ * k_ + 8.0 * n_) / (ms_kn / 1e3) / 1e9 / hbm_peak,
                "note": "each call timed alone incl. operand prep (pdist) / memset + sigma floor (kNN), 512 MB L2 flush before every rep"}
            del Dm, flush
        except Exception as ex:  # optional evidence
            secondary["alone_error"] = repr(ex)
        roofline["secondary"] = secondary
        for k_ in ("rips_stats_per_layer", "rips_reducer", "sgd_fired_per_layer", "stats_error"):
            if k_ in extra:
                roofline[k_] = extra[k_]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_res / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(a, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nl * a.points * a.dim * 4), "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / a.steps},
                "gpu_launches": launches, "clocks": clocks, "ms_each_step_rank0": each, "roofline": roofline}
        if dgms is not None:
            line["diagram_sets_gathered_per_step"] = len(dgms)
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_serial(a, a.cpu_budget_s)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
