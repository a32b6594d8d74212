"""CPU baseline table of BASELINE.md section 3, for the five configurations of BASELINE.json (C1..C5), on the host cores of the
machine it runs on.  TEST / MEASUREMENT INFRASTRUCTURE: it times the CPU oracle (or the real umap-learn + ripser when a box has
them: bench._real_libs), never the product path, and needs no GPU.

    python scripts/cpu_baseline_table.py [--timeout 600] [--out profiles/r02_cpu_baseline_table.json]

What is timed (wall clock, `time.perf_counter`, one untimed warm-up for the numba JIT, inputs = the seeded workloads of
tda_multimodal_b200/workloads.py that the GPU runs use):
  C1  one layer of the 48-sample set: UMAP(k=6, cosine, 3-D) + ripser(maxdim=1)            median of 3 layers
  C3  one layer 2000 x 4096: UMAP(k=15, cosine, 3-D) + ripser(maxdim=1)                     median of 3 layers
  C4  one bootstrap resample (1000 of the 2000 points of a C3 embedding): ripser(maxdim=1)  median of 3 resamples
  C2  2000-point torus in 4096-d: euclidean pdist + ripser(maxdim=2)        child process, --timeout, 24 GB address-space cap
  C5  100 000 x 4096: cosine pdist + row-argsort kNN (timed on a 4096-row block, extrapolated), ripser(maxdim=1) on 10 000
      greedy landmarks of the 3-D cloud                                       child process, --timeout, 24 GB address-space cap
plus the per-stage figures: sklearn pairwise_distances GFLOP/s (all BLAS threads) and the argsort kNN GB/s over the N x N matrix.
Serial layers/s = 1 / (umap_s + rips_s); the "x cpu_count processes" figure is bench.py --impl reference (not repeated here).
"""
import argparse
import json
import multiprocessing as mp
import os
import resource
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (for _real_libs / _cpu_kind: the same choice of implementation as the bench's CPU legs)
from tda_multimodal_b200 import workloads  # noqa: E402


def _impl():
    real = bench._real_libs()
    if real:
        return (lambda **kw: real[0].UMAP(**kw)), real[1]
    from oracle import umap_oracle as uo, rips as orips
    return (lambda **kw: uo.UMAPOracle(**kw)), (lambda X, **kw: orips.ripser(X, apparent=True, **kw))   # fastest mode of the port, as in bench.py


def _med(xs):
    return float(np.median(xs))


def _child(fn, args, q, mem_gb):
    lim = int(mem_gb * (1 << 30))
    resource.setrlimit(resource.RLIMIT_AS, (lim, lim))
    try:
        q.put(("ok", fn(*args)))
    except MemoryError:
        q.put(("oom", None))
    except Exception as e:  # noqa: BLE001
        q.put(("error", repr(e)))


def bounded(fn, args, timeout_s, mem_gb=24.0):
    """fn(*args) in a child process with a wall-clock limit and an address-space cap."""
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=_child, args=(fn, args, q, mem_gb))
    t0 = time.perf_counter()
    p.start()
    p.join(timeout_s)
    if p.is_alive():
        p.terminate()
        p.join()
        return {"status": f"> {timeout_s:.0f} s (timeout)", "seconds": None}
    dt = time.perf_counter() - t0
    try:
        status, val = q.get(timeout=5)
    except Exception:  # noqa: BLE001  (killed by the kernel / bad_alloc abort in the C++ oracle)
        return {"status": f"out of memory or crashed after {dt:.0f} s (address-space cap {mem_gb:.0f} GB, exit code {p.exitcode})", "seconds": None}
    if status != "ok":
        return {"status": status if status != "oom" else f"out of memory after {dt:.0f} s (cap {mem_gb:.0f} GB)", "seconds": None, "detail": val}
    return {"status": "ok", **val}


def c2_job(n):
    from sklearn.metrics import pairwise_distances
    _, rip = _impl()
    X = workloads.c2_torus(n=n)
    t0 = time.perf_counter()
    dm = pairwise_distances(X, metric="euclidean").astype(np.float32)
    t1 = time.perf_counter()
    r = rip(dm, maxdim=2, distance_matrix=True)
    t2 = time.perf_counter()
    return {"seconds": t2 - t0, "pdist_s": t1 - t0, "rips_s": t2 - t1, "rows": [int(len(d)) for d in r["dgms"]]}


def c5_landmark_job(n_land):
    from oracle import rips as orips
    _, rip = _impl()
    rng = np.random.default_rng(5001)
    # stand-in for the 3-D UMAP embedding of the 100k cloud (the CPU UMAP of 100k x 4096 does not fit the time box): 16 blobs
    lab = rng.integers(0, 16, 100000)
    Y = (rng.normal(0, 4.0, (16, 3))[lab] + rng.normal(0, 0.6, (100000, 3))).astype(np.float32)
    t0 = time.perf_counter()
    idx, _ = orips.greedy_permutation_points(Y, n_land)
    t1 = time.perf_counter()
    r = rip(Y[idx], maxdim=1)
    t2 = time.perf_counter()
    return {"seconds": t2 - t0, "greedy_perm_s": t1 - t0, "rips_s": t2 - t1, "rows": [int(len(d)) for d in r["dgms"]]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--timeout", type=float, default=600.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_cpu_baseline_table.json"))
    ap.add_argument("--skip-big", action="store_true", help="skip C2 and the C5 landmark Rips (the two time-boxed jobs)")
    a = ap.parse_args()
    mk_umap, rip = _impl()
    kind, what = bench._cpu_kind()
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:  # noqa: BLE001
        pass
    ram_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2**30
    out = {"host": {"cpu_count": os.cpu_count(), "model": cpu, "ram_gb": round(ram_gb, 1)}, "kind": kind, "implementation": what,
           "method": "wall clock around fit_transform + ripser only; numba JIT warmed on a tiny cloud first; seeded workloads of workloads.py"}
    bench._cpu_warm()

    # ---- C1: 48 samples x 4096-d per layer, k = 6 (debug_tda_pipeline.py:96-110)
    acts = workloads.c1_activations()
    ids = sorted(acts)
    tu, tr = [], []
    for layer in (0, 15, 31):
        X = np.stack([acts[i]["activations"][f"layer_{layer}"] for i in ids]).astype(np.float32)
        t0 = time.perf_counter()
        Y = mk_umap(n_neighbors=6, n_components=3, min_dist=0.1, metric="cosine", random_state=42).fit_transform(X)
        t1 = time.perf_counter()
        rip(Y, maxdim=1)
        t2 = time.perf_counter()
        tu.append(t1 - t0)
        tr.append(t2 - t1)
    out["C1"] = {"points": len(ids), "umap_s": _med(tu), "rips_s": _med(tr), "layers_per_s_serial": 1.0 / (_med(tu) + _med(tr)), "layers": [0, 15, 31]}
    print("C1", out["C1"], flush=True)

    # ---- C3: one layer 2000 x 4096, k = 15
    tu, tr, Y3 = [], [], None
    for layer in (0, 1, 2):
        X = workloads.c3_layer(layer, n=2000, d=4096, n_layers=32)
        t0 = time.perf_counter()
        Y = mk_umap(n_neighbors=15, n_components=3, min_dist=0.1, metric="cosine", random_state=42).fit_transform(X)
        t1 = time.perf_counter()
        rip(Y, maxdim=1)
        t2 = time.perf_counter()
        tu.append(t1 - t0)
        tr.append(t2 - t1)
        if layer == 0:
            Y3 = np.asarray(Y)
    out["C3"] = {"points": 2000, "umap_s": _med(tu), "rips_s": _med(tr), "per_layer": [(round(u, 2), round(r, 2)) for u, r in zip(tu, tr)],
                 "layers_per_s_serial": 1.0 / (_med(tu) + _med(tr)), "layers": [0, 1, 2]}
    print("C3", out["C3"], flush=True)

    # ---- C4: one 1000-point resample of the layer-0 embedding
    idx = workloads.c4_resample_indices(0, n_points=2000, n_resamples=3, size=1000)
    tr = []
    for r in range(3):
        t0 = time.perf_counter()
        rip(Y3[idx[r]], maxdim=1)
        tr.append(time.perf_counter() - t0)
    out["C4"] = {"points": 1000, "rips_s": _med(tr), "resamples_per_s_serial": 1.0 / _med(tr), "per_resample": [round(t, 2) for t in tr]}
    print("C4", out["C4"], flush=True)

    # ---- per-stage figures (section 3, last bullet)
    from sklearn.metrics import pairwise_distances
    X = workloads.c3_layer(0, n=2000, d=4096, n_layers=32)
    pairwise_distances(X[:256], metric="cosine")
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        D = pairwise_distances(X, metric="cosine")
        ts.append(time.perf_counter() - t0)
    t_pd = min(ts)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        np.argsort(D, axis=1, kind="stable")[:, :15]
        ts.append(time.perf_counter() - t0)
    t_knn = min(ts)
    out["stages"] = {"pdist_cosine_2000x4096": {"seconds": t_pd, "gflops": 2 * 2000 * 2000 * 4096 / t_pd / 1e9, "threads": os.cpu_count()},
                     "knn_row_argsort_2000": {"seconds": t_knn, "gb_per_s": D.dtype.itemsize * 2000 * 2000 / t_knn / 1e9,
                                              "note": f"numpy stable argsort over the {D.dtype} matrix sklearn returns, one thread"}}
    print("stages", out["stages"], flush=True)

    # ---- C5 front end: pdist + kNN of a 4096-row block against all 100 000 rows, extrapolated to the 100 000 rows
    n5, blk = 100000, 4096
    t0 = time.perf_counter()
    X5 = workloads.c5_cloud(n=n5)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    Db = pairwise_distances(X5[:blk], X5, metric="cosine")
    t1 = time.perf_counter()
    np.argpartition(Db, 15, axis=1)[:, :15]
    t2 = time.perf_counter()
    out["C5"] = {"points": n5, "pdist_block_s": t1 - t0, "knn_block_s": t2 - t1, "block_rows": blk,
                 "pdist_knn_extrapolated_s": (t2 - t0) * n5 / blk, "workload_generation_s": t_gen,
                 "umap": "not run: the oracle's dense N x N path needs 40 GB for the matrix alone; umap-learn would switch to NN-descent (approximate) at this size",
                 "note": "kNN by argpartition (unordered top-15) per row block; threads = all"}
    del X5, Db
    print("C5 front end", out["C5"], flush=True)

    if not a.skip_big:
        out["C2"] = {"points": 2000, "maxdim": 2, **bounded(c2_job, (2000,), a.timeout)}
        print("C2", out["C2"], flush=True)
        out["C5"]["landmark_rips_10000"] = bounded(c5_landmark_job, (10000,), a.timeout)
        print("C5 landmarks", out["C5"]["landmark_rips_10000"], flush=True)

    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("written", a.out)


if __name__ == "__main__":
    main()
