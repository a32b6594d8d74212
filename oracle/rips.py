"""oracle/rips.py -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline); the product never imports it.

Python front end of the CPU Rips oracle.  Restates what ``ripser.ripser`` (ripser.py, unpinned third
party dependency of the reference, README.md:28) does around its C++ core for the arguments the
reference uses (debug_tda_pipeline.py:109-110, analyze_tda_over_layers.py:76,
analyze_adversarial_tda.py:100-101): euclidean ``pairwise_distances`` on the point cloud, cast to
float32, dense Rips persistence over Z/2 up to ``maxdim`` with the enclosing-radius threshold, and a
result dict whose ``dgms`` are float64 ``(n_k, 2)`` arrays.  SURVEY.md Appendix B is the spec.

Parity: PINNED by tests/test_oracle_golden.py (32 shipped clouds -> summary_stats.json).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "librips_oracle.so")
    src = os.path.join(_HERE, "rips_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "librips_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        lib.rips_oracle_run.restype = ctypes.c_void_p
        lib.rips_oracle_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float]
        lib.rips_oracle_run2.restype = ctypes.c_void_p
        lib.rips_oracle_run2.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int]
        lib.rips_oracle_count.restype = ctypes.c_int64
        lib.rips_oracle_count.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.rips_oracle_get.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.rips_oracle_num_edges.restype = ctypes.c_int64
        lib.rips_oracle_num_edges.argtypes = [ctypes.c_void_p]
        lib.rips_oracle_thresh.restype = ctypes.c_float
        lib.rips_oracle_thresh.argtypes = [ctypes.c_void_p]
        lib.rips_oracle_stats.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.rips_oracle_dep_stats.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.rips_oracle_free.argtypes = [ctypes.c_void_p]
        lib.rips_oracle_set_compact.argtypes = [ctypes.c_int64]
        lib.rips_oracle_set_window.argtypes = [ctypes.c_int64, ctypes.c_int64]
        _LIB = lib
    return _LIB


def euclidean_dm_f32(X):
    """Distance matrix exactly as the GPU path defines it: squared distance by exact differences in
    float64, rounded to float32, then a correctly rounded float32 sqrt.  ripser.py gets its matrix from
    sklearn ``pairwise_distances`` (float32 input -> float64 ||x||^2+||y||^2-2xy -> float32 -> float32
    sqrt); on the reference's 32 shipped clouds both definitions reproduce summary_stats.json bit for
    bit (tests/test_oracle_golden.py)."""
    X = np.asarray(X, dtype=np.float64)
    diff = X[:, None, :] - X[None, :, :]
    d2 = np.einsum("ijk,ijk->ij", diff, diff).astype(np.float32)
    return np.sqrt(d2)


def greedy_permutation(dm, n_perm):
    """Furthest-point sampling as ripser.py's ``n_perm`` does it: start at index 0."""
    n = dm.shape[0]
    idx = np.zeros(n_perm, dtype=np.int64)
    lambdas = np.zeros(n_perm)
    ds = dm[0].astype(np.float64).copy()
    for i in range(1, n_perm):
        j = int(np.argmax(ds))
        idx[i] = j
        lambdas[i - 1] = ds[j]
        ds = np.minimum(ds, dm[j])
    lambdas[-1] = ds.max()
    return idx, lambdas


def greedy_permutation_points(X, n_perm):
    """The same furthest-point sampling (ripser.py getGreedyPerm semantics: start at index 0, np.argmax = lowest index on ties)
    straight from the points, one distance row at a time -- for clouds whose n x n matrix is too large (config C5: 1e5 points).
    Distances: float64 accumulation of the squared differences, float64 sqrt, rounded to float32 (what greedy_perm_kernel
    computes for point input)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    idx = np.zeros(n_perm, dtype=np.int64)
    lambdas = np.zeros(n_perm)

    def row(j):
        diff = X - X[j]
        return np.sqrt(np.einsum("ij,ij->i", diff, diff)).astype(np.float32)
    ds = row(0)
    for i in range(1, n_perm):
        j = int(np.argmax(ds))
        idx[i] = j
        lambdas[i - 1] = ds[j]
        ds = np.minimum(ds, row(j))
    lambdas[-1] = ds.max()
    return idx, lambdas


def set_compact(min_entries=1 << 25):
    """Entries a working column of the oracle may hold before its cancelling pairs are removed (tests lower it)."""
    _lib().rips_oracle_set_compact(int(min_entries))


def set_window(min_ranks=64, div=256):
    """First diameter window of the lean mode's working coboundary: max(min_ranks, edges / div) edge ranks (tests use 1 rank)."""
    _lib().rips_oracle_set_window(int(min_ranks), int(div))


def rips_dm(dm, maxdim=1, thresh=np.inf, with_simplices=False, with_stats=False, apparent=False):
    """Persistence of the Rips filtration of a full float32 distance matrix.  ``apparent=True`` switches on Ripser 1.2's
    zero-apparent-pair shortcut (same rows in the same order; far less memory in the top dimension: rips_oracle.cpp header)."""
    dm = np.ascontiguousarray(dm, dtype=np.float32)
    n = dm.shape[0]
    lib = _lib()
    h = lib.rips_oracle_run2(dm.ctypes.data, n, int(maxdim), float(thresh), int(bool(apparent)))
    try:
        dgms, simp, stats = [], [], []
        for q in range(maxdim + 1):
            c = lib.rips_oracle_count(h, q)
            p = np.zeros((c, 2), dtype=np.float64)
            s = np.zeros((c, 2), dtype=np.int64)
            if c:
                lib.rips_oracle_get(h, q, p.ctypes.data, s.ctypes.data)
            dgms.append(p)
            simp.append(s)
            st = np.zeros(7, dtype=np.int64)
            lib.rips_oracle_stats(h, q, st.ctypes.data)
            rec = dict(zip(["columns", "emergent", "reduced", "additions", "pops", "max_v", "cofacets"], st.tolist()))
            dp = np.zeros(3, dtype=np.int64)
            lib.rips_oracle_dep_stats(h, q, dp.ctypes.data)
            rec.update(dict(zip(["dep_total_steps", "dep_critical_steps", "dep_depth"], dp.tolist())))
            stats.append(rec)
        out = {"dgms": dgms, "num_edges": int(lib.rips_oracle_num_edges(h)), "thresh": float(lib.rips_oracle_thresh(h))}
        if with_simplices:
            out["simplices"] = simp
        if with_stats:
            out["stats"] = stats
        return out
    finally:
        lib.rips_oracle_free(h)


def ripser(X, maxdim=1, thresh=np.inf, coeff=2, distance_matrix=False, do_cocycles=False,
           metric="euclidean", n_perm=None, _sklearn_dm=False, **extra):
    """Oracle twin of ``ripser.ripser`` (same keyword surface, same result keys)."""
    if coeff != 2:
        raise NotImplementedError("oracle supports coeff=2 only")
    X = np.asarray(X)
    if distance_matrix:
        if X.shape[0] != X.shape[1]:
            raise ValueError("Distance matrix is not square")
        dm = X.astype(np.float32)
    elif metric == "euclidean" and not _sklearn_dm:
        dm = euclidean_dm_f32(X)
    else:
        from sklearn.metrics import pairwise_distances
        dm = pairwise_distances(X, metric=metric).astype(np.float32)
    n = dm.shape[0]
    idx_perm = np.arange(n)
    r_cover = 0.0
    dperm2all = dm
    if n_perm is not None and n_perm < n:
        idx_perm, lambdas = greedy_permutation(dm, n_perm)
        r_cover = float(lambdas[-1])
        dperm2all = dm[idx_perm, :]
        dm = dperm2all[:, idx_perm]
    res = rips_dm(dm, maxdim=maxdim, thresh=thresh, **extra)
    res.update({"cocycles": [[] for _ in range(maxdim + 1)], "dperm2all": dperm2all,
                "idx_perm": idx_perm, "r_cover": r_cover})
    return res


_MODEL = None


def model_h1(dm, window=None, kernel_steps=False, pend=None, percol=False, modes=None):
    """H1 diagram of a float32 distance matrix by the "substitute, then verify" model (oracle/rips_propagate_model.cpp): a CPU
    study of the next GPU reducer, checked against `rips_dm` in tests/test_reduction_model_cpu.py.  window=None: the whole view
    above the cursor is propagated on every pass; window=k: the kernel-shaped variant (windows of k ranks: substitute, verify,
    undo above the first failing row).  Returns (pairs [k,2] float64 in processing order, stats dict)."""
    global _MODEL
    if _MODEL is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "librips_model.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.dirname(path)])
        lib = ctypes.CDLL(path)
        lib.rips_model_h1.restype = ctypes.c_int64
        lib.rips_model_h1.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        lib.rips_model_h1_windowed.restype = ctypes.c_int64
        lib.rips_model_h1_windowed.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        lib.rips_model_h1_kernel.restype = ctypes.c_int64
        lib.rips_model_h1_kernel.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        lib.rips_model_h1_pend.restype = ctypes.c_int64
        lib.rips_model_h1_pend.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int64]
        lib.rips_model_h1_modes.restype = ctypes.c_int64
        lib.rips_model_h1_modes.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _MODEL = lib
    dm = np.ascontiguousarray(dm, dtype=np.float32)
    n = dm.shape[0]
    cap = max(16, n * n // 4)
    out = np.zeros((cap, 2), dtype=np.float64)
    st = np.zeros(16, dtype=np.int64)
    if modes is not None:   # (w0, wsparse, wmax, dense_min, dense_div): the two-mode control flow of Sweeper2 (csrc/rips.cu)
        k = int(_MODEL.rips_model_h1_modes(dm.ctypes.data, n, *[int(v) for v in modes], out.ctypes.data, cap, st.ctypes.data))
        names = ["residual_columns", "apparent_edges", "events", "flips", "flips_undone", "heavy_rows_verified", "windows", "rounds",
                 "spurious_rows", "rows_substituted", "rows_in_late_rounds", "pm_rows_moved", "dense_columns", "exact_row_checks"]
        if k < 0:
            raise RuntimeError(f"model_h1: modes model failed ({k})")
        return out[:k].copy(), dict(zip(names, st[:len(names)].tolist()))
    if pend is not None:   # (w0, wmax): growing windows, substitution over all apparent rows, superset-mask verification (the cluster reducer)
        pc = np.zeros((n * n // 4 + 16, 6), dtype=np.int64) if percol else None
        k = int(_MODEL.rips_model_h1_pend(dm.ctypes.data, n, int(pend[0]), int(pend[1]), out.ctypes.data, cap, st.ctypes.data,
                                          pc.ctypes.data if percol else None, len(pc) if percol else 0))
        names = ["residual_columns", "apparent_edges", "events", "flips", "flips_undone", "heavy_rows_verified", "windows", "rounds",
                 "spurious_stops", "rows_substituted", "max_window", "pm_rows_moved"]
        if k == -2:
            raise RuntimeError("model_h1: substitution made no progress")
        if k < 0:
            raise RuntimeError("model_h1: pair buffer too small")
        stats = dict(zip(names, st[:len(names)].tolist()))
        if percol:
            stats["columns"] = pc[:stats["residual_columns"]].copy()
        return out[:k].copy(), stats
    if window is None:
        k = int(_MODEL.rips_model_h1(dm.ctypes.data, n, out.ctypes.data, cap, st.ctypes.data))
        names = ["residual_columns", "apparent_edges", "events", "propagated_flips", "heavy_rows_verified", "max_v", "passes", "apparent_graph_depth",
                 "incremental_flips", "heavy_rows_in_v"]
    elif kernel_steps:   # the control flow of Sweeper<WPL, VERIFY=true> (csrc/rips.cu), chunk = window
        k = int(_MODEL.rips_model_h1_kernel(dm.ctypes.data, n, int(window), out.ctypes.data, cap, st.ctypes.data))
        names = ["residual_columns", "apparent_edges", "events", "flips", "flips_undone", "heavy_rows_verified", "refilter_passes", "rounds"]
        if k == -2:
            raise RuntimeError("model_h1: substitution made no progress")
    else:
        k = int(_MODEL.rips_model_h1_windowed(dm.ctypes.data, n, int(window), out.ctypes.data, cap, st.ctypes.data))
        names = ["residual_columns", "apparent_edges", "events", "flips", "flips_undone", "heavy_rows_verified", "max_v", "windows"]
    if k < 0:
        raise RuntimeError("model_h1: pair buffer too small")
    return out[:k].copy(), dict(zip(names, st[:len(names)].tolist()))
