"""Shared helpers for the parity tests (synthetic clouds, diagram comparison)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_ref_rips_golden():
    clouds = np.load(os.path.join(GOLDEN, "ref_clouds_3d.npy"))
    with open(os.path.join(GOLDEN, "ref_summary_stats.json")) as f:
        stats = json.load(f)
    return clouds, stats


def get_persistence(dgm):
    """Verbatim semantics of the reference helper (debug_tda_pipeline.py:79-89)."""
    if dgm.shape[0] == 0:
        return np.array([]), 0.0
    pers = dgm[:, 1] - dgm[:, 0]
    pers = pers[np.isfinite(pers)]
    if pers.shape[0] == 0:
        return np.array([]), 0.0
    return pers, np.max(pers)


def reference_stats(dgms):
    """The H0/H1 fields of one summary_stats.json record (debug_tda_pipeline.py:121-131)."""
    h0, m0 = get_persistence(dgms[0])
    h1, m1 = get_persistence(dgms[1])
    return {"n_h1_features": len(h1), "max_h1_persistence": float(m1), "all_h1_persistence_values": h1.tolist(),
            "n_h0_features": len(dgms[0]) - len(h0), "max_h0_persistence": float(m0)}


def torus3d(n, rng, noise=0.05):
    t, p = rng.uniform(0, 2 * np.pi, (2, n))
    X = np.c_[(3 + np.cos(p)) * np.cos(t), (3 + np.cos(p)) * np.sin(t), np.sin(p)]
    return (X + rng.normal(0, noise, X.shape)).astype(np.float32)


def blobs3d(n, rng, k=8):
    c = rng.normal(0, 5, (k, 3))
    return (c[rng.integers(0, k, n)] + rng.normal(0, 1, (n, 3))).astype(np.float32)


def circle2d(n, rng, noise=0.02):
    t = rng.uniform(0, 2 * np.pi, n)
    return (np.c_[np.cos(t), np.sin(t)] + rng.normal(0, noise, (n, 2))).astype(np.float32)


def sorted_rows(d):
    d = np.asarray(d, dtype=np.float64).reshape(-1, 2)
    return d[np.lexsort((d[:, 1], d[:, 0]))]


def same_diagram(a, b):
    """Equal as multisets of (birth, death) rows, bit for bit."""
    a, b = sorted_rows(a), sorted_rows(b)
    return a.shape == b.shape and np.array_equal(a, b)
