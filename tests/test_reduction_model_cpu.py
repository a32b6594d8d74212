"""CPU: the "propagate, then verify" model of the residual H1 reduction (oracle/rips_propagate_model.cpp, an algorithm study
for the next GPU reducer: apparent-pair additions as a forward substitution over a static two-parent graph, non-apparent
pivots found by verifying rows in order) must give the ripser restatement's H1 diagram -- same pairs, same order -- on the
reference's 32 shipped clouds and on random clouds."""
import numpy as np
import pytest

from oracle import rips as orips
from tests.helpers import blobs3d, circle2d, load_ref_rips_golden, torus3d


def test_model_equals_oracle_on_reference_clouds():
    clouds, _ = load_ref_rips_golden()
    for c in clouds:
        dm = orips.euclidean_dm_f32(c)
        want = orips.rips_dm(dm, maxdim=1)["dgms"][1]
        got, st = orips.model_h1(dm)
        assert np.array_equal(got, want)
        assert st["residual_columns"] >= len(want)


@pytest.mark.parametrize("gen,n,seed", [(torus3d, 150, 1), (blobs3d, 220, 2), (circle2d, 90, 3), (torus3d, 400, 4)])
def test_model_equals_oracle_on_random_clouds(gen, n, seed):
    X = gen(n, np.random.default_rng(seed))
    dm = orips.euclidean_dm_f32(X)
    want = orips.rips_dm(dm, maxdim=1, with_stats=True)
    got, st = orips.model_h1(dm)
    assert np.array_equal(got, want["dgms"][1])
    assert st["residual_columns"] == want["stats"][1]["reduced"]
    # the apparent-pair additions of the sequential algorithm became a substitution of depth << their number
    assert st["apparent_graph_depth"] < 64


@pytest.mark.parametrize("window", [1, 7, 64, 512, 1 << 30])
def test_windowed_model_equals_oracle(window):
    """The kernel-shaped variant: substitution and verification per window of ranks, flips above the first failing row undone.
    Any window size must give the same pairs (window 1 is the sequential algorithm, 2^30 one window per pass)."""
    clouds, _ = load_ref_rips_golden()
    for c in clouds[::4]:
        dm = orips.euclidean_dm_f32(c)
        got, _ = orips.model_h1(dm, window=window)
        assert np.array_equal(got, orips.rips_dm(dm, maxdim=1)["dgms"][1])
    for gen, n, seed in [(torus3d, 160, 5), (blobs3d, 250, 6), (torus3d, 420, 7)]:
        dm = orips.euclidean_dm_f32(gen(n, np.random.default_rng(seed)))
        want = orips.rips_dm(dm, maxdim=1, with_stats=True)
        got, st = orips.model_h1(dm, window=window)
        assert np.array_equal(got, want["dgms"][1]), (window, n)
        assert st["residual_columns"] == want["stats"][1]["reduced"]


@pytest.mark.parametrize("chunk", [4, 32, 512])
def test_kernel_step_model_equals_oracle(chunk):
    """The control flow of the GPU variant (Sweeper<WPL, VERIFY=true> in csrc/rips.cu), replayed on the CPU: heavy rows fixed by
    a filter per pass, substitution in dependency rounds, a whole-chunk re-filter when a flip touches a new vertex, undo of every
    recorded flip above the failing row."""
    clouds, _ = load_ref_rips_golden()
    for c in clouds[::5]:
        dm = orips.euclidean_dm_f32(c)
        got, _ = orips.model_h1(dm, window=chunk, kernel_steps=True)
        assert np.array_equal(got, orips.rips_dm(dm, maxdim=1)["dgms"][1])
    for gen, n, seed in [(torus3d, 170, 8), (blobs3d, 260, 9), (circle2d, 120, 10), (torus3d, 450, 11)]:
        dm = orips.euclidean_dm_f32(gen(n, np.random.default_rng(seed)))
        got, st = orips.model_h1(dm, window=chunk, kernel_steps=True)
        assert np.array_equal(got, orips.rips_dm(dm, maxdim=1)["dgms"][1]), (chunk, n, st)


def test_models_on_ties_and_duplicates():
    """Grid points (many equal edge lengths, duplicate points), circles and rounded Gaussians: all three model variants against
    the oracle (values compared as multisets where equal diameters leave the pair order open)."""
    rng = np.random.default_rng(123)
    for trial in range(80):
        n, dim = int(rng.integers(5, 70)), int(rng.integers(1, 4))
        kind = trial % 4
        if kind == 0:
            X = rng.normal(size=(n, dim)).astype(np.float32)
        elif kind == 1:
            X = rng.integers(0, 5, size=(n, dim)).astype(np.float32)
        elif kind == 2:
            t = rng.uniform(0, 2 * np.pi, n)
            X = np.c_[np.cos(t), np.sin(t)].astype(np.float32)
        else:
            X = np.round(rng.normal(size=(n, dim)) * 3).astype(np.float32) / 3
        dm = orips.euclidean_dm_f32(X)
        want = sorted(map(tuple, orips.rips_dm(dm, maxdim=1)["dgms"][1]))
        for kw in (dict(), dict(window=16), dict(window=8, kernel_steps=True), dict(window=512, kernel_steps=True)):
            got, _ = orips.model_h1(dm, **kw)
            assert sorted(map(tuple, got)) == want, (trial, kind, n, kw)


@pytest.mark.parametrize("cfg", [(1, 1, 1, 1, 8), (4, 16, 64, 2, 8), (32, 64, 64, 1, 1000000), (512, 8192, 65536, 64, 8)])
def test_two_mode_model_equals_oracle(cfg):
    """The control flow of the default GPU reducer (csrc/rips_sweep2.cuh): growing windows, substitution of every apparent row by
    rank in dependency rounds, sparse mode (exact lunes) / dense mode (superset mask Pend[c] & Pend[d], failing rows re-checked in
    ascending order with the exact lune, window end kept after an event).  cfg = (w0, wsparse, wmax, dense_min, dense_div)."""
    clouds, _ = load_ref_rips_golden()
    for c in clouds[::4]:
        dm = orips.euclidean_dm_f32(c)
        got, _ = orips.model_h1(dm, modes=cfg)
        assert np.array_equal(got, orips.rips_dm(dm, maxdim=1)["dgms"][1])
    for gen, n, seed in [(torus3d, 160, 5), (blobs3d, 250, 6), (circle2d, 120, 10), (torus3d, 420, 7)]:
        dm = orips.euclidean_dm_f32(gen(n, np.random.default_rng(seed)))
        got, st = orips.model_h1(dm, modes=cfg)
        assert np.array_equal(got, orips.rips_dm(dm, maxdim=1)["dgms"][1]), (cfg, n, st)
