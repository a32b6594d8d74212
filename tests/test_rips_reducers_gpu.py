"""GPU: every residual-H1 reducer of the library (tda_set_option("rips_reducer", ...)) must give the oracle's diagrams and
simplex pairs bit for bit: 0 "sweep2" (default: substitution by rank + window verification, csrc/rips_sweep2.cuh; CPU model
oracle/rips_propagate_model.cpp rips_model_h1_modes), 1 the row sweep with a sequential resolver, 2 its substitute-then-verify
variant, 3 the key bitset (the only one for n > 8192 besides sweep2).  Includes the bench size (n = 2000: a real C3 embedding, the
WPL=2 instantiation of reducers 1/2 and both the sparse and the dense mode of sweep2) and small windows that force every
control path of sweep2 (events inside windows, spurious stops, window regrowth, mode switches)."""
import numpy as np
import pytest

from tests.helpers import blobs3d, circle2d, load_ref_rips_golden, torus3d

pytestmark = pytest.mark.gpu
REDUCERS = {"sweep2": 0, "sweep": 1, "verify": 2, "bitset": 3}


def _check(X, reducer, tda_option, **opts):
    import torch
    from oracle import rips as orips
    from tda_multimodal_b200 import rips
    tda_option("rips_reducer", REDUCERS[reducer])
    for k, v in opts.items():
        tda_option(k, v)
    want = orips.ripser(X, maxdim=1, with_simplices=True)
    dm = rips.pdist_lowdim(torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda()[None])
    got = rips.rips_batch(dm, maxdim=1, want_simplices=True, want_stats=True)[0]
    assert np.array_equal(got["dgms"][0], want["dgms"][0])
    assert np.array_equal(got["dgms"][1], want["dgms"][1]), (reducer, opts, len(got["dgms"][1]), len(want["dgms"][1]))
    assert np.array_equal(got["simplices"][1][:, 0], want["simplices"][1][:, 0])   # birth edge of every H1 pair (the death triangle
    # of a pair is one of several of equal diameter: the library's reducers agree on it among themselves, see the last test)
    torch.cuda.synchronize()
    return got


@pytest.mark.parametrize("reducer", ["sweep2", "sweep", "verify", "bitset"])
@pytest.mark.parametrize("gen,n,seed", [(torus3d, 60, 0), (circle2d, 150, 1), (blobs3d, 400, 2), (torus3d, 1000, 3)])
def test_reducer_matches_oracle(tda_option, reducer, gen, n, seed):
    _check(gen(n, np.random.default_rng(seed)), reducer, tda_option)


@pytest.mark.parametrize("opts", [
    dict(rips_w0=32, rips_wsparse=32, rips_wmax=64, rips_dense_min=1, rips_dense_div=1000000),      # tiny windows, everything dense
    dict(rips_w0=32, rips_wsparse=64, rips_wmax=64, rips_dense_min=1000000, rips_dense_div=1),      # tiny windows, never dense
    dict(rips_w0=64, rips_wsparse=256, rips_wmax=4096, rips_dense_min=4, rips_dense_div=64),        # early mode switches
    dict(rips_w0=4096, rips_wsparse=65472, rips_wmax=65472, rips_dense_min=16, rips_dense_div=16),  # largest windows
])
@pytest.mark.parametrize("gen,n,seed", [(torus3d, 420, 7), (blobs3d, 700, 8), (circle2d, 300, 9)])
def test_sweep2_window_schedules(tda_option, opts, gen, n, seed):
    """the cluster engine alone (rips_warp_engine=0: every column goes through the windows)"""
    got = _check(gen(n, np.random.default_rng(seed)), "sweep2", tda_option, rips_warp_engine=0, **opts)
    st = got["stats"]
    assert st["windows"] > 0 and st["rows_substituted"] > 0 and st["columns_to_cluster"] == st["reduced"]


@pytest.mark.parametrize("gen,n,seed", [(torus3d, 420, 7), (blobs3d, 700, 8), (circle2d, 300, 9), (torus3d, 1500, 10)])
def test_sweep2_warp_engine(tda_option, gen, n, seed):
    """the default: every column first by a single warp (speculative, no owner look-ups), committed in ripser's order by one warp that
    resumes the columns whose tentative pivot is owned, and only the columns that outgrow a warp go through the windows"""
    got = _check(gen(n, np.random.default_rng(seed)), "sweep2", tda_option, rips_warp_engine=1, rips_w0=256)
    st = got["stats"]
    assert st["columns_to_cluster"] < st["reduced"]


def test_sweep2_dense_mode_is_exercised(tda_option):
    got = _check(torus3d(900, np.random.default_rng(21)), "sweep2", tda_option, rips_warp_engine=0, rips_w0=256, rips_wsparse=1024, rips_wmax=8192,
                 rips_dense_min=8, rips_dense_div=32)
    assert got["stats"]["dense_columns"] > 0 and got["stats"]["pm_rows_moved"] > 0


@pytest.mark.parametrize("reducer", ["sweep2", "sweep", "verify"])
def test_reducers_on_reference_clouds_batched(tda_option, reducer):
    import torch
    from oracle import rips as orips
    from tda_multimodal_b200 import rips
    tda_option("rips_reducer", REDUCERS[reducer])
    clouds, _ = load_ref_rips_golden()
    dm = rips.pdist_lowdim(torch.from_numpy(clouds.astype(np.float32)).cuda())
    res = rips.rips_batch(dm, maxdim=1)
    for i in range(len(clouds)):
        want = orips.ripser(clouds[i], maxdim=1)["dgms"]
        assert np.array_equal(res[i]["dgms"][0], want[0]) and np.array_equal(res[i]["dgms"][1], want[1])


@pytest.mark.parametrize("reducer", ["sweep2", "sweep", "verify"])
def test_bench_size_c3_embedding_matches_oracle(tda_option, reducer):
    """n = 2000, the size bench.py runs: a real C3 cloud (layer 0 of the synthetic sweep embedded by the GPU UMAP path) against the
    oracle, diagrams and simplex pairs bit for bit (the oracle needs ~10-40 s for such a cloud)."""
    import torch
    from tda_multimodal_b200 import umap_, workloads
    X = torch.from_numpy(workloads.c3_layers(layers=[0])).cuda()
    Y = umap_.umap_fit_batch(X, n_neighbors=15, n_components=3, metric="cosine", random_state=42)[0].cpu().numpy()
    got = _check(Y, reducer, tda_option)
    if reducer == "sweep2":
        assert got["stats"]["dense_columns"] > 0   # the long columns of a 2000-point cloud run in dense mode (cluster engine)
        assert 0 < got["stats"]["columns_to_cluster"] < got["stats"]["reduced"] // 2   # most columns never leave their warp


def test_sweep2_batch_of_bench_size_clouds_equals_sweep(tda_option):
    """8 blobs/torus clouds of 2000 points in one batched call: sweep2 and the sequential row sweep agree pair for pair."""
    import torch
    from tda_multimodal_b200 import rips
    rng = np.random.default_rng(5)
    X = np.stack([(torus3d if i % 2 else blobs3d)(2000, rng) for i in range(8)]).astype(np.float32)
    dm = rips.pdist_lowdim(torch.from_numpy(X).cuda())
    out = {}
    for reducer in ("sweep2", "sweep"):
        tda_option("rips_reducer", REDUCERS[reducer])
        out[reducer] = rips.rips_batch(dm, maxdim=1, want_simplices=True)
    for a, b in zip(out["sweep2"], out["sweep"]):
        assert np.array_equal(a["dgms"][1], b["dgms"][1]) and np.array_equal(a["simplices"][1], b["simplices"][1])
        assert np.array_equal(a["dgms"][0], b["dgms"][0])


@pytest.mark.parametrize("cluster", [1, 2, 8])
@pytest.mark.parametrize("gen,n,seed", [(torus3d, 500, 31), (blobs3d, 1200, 32)])
def test_sweep2_cluster_sizes(tda_option, cluster, gen, n, seed):
    """sweep2 deals every pass over a window to the warps of a thread-block cluster (default 4 CTAs per cloud); any cluster size
    must give the oracle's pairs."""
    for we in (0, 1):
        got = _check(gen(n, np.random.default_rng(seed)), "sweep2", tda_option, rips_cluster=cluster, rips_w0=256, rips_dense_min=16, rips_warp_engine=we)
    assert got["stats"]["reduced"] > 0
