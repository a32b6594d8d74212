"""GPU parity: libtda_b200 Rips (through the C ABI) vs the reference's shipped results and the CPU oracle."""
import numpy as np
import pytest

from tests.helpers import load_ref_rips_golden, reference_stats, torus3d, blobs3d, circle2d, same_diagram

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def test_reference_golden_clouds_batched(torch_cuda):
    """All 32 shipped clouds in ONE batched call; every summary_stats.json field bit-exact."""
    torch = torch_cuda
    from tda_multimodal_b200 import rips
    clouds, stats = load_ref_rips_golden()
    dm = rips.pdist_lowdim(torch.from_numpy(clouds).cuda())
    res = rips.rips_batch(dm, maxdim=1)
    for i in range(32):
        got = reference_stats(res[i]["dgms"])
        for key in ("n_h1_features", "max_h1_persistence", "all_h1_persistence_values", "n_h0_features", "max_h0_persistence"):
            assert got[key] == stats[i][key], (i, key, got[key], stats[i][key])


def test_reference_golden_via_shim(torch_cuda):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shims"))
    from ripser import ripser
    clouds, stats = load_ref_rips_golden()
    for i in (0, 13, 25, 31):
        r = ripser(clouds[i], maxdim=1)
        assert set(r) >= {"dgms", "cocycles", "num_edges", "dperm2all", "idx_perm", "r_cover"}
        assert r["dgms"][0].dtype == np.float64 and r["dgms"][1].dtype == np.float64
        got = reference_stats(r["dgms"])
        assert got["all_h1_persistence_values"] == stats[i]["all_h1_persistence_values"]
        assert got["max_h0_persistence"] == stats[i]["max_h0_persistence"]


@pytest.mark.parametrize("gen,n,seed", [(torus3d, 100, 0), (torus3d, 400, 1), (blobs3d, 300, 2), (blobs3d, 1000, 3),
                                        (circle2d, 257, 4), (torus3d, 1000, 5)])
def test_matches_oracle_bit_exact(torch_cuda, gen, n, seed):
    torch = torch_cuda
    from tda_multimodal_b200 import rips
    from oracle import rips as orips
    X = gen(n, np.random.default_rng(seed))
    want = orips.ripser(X, maxdim=1, with_simplices=True)
    dm = rips.pdist_lowdim(torch.from_numpy(X).cuda()[None])
    assert np.array_equal(dm[0].cpu().numpy(), orips.euclidean_dm_f32(X))  # distance matrix bit-exact
    got = rips.rips_batch(dm, maxdim=1, want_simplices=True, want_stats=True)[0]
    assert got["num_edges"] == want["num_edges"]
    assert np.float32(got["thresh"]) == np.float32(want["thresh"])
    # H0: same rows in the same order, same death-edge indices (tie-free data)
    assert np.array_equal(got["dgms"][0], want["dgms"][0])
    assert np.array_equal(got["simplices"][0][:-1, 1], want["simplices"][0][:-1, 1])
    # H1: same multiset of (birth, death), births descending, same birth edges
    assert same_diagram(got["dgms"][1], want["dgms"][1])
    assert np.all(np.diff(got["dgms"][1][:, 0]) <= 0)
    assert np.array_equal(got["dgms"][1], want["dgms"][1])
    assert np.array_equal(got["simplices"][1][:, 0], want["simplices"][1][:, 0])


def test_edge_cases(torch_cuda):
    torch = torch_cuda
    from tda_multimodal_b200 import rips
    from oracle import rips as orips
    sq = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float32)
    d = rips.ripser(sq, maxdim=1)["dgms"]
    assert d[1].shape == (1, 2) and d[1][0, 0] == 1.0 and d[1][0, 1] == np.float32(np.sqrt(np.float32(2.0)))
    # single point, two points
    one = rips.ripser(np.zeros((1, 3), np.float32), maxdim=1)["dgms"]
    assert one[0].shape == (1, 2) and np.isinf(one[0][0, 1]) and one[1].shape == (0, 2)
    two = rips.ripser(np.array([[0, 0, 0], [3, 4, 0]], np.float32), maxdim=1)["dgms"]
    assert np.array_equal(two[0], np.array([[0, 5.0], [0, np.inf]])) and two[1].shape == (0, 2)
    # duplicate points (zero-length edges) and ties: diagrams equal as multisets
    rng = np.random.default_rng(9)
    X = blobs3d(120, rng)
    X = np.concatenate([X, X[:20]])  # bootstrap-with-replacement style duplicates
    want = orips.ripser(X, maxdim=1)["dgms"]
    got = rips.ripser(X, maxdim=1)["dgms"]
    assert same_diagram(got[0], want[0]) and same_diagram(got[1], want[1])
    # integer grid: massive ties
    g = np.stack(np.meshgrid(np.arange(6), np.arange(6)), -1).reshape(-1, 2).astype(np.float32)
    want = orips.ripser(g, maxdim=1)["dgms"]
    got = rips.ripser(g, maxdim=1)["dgms"]
    assert same_diagram(got[0], want[0]) and same_diagram(got[1], want[1])
    # explicit threshold below the enclosing radius: several essential H0 classes
    X = blobs3d(200, rng)
    want = orips.ripser(X, maxdim=1, thresh=1.5)["dgms"]
    got = rips.ripser(X, maxdim=1, thresh=1.5)["dgms"]
    assert same_diagram(got[0], want[0]) and same_diagram(got[1], want[1])
    # distance-matrix input, non-square rejected
    dm = orips.euclidean_dm_f32(X)
    got = rips.ripser(dm, maxdim=1, distance_matrix=True)["dgms"]
    want = orips.ripser(dm, maxdim=1, distance_matrix=True)["dgms"]
    assert same_diagram(got[1], want[1])
    with pytest.raises(ValueError):
        rips.ripser(np.zeros((3, 4), np.float32), distance_matrix=True)
    with pytest.raises(NotImplementedError):
        rips.ripser(X, coeff=3)


def test_n_perm_landmarks(torch_cuda):
    from tda_multimodal_b200 import rips
    from oracle import rips as orips
    X = torus3d(600, np.random.default_rng(21))
    got = rips.ripser(X, maxdim=1, n_perm=150)
    want = orips.ripser(X, maxdim=1, n_perm=150)
    assert np.array_equal(got["idx_perm"], want["idx_perm"])
    assert np.float32(got["r_cover"]) == np.float32(want["r_cover"])
    assert same_diagram(got["dgms"][1], want["dgms"][1]) and same_diagram(got["dgms"][0], want["dgms"][0])


def test_large_cloud_properties(torch_cuda):
    """BASELINE-size cloud (n=2000): size-independent properties instead of the (slow) oracle."""
    torch = torch_cuda
    from tda_multimodal_b200 import rips
    rng = np.random.default_rng(2000)
    X = blobs3d(2000, rng)
    dm = rips.pdist_lowdim(torch.from_numpy(X).cuda()[None])
    r = rips.rips_batch(dm, maxdim=1)[0]
    d0, d1 = r["dgms"]
    assert d0.shape == (2000, 2) and np.isinf(d0[-1, 1]) and np.all(np.diff(d0[:-1, 1]) >= 0)
    from scipy.sparse.csgraph import minimum_spanning_tree
    mst = np.sort(minimum_spanning_tree(dm[0].cpu().numpy().astype(np.float64)).data)
    assert np.array_equal(mst.astype(np.float32), d0[:-1, 1].astype(np.float32))
    assert np.all(d1[:, 1] > d1[:, 0]) and np.all(np.diff(d1[:, 0]) <= 0)
    perm = rng.permutation(2000)
    r2 = rips.rips_batch(rips.pdist_lowdim(torch.from_numpy(X[perm]).cuda()[None]), maxdim=1)[0]
    assert same_diagram(r2["dgms"][1], d1) and same_diagram(r2["dgms"][0], d0)


def test_bootstrap_resamples_match_oracle(torch_cuda):
    """Config C4 semantics: resamples of one cloud, batched on the GPU == independent oracle calls on the same index sets."""
    torch = torch_cuda
    from oracle import rips as orips
    from tda_multimodal_b200 import pipeline, workloads
    rng = np.random.default_rng(21)
    Y = np.stack([torus3d(300, rng), blobs3d(300, rng)])
    res = pipeline.bootstrap_rips(torch.from_numpy(Y).cuda(), n_resamples=6, size=150, seed=4000, max_batch=4)
    assert len(res) == 2 and len(res[0]) == 6
    for l in range(2):
        idx = workloads.c4_resample_indices(l, 300, 6, 150, seed=4000)
        for r in range(6):
            want = orips.ripser(Y[l][idx[r]], maxdim=1)["dgms"]
            got = res[l][r]["dgms"]
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_async_jobs_and_chunked_sweep_equal_blocking(torch_cuda):
    torch = torch_cuda
    from tda_multimodal_b200 import rips
    rng = np.random.default_rng(22)
    X = np.stack([torus3d(200, rng) for _ in range(5)])
    dm = rips.pdist_lowdim(torch.from_numpy(X).cuda())
    blocking = rips.rips_batch(dm, maxdim=1)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        j1 = rips.rips_batch_launch(dm[:2], maxdim=1)
    with torch.cuda.stream(s2):
        j2 = rips.rips_batch_launch(dm[2:], maxdim=1)
    got = j1.finish() + j2.finish()
    for a, b in zip(got, blocking):
        assert np.array_equal(a["dgms"][0], b["dgms"][0]) and np.array_equal(a["dgms"][1], b["dgms"][1])
    # too small a capacity is detected by finish() and repaired by a synchronous re-run
    j3 = rips.rips_batch_launch(dm, maxdim=1, cap1=2)
    got3 = j3.finish()
    assert all(np.array_equal(a["dgms"][1], b["dgms"][1]) for a, b in zip(got3, blocking))


def test_full_size_invariants(torch_cuda):
    """BASELINE.json sizes (2000-point 3-D clouds): properties that do not need the (slow) oracle -- #H0 rows = n with one
    infinite bar, H0 deaths ascending = MST weights, H1 births descending, invariance under a permutation of the points and
    exact equivariance under scaling by a power of two."""
    torch = torch_cuda
    from scipy.sparse.csgraph import minimum_spanning_tree
    from tda_multimodal_b200 import rips
    rng = np.random.default_rng(33)
    X = torus3d(2000, rng)
    perm = rng.permutation(2000)
    batch = np.stack([X, X[perm], X * np.float32(4.0)])
    dm = rips.pdist_lowdim(torch.from_numpy(batch).cuda())
    a, b, c = rips.rips_batch(dm, maxdim=1)
    d0, d1 = a["dgms"]
    assert d0.shape == (2000, 2) and np.isinf(d0[-1, 1]) and np.isfinite(d0[:-1, 1]).all() and np.all(np.diff(d0[:-1, 1]) >= 0)
    mst = np.sort(minimum_spanning_tree(dm[0].double().cpu().numpy()).data)
    assert np.array_equal(mst.astype(np.float32), d0[:-1, 1].astype(np.float32))
    assert np.all(np.diff(d1[:, 0]) <= 0) and np.all(d1[:, 1] > d1[:, 0])
    assert same_diagram(b["dgms"][0], d0) and same_diagram(b["dgms"][1], d1)
    assert np.array_equal(c["dgms"][0][:-1] / 4.0, d0[:-1]) and same_diagram(c["dgms"][1] / 4.0, d1)
    pers = np.sort(d1[:, 1] - d1[:, 0])[::-1]
    assert pers[0] > 2 * pers[2]  # torus R=3, r=1: two dominant classes at most


def test_c4_full_size_resamples_spot_checked_against_oracle(torch_cuda):
    """The north-star's bootstrap leg at ITS size (config C4: resamples of 1000 of a 2000-point 3-D cloud, 256 per layer, batched 256
    problems per call): eight (layer, resample) units picked at random out of 2 x 256 are recomputed by the oracle on the same index
    sets and must agree row for row (H0 and H1, ripser's order)."""
    torch = torch_cuda
    from oracle import rips as orips
    from tda_multimodal_b200 import pipeline, workloads
    rng = np.random.default_rng(44)
    Y = np.stack([torus3d(2000, rng), blobs3d(2000, rng)])
    res = pipeline.bootstrap_rips(torch.from_numpy(Y).cuda(), n_resamples=256, size=1000, seed=4000, layer_ids=[3, 17])
    assert len(res) == 2 and len(res[0]) == 256
    pick = np.random.default_rng(45)
    for _ in range(8):
        l, r = int(pick.integers(0, 2)), int(pick.integers(0, 256))
        idx = workloads.c4_resample_indices([3, 17][l], 2000, 256, 1000, seed=4000)[r]
        want = orips.ripser(Y[l][idx], maxdim=1)["dgms"]
        got = res[l][r]["dgms"]
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (l, r)
