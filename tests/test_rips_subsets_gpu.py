"""GPU parity of the subset front end (tda_rips_sort_edges + tda_rips_subsets_launch: config C4, bootstrap resamples of one
cloud): the persistence of points[idx] taken from the parent's sorted edge list must be the SAME BITS -- diagrams, simplex
indices, thresholds, edge counts -- as the regular path on tda_pdist_lowdim(points[idx]), ties and duplicate points included."""
import numpy as np
import pytest

from tests.helpers import torus3d, blobs3d

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _subsets(rng, n_parent, m, B):
    return np.stack([np.sort(rng.choice(n_parent, size=m, replace=False)) for _ in range(B)]).astype(np.int32)


def _both_paths(torch, pts, idx, thresh=float("inf")):
    from tda_multimodal_b200 import rips
    P = torch.from_numpy(pts.astype(np.float32)).cuda()
    I = torch.from_numpy(idx).cuda()
    parent_dm = rips.pdist_lowdim(P[None])
    ends, sdist = rips.rips_sort_edges(parent_dm)
    sub = rips.rips_subsets_launch(ends, sdist, parent_dm, I, maxdim=1, thresh=thresh, want_simplices=True).finish()
    reg = rips.rips_batch(rips.pdist_lowdim(P[I.long()].contiguous()), maxdim=1, thresh=thresh, want_simplices=True)
    return sub, reg


def _assert_same(sub, reg):
    assert len(sub) == len(reg)
    for b, (s, r) in enumerate(zip(sub, reg)):
        assert s["num_edges"] == r["num_edges"], b
        assert np.float32(s["thresh"]) == np.float32(r["thresh"]), b
        for d in range(2):
            assert np.array_equal(s["dgms"][d], r["dgms"][d]), (b, d)
            assert np.array_equal(s["simplices"][d], r["simplices"][d]), (b, d)


def test_sorted_edges_are_ripsers_order(torch_cuda):
    torch = torch_cuda
    from tda_multimodal_b200 import rips
    rng = np.random.default_rng(70)
    pts = np.round(rng.normal(size=(2, 90, 3)) * 4) / 4          # many exactly equal lengths
    dm = rips.pdist_lowdim(torch.from_numpy(pts.astype(np.float32)).cuda())
    ends, sdist = rips.rips_sort_edges(dm)
    dmh = dm.cpu().numpy()
    for b in range(2):
        e = ends[b].cpu().numpy().astype(np.int64) & 0xffffffff
        i, j = e >> 16, e & 0xffff
        assert (i > j).all()
        d = sdist[b].cpu().numpy()
        assert np.array_equal(d, dmh[b][i, j])
        index = i * (i - 1) // 2 + j
        order = np.lexsort((-index, d))                          # length ascending, edge index descending
        assert np.array_equal(order, np.arange(len(d)))


@pytest.mark.parametrize("n_parent,m,B", [(300, 150, 5), (257, 256, 2), (64, 40, 3), (500, 333, 4)])
def test_subsets_equal_regular_path(torch_cuda, n_parent, m, B):
    rng = np.random.default_rng(71 + n_parent)
    pts = torus3d(n_parent, rng) if n_parent % 2 == 0 else blobs3d(n_parent, rng)
    _assert_same(*_both_paths(torch_cuda, pts, _subsets(rng, n_parent, m, B)))


def test_subsets_with_ties_and_duplicate_points(torch_cuda):
    rng = np.random.default_rng(72)
    g = np.stack(np.meshgrid(np.arange(6), np.arange(6), np.arange(5), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)   # 180 lattice points
    pts = np.concatenate([g, g[rng.choice(len(g), 40, replace=False)]])     # + 40 duplicates: zero-length edges
    pts = pts[rng.permutation(len(pts))]
    _assert_same(*_both_paths(torch_cuda, pts, _subsets(rng, len(pts), 120, 6)))


def test_subsets_with_user_threshold_and_tiny_sets(torch_cuda):
    rng = np.random.default_rng(73)
    pts = torus3d(200, rng)
    _assert_same(*_both_paths(torch_cuda, pts, _subsets(rng, 200, 80, 3), thresh=1.1))
    _assert_same(*_both_paths(torch_cuda, pts, _subsets(rng, 200, 200, 1)))            # the whole cloud
    sub, reg = _both_paths(torch_cuda, pts[:40], _subsets(rng, 40, 12, 2))              # few points: E_sub * 8 bytes still hold the vertex map
    _assert_same(sub, reg)
    # a finite threshold needs no parent distance matrix (it is only read for the enclosing radius)
    from tda_multimodal_b200 import rips
    torch = torch_cuda
    P = torch.from_numpy(pts.astype(np.float32)).cuda()
    idx = _subsets(rng, 200, 90, 2)
    ends, sdist = rips.rips_sort_edges(rips.pdist_lowdim(P[None]))
    got = rips.rips_subsets_launch(ends, sdist, None, torch.from_numpy(idx).cuda(), thresh=0.9, want_simplices=True).finish()
    want = rips.rips_batch(rips.pdist_lowdim(P[torch.from_numpy(idx).cuda().long()].contiguous()), thresh=0.9, want_simplices=True)
    _assert_same(got, want)
    with pytest.raises(ValueError):
        rips.rips_subsets_launch(ends, sdist, None, torch.from_numpy(idx).cuda())


def test_bootstrap_both_paths_and_replacement(torch_cuda):
    """pipeline.bootstrap_rips: subsets path (default) == per-resample path; sampling with replacement takes the per-resample path."""
    torch = torch_cuda
    from tda_multimodal_b200 import pipeline
    rng = np.random.default_rng(74)
    Y = torch.from_numpy(np.stack([torus3d(400, rng), blobs3d(400, rng)]).astype(np.float32)).cuda()
    a = pipeline.bootstrap_rips(Y, n_resamples=10, size=200, seed=4100, max_batch=4)
    b = pipeline.bootstrap_rips(Y, n_resamples=10, size=200, seed=4100, max_batch=4, subsets=False)
    for l in range(2):
        for r in range(10):
            for d in range(2):
                assert np.array_equal(a[l][r]["dgms"][d], b[l][r]["dgms"][d]), (l, r, d)
            assert a[l][r]["num_edges"] == b[l][r]["num_edges"]
    c = pipeline.bootstrap_rips(Y, n_resamples=3, size=150, seed=4100, replace=True)
    assert len(c[0]) == 3 and 100 < c[0][0]["dgms"][0].shape[0] <= 150    # (duplicate points: their zero-length H0 bars are dropped, as ripser does)
    with pytest.raises(ValueError):
        pipeline.bootstrap_rips(Y, n_resamples=3, size=150, seed=4100, replace=True, subsets=True)


@pytest.mark.parametrize("chunked", [1, 0])
def test_h0_from_edge_list_equals_matrix_boruvka(torch_cuda, chunked):
    """H0 by chunks of the sorted edge list (one launch) and by Boruvka rounds on the rank matrix: same rows, same death edges --
    on clustered clouds (the forest's long edges lie far beyond the first chunk), with a user threshold that leaves several
    components, on ties and duplicates -- and both equal to the oracle."""
    torch = torch_cuda
    from oracle import rips as orips
    from tda_multimodal_b200 import rips, _lib
    rng = np.random.default_rng(75)
    far = np.concatenate([rng.normal(size=(200, 3)) * 0.05 + c for c in ((0, 0, 0), (9, 0, 0), (0, 7, 0), (5, 5, 5), (-8, 2, 1))])   # 1000 points: the edges between the clusters start at rank ~99 500 (fourth chunk)
    lattice = np.stack(np.meshgrid(np.arange(7), np.arange(7), np.arange(6), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    lattice = np.concatenate([lattice, lattice[:30]])[rng.permutation(len(lattice) + 30)]
    _lib.set_option("rips_h0_chunked", chunked)
    try:
        for X, th in ((far, float("inf")), (far, 2.0), (lattice, float("inf")), (torus3d(700, rng), float("inf"))):
            X = X.astype(np.float32)
            want = orips.ripser(X, maxdim=1, thresh=th, with_simplices=True)
            got = rips.rips_batch(rips.pdist_lowdim(torch.from_numpy(X).cuda()[None]), maxdim=1, thresh=th, want_simplices=True)[0]
            assert np.array_equal(got["dgms"][0], want["dgms"][0])
            fin = np.isfinite(want["dgms"][0][:, 1])
            assert np.array_equal(got["simplices"][0][fin, 1], want["simplices"][0][fin, 1])
            assert np.array_equal(got["dgms"][1], want["dgms"][1])
    finally:
        _lib.set_option("rips_h0_chunked", 1)
