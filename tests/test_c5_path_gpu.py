"""GPU: the large-single-cloud path of config C5 at a reduced size (the full 100k x 4096 needs minutes): row-sharded exact
kNN (never materialising the n x n matrix) -> UMAP from that kNN -> landmark Rips (ripser's n_perm semantics)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_row_sharded_knn_equals_full_matrix_knn():
    import torch
    from tda_multimodal_b200 import pipeline, umap_, workloads
    X = torch.from_numpy(workloads.c5_cloud(n=3001, d=256, seed=5000, latent_dim=6, n_mix=5)).cuda()
    D = umap_.distance_matrix(X[None], metric="cosine")
    want = umap_.knn_smooth(D, 15)
    # two "ranks", run one after the other on this GPU, each with small row blocks; concatenated == the full-matrix result
    parts = [pipeline.knn_row_sharded(X, 15, metric="cosine", rank=r, world=2, row_block=700) for r in range(2)]
    for q in range(4):
        got = torch.cat([parts[r][q][0] for r in range(2)], dim=0)
        if q == 0:
            assert torch.equal(got, want[q][0])                      # indices bit-exact (same GEMM accumulation order per element)
        else:
            assert torch.allclose(got, want[q][0], rtol=1e-6, atol=1e-7)
    # single-rank call returns the gathered layout directly
    one = pipeline.knn_row_sharded(X, 15, metric="cosine", row_block=1024)
    assert one[0].shape == (1, 3001, 15) and torch.equal(one[0], want[0])


def test_umap_from_sharded_knn_and_landmark_rips():
    import torch
    from sklearn.manifold import trustworthiness
    from oracle import rips as orips
    from tda_multimodal_b200 import pipeline, umap_, rips, workloads
    Xh = workloads.c5_cloud(n=5000, d=128, seed=5001, latent_dim=5, n_mix=4)
    X = torch.from_numpy(Xh).cuda()
    knn = pipeline.knn_row_sharded(X, 15, metric="cosine", row_block=2048)
    Y = umap_.umap_fit_batch(X[None], n_neighbors=15, n_components=3, metric="cosine", random_state=42, n_epochs=200, knn=knn)[0]
    Yh = Y.cpu().numpy()
    sub = np.random.default_rng(0).choice(5000, 1500, replace=False)
    assert trustworthiness(Xh[sub], Yh[sub], n_neighbors=10, metric="cosine") > 0.80
    # landmarks: greedy furthest-point permutation from index 0 (ripser's n_perm), Rips on the landmarks only
    r = rips.ripser(Yh, maxdim=1, n_perm=400)
    assert r["idx_perm"].shape == (400,) and r["idx_perm"][0] == 0 and len(set(r["idx_perm"].tolist())) == 400
    assert r["dperm2all"].shape == (400, 5000) and r["r_cover"] > 0
    want = orips.ripser(Yh[r["idx_perm"]], maxdim=1)["dgms"]
    assert np.array_equal(r["dgms"][0], want[0]) and np.array_equal(r["dgms"][1], want[1])
    # furthest-point property: every landmark was the farthest point from the previous ones
    d = np.linalg.norm(Yh[:, None, :] - Yh[r["idx_perm"][:50]][None, :, :], axis=2)
    for i in range(1, 50):
        assert np.argmax(d[:, :i].min(axis=1)) == r["idx_perm"][i]


def test_landmark_selection_at_c5_scale_matches_oracle():
    """Config C5's landmark step at n = 20 000 / 100 000 points (3-D, the UMAP output's shape): idx_perm of the one-launch cluster
    kernel equals the oracle's furthest-point sampling index for index, r_cover and every lambda bit for bit (ripser.py
    getGreedyPerm semantics: start at 0, lowest index on ties)."""
    import torch
    from oracle import rips as orips
    from tda_multimodal_b200 import rips
    rng = np.random.default_rng(77)
    for n, n_perm in ((20000, 2000), (100000, 1000)):
        Y = (rng.normal(0, 1, (n, 3)) * np.array([3.0, 2.0, 1.0]) + rng.integers(0, 4, (n, 1)) * 2.5).astype(np.float32)
        idx_t, lam_t = rips.greedy_permutation_points(torch.from_numpy(Y).cuda(), n_perm)
        want_idx, want_lam = orips.greedy_permutation_points(Y, n_perm)
        assert np.array_equal(idx_t.cpu().numpy(), want_idx)
        assert np.array_equal(lam_t.cpu().numpy(), want_lam.astype(np.float32))
    # through the ripser() shim: r_cover and the landmark diagrams (600 landmarks of 20 000 points) against the oracle on the same landmarks
    r = rips.ripser(Y[:20000], maxdim=1, n_perm=600)
    wi, wl = orips.greedy_permutation_points(Y[:20000], 600)
    assert np.array_equal(r["idx_perm"], wi) and r["r_cover"] == float(np.float32(wl[-1]))
    want = orips.ripser(Y[:20000][wi], maxdim=1)["dgms"]
    assert np.array_equal(r["dgms"][0], want[0]) and np.array_equal(r["dgms"][1], want[1])
