"""CPU: the UMAP oracle (oracle/umap_oracle.py) against the stage invariants of SURVEY.md Appendix A.  The reference
pins nothing for these stages (inputs git-ignored, umap-learn unpinned) -- "parity unpinned" -- so the oracle is held
to the algorithm's defining properties instead."""
import numpy as np
import pytest

from oracle import umap_oracle as uo
from tda_multimodal_b200 import workloads


@pytest.fixture(scope="module")
def fitted():
    rng = np.random.default_rng(3)
    X = workloads._embed(workloads.torus_latent(300, rng, 0.05), 128, rng, noise=0.02, scale=7.0, offset=0.4)
    um = uo.UMAPOracle(n_neighbors=15, n_components=3, min_dist=0.1, metric="cosine", random_state=42)
    um.n_epochs = 60  # keep the CPU suite fast; the schedule logic is the same
    return X, um.fit(X)


def test_ab_params():
    a, b = uo.find_ab_params(1.0, 0.1)
    assert abs(a - 1.57694346) < 1e-4 and abs(b - 0.89506088) < 1e-4


def test_knn_and_smooth_invariants(fitted):
    X, um = fitted
    k = 15
    assert (um._knn_indices[:, 0] == np.arange(300)).all() and (um._knn_dists[:, 0] == 0).all()
    assert np.all(np.diff(um._knn_dists, axis=1) >= 0)
    d = um._knn_dists.astype(np.float64)
    assert np.array_equal(um._rhos, um._knn_dists[:, 1])  # smallest positive distance
    psum = np.exp(-np.maximum(d[:, 1:] - um._rhos[:, None], 0) / um._sigmas[:, None]).sum(1)
    assert np.abs(psum - np.log2(k)).max() < 1e-4


def test_graph_invariants(fitted):
    X, um = fitted
    G = um.graph_
    assert abs(G - G.T).max() < 1e-7 and G.data.min() > 0 and G.data.max() <= 1.0 + 1e-6
    assert np.allclose(G.max(axis=1).toarray().ravel(), 1.0, atol=1e-6)  # weight-1 edge to the nearest neighbour
    assert um._eps.min() >= 1.0 - 1e-9 and abs(um._eps[um._eps > 0].min() - 1.0) < 1e-9


def test_embedding_quality(fitted):
    from sklearn.manifold import trustworthiness
    X, um = fitted
    Y = um.embedding_
    assert Y.shape == (300, 3) and Y.dtype == np.float32 and np.isfinite(Y).all()
    assert trustworthiness(X, Y, n_neighbors=10, metric="cosine") > 0.85
    Yt = um.transform(X[:20] + 1e-4)
    assert Yt.shape == (20, 3) and np.isfinite(Yt).all()
    assert um.transform(X) is um.embedding_
