"""GPU parity: tcgen05 3xTF32 distance GEMM (through the C ABI) vs float64 numpy / sklearn."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# north_star tolerance: distances within 1e-5 relative (taken relative to the matrix scale, plus 1e-5 elementwise)
RTOL = 1e-5


def activations(n, d, rng, scale=30.0, offset=5.0):
    """LLM-like activations: low-rank structure + noise + a large common offset (hard case for ||x||^2+||y||^2-2xy)."""
    z = rng.normal(size=(n, 8))
    Q = np.linalg.qr(rng.normal(size=(d, 8)))[0]
    return (scale * (z @ Q.T) + rng.normal(0, 0.5, (n, d)) + offset).astype(np.float32)


def ref64(X, Y, metric):
    X = X.astype(np.float64)
    Y = X if Y is None else Y.astype(np.float64)
    if metric == "cosine":
        Xn = X / np.maximum(np.linalg.norm(X, axis=1, keepdims=True), 1e-300)
        Yn = Y / np.maximum(np.linalg.norm(Y, axis=1, keepdims=True), 1e-300)
        return np.clip(1.0 - Xn @ Yn.T, 0, 2)
    if metric == "dot":
        return X @ Y.T
    d2 = np.maximum((X * X).sum(1)[:, None] + (Y * Y).sum(1)[None, :] - 2 * X @ Y.T, 0)
    return d2 if metric == "sqeuclidean" else np.sqrt(d2)


@pytest.mark.parametrize("metric", ["cosine", "euclidean", "sqeuclidean", "dot"])
@pytest.mark.parametrize("B,n,d", [(1, 36, 4096), (3, 300, 4096), (2, 129, 100), (1, 2000, 4096), (4, 7, 33)])
def test_pdist_symmetric(metric, B, n, d):
    import torch
    from tda_multimodal_b200.pdist import pdist
    rng = np.random.default_rng(B * 1000 + n)
    X = np.stack([activations(n, d, rng) for _ in range(B)])
    D = pdist(torch.from_numpy(X).cuda(), metric=metric).cpu().numpy()
    assert D.shape == (B, n, n) and np.isfinite(D).all()
    for b in range(B):
        want = ref64(X[b], None, metric)
        if metric != "dot":
            np.fill_diagonal(want, 0.0)
            assert (np.diag(D[b]) == 0).all()
        scale = np.abs(want).max()
        err = np.abs(D[b] - want)
        assert (err <= RTOL * scale + RTOL * np.abs(want)).all(), (metric, err.max(), scale)
        assert np.array_equal(D[b], D[b].T) or np.abs(D[b] - D[b].T).max() <= RTOL * scale


def test_pdist_matches_sklearn_cosine_and_knn_order():
    import torch
    from sklearn.metrics import pairwise_distances
    from tda_multimodal_b200.pdist import pdist
    rng = np.random.default_rng(7)
    X = activations(500, 4096, rng, offset=0.0)
    D = pdist(torch.from_numpy(X).cuda(), metric="cosine").cpu().numpy()
    S = pairwise_distances(X, metric="cosine")
    assert np.abs(D - S).max() <= 1e-5
    # neighbour order of the 15 nearest agrees with the float64 oracle wherever the oracle's gaps exceed 2e-6
    want = ref64(X, None, "cosine")
    np.fill_diagonal(want, 0)
    oi = np.argsort(want, axis=1)[:, :16]
    od = np.take_along_axis(want, oi, 1)
    clear = np.diff(od, axis=1).min(axis=1) > 2e-6
    gi = np.argsort(D, axis=1, kind="stable")[:, :15]
    assert clear.mean() > 0.9 and np.array_equal(gi[clear], oi[clear][:, :15])


def test_pdist_asymmetric_and_disconnect():
    import torch
    from tda_multimodal_b200.pdist import pdist
    rng = np.random.default_rng(8)
    X, Y = activations(70, 512, rng, offset=0.0), activations(200, 512, rng, offset=0.0)
    D = pdist(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda(), metric="cosine").cpu().numpy()
    want = ref64(X, Y, "cosine")
    assert D.shape == (70, 200) and np.abs(D - want).max() <= 1e-5
    # antipodal rows are "disconnected" (umap-learn: cosine distance >= 2 -> inf)
    Z = np.concatenate([X[:4], -X[:4]])
    Dz = pdist(torch.from_numpy(Z).cuda(), metric="cosine", disconnect=2.0 - 1e-5).cpu().numpy()
    assert np.isinf(Dz[0, 4]) and np.isfinite(Dz[0, 1]) and Dz[0, 0] == 0
