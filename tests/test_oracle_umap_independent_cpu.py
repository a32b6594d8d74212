"""CPU: stages of the UMAP oracle (oracle/umap_oracle.py) against INDEPENDENT formulations of the published algorithm
(McInnes, Healy, Melville 2018, sections 3.1-3.2 and Algorithms 2-5).  The reference pins no UMAP
inputs or outputs ("parity unpinned"), so the oracle cannot be checked against umap-learn's numbers; what can
be checked is that each stage computes what the paper defines, by a second route that shares no code with the oracle:
  * sigma_i as the root of sum_j exp(-(d_ij - rho_i) / sigma_i) = log2(k), found by scipy's brentq instead of the bisection;
  * the fuzzy union as the dense matrix expression P + P^T - P o P^T instead of sparse COO arithmetic;
  * the spectral initialisation as the eigenvectors of the dense normalised Laplacian (LAPACK eigh) instead of ARPACK;
  * the sampling schedule as the closed form n_epochs * w / w_max firings per edge;
  * the SGD update as the numerical derivative of the cross-entropy terms log(phi) and log(1 - phi), phi = 1 / (1 + a d^2b),
    instead of the closed-form coefficients, on a system small enough to follow by hand."""
import numpy as np
import pytest

from oracle import umap_oracle as uo
from tda_multimodal_b200 import workloads


@pytest.fixture(scope="module")
def fitted():
    rng = np.random.default_rng(5)
    X = workloads._embed(workloads.torus_latent(240, rng, 0.05), 96, rng, noise=0.02, scale=5.0, offset=0.3)
    um = uo.UMAPOracle(n_neighbors=12, n_components=3, min_dist=0.1, metric="cosine", random_state=42)
    um.n_epochs = 30
    return X, um.fit(X)


def test_sigma_is_the_root_of_the_smoothing_equation(fitted):
    from scipy.optimize import brentq
    _, um = fitted
    k = um._n_neighbors
    d = um._knn_dists.astype(np.float64)
    for i in range(0, d.shape[0], 7):
        rho = d[i, 1:][d[i, 1:] > 0].min()
        assert um._rhos[i] == np.float32(rho)

        def f(s):
            return np.exp(-np.maximum(d[i, 1:] - rho, 0.0) / s).sum() - np.log2(k)
        root = brentq(f, 1e-8, 1e3, xtol=1e-14)
        assert abs(um._sigmas[i] - root) <= 2e-4 * root, (i, um._sigmas[i], root)   # the bisection stops at |sum - log2 k| < 1e-5


def test_fuzzy_union_equals_the_dense_matrix_expression(fitted):
    _, um = fitted
    n, k = um._knn_indices.shape
    P = np.zeros((n, n))
    d = um._knn_dists.astype(np.float64)
    for i in range(n):
        for jj in range(k):
            j = um._knn_indices[i, jj]
            if j == i:
                continue
            x = d[i, jj] - um._rhos[i]
            P[i, j] = 1.0 if (x <= 0 or um._sigmas[i] == 0) else np.exp(-x / um._sigmas[i])
    W = P + P.T - P * P.T
    G = um.graph_.toarray()
    assert np.array_equal(G != 0, W != 0)
    assert np.abs(G - W).max() < 3e-6


def test_sampling_schedule_is_the_closed_form(fitted):
    _, um = fitted
    w = um.graph_.tocoo()
    w.sum_duplicates()
    data = w.data.copy()
    n_epochs = 30
    data[data < data.max() / n_epochs] = 0
    data = data[data > 0]
    firings = n_epochs * data / data.max()                 # the paper: an edge of weight w is sampled with probability ~ w
    assert um._eps.shape == data.shape
    assert np.allclose(n_epochs / um._eps, firings, rtol=1e-12) and um._eps.min() == 1.0


def test_spectral_init_spans_the_bottom_eigenvectors_of_the_dense_laplacian(fitted):
    import scipy.sparse.csgraph
    X, um = fitted
    g = um.graph_.tocoo()
    g.sum_duplicates()
    g.data[g.data < g.data.max() / 30.0] = 0.0
    g.eliminate_zeros()
    ncomp, _ = scipy.sparse.csgraph.connected_components(g)
    assert ncomp == 1
    A = g.toarray().astype(np.float64)
    deg = A.sum(0)
    L = np.eye(len(A)) - A / np.sqrt(deg)[:, None] / np.sqrt(deg)[None, :]
    vals, vecs = np.linalg.eigh(L)
    assert abs(vals[0]) < 1e-10 and vals[4] - vals[3] > 1e-6      # eigenvalues 1..3 are separated from the rest
    Q = vecs[:, 1:4]
    init = uo.spectral_layout(X, g, 3, np.random.RandomState(42), metric="cosine").astype(np.float64)
    init /= np.linalg.norm(init, axis=0)
    resid = init - Q @ (Q.T @ init)                                 # what of the oracle's vectors lies outside the dense eigenspace
    assert np.abs(resid).max() < 5e-3, np.abs(resid).max()          # ARPACK runs with tol = 1e-4 in umap-learn


def _phi(d2, a, b):
    return 1.0 / (1.0 + a * d2 ** b)


def _num_grad(f, y, h=1e-6):
    g = np.zeros_like(y)
    for t in range(len(y)):
        e = np.zeros_like(y)
        e[t] = h
        g[t] = (f(y + e) - f(y - e)) / (2 * h)
    return g


def test_sgd_steps_are_gradient_steps_of_the_cross_entropy():
    """One head point, one edge to a fixed tail, every negative sample lands on the same fixed point (the tail array holds one
    position for every vertex, so the random index does not matter).  With epochs_per_sample = 1 the edge first fires in epoch 1
    (`epoch_of_next_sample <= n`): one attractive step, then the (1 - 1/5) / (1/5) = 4 negative samples that came due.  Each step
    must be alpha times the (clipped) gradient of log(phi) resp. gamma * log(1 - phi) at the current position, taken numerically
    here."""
    a, b = uo.find_ab_params(1.0, 0.1)
    alpha0, gamma = 0.05, 1.0
    y0 = np.array([[0.7, -0.4, 0.3]], dtype=np.float32)
    p = np.array([1.9, 0.8, -0.6], dtype=np.float32)
    n_vertices = 4
    tail = np.tile(p, (n_vertices, 1)).astype(np.float32)
    eps = np.array([1.0])
    out = uo.optimize_layout_euclidean(y0.copy(), tail.copy(), np.array([0], dtype=np.int64), np.array([2], dtype=np.int64), 2, n_vertices,
                                       eps, a, b, np.array([11, 22, 33], dtype=np.int64), gamma, alpha0, 5.0, False)
    y = y0[0].astype(np.float64)
    pp = p.astype(np.float64)

    def attract(y):
        return _num_grad(lambda z: np.log(_phi(((z - pp) ** 2).sum(), a, b)), y)

    def repel(y):
        return _num_grad(lambda z: gamma * np.log(1.0 - _phi(((z - pp) ** 2).sum(), a, b)), y)
    clip = lambda v: np.clip(v, -4.0, 4.0)                           # noqa: E731
    f32 = lambda v: v.astype(np.float32).astype(np.float64)          # noqa: E731  (the embedding is stored in float32)
    alpha1 = alpha0 * (1.0 - 0.0 / 2.0)                              # epoch 0: nothing is due; alpha after it
    y = f32(y + alpha1 * clip(attract(y)))                           # epoch 1
    n_neg = int((1 - 1.0 / 5.0) / (1.0 / 5.0))                       # (n - next_neg) / eps_neg with eps_neg = 1/5
    assert n_neg == 4
    for _ in range(n_neg):
        y = f32(y + alpha1 * clip(repel(y)))
    # the closed-form repulsion carries +0.001 in its denominator (numerical guard of the published code): 1e-3 / d^2 relative
    assert np.abs(out[0] - y).max() < 2e-4 * np.abs(y).max(), (out[0], y)
    # and the attraction really pulls towards the tail, the repulsion pushes away from it
    assert np.dot(attract(y0[0].astype(np.float64)), pp - y0[0]) > 0 > np.dot(repel(y0[0].astype(np.float64)), pp - y0[0])


def test_exact_knn_equals_sklearn_brute_force(fitted):
    """The oracle's kNN (row-wise stable argsort of sklearn pairwise_distances) against sklearn's own brute-force neighbour search."""
    from sklearn.neighbors import NearestNeighbors
    X, um = fitted
    k = um._n_neighbors
    dist, idx = NearestNeighbors(n_neighbors=k, metric="cosine", algorithm="brute").fit(X).kneighbors(X)
    # the self-distance is ~1e-16 rather than exactly 0 in sklearn's search, so position 0 can hold a near-duplicate: compare as sets
    same_rows = sum(set(idx[i]) == set(um._knn_indices[i]) for i in range(len(X)))
    assert same_rows == len(X)
    assert np.abs(np.sort(dist, axis=1)[:, 1:] - um._knn_dists[:, 1:]).max() < 1e-6
