"""oracle/umap_oracle.py -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline); the product never imports it.

CPU restatement (numpy / scipy / scikit-learn / numba) of the UMAP stages the reference reaches through
``umap.UMAP(n_neighbors, n_components=3, min_dist=0.1, random_state=42, metric='cosine')
.fit_transform / .fit / .transform`` (debug_tda_pipeline.py:96-104, analyze_tda_over_layers.py:38-44,69,72,
analyze_adversarial_tda.py:85-93).  The arithmetic lives in the third-party package ``umap-learn``
(UNPINNED, README.md:28; not vendored, not installed here), so this file restates its published algorithm
(McInnes, Healy, Melville 2018; umap-learn 0.5.x defaults): exact kNN on the
``sklearn.pairwise_distances`` matrix, ``smooth_knn_dist``, ``compute_membership_strengths`` +
fuzzy union, ``find_ab_params``, spectral initialisation (incl. the multi-component layout), the epoch
schedule and the serial ``optimize_layout_euclidean`` SGD with the Tausworthe RNG.

PARITY UNPINNED: the reference ships neither UMAP inputs (all_activations.pt is git-ignored) nor a pinned
umap-learn version, so no golden vector exists for these stages; the deterministic stages (distances, kNN,
sigma/rho, membership graph, schedule) are checked by their defining invariants, the stochastic ones by
trustworthiness and downstream diagrams, as BASELINE.json's north_star prescribes.
"""
import warnings

import numba
import numpy as np
import scipy.sparse
import scipy.sparse.csgraph
import scipy.sparse.linalg
from scipy.optimize import curve_fit
from sklearn.metrics import pairwise_distances

SMOOTH_K_TOLERANCE = 1e-5
MIN_K_DIST_SCALE = 1e-3
INT32_MIN = np.iinfo(np.int32).min + 1
INT32_MAX = np.iinfo(np.int32).max - 1
DISCONNECTION_DISTANCES = {"correlation": 2, "cosine": 2, "hellinger": 1, "jaccard": 1, "dice": 1}


def find_ab_params(spread, min_dist):
    def curve(x, a, b):
        return 1.0 / (1.0 + a * x ** (2 * b))

    xv = np.linspace(0, spread * 3, 300)
    yv = np.zeros(xv.shape)
    yv[xv < min_dist] = 1.0
    yv[xv >= min_dist] = np.exp(-(xv[xv >= min_dist] - min_dist) / spread)
    params, _ = curve_fit(curve, xv, yv)
    return params[0], params[1]


def exact_knn(dmat, k):
    """fast_knn_indices + gather: per-row argsort, first k (self included at rank 0)."""
    idx = np.argsort(dmat, axis=1, kind="stable")[:, :k]
    dist = np.take_along_axis(dmat, idx, axis=1)
    idx = idx.astype(np.int64)
    idx[dist == np.inf] = -1
    return idx, dist.astype(np.float32)


@numba.njit(cache=True)
def smooth_knn_dist(distances, k, n_iter=64, local_connectivity=1.0, bandwidth=1.0):
    target = np.log2(k) * bandwidth
    rho = np.zeros(distances.shape[0], dtype=np.float32)
    result = np.zeros(distances.shape[0], dtype=np.float32)
    mean_distances = np.mean(distances)
    for i in range(distances.shape[0]):
        lo = 0.0
        hi = np.inf
        mid = 1.0
        ith = distances[i]
        nz = ith[ith > 0.0]
        if nz.shape[0] >= local_connectivity:
            index = int(np.floor(local_connectivity))
            interpolation = local_connectivity - index
            if index > 0:
                rho[i] = nz[index - 1]
                if interpolation > SMOOTH_K_TOLERANCE:
                    rho[i] += interpolation * (nz[index] - nz[index - 1])
            else:
                rho[i] = interpolation * nz[0]
        elif nz.shape[0] > 0:
            rho[i] = np.max(nz)
        for _ in range(n_iter):
            psum = 0.0
            for j in range(1, distances.shape[1]):
                d = distances[i, j] - rho[i]
                if d > 0:
                    psum += np.exp(-(d / mid))
                else:
                    psum += 1.0
            if np.fabs(psum - target) < SMOOTH_K_TOLERANCE:
                break
            if psum > target:
                hi = mid
                mid = (lo + hi) / 2.0
            else:
                lo = mid
                if hi == np.inf:
                    mid *= 2
                else:
                    mid = (lo + hi) / 2.0
        result[i] = mid
        if rho[i] > 0.0:
            mean_ith = np.mean(ith)
            if result[i] < MIN_K_DIST_SCALE * mean_ith:
                result[i] = MIN_K_DIST_SCALE * mean_ith
        else:
            if result[i] < MIN_K_DIST_SCALE * mean_distances:
                result[i] = MIN_K_DIST_SCALE * mean_distances
    return result, rho


@numba.njit(cache=True)
def compute_membership_strengths(knn_indices, knn_dists, sigmas, rhos, bipartite=False):
    n, k = knn_indices.shape
    rows = np.zeros(n * k, dtype=np.int32)
    cols = np.zeros(n * k, dtype=np.int32)
    vals = np.zeros(n * k, dtype=np.float32)
    for i in range(n):
        for j in range(k):
            if knn_indices[i, j] == -1:
                continue
            if (not bipartite) and knn_indices[i, j] == i:
                val = 0.0
            elif knn_dists[i, j] - rhos[i] <= 0.0 or sigmas[i] == 0.0:
                val = 1.0
            else:
                val = np.exp(-((knn_dists[i, j] - rhos[i]) / (sigmas[i])))
            rows[i * k + j] = i
            cols[i * k + j] = knn_indices[i, j]
            vals[i * k + j] = val
    return rows, cols, vals


def fuzzy_simplicial_set(knn_indices, knn_dists, n_samples, k, set_op_mix_ratio=1.0, local_connectivity=1.0):
    sigmas, rhos = smooth_knn_dist(knn_dists, float(k), local_connectivity=float(local_connectivity))
    rows, cols, vals = compute_membership_strengths(knn_indices, knn_dists, sigmas, rhos)
    result = scipy.sparse.coo_matrix((vals, (rows, cols)), shape=(n_samples, n_samples))
    result.eliminate_zeros()
    transpose = result.transpose()
    prod = result.multiply(transpose)
    result = set_op_mix_ratio * (result + transpose - prod) + (1.0 - set_op_mix_ratio) * prod
    result.eliminate_zeros()
    return result.tocsr(), sigmas, rhos


def make_epochs_per_sample(weights, n_epochs):
    result = -1.0 * np.ones(weights.shape[0], dtype=np.float64)
    n_samples = n_epochs * (weights / weights.max())
    result[n_samples > 0] = float(n_epochs) / np.float64(n_samples[n_samples > 0])
    return result


def _spectral_one(graph, dim, n_total, random_state):
    """Bottom non-trivial eigenvectors of the symmetric normalised Laplacian of one connected graph."""
    diag = np.asarray(graph.sum(axis=0)).ravel()
    n = graph.shape[0]
    D = scipy.sparse.spdiags(1.0 / np.sqrt(diag), 0, n, n)
    L = scipy.sparse.identity(n, dtype=np.float64) - D * graph * D
    k = dim + 1
    ncv = max(2 * k + 1, int(np.sqrt(n)))
    try:
        vals, vecs = scipy.sparse.linalg.eigsh(L, k, which="SM", ncv=ncv, tol=1e-4, v0=np.ones(n), maxiter=n_total * 5)
        order = np.argsort(vals)[1:k]
        return vecs[:, order]
    except (scipy.sparse.linalg.ArpackError, scipy.sparse.linalg.ArpackNoConvergence):
        warnings.warn("spectral initialisation failed; falling back to random initialisation")
        return None


def component_layout(data, n_components, labels, dim, random_state, metric):
    from sklearn.manifold import SpectralEmbedding
    centroids = np.empty((n_components, data.shape[1]), dtype=np.float64)
    for label in range(n_components):
        centroids[label] = data[labels == label].mean(axis=0)
    dm = pairwise_distances(centroids, metric=metric)
    affinity = np.exp(-(dm ** 2))
    emb = SpectralEmbedding(n_components=dim, affinity="precomputed", random_state=random_state).fit_transform(affinity)
    emb /= emb.max()
    return emb


def spectral_layout(data, graph, dim, random_state, metric="euclidean"):
    n_comp, labels = scipy.sparse.csgraph.connected_components(graph)
    n = graph.shape[0]
    if n_comp == 1:
        emb = _spectral_one(graph.tocsr().astype(np.float64), dim, n, random_state)
        if emb is None:
            return random_state.uniform(low=-10.0, high=10.0, size=(n, dim))
        return emb
    result = np.empty((n, dim), dtype=np.float32)
    if n_comp > 2 * dim:
        meta = component_layout(data, n_comp, labels, dim, random_state, metric)
    else:
        k = int(np.ceil(n_comp / 2.0))
        base = np.hstack([np.eye(k), np.zeros((k, dim - k))])
        meta = np.vstack([base, -base])[:n_comp]
    csr = graph.tocsr()
    for label in range(n_comp):
        mask = labels == label
        cg = csr[mask, :].tocsc()[:, mask].tocsr().astype(np.float64)
        dists = pairwise_distances([meta[label]], meta)
        data_range = dists[dists > 0.0].min() / 2.0
        m = cg.shape[0]
        if m < 2 * dim or m <= dim + 1:
            result[mask] = random_state.uniform(low=-data_range, high=data_range, size=(m, dim)) + meta[label]
            continue
        emb = _spectral_one(cg, dim, n, random_state)
        if emb is None:
            result[mask] = random_state.uniform(low=-data_range, high=data_range, size=(m, dim)) + meta[label]
        else:
            emb = emb * (data_range / np.max(np.abs(emb)))
            result[mask] = emb + meta[label]
    return result


@numba.njit(cache=True)
def tau_rand_int(state):
    state[0] = (((state[0] & 4294967294) << 12) & 0xFFFFFFFF) ^ ((((state[0] << 13) & 0xFFFFFFFF) ^ state[0]) >> 19)
    state[1] = (((state[1] & 4294967288) << 4) & 0xFFFFFFFF) ^ ((((state[1] << 2) & 0xFFFFFFFF) ^ state[1]) >> 25)
    state[2] = (((state[2] & 4294967280) << 17) & 0xFFFFFFFF) ^ ((((state[2] << 3) & 0xFFFFFFFF) ^ state[2]) >> 11)
    return state[0] ^ state[1] ^ state[2]


@numba.njit(cache=True)
def _clip(v):
    if v > 4.0:
        return 4.0
    elif v < -4.0:
        return -4.0
    return v


@numba.njit(cache=True)
def optimize_layout_euclidean(head_emb, tail_emb, head, tail, n_epochs, n_vertices, epochs_per_sample, a, b,
                              rng_state, gamma, initial_alpha, negative_sample_rate, move_other):
    dim = head_emb.shape[1]
    alpha = initial_alpha
    eps_neg = epochs_per_sample / negative_sample_rate
    next_neg = eps_neg.copy()
    next_pos = epochs_per_sample.copy()
    # per-vertex RNG state: global state perturbed by the bits of the first coordinate
    per = np.empty((head_emb.shape[0], 3), dtype=np.int64)
    first = head_emb[:, 0].astype(np.float64).view(np.int64)
    for v in range(head_emb.shape[0]):
        for t in range(3):
            per[v, t] = rng_state[t] + first[v]
    for n in range(n_epochs):
        for i in range(epochs_per_sample.shape[0]):
            if next_pos[i] <= n:
                j = head[i]
                k = tail[i]
                cur = head_emb[j]
                oth = tail_emb[k]
                d2 = 0.0
                for d in range(dim):
                    d2 += (cur[d] - oth[d]) ** 2
                if d2 > 0.0:
                    g = -2.0 * a * b * pow(d2, b - 1.0)
                    g /= a * pow(d2, b) + 1.0
                else:
                    g = 0.0
                for d in range(dim):
                    gd = _clip(g * (cur[d] - oth[d]))
                    cur[d] += gd * alpha
                    if move_other:
                        oth[d] += -gd * alpha
                next_pos[i] += epochs_per_sample[i]
                n_neg = int((n - next_neg[i]) / eps_neg[i])
                for _ in range(n_neg):
                    k = tau_rand_int(per[j]) % n_vertices
                    oth = tail_emb[k]
                    d2 = 0.0
                    for d in range(dim):
                        d2 += (cur[d] - oth[d]) ** 2
                    if d2 > 0.0:
                        g = 2.0 * gamma * b
                        g /= (0.001 + d2) * (a * pow(d2, b) + 1)
                    elif j == k:
                        continue
                    else:
                        g = 0.0
                    for d in range(dim):
                        if g > 0.0:
                            gd = _clip(g * (cur[d] - oth[d]))
                        else:
                            gd = 0.0
                        cur[d] += gd * alpha
                next_neg[i] += n_neg * eps_neg[i]
        alpha = initial_alpha * (1.0 - (float(n) / float(n_epochs)))
    return head_emb


class UMAPOracle:
    """Oracle twin of umap.UMAP for the arguments the reference passes (everything else at its default)."""

    def __init__(self, n_neighbors=15, n_components=2, metric="euclidean", n_epochs=None, learning_rate=1.0,
                 init="spectral", min_dist=0.1, spread=1.0, set_op_mix_ratio=1.0, local_connectivity=1.0,
                 repulsion_strength=1.0, negative_sample_rate=5, random_state=None):
        self.n_neighbors, self.n_components, self.metric, self.n_epochs = n_neighbors, n_components, metric, n_epochs
        self.learning_rate, self.init, self.min_dist, self.spread = learning_rate, init, min_dist, spread
        self.set_op_mix_ratio, self.local_connectivity = set_op_mix_ratio, local_connectivity
        self.repulsion_strength, self.negative_sample_rate, self.random_state = repulsion_strength, negative_sample_rate, random_state

    def distance_matrix(self, X, Y=None):
        dmat = pairwise_distances(X, Y, metric=self.metric)
        disc = DISCONNECTION_DISTANCES.get(self.metric, np.inf)
        dmat[dmat >= disc] = np.inf
        return dmat

    def fit(self, X):
        X = np.ascontiguousarray(X, dtype=np.float32)
        n = X.shape[0]
        rs = np.random.RandomState(self.random_state) if not isinstance(self.random_state, np.random.RandomState) else self.random_state
        self._a, self._b = find_ab_params(self.spread, self.min_dist)
        k = self.n_neighbors
        if k >= n:
            warnings.warn("n_neighbors is larger than the dataset size; truncating to X.shape[0] - 1")
            k = n - 1
        self._n_neighbors = k
        self._raw_data = X
        dmat = self.distance_matrix(X)
        self._knn_indices, self._knn_dists = exact_knn(dmat, k)
        self.graph_, self._sigmas, self._rhos = fuzzy_simplicial_set(self._knn_indices, self._knn_dists, n, k,
                                                                    self.set_op_mix_ratio, self.local_connectivity)
        graph = self.graph_.tocoo(copy=True)  # graph_ itself stays unpruned
        graph.sum_duplicates()
        n_epochs = self.n_epochs if self.n_epochs is not None else (500 if n <= 10000 else 200)
        graph.data[graph.data < (graph.data.max() / float(n_epochs))] = 0.0
        graph.eliminate_zeros()
        if isinstance(self.init, str) and self.init == "spectral":
            init = spectral_layout(X, graph, self.n_components, rs, metric=self.metric)
            expansion = 10.0 / np.abs(init).max()
            emb = (init * expansion).astype(np.float32) + rs.normal(scale=0.0001, size=[n, self.n_components]).astype(np.float32)
        elif isinstance(self.init, str) and self.init == "random":
            emb = rs.uniform(low=-10.0, high=10.0, size=(n, self.n_components)).astype(np.float32)
        else:
            emb = np.array(self.init, dtype=np.float32)
        self._init_embedding = emb.copy()
        eps = make_epochs_per_sample(graph.data, n_epochs)
        self._head, self._tail, self._eps = graph.row.copy(), graph.col.copy(), eps
        rng_state = rs.randint(INT32_MIN, INT32_MAX, 3).astype(np.int64)
        emb = (10.0 * (emb - np.min(emb, 0)) / (np.max(emb, 0) - np.min(emb, 0))).astype(np.float32, order="C")
        self.embedding_ = optimize_layout_euclidean(emb, emb, graph.row.astype(np.int64), graph.col.astype(np.int64), n_epochs, n,
                                                    eps, self._a, self._b, rng_state, self.repulsion_strength,
                                                    self.learning_rate, float(self.negative_sample_rate), True)
        self._rs = rs
        return self

    def fit_transform(self, X):
        return self.fit(X).embedding_

    def transform(self, X):
        X = np.ascontiguousarray(X, dtype=np.float32)
        if X.shape == self._raw_data.shape and np.array_equal(X, self._raw_data):
            return self.embedding_
        k = self._n_neighbors
        dmat = pairwise_distances(X, self._raw_data, metric=self.metric)
        idx = np.argsort(dmat, axis=1, kind="stable")[:, :k]
        dists = np.take_along_axis(dmat, idx, axis=1).astype(np.float32)
        sigmas, rhos = smooth_knn_dist(dists, float(k), local_connectivity=float(max(0.0, self.local_connectivity - 1.0)))
        rows, cols, vals = compute_membership_strengths(idx.astype(np.int64), dists, sigmas, rhos, bipartite=True)
        graph = scipy.sparse.coo_matrix((vals, (rows, cols)), shape=(X.shape[0], self._raw_data.shape[0]))
        from sklearn.preprocessing import normalize
        csr = normalize(graph.tocsr(), norm="l1")
        inds = csr.indices.reshape(X.shape[0], k)
        weights = csr.data.reshape(X.shape[0], k)
        emb = np.einsum("ik,ikd->id", weights, self.embedding_[inds]).astype(np.float32)
        n_epochs = 100 if graph.shape[0] <= 10000 else 30
        if self.n_epochs is not None:
            n_epochs = int(self.n_epochs // 3.0)
        graph.data[graph.data < (graph.data.max() / float(n_epochs))] = 0.0
        graph.eliminate_zeros()
        eps = make_epochs_per_sample(graph.data, n_epochs)
        rng_state = self._rs.randint(INT32_MIN, INT32_MAX, 3).astype(np.int64)
        return optimize_layout_euclidean(emb, self.embedding_.astype(np.float32, copy=True), graph.row.astype(np.int64),
                                         graph.col.astype(np.int64), n_epochs, graph.shape[1], eps, self._a, self._b, rng_state,
                                         self.repulsion_strength, self.learning_rate / 4.0, float(self.negative_sample_rate), False)
