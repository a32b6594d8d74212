"""GPU parity: the UMAP stage kernels (through the C ABI) vs the CPU restatement oracle/umap_oracle.py.

Tolerances are BASELINE.json's: kNN indices bit-exact on tie-free data, deterministic float stages within 1e-5
relative (sigma: bisection end points may differ by one step -> 1e-4), embeddings by trustworthiness and by the
downstream diagrams (the SGD is stochastic in both implementations)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def activations(n, d, rng, kind="torus"):
    from tda_multimodal_b200 import workloads
    if kind == "torus":
        z = workloads.torus_latent(n, rng, 0.05)
    else:
        c = rng.normal(0, 2.0, (6, 4))
        z = c[rng.integers(0, 6, n)] + rng.normal(0, 0.3, (n, 4))
    return workloads._embed(z, d, rng, noise=0.02, scale=7.0, offset=0.4)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.parametrize("n,k,metric", [(500, 15, "cosine"), (36, 6, "cosine"), (300, 40, "euclidean"), (130, 129, "cosine"),
                                        (501, 15, "cosine"), (37, 2, "euclidean"), (131, 16, "cosine"), (203, 17, "euclidean"),
                                        (1028, 15, "cosine")])
def test_knn_smooth_matches_oracle(torch_cuda, n, k, metric):
    torch = torch_cuda
    from oracle import umap_oracle as uo
    from tda_multimodal_b200 import umap_
    rng = np.random.default_rng(n + k)
    X = activations(n, 256, rng)
    orc = uo.UMAPOracle(n_neighbors=k, metric=metric)
    dmat = orc.distance_matrix(X).astype(np.float32)
    oidx, odist = uo.exact_knn(dmat, k)
    osig, orho = uo.smooth_knn_dist(odist, float(k))
    # the kernel under test gets the SAME float32 matrix, so indices must agree bit for bit
    idx, dist, sigma, rho = umap_.knn_smooth(torch.from_numpy(dmat).cuda()[None], k)
    assert np.array_equal(idx[0].cpu().numpy(), oidx)
    assert np.array_equal(dist[0].cpu().numpy(), odist)
    assert np.array_equal(rho[0].cpu().numpy(), orho)
    np.testing.assert_allclose(sigma[0].cpu().numpy(), osig, rtol=1e-4)
    # defining invariant of sigma (Appendix A.4), checked on the kernel's own output
    s, r, dd = sigma[0].cpu().numpy().astype(np.float64), rho[0].cpu().numpy().astype(np.float64), dist[0].cpu().numpy().astype(np.float64)
    psum = np.exp(-np.maximum(dd[:, 1:] - r[:, None], 0) / s[:, None]).sum(1)
    floored = s <= 1e-3 * dd.mean(1) * (1 + 1e-6)
    assert np.all((np.abs(psum - np.log2(k)) < 1e-3) | floored)


@pytest.mark.parametrize("rows,cols,k,batch", [(301, 500, 15, 3), (64, 2000, 15, 2), (77, 333, 8, 2), (50, 402, 20, 2),
                                               (40, 5000, 15, 1), (33, 3001, 10, 2), (20, 20000, 16, 1)])
def test_knn_smooth_batched_rectangular(torch_cuda, rows, cols, k, batch):
    """Row pairs / odd row counts / vector and scalar loads / several problems per launch (the transform and row-block
    callers pass rectangular blocks): indices, distances, rho exact; sigma equal to a float64 host replay of the bisection."""
    torch = torch_cuda
    from oracle import umap_oracle as uo
    from tda_multimodal_b200 import umap_
    rng = np.random.default_rng(rows * 7 + cols)
    dmat = rng.random((batch, rows, cols), dtype=np.float32)
    dmat[:, :, 0][:, ::5] = 0.0          # some zero distances (rho skips them)
    dmat[0, 3, :] = np.inf               # a row with fewer than k finite entries
    dmat[0, 3, 5:9] = [0.5, 0.25, 0.75, 0.125]
    dmat[batch - 1, 7, 10:14] = 0.3      # ties: smaller column first
    idx, dist, sigma, rho = umap_.knn_smooth(torch.from_numpy(dmat).cuda(), k)
    for b in range(batch):
        oidx, odist = uo.exact_knn(dmat[b], k)
        osig, orho = uo.smooth_knn_dist(odist, float(k))
        assert np.array_equal(idx[b].cpu().numpy(), oidx)
        assert np.array_equal(dist[b].cpu().numpy(), odist)
        assert np.array_equal(rho[b].cpu().numpy(), orho)
        fin = np.isfinite(odist).all(1)
        np.testing.assert_allclose(sigma[b].cpu().numpy()[fin], osig[fin], rtol=1e-5)


def test_knn_from_tensor_core_distances(torch_cuda):
    """kNN on the 3xTF32 distance matrix: indices equal the float64 oracle's wherever its gaps exceed the GEMM error."""
    torch = torch_cuda
    from tda_multimodal_b200 import umap_
    rng = np.random.default_rng(5)
    X = activations(700, 4096, rng)
    D = umap_.distance_matrix(torch.from_numpy(X).cuda()[None], metric="cosine")
    idx = umap_.knn_smooth(D, 15)[0][0].cpu().numpy()
    Xn = X.astype(np.float64)
    Xn /= np.linalg.norm(Xn, axis=1, keepdims=True)
    want = np.clip(1 - Xn @ Xn.T, 0, 2)
    np.fill_diagonal(want, 0)
    oi = np.argsort(want, axis=1, kind="stable")[:, :16]
    od = np.take_along_axis(want, oi, 1)
    clear = np.diff(od, axis=1).min(axis=1) > 4e-6
    assert clear.mean() > 0.8
    assert np.array_equal(idx[clear], oi[clear][:, :15])
    assert (idx[:, 0] == np.arange(700)).all()  # self first, at distance exactly 0


@pytest.mark.parametrize("n,k", [(400, 15), (36, 6)])
def test_fuzzy_graph_and_schedule_match_oracle(torch_cuda, n, k):
    torch = torch_cuda
    from oracle import umap_oracle as uo
    from tda_multimodal_b200 import umap_
    rng = np.random.default_rng(17 + n)
    X = activations(n, 128, rng, kind="clusters")
    orc = uo.UMAPOracle(n_neighbors=k, metric="cosine")
    dmat = orc.distance_matrix(X).astype(np.float32)
    oidx, odist = uo.exact_knn(dmat, k)
    graph, osig, orho = uo.fuzzy_simplicial_set(oidx, odist, n, k)
    idx, dist, sigma, rho = umap_.knn_smooth(torch.from_numpy(dmat).cuda()[None], k)
    head, tail, weight, eps = umap_.fuzzy_graph(idx, dist, sigma, rho, 500)
    h, t, w, e = (a[0].cpu().numpy() for a in (head, tail, weight, eps))
    import scipy.sparse
    keep = w > 0
    got = scipy.sparse.coo_matrix((w[keep], (h[keep], t[keep])), shape=(n, n)).tocsr()
    assert (got - got.T).nnz == 0 or abs(got - got.T).max() < 1e-7  # symmetric
    diff = abs(got - graph.astype(np.float32))
    assert got.nnz == graph.nnz and diff.max() < 2e-5
    # schedule: eps = max_w / w for entries >= max_w / n_epochs, -1 otherwise
    g = graph.tocoo()
    o_eps = uo.make_epochs_per_sample(np.where(g.data < g.data.max() / 500.0, 0.0, g.data), 500)
    o_map = {(int(r), int(c)): v for r, c, v in zip(g.row, g.col, o_eps)}
    for hh, tt, ee, ww in zip(h[keep], t[keep], e[keep], w[keep]):
        oe = o_map[(int(hh), int(tt))]
        assert (ee < 0 and oe < 0) or abs(ee - oe) <= 1e-4 * abs(oe) or abs(ww - g.data.max() / 500.0) < 1e-6
    assert (e[~keep] < 0).all() and (e[keep][e[keep] > 0] >= 1.0 - 1e-6).all()


def test_components_and_spectral_vs_scipy(torch_cuda):
    torch = torch_cuda
    import scipy.sparse
    import scipy.sparse.csgraph
    import scipy.sparse.linalg
    from tda_multimodal_b200 import umap_, _lib
    rng = np.random.default_rng(23)
    n, k, dim = 600, 10, 3
    X = np.concatenate([activations(400, 64, rng), activations(200, 64, rng, kind="clusters") + 50.0])
    Xd = torch.from_numpy(X).cuda()[None]
    D = umap_.distance_matrix(Xd, metric="euclidean")
    idx, dist, sigma, rho = umap_.knn_smooth(D, k)
    head, tail, weight, eps = umap_.fuzzy_graph(idx, dist, sigma, rho, 500)
    L = _lib.lib()
    comp = torch.empty((1, n), dtype=torch.int32, device="cuda")
    ncomp = torch.empty((1,), dtype=torch.int32, device="cuda")
    csize = torch.empty((1, n), dtype=torch.int32, device="cuda")
    deg = torch.empty((1, n), dtype=torch.float32, device="cuda")
    ws0 = torch.empty(12 * n, dtype=torch.uint8, device="cuda")
    _lib.check(L.tda_graph_components(_lib.ptr(head), _lib.ptr(tail), _lib.ptr(weight), _lib.ptr(eps), head.shape[1], n, 1, _lib.ptr(comp),
                                      _lib.ptr(ncomp), _lib.ptr(csize), _lib.ptr(deg), _lib.ptr(ws0), 12 * n, _lib.stream_ptr()))
    h, t, w, e = (a[0].cpu().numpy() for a in (head, tail, weight, eps))
    keep = e > 0
    G = scipy.sparse.coo_matrix((w[keep].astype(np.float64), (h[keep], t[keep])), shape=(n, n)).tocsr()
    nc, labels = scipy.sparse.csgraph.connected_components(G)
    assert int(ncomp[0]) == nc and nc >= 2
    assert np.array_equal(comp[0].cpu().numpy(), labels)  # scipy numbers components by smallest member too
    np.testing.assert_allclose(deg[0].cpu().numpy(), np.asarray(G.sum(1)).ravel(), rtol=1e-5)
    # eigenvectors of every large component: residual of A v = lambda v and agreement of lambda with ARPACK
    maxcomp = nc
    ws_bytes = int(L.tda_spectral_workspace_bytes(n, 1, maxcomp, head.shape[1]))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    Y = torch.zeros((1, n, dim), dtype=torch.float32, device="cuda")
    ev = torch.zeros((1, maxcomp, 4), dtype=torch.float32, device="cuda")
    _lib.check(L.tda_spectral_embed(_lib.ptr(head), _lib.ptr(tail), _lib.ptr(weight), _lib.ptr(eps), head.shape[1], n, dim, 1, _lib.ptr(comp),
                                    _lib.ptr(ncomp), _lib.ptr(csize), _lib.ptr(deg), maxcomp, 2 * dim, 1, _lib.ptr(Y), _lib.ptr(ev),
                                    _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
    Yh, evh = Y[0].cpu().numpy().astype(np.float64), ev[0].cpu().numpy()
    for c in range(nc):
        mask = labels == c
        if mask.sum() < 50:
            continue
        Gc = G[mask][:, mask]
        d = np.asarray(Gc.sum(1)).ravel()
        A = scipy.sparse.diags(1 / np.sqrt(d)) @ Gc @ scipy.sparse.diags(1 / np.sqrt(d))
        vals = np.sort(scipy.sparse.linalg.eigsh(A, k=dim + 1, which="LA")[0])[::-1][1:]
        np.testing.assert_allclose(evh[c, :dim], vals, atol=2e-3)
        V = Yh[mask]
        assert np.allclose(np.linalg.norm(V, axis=0), 1.0, atol=1e-3)
        assert np.abs(V.T @ V - np.eye(dim)).max() < 1e-2
        for a in range(dim):
            res = A @ V[:, a] - evh[c, a] * V[:, a]
            assert np.linalg.norm(res) < 3e-2, (c, a, np.linalg.norm(res))
        # the reported Ritz residual estimate (last slot of evals) bounds what was just measured, within fp32 rounding of the vectors
        worst = max(np.linalg.norm(A @ V[:, a] - evh[c, a] * V[:, a]) for a in range(dim))
        assert evh[c, 3] >= 0 and abs(evh[c, 3] - worst) < 5e-3 + 0.5 * worst, (c, evh[c, 3], worst)


def _trust(X, Y, metric):
    from sklearn.manifold import trustworthiness
    return trustworthiness(X, Y, n_neighbors=10, metric=metric)


@pytest.mark.parametrize("kind,n,k", [("torus", 800, 15), ("clusters", 600, 15), ("torus", 36, 6)])
def test_fit_transform_trustworthiness_vs_oracle(torch_cuda, kind, n, k):
    from oracle import umap_oracle as uo
    from tda_multimodal_b200.umap_ import UMAP
    rng = np.random.default_rng(31 + n)
    X = activations(n, 512, rng, kind=kind)
    Yo = uo.UMAPOracle(n_neighbors=k, n_components=3, min_dist=0.1, metric="cosine", random_state=42).fit_transform(X)
    um = UMAP(n_neighbors=k, n_components=3, min_dist=0.1, random_state=42, metric="cosine")
    Yg = um.fit_transform(X.astype(np.float64))  # the reference passes float64 (debug_tda_pipeline.py:64)
    assert Yg.shape == (n, 3) and Yg.dtype == np.float32 and np.isfinite(Yg).all()
    if n > 50:
        to, tg = _trust(X, Yo, "cosine"), _trust(X, Yg, "cosine")
        assert tg >= to - 0.03, (tg, to)
    assert um.graph_.shape == (n, n) and um._sigmas.shape == (n,) and um.embedding_ is Yg


@pytest.mark.parametrize("cluster", [8, 4, 0])
def test_cluster_lanczos_on_connected_graph(torch_cuda, tda_option, cluster):
    """Connected graph (maxcomp = 1): the thread-block-cluster Lanczos kernel (spectral_cluster = 8 / 4) and the one-CTA kernel (0)
    against ARPACK: eigenvalues, orthonormality, residuals; the cluster kernel is bit-reproducible."""
    torch = torch_cuda
    import scipy.sparse
    import scipy.sparse.linalg
    from tda_multimodal_b200 import umap_, _lib
    tda_option("spectral_cluster", cluster)
    rng = np.random.default_rng(29)
    n, k, dim = 1500, 15, 3
    Xd = torch.from_numpy(np.stack([activations(n, 64, rng), activations(n, 64, rng)])).cuda()
    D = umap_.distance_matrix(Xd, metric="cosine")
    idx, dist, sigma, rho = umap_.knn_smooth(D, k)
    head, tail, weight, eps = umap_.fuzzy_graph(idx, dist, sigma, rho, 500)
    L = _lib.lib()
    B = 2
    comp = torch.empty((B, n), dtype=torch.int32, device="cuda")
    ncomp = torch.empty((B,), dtype=torch.int32, device="cuda")
    csize = torch.empty((B, n), dtype=torch.int32, device="cuda")
    deg = torch.empty((B, n), dtype=torch.float32, device="cuda")
    ws0 = torch.empty(12 * B * n, dtype=torch.uint8, device="cuda")
    _lib.check(L.tda_graph_components(_lib.ptr(head), _lib.ptr(tail), _lib.ptr(weight), _lib.ptr(eps), head.shape[1], n, B, _lib.ptr(comp),
                                      _lib.ptr(ncomp), _lib.ptr(csize), _lib.ptr(deg), _lib.ptr(ws0), 12 * B * n, _lib.stream_ptr()))
    assert ncomp.cpu().tolist() == [1, 1]
    ws_bytes = int(L.tda_spectral_workspace_bytes(n, B, 1, head.shape[1]))
    outs = []
    for rep in range(2):
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        Y = torch.zeros((B, n, dim), dtype=torch.float32, device="cuda")
        ev = torch.zeros((B, 1, 4), dtype=torch.float32, device="cuda")
        _lib.check(L.tda_spectral_embed(_lib.ptr(head), _lib.ptr(tail), _lib.ptr(weight), _lib.ptr(eps), head.shape[1], n, dim, B, _lib.ptr(comp),
                                        _lib.ptr(ncomp), _lib.ptr(csize), _lib.ptr(deg), 1, 1, 1, _lib.ptr(Y), _lib.ptr(ev),
                                        _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
        outs.append((Y.cpu().numpy(), ev.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0])   # fixed summation order / fixed-point accumulation: reproducible
    for b in range(B):
        h, t, w, e = (a[b].cpu().numpy() for a in (head, tail, weight, eps))
        keep = e > 0
        G = scipy.sparse.coo_matrix((w[keep].astype(np.float64), (h[keep], t[keep])), shape=(n, n)).tocsr()
        d = np.asarray(G.sum(1)).ravel()
        A = scipy.sparse.diags(1 / np.sqrt(d)) @ G @ scipy.sparse.diags(1 / np.sqrt(d))
        vals = np.sort(scipy.sparse.linalg.eigsh(A, k=dim + 1, which="LA")[0])[::-1][1:]
        V, evh = outs[0][0][b].astype(np.float64), outs[0][1][b, 0]
        np.testing.assert_allclose(evh[:dim], vals, atol=2e-3)
        assert np.abs(V.T @ V - np.eye(dim)).max() < 1e-2
        for a in range(dim):
            assert np.linalg.norm(A @ V[:, a] - evh[a] * V[:, a]) < 3e-2


def test_spectral_init_on_device_handles_components(torch_cuda):
    """tda_spectral_init (no host round trip; the path pipeline.layer_sweep takes): a batch mixing a connected cloud, a cloud with
    three far-apart pieces and one with a piece too small for a spectral layout.  Component counts equal scipy's, every cloud is
    laid out on the device (status 0), the pieces of a cloud sit around different meta positions (multi_component_layout), and the
    embedding that follows is as trustworthy as the one of the host path."""
    torch = torch_cuda
    from tda_multimodal_b200 import umap_
    rng = np.random.default_rng(31)
    n, d = 900, 64
    a = activations(n, d, rng)
    b = np.concatenate([activations(300, d, rng), activations(300, d, rng) + 40.0, activations(300, d, rng) - 40.0]).astype(np.float32)
    c = np.concatenate([activations(n - 4, d, rng), activations(4, d, rng) * 0.01 + 90.0]).astype(np.float32)
    X = np.stack([a, b, c])
    Xd = torch.from_numpy(X).cuda()
    Y, status = umap_.umap_fit_batch(Xd, n_neighbors=10, n_components=3, metric="euclidean", random_state=42, defer_component_check=True)
    assert status.cpu().tolist() == [0, 0, 0]
    Y = Y.cpu().numpy()
    Yh = umap_.umap_fit_batch(Xd, n_neighbors=10, n_components=3, metric="euclidean", random_state=42, host_spectral_init=True).cpu().numpy()
    Y2 = umap_.umap_fit_batch(Xd, n_neighbors=10, n_components=3, metric="euclidean", random_state=42).cpu().numpy()
    assert np.array_equal(Y, Y2), "fit with and without the deferred component check must be the same embedding, bit for bit"
    assert np.isfinite(Y).all()
    for i in range(3):
        assert _trust(X[i], Y[i], "euclidean") >= _trust(X[i], Yh[i], "euclidean") - 0.03
    # the three pieces of cloud b end up apart: centroid distances well above the pieces' radii
    cents = np.stack([Y[1][j * 300:(j + 1) * 300].mean(0) for j in range(3)])
    rad = max(np.linalg.norm(Y[1][j * 300:(j + 1) * 300] - cents[j], axis=1).mean() for j in range(3))
    dmin = min(np.linalg.norm(cents[i] - cents[j]) for i in range(3) for j in range(i))
    assert dmin > 1.5 * rad, (dmin, rad)


def test_sgd_deterministic_kernel_matches_atomic_kernel(torch_cuda, tda_option):
    """The default SGD (sgd_mode 0: a thread-block cluster per cloud, every vertex sums its own displacement in list order, no
    atomics, all epochs in one launch) against the per-epoch kernels with float atomics (sgd_mode 3) on a batch of 8 clouds:
    same schedule and update rule, so both embeddings must be finite, inside the clip box and equally trustworthy -- and the
    default one must be BIT-IDENTICAL from run to run (random_state=42 reproduces, as in the reference; debug_tda_pipeline.py:100)."""
    torch = torch_cuda
    from tda_multimodal_b200 import umap_
    rng = np.random.default_rng(91)
    X = np.stack([activations(400, 128, rng, kind="torus" if i % 2 else "clusters") for i in range(8)])
    Xd = torch.from_numpy(X).cuda()
    out = {}
    for mode in (0, 3, 0):
        tda_option("sgd_mode", mode)
        Y = umap_.umap_fit_batch(Xd, n_neighbors=15, n_components=3, metric="cosine", random_state=42).cpu().numpy()
        assert Y.shape == (8, 400, 3) and np.isfinite(Y).all() and np.abs(Y).max() < 100
        if mode == 0 and 0 in out:
            assert np.array_equal(out[0], Y), "deterministic SGD differs between two runs with the same seed"
        out[mode] = Y
    t0 = [_trust(X[i], out[0][i], "cosine") for i in range(8)]
    t3 = [_trust(X[i], out[3][i], "cosine") for i in range(8)]
    assert min(t0) > 0.8 and np.mean(t0) >= np.mean(t3) - 0.02, (t0, t3)
    rel = np.linalg.norm(out[0] - out[3], axis=2).mean() / np.linalg.norm(out[3] - out[3].mean(1, keepdims=True), axis=2).mean()
    print("mean point displacement between the two kernels / cloud radius:", rel, "trust", np.mean(t0), np.mean(t3))


@pytest.mark.parametrize("cluster", [1, 2, 8])
def test_sgd_cluster_sizes(torch_cuda, tda_option, cluster):
    """The deterministic kernel deals the vertices to the CTAs of a cluster in whole tiles, so a tile's fired entries fall into the
    same 32-lane batches whatever the cluster size: the embedding is BIT-IDENTICAL for 1, 2, 4 and 8 CTAs per cloud (the library
    picks 8 for small batches and 4 otherwise: the same cloud must not depend on what else is in the batch)."""
    torch = torch_cuda
    from tda_multimodal_b200 import umap_
    rng = np.random.default_rng(17)
    X = np.stack([activations(333, 64, rng, kind="torus") for _ in range(3)])
    Xd = torch.from_numpy(X).cuda()
    tda_option("sgd_cluster", 4)
    want = umap_.umap_fit_batch(Xd, n_neighbors=10, n_components=3, metric="cosine", random_state=7).cpu().numpy()
    tda_option("sgd_cluster", cluster)
    got = umap_.umap_fit_batch(Xd, n_neighbors=10, n_components=3, metric="cosine", random_state=7).cpu().numpy()
    assert np.array_equal(got, want)
    # ... and a cloud embedded alone equals the same cloud embedded inside a batch
    tda_option("sgd_cluster", 0)
    alone = umap_.umap_fit_batch(Xd[1:2], n_neighbors=10, n_components=3, metric="cosine", random_state=7).cpu().numpy()
    assert np.array_equal(alone[0], want[1])
    tg = [_trust(X[i], got[i], "cosine") for i in range(3)]
    assert min(tg) > 0.8


def test_transform_deterministic_and_equals_atomic_kernel(torch_cuda, tda_option):
    """transform(): the warp-per-point kernel (sgd_mode 0) gives bit-identical results run to run and the same quality as the
    per-epoch atomic kernel (both optimise every query point against the fixed training embedding)."""
    torch = torch_cuda
    from tda_multimodal_b200 import umap_
    rng = np.random.default_rng(5)
    Xtr, Xq = activations(500, 96, rng, kind="torus"), activations(200, 96, np.random.default_rng(6), kind="torus")
    um = umap_.UMAP(n_neighbors=15, n_components=3, metric="cosine", random_state=42).fit(Xtr)
    out = {}
    for mode in (0, 3, 0):
        tda_option("sgd_mode", mode)
        Yq = um.transform(Xq)
        assert Yq.shape == (200, 3) and np.isfinite(Yq).all()
        if mode == 0 and 0 in out:
            assert np.array_equal(out[0], Yq)
        out[mode] = Yq
    rad = np.linalg.norm(um.embedding_ - um.embedding_.mean(0), axis=1).mean()
    assert np.median(np.linalg.norm(out[0] - out[3], axis=1)) < 0.25 * rad


def test_downstream_diagrams_clusters(torch_cuda):
    """Downstream-diagram parity (north_star: UMAP is compared by trustworthiness and by the diagrams that follow):
    5 well separated clusters in 512-d -> UMAP 3-D -> Rips.  Both implementations must see exactly 4 dominant H0
    gaps (the cluster structure), and Rips on the SAME cloud is bit-exact between the GPU and the oracle."""
    from oracle import umap_oracle as uo, rips as orips
    from tda_multimodal_b200.umap_ import UMAP
    from tda_multimodal_b200.rips import ripser
    from tda_multimodal_b200 import workloads
    rng = np.random.default_rng(77)
    centers = rng.normal(0, 1.0, (5, 6))
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
    z = centers[rng.integers(0, 5, 400)] + rng.normal(0, 0.03, (400, 6))
    X = workloads._embed(z, 512, rng, noise=0.002, scale=5.0)
    Yg = UMAP(n_neighbors=15, n_components=3, random_state=42, metric="cosine").fit_transform(X)
    Yo = uo.UMAPOracle(n_neighbors=15, n_components=3, metric="cosine", random_state=42).fit_transform(X)
    for Y, rip in ((Yg, ripser), (Yo, orips.ripser)):
        d0 = rip(Y, maxdim=1)["dgms"][0]
        deaths = np.sort(d0[np.isfinite(d0[:, 1]), 1])[::-1]
        assert deaths[3] > 3 * deaths[4], deaths[:6]  # 4 big merges, then within-cluster scale
    a, b = ripser(Yg, maxdim=1)["dgms"], orips.ripser(Yg, maxdim=1)["dgms"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_transform_fit_once_transform_many(torch_cuda):
    """analyze_tda_over_layers.py:67-72: fit on one layer, transform the others."""
    from tda_multimodal_b200.umap_ import UMAP
    rng = np.random.default_rng(41)
    X = activations(500, 256, rng)
    Xnew = X[:100] + rng.normal(0, 1e-3, (100, 256)).astype(np.float32)
    um = UMAP(n_neighbors=18, n_components=3, min_dist=0.1, random_state=42, metric="cosine").fit(X)
    assert um.transform(X) is um.embedding_  # the training data itself returns the stored embedding
    Yn = um.transform(Xnew)
    assert Yn.shape == (100, 3) and Yn.dtype == np.float32
    # a slightly perturbed training point lands close to where that training point sits
    spread = np.linalg.norm(um.embedding_ - um.embedding_.mean(0), axis=1).mean()
    assert np.median(np.linalg.norm(Yn - um.embedding_[:100], axis=1)) < 0.25 * spread


def test_validation_and_unsupported_options(torch_cuda):
    from tda_multimodal_b200.umap_ import UMAP
    X = np.random.default_rng(0).normal(size=(30, 16)).astype(np.float32)
    with pytest.raises(ValueError):
        UMAP(n_neighbors=1).fit(X)
    with pytest.raises(ValueError):
        UMAP(min_dist=2.0, spread=1.0).fit(X)
    with pytest.raises(NotImplementedError):
        UMAP(densmap=True).fit(X)
    with pytest.raises(NotImplementedError):
        UMAP(metric="manhattan").fit(X)
    with pytest.warns(UserWarning):
        Y = UMAP(n_neighbors=50, n_components=3).fit_transform(X)  # n_neighbors > n: truncated with a warning
    assert Y.shape == (30, 3)


def test_pipeline_layer_sweep_matches_oracle_rips(torch_cuda):
    """The batched sweep: per layer, Rips of the device pipeline == oracle Rips on the same embedding (bit-exact),
    and the stats record has the reference's schema."""
    torch = torch_cuda
    import json
    from oracle import rips as orips
    from tda_multimodal_b200 import pipeline, workloads
    X = workloads.c3_layers(layers=[0, 13, 31], n=400, d=512)
    out = pipeline.layer_sweep_host(X, n_neighbors=15)
    assert out["embedding"].shape == (3, 400, 3)
    for l in range(3):
        want = orips.ripser(out["embedding"][l], maxdim=1)["dgms"]
        got = out["results"][l]["dgms"]
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        rec = pipeline.stats_record(l, got)
        assert list(rec) == ["layer", "n_h1_features", "max_h1_persistence", "all_h1_persistence_values", "n_h0_features", "max_h0_persistence"]
        json.dumps(rec)
        assert rec["n_h0_features"] == 1
