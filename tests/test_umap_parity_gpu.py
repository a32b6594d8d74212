"""GPU: UMAP parity beyond trustworthiness (north_star: "UMAP embeddings are compared by trustworthiness and by the downstream
diagrams rather than elementwise").  The CPU side is oracle/umap_oracle.py (umap-learn 0.5 restated; parity unpinned: the
reference ships no UMAP inputs), the GPU side the library's default kernels.
  * downstream H1: a torus embedded by both; the dominant H1 bars of the two embeddings (persistence / cloud diameter) must agree
    in number and size (medians over seeds; UMAP is stochastic, single embeddings of the SAME implementation differ by +-0.03);
  * transform: the oracle's transform() run on the GPU fit's own state (training data, embedding, a, b) must put the query points
    where the GPU transform puts them;
  * multi-component initialisation: component meta positions of the device path, the host path and the oracle's spectral_layout."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _norm_pers(Y, dgm1):
    diam = float(np.linalg.norm(Y.max(0) - Y.min(0)))
    p = np.sort(dgm1[:, 1] - dgm1[:, 0])[::-1] / diam
    return np.concatenate([p, np.zeros(4)])[:4]


def test_downstream_h1_bars_torus():
    import torch
    from sklearn.manifold import trustworthiness
    from oracle import umap_oracle as uo
    from tda_multimodal_b200 import rips, umap_, workloads
    seeds = (1, 2, 3)
    X = np.stack([workloads._embed(workloads.torus_latent(900, np.random.default_rng(s), 0.03), 64, np.random.default_rng(100 + s), noise=0.01)
                  for s in seeds])
    Yg = umap_.umap_fit_batch(torch.from_numpy(X).cuda(), n_neighbors=15, n_components=3, min_dist=0.1, metric="euclidean",
                              random_state=42).cpu().numpy()
    Yo = np.stack([uo.UMAPOracle(n_neighbors=15, n_components=3, min_dist=0.1, metric="euclidean", random_state=42).fit_transform(X[i])
                   for i in range(len(seeds))])
    both = np.concatenate([Yg, Yo]).astype(np.float32)
    res = rips.rips_batch(rips.pdist_lowdim(torch.from_numpy(both).cuda()), maxdim=1)
    pg = np.stack([_norm_pers(Yg[i], res[i]["dgms"][1]) for i in range(len(seeds))])
    po = np.stack([_norm_pers(Yo[i], res[len(seeds) + i]["dgms"][1]) for i in range(len(seeds))])
    print("GPU   top-4 H1 persistence / diameter per seed:\n", np.round(pg, 3), "\noracle:\n", np.round(po, 3))
    mg, mo = np.median(pg, axis=0), np.median(po, axis=0)
    # the torus has two generators: two dominant bars in both, of the same size; the third bar clearly smaller in both
    assert abs(mg[0] - mo[0]) <= 0.05 and abs(mg[1] - mo[1]) <= 0.05, (mg, mo)
    assert mg[1] >= 0.08 and mo[1] >= 0.08
    assert int((mg >= 0.08).sum()) == int((mo >= 0.08).sum()) or abs(mg[2] - mo[2]) <= 0.04, (mg, mo)
    tg = [trustworthiness(X[i], Yg[i], n_neighbors=10) for i in range(len(seeds))]
    to = [trustworthiness(X[i], Yo[i], n_neighbors=10) for i in range(len(seeds))]
    assert np.mean(tg) >= np.mean(to) - 0.02, (tg, to)   # +-0.02


def test_transform_against_oracle_on_same_state():
    from sklearn.manifold import trustworthiness
    from oracle import umap_oracle as uo
    from tda_multimodal_b200 import umap_, workloads
    rng = np.random.default_rng(7)
    Xtr = workloads._embed(workloads.torus_latent(700, rng, 0.03), 48, rng, noise=0.01)
    um = umap_.UMAP(n_neighbors=15, n_components=3, min_dist=0.1, metric="euclidean", random_state=42).fit(Xtr)
    # queries = perturbed training points (on the same manifold)
    pick = np.random.default_rng(9).choice(700, 250, replace=False)
    Xq = (Xtr[pick] + np.random.default_rng(10).normal(0, 0.01, (250, 48))).astype(np.float32)
    Yq_gpu = um.transform(Xq)
    orc = uo.UMAPOracle(n_neighbors=15, n_components=3, min_dist=0.1, metric="euclidean", random_state=42)
    orc._raw_data = np.ascontiguousarray(Xtr, dtype=np.float32)
    orc.embedding_ = np.ascontiguousarray(um.embedding_, dtype=np.float32)
    orc._a, orc._b, orc._n_neighbors = um._a, um._b, 15
    orc._rs = np.random.RandomState(42)
    Yq_or = orc.transform(Xq)
    rad = np.linalg.norm(um.embedding_ - um.embedding_.mean(0), axis=1).mean()
    disp = np.linalg.norm(Yq_gpu - Yq_or, axis=1)
    print("median |GPU - oracle| / cloud radius:", np.median(disp) / rad, "90th pct:", np.percentile(disp, 90) / rad)
    assert np.median(disp) <= 0.05 * rad and np.percentile(disp, 90) <= 0.15 * rad
    # and both keep the query points next to their training neighbours
    joint_g = np.concatenate([um.embedding_, Yq_gpu]); joint_o = np.concatenate([um.embedding_, Yq_or]); Xj = np.concatenate([Xtr, Xq])
    assert trustworthiness(Xj, joint_g, n_neighbors=10) >= trustworthiness(Xj, joint_o, n_neighbors=10) - 0.02


def test_multi_component_init_positions():
    """three far-apart pieces: the device path (tda_spectral_init), the host path and the oracle's spectral_layout put the pieces
    around the same meta positions (+-e_k, after the common [0,10] rescale of the initialisation)."""
    import torch
    from oracle import umap_oracle as uo
    from tda_multimodal_b200 import umap_, workloads
    rng = np.random.default_rng(3)
    pieces = [workloads._embed(workloads.torus_latent(200, rng, 0.05), 32, rng, noise=0.01) + off for off in (0.0, 30.0, -30.0)]
    X = np.concatenate(pieces).astype(np.float32)
    Xd = torch.from_numpy(X).cuda()[None]
    lab = np.repeat(np.arange(3), 200)
    _, st_host = umap_.umap_fit_batch(Xd, n_neighbors=10, n_components=3, metric="euclidean", random_state=42, n_epochs=11, return_state=True, host_spectral_init=True)
    init_host = st_host["init"][0].cpu().numpy()
    # device path: the initialisation is the embedding after zero epochs
    Yd, status = umap_.umap_fit_batch(Xd, n_neighbors=10, n_components=3, metric="euclidean", random_state=42, n_epochs=0, defer_component_check=True)
    assert int(status.max()) == 0
    init_dev = Yd[0].cpu().numpy()
    orc = uo.UMAPOracle(n_neighbors=10, n_components=3, metric="euclidean", random_state=42, n_epochs=11).fit(X)
    e = orc._init_embedding
    init_or = 10.0 * (e - e.min(0)) / (e.max(0) - e.min(0))

    def cents(Y):
        return np.stack([Y[lab == c].mean(0) for c in range(3)])
    ch, cd, co = cents(init_host), cents(init_dev), cents(init_or)
    print("component centroids host / device / oracle:\n", np.round(ch, 2), "\n", np.round(cd, 2), "\n", np.round(co, 2))
    assert np.abs(ch - co).max() < 1.0 and np.abs(cd - co).max() < 1.0   # [0,10] box: meta positions 5 apart


def _nine_pieces():
    from tda_multimodal_b200 import workloads
    rng = np.random.default_rng(5)
    offs = rng.normal(0, 1.0, (9, 24)) * 0.25       # centroid distances ~1.7: affinities exp(-d^2) of a few percent
    pieces = [workloads._embed(workloads.torus_latent(120, rng, 0.05), 24, rng, noise=0.01) * 0.05 + off for off in offs]
    return pieces, np.concatenate(pieces).astype(np.float32), np.repeat(np.arange(9), 120)


def test_many_component_init_positions():
    """nine separate pieces (> 2*dim components: umap-learn's component_layout = spectral embedding of the component centroids):
    the device path (tda_spectral_init with the data: centroid_kernel + meta_layout_kernel, no host round trip), the host path and
    the oracle's spectral_layout put the pieces at the same meta positions after the common [0,10] rescale of the initialisation.
    The pieces sit at generic positions, so the centroid Laplacian has distinct eigenvalues and the layout is unique up to the
    sign convention (sklearn's deterministic flip, which all three follow)."""
    import torch
    from oracle import umap_oracle as uo
    from tda_multimodal_b200 import umap_
    pieces, X, lab = _nine_pieces()
    Xd = torch.from_numpy(X).cuda()[None]
    _, st_host = umap_.umap_fit_batch(Xd, n_neighbors=10, n_components=3, metric="euclidean", random_state=42, n_epochs=11, return_state=True, host_spectral_init=True)
    init_host = st_host["init"][0].cpu().numpy()
    Yd, status = umap_.umap_fit_batch(Xd, n_neighbors=10, n_components=3, metric="euclidean", random_state=42, n_epochs=0, defer_component_check=True)
    assert int(status.max()) == 0, "the device path must lay out 9 components itself"
    init_dev = Yd[0].cpu().numpy()
    orc = uo.UMAPOracle(n_neighbors=10, n_components=3, metric="euclidean", random_state=42, n_epochs=11).fit(X)
    e = orc._init_embedding
    init_or = 10.0 * (e - e.min(0)) / (e.max(0) - e.min(0))

    def cents(Y):
        return np.stack([Y[lab == c].mean(0) for c in range(9)])
    ch, cd, co = cents(init_host), cents(init_dev), cents(init_or)
    print("component centroids host / device / oracle:\n", np.round(ch, 2), "\n", np.round(cd, 2), "\n", np.round(co, 2))
    assert np.abs(cd - ch).max() < 0.25, "device and host component_layout differ"
    assert np.abs(cd - co).max() < 1.0    # the oracle runs sklearn's SpectralEmbedding (ARPACK): same vectors up to solver tolerance
    # cosine metric, batch of two (a connected cloud and one with nine pieces): statuses and finiteness
    Xb = torch.from_numpy(np.stack([np.random.default_rng(2).normal(0, 1, X.shape).astype(np.float32), X + 0.5])).cuda()
    Yb, st = umap_.umap_fit_batch(Xb, n_neighbors=10, n_components=3, metric="cosine", random_state=42, defer_component_check=True)
    assert st.cpu().tolist() == [0, 0] and bool(torch.isfinite(Yb).all())
