"""GPU parity: tda_multimodal_b200.metrics vs a plain-torch restatement of the reference's metrics.py functions
(metrics.py:112-206, 211-265, 345-399).  /root/reference is not available on the GPU box, so the reference semantics are
restated here with torch.cdist / topk / matmul exactly as the reference calls them (float32; tolerance 1e-4 relative on the
dimension estimate: it is a ratio of sums of logs of distance ratios)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ref_id(data, discard_fraction=0.1, eps=1e-10):
    import torch
    B, n, _ = data.shape
    out = torch.full((B,), float("nan"))
    if n <= 5:
        return out
    data = data.to(torch.float64)                         # float64 reference of the same formula
    d = torch.cdist(data, data, p=2.0)
    d.diagonal(dim1=-2, dim2=-1).fill_(float("inf"))
    k2, _ = torch.topk(d, k=2, dim=-1, largest=False, sorted=True)
    r1, r2 = k2[..., 0], k2[..., 1]
    mu = torch.where((r1 > eps) & (r2 > eps), r2 / r1, torch.tensor(float("inf"), dtype=torch.float64))
    for b in range(B):
        mv = mu[b][torch.isfinite(mu[b])]
        if len(mv) < 5:
            continue
        ms, _ = torch.sort(mv)
        nk = max(int(len(ms) * (1.0 - discard_fraction)), 5)
        mk = ms[:nk]
        f = torch.arange(1, nk + 1, dtype=torch.float64) / float(n)
        x, y = torch.log(mk + eps), -torch.log(1.0 - f + eps)
        if torch.var(x) < eps or torch.var(y) < eps:
            continue
        s = torch.sum(x * y) / torch.sum(x * x)
        if torch.isfinite(s) and 0 < s < 1000:
            out[b] = s.float()
    return out


def test_intrinsic_dimensionality_matches_reference_formula():
    import torch
    from tda_multimodal_b200 import metrics
    g = torch.Generator().manual_seed(0)
    for (B, n, E, k) in [(3, 200, 64, 5), (2, 500, 512, 12), (4, 40, 16, 3)]:
        z = torch.randn(B, n, k, generator=g)
        data = torch.cat([z, torch.zeros(B, n, E - k)], dim=-1) @ torch.linalg.qr(torch.randn(E, E, generator=g))[0]
        got = metrics.compute_intrinsic_dimensionality(data.cuda()).cpu()
        want = ref_id(data)
        assert got.shape == (B,) and torch.isfinite(got).all()
        assert torch.allclose(got, want, rtol=1e-4), (got, want)
        assert (got > 0.5 * k).all() and (got < 2.0 * k).all()      # the estimate tracks the latent dimension
    assert torch.isnan(metrics.compute_intrinsic_dimensionality(torch.randn(2, 5, 8).cuda())).all()
    w = metrics.compute_fixed_window_id(data.cuda(), 4)
    assert w.shape == (4, 4) and torch.isfinite(w).all()
    assert torch.isnan(metrics.compute_fixed_window_id(data.cuda(), 30)).all()   # 40 // 30 = 1 sample per window < 6


def test_matrix_entropy_matches_torch():
    import torch
    from tda_multimodal_b200 import metrics
    g = torch.Generator().manual_seed(1)
    Z = torch.randn(3, 60, 128, generator=g)

    def ref(Z, alpha):
        K = (Z.double() @ Z.double().transpose(-2, -1))
        ev = torch.clamp(torch.linalg.eigvalsh(K), min=0)
        p = ev / (ev.sum(-1, keepdim=True) + 1e-10)
        if abs(alpha - 1.0) < 1e-10:
            return -torch.sum(torch.xlogy(p, p), dim=-1)
        return torch.log(torch.sum(p ** alpha, dim=-1)) / (1.0 - alpha)

    for alpha in (1.0, 2.0):
        got = metrics.matrix_entropy(Z.cuda(), alpha=alpha).cpu().double()
        assert torch.allclose(got, ref(Z, alpha), rtol=1e-4, atol=1e-5), (alpha, got, ref(Z, alpha))


def test_silhouette_matches_sklearn():
    from sklearn.metrics import silhouette_score as sk_sil
    from tda_multimodal_b200 import pipeline
    rng = np.random.default_rng(5)
    shapes = ["circle", "square", "triangle", "star", "hexagon", "diamond"]
    for n, k in [(36, 6), (500, 6), (301, 17)]:
        centers = rng.normal(0, 3, (k, 3))
        lab_i = rng.integers(0, k, n)
        lab_i[:k] = np.arange(k)
        Y = (centers[lab_i] + rng.normal(0, 1.0, (n, 3))).astype(np.float32)
        labels = [shapes[i] if k == 6 else f"c{i}" for i in lab_i]      # string labels, as in the reference
        got = pipeline.silhouette_score(Y, labels)
        want = sk_sil(Y, labels)
        assert isinstance(got, float) and abs(got - want) <= 1e-5, (n, k, got, want)
    # batched, shared labels; singleton label contributes 0
    Yb = rng.normal(size=(3, 50, 3)).astype(np.float32)
    lb = np.r_[np.zeros(24, int), np.ones(25, int), [2]]
    got = pipeline.silhouette_score(Yb, lb)
    for b in range(3):
        assert abs(got[b] - sk_sil(Yb[b], lb)) <= 1e-5
    with pytest.raises(ValueError):
        pipeline.silhouette_score(Yb[0], np.zeros(50, int))
