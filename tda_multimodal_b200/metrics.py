"""The pdist / kNN / Gram pieces of the reference's ``metrics.py`` on libtda_b200 kernels.

Mirrors, with the reference's names, arguments and NaN conventions:
  compute_intrinsic_dimensionality  metrics.py:112-206  TwoNN: batched euclidean pdist + diag=inf + 2 nearest (metrics.py:143-151)
                                    -> tda_pdist (tcgen05 GEMM) + tda_knn_smooth (k=3, self at rank 0); the per-batch Python
                                    loop of sort / discard / regression through the origin (metrics.py:160-206) is done for all
                                    batch items at once with device-side torch ops
  compute_fixed_window_id           metrics.py:211-265  window reshape + the function above
  matrix_entropy                    metrics.py:345-399  Gram matrix K = Z Z^T (metrics.py:368) -> tda_pdist(metric='dot');
                                    the eigenvalue solve stays torch.linalg.eigvalsh (a dense LAPACK-style factorisation, out
                                    of the scope of this path)
Inputs are CUDA tensors (CPU tensors are moved); there is no CPU fallback.
"""
from . import _lib
from .pdist import pdist
from .umap_ import knn_smooth


def compute_intrinsic_dimensionality(data, discard_fraction=0.1, eps=1e-10):
    torch = _lib.require_cuda()
    batch_size, n_samples, _ = data.shape
    dev = data.device if data.is_cuda else torch.device("cuda")
    if n_samples <= 5:
        return torch.full((batch_size,), float("nan"), device=dev, dtype=torch.float32)
    X = data.to(device=dev, dtype=torch.float32).contiguous()
    D = pdist(X, metric="euclidean")                       # exact-zero diagonal
    _, kd, _, _ = knn_smooth(D, 3)                         # rank 0 is a zero distance (self, or a duplicate with a smaller index)
    r1, r2 = kd[..., 1], kd[..., 2]
    valid = (r1 > eps) & (r2 > eps)
    mu = torch.where(valid, r2 / r1, torch.full_like(r1, float("inf")))
    # ---- all batch items at once: sort, keep the smallest n_keep finite ratios, regression through the origin
    mu_sorted, _ = torch.sort(mu, dim=1)                   # infinities last
    n_valid = torch.isfinite(mu).sum(dim=1)
    n_keep = torch.clamp((n_valid.to(torch.float64) * (1.0 - discard_fraction)).floor().to(torch.int64), min=5)
    n_keep = torch.minimum(n_keep, n_valid)                # never reach into the infinities
    idx = torch.arange(n_samples, device=dev)[None, :]
    keep = idx < n_keep[:, None]
    f_emp = (idx + 1).to(torch.float32) / float(n_samples)
    x = torch.where(keep, torch.log(torch.where(keep, mu_sorted, torch.ones_like(mu_sorted)) + eps), torch.zeros_like(mu_sorted))
    y = torch.where(keep, -torch.log(1.0 - f_emp + eps).expand_as(x), torch.zeros_like(x))
    cnt = n_keep.to(torch.float32).clamp_min(2)

    def var(v):                                            # unbiased variance over the kept entries (torch.var semantics)
        mean = v.sum(1) / cnt
        return (torch.where(keep, (v - mean[:, None]) ** 2, torch.zeros_like(v))).sum(1) / (cnt - 1)

    num, den = (x * y).sum(1), (x * x).sum(1)
    slope = num / den
    ok = (n_valid >= 5) & (n_keep >= 5) & (var(x) >= eps) & (var(y) >= eps) & (den.abs() >= eps) & torch.isfinite(slope) & (slope > 0.0) & (slope < 1000.0)
    return torch.where(ok, slope, torch.full_like(slope, float("nan")))


def compute_fixed_window_id(activations_batch, n_windows, discard_fraction=0.1):
    torch = _lib.require_cuda()
    batch_size, seq_len, embed_dim = activations_batch.shape
    dev = activations_batch.device if activations_batch.is_cuda else torch.device("cuda")
    nan = lambda: torch.full((batch_size, max(n_windows, 0)), float("nan"), device=dev, dtype=torch.float32)  # noqa: E731
    if n_windows <= 0 or seq_len < n_windows or seq_len < 6:
        return nan()
    window_size = seq_len // n_windows
    if window_size < 6:
        return nan()
    trunc = activations_batch[:, :n_windows * window_size, :]
    windows = trunc.reshape(batch_size, n_windows, window_size, embed_dim).permute(1, 0, 2, 3).reshape(n_windows * batch_size, window_size, embed_dim)
    flat = compute_intrinsic_dimensionality(windows, discard_fraction)
    return flat.view(n_windows, batch_size).permute(1, 0)


def matrix_entropy(matrix, alpha=1.0, eps=1e-10):
    torch = _lib.require_cuda()
    dev = matrix.device if matrix.is_cuda else torch.device("cuda")
    Z = matrix.to(device=dev, dtype=torch.float32)
    lead = Z.shape[:-2]
    K = pdist(Z.reshape((-1,) + Z.shape[-2:]).contiguous(), metric="dot").reshape(lead + (Z.shape[-2], Z.shape[-2]))
    K = 0.5 * (K + K.transpose(-2, -1))
    ev = torch.clamp(torch.linalg.eigvalsh(K), min=0)
    p = ev / (ev.sum(dim=-1) + eps).unsqueeze(-1)
    if abs(alpha - 1.0) < eps:
        return -torch.sum(torch.xlogy(p, p), dim=-1)
    return torch.log(torch.sum(torch.pow(p, alpha), dim=-1)) / (1.0 - alpha)
