"""Pairwise distances on the tensor cores: host side of kernel (1) (sklearn.pairwise_distances as reached from
umap-learn / ripser.py / metrics.py -- see include/tda_b200.h)."""
from . import _lib

METRICS = {"sqeuclidean": 0, "euclidean": 1, "cosine": 2, "dot": 3}


def pdist(X, Y=None, metric="euclidean", disconnect=float("inf"), out=None):
    """X [B,n,d] (and optionally Y [B,m,d]) float32 CUDA tensors -> [B,n,m] float32 distance matrices."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    if metric not in METRICS:
        raise ValueError(f"tda_multimodal_b200.pdist: unsupported metric {metric!r} (supported: {sorted(METRICS)})")
    squeeze = X.dim() == 2
    if squeeze:
        X = X[None]
        Y = Y[None] if Y is not None else None
    assert X.is_cuda and X.dtype == torch.float32
    X = X.contiguous()
    B, n, d = X.shape
    m = n
    if Y is not None:
        Y = Y.contiguous()
        assert Y.shape[0] == B and Y.shape[2] == d and Y.dtype == torch.float32
        m = Y.shape[1]
    with torch.cuda.device(X.device):
        ws_bytes = int(L.tda_pdist_workspace_bytes(n, m, d, B, 1 if Y is None else 0))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=X.device)
        D = out if out is not None else torch.empty((B, n, m), dtype=torch.float32, device=X.device)
        _lib.check(L.tda_pdist(_lib.ptr(X), _lib.ptr(Y), n, m, d, B, METRICS[metric], float(disconnect), _lib.ptr(D), _lib.ptr(ws),
                               ws_bytes, _lib.stream_ptr()))
    return D[0] if squeeze else D
