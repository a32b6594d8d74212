"""Vietoris-Rips persistence on the GPU: host side of `ripser(X, maxdim)` (reference call sites
debug_tda_pipeline.py:109-110, analyze_tda_over_layers.py:76, analyze_adversarial_tda.py:100-101).

`rips_batch` is the batched device entry (layers x bootstrap resamples in one call); `ripser` mirrors the
keyword surface and the result dict of ripser.py's `ripser` for a single cloud.
"""
import os
import warnings

import numpy as np

from . import _lib


def _next_pow2(x):
    p = 1
    while p < x:
        p <<= 1
    return p


def pdist_lowdim(pts):
    """[B,n,d] float32 cuda tensor -> [B,n,n] float32 euclidean distances (ripser.py front-end definition)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    B, n, d = pts.shape
    dm = torch.empty((B, n, n), dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        _lib.check(L.tda_pdist_lowdim(_lib.ptr(pts), n, d, B, _lib.ptr(dm), _lib.stream_ptr()))
    return dm


# names of the device counters per reducer (include/tda_b200.h: tda_rips_stats)
_STAT_NAMES = {
    "sweep2": ["columns", "apparent", "reduced", "additions", "rows_substituted", "pivots", "windows", "max_v", "cyc_warp_stage", "cyc_commit_loop",
               "cyc_substitute", "cyc_verify", "cyc_events", "cyc_pm_finalise", "edges_via_columns", "heavy_rows", "late_rounds", "late_rows",
               "pm_rows_moved", "dense_columns", "columns_resumed", "columns_to_cluster", "barrier_cycles", "barriers"],
    "sweep": ["columns", "apparent", "reduced", "additions", "rows_streamed", "pivots", "restarts", "max_v", "cyc_filter", "row_groups",
              "cyc_resolve", "cyc_column_add", "cyc_flip_patch", "cyc_dense_final", "edges_via_columns", "heavy_rows"],
    "bitset": ["columns", "apparent", "reduced", "additions", "toggles", "pivots", "slides", "max_v", "cyc_scan", "cyc_owner", "cyc_gen",
               "cyc_column_add", "cyc_slide", "cyc_final", "edges_via_columns", "far_keys"],
}
_STAT_NAMES["verify"] = _STAT_NAMES["sweep"]


class RipsJob:
    """A batch of Rips problems enqueued on the current CUDA stream (tda_rips_launch).  `finish()` synchronises that
    stream, checks the per-problem status and returns the list of result dicts; on a capacity overflow it re-runs the
    batch synchronously with larger buffers."""

    def __init__(self, dm, maxdim, thresh, cap1, pool_bytes, want_simplices, want_stats, subset=None):
        torch = _lib.require_cuda()
        L = _lib.lib()
        self.dm, self.maxdim, self.thresh = dm, maxdim, float(thresh)
        self.cap1, self.pool_bytes = cap1, pool_bytes
        self.want_simplices, self.want_stats = want_simplices, want_stats
        self.subset = subset           # (parent_ends, parent_sdist, parent_dm, n_parent, idx [B,m] int32): clouds = subsets of one parent
        if subset is None:
            B, n, _ = dm.shape
            dev = dm.device
        else:
            B, n = subset[4].shape
            dev = subset[4].device
        self.B, self.n, self.dev = B, n, dev
        self.stream = torch.cuda.current_stream(dev)
        self.ws_bytes = int(L.tda_rips_workspace_bytes(n, B, maxdim, cap1, pool_bytes))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.h0 = torch.empty((B, n, 2), dtype=torch.float32, device=dev)
        self.h0s = torch.empty((B, n, 2), dtype=torch.int64, device=dev) if want_simplices else None
        self.h1 = torch.empty((B, cap1, 2), dtype=torch.float32, device=dev) if maxdim >= 1 else None
        self.h1s = torch.empty((B, cap1, 2), dtype=torch.int64, device=dev) if (want_simplices and maxdim >= 1) else None
        self.counts = torch.zeros((B, 4), dtype=torch.int32, device=dev)
        self.th = torch.empty((B,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if subset is None:
                _lib.check(L.tda_rips_launch(_lib.ptr(dm), n, B, maxdim, self.thresh, _lib.ptr(self.h0), _lib.ptr(self.h0s), _lib.ptr(self.h1),
                                             _lib.ptr(self.h1s), cap1, _lib.ptr(self.counts), _lib.ptr(self.th), _lib.ptr(self.ws),
                                             self.ws_bytes, pool_bytes, _lib.stream_ptr()))
            else:
                pe, ps, pdm, n_parent, idx = subset
                _lib.check(L.tda_rips_subsets_launch(_lib.ptr(pe), _lib.ptr(ps), _lib.ptr(pdm), int(n_parent), _lib.ptr(idx), n, B, maxdim,
                                                     self.thresh, _lib.ptr(self.h0), _lib.ptr(self.h0s), _lib.ptr(self.h1), _lib.ptr(self.h1s),
                                                     cap1, _lib.ptr(self.counts), _lib.ptr(self.th), _lib.ptr(self.ws), self.ws_bytes,
                                                     pool_bytes, _lib.stream_ptr()))
            # the results follow the kernels on the same stream into pinned host buffers (torch's caching host allocator): when the
            # stream has drained they are already on the host, and finish() has no copy left to wait for
            self._host = {k: (t.to("cpu", non_blocking=True) if t is not None else None)
                          for k, t in (("counts", self.counts), ("h0", self.h0), ("h1", self.h1), ("th", self.th), ("h0s", self.h0s), ("h1s", self.h1s))}

    def finish(self):
        torch = _lib.require_cuda()
        L = _lib.lib()
        dm = self.dm
        B, n = self.B, self.n
        self.stream.synchronize()
        host = self._host
        counts_h = host["counts"].numpy()
        if (counts_h[:, 3] != 0).any():
            # some problem overflowed cap1 / the column pool: run the whole batch again, synchronously, with larger buffers
            del self.ws
            with torch.cuda.stream(self.stream):
                if self.subset is not None:
                    if self.cap1 > (1 << 24):
                        raise RuntimeError("tda_multimodal_b200: Rips job keeps overflowing (cap1=%d)" % self.cap1)
                    return RipsJob(None, self.maxdim, self.thresh, 2 * self.cap1, 2 * self.pool_bytes, self.want_simplices, self.want_stats,
                                   subset=self.subset).finish()
                return rips_batch(dm, maxdim=self.maxdim, thresh=self.thresh, cap1=2 * self.cap1, pool_bytes=2 * self.pool_bytes,
                                  want_simplices=self.want_simplices, want_stats=self.want_stats)
        stats = None
        if self.want_stats and self.maxdim >= 1:
            stats = np.zeros((B, _lib.RIPS_STATS), dtype=np.int64)
            with torch.cuda.device(self.dev):
                _lib.check(L.tda_rips_stats(_lib.ptr(self.ws), n, B, self.maxdim, self.cap1, self.pool_bytes, stats.ctypes.data))
        h0_h = host["h0"].numpy()
        h1_h = host["h1"].numpy() if host["h1"] is not None else None
        th_h = host["th"].numpy()
        h0s_h = host["h0s"].numpy() if host["h0s"] is not None else None
        h1s_h = host["h1s"].numpy() if host["h1s"] is not None else None
        out = []
        for p in range(B):
            c0, c1 = int(counts_h[p, 0]), int(counts_h[p, 1])
            dgms = [h0_h[p, :c0].astype(np.float64)]
            if self.maxdim >= 1:
                dgms.append(h1_h[p, :c1].astype(np.float64))
            r = {"dgms": dgms, "num_edges": int(counts_h[p, 2]), "thresh": float(th_h[p])}
            if self.want_simplices:
                r["simplices"] = [h0s_h[p, :c0]] + ([h1s_h[p, :c1]] if self.maxdim >= 1 else [])
            if stats is not None:
                names = _STAT_NAMES["bitset" if n > 8192 and _lib.rips_reducer() in ("sweep", "verify") else _lib.rips_reducer()]
                r["stats"] = dict(zip(names, stats[p].tolist()))
            out.append(r)
        return out


def _free_bytes_estimate(torch, dev):
    """Memory this process can still take: the device's total minus what torch's allocator holds.  Host-side bookkeeping only --
    cudaMemGetInfo is an ioctl into the kernel driver, and on a shared 8-GPU host it stalled the launching thread for 20-90 ms at
    random (measured: profiles/r02_step_variance.txt), right in the middle of a step's kernel enqueue."""
    return max(1 << 30, int(torch.cuda.get_device_properties(dev).total_memory * 0.95) - int(torch.cuda.memory_reserved(dev))
               + int(torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)))


def _default_sizes(torch, dm, cap1, pool_bytes, shape=None, device=None):
    B, n = (dm.shape[0], dm.shape[1]) if shape is None else shape
    dev = dm.device if device is None else device
    cap1 = _next_pow2(cap1 or max(64, 4 * n))
    free_bytes = _free_bytes_estimate(torch, dev)
    if pool_bytes is None:
        E = n * (n - 1) // 2
        if _lib.rips_reducer() == "bitset":
            # half of the pool holds one key window (bitset over the E*n triangle keys, <= 2^32 bits) per resident CTA,
            # the other half the reduction columns of finished columns
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            grid = min(B, 2 * sms)
            window = min(-(-(E * n) // 8), 1 << 29)
            pool_bytes = max(64 << 20, min(2 * grid * window, int(0.35 * free_bytes)))
        else:
            # row-sweep reducer: the pool only stores the reduction columns (edge lists) of finished columns
            pool_bytes = max(64 << 20, min(B * 16 * E * 4, int(0.35 * free_bytes)))
    return cap1, int(pool_bytes), free_bytes


def rips_batch_launch(dm, maxdim=1, thresh=float("inf"), cap1=None, pool_bytes=None, want_simplices=False, want_stats=False):
    """Enqueue the persistence of `B` dense distance matrices `dm` [B,n,n] (float32, CUDA) on the current stream and return a
    RipsJob; nothing is synchronised until `job.finish()`."""
    torch = _lib.require_cuda()
    if maxdim > 1:
        raise NotImplementedError("tda_multimodal_b200: asynchronous Rips jobs cover maxdim <= 1 (use rips_batch for maxdim=2)")
    assert dm.is_cuda and dm.dtype == torch.float32 and dm.dim() == 3 and dm.shape[1] == dm.shape[2]
    dm = dm.contiguous()
    cap1, pool_bytes, _ = _default_sizes(torch, dm, cap1, pool_bytes)
    return RipsJob(dm, maxdim, thresh, cap1, pool_bytes, want_simplices, want_stats)


def rips_sort_edges(dm):
    """The edges of `B` clouds in filtration order -- ALL of them, no threshold (tda_rips_sort_edges): dm [B,n,n] float32 CUDA ->
    (ends [B,E] int32 holding (i << 16 | j), i > j; sdist [B,E] float32), ripser's order (length ascending, edge index descending).
    Input of rips_subsets_launch."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    assert dm.is_cuda and dm.dtype == torch.float32 and dm.dim() == 3 and dm.shape[1] == dm.shape[2]
    dm = dm.contiguous()
    B, n, _ = dm.shape
    E = n * (n - 1) // 2
    ends = torch.empty((B, E), dtype=torch.int32, device=dm.device)
    sdist = torch.empty((B, E), dtype=torch.float32, device=dm.device)
    with torch.cuda.device(dm.device):
        ws_bytes = int(L.tda_rips_sort_edges_workspace_bytes(n, B))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dm.device)
        _lib.check(L.tda_rips_sort_edges(_lib.ptr(dm), n, B, _lib.ptr(ends), _lib.ptr(sdist), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
        ws.record_stream(torch.cuda.current_stream(dm.device))
    return ends, sdist


def rips_subsets_launch(parent_ends, parent_sdist, parent_dm, idx, maxdim=1, thresh=float("inf"), cap1=None, pool_bytes=None,
                        want_simplices=False, want_stats=False):
    """Persistence of `B` SUBSETS of one parent cloud, enqueued on the current stream (tda_rips_subsets_launch): parent_ends /
    parent_sdist [E_parent] = rips_sort_edges(parent_dm[None]) of that cloud, parent_dm [n_parent, n_parent] its distance matrix
    (used for the enclosing radius of each subset), idx [B,m] int32 CUDA: parent indices of every subset, STRICTLY ASCENDING in each
    row.  The result equals rips_batch_launch(pdist_lowdim(points[idx])) -- same diagrams, same simplex indices -- without the
    pairwise distances, the key generation and the radix sort of every subset: a subset keeps the parent's edge order, so its
    ranks are a flag + prefix count over the parent's sorted edge list."""
    torch = _lib.require_cuda()
    if maxdim > 1:
        raise NotImplementedError("tda_multimodal_b200: asynchronous Rips jobs cover maxdim <= 1")
    assert idx.is_cuda and idx.dtype == torch.int32 and idx.dim() == 2
    pe, ps = parent_ends.reshape(-1).contiguous(), parent_sdist.reshape(-1).contiguous()
    if parent_dm is None:      # only read for the enclosing radius: a finite threshold does without it
        if not np.isfinite(thresh):
            raise ValueError("rips_subsets_launch: parent_dm is needed unless thresh is finite (enclosing radius of every subset)")
        n_parent = (1 + int(round((1 + 8 * pe.numel()) ** 0.5))) // 2
        pdm = None
    else:
        n_parent = int(parent_dm.shape[-1])
        pdm = parent_dm.reshape(n_parent, n_parent).contiguous()
    assert pe.numel() == n_parent * (n_parent - 1) // 2 == ps.numel()
    idx = idx.contiguous()
    cap1, pool_bytes, _ = _default_sizes(torch, None, cap1, pool_bytes, shape=tuple(idx.shape), device=idx.device)
    return RipsJob(None, maxdim, thresh, cap1, pool_bytes, want_simplices, want_stats, subset=(pe, ps, pdm, n_parent, idx))


def rips_batch(dm, maxdim=1, thresh=float("inf"), cap1=None, pool_bytes=None, want_simplices=False, want_stats=False):
    """Persistence of `B` dense distance matrices `dm` [B,n,n] (float32, CUDA).

    Returns a list of B dicts: {'dgms': [float64 (n_k,2)]*(maxdim+1), 'num_edges': int, 'thresh': float
    [, 'simplices': [...], 'stats': {...}]}.  Grows cap1 / the column pool and retries when the library
    reports TDA_ERR_CAPACITY.
    """
    torch = _lib.require_cuda()
    L = _lib.lib()
    if maxdim > 2:
        raise NotImplementedError("tda_multimodal_b200: Rips persistence is implemented for maxdim <= 2")
    assert dm.is_cuda and dm.dtype == torch.float32 and dm.dim() == 3 and dm.shape[1] == dm.shape[2]
    dm = dm.contiguous()
    B, n, _ = dm.shape
    dev = dm.device
    if maxdim == 2 and n > 2048:
        raise NotImplementedError("tda_multimodal_b200: maxdim=2 is implemented for clouds of at most 2048 points")
    want_h2 = maxdim == 2
    maxdim = min(maxdim, 1)    # the library's H0/H1 stage; H2 runs on top of its workspace
    cap1, pool_bytes, free_bytes = _default_sizes(torch, dm, cap1, pool_bytes)
    with torch.cuda.device(dev):
        while True:
            ws_bytes = int(L.tda_rips_workspace_bytes(n, B, maxdim, cap1, pool_bytes))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            h0 = torch.empty((B, n, 2), dtype=torch.float32, device=dev)
            h0s = torch.empty((B, n, 2), dtype=torch.int64, device=dev) if want_simplices else None
            h1 = torch.empty((B, cap1, 2), dtype=torch.float32, device=dev) if maxdim >= 1 else None
            h1s = torch.empty((B, cap1, 2), dtype=torch.int64, device=dev) if (want_simplices and maxdim >= 1) else None
            counts = torch.zeros((B, 4), dtype=torch.int32, device=dev)
            th = torch.empty((B,), dtype=torch.float32, device=dev)
            code = L.tda_rips(_lib.ptr(dm), n, B, maxdim, float(thresh), _lib.ptr(h0), _lib.ptr(h0s), _lib.ptr(h1), _lib.ptr(h1s),
                              cap1, _lib.ptr(counts), _lib.ptr(th), _lib.ptr(ws), ws_bytes, pool_bytes, _lib.stream_ptr())
            if code == _lib.TDA_ERR_CAPACITY and pool_bytes < free_bytes // 2:
                cap1 *= 2
                pool_bytes *= 2
                del ws
                continue
            _lib.check(code)
            break
        stats = None
        if want_stats and maxdim >= 1:
            stats = np.zeros((B, _lib.RIPS_STATS), dtype=np.int64)
            _lib.check(L.tda_rips_stats(_lib.ptr(ws), n, B, maxdim, cap1, pool_bytes, stats.ctypes.data))
        h2_h = c2_h = None
        if want_h2:
            E = n * (n - 1) // 2
            cap2 = _next_pow2(max(1024, 32 * n, n * n // 8))   # residual triangle columns: ~7 k at n=500, ~3e5 expected at n=2000
            pool2 = max(64 << 20, min(2 * min(B, 296) * min(-(-(E * n * n) // 8), 1 << 29), int(0.35 * free_bytes)))
            pool2 = int(os.environ.get("TDA_H2_POOL_BYTES", "0")) or pool2
            # tetrahedron keys beyond one 2^32-bit window of the working column (n > ~300) are parked in far buckets, 8 B per key
            far2 = int(os.environ.get("TDA_H2_FAR_BYTES", "-1"))
            if far2 < 0:
                far2 = 0 if E * n * n <= (1 << 32) else min(int(0.4 * free_bytes), min(B, 296) * (64 << 30))
            while True:
                ws2_bytes = int(L.tda_rips_h2_workspace_bytes(n, B, cap2, pool2, far2))
                ws2 = torch.empty(ws2_bytes, dtype=torch.uint8, device=dev)
                h2 = torch.empty((B, cap2, 2), dtype=torch.float32, device=dev)
                c2 = torch.zeros((B, 4), dtype=torch.int32, device=dev)
                code = L.tda_rips_h2(_lib.ptr(ws), n, B, cap1, pool_bytes, _lib.ptr(h2), cap2, _lib.ptr(c2), _lib.ptr(ws2), ws2_bytes, pool2,
                                     far2, _lib.stream_ptr())
                if code == _lib.TDA_ERR_CAPACITY and pool2 < free_bytes // 2:
                    cap2 *= 4
                    pool2 *= 2
                    del ws2
                    continue
                _lib.check(code)
                break
            h2_h, c2_h = h2.cpu().numpy(), c2.cpu().numpy()
    counts_h = counts.cpu().numpy()
    h0_h = h0.cpu().numpy()
    h1_h = h1.cpu().numpy() if h1 is not None else None
    th_h = th.cpu().numpy()
    h0s_h = h0s.cpu().numpy() if h0s is not None else None
    h1s_h = h1s.cpu().numpy() if h1s is not None else None
    out = []
    for p in range(B):
        c0, c1 = int(counts_h[p, 0]), int(counts_h[p, 1])
        dgms = [h0_h[p, :c0].astype(np.float64)]
        if maxdim >= 1:
            dgms.append(h1_h[p, :c1].astype(np.float64))
        if want_h2:
            dgms.append(h2_h[p, :int(c2_h[p, 1])].astype(np.float64))
        r = {"dgms": dgms, "num_edges": int(counts_h[p, 2]), "thresh": float(th_h[p])}
        if want_simplices:
            r["simplices"] = [h0s_h[p, :c0]] + ([h1s_h[p, :c1]] if maxdim >= 1 else [])
        if stats is not None:
            r["stats"] = dict(zip(_STAT_NAMES[_lib.rips_reducer()], stats[p].tolist()))
        out.append(r)
    return out


def _as_cuda_f32(X):
    torch = _lib.require_cuda()
    if isinstance(X, torch.Tensor):
        return X.to(device="cuda", dtype=torch.float32)
    return torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda()


def _greedy_perm(X, n_perm, is_matrix):
    torch = _lib.require_cuda()
    L = _lib.lib()
    X = X.contiguous()
    n, d = X.shape
    idx = torch.empty(n_perm, dtype=torch.int32, device=X.device)
    lambdas = torch.empty(n_perm, dtype=torch.float32, device=X.device)
    ws_bytes = int(L.tda_greedy_perm_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=X.device)
    with torch.cuda.device(X.device):
        _lib.check(L.tda_greedy_perm(_lib.ptr(X), n, d, int(n_perm), int(is_matrix), _lib.ptr(idx), _lib.ptr(lambdas), _lib.ptr(ws), ws_bytes,
                                     _lib.stream_ptr()))
    return idx.long(), lambdas


def greedy_permutation_device(dm, n_perm):
    """Furthest-point sampling with ripser.py's `n_perm` semantics (start at index 0, lowest index on ties) from a distance
    matrix [n,n] on the device: one launch of tda_greedy_perm.  Returns (idx_perm [n_perm] int64, lambdas [n_perm])."""
    return _greedy_perm(dm, n_perm, True)


def greedy_permutation_points(pts, n_perm):
    """The same furthest-point sampling straight from the points (euclidean, d <= 16): no n x n matrix (clouds of 1e5 points)."""
    return _greedy_perm(pts, n_perm, False)


def ripser(X, maxdim=1, thresh=np.inf, coeff=2, distance_matrix=False, do_cocycles=False, metric="euclidean", n_perm=None):
    """Drop-in for ``ripser.ripser`` (ripser.py): same arguments, same result keys; `dgms` are float64
    ``(n_k, 2)`` arrays (debug_tda_pipeline.py:124-127 json-dumps np.max of them, which needs float64)."""
    torch = _lib.require_cuda()
    if coeff != 2:
        raise NotImplementedError("tda_multimodal_b200.ripser: only coeff=2 (the reference's setting) is implemented")
    if do_cocycles:
        raise NotImplementedError("tda_multimodal_b200.ripser: do_cocycles=True is not implemented")
    if hasattr(X, "tocoo"):
        raise NotImplementedError("tda_multimodal_b200.ripser: sparse distance matrices are not implemented")
    is_tensor = isinstance(X, torch.Tensor)
    shape = tuple(X.shape)
    if len(shape) != 2:
        raise ValueError("ripser expects a 2-D array")
    if distance_matrix:
        if shape[0] != shape[1]:
            raise ValueError("Distance matrix is not square")
    else:
        if shape[0] == shape[1]:
            warnings.warn("The input matrix is square, but the distance_matrix flag is off.  Did you mean to indicate that "
                          "this was a distance matrix?")
        elif shape[0] < shape[1]:
            warnings.warn("The input point cloud has more columns than rows; did you mean to transpose?")
    n = shape[0]
    if n_perm is not None:
        if n_perm > n:
            raise ValueError("Number of points in greedy permutation is greater than number of points in the point cloud")
        if n_perm < 0:
            raise ValueError("Should be a strictly positive number of points in the greedy permutation")
    Xd = _as_cuda_f32(X)
    if (n_perm is not None and n_perm < n and not distance_matrix and metric == "euclidean" and shape[1] <= 16 and n > 4096):
        # large cloud + landmarks: sample from the points, build only the landmark matrix (and landmark-to-all distances)
        idx_t, lambdas = greedy_permutation_points(Xd, n_perm)
        dm = pdist_lowdim(Xd[idx_t][None])[0]
        dperm2all = torch.cdist(Xd[idx_t].double(), Xd.double()).to(torch.float32)
        res = rips_batch(dm[None], maxdim=maxdim, thresh=float(thresh))[0]
        return {"dgms": res["dgms"], "cocycles": [[] for _ in range(maxdim + 1)], "num_edges": res["num_edges"],
                "dperm2all": dperm2all if is_tensor else dperm2all.cpu().numpy(), "idx_perm": idx_t.cpu().numpy(),
                "r_cover": float(lambdas[-1])}
    if distance_matrix:
        dm = Xd
    elif metric == "euclidean" and shape[1] <= 64:
        dm = pdist_lowdim(Xd[None])[0]
    else:
        from .pdist import pdist  # tensor-core distance matrix for high-dimensional clouds
        dm = pdist(Xd[None], metric=metric)[0]
    idx_perm = np.arange(n)
    r_cover = 0.0
    dperm2all = dm
    if n_perm is not None and n_perm < n:
        idx_t, lambdas = greedy_permutation_device(dm, n_perm)
        r_cover = float(lambdas[-1])
        dperm2all = dm[idx_t, :]
        dm = dperm2all[:, idx_t].contiguous()
        idx_perm = idx_t.cpu().numpy()
    res = rips_batch(dm[None], maxdim=maxdim, thresh=float(thresh))[0]
    return {
        "dgms": res["dgms"],
        "cocycles": [[] for _ in range(maxdim + 1)],
        "num_edges": res["num_edges"],
        "dperm2all": dperm2all if is_tensor else dperm2all.cpu().numpy(),
        "idx_perm": idx_perm,
        "r_cover": r_cover,
    }
