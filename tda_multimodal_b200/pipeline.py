"""The per-layer TDA sweep of the reference as ONE batched device pipeline.

Reference: the hot loop ``for i in tqdm(range(N_LAYERS))`` of debug_tda_pipeline.py:92-131 (same body in
analyze_adversarial_tda.py:82-123): per layer  X[N,hidden] -> umap.UMAP(n_neighbors, n_components=3, min_dist=0.1,
random_state=42, metric='cosine').fit_transform -> ripser(Y, maxdim=1)['dgms'] -> get_persistence -> stats record.
The layers are independent (no carried state but the append-only stats list, :131), so here all layers of a
rank go through each kernel together (batch dimension = layer), and ranks split the layers (layer l -> rank
l mod G) with a single gather of the diagrams at the end.
"""
import gc
import os

import numpy as np

from . import _lib
from .rips import pdist_lowdim, rips_batch, rips_batch_launch
from .umap_ import umap_fit_batch


def get_persistence(dgm):
    """debug_tda_pipeline.py:79-89 (verbatim semantics): finite persistence values and their max."""
    if dgm.shape[0] == 0:
        return np.array([]), 0.0
    pers = dgm[:, 1] - dgm[:, 0]
    pers = pers[np.isfinite(pers)]
    if pers.shape[0] == 0:
        return np.array([]), 0.0
    return pers, np.max(pers)


def stats_record(layer, dgms):
    """The H0/H1 fields of one summary_stats.json record, in the reference's key order (debug_tda_pipeline.py:121-131);
    the two silhouette fields are appended by the caller that has the labels."""
    h0_pers, max_h0 = get_persistence(dgms[0])
    h1_pers, max_h1 = get_persistence(dgms[1]) if len(dgms) > 1 else (np.array([]), 0.0)
    return {"layer": int(layer), "n_h1_features": len(h1_pers), "max_h1_persistence": float(max_h1),
            "all_h1_persistence_values": h1_pers.tolist(), "n_h0_features": len(dgms[0]) - len(h0_pers),
            "max_h0_persistence": float(max_h0)}


def layer_sweep(X, n_neighbors=15, n_components=3, min_dist=0.1, metric="cosine", random_state=42, maxdim=1, n_epochs=None,
                return_embedding=True, chunks=None):
    """UMAP + Rips for a stack of layers resident on the device.  X [L,n,d] float32 CUDA tensor.
    Returns {'embedding': [L,n,n_components] CUDA tensor, 'results': [L dicts with 'dgms', 'num_edges', 'thresh']}.

    The layers are cut into `chunks` groups (an int, or a list of group sizes; default: _default_groups -- four groups for
    L >= 24, two for L >= 8; env TDA_SWEEP_CHUNKS overrides)
    that run on their own CUDA streams:
    the Rips reduction of a group (one SM per cloud, latency bound) overlaps the UMAP stages of the next groups, and the
    reductions of all groups overlap each other (tda_rips_launch does not synchronise)."""
    torch = _lib.require_cuda()
    Lc = X.shape[0]
    on_host = not X.is_cuda     # a (pinned) host tensor: every chunk copies its own layers on its own stream, so the copy of
    dev = torch.device("cuda", torch.cuda.current_device()) if on_host else X.device   # one chunk overlaps the compute of another
    if chunks is None:
        chunks = int(os.environ.get("TDA_SWEEP_CHUNKS", "0")) or _default_groups(Lc, on_host)
    if isinstance(chunks, (list, tuple)):      # explicit group sizes (they must add up to L)
        sizes = [int(c) for c in chunks if int(c) > 0]
        assert sum(sizes) == Lc, "chunk sizes must add up to the number of layers"
        chunks = len(sizes)
    else:
        chunks = max(1, min(int(chunks), Lc))
        sizes = [(Lc * (c + 1)) // chunks - (Lc * c) // chunks for c in range(chunks)]
    cur = torch.cuda.current_stream(dev)
    bounds = [0]
    for sz in sizes:
        bounds.append(bounds[-1] + sz)
    streams = _sweep_streams(dev, chunks, staggered=on_host)
    Ys, jobs, checks, Xcs = [], [], [], []
    kw = dict(n_neighbors=n_neighbors, n_components=n_components, metric=metric, min_dist=min_dist, random_state=random_state, n_epochs=n_epochs)
    # host input: the H2D copies of all groups go through ONE copy stream, in group order -- copies issued on several streams share
    # the link and all finish late; in order, group 0 is on the device after 1/chunks of the transfer and its kernels start then
    copy_stream, staged = None, []
    if on_host:
        copy_stream = _copy_stream(dev)
        copy_stream.wait_stream(cur)
        with torch.cuda.stream(copy_stream):
            for c in range(chunks):
                Xc = X[bounds[c]:bounds[c + 1]].to(device=dev, dtype=torch.float32, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                staged.append((Xc, ev))
    # no garbage collection while the groups' kernels are being enqueued: a generation-1/2 pass of CPython's collector takes
    # 25-60 ms in a process with torch loaded -- as long as half a step -- and the device idles behind it.  Once everything is
    # queued a collection costs nothing (the host only waits for the device then).
    gc_was_enabled = gc.isenabled()
    gc.disable()
    try:
        _enqueue_groups(torch, X, on_host, chunks, bounds, streams, staged, cur, kw, maxdim, Ys, jobs, checks, Xcs)
    finally:
        if gc_was_enabled:
            gc.enable()
    res = []
    for c, job in enumerate(jobs):
        r = job.finish()
        if checks[c] is not None and int(checks[c].max().item()) > 0:   # clouds with more than 32 components: the host path, for them only
            bad = torch.nonzero(checks[c] > 0).flatten()
            with torch.cuda.stream(streams[c]):
                Yb = umap_fit_batch(Xcs[c][bad].contiguous(), **kw)
                Ys[c][bad] = Yb
                rb = rips_batch(pdist_lowdim(Yb), maxdim=maxdim)
            for j, q in enumerate(bad.tolist()):
                r[q] = rb[j]
        res += r
    del Xcs, staged
    for st in streams[:chunks]:
        cur.wait_stream(st)
    Yall = None
    if return_embedding:
        Yall = torch.cat(Ys, dim=0)
        for Y, st in zip(Ys, streams):
            Y.record_stream(cur)
    return {"embedding": Yall, "results": res}


def _enqueue_groups(torch, X, on_host, chunks, bounds, streams, staged, cur, kw, maxdim, Ys, jobs, checks, Xcs):
    """Enqueues the whole chain of every group on its stream; nothing here synchronises with the device."""
    for c in range(chunks):
        st = streams[c]
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            if on_host:
                Xc, ev = staged[c]
                st.wait_event(ev)
                Xc.record_stream(st)
            else:
                Xc = X[bounds[c]:bounds[c + 1]]
                Xc.record_stream(st)
            # nothing below synchronises with the device: every chunk's whole chain is enqueued before the first result is awaited
            # (tda_spectral_init handles up to 32 components per cloud on the device; its status is checked after the sweep)
            # the LAST group's chain ends the sweep with nothing left to overlap it: its kernels may take more SMs per cloud
            tail = TAIL_OPTIONS if (c == chunks - 1 and chunks > 1) else {}
            saved = {k_: _lib.get_option(k_) for k_ in tail}
            for k_, v_ in tail.items():
                _lib.set_option(k_, v_)
            try:
                Y, ncomp = umap_fit_batch(Xc, defer_component_check=True, **kw)
                jobs.append(rips_batch_launch(pdist_lowdim(Y), maxdim=maxdim))
            finally:
                for k_, v_ in saved.items():
                    _lib.set_option(k_, v_)
            Ys.append(Y)
            checks.append(ncomp)
            Xcs.append(Xc)


def silhouette_score(Y, labels, dm=None):
    """sklearn.metrics.silhouette_score(Y, labels) (euclidean) as the reference calls it (debug_tda_pipeline.py:117-118) on the
    GPU.  Y [n,dim] or [B,n,dim] (numpy or CUDA tensor), labels: one sequence of n hashable labels (strings in the reference)
    shared by all clouds, or [B,n].  Returns a Python float (single cloud) or a float32 numpy array [B]."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    single = (Y.ndim == 2)
    Yt = Y if isinstance(Y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(Y, dtype=np.float32))
    Yt = Yt.to(device="cuda", dtype=torch.float32)
    if single:
        Yt = Yt[None]
    B, n, _ = Yt.shape
    lab = np.asarray(labels)
    if lab.ndim == 1:
        lab = np.broadcast_to(lab, (B, n))
    uniq, inv = np.unique(lab.reshape(-1), return_inverse=True)
    for b in range(B):   # scikit-learn's check_number_of_labels, per cloud
        nl = len(np.unique(lab[b]))
        if not (2 <= nl <= n - 1):
            raise ValueError("Number of labels is %d. Valid values are 2 to n_samples - 1 (inclusive)" % nl)
    lab_d = torch.from_numpy(inv.reshape(B, n).astype(np.int32)).to(Yt.device)
    if dm is None:
        dm = pdist_lowdim(Yt.contiguous())
    score = torch.empty((B,), dtype=torch.float32, device=Yt.device)
    ws = torch.empty(8 * B, dtype=torch.uint8, device=Yt.device)
    with torch.cuda.device(Yt.device):
        _lib.check(L.tda_silhouette(_lib.ptr(dm), _lib.ptr(lab_d), n, B, int(len(uniq)), _lib.ptr(score), _lib.ptr(ws), 8 * B, _lib.stream_ptr()))
    out = score.cpu().numpy()
    return float(out[0]) if single else out


def knn_row_sharded(X, k, metric="cosine", rank=None, world=None, group=None, row_block=8192, local_connectivity=1.0):
    """Exact kNN (+ sigma/rho) of ONE large cloud X [n,d] (CUDA, replicated on every rank) with the rows of the distance
    matrix sharded over the ranks (config C5 of BASELINE.json): rank r computes rows
    [r*n/G, (r+1)*n/G) x all columns, `row_block` rows at a time (the n x n matrix is never materialised), then the
    [n/G, k] index / distance / sigma / rho blocks are exchanged with one all_gather each -- the only traffic of the path.
    Returns (idx [1,n,k] int32, dist [1,n,k], sigma [1,n], rho [1,n]) on every rank, ready for umap_fit_batch(knn=...).
    (The sigma floor of rows without a positive neighbour distance uses the mean distance of the row block, not of the
    whole matrix; such rows only occur for duplicated points.)"""
    torch = _lib.require_cuda()
    import torch.distributed as dist
    from .pdist import METRICS
    from .umap_ import DISCONNECTION_DISTANCES
    n = X.shape[0]
    if rank is None:
        rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
    if world is None:
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    r0, r1 = (n * rank) // world, (n * (rank + 1)) // world
    disc = float(DISCONNECTION_DISTANCES.get(metric, float("inf")))
    Xc = X.to(torch.float32).contiguous()
    d = Xc.shape[1]
    L = _lib.lib()
    rows = r1 - r0
    idx = torch.empty((rows, k), dtype=torch.int32, device=X.device)
    dst = torch.empty((rows, k), dtype=torch.float32, device=X.device)
    sig = torch.empty((rows,), dtype=torch.float32, device=X.device)
    rho = torch.empty((rows,), dtype=torch.float32, device=X.device)
    if rows > 0:
        with torch.cuda.device(X.device):   # one C call: GEMM row block -> top-k, block after block (tda_knn_fused)
            ws_bytes = int(L.tda_knn_fused_workspace_bytes(n, d, int(row_block)))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=X.device)
            _lib.check(L.tda_knn_fused(_lib.ptr(Xc), n, d, r0, r1, int(k), METRICS[metric], disc, float(local_connectivity), _lib.ptr(idx),
                                       _lib.ptr(dst), _lib.ptr(sig), _lib.ptr(rho), int(row_block), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
    local = [idx, dst, sig, rho]                                                       # rows [r0, r1)
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return tuple(t[None] for t in local)      # no process group: the caller gets this rank's rows only
    # ragged row counts: pad every block to the largest, gather, trim
    per = [(n * (r + 1)) // world - (n * r) // world for r in range(world)]
    mx = max(per)
    out = []
    for t in local:
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        out.append(torch.cat([bufs[r][:per[r]] for r in range(world)], dim=0)[None])
    return tuple(out)


def bootstrap_rips(Y, n_resamples=256, size=1000, seed=4000, layer_ids=None, replace=False, maxdim=1, max_batch=256, subsets=None):
    """Config C4 of BASELINE.json: for every 3-D cloud Y[l] ([L,n,dim] CUDA tensor, e.g. the UMAP output of layer l),
    `n_resamples` bootstrap resamples of `size` points -> Rips H0/H1 of each, batched `max_batch` problems per call (the
    batch dimension of every Rips kernel is the resample).  The reference has no bootstrap code: the semantics are
    L * n_resamples independent ``ripser(Y_l[idx], maxdim=1)`` calls; index sets come from
    workloads.c4_resample_indices (seeded per (layer, resample)).  Returns results[l][r] = dict with 'dgms', ...

    subsets (default: whenever the index sets are strictly ascending, i.e. sampled without replacement): the edges of the layer's
    cloud are sorted ONCE and every resample takes its filtration ranks from that order (rips_subsets_launch) instead of
    computing and sorting its own distance matrix; the results are the same bits as the per-resample path (subsets=False)."""
    torch = _lib.require_cuda()
    from . import workloads
    from .rips import rips_sort_edges, rips_subsets_launch
    Lc, n, dim = Y.shape
    layer_ids = list(range(Lc)) if layer_ids is None else list(layer_ids)
    dev = Y.device
    cur = torch.cuda.current_stream(dev)
    streams = _sweep_streams(dev, 2)
    out = [[] for _ in layer_ids]
    pending = []       # (layer slot, job): at most two batches in flight, on alternating streams -- the host unpacks the
    k = 0              # diagrams of one batch while the device works on the next
    for li, layer in enumerate(layer_ids):
        idx_h = workloads.c4_resample_indices(layer, n, n_resamples, size, seed=seed, replace=replace)
        ascending = size >= 2 and bool((np.diff(idx_h, axis=1) > 0).all())
        use_sub = ascending and maxdim <= 1 and size * (size - 1) >= n if subsets is None else bool(subsets)
        if use_sub and not ascending:
            raise ValueError("bootstrap_rips(subsets=True) needs strictly ascending index sets (sampling without replacement)")
        if use_sub:   # the layer's cloud: distance matrix and ALL its edges in filtration order, once for its resamples
            parent_dm = pdist_lowdim(Y[li:li + 1].contiguous())
            parent_ends, parent_sdist = rips_sort_edges(parent_dm)
        for r0 in range(0, n_resamples, max_batch):
            st = streams[k % 2]
            k += 1
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                if use_sub:
                    idx = torch.from_numpy(idx_h[r0:r0 + max_batch].astype(np.int32)).to(dev, non_blocking=True)
                    pending.append((li, rips_subsets_launch(parent_ends, parent_sdist, parent_dm, idx, maxdim=maxdim)))
                else:
                    idx = torch.from_numpy(np.array(idx_h[r0:r0 + max_batch])).to(dev, non_blocking=True)   # (a copy: the cached sets are read-only)
                    pts = Y[li][idx].contiguous()                      # [b, size, dim]
                    pending.append((li, rips_batch_launch(pdist_lowdim(pts), maxdim=maxdim)))
            if len(pending) == 2:
                lj, job = pending.pop(0)
                out[lj] += job.finish()
    for lj, job in pending:
        out[lj] += job.finish()
    for st in streams[:2]:
        cur.wait_stream(st)
    return out


_STREAMS = {}
_COPY_STREAMS = {}
# library options (tda_set_option) applied to the launches of the LAST group of a sweep only: its Rips reduction ends the sweep with
# nothing left to overlap, so it takes 8 CTAs per cloud instead of 4 (measured on C3: 50.8 -> 48.9 ms per step)
TAIL_OPTIONS = {"rips_cluster": 8}


def _default_groups(L, on_host):
    """Number of (equal) groups of a sweep over L layers: from 24 layers on, three groups for resident input and four for host
    input (the H2D copy of a group hides behind the compute of the previous ones); two from 8 layers on.  Measured on C3
    (profiles/r02_tune_groups.txt): resident 3 groups 45.7 ms, 4 groups 47.1 ms; host input 4 groups 55.1 ms, 3 groups 57+ ms.
    Uneven splits (10/9/8/5 ...) were up to 3 % faster on one set of clouds and 5 % slower, with a 10 % step-to-step spread, on
    another: not used."""
    if L < 8:
        return 1
    if L < 24:
        return 2
    return 4 if on_host else 3


def _copy_stream(device):
    torch = _lib.require_cuda()
    key = (device.type, device.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


def _sweep_streams(device, k, staggered=False):
    """k streams for the groups of a sweep.  staggered=True (host input: group c's data arrives after group c-1's): later groups get
    the higher priority, so the chain that starts last goes first wherever two groups compete for SMs (measured: e2e 62.5 -> 60.3 ms
    per 32-layer step; with the data already resident the same priorities cost 2 ms, so the resident sweep uses equal priorities)."""
    torch = _lib.require_cuda()
    key = (device.type, device.index, bool(staggered))
    have = _STREAMS.setdefault(key, [])
    while len(have) < k:
        have.append(torch.cuda.Stream(device=device, priority=-1 if (staggered and len(have) >= 1) else 0))
    return have


def layer_sweep_host(X_host, device=None, **kw):
    """The call a user of the reference makes, with HOST buffers: X_host [L,n,d] float32/float64 numpy array or a
    (pinned) CPU torch tensor.  Copies to the device, runs layer_sweep, returns host results
    ({'embedding': float32 ndarray [L,n,3], 'results': [...]})."""
    torch = _lib.require_cuda()
    Xt = X_host if isinstance(X_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(X_host))
    if Xt.dtype != torch.float32:
        Xt = Xt.to(torch.float32)     # check_array(dtype=float32) of umap-learn, on the host
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    with torch.cuda.device(dev):
        out = layer_sweep(Xt, **kw)
    out["embedding"] = out["embedding"].cpu().numpy() if out["embedding"] is not None else None
    return out


# ---- diagram exchange between ranks ---------------------------------------------------------------------------
def pack_diagrams(results, maxdim=1):
    """list of per-unit result dicts -> (counts [U, maxdim+1] int32, payload [sum,2] float32), unit-major, dim-minor."""
    counts = np.zeros((len(results), maxdim + 1), dtype=np.int32)
    rows = []
    for u, r in enumerate(results):
        for q in range(maxdim + 1):
            d = np.asarray(r["dgms"][q], dtype=np.float32).reshape(-1, 2)
            counts[u, q] = d.shape[0]
            rows.append(d)
    payload = np.concatenate(rows, axis=0) if rows else np.zeros((0, 2), np.float32)
    return counts, np.ascontiguousarray(payload, dtype=np.float32)


def unpack_diagrams(counts, payload):
    out, off = [], 0
    for u in range(counts.shape[0]):
        dgms = []
        for q in range(counts.shape[1]):
            c = int(counts[u, q])
            dgms.append(payload[off:off + c].astype(np.float64))
            off += c
        out.append(dgms)
    return out


def shard_units(n_units, rank, world):
    """unit u -> rank u mod world."""
    return list(range(rank, n_units, world))


def gather_diagrams(local_units, local_results, n_units, maxdim=1, group=None, device=None):
    """All ranks contribute the diagrams of their units; every rank gets the full list (unit order).  Two
    collectives: all_gather of the per-unit counts, all_gather of the padded [pairs,2] payload.  Works on NCCL
    (device tensors) and on gloo (CPU tensors; used by the CPU tests)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        full = [None] * n_units
        for u, r in zip(local_units, local_results):
            full[u] = r["dgms"]
        return full
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))
    counts, payload = pack_diagrams(local_results, maxdim)
    per_rank = (n_units + world - 1) // world
    cpad = np.zeros((per_rank, maxdim + 1), dtype=np.int32)
    cpad[:counts.shape[0]] = counts
    ct = torch.from_numpy(cpad).to(dev)
    all_counts = [torch.empty_like(ct) for _ in range(world)]
    dist.all_gather(all_counts, ct, group=group)
    all_counts = [c.cpu().numpy() for c in all_counts]
    max_rows = max(int(c.sum()) for c in all_counts)
    ppad = np.zeros((max(max_rows, 1), 2), dtype=np.float32)
    ppad[:payload.shape[0]] = payload
    pt = torch.from_numpy(ppad).to(dev)
    all_payload = [torch.empty_like(pt) for _ in range(world)]
    dist.all_gather(all_payload, pt, group=group)
    full = [None] * n_units
    for r in range(world):
        units = shard_units(n_units, r, world)
        dg = unpack_diagrams(all_counts[r][:len(units)], all_payload[r].cpu().numpy())
        for u, d in zip(units, dg):
            full[u] = d
    return full


# ------------------------------------------------------------------------------------------------------------------
# fit once / transform many (analyze_tda_over_layers.py:67-72): all layers depend on ONE fit, so the fit runs on one rank,
# the fitted state is broadcast, and the transforms shard by layer.
FITTED_STATE_KEYS = ("raw_data", "embedding", "scalars")


def pack_fitted_state(raw_data, embedding, a, b, n_neighbors):
    """The three tensors a rank needs to run UMAP.transform: training data [n,d] f32, its embedding [n,dim] f32, and
    (a, b, n_neighbors, n, d, dim) as one float64 vector."""
    import torch
    raw = raw_data.reshape(raw_data.shape[-2], raw_data.shape[-1]).to(torch.float32).contiguous()
    emb = embedding.reshape(embedding.shape[-2], embedding.shape[-1]).to(torch.float32).contiguous()
    sc = torch.tensor([float(a), float(b), float(n_neighbors), float(raw.shape[0]), float(raw.shape[1]), float(emb.shape[1])], dtype=torch.float64)
    return {"raw_data": raw, "embedding": emb, "scalars": sc}


def broadcast_fitted_state(state, src=0, group=None, device=None):
    """`state` = pack_fitted_state(...) on rank `src` OF THE GROUP, None elsewhere.  Every rank returns the same dict (tensors on
    `device`: the current CUDA device under NCCL, the CPU under gloo).  Two broadcasts: the shape/scalar vector, then data +
    embedding as one flat float32 buffer."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return state
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))
    gsrc = dist.get_global_rank(group, src) if group is not None else src   # dist.broadcast takes the GLOBAL rank of the source
    sc = state["scalars"].to(dev) if rank == src else torch.zeros(6, dtype=torch.float64, device=dev)
    dist.broadcast(sc, src=gsrc, group=group)
    n, d, dim = int(sc[3].item()), int(sc[4].item()), int(sc[5].item())
    flat = torch.empty(n * d + n * dim, dtype=torch.float32, device=dev)
    if rank == src:
        flat[:n * d] = state["raw_data"].to(dev).reshape(-1)
        flat[n * d:] = state["embedding"].to(dev).reshape(-1)
    dist.broadcast(flat, src=gsrc, group=group)
    return {"raw_data": flat[:n * d].reshape(n, d), "embedding": flat[n * d:].reshape(n, dim), "scalars": sc.cpu()}


def fit_once_transform_many(X_layers, fit_layer=-1, n_neighbors=15, n_components=3, min_dist=0.1, metric="cosine", random_state=42,
                            maxdim=1, group=None):
    """analyze_tda_over_layers.py:56-77 over the ranks: UMAP is fitted on layer `fit_layer` (default: the last one, as
    analyze_tda_over_layers.py:68 does) by rank 0, its state is broadcast,
    rank r transforms layers r, r+G, ... (the fit layer returns the stored embedding, as umap-learn does for its own training
    data) and runs Rips on each; the diagrams are gathered.  X_layers [L,n,d] float32 CUDA tensor (replicated, or only this
    rank's layers filled in).  Returns (embeddings of this rank's layers {layer: [n,dim] CUDA tensor}, diagrams of ALL layers)."""
    torch = _lib.require_cuda()
    import torch.distributed as dist
    from .umap_ import UMAP, find_ab_params, umap_transform_batch
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else 0
    world = dist.get_world_size(group) if multi else 1
    L = X_layers.shape[0]
    fit_layer = fit_layer % L
    state = None
    if rank == 0:
        um = UMAP(n_neighbors=n_neighbors, n_components=n_components, min_dist=min_dist, metric=metric, random_state=random_state)
        um.fit(X_layers[fit_layer])
        state = pack_fitted_state(um._raw_data[0], um._embedding_dev[0], um._a, um._b, um._n_neighbors)
    state = broadcast_fitted_state(state, src=0, group=group)
    a, b, k = float(state["scalars"][0]), float(state["scalars"][1]), int(state["scalars"][2])
    mine = shard_units(L, rank, world)
    emb, results = {}, []
    for l in mine:
        if l == fit_layer:
            Y = state["embedding"]
        else:
            Y = umap_transform_batch(X_layers[l][None].to(torch.float32), state["raw_data"][None], state["embedding"][None], k, metric=metric,
                                     a=a, b=b, seed=42)[0]
        emb[l] = Y
        results.append(rips_batch(pdist_lowdim(Y[None].contiguous()), maxdim=maxdim)[0])
    return emb, gather_diagrams(mine, results, L, maxdim=maxdim, group=group)
