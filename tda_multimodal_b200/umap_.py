"""UMAP on the GPU: host side of ``umap.UMAP`` as the reference uses it
(``umap.UMAP(n_neighbors=6, n_components=3, min_dist=0.1, random_state=42, metric='cosine').fit_transform(X)``,
debug_tda_pipeline.py:96-104; ``.fit`` / ``.transform``, analyze_tda_over_layers.py:38-44,69,72;
analyze_adversarial_tda.py:85-93).  Stage order and defaults follow umap-learn 0.5.x;
all arithmetic runs in libtda_b200.so.  ``umap_fit_batch`` is the batched device entry (one call for many
layers); ``UMAP`` mirrors the estimator's constructor, methods and fitted attributes for one cloud.
"""
import warnings

import numpy as np

from . import _lib
from .pdist import METRICS, pdist

try:  # sklearn is present in the image; keep get_params/set_params/clone working like umap-learn's estimator
    from sklearn.base import BaseEstimator
except Exception:  # pragma: no cover
    BaseEstimator = object

DISCONNECTION_DISTANCES = {"correlation": 2.0, "cosine": 2.0, "hellinger": 1.0, "jaccard": 1.0, "dice": 1.0}
_AB_CACHE = {}
DEVICE_MAX_COMPONENTS = 32   # components per cloud tda_spectral_init lays out on the device (kMetaMaxComp in csrc/spectral.cu)


def find_ab_params(spread, min_dist):
    """umap-learn's find_ab_params: least-squares fit of 1/(1+a x^(2b)) to the min_dist/spread target curve."""
    key = (float(spread), float(min_dist))
    if key not in _AB_CACHE:
        from scipy.optimize import curve_fit

        def curve(x, a, b):
            return 1.0 / (1.0 + a * x ** (2 * b))

        xv = np.linspace(0, spread * 3, 300)
        yv = np.zeros(xv.shape)
        yv[xv < min_dist] = 1.0
        yv[xv >= min_dist] = np.exp(-(xv[xv >= min_dist] - min_dist) / spread)
        params, _ = curve_fit(curve, xv, yv)
        _AB_CACHE[key] = (float(params[0]), float(params[1]))
    return _AB_CACHE[key]


def _seed_from(random_state):
    if random_state is None:
        return int(np.random.SeedSequence().generate_state(1)[0])
    if isinstance(random_state, (int, np.integer)):
        return int(random_state) & 0xFFFFFFFF
    if isinstance(random_state, np.random.RandomState):
        return int(random_state.randint(0, 2 ** 31 - 1))
    raise ValueError("random_state must be None, an int or a numpy RandomState")


def distance_matrix(X, Y=None, metric="euclidean", disconnect=None):
    """[B,n,d] (x [B,m,d]) float32 CUDA -> [B,n,m] distances, umap-learn small-data semantics."""
    torch = _lib.require_cuda()
    if disconnect is None:
        disconnect = DISCONNECTION_DISTANCES.get(metric, float("inf"))
    if metric == "euclidean" and X.shape[-1] <= 16 and Y is None:
        from .rips import pdist_lowdim
        return pdist_lowdim(X.contiguous())
    if metric not in ("euclidean", "sqeuclidean", "cosine"):
        raise NotImplementedError(f"tda_multimodal_b200.UMAP: metric={metric!r} is not implemented (cosine, euclidean, sqeuclidean are)")
    return pdist(X, Y, metric=metric, disconnect=disconnect)


def knn_smooth(D, k, local_connectivity=1.0, bandwidth=1.0, n_iter=64):
    torch = _lib.require_cuda()
    L = _lib.lib()
    B, n, m = D.shape
    dev = D.device
    idx = torch.empty((B, n, k), dtype=torch.int32, device=dev)
    dist = torch.empty((B, n, k), dtype=torch.float32, device=dev)
    sigma = torch.empty((B, n), dtype=torch.float32, device=dev)
    rho = torch.empty((B, n), dtype=torch.float32, device=dev)
    ws = torch.empty(8 * B, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.tda_knn_smooth(_lib.ptr(D), n, m, B, int(k), float(local_connectivity), float(bandwidth), int(n_iter), _lib.ptr(idx),
                                    _lib.ptr(dist), _lib.ptr(sigma), _lib.ptr(rho), _lib.ptr(ws), 8 * B, _lib.stream_ptr()))
    return idx, dist, sigma, rho


def fuzzy_graph(idx, dist, sigma, rho, n_epochs, set_op_mix_ratio=1.0):
    torch = _lib.require_cuda()
    L = _lib.lib()
    B, n, k = idx.shape
    dev = idx.device
    slots = 2 * n * k
    head = torch.empty((B, slots), dtype=torch.int32, device=dev)
    tail = torch.empty((B, slots), dtype=torch.int32, device=dev)
    weight = torch.empty((B, slots), dtype=torch.float32, device=dev)
    eps = torch.empty((B, slots), dtype=torch.float32, device=dev)
    maxw = torch.empty((B,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.tda_fuzzy_graph(_lib.ptr(idx), _lib.ptr(dist), _lib.ptr(sigma), _lib.ptr(rho), n, k, B, float(set_op_mix_ratio),
                                     int(n_epochs), _lib.ptr(head), _lib.ptr(tail), _lib.ptr(weight), _lib.ptr(eps), _lib.ptr(maxw),
                                     _lib.stream_ptr()))
    return head, tail, weight, eps


def _component_meta_layout(Xp, comp, ncomp, dim, metric, torch):
    """meta-embedding of component centroids when there are more than 2*dim components (umap-learn component_layout:
    spectral embedding of exp(-d^2) between centroids); tiny (ncomp x ncomp), done with device-side torch ops."""
    d = Xp.shape[1]
    onehot = (comp.long()[None, :] == torch.arange(ncomp, device=Xp.device)[:, None]).to(torch.float64)   # (no index_add_: its float
    cent = (onehot @ Xp.double()) / onehot.sum(1).clamp_min(1.0)[:, None]                                # atomics are not reproducible)
    if metric == "cosine":
        cn = torch.nn.functional.normalize(cent, dim=1)
        dm = (1.0 - cn @ cn.T).clamp(0, 2)
    else:
        dm = torch.cdist(cent, cent)
        if metric == "sqeuclidean":
            dm = dm * dm
    aff = torch.exp(-(dm ** 2))
    aff.fill_diagonal_(0.0)
    deg = aff.sum(1).clamp_min(1e-300)
    isd = deg.rsqrt()
    lap = torch.eye(ncomp, dtype=torch.float64, device=Xp.device) - isd[:, None] * aff * isd[None, :]
    vals, vecs = torch.linalg.eigh(lap)
    emb = vecs[:, 1:dim + 1] * isd[:, None]
    big = emb.abs().argmax(dim=0)                       # sklearn _deterministic_vector_sign_flip: largest entry of every vector positive
    emb = emb * torch.sign(emb[big, torch.arange(emb.shape[1], device=emb.device)])[None, :]
    if emb.shape[1] < dim:
        emb = torch.cat([emb, torch.zeros((ncomp, dim - emb.shape[1]), dtype=emb.dtype, device=emb.device)], 1)
    emb = emb / emb.max().clamp_min(1e-300)             # umap-learn: component_embedding /= component_embedding.max()
    return emb.to(torch.float32)


def spectral_init(X, head, tail, weight, eps, n, dim, seed, metric, speculate_connected=False):
    """spectral_layout / multi_component_layout.  Returns Y [B,n,dim] (not yet noisy-scaled).
    speculate_connected=True: no host synchronisation -- tda_spectral_init lays out every cloud with up to DEVICE_MAX_COMPONENTS
    components on the device (component_layout of the centroids included) and returns (Y, status); status [B] (device) is 1 for a
    cloud with more components: the caller checks it when it synchronises anyway and repeats the fit of such a cloud through the
    path below."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    B, slots = head.shape
    dev = head.device
    if speculate_connected:   # everything on the device; `status` (1 = that cloud needs the host path below) is checked by the caller later
        ncomp = torch.empty((B,), dtype=torch.int32, device=dev)
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        Y = torch.empty((B, n, dim), dtype=torch.float32, device=dev)
        maxcomp = DEVICE_MAX_COMPONENTS
        have_x = X is not None and metric in METRICS and METRICS[metric] <= 2
        d = int(X.shape[-1]) if have_x else 0
        with torch.cuda.device(dev):
            ws_bytes = int(L.tda_spectral_init_workspace_bytes(n, B, maxcomp, slots, d))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(L.tda_spectral_init(_lib.ptr(head), _lib.ptr(tail), _lib.ptr(weight), _lib.ptr(eps), slots, n, dim, B, maxcomp, int(seed),
                                           _lib.ptr(X) if have_x else None, d, METRICS[metric] if have_x else 0,
                                           _lib.ptr(Y), _lib.ptr(ncomp), _lib.ptr(status), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
        return Y, status
    comp = torch.empty((B, n), dtype=torch.int32, device=dev)
    ncomp = torch.empty((B,), dtype=torch.int32, device=dev)
    csize = torch.empty((B, n), dtype=torch.int32, device=dev)
    deg = torch.empty((B, n), dtype=torch.float32, device=dev)
    ws0 = torch.empty(12 * B * n, dtype=torch.uint8, device=dev)
    Y = torch.zeros((B, n, dim), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.tda_graph_components(_lib.ptr(head), _lib.ptr(tail), _lib.ptr(weight), _lib.ptr(eps), slots, n, B, _lib.ptr(comp),
                                          _lib.ptr(ncomp), _lib.ptr(csize), _lib.ptr(deg), _lib.ptr(ws0), 12 * B * n, _lib.stream_ptr()))
        ncomp_h = ncomp.cpu().numpy()
        maxcomp = int(ncomp_h.max())
        min_size = 1 if maxcomp == 1 else max(2 * dim, dim + 2)
        ws_bytes = int(L.tda_spectral_workspace_bytes(n, B, maxcomp, slots))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(L.tda_spectral_embed(_lib.ptr(head), _lib.ptr(tail), _lib.ptr(weight), _lib.ptr(eps), slots, n, dim, B, _lib.ptr(comp),
                                        _lib.ptr(ncomp), _lib.ptr(csize), _lib.ptr(deg), maxcomp, min_size, int(seed), _lib.ptr(Y), None,
                                        _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
    if maxcomp == 1:
        return Y
    # multi-component layout (umap-learn multi_component_layout): components around meta positions
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) + 7919)
    csize_h = csize.cpu().numpy()
    for p in range(B):
        nc = int(ncomp_h[p])
        if nc == 1:
            continue
        if nc > 2 * dim:
            meta = _component_meta_layout(X[p], comp[p], nc, dim, metric, torch)
        else:
            k = int(np.ceil(nc / 2.0))
            base = np.hstack([np.eye(k), np.zeros((k, dim - k))])
            meta = torch.from_numpy(np.vstack([base, -base])[:nc].astype(np.float32)).to(dev)
        md = torch.cdist(meta, meta)
        md[md <= 0] = float("inf")
        data_range = (md.min(dim=1).values / 2.0)
        data_range[~torch.isfinite(data_range)] = 1.0
        cp = comp[p].long()
        small = torch.from_numpy(csize_h[p, :nc] < min_size).to(dev)
        rng_pts = (torch.rand((n, dim), generator=gen, device=dev) * 2 - 1) * data_range[cp][:, None]
        # per-component max |coordinate| of the spectral part
        amax = torch.zeros(nc, dtype=torch.float32, device=dev).scatter_reduce_(0, cp, Y[p].abs().max(dim=1).values, reduce="amax")
        scaled = Y[p] * (data_range / amax.clamp_min(1e-30))[cp][:, None]
        Y[p] = torch.where(small[cp][:, None], rng_pts, scaled) + meta[cp]
    return Y


def umap_fit_batch(X, n_neighbors=15, n_components=2, metric="euclidean", n_epochs=None, learning_rate=1.0, init="spectral",
                   min_dist=0.1, spread=1.0, set_op_mix_ratio=1.0, local_connectivity=1.0, repulsion_strength=1.0,
                   negative_sample_rate=5, random_state=None, a=None, b=None, return_state=False, knn=None, defer_component_check=False,
                   host_spectral_init=False):
    """fit_transform of B clouds at once.  X [B,n,d] float32 CUDA tensor -> embedding [B,n,n_components] (CUDA).
    `knn` = (idx [B,n,k] int32, dist, sigma, rho) skips the distance / kNN stages (e.g. the row-sharded exact kNN of
    pipeline.knn_row_sharded for clouds whose distance matrix should not be materialised on one GPU).
    defer_component_check=True (spectral init only): the call never synchronises with the device; it returns (Y, status) and the
    caller must repeat the fit with the default setting if status.max() > 0 (see spectral_init).
    host_spectral_init=True: components, meta layout and placement through the host-side path (one synchronisation)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    assert X.is_cuda and X.dim() == 3
    X = X.to(torch.float32).contiguous()
    B, n, d = X.shape
    dev = X.device
    k = int(n_neighbors)
    if k >= n:
        warnings.warn("n_neighbors is larger than the dataset size; truncating to X.shape[0] - 1")
        k = n - 1
    if a is None or b is None:
        a, b = find_ab_params(spread, min_dist)
    seed = _seed_from(random_state)
    n_ep = int(n_epochs) if n_epochs is not None else (500 if n <= 10000 else 200)
    if knn is None:
        D = distance_matrix(X, metric=metric)
        idx, dist, sigma, rho = knn_smooth(D, k, local_connectivity=local_connectivity)
        del D
    else:
        idx, dist, sigma, rho = (t.contiguous() for t in knn)
        assert idx.shape == (B, n, k), "knn must hold n_neighbors entries per point (self included)"
    # umap-learn prunes with max(n_epochs, default) > 10 semantics: weights below max/n_epochs are dropped
    head, tail, weight, eps = fuzzy_graph(idx, dist, sigma, rho, n_ep if n_ep > 10 else (500 if n <= 10000 else 200), set_op_mix_ratio)
    with torch.cuda.device(dev):
        if isinstance(init, str) and init == "spectral":
            # one path for every caller: tda_spectral_init lays out clouds with up to DEVICE_MAX_COMPONENTS components on the device
            # (bit-reproducible: fixed summation orders).  Clouds with more components (status 1) go through the host path --
            # here, unless the caller asked to defer that check to its own synchronisation point.
            if host_spectral_init:      # (tests: the host-side multi_component_layout against the device kernels)
                Y, ncomp_dev = spectral_init(X, head, tail, weight, eps, n, n_components, seed, metric), None
            else:
                Y, ncomp_dev = spectral_init(X, head, tail, weight, eps, n, n_components, seed, metric, speculate_connected=True)
            if not defer_component_check and not host_spectral_init:
                bad = torch.nonzero(ncomp_dev > 0).flatten()
                if bad.numel():
                    Y[bad] = spectral_init(X[bad].contiguous(), head[bad].contiguous(), tail[bad].contiguous(), weight[bad].contiguous(),
                                           eps[bad].contiguous(), n, n_components, seed, metric)
            _lib.check(L.tda_umap_rescale(_lib.ptr(Y), n, n_components, B, 1e-4, seed + 1, _lib.stream_ptr()))
        elif isinstance(init, str) and init == "random":
            Y = torch.empty((B, n, n_components), dtype=torch.float32, device=dev)
            _lib.check(L.tda_umap_init_random(_lib.ptr(Y), n, n_components, B, -10.0, 10.0, seed, _lib.stream_ptr()))
            _lib.check(L.tda_umap_rescale(_lib.ptr(Y), n, n_components, B, 0.0, seed + 1, _lib.stream_ptr()))
        else:
            Y0 = torch.as_tensor(np.asarray(init, dtype=np.float32) if not isinstance(init, torch.Tensor) else init, device=dev)
            Y = Y0.to(torch.float32).reshape(B, n, n_components).contiguous().clone()
            _lib.check(L.tda_umap_rescale(_lib.ptr(Y), n, n_components, B, 0.0, seed + 1, _lib.stream_ptr()))
        init_embedding = Y.clone() if return_state else None
        ws_sgd = torch.empty(int(L.tda_umap_sgd_workspace_bytes(head.shape[1], n, n, n_components, B, 1)), dtype=torch.uint8, device=dev)
        _lib.check(L.tda_umap_sgd(_lib.ptr(Y), None, _lib.ptr(head), _lib.ptr(tail), _lib.ptr(eps), head.shape[1], n, n, n_components, B,
                                  n_ep, float(a), float(b), float(repulsion_strength), float(learning_rate), float(negative_sample_rate), 1,
                                  seed + 2, _lib.ptr(ws_sgd), ws_sgd.numel(), _lib.stream_ptr()))
    if return_state:
        return Y, {"knn_indices": idx, "knn_dists": dist, "sigmas": sigma, "rhos": rho, "head": head, "tail": tail, "weight": weight,
                   "eps": eps, "a": a, "b": b, "n_neighbors": k, "n_epochs": n_ep, "init": init_embedding, "seed": seed}
    if defer_component_check:
        return Y, (ncomp_dev if (isinstance(init, str) and init == "spectral") else None)
    return Y


def umap_transform_batch(Xq, Xtrain, train_embedding, n_neighbors, metric="euclidean", n_epochs=None, learning_rate=1.0,
                         local_connectivity=1.0, repulsion_strength=1.0, negative_sample_rate=5, a=None, b=None, seed=0):
    """transform() of B query sets against their fitted clouds: Xq [B,m,d], Xtrain [B,n,d], train_embedding [B,n,dim]."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    B, m, d = Xq.shape
    n = Xtrain.shape[1]
    dim = train_embedding.shape[2]
    dev = Xq.device
    k = int(n_neighbors)
    D = distance_matrix(Xq.contiguous(), Xtrain.contiguous(), metric=metric, disconnect=float("inf"))
    idx, dist, sigma, rho = knn_smooth(D, k, local_connectivity=max(0.0, local_connectivity - 1.0))
    del D
    if n_epochs is None:
        n_ep = 100 if m <= 10000 else 30
    else:
        n_ep = int(n_epochs // 3.0)
    slots = m * k
    head = torch.empty((B, slots), dtype=torch.int32, device=dev)
    tail = torch.empty((B, slots), dtype=torch.int32, device=dev)
    weight = torch.empty((B, slots), dtype=torch.float32, device=dev)
    eps = torch.empty((B, slots), dtype=torch.float32, device=dev)
    maxw = torch.empty((B,), dtype=torch.float32, device=dev)
    Y = torch.empty((B, m, dim), dtype=torch.float32, device=dev)
    te = train_embedding.to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        _lib.check(L.tda_umap_transform_init(_lib.ptr(idx), _lib.ptr(dist), _lib.ptr(sigma), _lib.ptr(rho), _lib.ptr(te), m, n, k, dim, B,
                                             n_ep, _lib.ptr(Y), _lib.ptr(head), _lib.ptr(tail), _lib.ptr(weight), _lib.ptr(eps),
                                             _lib.ptr(maxw), _lib.stream_ptr()))
        ws_sgd = torch.empty(int(L.tda_umap_sgd_workspace_bytes(slots, m, n, dim, B, 0)), dtype=torch.uint8, device=dev)
        _lib.check(L.tda_umap_sgd(_lib.ptr(Y), _lib.ptr(te), _lib.ptr(head), _lib.ptr(tail), _lib.ptr(eps), slots, m, n, dim, B, n_ep,
                                  float(a), float(b), float(repulsion_strength), float(learning_rate) / 4.0, float(negative_sample_rate), 0,
                                  int(seed), _lib.ptr(ws_sgd), ws_sgd.numel(), _lib.stream_ptr()))
    return Y


class UMAP(BaseEstimator):
    """Drop-in for ``umap.UMAP`` for the arguments the reference scripts pass; unsupported umap-learn options raise."""

    def __init__(self, n_neighbors=15, n_components=2, metric="euclidean", metric_kwds=None, output_metric="euclidean",
                 output_metric_kwds=None, n_epochs=None, learning_rate=1.0, init="spectral", min_dist=0.1, spread=1.0,
                 low_memory=True, n_jobs=-1, set_op_mix_ratio=1.0, local_connectivity=1.0, repulsion_strength=1.0,
                 negative_sample_rate=5, transform_queue_size=4.0, a=None, b=None, random_state=None, angular_rp_forest=False,
                 target_n_neighbors=-1, target_metric="categorical", target_metric_kwds=None, target_weight=0.5,
                 transform_seed=42, transform_mode="embedding", force_approximation_algorithm=False, verbose=False,
                 tqdm_kwds=None, unique=False, densmap=False, dens_lambda=2.0, dens_frac=0.3, dens_var_shift=0.1,
                 output_dens=False, disconnection_distance=None, precomputed_knn=(None, None, None)):
        self.n_neighbors = n_neighbors
        self.n_components = n_components
        self.metric = metric
        self.metric_kwds = metric_kwds
        self.output_metric = output_metric
        self.output_metric_kwds = output_metric_kwds
        self.n_epochs = n_epochs
        self.learning_rate = learning_rate
        self.init = init
        self.min_dist = min_dist
        self.spread = spread
        self.low_memory = low_memory
        self.n_jobs = n_jobs
        self.set_op_mix_ratio = set_op_mix_ratio
        self.local_connectivity = local_connectivity
        self.repulsion_strength = repulsion_strength
        self.negative_sample_rate = negative_sample_rate
        self.transform_queue_size = transform_queue_size
        self.a = a
        self.b = b
        self.random_state = random_state
        self.angular_rp_forest = angular_rp_forest
        self.target_n_neighbors = target_n_neighbors
        self.target_metric = target_metric
        self.target_metric_kwds = target_metric_kwds
        self.target_weight = target_weight
        self.transform_seed = transform_seed
        self.transform_mode = transform_mode
        self.force_approximation_algorithm = force_approximation_algorithm
        self.verbose = verbose
        self.tqdm_kwds = tqdm_kwds
        self.unique = unique
        self.densmap = densmap
        self.dens_lambda = dens_lambda
        self.dens_frac = dens_frac
        self.dens_var_shift = dens_var_shift
        self.output_dens = output_dens
        self.disconnection_distance = disconnection_distance
        self.precomputed_knn = precomputed_knn

    # -- umap-learn's _validate_parameters, for the options this implementation honours
    def _validate_parameters(self):
        if self.set_op_mix_ratio < 0.0 or self.set_op_mix_ratio > 1.0:
            raise ValueError("set_op_mix_ratio must be between 0.0 and 1.0")
        if self.repulsion_strength < 0.0:
            raise ValueError("repulsion_strength cannot be negative")
        if self.min_dist > self.spread:
            raise ValueError("min_dist must be less than or equal to spread")
        if self.min_dist < 0.0:
            raise ValueError("min_dist cannot be negative")
        if not isinstance(self.init, str) and not hasattr(self.init, "shape"):
            raise ValueError("init must be a string or ndarray")
        if isinstance(self.init, str) and self.init not in ("spectral", "random"):
            raise ValueError('string init values must be one of: "spectral", "random" (pca/tswspectral are not implemented)')
        if self.negative_sample_rate < 0:
            raise ValueError("negative sample rate must be positive")
        if self.learning_rate < 0.0:
            raise ValueError("learning_rate must be positive")
        if self.n_neighbors < 2:
            raise ValueError("n_neighbors must be greater than 1")
        if not isinstance(self.n_components, (int, np.integer)) or self.n_components < 1:
            raise ValueError("n_components must be an int greater than 0")
        if self.n_components > 4:
            raise NotImplementedError("tda_multimodal_b200.UMAP: n_components > 4 is not implemented")
        if self.n_epochs is not None and (not isinstance(self.n_epochs, (int, np.integer)) or self.n_epochs < 0):
            raise ValueError("n_epochs must be a nonnegative integer")
        for name, default in (("densmap", False), ("output_dens", False), ("unique", False), ("output_metric", "euclidean"),
                              ("transform_mode", "embedding")):
            if getattr(self, name) != default:
                raise NotImplementedError(f"tda_multimodal_b200.UMAP: {name}={getattr(self, name)!r} is not implemented")
        if self.metric_kwds:
            raise NotImplementedError("tda_multimodal_b200.UMAP: metric_kwds is not implemented")

    def fit(self, X, y=None, force_all_finite=True):
        torch = _lib.require_cuda()
        if y is not None:
            raise NotImplementedError("tda_multimodal_b200.UMAP: supervised fitting (y) is not implemented")
        self._validate_parameters()
        is_tensor = isinstance(X, torch.Tensor)
        Xh = X if is_tensor else np.ascontiguousarray(np.asarray(X), dtype=np.float32)  # check_array(dtype=float32, order='C')
        if Xh.ndim != 2:
            raise ValueError("Expected 2D array")
        if not is_tensor and force_all_finite and not np.isfinite(Xh).all():
            raise ValueError("Input contains NaN or infinity")
        Xd = (Xh if is_tensor else torch.from_numpy(Xh)).to(device="cuda", dtype=torch.float32)[None]
        n = Xd.shape[1]
        if n <= 1:
            self.embedding_ = np.zeros((n, self.n_components), dtype=np.float32)
            return self
        self._raw_data = Xd
        self._a, self._b = (self.a, self.b) if (self.a is not None and self.b is not None) else find_ab_params(self.spread, self.min_dist)
        Y, st = umap_fit_batch(Xd, n_neighbors=self.n_neighbors, n_components=self.n_components, metric=self.metric,
                               n_epochs=self.n_epochs, learning_rate=self.learning_rate, init=self.init, min_dist=self.min_dist,
                               spread=self.spread, set_op_mix_ratio=self.set_op_mix_ratio, local_connectivity=self.local_connectivity,
                               repulsion_strength=self.repulsion_strength, negative_sample_rate=self.negative_sample_rate,
                               random_state=self.random_state, a=self._a, b=self._b, return_state=True)
        self._state = st
        self._n_neighbors = st["n_neighbors"]
        self._embedding_dev = Y
        self.embedding_ = Y[0].cpu().numpy()
        self._sigmas = st["sigmas"][0].cpu().numpy()
        self._rhos = st["rhos"][0].cpu().numpy()
        self._knn_indices = st["knn_indices"][0].cpu().numpy()
        self._knn_dists = st["knn_dists"][0].cpu().numpy()
        self._graph = None
        self._input_hash = None if is_tensor else hash(Xh.tobytes())
        return self

    @property
    def graph_(self):
        """fuzzy simplicial set as a scipy CSR matrix (built lazily from the device slot table)."""
        if self._graph is None:
            import scipy.sparse
            st = self._state
            w = st["weight"][0].cpu().numpy()
            keep = w > 0
            n = self.embedding_.shape[0]
            self._graph = scipy.sparse.coo_matrix((w[keep], (st["head"][0].cpu().numpy()[keep], st["tail"][0].cpu().numpy()[keep])),
                                                  shape=(n, n)).tocsr()
        return self._graph

    def fit_transform(self, X, y=None, force_all_finite=True):
        self.fit(X, y, force_all_finite)
        return self.embedding_

    def transform(self, X, force_all_finite=True):
        torch = _lib.require_cuda()
        if not hasattr(self, "embedding_"):
            raise ValueError("This UMAP instance is not fitted yet")
        is_tensor = isinstance(X, torch.Tensor)
        Xh = X if is_tensor else np.ascontiguousarray(np.asarray(X), dtype=np.float32)
        if not is_tensor and self._input_hash is not None and hash(Xh.tobytes()) == self._input_hash:
            return self.embedding_  # umap-learn returns the stored embedding for the training data itself
        Xq = (Xh if is_tensor else torch.from_numpy(Xh)).to(device="cuda", dtype=torch.float32)[None]
        Y = umap_transform_batch(Xq, self._raw_data, self._embedding_dev, self._n_neighbors, metric=self.metric, n_epochs=self.n_epochs,
                                 learning_rate=self.learning_rate, local_connectivity=self.local_connectivity,
                                 repulsion_strength=self.repulsion_strength, negative_sample_rate=self.negative_sample_rate,
                                 a=self._a, b=self._b, seed=_seed_from(self.transform_seed))
        return Y[0].cpu().numpy()
