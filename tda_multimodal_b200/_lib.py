"""ctypes binding of libtda_b200.so (the C ABI declared in include/tda_b200.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtda_b200.so")

c_void_p, c_int, c_float, c_size_t, c_int64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_int64

# name -> (restype, argtypes); every symbol include/tda_b200.h declares must be listed here
PROTOTYPES = {
    "tda_version": (c_int, []),
    "tda_last_error": (ctypes.c_char_p, []),
    "tda_launch_count": (c_int64, []),
    "tda_launch_count_reset": (None, []),
    "tda_set_option": (c_int, [ctypes.c_char_p, ctypes.c_longlong]),
    "tda_get_option": (ctypes.c_longlong, [ctypes.c_char_p]),
    "tda_stage_timing_enable": (None, [c_int]),
    "tda_stage_timing_reset": (None, []),
    "tda_stage_timing_read": (c_int, [c_void_p, c_void_p, c_int]),
    "tda_stage_timeline_read": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "tda_pdist_lowdim": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tda_pdist_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "tda_pdist": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tda_knn_smooth": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_size_t, c_void_p]),
    "tda_fuzzy_graph": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "tda_umap_sgd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                             c_float, c_float, c_float, c_float, c_int, ctypes.c_uint64, c_void_p, c_size_t, c_void_p]),
    "tda_umap_sgd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "tda_umap_init_random": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, ctypes.c_uint64, c_void_p]),
    "tda_umap_rescale": (c_int, [c_void_p, c_int, c_int, c_int, c_float, ctypes.c_uint64, c_void_p]),
    "tda_umap_transform_init": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tda_knn_fused_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "tda_knn_fused": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                              c_void_p, c_size_t, c_void_p]),
    "tda_spectral_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tda_spectral_init_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "tda_spectral_init": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.c_uint64, c_void_p, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tda_graph_components": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_size_t, c_void_p]),
    "tda_spectral_embed": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_int, c_int, ctypes.c_uint64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tda_silhouette": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tda_greedy_perm_workspace_bytes": (c_size_t, [c_int]),
    "tda_greedy_perm": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tda_rips_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_size_t]),
    "tda_rips": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                         c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]),
    "tda_rips_launch": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]),
    "tda_rips_sort_edges_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tda_rips_sort_edges": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tda_rips_subsets_launch": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]),
    "tda_rips_h2_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_size_t, c_size_t]),
    "tda_rips_h2": (c_int, [c_void_p, c_int, c_int, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
    "tda_rips_stats": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_size_t, c_void_p]),
}

TDA_ERR_CAPACITY = -4
RIPS_STATS = 24   # TDA_RIPS_STATS

# The library itself reads no environment variables; this host glue maps the historical TDA_* variables onto
# tda_set_option when the library is loaded (scripts/ and the A/B runs in profiles/ use them).
_REDUCERS = {"sweep2": 0, "sweep": 1, "verify": 2, "bitset": 3}
_ENV_OPTIONS = {
    "TDA_RIPS_REDUCER": ("rips_reducer", lambda v: _REDUCERS[v]),
    "TDA_RIPS_W0": ("rips_w0", int), "TDA_RIPS_WSPARSE": ("rips_wsparse", int), "TDA_RIPS_WMAX": ("rips_wmax", int),
    "TDA_RIPS_DENSE_MIN": ("rips_dense_min", int), "TDA_RIPS_DENSE_DIV": ("rips_dense_div", int),
    "TDA_RIPS_CLUSTER": ("rips_cluster", int), "TDA_RIPS_WARP_ENGINE": ("rips_warp_engine", int), "TDA_SWEEP_EXCLUSIVE": ("sweep_exclusive", int),
    "TDA_SGD_MODE": ("sgd_mode", int), "TDA_SGD_CLUSTER": ("sgd_cluster", int), "TDA_SPECTRAL_CLUSTER": ("spectral_cluster", int),
    "TDA_KNN_LOADS": ("knn_loads", int), "TDA_DEBUG_SYNC": ("debug_sync", lambda v: 1), "TDA_H2_STATS": ("h2_stats", lambda v: 1),
}


def set_option(name, value):
    """Process-wide tuning option of the library (include/tda_b200.h: tda_set_option)."""
    check(lib().tda_set_option(name.encode(), int(value)))


def get_option(name):
    return int(lib().tda_get_option(name.encode()))


def rips_reducer():
    return {v: k for k, v in _REDUCERS.items()}[get_option("rips_reducer")]
STAGES = ["pdist_prep", "pdist_gemm", "knn_smooth", "fuzzy_graph", "spectral_init", "umap_sgd", "rips_pdist", "rips_edge_sort",
          "rips_h0", "rips_apparent", "rips_reduce"]


def stage_times():
    """{stage: (milliseconds, calls)} accumulated since the last reset (synchronises the recorded events)."""
    import numpy as np
    ms = np.zeros(len(STAGES), dtype=np.float64)
    calls = np.zeros(len(STAGES), dtype=np.int64)
    lib().tda_stage_timing_read(ms.ctypes.data, calls.ctypes.data, len(STAGES))
    return {s: (float(ms[i]), int(calls[i])) for i, s in enumerate(STAGES)}


def stage_timeline(cap=65536):
    """[(stage name, start ms, end ms)] of every timed stage span since the last reset (spans of different streams overlap)."""
    import numpy as np
    st = np.zeros(cap, dtype=np.int32)
    t0 = np.zeros(cap, dtype=np.float64)
    t1 = np.zeros(cap, dtype=np.float64)
    k = min(cap, int(lib().tda_stage_timeline_read(st.ctypes.data, t0.ctypes.data, t1.ctypes.data, cap)))
    return [(STAGES[int(st[i])], float(t0[i]), float(t1[i])) for i in range(k)]


class TdaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libtda_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Loads the library; fails loudly when it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C tda_multimodal_b200/csrc`). There is no CPU fallback for this path.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
        for env, (opt, conv) in _ENV_OPTIONS.items():
            v = os.environ.get(env)
            if v not in (None, ""):
                check(l.tda_set_option(opt.encode(), int(conv(v))))
    return _lib


def check(code):
    if code != 0:
        raise TdaError(code, lib().tda_last_error().decode("utf-8", "replace"))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("tda_multimodal_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
