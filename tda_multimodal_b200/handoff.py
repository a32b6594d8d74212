"""Device-resident hand-off between activation extraction and the TDA sweep.

The reference moves every hidden state to the host inside the forward hook (``output[0].detach().cpu()``,
extract_activations.py:34-40), keeps one ``[hidden]`` vector per (sample, layer) in a dict of dicts, ``torch.save``s it
(:126-141), and the analysis scripts ``torch.load`` it and re-stack one ``[N, hidden]`` float64 cloud per layer
(debug_tda_pipeline.py:46-65).  Here the last-token vector goes straight from the hook's output into a preallocated
``[n_layers, n_samples, hidden]`` float32 tensor on the device the model runs on; that tensor is what
``pipeline.layer_sweep`` consumes, so nothing crosses PCIe between the forward pass and the first distance GEMM.

  ActivationCollector  the hooks + the [L, N, hidden] buffer (same hook registration and last-token rule as the reference)
  stack_clouds         the reference's cloud assembly for an ``all_activations.pt``-style dict (for files that already exist)

Both produce the layout ``X[layer, sample, hidden]`` with samples ordered by sorted id, as debug_tda_pipeline.py:46-49 does.
"""
import numpy as np


def _torch():
    import torch
    return torch


class ActivationCollector:
    """Forward hooks on the decoder layers that keep the last-token hidden state ON THE DEVICE.

    Usage, mirroring extract_activations.py:42-132::

        col = ActivationCollector(model.transformer.h, n_samples=len(metadata), hidden=model.config.hidden_size)
        for item in metadata:
            col.begin(item["id"], item)            # layer_activations.clear()
            with torch.no_grad():
                model(**inputs)
            col.commit(last_token_idx)             # activation_tensor[0, last_token_idx, :] of every layer
        col.remove()
        ids, X = col.clouds(lambda meta: meta["type"] == "bound")   # X [L, N, hidden] float32 on the device
        out = pipeline.layer_sweep(X, n_neighbors=6)

    ``to_all_results()`` returns the reference's ``all_results`` dict (``{id: {"metadata", "activations": {"layer_i": vec}}}``,
    extract_activations.py:129-132) with host tensors, for scripts that still want ``all_activations.pt``.
    """

    def __init__(self, layers, n_samples, hidden=None, device=None, dtype=None):
        torch = _torch()
        self.layers = list(layers)
        self.n_layers = len(self.layers)
        self.n_samples = int(n_samples)
        self.hidden = hidden
        self.device = device
        self.dtype = dtype or torch.float32
        self.buffer = None                      # [L, n_samples, hidden], allocated at the first commit if hidden is unknown
        self.ids, self.meta = [], []
        self._pending = {}
        self._current = None
        self.handles = [layer.register_forward_hook(self._hook(i)) for i, layer in enumerate(self.layers)]

    def _hook(self, index):
        def hook(module, inputs, output):
            # output[0] holds the hidden states [1, seq, hidden] (extract_activations.py:37-39); keep a view, no copy, no .cpu()
            hs = output[0] if isinstance(output, (tuple, list)) else output
            self._pending[index] = hs.detach()
        return hook

    def begin(self, sample_id, metadata=None):
        if len(self.ids) >= self.n_samples:
            raise ValueError(f"ActivationCollector: more than n_samples={self.n_samples} samples")
        self._pending.clear()
        self._current = (sample_id, metadata)

    def commit(self, last_token_idx):
        """Copy ``hidden[0, last_token_idx, :]`` of every hooked layer into row ``len(ids)`` of the buffer (device to device).
        An index past the sequence falls back to -1, as extract_activations.py:121-123 does.  Returns False (and records
        nothing) when no hook fired, the reference's "No activations captured" case (:112-114)."""
        torch = _torch()
        if self._current is None:
            raise RuntimeError("ActivationCollector.commit() without begin()")
        if not self._pending:
            self._current = None
            return False
        if len(self._pending) != self.n_layers:
            missing = sorted(set(range(self.n_layers)) - set(self._pending))
            raise RuntimeError(f"ActivationCollector: layers {missing} produced no output for sample {self._current[0]!r}")
        first = self._pending[0]
        if self.buffer is None:
            self.hidden = int(first.shape[-1]) if self.hidden is None else int(self.hidden)
            self.device = first.device if self.device is None else torch.device(self.device)
            self.buffer = torch.empty((self.n_layers, self.n_samples, self.hidden), dtype=self.dtype, device=self.device)
        row = len(self.ids)
        for i in range(self.n_layers):
            hs = self._pending[i]
            if hs.dim() == 2:
                hs = hs[None]
            idx = last_token_idx
            if idx >= hs.shape[1]:
                idx = -1
            self.buffer[i, row].copy_(hs[0, idx, :], non_blocking=True)
        self.ids.append(self._current[0])
        self.meta.append(self._current[1])
        self._pending.clear()
        self._current = None
        return True

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []

    def clouds(self, select=None):
        """(sample ids sorted as debug_tda_pipeline.py:46-49 sorts them, X [L, N, hidden] on the device).  ``select`` filters
        on the per-sample metadata (the reference's ``data["metadata"]["type"] == POINT_CLOUD_TYPE``)."""
        torch = _torch()
        if self.buffer is None:
            raise RuntimeError("ActivationCollector: nothing collected")
        keep = [i for i in range(len(self.ids)) if select is None or select(self.meta[i])]
        keep.sort(key=lambda i: self.ids[i])
        index = torch.as_tensor(keep, dtype=torch.long, device=self.buffer.device)
        X = self.buffer[:, :len(self.ids)].index_select(1, index).contiguous()
        return [self.ids[i] for i in keep], X

    def to_all_results(self):
        host = self.buffer[:, :len(self.ids)].to("cpu")
        out = {}
        for s, (sid, meta) in enumerate(zip(self.ids, self.meta)):
            out[sid] = {"metadata": meta, "activations": {f"layer_{i}": host[i, s].clone() for i in range(self.n_layers)}}
        return out


def stack_clouds(all_data, point_cloud_type=None, n_layers=None, device=None):
    """The cloud assembly of debug_tda_pipeline.py:46-65 (same in analyze_tda_over_layers.py:47-51,
    analyze_adversarial_tda.py:73-78) for an ``all_activations.pt``-style dict: ids of the requested metadata type, sorted;
    one stacked cloud per layer.  Returns (sample_ids, X [L, N, hidden] float32 on ``device``) -- float32 because that is what
    UMAP computes in (the reference's float64 copy is cast back by ``check_array``)."""
    torch = _torch()
    ids = sorted(i for i, d in all_data.items() if point_cloud_type is None or d["metadata"]["type"] == point_cloud_type)
    if not ids:
        raise ValueError(f"stack_clouds: no samples of type {point_cloud_type!r}")
    if n_layers is None:
        n_layers = len(all_data[ids[0]]["activations"])
    layers = []
    for i in range(n_layers):
        name = f"layer_{i}"
        layers.append(torch.stack([torch.as_tensor(np.asarray(all_data[s]["activations"][name])) if not hasattr(all_data[s]["activations"][name], "dim")
                                   else all_data[s]["activations"][name] for s in ids]))
    X = torch.stack(layers).to(dtype=torch.float32)
    if device is not None:
        X = X.to(device, non_blocking=True)
    return ids, X
