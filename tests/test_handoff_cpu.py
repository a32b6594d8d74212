"""Host logic of the device-resident hand-off (tda_multimodal_b200/handoff.py) against the reference's own flow:
hooks that .cpu() every hidden state, the all_results dict (extract_activations.py:34-40,110-132) and the per-layer cloud
assembly of debug_tda_pipeline.py:46-65.  Runs on the CPU with a toy decoder (no GPU, no Qwen weights)."""
import numpy as np
import torch

from tda_multimodal_b200 import handoff


class ToyLayer(torch.nn.Module):
    def __init__(self, hidden, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w = torch.nn.Parameter(torch.randn(hidden, hidden, generator=g) / hidden ** 0.5)

    def forward(self, x):
        return (torch.tanh(x @ self.w) + x, None)       # a decoder layer returns a tuple; [0] = hidden states


class ToyModel(torch.nn.Module):
    def __init__(self, n_layers=5, hidden=24):
        super().__init__()
        self.h = torch.nn.ModuleList([ToyLayer(hidden, 100 + i) for i in range(n_layers)])

    def forward(self, x):
        for layer in self.h:
            x = layer(x)[0]
        return x


def _inputs(n, hidden):
    g = torch.Generator().manual_seed(7)
    items = []
    for s in range(n):
        seq = 3 + (s * 5) % 7
        items.append({"id": f"s{(n - s):03d}", "type": "bound" if s % 3 else "unbound", "x": torch.randn(1, seq, hidden, generator=g),
                      "last": seq - 2 if s % 4 else seq + 3})   # some indices past the end: the reference falls back to -1
    return items


def _reference_flow(model, items):
    """extract_activations.py, literally: hook -> output[0].detach().cpu(); dict of dicts; last-token vector per layer."""
    layer_activations = {}

    def get_hook(name):
        def hook(module, inp, out):
            layer_activations[name] = out[0].detach().cpu()
        return hook
    handles = [layer.register_forward_hook(get_hook(f"layer_{i}")) for i, layer in enumerate(model.h)]
    all_results = {}
    for item in items:
        layer_activations.clear()
        with torch.no_grad():
            model(item["x"])
        last = item["last"]
        sample = {}
        for name, act in layer_activations.items():
            if last >= act.shape[1]:
                last = -1
            sample[name] = act[0, last, :].clone()
        all_results[item["id"]] = {"metadata": {"id": item["id"], "type": item["type"]}, "activations": sample}
    for h in handles:
        h.remove()
    return all_results


def _reference_clouds(all_data, kind, n_layers):
    """debug_tda_pipeline.py:46-65."""
    sample_ids = sorted([i for i, d in all_data.items() if d["metadata"]["type"] == kind])
    clouds = []
    for i in range(n_layers):
        cloud = [all_data[s]["activations"][f"layer_{i}"] for s in sample_ids]
        clouds.append(torch.stack(cloud).numpy().astype(np.float64))
    return sample_ids, clouds


def test_collector_equals_reference_flow():
    model = ToyModel().eval()
    items = _inputs(11, 24)
    want = _reference_flow(model, items)
    col = handoff.ActivationCollector(model.h, n_samples=len(items))
    for item in items:
        col.begin(item["id"], {"id": item["id"], "type": item["type"]})
        with torch.no_grad():
            model(item["x"])
        assert col.commit(item["last"])
    col.remove()
    assert not any(layer._forward_hooks for layer in model.h)
    got = col.to_all_results()
    assert list(got) == list(want)
    for sid in want:
        assert got[sid]["metadata"] == want[sid]["metadata"]
        for name, vec in want[sid]["activations"].items():
            assert torch.equal(got[sid]["activations"][name], vec)
    ids_ref, clouds_ref = _reference_clouds(want, "bound", 5)
    ids, X = col.clouds(lambda m: m["type"] == "bound")
    assert ids == ids_ref and X.shape == (5, len(ids), 24) and X.dtype == torch.float32 and X.is_contiguous()
    for i in range(5):
        assert np.array_equal(X[i].numpy().astype(np.float64), clouds_ref[i])
    # the same assembly from the saved-dict format
    ids2, X2 = handoff.stack_clouds(want, "bound")
    assert ids2 == ids_ref and torch.equal(X2, X)
    ids3, X3 = handoff.stack_clouds(want)
    assert len(ids3) == 11 and X3.shape == (5, 11, 24)


def test_collector_edge_cases():
    model = ToyModel(n_layers=2, hidden=8).eval()
    col = handoff.ActivationCollector(model.h, n_samples=1)
    col.begin("a")
    assert col.commit(0) is False            # no forward pass ran: nothing captured (extract_activations.py:112-114)
    col.begin("a")
    with torch.no_grad():
        model(torch.zeros(1, 2, 8))
    assert col.commit(1) is True
    try:
        col.begin("b")
        raise AssertionError("expected ValueError")
    except ValueError:
        pass
    try:
        handoff.stack_clouds({"a": {"metadata": {"type": "x"}, "activations": {}}}, "y")
        raise AssertionError("expected ValueError")
    except ValueError:
        pass
