"""Pinned host -> device bandwidth of this box (what bounds the e2e figure of bench.py: 1.05 GB of activations per 32-layer step)."""
import torch
x = torch.empty((32, 2000, 4096), dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for nb in (1, 4, 8, 32):
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d[:nb].copy_(x[:nb], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(f"H2D {nb} layers ({nb * 32.768:.0f} MB): {ms:.2f} ms = {nb * 0.032768 / (ms / 1e3):.1f} GB/s")
