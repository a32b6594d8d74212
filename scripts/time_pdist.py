import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tda_multimodal_b200.pdist import pdist
for B, n, d in [(1, 2000, 4096), (32, 2000, 4096), (1, 16384, 4096)]:
    X = torch.randn(B, n, d, device="cuda")
    out = torch.empty(B, n, n, device="cuda")
    for _ in range(2): pdist(X, metric="cosine", out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): pdist(X, metric="cosine", out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * B * n * n * d
    print(f"B={B} n={n} d={d}: {ms:.3f} ms  useful {fl/ms/1e9:.1f} TFLOP/s  issued {3*fl/ms/1e9:.1f} TFLOP/s (incl. prep)", flush=True)
