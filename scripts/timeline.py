"""One C3 step (32 layers, 2 chunks on 2 streams) as a timeline of stage spans: which stage runs when, on the device clock
(CUDA events recorded by the library's stage timer).  Shows what overlaps and what the critical path of the step is."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tda_multimodal_b200 import _lib, pipeline, workloads
L = _lib.lib()
X = torch.from_numpy(workloads.c3_layers(n_layers=32)).cuda()
for _ in range(3):
    pipeline.layer_sweep(X)
torch.cuda.synchronize()
L.tda_stage_timing_enable(1)
L.tda_stage_timing_reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pipeline.layer_sweep(X)
e1.record()
torch.cuda.synchronize()
spans = _lib.stage_timeline()
L.tda_stage_timing_enable(0)
print(f"step {e0.elapsed_time(e1):.1f} ms; {len(spans)} spans")
t00 = min(s[1] for s in spans)
# merge consecutive spans of the same stage (e.g. 500 SGD epochs) per enqueue order
merged = []
for name, a, b in spans:
    if merged and merged[-1][0] == name and a - merged[-1][2] < 0.5:
        merged[-1][2] = max(merged[-1][2], b)
    else:
        merged.append([name, a, b])
for name, a, b in sorted(merged, key=lambda s: s[1]):
    print(f"{a - t00:8.2f} -> {b - t00:8.2f}  ({b - a:7.2f} ms)  {name}")
