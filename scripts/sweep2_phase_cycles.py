"""Counters and phase cycles of the sweep2 reducer per cloud (option rips_debug) on the first 11 layers of the C3 workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tda_multimodal_b200 import _lib, umap_, rips, workloads
X = torch.from_numpy(workloads.c3_layers(n_layers=32, layers=range(11))).cuda()
Y = umap_.umap_fit_batch(X, n_neighbors=15, n_components=3, metric="cosine", random_state=42)
dm = rips.pdist_lowdim(Y)
rips.rips_batch(dm, maxdim=1)
torch.cuda.synchronize()
_lib.set_option("rips_debug", 1)
rips.rips_batch(dm, maxdim=1)
torch.cuda.synchronize()
