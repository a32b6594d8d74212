import sys
sys.path.insert(0, "/root/repo")
import torch
from tda_multimodal_b200 import _lib, umap_, workloads
X = torch.from_numpy(workloads.c3_layers(n_layers=32, layers=[0, 1, 9, 31])).cuda()
_lib.set_option("spectral_debug", 1)
for l in range(4):
    print("layer", [0, 1, 9, 31][l]); sys.stdout.flush()
    Y, st = umap_.umap_fit_batch(X[l:l + 1], n_neighbors=15, n_components=3, metric="cosine", random_state=42, n_epochs=0, defer_component_check=True)
    torch.cuda.synchronize()
