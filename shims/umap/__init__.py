"""Import-compatible front door: `import umap; umap.UMAP(...)` (debug_tda_pipeline.py:9,96-104;
analyze_tda_over_layers.py:10,38-44; analyze_adversarial_tda.py:11,85-93) resolves here when `<repo>/shims` is on
sys.path, and runs on libtda_b200.so."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from tda_multimodal_b200.umap_ import UMAP, find_ab_params  # noqa: E402

__all__ = ["UMAP", "find_ab_params"]
