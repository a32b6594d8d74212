"""Packs the reference's own shipped Rips artefacts into one small fixture.

Source (read-only reference tree, only available in the build container):
  /root/reference/tda-output/point_clouds_3d/layer_{0..31}_cloud.npy   (36,3) float32 UMAP outputs
  /root/reference/tda-output/summary_stats.json                        ripser(maxdim=1) statistics of those clouds
Both were produced by the reference authors running debug_tda_pipeline.py:104-131 with the real
`ripser` package, so they are known-answer vectors for the Rips stage.
Run:  python tests/golden/make_rips_golden.py
"""
import json
import os

import numpy as np

REF = "/root/reference/tda-output"
HERE = os.path.dirname(os.path.abspath(__file__))

clouds = np.stack([np.load(os.path.join(REF, "point_clouds_3d", f"layer_{i}_cloud.npy")) for i in range(32)])
assert clouds.shape == (32, 36, 3) and clouds.dtype == np.float32
np.save(os.path.join(HERE, "ref_clouds_3d.npy"), clouds)
with open(os.path.join(REF, "summary_stats.json")) as f:
    stats = json.load(f)
with open(os.path.join(HERE, "ref_summary_stats.json"), "w") as f:
    json.dump(stats, f, indent=2)
print("wrote", clouds.shape, len(stats))
