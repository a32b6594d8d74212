"""Import-compatible front door: `from ripser import ripser` (debug_tda_pipeline.py:10) resolves here when
`<repo>/shims` is on sys.path, and runs on libtda_b200.so."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from tda_multimodal_b200.rips import ripser  # noqa: E402


class Rips:  # minimal sklearn-style wrapper ripser.py also exports
    def __init__(self, maxdim=1, thresh=float("inf"), coeff=2, do_cocycles=False, n_perm=None, verbose=True):
        self.maxdim, self.thresh, self.coeff, self.do_cocycles, self.n_perm = maxdim, thresh, coeff, do_cocycles, n_perm

    def fit_transform(self, X, distance_matrix=False, metric="euclidean"):
        r = ripser(X, maxdim=self.maxdim, thresh=self.thresh, coeff=self.coeff, do_cocycles=self.do_cocycles,
                   distance_matrix=distance_matrix, metric=metric, n_perm=self.n_perm)
        self.dgms_ = r["dgms"]
        self.__dict__.update({k + "_": v for k, v in r.items() if k != "dgms"})
        return self.dgms_


__all__ = ["ripser", "Rips"]
