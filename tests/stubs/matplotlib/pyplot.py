"""TEST STUB of matplotlib.pyplot: every call is accepted and ignored; subplots() returns arrays of stub axes."""
import numpy as _np


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


class _Axes(_Anything):
    pass


_CUR = _Axes()


def gca():
    return _CUR


def gcf():
    return _Anything()


def figure(*a, **k):
    return _Anything()


def subplots(nrows=1, ncols=1, *a, **k):
    if nrows == 1 and ncols == 1:
        return _Anything(), _Axes()
    axes = _np.empty((nrows, ncols), dtype=object)
    for i in range(nrows):
        for j in range(ncols):
            axes[i, j] = _Axes()
    if nrows == 1 or ncols == 1:
        axes = axes.reshape(-1)
    return _Anything(), axes


def savefig(path, *a, **k):
    open(path, "wb").close()


def __getattr__(name):   # plot, title, subplot, xlabel, ylabel, legend, grid, tight_layout, close, show, ...
    return lambda *a, **k: _Anything()
