"""TEST STUB of matplotlib (the GPU image has none): records nothing, draws nothing; `savefig` creates an empty file so that
scripts which list their outputs still find them.  Only used by tests that execute the reference's scripts (plotting is outside
the hot path)."""
__version__ = "0.0-stub"


def use(*a, **k):
    pass
