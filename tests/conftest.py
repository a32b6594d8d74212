import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture
def tda_option():
    """set_option(name, value) for the duration of a test (process-wide library options, restored afterwards)."""
    from tda_multimodal_b200 import _lib
    saved = {}

    def setter(name, value):
        if name not in saved:
            saved[name] = _lib.get_option(name)
        _lib.set_option(name, value)
    yield setter
    for k, v in saved.items():
        _lib.set_option(k, v)
