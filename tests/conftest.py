import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never pass silently on a box without a device: they are skipped only when the marker
    expression does not select them; if selected without a device they fail inside the library."""
    pass


@pytest.fixture(scope="session")
def ref():
    """the wrapped, unmodified reference (oracle/_ref/libdgref.so) or None when it has not been built"""
    from oracle import refwrap
    return refwrap if refwrap.available() else None


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    return np.load(path)
