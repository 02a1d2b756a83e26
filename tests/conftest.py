import os, sys
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return lambda name: np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    """the oracle's C pieces are built from source on demand (gcc); _ref only where the reference exists"""
    from oracle.core import ensure_built
    ensure_built()
