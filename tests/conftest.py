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

    def load(name):
        d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        if "rx_in_int16" in d:          # off-air recording: int16 -> (x, 0) complex, unscaled (int16tof32.py --zeropad)
            d["rx_in"] = d["rx_in_int16"].astype(np.float32).astype(np.complex64)
        return d
    return load


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    """the oracle's C pieces are built from source on demand (gcc); _ref only where the reference exists"""
    from oracle.core import ensure_built
    ensure_built()
