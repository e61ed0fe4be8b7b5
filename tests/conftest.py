import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_synth():
    return dict(np.load(os.path.join(GOLDEN, "synth_c1.npz")))


@pytest.fixture(scope="session")
def golden_real():
    return dict(np.load(os.path.join(GOLDEN, "real_pair.npz")))


@pytest.fixture(scope="session")
def ctx():
    """One GPU context for the whole session; fails loudly when no GPU / library is present."""
    from laser_3d_reconstruction_b200 import _native
    c = _native.Context(0)
    yield c
    c.close()
