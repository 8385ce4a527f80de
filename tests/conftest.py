import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only; built on demand with gcc)."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def lead_urban():
    d = np.load(os.path.join(GOLDEN, "lead_TO01_EAD.npz"))
    return np.stack([d["s_tv"], d["v_tv"]], 1)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))
