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
def golden_orders():
    return np.load(os.path.join(GOLDEN, "orders.npz"))


@pytest.fixture(scope="session")
def golden_scanmerge():
    return np.load(os.path.join(GOLDEN, "scanmerge.npz"))


@pytest.fixture(scope="session")
def golden_dct():
    return np.load(os.path.join(GOLDEN, "dct.npz"))


@pytest.fixture(scope="session")
def golden_ablation():
    return np.load(os.path.join(GOLDEN, "ablation_orders.npz"))
