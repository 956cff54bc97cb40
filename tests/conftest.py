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


# ----------------------------------------------------------------------------------------------------------------
# TRAMBA_GPU_TESTS_ON_SIM=1: run the `-m gpu` tests WITHOUT a GPU, against the product's own C ABI and kernels executed
# on the CPU simulator (tools/cpu_sim; tests/_abi_emulator.py: installed("sim")).  "cuda" devices are mapped to the
# CPU for the session.  Meant for the small / medium cases (`-k ...`): BASELINE-size tests take minutes to hours there.
#     TRAMBA_GPU_TESTS_ON_SIM=1 python -m pytest tests/test_z_gpu_ablation.py -m gpu -q
# ----------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="session", autouse=True)
def _gpu_tests_on_the_simulator():
    mode = os.environ.get("TRAMBA_GPU_TESTS_ON_SIM", "")
    if mode != "1":
        yield
        return
    import functools
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _abi_emulator import installed

    def cpu_device(d):
        return "cpu" if (isinstance(d, str) and d.startswith("cuda")) or (
            isinstance(d, torch.device) and d.type == "cuda") else d

    def wrap(fn):
        @functools.wraps(fn)
        def inner(*a, **k):
            if "device" in k:
                k["device"] = cpu_device(k["device"])
            return fn(*a, **k)
        return inner

    names = ["zeros", "ones", "empty", "randn", "rand", "randint", "tensor", "arange", "full", "zeros_like",
             "ones_like", "empty_like", "randn_like", "eye", "linspace"]
    saved = {n: getattr(torch, n) for n in names}
    saved_to, saved_sync = torch.Tensor.to, torch.cuda.synchronize
    for n in names:
        setattr(torch, n, wrap(saved[n]))

    def to(self, *a, **k):
        moved = any(cpu_device(x) is not x for x in a) or ("device" in k and cpu_device(k["device"]) is not k["device"])
        a = tuple(cpu_device(x) for x in a)
        if "device" in k:
            k["device"] = cpu_device(k["device"])
        out = saved_to(self, *a, **k)
        return out.clone() if moved and out is self else out       # a host -> device move is a copy

    torch.Tensor.to = to
    torch.cuda.synchronize = lambda *a, **k: None
    with installed("sim", sms=4):
        yield
    for n in names:
        setattr(torch, n, saved[n])
    torch.Tensor.to, torch.cuda.synchronize = saved_to, saved_sync
