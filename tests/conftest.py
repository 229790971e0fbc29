import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# FP64 parity bar of BASELINE.json:north_star -- relative L-infinity <= 1e-10
RTOL_LINF = 1e-10


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_linf(a, b):
    """max|a-b| / max|b| (0 if both are identically zero)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b)) if b.size else 0.0
    num = np.max(np.abs(a - b)) if b.size else 0.0
    return 0.0 if num == 0.0 else num / max(den, np.finfo(float).tiny)


def assert_close(a, b, tol=RTOL_LINF, what=""):
    err = rel_linf(a, b)
    assert err <= tol, f"{what}: relative Linf error {err:.3e} > {tol:.1e}"


@pytest.fixture(scope="session")
def have_gpu():
    import torch

    return torch.cuda.is_available()
