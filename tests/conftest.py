import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cge():
    """The product package (directory name contains a hyphen, hence importlib)."""
    return importlib.import_module("computer-graphics-engine_b200")


@pytest.fixture(scope="session")
def ref():
    """ctypes binding of the unmodified reference engine (oracle/_ref, test infrastructure)."""
    import refharness
    if not refharness.available():
        pytest.skip("oracle/_ref/libcge_ref.so not built (needs /root/reference at build time)")
    return refharness


GOLDEN = ROOT / "tests" / "golden"


def compare_images(rgb_a, rgb_b):
    """NaN-aware comparison (the reference image contains NaN pixels).
    Returns (max_abs_err over finite pixels, number of pixels whose NaN-ness differs)."""
    a = np.asarray(rgb_a, np.float32).reshape(-1, 3)
    b = np.asarray(rgb_b, np.float32).reshape(-1, 3)
    na, nb = np.isnan(a), np.isnan(b)
    nan_mismatch = int((na != nb).any(axis=1).sum())
    both = ~(na | nb)
    err = np.abs(np.where(both, a, 0) - np.where(both, b, 0))
    inf_mismatch = int((np.isinf(a) != np.isinf(b)).sum())
    err = np.where(np.isinf(a) & np.isinf(b) & (a == b), 0, err)
    return float(np.nanmax(err)) if err.size else 0.0, nan_mismatch + inf_mismatch
