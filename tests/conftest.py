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


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def bits_equal(a, b):
    """Bitwise equality of two float64 arrays (NaN-safe, sign-of-zero-sensitive)."""
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()  # builds liboracle.so on first use if it is missing
    return o


@pytest.fixture(scope="session")
def cpp_ls():
    """The product's Python boundary.  GPU tests call through it (ctypes -> C ABI -> CUDA)."""
    from movie_recommender_b200 import cpp_ls as m
    return m


@pytest.fixture(scope="session")
def require_gpu(cpp_ls):
    from movie_recommender_b200 import _lib
    if _lib.dll.mrb_device_count() < 1:
        pytest.fail("gpu-marked test but no CUDA device is visible (no CPU fallback exists)")
