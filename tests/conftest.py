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


# The Python mirrors bind the same five symbols as the reference's cpp_ls.py, so on a CPU box they
# can be pointed at the UNMODIFIED reference library (oracle/_ref/cpp_ls_lib.so): host logic
# (marshalling, RNG draw order, files in and out) is then checked without a GPU.
@pytest.fixture()
def cpp_ls_on_reference(monkeypatch):
    import ctypes

    from oracle import oracle
    if not oracle.has_ref():
        pytest.skip("oracle/_ref/cpp_ls_lib.so not built (reference sources absent)")
    from movie_recommender_b200 import _lib, cpp_ls
    ref = ctypes.CDLL(oracle.ref_path())
    c_int, c_double, I, D = ctypes.c_int, ctypes.c_double, _lib._I, _lib._D
    ref.set_thread_count.restype = None
    ref.set_thread_count.argtypes = [c_int]
    ref.get_thread_count.restype = c_int
    for name in ("cg_least_squares_from_python", "cg_least_squares2_from_python"):
        getattr(ref, name).restype = c_int
        getattr(ref, name).argtypes = [c_int, c_int, I, I, D, c_int, D, c_int, D, c_double, c_int, D]
    ref.als_from_python.restype = c_int
    ref.als_from_python.argtypes = [I, I, c_int, D, c_int, c_int, D, c_int, D, c_double, c_int, c_int]
    monkeypatch.setattr(cpp_ls, "_dll", ref)
    return cpp_ls
