import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cs():
    """The product library, initialised on cuda:0.  GPU tests fail loudly without it."""
    import cudasift_b200 as m
    m.InitCuda(0)
    return m


@pytest.fixture(scope="session")
def reflib(cs):
    """The unmodified reference built into oracle/_ref (None when it did not travel)."""
    import reflib as r
    return r.load_reference()


@pytest.fixture(scope="session")
def selflib(cs):
    """libcudasift_b200.so driven through the reference's own C++ (mangled) API."""
    import reflib as r
    from cudasift_b200 import build
    return r.CxxSiftLib(build.LIB)
