import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as ge
    ge.build()
    return ge


@pytest.fixture(scope="session")
def host_shim(built):
    import ctypes
    return ctypes.CDLL(os.path.join(ROOT, "tests", "_build", "host_shim.so"))


@pytest.fixture(scope="session")
def gpu(built):
    import gync_b200
    from gync_b200 import lib as L
    L.check(L.load().gync_init(0))
    return gync_b200
