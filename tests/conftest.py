import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_ok():
    try:
        import ctypes
        from voxel_ma_b200 import api
        lib = api.load_library()
        h = ctypes.c_void_p()
        if lib.vc_ctx_create(0, ctypes.byref(h)) != 0:
            return False
        lib.vc_ctx_destroy(h)
        return True
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx_factory():
    """GPU tests construct contexts through this; a missing library or device FAILS (no silent
    skip, no CPU fallback) -- `-m gpu` is only selected on a GPU box."""
    from voxel_ma_b200 import api

    made = []

    def make(device=0):
        c = api.Context(device)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()
