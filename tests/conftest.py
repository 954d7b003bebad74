import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: long CPU test")


@pytest.fixture(scope="session")
def fx():
    from tests.helpers import Fixtures
    return Fixtures()


def _cuda_device_present() -> bool:
    """True when libbevk.so can create a context (a CUDA device is visible).  A missing library is NOT a reason to skip:
    on a GPU box that must fail loudly."""
    try:
        import ctypes
        from cameracalibration_b200 import _lib as L
        h = ctypes.c_void_p()
        lib = L.load()
        if lib.bevk_ctx_create(0, ctypes.byref(h)) != 0:
            return False
        lib.bevk_ctx_destroy(h)
        return True
    except RuntimeError:
        raise
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (libbevk.so has no CPU fallback); run under gpurun")
    for it in gpu_items:
        it.add_marker(skip)
