import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # scipy.signal.butter is called with a one-element list of corners, exactly as pyrocko does; numpy >= 1.25 warns about it
    config.addinivalue_line("filterwarnings", "ignore:Conversion of an array with ndim > 0 to a scalar:DeprecationWarning")


def _cuda_device_present():
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests/` on a box without a GPU: gpu-marked tests are skipped, not failed (the product itself never falls
    back -- it raises; the skip only concerns the test run)."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the reference's own code (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The C checkers are built on demand (gcc only); the product never loads them."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "oracle", "libfsport.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    # libbeatgpu.so is git-ignored: a fresh checkout has to compile it before any test loads it (host-only entries such as
    # beatgpu_trace_append are used by CPU tests, too).  No-op when the library is up to date; left alone without nvcc.
    try:
        from beat_b200.build import build, find_nvcc
        find_nvcc()
    except Exception:
        return
    build()
