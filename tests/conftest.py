import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    try:
        import ctypes

        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch

            return torch.cuda.device_count()
        except Exception:
            return 0
    n = ctypes.c_int(0)
    return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a box without a GPU: skip the gpu-marked tests instead of failing."""
    if not any("gpu" in it.keywords for it in items):
        return
    try:
        import torch

        have = torch.cuda.is_available()
    except Exception:
        have = _cuda_device_count() > 0
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (pogs_b200 has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    return np.load(path)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_ctypes as O

    O.lib()   # builds oracle/liboracle.so with gcc on first use
    return O


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
