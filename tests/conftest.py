import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices():
    try:
        import dspb200
        return dspb200.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a device: on a box without one (this build container) they are skipped, not failed."""
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_devices() >= 1:
        return
    skip = pytest.mark.skip(reason="no CUDA device (dspb200 has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def goldens():
    """Reference golden vectors (see tests/golden/import_reference_goldens.py)."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_goldens.npz")))


def relerr(a, b):
    """Norm-relative error, the measure behind Julia's `isapprox` (||a-b|| / max(||a||,||b||))."""
    a = np.asarray(a)
    b = np.asarray(b)
    na = np.linalg.norm(a.ravel().astype(np.complex128))
    nb = np.linalg.norm(b.ravel().astype(np.complex128))
    d = np.linalg.norm((a.astype(np.complex128) - b.astype(np.complex128)).ravel())
    m = max(na, nb)
    return d / m if m > 0 else d


def approx(a, b, rtol=None):
    """Julia's `a ≈ b`: rtol defaults to sqrt(eps) of the (narrower) eltype."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return False
    if rtol is None:
        ea = np.finfo(a.dtype).eps if np.issubdtype(a.dtype, np.inexact) else 0.0
        eb = np.finfo(b.dtype).eps if np.issubdtype(b.dtype, np.inexact) else 0.0
        rtol = np.sqrt(max(ea, eb)) if max(ea, eb) > 0 else 0.0
    return relerr(a, b) <= rtol
