import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def scans():
    """The reference's shipped scans (tests/golden/scans_*.npz, written by tests/golden/make_golden.py)."""
    out = {}
    for name in ("sim_structured", "sim_unstructured"):
        z = np.load(os.path.join(GOLDEN, "scans_%s.npz" % name))
        out[name] = (z["pts"], z["origins"])
    return out


def golden(name):
    return np.load(os.path.join(GOLDEN, name))
