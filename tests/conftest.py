import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the host-side index paths (Lifter tables on NumPy vectors, exchange plans on CPU tensors over gloo) are test-only:
# they raise unless switched on; spawned gloo ranks inherit the variable
os.environ.setdefault("TATVA_B200_HOST_TABLES", "1")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a CUDA device skips the GPU tests instead of failing them."""
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference, written by tests/golden/make_golden.py."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
