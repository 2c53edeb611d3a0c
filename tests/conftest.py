import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device skips the gpu-marked tests instead of failing them."""
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything native is built in-tree once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as entry
    entry.build()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def has_gpu():
    """A usable CUDA device?  Asked of the library itself (cudaGetDeviceCount through pdp_device_count), so that GPU runs do
    not pay for `import torch`; torch is the fallback when the library is not built yet."""
    try:
        from pyro_b200.engine import device_count
        return device_count() > 0
    except Exception:
        pass
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
