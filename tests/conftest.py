import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_package  # noqa: E402

load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    """A plain ``pytest tests`` on a box without a GPU (or without the built library) skips the gpu-marked tests
    instead of failing them; ``-m gpu`` on the B200 box runs them."""
    import pytest
    import torch
    from switchfl_b200 import backend
    reason = None
    if not torch.cuda.is_available():
        reason = "needs a CUDA device"
    elif not os.path.exists(backend.LIB_PATH):
        reason = "libswitchfl_b200.so is not built"
    if reason:
        skip = pytest.mark.skip(reason=reason)
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)
