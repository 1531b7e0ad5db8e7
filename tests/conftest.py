import os
import sys

import pytest

# several ranks share ONE GPU in the single-process harness (tests/gpu_util.py): their collective
# kernels must be able to run concurrently, so give every stream its own hardware queue
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bluebottle-3.0_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here; GPU tests run under gpurun")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
