import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "gpu_unverified: GPU test written after the round's GPU budget was spent; it has never run "
                            "on a device and only runs with MAGIC_UNVERIFIED_GPU=1 (first GPU call of the next round)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not os.environ.get("MAGIC_UNVERIFIED_GPU"):
        hold = pytest.mark.skip(reason="never run on a device yet: set MAGIC_UNVERIFIED_GPU=1")
        for item in items:
            if "gpu_unverified" in item.keywords:
                item.add_marker(hold)
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
