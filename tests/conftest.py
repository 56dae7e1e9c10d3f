import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "experimental: code paths that are not enabled by default; run with ROITR_EXPERIMENTAL=1 on a GPU")


def pytest_collection_modifyitems(config, items):
    import torch
    if os.environ.get("ROITR_EXPERIMENTAL") != "1":
        skip_x = pytest.mark.skip(reason="experimental path: set ROITR_EXPERIMENTAL=1 (GPU box)")
        for item in items:
            if "experimental" in item.keywords:
                item.add_marker(skip_x)
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
