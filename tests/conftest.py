import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))   # tests may import the oracle (checker only)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "ref_golden.pt"), map_location="cpu")


@pytest.fixture(scope="session")
def golden_infer():
    """Optimizer moments + encoder-only inference outputs of the real reference (oracle/make_golden.py section 8)."""
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "ref_golden_infer.pt"), map_location="cpu")
