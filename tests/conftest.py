import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _no_fragment_cache():
    """Every test exercises the kernel it names: Fragments are not reused between back-to-back identical renders
    unless a test switches the cache on (tests/test_gpu_parity.py::test_fragment_cache_*)."""
    try:
        from torch_renderer_b200.rasterizer import set_fragment_cache
    except Exception:
        yield
        return
    set_fragment_cache(False)
    yield
    set_fragment_cache(False)
