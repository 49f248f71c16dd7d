import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a CPU-only host skips the gpu-marked tests (the driver selects them with -m gpu on the B200 box,
    where a missing device or library must FAIL, not skip)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run with -m gpu on a B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if stale) and load the C-ABI library; GPU tests depend on it and fail loudly without it -- on a box WITH a GPU.
    On a CPU-only host a plain `pytest` (no -m filter) skips them instead of erroring in the fixture."""
    from pagnerf_b200 import build, _lib
    build.build()
    return _lib.load()
