"""pytest configuration: registers the `gpu` marker and makes the repo root importable.

`-m "not gpu"` tests: the oracle against its pins (fp64 shadow, torch-fp64 autograd, identities, golden
fixtures), host logic, and that the C-ABI library loads and exports every declared symbol.
`-m gpu` tests: parity of the CUDA path (through the C ABI) against the oracle.
"""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(autouse=True)
def _release_device_tensors():
    yield
    try:
        from tests import gpu_util

        gpu_util.release()
    except Exception:
        pass
