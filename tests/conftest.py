import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present() -> bool:
    try:
        import torch

        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped (not errors), so CPU regressions stay visible."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200): run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree libraries exist (content-stamped: no rebuild when up to date)."""
    from mfmg_b200 import build

    build.build_host()
    build.build_oracle()
    build.build_cuda()
    yield


@pytest.fixture(scope="session")
def handle():
    from mfmg_b200.device import CudaHandle

    h = CudaHandle(0)
    yield h
    h.close()
