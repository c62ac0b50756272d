import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
