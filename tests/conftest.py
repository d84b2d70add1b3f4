import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def pkg():
    from pnpb200_loader import load_package
    return load_package()


@pytest.fixture(scope="session")
def gpu_lib(pkg):
    L = pkg.lib()
    if L.pnpb200_device_count() < 1:
        pytest.fail("no CUDA device visible: the -m gpu tests must run on the GPU box (no CPU fallback)")
    return L
