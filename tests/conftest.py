import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_dev():
    """CudaTensor.Dev, initialised on device 0. Fails loudly if the native library is not built."""
    from deepnet_b200 import CudaTensor
    dev = CudaTensor.dev()
    dev.Init(0)
    return dev


@pytest.fixture(scope="session")
def host_dev():
    from oracle.host_tensor import HostTensor
    return HostTensor.Dev
