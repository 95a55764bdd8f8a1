import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    return oracle.load()


@pytest.fixture(scope="session")
def host_backend():
    import ne_b200
    return ne_b200.NumpyHostBackend()


@pytest.fixture(scope="session")
def cuda_backend():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ne_b200
    return ne_b200.TorchCudaBackend("cuda:0")


@pytest.fixture(scope="session")
def cuda_lib():
    import ne_b200
    lib = ne_b200.get_library()   # raises if the extension is missing: no CPU fallback
    assert lib.device_count() > 0, "libne_b200.so sees no CUDA device"
    return lib
