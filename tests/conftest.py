import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.build()
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def core():
    """The CUDA product path. Fails loudly (no fallback) if the library or the device is missing."""
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from nrslam_b200 import api
    c = api.Core()
    yield c
    c.close()
