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
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def cgb():
    """The CUDA engine through its C ABI.  Fails loudly when the extension or the GPU is missing."""
    import torch

    import cognn_b200

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    ctx = cognn_b200.Context(0)
    yield ctx
    ctx.close()
