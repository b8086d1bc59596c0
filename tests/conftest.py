import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu` on the GPU box")


@pytest.fixture(scope="session")
def cuda():
    """GPU tests FAIL (they do not skip) without a device: the product has no CPU path."""
    import torch
    assert torch.cuda.is_available(), "this test needs a CUDA device (run it through gpurun)"
    return torch.device("cuda:0")
