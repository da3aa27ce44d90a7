import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_sessionstart(session):
    """The C-ABI library is git-ignored (history stays source-only): on a fresh checkout build it once, exactly as
    ``__graft_entry__.build()`` does, so that the load / export tests exercise the real product library."""
    lib = os.path.join(ROOT, "gipmed-project-self-supervised-vit_b200", "libb200ssl.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with `pytest -m gpu` under gpurun")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
