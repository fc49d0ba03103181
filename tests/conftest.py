import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (oracle/ptg_oracle.c) -- test infrastructure only."""
    from oracle import ptg_oracle
    ptg_oracle.lib()
    return ptg_oracle


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference module built into oracle/_ref (None when it is not there)."""
    try:
        from oracle import build_ref
        if build_ref.build(verbose=False) is None:
            return None
        from oracle._ref import _particlestogrid
        return _particlestogrid
    except Exception:
        return None
