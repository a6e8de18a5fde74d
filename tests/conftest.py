import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with `-m gpu` on a B200)')


@pytest.fixture
def torch_backend():
    """Route cgc_net_b200.ops through the per-operator CPU formulas (tests only)."""
    from cgc_net_b200 import _backend
    from oracle import op_reference
    prev = _backend._active
    _backend.use_backend_for_testing(op_reference.TorchBackend())
    yield
    _backend.use_backend_for_testing(prev)


@pytest.fixture
def cuda_backend():
    import torch
    from cgc_net_b200 import _backend
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    prev = _backend._active
    _backend.use_backend_for_testing(None)
    be = _backend.get()          # raises if libcgc_sm100.so is missing: no fallback
    yield be
    _backend.use_backend_for_testing(prev)
