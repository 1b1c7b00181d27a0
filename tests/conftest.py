import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dinov3-jax_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def native():
    """The C-ABI library bound to cuda:0 — GPU tests fail loudly (no fallback) when it is missing."""
    import torch
    from dinov3_jax import _native
    assert torch.cuda.is_available(), "GPU test selected without a GPU"
    _native.init(0)
    return _native


GOLDEN = os.path.join(ROOT, "tests", "golden")
