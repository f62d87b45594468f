import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def vr_ctx():
    """One CUDA context for the whole GPU test session.  Fails loudly (no skip, no CPU fallback) when the
    extension or the device is missing."""
    from cl_volume_renderer_b200 import api
    ctx = api.Context(0)
    yield ctx
    ctx.close()
