import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "reference: needs the reference checkout at /root/reference")


@pytest.fixture
def sim(monkeypatch):
    """Swap the CUDA primitives for their torch semantic models (host-logic tests on CPU)."""
    from tests import sim_backend
    sim_backend.install(monkeypatch)
    return sim_backend
