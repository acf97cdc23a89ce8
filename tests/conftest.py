import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# cluster-id parity with the reference needs its serial order
os.environ.setdefault("OMP_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ctx():
    """One rtr_ctx for the whole GPU session; building the library is part of the fixture."""
    from realtimeraytracing_b200 import build, capi
    build.build()
    c = capi.Context(0)
    yield c
    c.close()
