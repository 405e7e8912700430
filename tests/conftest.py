import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA B200 device (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One qcp2 context on cuda:0.  No skip, no fallback: without the CUDA library and a
    device this raises, which is the behaviour the product path must have."""
    import qcp2_b200
    c = qcp2_b200.Context(0)
    yield c
    c.close()
