import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def backend():
    """The CUDA backend through its C ABI.  No fallback: a missing library or GPU is a failure."""
    import bpperm_b200

    be = bpperm_b200.Backend(0)
    yield be
    be.close()
