import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; compiled on demand with gcc)."""
    from oracle import kbo
    kbo.build()
    return kbo


@pytest.fixture(scope="session")
def native():
    """The product library; GPU tests fail loudly if it is missing (no fallback)."""
    from gym_kilobots_b200 import _native
    _native.load()
    return _native
