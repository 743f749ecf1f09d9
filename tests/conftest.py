import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def built_lib():
    """libalps_b200.so built in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    from alps_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    return _lib.SO_PATH
