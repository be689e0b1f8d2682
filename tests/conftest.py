import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure). Built on demand with oracle/Makefile."""
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def vmp_lib():
    """The CUDA library; the gpu tests call through its C ABI only."""
    from voxelmapplus_fastlio2_b200 import bindings
    return bindings.load_library()
