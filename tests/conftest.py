import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))  # shared helpers (real_export.py)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture(scope="session")
def native_lib():
    """The engine's shared library, built on demand (nvcc cross-compiles without a GPU)."""
    from smelter_b200 import build as engine_build
    from smelter_b200 import _lib

    engine_build.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def host_oracle():
    """oracle/host_oracle.c as a ctypes library (test infrastructure only)."""
    import ctypes

    from oracle import build as oracle_build

    return ctypes.CDLL(oracle_build.build())


@pytest.fixture(scope="session")
def ctx(native_lib):
    from smelter_b200.api import Context

    c = Context(0)
    yield c
    c.close()
