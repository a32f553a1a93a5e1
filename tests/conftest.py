import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): builds oracle/build/liboracle.so on demand."""
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def engine():
    """One engine on cuda:0, through the C ABI.  Fails loudly when the .so or the GPU is missing."""
    from sequencealigning_b200 import Engine
    from sequencealigning_b200.build import build_all
    build_all()
    eng = Engine(0)
    yield eng
    eng.close()
