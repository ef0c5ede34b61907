import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pmv():
    import pmv_b200
    return pmv_b200


@pytest.fixture(scope="session")
def ctx(pmv):
    """One device context for the GPU tests.  Fails loudly when the library or GPU is missing."""
    c = pmv.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def synth(pmv):
    from harness import synth as s
    return s
