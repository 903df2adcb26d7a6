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
def oracle():
    """The CPU restatement (test infrastructure)."""
    from tests import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def capi():
    from phylommand_b200 import build, capi as _capi
    build.build_library()
    _capi.load()
    return _capi


@pytest.fixture(scope="session")
def gpu(capi):
    """Initialised CUDA context; the product path has no CPU fallback, so a missing device is an error."""
    capi.init()
    yield capi
    capi.shutdown()
