import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge

    return ge.load_package()


@pytest.fixture(scope="session")
def oracle():
    import __graft_entry__ as ge

    O = ge.load_oracle()
    O.build()
    return O


@pytest.fixture(scope="session")
def grid11(pkg):
    """1x1 grid on the current CUDA device (GPU tests only)."""
    pkg.initialize()
    ctx = pkg.create_grid(None, 1, 1, "R")
    yield ctx
    pkg.free_grid(ctx)
