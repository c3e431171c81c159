import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # (pytest-timeout registers this one itself where it is installed; without the plug-in the marker is simply inert)
    config.addinivalue_line("markers", "timeout(seconds): upper bound for one test (pytest-timeout)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    import __graft_entry__ as ge
    return ge.load_oracle()


@pytest.fixture(scope="session")
def wb():
    """The product package; on a GPU box the library must load and find an sm_100 device (no fallback)."""
    import __graft_entry__ as ge
    pkg = ge.load_package()
    return pkg


@pytest.fixture(scope="session")
def gpu(wb):
    wb.init(0)
    return wb
