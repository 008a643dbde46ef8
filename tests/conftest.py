import os
import sys
from pathlib import Path

import pytest

os.environ.setdefault("OMP_NUM_THREADS", "4")
ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_addoption(parser):
    parser.addoption("--runslow", action="store_true", default=False, help="run tests that recompute base flows")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU test, only with --runslow")


def pytest_collection_modifyitems(config, items):
    if config.getoption("--runslow"):
        return
    skip = pytest.mark.skip(reason="needs --runslow")
    for item in items:
        if "slow" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def root():
    return ROOT


@pytest.fixture(scope="session")
def built_lib():
    """Make sure libfcb200.so exists (compiled in-tree by __graft_entry__.build)."""
    import __graft_entry__ as g

    g.build()
    from flowcontrol_b200 import libfcb

    return libfcb.load()
