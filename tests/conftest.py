import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CUDA library (cross-compiles without a GPU) and the oracle once per session."""
    import __graft_entry__ as graft

    # always: build() returns at once when nothing is newer than the libraries (build.needs_build, make's own dependency
    # check), and a stale library would let the suites pass without exercising an edited kernel
    graft.build()
    graft.load_package()
    yield
