import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """The in-tree C-ABI library and the oracle, built if missing (no GPU needed to build)."""
    from colate_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    from oracle import pyoracle as po
    po.lib()
    return True


@pytest.fixture(scope="session")
def handle(built):
    from colate_b200 import api
    h = api.Handle(0)
    yield h
    h.close()
