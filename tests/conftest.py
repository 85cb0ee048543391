import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def ctx():
    from himut_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()
