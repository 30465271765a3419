import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    import oracle
    return oracle.oracle()


@pytest.fixture(scope="session")
def ref():
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libref_shim.so not built (needs /root/reference at build time)")
    return oracle.reference()


@pytest.fixture(scope="session")
def ctx():
    import rattle_b200
    c = rattle_b200.Context(0)
    yield c
    c.close()
