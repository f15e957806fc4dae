import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import helpers
    return helpers.load_oracle()


@pytest.fixture(scope="session")
def emu():
    import helpers
    return helpers.load_emu()


@pytest.fixture(scope="session")
def ref():
    import helpers
    lib = helpers.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    return lib
