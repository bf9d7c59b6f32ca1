import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """oracle/oracle.py -- the checker.  Only tests may use it."""
    import __graft_entry__ as ge
    return ge.load_oracle()


@pytest.fixture(scope="session")
def port(orc):
    if not os.path.exists(orc.PORT_SO):
        pytest.skip("oracle/_ref/liboracle_port.so not built (run __graft_entry__.build())")
    return orc.Port()


@pytest.fixture(scope="session")
def capi():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as ge
    return ge.load_package().capi


@pytest.fixture(scope="session")
def have_ref(orc):
    return os.path.exists(orc.REF_SO)
