import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _build_once():
    import __graft_entry__ as ge
    ge.build_cpu_side()


@pytest.fixture(scope="session")
def built():
    _build_once()
    return True


@pytest.fixture(scope="session")
def port(built):
    from oracle.oracle import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def ref(built):
    """The reference's own lz4.c if its build is present, else the restatement
    (which test_oracle.py pins against it wherever the reference build exists)."""
    from oracle.oracle import Oracle, available
    return Oracle("reference" if available("reference") else "port")


@pytest.fixture(scope="session")
def ctx(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from streamly_lz4_b200 import Context
    c = Context(0)
    yield c
    c.close()
