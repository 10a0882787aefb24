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
    """The reference's own lz4.c, compiled by oracle/Makefile into oracle/_ref (prebuilt here, it travels to the GPU
    box).  There is NO silent downgrade: without it the parity tests fail, unless B200LZ4_ALLOW_PORT_ORACLE=1
    explicitly selects the restatement (which test_oracle.py pins against the reference wherever both exist)."""
    from oracle.oracle import Oracle, available
    if available("reference"):
        return Oracle("reference")
    if os.environ.get("B200LZ4_ALLOW_PORT_ORACLE") == "1":
        import warnings
        warnings.warn("oracle/_ref is missing: parity is checked against the RESTATEMENT (oracle/lz4_oracle.c), not the reference")
        return Oracle("port")
    pytest.fail("oracle/_ref/libreflz4.so is missing (build it here with `make -C oracle`; it is git-ignored but travels "
                "with the snapshot).  Set B200LZ4_ALLOW_PORT_ORACLE=1 to run against the restatement instead.")


@pytest.fixture(scope="session")
def ctx(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from streamly_lz4_b200 import Context
    c = Context(0)
    yield c
    c.close()
