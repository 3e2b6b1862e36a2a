import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture
def fake_device():
    """Host-logic tests: the C-ABI replaced by the numpy test double in tests/fake_device.py."""
    from tests import fake_device as fd
    dev, undo = fd.install()
    try:
        yield dev
    finally:
        undo()


@pytest.fixture(scope="session")
def cuda():
    """The real device backend; fails (does not skip) when the library or the GPU is missing."""
    from lightgrad_b200.autograd.cuda import runtime as rt
    rt.ensure_device()
    return rt
