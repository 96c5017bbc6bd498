import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")
    config.addinivalue_line("markers", "ref: needs the compiled reference in oracle/_ref")


@pytest.fixture(scope="session")
def lib():
    """The product library must be present and export the whole C ABI (built by __graft_entry__.build())."""
    from source_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _cabi.load()


@pytest.fixture(scope="session")
def reference():
    from oracle import harness
    if not harness.available():
        pytest.skip("oracle/_ref not built (needs /root/reference): golden fixtures cover this case")
    return harness


@pytest.fixture(scope="session")
def device(lib):
    from source_b200.engine import Device
    return Device(0)
