import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def checkers():
    """Build oracle/liboracle.so (+ oracle/_ref when /root/reference exists)."""
    import _bind
    _bind.build_checkers()
    return _bind


@pytest.fixture(scope="session")
def oracle(checkers):
    return checkers.Oracle()


@pytest.fixture(scope="session")
def reference(checkers):
    if not checkers.have_reference("det"):
        pytest.skip("oracle/_ref not built (no /root/reference and no prebuilt .so)")
    return checkers.Reference("det")


@pytest.fixture(scope="session")
def product():
    """The CUDA product library; built in-tree if missing.  Never falls back to a CPU path."""
    from youtokentome_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()
