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
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def lib():
    import tron_b200
    from tron_b200 import build
    build.build()
    return tron_b200.load_library()


@pytest.fixture(scope="session")
def reflib():
    """The unmodified reference compiled in place (GPU only)."""
    from oracle.oracle import RefLib
    try:
        return RefLib()
    except (FileNotFoundError, OSError) as e:
        pytest.skip("oracle/_ref not built: %s" % e)


@pytest.fixture(scope="session")
def reflib_wide():
    from oracle.oracle import RefLib
    try:
        return RefLib(widened=True)
    except (FileNotFoundError, OSError) as e:
        pytest.skip("oracle/_ref not built: %s" % e)
