"""pytest configuration: the ``gpu`` marker, import path, and shared fixtures."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def engine():
    """One Engine for the whole GPU session.  No skip-on-missing: a GPU test without the CUDA extension
    (or without a device) must FAIL, never pass on a fallback."""
    import giql_b200

    return giql_b200.Engine(0)


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as co

    co.build()
    return co
