import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def pcreg():
    """The initialised product library.  GPU tests only: fails loudly (no skip, no fallback) when the
    CUDA library is missing or no device is usable."""
    import pcreg_b200
    pcreg_b200.init()
    return pcreg_b200
