import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests" / "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    return o


@pytest.fixture(scope="session")
def mm():
    """The product package; on a GPU box also asserts the device is usable."""
    import mmrs_b200
    return mmrs_b200
