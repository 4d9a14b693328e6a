import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
os.environ.setdefault("FITV2_POISON_WORKSPACE", "1")     # scratch memory starts as NaN whenever the (rows, tokens) layout changes


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; parity tests through the C ABI")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and return the path of the in-tree C-ABI library."""
    import __graft_entry__ as g
    g.build()
    return g.OUT
