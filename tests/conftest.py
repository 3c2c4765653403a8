import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib_built():
    """Build (if needed) the CUDA library and the oracle once per session."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture()
def ctx(lib_built):
    import mosaic_b200
    c = mosaic_b200.Context(0)
    yield c
    c.close()
