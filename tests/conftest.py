import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_usable():
    """True when a context can be created (a B200 is visible and the library is built)."""
    try:
        import mosaic_b200
        mosaic_b200.Context(0).close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests/` on a box without a GPU skips the GPU tests instead of failing them (the driver's own
    runs select with -m gpu / -m "not gpu"; there nothing is skipped that could have run)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _gpu_usable():
        return
    skip = pytest.mark.skip(reason="needs a B200 (mosaic_ctx_create failed)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def lib_built():
    """Build (if needed) the CUDA library and the oracle once per session."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture()
def ctx(lib_built):
    import mosaic_b200
    try:
        c = mosaic_b200.Context(0)
    except mosaic_b200.MosaicError as e:
        if e.code == -2:
            pytest.skip("no usable CUDA device")
        raise
    yield c
    c.close()
