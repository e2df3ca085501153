import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree native pieces exist (a fresh checkout has none of the git-ignored .so files)."""
    from misaki_render_b200 import capi
    from oracle import pyoracle
    from misaki_render_b200 import host_api
    need = [capi.LIB_PATH, host_api.LIB_PATH, pyoracle.LIB, ROOT / "misaki_render_b200" / "data" / "srgb.coeff"]
    if not all(p.exists() for p in need):
        import __graft_entry__
        __graft_entry__.build()
    yield


@pytest.fixture(scope="session")
def gpu_ctx():
    from misaki_render_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()
