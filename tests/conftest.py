import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def gpu_ctx():
    """One CUDA context for the whole GPU test session (fails loudly without the library / a GPU)."""
    from distgcn_b200 import engine
    ctx = engine.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(autouse=True)
def _library_env_follows_monkeypatch(monkeypatch):
    """The library reads its DG_* options once per context (never on the solve path): every setenv / delenv of a test, and
    the start of every test (the previous test's changes have been undone by then), is followed by a reload on the live
    contexts."""
    import sys
    eng = sys.modules.get("distgcn_b200.engine")
    if eng is not None:
        eng.reload_env()
    orig_set, orig_del = monkeypatch.setenv, monkeypatch.delenv

    def setenv(name, value, prepend=None):
        orig_set(name, value, prepend)
        e = sys.modules.get("distgcn_b200.engine")
        if e is not None and name.startswith("DG_"):
            e.reload_env()

    def delenv(name, raising=True):
        orig_del(name, raising)
        e = sys.modules.get("distgcn_b200.engine")
        if e is not None and name.startswith("DG_"):
            e.reload_env()
    monkeypatch.setenv, monkeypatch.delenv = setenv, delenv
    yield
