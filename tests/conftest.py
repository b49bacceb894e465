import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mm-unet_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    cases = {}
    for k in z.files:
        case, _, field = k.partition(".")
        cases.setdefault(case, {})[field] = z[k]
    return cases


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(autouse=True)
def _mmu_knobs(monkeypatch):
    """The library reads its MMU_* knobs once: reload them at the start of every test (the previous test's environment has
    been restored by then) and whenever a test changes one through monkeypatch."""
    try:
        from mmunet_b200 import _lib
        reload = _lib.reload_knobs
        reload()
    except Exception:       # library not built: the tests that need it fail on their own
        yield
        return
    setenv, delenv = monkeypatch.setenv, monkeypatch.delenv

    def setenv_(name, value, *a, **k):
        setenv(name, value, *a, **k)
        if name.startswith("MMU_"):
            reload()

    def delenv_(name, *a, **k):
        delenv(name, *a, **k)
        if name.startswith("MMU_"):
            reload()

    monkeypatch.setenv, monkeypatch.delenv = setenv_, delenv_
    yield
