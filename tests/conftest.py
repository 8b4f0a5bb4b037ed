import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both shared libraries must exist; building is the job of __graft_entry__.build()."""
    import simmr_b200
    from oracle import pyoracle
    if not os.path.exists(simmr_b200._native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    pyoracle.lib()


@pytest.fixture(scope="session")
def engine():
    import simmr_b200 as S
    eng = S.Engine(0)       # raises loudly when there is no GPU / no library
    yield eng
    eng.close()
