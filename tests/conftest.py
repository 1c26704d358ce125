import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build them once, like the driver's build()."""
    need = [os.path.join(ROOT, "hyperex_b200", "_lib", "libhyperex_b200.so"), os.path.join(ROOT, "bin", "hyperex"),
            os.path.join(ROOT, "oracle", "_build", "libhxoracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def ctx():
    """one hx_ctx on cuda:0 for the whole GPU session; raises (never falls back) without a GPU"""
    import hyperex_b200 as hx
    c = hx.Context((0,))
    yield c
    c.close()
