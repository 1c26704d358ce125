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


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped (a plain `pytest` on a CPU box must not fail in them);
    on a GPU box nothing is skipped here, and a broken device path fails loudly in the tests themselves."""
    if not any("gpu" in it.keywords for it in items):
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (there is no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


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
