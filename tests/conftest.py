import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a device: skip them (instead of failing in Pool()) where there is none."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Native artefacts are built once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def hostemu_lib():
    """The tick kernel's row body compiled by g++ (tests/hostemu) — CPU debugging aid only."""
    from consul_b200 import _lib
    return _lib.load(os.path.join(ROOT, "tests", "hostemu", "libgsim_hostemu.so"))


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library.  Pool creation raises without a B200: no fallback."""
    from consul_b200 import _lib
    return _lib.lib()
