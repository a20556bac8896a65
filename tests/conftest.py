import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    if os.environ.get("PHB200_TEST_HOST_EMUL") == "1":
        # TEST INFRASTRUCTURE: exercise the `-m gpu` suite where there is no GPU, against the whole product library
        # compiled for the host (tests/host_emul/fullhost: fake CUDA runtime, one fiber per CUDA thread).  Only the
        # test process is redirected; phasta_b200/ knows nothing about it and has no CPU fallback.
        sys.path.insert(0, os.path.join(ROOT, "tests", "host_emul"))
        from build_fullhost import build
        import phasta_b200.lib as lib
        lib.LIB_PATH = build()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip silently;
    # a plain run without -m skips GPU tests when there is no device.
    if config.getoption("-m"):
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
