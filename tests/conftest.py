import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "off-policy_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="module")
def emu_engine():
    """CPU fiber-emulated build of the CUDA kernels (tests/emu) bound into the Python mirror: kernel-logic tests only."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    from build_emu import build
    from offpolicy._b200 import capi
    path = build()
    capi._install_for_tests(path)
    yield capi
    capi._uninstall_for_tests()


@pytest.fixture(scope="module")
def gpu_engine():
    """The real nvcc-built library on cuda:0 (fails loudly if it is missing)."""
    from offpolicy._b200 import capi
    capi._uninstall_for_tests()
    capi.lib()
    yield capi
