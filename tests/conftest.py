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
        # a kernel that never finishes must fail the run, not hang the box: per-test limit enforced from a watchdog thread (the main
        # thread may be blocked inside a CUDA call, where a signal would not be delivered); on expiry pytest-timeout dumps the stacks
        # and exits the process, which tears the CUDA context down
        try:
            import pytest_timeout  # noqa: F401
            for it in items:
                if "gpu" in it.keywords and not any(m.name == "timeout" for m in it.iter_markers()):
                    it.add_marker(pytest.mark.timeout(600, method="thread"))
        except Exception:
            pass
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
