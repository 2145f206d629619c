"""examples/c_host.c: a plain-C99 consumer of the C-ABI (no Python, no torch types).  CPU: it compiles against include/marl_b200.h
with gcc -std=c99 -Wall -Werror and links against the nvcc-built library.  GPU (tests/test_gpu_zz_late_additions.py): it runs the
sample -> train -> soft-update loop."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
LIBDIR = os.path.join(ROOT, "off-policy_b200", "lib")


def build_c_host(out):
    if not os.path.exists(os.path.join(LIBDIR, "libmarl_b200.so")):
        pytest.skip("libmarl_b200.so not built (python __graft_entry__.py build)")
    if shutil.which("gcc") is None or not os.path.isdir(os.path.join(CUDA, "include")):
        pytest.skip("gcc / CUDA headers not available")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"),
           os.path.join(ROOT, "examples", "c_host.c"), "-L" + LIBDIR, "-lmarl_b200", "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + LIBDIR, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_c_host_example_compiles_and_links_as_c99(tmp_path):
    build_c_host(str(tmp_path / "c_host"))
