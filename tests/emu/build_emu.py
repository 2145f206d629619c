"""Build the CPU-emulated unit-test library (tests only; see emu_runtime.h)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "off-policy_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libmarl_b200_emu.so")


def build(verbose=False):
    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("emu_runtime.h", "emu_runtime.cpp")] + \
           [os.path.join(ROOT, "include", "marl_b200.h")]
    newest = max(os.path.getmtime(d) for d in deps)
    objs = []
    procs = []
    for f in srcs + ["emu_runtime.cpp"]:
        src = os.path.join(CSRC if f.endswith(".cu") else HERE, f)
        obj = os.path.join(OUT, f + ".o")
        objs.append(obj)
        if os.path.exists(obj) and os.path.getmtime(obj) > newest:
            continue
        cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-DMARL_EMU", "-I", HERE, "-I", CSRC, "-x", "c++", "-c", src, "-o", obj,
               "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas", "-Wno-sign-compare", "-Wno-unused-variable"]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    fail = False
    for f, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            fail = True
            sys.stderr.write("== %s ==\n%s\n" % (f, out))
        elif verbose and out.strip():
            sys.stderr.write("== %s ==\n%s\n" % (f, out))
    if fail:
        raise RuntimeError("emu build failed")
    if procs or not os.path.exists(LIB):
        subprocess.check_call(["g++", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
