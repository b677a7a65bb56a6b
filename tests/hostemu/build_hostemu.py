"""Builds tests/hostemu/libhostemu.so: the device pipeline header compiled as plain C++ for the
CPU-side unit tests (see hostemu.cc). Needs g++ and the reference headers; test-only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "mujoco_inversedynamicstest_b200", "csrc")
LIB = os.path.join(HERE, "libhostemu.so")


def include_dir():
    for cand in (os.environ.get("MUJOCO_INCLUDE"), "/root/reference/include"):
        if cand and os.path.exists(os.path.join(cand, "mujoco", "mujoco.h")):
            return cand
    return None


def build(force=False):
    inc = include_dir()
    srcs = [os.path.join(HERE, "hostemu.cc"), os.path.join(CSRC, "mjb_upload.cc")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    if os.path.exists(LIB) and not force:
        if inc is None or os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
            return LIB
    if inc is None:
        raise RuntimeError("reference headers not found and no prebuilt libhostemu.so")
    # -ffp-contract=off: same arithmetic as the CPU reference build, so predicates agree exactly
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I" + CSRC,
           "-I" + inc] + srcs + ["-o", LIB]
    subprocess.run(cmd, check=True)
    return LIB
