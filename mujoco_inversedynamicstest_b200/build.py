"""In-tree build of libmjb.so (hand-written sm_100a kernels + C-ABI) with nvcc.

    python -m mujoco_inversedynamicstest_b200.build [--force]

The reference's public headers (<mujoco/mujoco.h>) are an include dependency of the drop-in
boundary (mjb.h takes `const mjModel*`); they are taken from $MUJOCO_INCLUDE or
/root/reference/include and are never copied into this repository. On a box without them (the GPU
box) the prebuilt library that travelled with the snapshot is used as is.
"""
import os
import shutil
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmjb.so")
OBJ_DIR = os.path.join(_ROOT, "build", "mjb")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
SOURCES = ["mjb_kernels.cu", "mjb_api.cu", "mjb_upload.cc", "mjb_modelio.cc"]


def mujoco_include():
    for cand in (os.environ.get("MUJOCO_INCLUDE"), "/root/reference/include"):
        if cand and os.path.exists(os.path.join(cand, "mujoco", "mujoco.h")):
            return cand
    return None


def _newest_source_mtime():
    t = 0.0
    for d in (CSRC, os.path.join(_ROOT, "include")):
        for f in os.listdir(d):
            t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def up_to_date():
    return os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime()


def build(force=False, verbose=False):
    """Compile libmjb.so if sources are newer than the library. Returns the library path."""
    inc = mujoco_include()
    if inc is None:
        if os.path.exists(LIB_PATH):
            return LIB_PATH          # prebuilt library, headers not present on this box
        raise RuntimeError("MuJoCo headers not found (set MUJOCO_INCLUDE) and no prebuilt libmjb.so")
    if not force and up_to_date():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden",
              "-I" + inc, "-I" + os.path.join(_ROOT, "include"), "-I" + CSRC]
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        cmd = [nvcc] + ARCH + common + ["-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                            text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc] + ARCH + ["-shared", "-o", LIB_PATH] + objs + ["-cudart", "static"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(OBJ_DIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
