#!/usr/bin/env python
"""Fill the model-specialisation cache (mjb_precompile: NVRTC, no GPU needed) for committed models.

    python tools/precompile.py [--torch] [name ...]        # names of tests/golden/*.mjb.gz; default: all

--torch: import torch first. A process that has torch loaded resolves libnvrtc to the copy bundled with
torch (another NVRTC version, hence another cache key) -- bench.py is such a process, the tests are not.
"""
import ctypes
import glob
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(name):
    if os.environ.get("MJB_PRECOMPILE_TORCH"):
        import torch  # noqa: F401
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200._lib import lib
    m = mjb.Model.from_mjb(os.path.join(ROOT, "tests", "golden", name + ".mjb.gz"))
    info = ctypes.create_string_buffer(8192)
    rc = lib().mjb_precompile(m.ptr, info, 8192)
    return name, rc, info.value.decode()


def main(names):
    if "--torch" in names:
        names = [n for n in names if n != "--torch"]
        os.environ["MJB_PRECOMPILE_TORCH"] = "1"
    if not names:
        names = sorted(os.path.basename(p)[:-7] for p in glob.glob(os.path.join(ROOT, "tests", "golden", "*.mjb.gz")))
    bad = 0
    with ProcessPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for name, rc, info in ex.map(one, names):
            print(f"{name:28s} {'ok ' if rc == 0 else 'skip'} {info.splitlines()[0] if info else ''}")
            bad += rc != 0 and "too large" not in info
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
