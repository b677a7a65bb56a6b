"""The C-ABI shared library loads and exports every entry point declared in include/*.h.
No compute call is made: this runs on a box without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


def declared_symbols():
    names = []
    for fn in sorted(os.listdir(INCLUDE)):
        text = open(os.path.join(INCLUDE, fn)).read()
        names += re.findall(r"MJB_API\s+[\w\s\*]+?\b(mjb_\w+)\s*\(", text)
    return sorted(set(names))


def test_headers_declare_the_expected_boundary():
    names = declared_symbols()
    for required in ("mjb_makeData", "mjb_deleteData", "mjb_setState", "mjb_inverse",
                     "mjb_getQfrcInverse", "mjb_loadModel"):
        assert required in names
    assert len(names) >= 25


def test_library_exports_every_declared_symbol():
    from mujoco_inversedynamicstest_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libmjb.so is not built"
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(L, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_python_binding_covers_every_declared_symbol():
    from mujoco_inversedynamicstest_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.lib()


def test_kernels_are_compiled_for_sm_100a_with_tma():
    """cuobjdump: the fused kernel exists for sm_100a and stages the model with a TMA bulk copy."""
    import shutil
    import subprocess
    from mujoco_inversedynamicstest_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "smooth_kernel" in sass and "contact_kernel" in sass and "backward_kernel" in sass
    assert "UBLKCP" in sass, "model blob staging is not a TMA bulk copy"
    assert "DFMA" in sass


def test_no_cpu_fallback_without_a_device():
    """On a box without CUDA the product refuses to run instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import mujoco_inversedynamicstest_b200 as mjb
    model = mjb.Model.from_mjb(os.path.join(ROOT, "tests", "golden", "humanoid.mjb.gz"))
    with pytest.raises(mjb.MjbError, match="no CUDA device"):
        mjb.BatchData(model, 4)


def test_product_never_imports_the_oracle():
    """The oracle and the host emulation are test infrastructure: no file of the product package
    may import, link or open them."""
    pkg = os.path.join(ROOT, "mujoco_inversedynamicstest_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".c")):
                continue
            text = open(os.path.join(dirpath, f), errors="ignore").read()
            for needle in ("import oracle", "from oracle", "libmujoco_ref", "reflib", "libhostemu",
                           "hostemu_inverse"):
                assert needle not in text, f"{f} mentions {needle}"
