"""The compiled-C drop-in test (tests/cabi/dropin.c): one process links the reference library and
libmjb.so, loads the model with the reference's own mj_loadModel, runs the mj_inverse loop and
mjb_inverse side by side and compares at the strict bound; also m->opt changes between calls and
the refusal of active global callbacks (include/mujoco/mujoco.h:131-137,
src/inverse/inverse_test.cpp:43-112, src/engine/engine_callback.c:21-28)."""
import gzip
import os
import subprocess
import tempfile

import pytest

import util

pytestmark = pytest.mark.gpu

BIN = os.path.join(util.ROOT, "tests", "cabi", "_bin", "dropin")


@pytest.mark.skipif(not os.path.exists(BIN), reason="tests/cabi/_bin/dropin not built (python tests/cabi/build_dropin.py)")
# weld: components that cancel between large weld forces (util.STRICT_EXCEPTIONS); a few entries allowed
@pytest.mark.parametrize("name,n,allowed", [("humanoid", 2048, 0), ("humanoid_nocontact", 512, 0), ("zoo", 512, 0),
                                            ("weld", 256, 16)])
def test_compiled_c_dropin(name, n, allowed):
    with tempfile.NamedTemporaryFile(suffix=".mjb", delete=False) as tf:
        tf.write(gzip.open(os.path.join(util.GOLDEN, name + ".mjb.gz"), "rb").read())
    try:
        r = subprocess.run([BIN, tf.name, str(n), str(allowed)], capture_output=True, text=True, timeout=600)
    finally:
        os.remove(tf.name)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "DROPIN OK" in r.stdout, r.stdout + r.stderr
