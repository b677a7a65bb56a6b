"""The checker itself: oracle/_ref is the UNMODIFIED reference engine compiled from /root/reference.
These tests pin it against the reference's own test properties and show that the
committed golden fixtures are exactly what it produces."""
import gzip

import numpy as np
import pytest

import util

pytestmark = pytest.mark.skipif(not util.ref_available(), reason="oracle/_ref not built")


def _ref_model(name, tmp_path):
    from oracle import reflib
    raw = tmp_path / (name + ".mjb")
    raw.write_bytes(gzip.open(util.golden(name)[0], "rb").read())
    return reflib.Model.from_mjb(str(raw))


@pytest.mark.parametrize("name", ["humanoid", "humanoid_elliptic", "humanoid_nocontact",
                                  "slider_crank_nocontact", "inverse_test"])
def test_golden_fixtures_are_the_reference_output(name, tmp_path):
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref = util.golden(name)
    rm = _ref_model(name, tmp_path)
    n = int(ref["nstate"])
    qpos, qvel, qacc = generate_states(rm, n, z_range=tuple(ref["z_range"]))
    out, _ = rm.inverse_batch(qpos, qvel, qacc, fields={
        "ncon": 1, "nefc": 1, "contact_geom": int(ref["nconmax"]), "efc_type": int(ref["njmax"])})
    np.testing.assert_array_equal(out["qfrc_inverse"], ref["qfrc_inverse"])   # bit-exact
    np.testing.assert_array_equal(out["ncon"], ref["ncon"])
    np.testing.assert_array_equal(out["nefc"], ref["nefc"])
    np.testing.assert_array_equal(out["contact_geom"], ref["contact_geom"])
    np.testing.assert_array_equal(out["efc_type"][:, :, 0], ref["efc_type"])


def test_thread_pool_loop_is_bit_identical_to_single_thread(tmp_path):
    """test/engine/engine_thread_test.cc:36-88 bar: threading must not change a single bit."""
    from mujoco_inversedynamicstest_b200.states import generate_states
    rm = _ref_model("humanoid", tmp_path)
    qpos, qvel, qacc = generate_states(rm, 600)
    a, _ = rm.inverse_batch(qpos, qvel, qacc, nthread=1)
    b, _ = rm.inverse_batch(qpos, qvel, qacc, nthread=4)
    np.testing.assert_array_equal(a["qfrc_inverse"], b["qfrc_inverse"])


def test_forward_inverse_match(tmp_path):
    """test/engine/engine_inverse_test.cc:35-56 restated on humanoid: after 70 steps,
    mj_compareFwdInv's two discrepancy norms stay < 1e-10 relative to the force scale."""
    rm = _ref_model("humanoid", tmp_path)
    qpos0 = rm.array("qpos0").ravel().copy()
    res = rm.compare_fwdinv(qpos0, np.zeros(rm.int("nv")), nstep=70)
    assert res[0] < 1e-6 and res[1] < 1e-6   # absolute norms; forces are O(1e3) here
    rm2 = _ref_model("inverse_test", tmp_path)
    res2 = rm2.compare_fwdinv(rm2.array("qpos0").ravel().copy(), np.zeros(rm2.int("nv")), nstep=70)
    assert res2[0] < 1e-10 and res2[1] < 1e-10


def test_ldl_equals_m_on_reference(tmp_path):
    """test/engine/engine_core_smooth_test.cc:466-511: L' D L == M to 1e-12 (relative here)."""
    from mujoco_inversedynamicstest_b200.states import generate_states
    rm = _ref_model("humanoid", tmp_path)
    nv = rm.int("nv")
    qpos, qvel, qacc = generate_states(rm, 4)
    out, _ = rm.inverse_batch(qpos, qvel, qacc, fields={"qM": None, "qLD": None})
    parent = rm.array("dof_parentid").ravel()
    madr = rm.array("dof_Madr").ravel()
    for s in range(4):
        qM, qLD = out["qM"][s, :, 0], out["qLD"][s, :, 0]
        M = np.zeros((nv, nv)); L = np.eye(nv); D = np.zeros(nv)
        adr = 0
        for i in range(nv):
            chain = []
            j = i
            while j >= 0:
                chain.append(j); j = parent[j]
            for t, j in enumerate(chain):
                M[i, j] = M[j, i] = qM[madr[i] + t]
            for t, j in enumerate(chain[::-1]):
                if j == i: D[i] = qLD[adr + t]
                else: L[i, j] = qLD[adr + t]
            adr += len(chain)
        np.testing.assert_allclose(L.T @ np.diag(D) @ L, M, rtol=1e-11, atol=1e-11)
