"""Pins oracle/mjinv_oracle.c (the plain-C restatement in the reference's own dense-Jacobian
formulation) against the dumps of the unmodified reference engine (tests/golden). Bit-exact on
every discrete output; the continuous ones agree to rounding of a different-but-equivalent
summation order (both builds use -ffp-contract=off)."""
import numpy as np
import pytest

import util


def _bind(name):
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    from oracle import restatement
    path, ref = util.golden(name)
    model = mjb.Model.from_mjb(path)
    ptr = ctypes_opt(model)
    r = restatement.Restatement(model, ptr)
    n = int(ref["nstate"])
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(ref["z_range"]))
    return r, ref, qpos, qvel, qacc


def ctypes_opt(model):
    from mujoco_inversedynamicstest_b200._lib import lib
    import ctypes
    n = ctypes.c_int()
    def num(name):
        p = lib().mjb_modelOptNum(model.ptr, name.encode(), ctypes.byref(n))
        return [p[i] for i in range(n.value)]
    return {"disableflags": model.get_opt_int("disableflags"), "cone": model.get_opt_int("cone"),
            "timestep": num("timestep")[0], "impratio": num("impratio")[0], "gravity": num("gravity")}


@pytest.mark.parametrize("name", ["humanoid", "humanoid_elliptic", "humanoid_nocontact",
                                  "slider_crank_nocontact", "inverse_test"])
def test_restatement_matches_reference_dump(name):
    r, ref, qpos, qvel, qacc = _bind(name)
    out = r.inverse_batch(qpos, qvel, qacc, maxcon=int(ref["nconmax"]), maxefc=int(ref["njmax"]))
    for k in ("ncon", "ne", "nf", "nl", "nefc"):
        np.testing.assert_array_equal(out[k], ref[k], err_msg=k)
    np.testing.assert_array_equal(out["contact_geom"], ref["contact_geom"])
    np.testing.assert_array_equal(out["efc_type"], ref["efc_type"])
    np.testing.assert_array_equal(out["efc_id"], ref["efc_id"])
    nviol, worst = util.qfrc_violations_scaled(out["qfrc_inverse"], ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)
    scale = max(1.0, np.abs(ref["efc_force"]).max())
    np.testing.assert_allclose(out["efc_force"], ref["efc_force"], rtol=1e-9, atol=1e-13 * scale)
    np.testing.assert_array_equal(out["qM"], ref["qM"])
    np.testing.assert_allclose(out["qLD"], ref["qLD"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(out["qLDiagInv"], ref["qLDiagInv"], rtol=1e-12, atol=1e-13)


def test_restatement_matches_reference_on_22_humanoids():
    r, ref, qpos, qvel, qacc = _bind("humanoids22")
    out = r.inverse_batch(qpos[:2], qvel[:2], qacc[:2], maxcon=int(ref["nconmax"]),
                          maxefc=int(ref["njmax"]), inertia=False)
    np.testing.assert_array_equal(out["ncon"], ref["ncon"][:2])
    np.testing.assert_array_equal(out["nefc"], ref["nefc"][:2])
    np.testing.assert_array_equal(out["contact_geom"], ref["contact_geom"][:2])
    np.testing.assert_array_equal(out["efc_type"], ref["efc_type"][:2])
    nviol, worst = util.qfrc_violations_scaled(out["qfrc_inverse"], ref["qfrc_inverse"][:2])
    assert nviol == 0, (nviol, worst)
