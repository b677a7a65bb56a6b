"""GPU parity: the CUDA path, called through the C-ABI (libmjb.so via ctypes), against
  (1) the committed golden dumps of the reference's own CPU mj_inverse (tests/golden), and
  (2) the reference library itself (oracle/_ref) run live on the same seeded inputs,
plus size-independent properties at BASELINE.json's full batch size.

Bars (north_star): bit-exact ncon / contact geom pairs / efc_type / efc_id / counters;
qfrc_inverse within 1e-9 relative + 1e-12 absolute.
"""
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

CASES = ["humanoid", "humanoid_elliptic", "humanoid_nocontact", "humanoids22",
         "slider_crank_nocontact", "inverse_test", "arm26", "weld", "connect", "zoo", "zoo_elliptic",
         "gravcomp", "humanoid_invdiscrete", "capsbox", "capsbox_elliptic", "boxes", "boxes_elliptic", "tendons",
         "sensors", "mocap", "touch", "touch_elliptic", "camlight", "transmission", "sensors2",
         "humanoid_invdiscrete_fast", "implicitfast", "humanoid_invdiscrete_implicit", "implicit", "adhesion",
         "adhesion_elliptic"]


def _run(mjb, name, gold, outmask):
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref = util.golden(name)
    model = mjb.Model.from_mjb(path)
    n = int(ref["nstate"])
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(ref["z_range"]))
    bd = mjb.BatchData(model, n, outmask=outmask, nconmax=int(ref["nconmax"]), njmax=int(ref["njmax"]))
    bd.set_state(qpos, qvel, qacc)
    nbad = bd.inverse()
    return model, bd, ref, nbad, (qpos, qvel, qacc)


@pytest.mark.parametrize("name", CASES)
def test_golden_discrete_outputs_bit_exact(name):
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, name, True, mjb.OUT_COUNTS | mjb.OUT_CONTACT | mjb.OUT_EFC)
    assert nbad == 0
    cnt = bd.counts()
    for k in ("ncon", "ne", "nf", "nl", "nefc"):
        np.testing.assert_array_equal(cnt[k], ref[k], err_msg=k)
    con = bd.contacts()
    np.testing.assert_array_equal(con["geom"], ref["contact_geom"])
    np.testing.assert_array_equal(con["dim"], ref["contact_dim"])
    np.testing.assert_array_equal(con["exclude"], ref["contact_exclude"])
    np.testing.assert_array_equal(con["efc_address"], ref["contact_efc_address"])
    efc = bd.efc()
    np.testing.assert_array_equal(efc["type"], ref["efc_type"])
    np.testing.assert_array_equal(efc["id"], ref["efc_id"])
    np.testing.assert_array_equal(efc["state"], ref["efc_state"])


@pytest.mark.parametrize("name", CASES)
def test_golden_qfrc_inverse_within_tolerance(name):
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, name, True, mjb.OUT_QFRC | mjb.OUT_CONTACT | mjb.OUT_EFC)
    got = bd.qfrc_inverse()
    nviol, worst = util.qfrc_violations_scaled(got, ref["qfrc_inverse"])
    assert nviol == 0, f"{nviol} entries outside 1e-9 rel / 1e-12 abs (worst ratio {worst:.3g})"
    # strict element-wise bound: required where no large constraint forces cancel (no contacts, no
    # equality constraints); elsewhere the state-scaled bound above is the criterion
    nstrict, wstrict = util.qfrc_violations(got, ref["qfrc_inverse"])
    if ref["ncon"].max() == 0 and ref["ne"].max() == 0:
        assert nstrict == 0, f"strict bound: {nstrict} violations, worst {wstrict:.3g}"
    con = bd.contacts()
    np.testing.assert_allclose(con["dist"], ref["contact_dist"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(con["pos"], ref["contact_pos"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(con["frame"], ref["contact_frame"], rtol=1e-9, atol=1e-11)
    efc = bd.efc()
    scale = np.abs(ref["efc_force"]).max()
    np.testing.assert_allclose(efc["force"], ref["efc_force"], rtol=1e-9, atol=1e-12 + 1e-13 * scale)
    np.testing.assert_allclose(efc["pos"], ref["efc_pos"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(bd.get(mjb.F_QFRC_PASSIVE), ref["qfrc_passive"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", ["humanoid", "humanoids22"])
def test_golden_inertia_outputs(name):
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, name, True, mjb.OUT_INERTIA | mjb.OUT_INTERNAL)
    for f, k in ((mjb.F_QM, "qM"), (mjb.F_QLD, "qLD"), (mjb.F_QLDIAGINV, "qLDiagInv")):
        np.testing.assert_allclose(bd.get(f), ref[k], rtol=1e-9, atol=1e-12, err_msg=k)
    n = int(ref["nstate"])
    np.testing.assert_allclose(bd.internal("xpos").reshape(n, -1, 3), ref["xpos"], rtol=1e-12, atol=1e-13)
    # spatial vectors live about the tree origin here, about the tree's centre of mass in the
    # reference: compare after the change of origin
    nb = model.int("nbody")
    frames = (bd.internal("xpos").reshape(n, nb, 3), bd.internal("xquat").reshape(n, nb, 4),
              bd.internal("origin").reshape(n, nb, 3))
    cvel = util.to_com_frame(model, *frames, bd.internal("cvel").reshape(n, -1, 6), "body")
    cdof = util.to_com_frame(model, *frames, bd.internal("cdof").reshape(n, -1, 6), "dof")
    np.testing.assert_allclose(cvel, ref["cvel"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(cdof, ref["cdof"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", ["humanoid", "humanoid_nocontact", "arm26", "zoo", "weld", "capsbox", "gravcomp",
                                  "ref_inertia", "ref_dofless_weld", "ref_humanoid_sparse", "tendons"])
def test_subwarp_inertia_kernel(name, monkeypatch):
    """mj_crb + mj_factorM by the sub-warp kernel (8 lanes per state, intermediates in shared
    memory; MJB_INERTIA=subwarp selects it for A/B runs) against the reference's qM / qLD / qLDiagInv,
    and against the thread-per-state kernel's."""
    import mujoco_inversedynamicstest_b200 as mjb
    import edge_cases
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref = util.golden(name)
    model = mjb.Model.from_mjb(path)
    n = int(ref["nstate"])
    if name.startswith("ref_"):
        qpos, qvel, qacc = edge_cases.states(model, ref)
    else:
        qpos, qvel, qacc = generate_states(model, n, z_range=tuple(ref["z_range"]))
    got = {}
    for mode in ("thread", "subwarp"):
        monkeypatch.setenv("MJB_INERTIA", mode)
        bd = mjb.BatchData(model, n, outmask=mjb.OUT_INERTIA)
        bd.set_state(qpos, qvel, qacc)
        bd.inverse()
        got[mode] = {k: bd.get(f) for f, k in ((mjb.F_QM, "qM"), (mjb.F_QLD, "qLD"), (mjb.F_QLDIAGINV, "qLDiagInv"))}
        bd.close()
    for k in ("qM", "qLD", "qLDiagInv"):
        # entries that cancel to ~0 (off-diagonal terms of symmetric bodies) carry the rounding of the
        # entries they are made of: absolute term relative to the largest entry of the array
        atol = 1e-12 + 1e-10 * float(np.abs(ref[k]).max())
        np.testing.assert_allclose(got["subwarp"][k], ref[k], rtol=1e-9, atol=atol, err_msg=k)
        np.testing.assert_allclose(got["subwarp"][k], got["thread"][k], rtol=1e-9, atol=atol,
                                   err_msg=k + " vs thread-per-state")


@pytest.mark.parametrize("name", util.POST_CASES)
def test_golden_rne_post_constraint(name):
    """mjbOUT_RNEPOST: cacc, cfrc_int, cfrc_ext of the reference's mj_rnePostConstraint after
    mj_inverse (engine_core_smooth.c:2027-2181), 1e-9 relative / 1e-12 absolute."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref, n, zr = util.post_fixture(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=zr)
    bd = mjb.BatchData(model, n, outmask=mjb.OUT_RNEPOST)
    bd.set_state(qpos, qvel, qacc)
    assert bd.inverse() == 0
    got = bd.rne_post_constraint()
    for k in ("cacc", "cfrc_int", "cfrc_ext"):
        nviol, worst = util.spatial_violations(got[k], ref[k])
        assert nviol == 0, (k, nviol, worst)
    # bodies that no contact or equality constraint touches carry exactly zero (single components
    # that cancel exactly on the CPU may keep an FMA residual here, inside the tolerance above)
    assert (got["cfrc_ext"][(ref["cfrc_ext"] == 0).all(axis=-1)] == 0).all()
    # the extra outputs do not change qfrc_inverse by a bit
    plain = mjb.BatchData(model, n)
    plain.set_state(qpos, qvel, qacc)
    plain.inverse()
    np.testing.assert_array_equal(bd.qfrc_inverse(), plain.qfrc_inverse())


@pytest.mark.parametrize("name", util.BIAS_CASES)
def test_golden_qfrc_bias(name):
    """mjbF_QFRC_BIAS (mjbOUT_QFRC): mj_rne without accelerations, engine_forward.c:228."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref, n, zr = util.post_fixture(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=zr)
    bd = mjb.BatchData(model, n, outmask=mjb.OUT_QFRC)
    bd.set_state(qpos, qvel, qacc)
    assert bd.inverse() == 0
    nviol, worst = util.qfrc_violations(bd.get(mjb.F_QFRC_BIAS), ref["qfrc_bias"].reshape(n, -1))
    assert nviol == 0, (nviol, worst)
    # bias forces do not depend on the accelerations
    bd.set_state(qpos, qvel, -2.0 * qacc)
    bd.inverse()
    nviol, worst = util.qfrc_violations(bd.get(mjb.F_QFRC_BIAS), ref["qfrc_bias"].reshape(n, -1))
    assert nviol == 0, (nviol, worst)


@pytest.mark.parametrize("name", util.FWDINV_CASES)
def test_golden_compare_fwdinv(name):
    """mjb_compareFwdInv against the reference's mj_forward + mj_compareFwdInv."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, z = util.fwdinv_fixture(name)
    model = mjb.Model.from_mjb(path)
    n = int(z["nstate"])
    qpos, qvel, _ = generate_states(model, n, z_range=tuple(z["z_range"]))
    bd = mjb.BatchData(model, n)
    bd.set_state(qpos, qvel, z["qacc"])
    got = bd.compare_fwdinv(z["qfrc_constraint"], qfrc_applied=z["qfrc_applied"],
                            qfrc_actuator=z["qfrc_actuator"], xfrc_applied=z["xfrc_applied"])
    nviol, worst = util.fwdinv_violations(got, z)
    assert nviol == 0, (nviol, worst)
    # the check leaves the plain path as it was
    plain = mjb.BatchData(model, n)
    plain.set_state(qpos, qvel, z["qacc"])
    plain.inverse()
    bd.inverse()
    np.testing.assert_array_equal(bd.qfrc_inverse(), plain.qfrc_inverse())


def test_golden_per_state_mocap_poses():
    """mjb_setMocap: per-state d->mocap_pos / d->mocap_quat (mj_kinematics, engine_core_smooth.c:70-86)."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    z = np.load(os.path.join(util.GOLDEN, "mocap_moved.npz"))
    model = mjb.Model.from_mjb(os.path.join(util.GOLDEN, "mocap.mjb.gz"))
    n = int(z["nstate"])
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(z["z_range"]))
    bd = mjb.BatchData(model, n, outmask=mjb.OUT_COUNTS | mjb.OUT_CONTACT | mjb.OUT_EFC,
                       nconmax=int(z["nconmax"]), njmax=int(z["njmax"]))
    bd.set_state(qpos, qvel, qacc)
    bd.set_mocap(z["mocap_pos"], z["mocap_quat"])
    assert bd.inverse() == 0
    cnt = bd.counts()
    for k in ("ncon", "ne", "nf", "nl", "nefc"):
        np.testing.assert_array_equal(cnt[k], z[k], err_msg=k)
    np.testing.assert_array_equal(bd.contacts()["geom"], z["contact_geom"])
    efc = bd.efc()
    np.testing.assert_array_equal(efc["type"], z["efc_type"])
    np.testing.assert_array_equal(efc["state"], z["efc_state"])
    nviol, worst = util.qfrc_violations_scaled(bd.qfrc_inverse(), z["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)
    # back to the model pose: the plain fixture
    _, ref = util.golden("mocap")
    bd.set_mocap(None, None)
    bd.inverse()
    nviol, worst = util.qfrc_violations_scaled(bd.qfrc_inverse(), ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)


@pytest.mark.parametrize("case", ["touch", "touch_elliptic"])
def test_golden_touch_sensors(case):
    """touch sensors (engine_sensor.c:750-793): site volumes of every shape, both cones; the contact and
    efc outputs they read are switched on by mjb_makeData itself."""
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, case, True, 0)
    assert nbad == 0
    got = bd.sensordata()
    nviol, worst = util.sensor_violations(model, got, ref["sensordata"])
    assert nviol == 0, (nviol, worst)
    np.testing.assert_array_equal(got > 0, ref["sensordata"] > 0)
    assert (ref["sensordata"] > 0).mean() > 0.05


def test_golden_sensordata():
    """sensordata of mj_inverse (mj_sensorPos / Vel / Acc) through the C-ABI, and mj_inverseSkip's
    skipsensor."""
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, (qpos, qvel, qacc) = _run(mjb, "sensors", True, 0)
    assert nbad == 0
    got = bd.sensordata()
    nviol, worst = util.sensor_violations(model, got, ref["sensordata"])
    assert nviol == 0, (nviol, worst)
    # the sensors bring mj_rnePostConstraint's outputs with them
    post = bd.rne_post_constraint()
    assert np.isfinite(post["cfrc_int"]).all()
    # skipsensor = 1 leaves sensordata as it is (engine_inverse.c:206-246)
    bd.set_state(qpos[::-1].copy(), qvel[::-1].copy(), qacc[::-1].copy())
    bd.inverse_skip(0, 1)
    np.testing.assert_array_equal(bd.sensordata(), got)
    bd.inverse_skip(0, 0)
    np.testing.assert_array_equal(bd.sensordata(), got[::-1])


def test_golden_sensordata_camlight_transmission_energy():
    """Sensors that read other output-only stages: camprojection (mj_camlight poses, cam_project:
    engine_sensor.c:126-215), actuatorpos / actuatorvel (mj_transmission), e_potential / e_kinetic
    (mj_energyPos / mj_energyVel), magnetometer, clock; mjb_makeData adds the stages they need."""
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, "sensors2", True, 0)
    assert nbad == 0
    nviol, worst = util.sensor_violations(model, bd.sensordata(), ref["sensordata"])
    assert nviol == 0, (nviol, worst)
    # the stages came with the sensors
    assert np.isfinite(bd.camlight()["cam_xmat"]).all()
    assert np.isfinite(bd.transmission()["actuator_moment"]).all()


@pytest.mark.parametrize("name", util.KNOWN_ANSWER_CASES)
def test_reference_known_answers(name):
    """Known-answer tests the reference holds, restated through the C-ABI: force / torque sensor readings of
    bodies hanging on connect / weld constraints against the values written in the reference's model files
    (test/engine/engine_core_smooth_test.cc:160-300), potential energy (engine_sensor_test.cc:400-455), camera
    projection (:595-634), ray distances (engine_ray_test.cc:79-165)."""
    import mujoco_inversedynamicstest_b200 as mjb
    path, z = util.known_answer_fixture(name)
    model = mjb.Model.from_mjb(path)
    n = z["qpos"].shape[0]
    bd = mjb.BatchData(model, n)
    bd.set_state(z["qpos"], z["qvel"], z["qacc"])
    assert bd.inverse() == 0
    sensordata = bd.sensordata() if model.int("nsensordata") > 0 else None
    energy = bd.energy() if name == "ka_enable_energy" else None
    util.check_known_answer(name, z, sensordata, energy)


@pytest.mark.parametrize("name", util.EQACTIVE_CASES)
def test_golden_per_state_eq_active(name):
    """mjb_setEqActive: per-state d->eq_active (mj_instantiateEquality skips inactive constraints,
    engine_core_constraint.c:493-763; every later row moves up). Discrete outputs bit-identical, forces to
    rounding against the reference; NULL returns to eq_active0."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref, n, zr = util.eqactive_fixture(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=zr)
    bd = mjb.BatchData(model, n, outmask=mjb.OUT_QFRC | mjb.OUT_COUNTS | mjb.OUT_CONTACT | mjb.OUT_EFC,
                       nconmax=int(ref["nconmax"]), njmax=int(ref["njmax"]))
    bd.set_state(qpos, qvel, qacc)
    bd.set_eq_active(util.eq_active_samples(model, n))
    assert bd.inverse() == 0
    cnt = bd.counts()
    for k in ("ncon", "ne", "nf", "nl", "nefc"):
        np.testing.assert_array_equal(cnt[k], ref[k], err_msg=k)
    np.testing.assert_array_equal(bd.contacts()["geom"], ref["contact_geom"])
    np.testing.assert_array_equal(bd.contacts()["efc_address"], ref["contact_efc_address"])
    efc = bd.efc()
    for k in ("type", "id", "state"):
        np.testing.assert_array_equal(efc[k], ref["efc_" + k], err_msg=k)
    scale = max(1.0, np.abs(ref["efc_force"]).max())
    np.testing.assert_allclose(efc["force"], ref["efc_force"], rtol=1e-9, atol=1e-13 * scale)
    nviol, worst = util.qfrc_violations_scaled(bd.qfrc_inverse(), ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)
    ne_set = cnt["ne"].copy()
    bd.set_eq_active(None)                    # back to the model's eq_active0
    assert bd.inverse() == 0
    ne0 = bd.counts()["ne"]                   # a model constant again (the equality rows of eq_active0)
    assert (ne0 == ne0[0]).all() and (ne0 != ne_set).any()


@pytest.mark.parametrize("name", util.XFRC_CASES)
def test_golden_rne_post_constraint_with_xfrc_applied(name):
    """mjb_setXfrcApplied: per-state d->xfrc_applied enters cfrc_ext / cfrc_int of mj_rnePostConstraint
    (engine_core_smooth.c:2039-2049) and the force / torque sensors, not qfrc_inverse."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref, n, zr = util.xfrc_fixture(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=zr)
    bd = mjb.BatchData(model, n, outmask=mjb.OUT_RNEPOST, nconmax=704, njmax=1408)
    bd.set_state(qpos, qvel, qacc)
    bd.set_xfrc_applied(util.xfrc_samples(model, n))
    assert bd.inverse() == 0
    post = bd.rne_post_constraint()
    for k in ("cacc", "cfrc_int", "cfrc_ext"):
        nviol, worst = util.spatial_violations(post[k], ref[k])
        assert nviol == 0, (k, nviol, worst)
    nviol, worst = util.qfrc_violations_scaled(bd.qfrc_inverse(), ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)
    if "sensordata" in ref:
        nviol, worst = util.sensor_violations(model, bd.sensordata(), ref["sensordata"])
        assert nviol == 0, (nviol, worst)
    with_x = post["cfrc_ext"].copy()
    bd.set_xfrc_applied(None)                 # back to zero applied wrenches
    assert bd.inverse() == 0
    assert np.abs(bd.rne_post_constraint()["cfrc_ext"] - with_x).max() > 1


def test_rne_post_constraint_newton_euler_balance():
    """Size-independent property on 2^16 humanoid states: for a free-floating tree the root's
    cfrc_int is the wrench its (force-free) free joint transmits, so it equals qfrc_inverse of the
    free joint's dofs: force rows in world axes, torque rows in body axes about the root body's
    frame -- here checked through the force part, which is independent of the reference point."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, _ = util.golden("humanoid")
    model = mjb.Model.from_mjb(path)
    n = 1 << 16
    qpos, qvel, qacc = generate_states(model, n, z_range=(0.0, 1.5))
    bd = mjb.BatchData(model, n, outmask=mjb.OUT_RNEPOST)
    bd.set_state(qpos, qvel, qacc)
    assert bd.inverse() == 0
    post = bd.rne_post_constraint()
    q = bd.qfrc_inverse()
    f = post["cfrc_int"][:, 1, 3:6]
    scale = np.maximum(np.abs(f).max(axis=1, keepdims=True), 1.0)
    assert (np.abs(f - q[:, 0:3]) <= 1e-9 * scale).all()
    # the world body's row is the sum of the trees' root rows (engine_core_smooth.c:2178-2180)
    np.testing.assert_array_equal(post["cfrc_int"][:, 0], post["cfrc_int"][:, 1])
    assert np.isfinite(post["cacc"]).all() and np.isfinite(post["cfrc_ext"]).all()


def test_rne_post_constraint_refused_for_force_carrying_spatial_tendons():
    import mujoco_inversedynamicstest_b200 as mjb
    model = mjb.Model.from_mjb(util.golden("tendons")[0])
    with pytest.raises(mjb.MjbError, match="RNEPOST"):
        mjb.BatchData(model, 32, outmask=mjb.OUT_RNEPOST)


def test_ldl_reconstructs_mass_matrix():
    """L' D L == M (test/engine/engine_core_smooth_test.cc:466-511, 1e-12 there on a small model)."""
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, _, _ = _run(mjb, "humanoid", True, mjb.OUT_INERTIA)
    nv = model.int("nv")
    parent = model.array("dof_parentid").ravel()
    madr = model.array("dof_Madr").ravel()
    qM, qLD = bd.get(mjb.F_QM), bd.get(mjb.F_QLD)
    for s in range(0, qM.shape[0], 37):
        M = np.zeros((nv, nv)); L = np.eye(nv); D = np.zeros(nv)
        adrC = 0
        for i in range(nv):
            chain = []
            j = i
            while j >= 0:
                chain.append(j); j = parent[j]
            for t, j in enumerate(chain):
                M[i, j] = M[j, i] = qM[s, madr[i] + t]
            cols = chain[::-1]
            for t, j in enumerate(cols):
                if j == i: D[i] = qLD[s, adrC + t]
                else: L[i, j] = qLD[s, adrC + t]
            adrC += len(chain)
        np.testing.assert_allclose(L.T @ np.diag(D) @ L, M, rtol=1e-10, atol=1e-10)


@pytest.mark.skipif(not util.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name,n", [("humanoid", 4096), ("humanoid_elliptic", 4096)])
def test_live_reference_4096_states(name, n):
    """BASELINE config 1: 4096 random states, every discrete output bit-exact, qfrc within bound."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    from oracle import reflib
    path, _ = util.golden(name)
    import gzip, tempfile, os
    with tempfile.NamedTemporaryFile(suffix=".mjb", delete=False) as tf:
        tf.write(gzip.open(path, "rb").read())
    try:
        rm = reflib.Model.from_mjb(tf.name)
    finally:
        os.remove(tf.name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, first=100000)
    ref, _ = rm.inverse_batch(qpos, qvel, qacc, nthread=4, fields={
        "ncon": 1, "nefc": 1, "nl": 1, "contact_geom": 64, "efc_type": 256, "efc_id": 256})
    bd = mjb.BatchData(model, n, outmask=mjb.OUT_COUNTS | mjb.OUT_CONTACT | mjb.OUT_EFC, nconmax=64, njmax=256)
    bd.set_state(qpos, qvel, qacc)
    assert bd.inverse() == 0
    cnt = bd.counts()
    np.testing.assert_array_equal(cnt["ncon"], ref["ncon"])
    np.testing.assert_array_equal(cnt["nefc"], ref["nefc"])
    np.testing.assert_array_equal(cnt["nl"], ref["nl"])
    np.testing.assert_array_equal(bd.contacts()["geom"], ref["contact_geom"])
    efc = bd.efc()
    np.testing.assert_array_equal(efc["type"], ref["efc_type"][:, :, 0])
    np.testing.assert_array_equal(efc["id"], ref["efc_id"][:, :, 0])
    nviol, worst = util.qfrc_violations_scaled(bd.qfrc_inverse(), ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)


def test_full_size_properties_1m_states():
    """2^20 humanoid states (BASELINE configs 2/3): size-independent properties.
    - determinism: two launches give bit-identical qfrc_inverse
    - batch-position independence: a state gives the same bits wherever it sits in the batch
    - the first 256 states equal the golden dump within tolerance
    - no status flags on valid inputs; NaN / huge inputs are flagged per state"""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref = util.golden("humanoid")
    model = mjb.Model.from_mjb(path)
    n = 1 << 20
    qpos, qvel, qacc = generate_states(model, n)
    bd = mjb.BatchData(model, n, outmask=mjb.OUT_COUNTS)
    bd.set_state(qpos, qvel, qacc)
    assert bd.inverse() == 0
    a = bd.qfrc_inverse()
    assert np.isfinite(a).all()
    assert bd.inverse() == 0
    b = bd.qfrc_inverse()
    assert np.array_equal(a, b)
    nviol, worst = util.qfrc_violations_scaled(a[:256], ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)
    np.testing.assert_array_equal(bd.counts()["ncon"][:256], ref["ncon"])
    # reversed batch order
    bd.set_state(qpos[::-1], qvel[::-1], qacc[::-1])
    assert bd.inverse() == 0
    c = bd.qfrc_inverse()[::-1]
    assert np.array_equal(a, c)
    # bad inputs
    qpos2 = qpos[:1024].copy(); qvel2 = qvel[:1024].copy(); qacc2 = qacc[:1024].copy()
    qpos2[3, 5] = np.nan; qvel2[7, 0] = 1e11; qacc2[9, 2] = -np.inf
    bd.set_state(qpos2, qvel2, qacc2)
    assert bd.inverse() == 3
    st = bd.status()
    assert st[3] & 1 and st[7] & 2 and st[9] & 4 and (st != 0).sum() == 3


def test_empty_and_ragged_batches():
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref = util.golden("humanoid")
    model = mjb.Model.from_mjb(path)
    bd = mjb.BatchData(model, 1000)
    qpos, qvel, qacc = generate_states(model, 256)
    for n in (0, 1, 31, 33, 129, 256):
        bd.set_state(qpos[:n], qvel[:n], qacc[:n])
        assert bd.inverse() == 0
        got = bd.qfrc_inverse()
        assert got.shape == (n, model.int("nv"))
        if n:
            nviol, worst = util.qfrc_violations_scaled(got, ref["qfrc_inverse"][:n])
            assert nviol == 0, (n, nviol, worst)
    with pytest.raises(mjb.MjbError):
        bd.set_state(np.zeros((2000, 28)), np.zeros((2000, 27)), np.zeros((2000, 27)))


def test_output_capacity_overflow_is_flagged():
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, "humanoid", True, mjb.OUT_COUNTS)
    path, _ = util.golden("humanoid")
    from mujoco_inversedynamicstest_b200.states import generate_states
    qpos, qvel, qacc = generate_states(model, 256)
    small = mjb.BatchData(model, 256, outmask=mjb.OUT_COUNTS | mjb.OUT_CONTACT | mjb.OUT_EFC, nconmax=4, njmax=8)
    small.set_state(qpos, qvel, qacc)
    nbad = small.inverse()
    st = small.status()
    expect = ((ref["ncon"] > 4) * 8) | ((ref["nefc"] > 8) * 16)
    np.testing.assert_array_equal(st, expect)
    assert nbad == int((expect != 0).sum())
    # the physics is not truncated by the output capacity
    nviol, worst = util.qfrc_violations_scaled(small.qfrc_inverse(), ref["qfrc_inverse"])
    assert nviol == 0


def test_unsupported_models_are_rejected_at_upload(tmp_path):
    import mujoco_inversedynamicstest_b200 as mjb
    path, _ = util.golden("slider_crank_nocontact")
    model = mjb.Model.from_mjb(path)
    mjb.BatchData(model, 8)                       # contacts disabled: accepted
    model.set_opt_int("disableflags", 0)          # contacts on: capsule-cylinder pairs -> mjc_Convex (GJK / EPA)
    mjb.BatchData(model, 8)
    model.set_opt_int("enableflags", 1 << 4)      # mjENBL_MULTICCD
    with pytest.raises(mjb.MjbError, match="MULTICCD"):
        mjb.BatchData(model, 8)
    model.set_opt_int("enableflags", 0)
    model.array("geom_type")[1] = 7               # a mesh geom: no collision function here
    with pytest.raises(mjb.MjbError, match="collision function"):
        mjb.BatchData(model, 8)


def test_cuda_vs_c_restatement_on_fresh_states():
    """Third leg of the three-way check: CUDA path vs oracle/mjinv_oracle.c (dense-Jacobian
    restatement) on states that are in no fixture."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    from oracle import restatement
    from test_oracle_restatement import ctypes_opt
    for name in ("humanoid", "humanoid_elliptic"):
        model = mjb.Model.from_mjb(util.golden(name)[0])
        r = restatement.Restatement(model, ctypes_opt(model))
        n = 512
        qpos, qvel, qacc = generate_states(model, n, first=7_000_000)
        ref = r.inverse_batch(qpos, qvel, qacc, inertia=False)
        bd = mjb.BatchData(model, n, outmask=mjb.OUT_COUNTS | mjb.OUT_CONTACT | mjb.OUT_EFC, nconmax=64, njmax=256)
        bd.set_state(qpos, qvel, qacc)
        assert bd.inverse() == 0
        cnt = bd.counts()
        np.testing.assert_array_equal(cnt["ncon"], ref["ncon"])
        np.testing.assert_array_equal(cnt["nefc"], ref["nefc"])
        np.testing.assert_array_equal(bd.contacts()["geom"], ref["contact_geom"])
        np.testing.assert_array_equal(bd.efc()["type"], ref["efc_type"])
        nviol, worst = util.qfrc_violations_scaled(bd.qfrc_inverse(), ref["qfrc_inverse"])
        assert nviol == 0, (name, nviol, worst)


def test_pipelined_host_entry_point_matches_staged_calls():
    """mjb_inverseHost (three-stream pipeline over pieces) == setState + inverse + get, bit for bit,
    including a batch that is not a multiple of the piece size and repeated use of the buffers."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    model = mjb.Model.from_mjb(util.golden("humanoid")[0])
    n = 300_000
    qpos, qvel, qacc = generate_states(model, n, first=12345)
    bd = mjb.BatchData(model, n)
    bd.set_state(qpos, qvel, qacc)
    assert bd.inverse() == 0
    a = bd.qfrc_inverse()
    for _ in range(2):
        b = bd.inverse_host_arrays(qpos, qvel, qacc)
        assert np.array_equal(a, b)
    c = bd.inverse_host_arrays(qpos[:1000], qvel[:1000], qacc[:1000])
    assert np.array_equal(a[:1000], c)
    # consecutive calls overlap (the copy-in of a call runs under the kernels of the previous one):
    # three calls on different inputs and outputs without a synchronisation in between
    import torch
    nv = model.int("nv")
    sets = []
    for k in range(3):
        sl = slice(k * 70_001, k * 70_001 + 150_000)
        ins = [torch.from_numpy(np.ascontiguousarray(x[sl])).pin_memory() for x in (qpos, qvel, qacc)]
        out = torch.zeros((150_000, nv), dtype=torch.float64).pin_memory()
        sets.append((sl, ins, out))
    for sl, ins, out in sets:
        bd.inverse_host(150_000, ins[0].data_ptr(), ins[1].data_ptr(), ins[2].data_ptr(), out.data_ptr())
    bd.synchronize()
    for sl, ins, out in sets:
        assert np.array_equal(out.numpy(), a[sl])


def test_inverse_skip_equals_inverse():
    """mj_inverseSkip(m, d, mjSTAGE_VEL, 1) as the fork's driver calls it (src/inverse/inverse_test.cpp:93):
    with unchanged inputs it must return what mj_inverse returns."""
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, "humanoid", True, 0)
    full = bd.qfrc_inverse().copy()
    for stage in (0, 1, 2):
        assert bd.inverse_skip(skipstage=stage, skipsensor=1) == 0
        np.testing.assert_array_equal(bd.qfrc_inverse(), full)
    with pytest.raises(mjb.MjbError):
        bd.inverse_skip(skipstage=7)


def test_item_parallel_and_pooled_contact_paths_agree():
    """The contact phase has two implementations: item-parallel kernels over global lists (default)
    and the warp-pooled kernel (on-device fallback when the lists overflow, MJB_CONTACT_PATH=pooled).
    Both apply J'f per (state, body) in contact order starting from the accumulator row, so they
    must agree bit for bit, on discrete outputs and on qfrc_inverse."""
    import os
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    for name in ("humanoid", "humanoid_elliptic", "boxes"):
        path, ref = util.golden(name)
        model = mjb.Model.from_mjb(path)
        n = 4096
        qpos, qvel, qacc = generate_states(model, n, z_range=tuple(ref["z_range"]))
        res = {}
        for mode in ("items", "pooled"):
            if mode == "pooled":
                os.environ["MJB_CONTACT_PATH"] = "pooled"
            try:
                bd = mjb.BatchData(model, n, outmask=mjb.OUT_COUNTS | mjb.OUT_EFC, njmax=int(ref["njmax"]))
            finally:
                os.environ.pop("MJB_CONTACT_PATH", None)
            bd.set_state(qpos, qvel, qacc)
            assert bd.inverse() == 0
            res[mode] = (bd.qfrc_inverse().copy(), bd.counts(), bd.efc())
        np.testing.assert_array_equal(res["items"][1]["ncon"], res["pooled"][1]["ncon"])
        np.testing.assert_array_equal(res["items"][1]["nefc"], res["pooled"][1]["nefc"])
        np.testing.assert_array_equal(res["items"][2]["type"], res["pooled"][2]["type"])
        np.testing.assert_array_equal(res["items"][2]["state"], res["pooled"][2]["state"])
        np.testing.assert_array_equal(res["items"][0], res["pooled"][0], err_msg=name)


def test_contact_list_overflow_falls_back_on_the_device():
    """A chunk whose contacts do not fit the global lists (22 interpenetrating humanoids: ~290
    contacts per state against a forced capacity of 16 per state) is handled by the pooled kernel;
    the counters of the item path show the overflow and the results are the reference's. With the
    per-model list sizing the same scene runs on the item-parallel path (next test)."""
    import ctypes
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200._lib import lib
    import os
    os.environ["MJB_ITEMS_PER_STATE"] = "48"            # the sizing of a single articulated figure
    os.environ["MJB_CONTACTS_PER_STATE"] = "16"
    try:
        model, bd, ref, nbad, _ = _run(mjb, "humanoids22", True, mjb.OUT_COUNTS)
    finally:
        os.environ.pop("MJB_ITEMS_PER_STATE", None)
        os.environ.pop("MJB_CONTACTS_PER_STATE", None)
    out = (ctypes.c_int * 4)()
    assert lib().mjb_debugQueue(bd._d, out) == 0
    assert out[2] != 0, list(out)                       # overflow flag raised on the device
    np.testing.assert_array_equal(bd.counts()["ncon"], ref["ncon"])
    nviol, worst = util.qfrc_violations_scaled(bd.qfrc_inverse(), ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)


def test_multi_tree_scene_runs_on_the_item_parallel_path():
    """22 humanoids (BASELINE config 5), 256 states: the lists are sized from the geom count, so the
    item-parallel contact kernels handle the scene (no overflow, no pooled fallback); contact lists
    and efc ordering bit-exact against the 256-state dump."""
    import ctypes
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200._lib import lib
    model, bd, ref, nbad, _ = _run(mjb, "humanoids22_256", True, mjb.OUT_COUNTS | mjb.OUT_CONTACT | mjb.OUT_EFC)
    assert nbad == 0
    out = (ctypes.c_int * 4)()
    assert lib().mjb_debugQueue(bd._d, out) == 0
    assert out[2] == 0, list(out)
    assert out[0] == int(bd.counts()["ncon"].size and out[0])   # items were appended
    cnt = bd.counts()
    for k in ("ncon", "ne", "nf", "nl", "nefc"):
        np.testing.assert_array_equal(cnt[k], ref[k], err_msg=k)
    np.testing.assert_array_equal(bd.contacts()["geom"], ref["contact_geom"])
    efc = bd.efc()
    for k in ("type", "id", "state"):
        np.testing.assert_array_equal(efc[k], ref["efc_" + k], err_msg=k)


def _run_fixture(mjb, path, n, zr, outmask):
    from mujoco_inversedynamicstest_b200.states import generate_states
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=zr)
    bd = mjb.BatchData(model, n, outmask=outmask)
    bd.set_state(qpos, qvel, qacc)
    assert bd.inverse() == 0
    return model, bd


@pytest.mark.parametrize("name", ["camlight_cl", "humanoid_cl"])
def test_golden_camlight(name):
    """mjbOUT_CAMLIGHT: cam_xpos / cam_xmat / light_xpos / light_xdir as mj_camlight leaves them inside
    mj_invPosition (engine_core_smooth.c:275-389), every mjtCamLight mode, 1e-9 relative / 1e-12 absolute."""
    import mujoco_inversedynamicstest_b200 as mjb
    path, ref, n, zr = util.camlight_fixture(name)
    model, bd = _run_fixture(mjb, path, n, zr, mjb.OUT_CAMLIGHT)
    got = bd.camlight()
    for k in ("cam_xpos", "cam_xmat", "light_xpos", "light_xdir"):
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-9, atol=1e-12, err_msg=k)


@pytest.mark.parametrize("name", ["transmission_trn", "humanoid_trn", "arm26_trn", "slider_crank_trn", "adhesion_trn",
                                  "adhesion_elliptic_trn"])
def test_golden_transmission(name):
    """mjbOUT_TRANSMISSION: actuator_length, actuator_moment (dense) and actuator_velocity as mj_transmission
    and mj_fwdVelocity leave them inside mj_inverse (engine_core_smooth.c:865-1346, engine_forward.c:216);
    joint / jointinparent / slider-crank / tendon / site transmissions, 1e-9 relative / 1e-12 absolute."""
    import mujoco_inversedynamicstest_b200 as mjb
    path, ref, n, zr = util.transmission_fixture(name)
    model, bd = _run_fixture(mjb, path, n, zr, mjb.OUT_TRANSMISSION)
    got = bd.transmission()
    np.testing.assert_allclose(got["actuator_length"], ref["actuator_length"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got["actuator_moment"], ref["actuator_moment"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got["actuator_velocity"], ref["actuator_velocity"], rtol=1e-9, atol=1e-11)


def test_outputs_only_rows_leave_qfrc_inverse_unchanged():
    """Requesting the output-only rows (camlight, transmission) does not change qfrc_inverse."""
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, "humanoid", True, mjb.OUT_CAMLIGHT | mjb.OUT_TRANSMISSION)
    assert nbad == 0
    nviol, worst = util.qfrc_violations(bd.qfrc_inverse(), ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)


@pytest.mark.parametrize("name", ["humanoid_energy", "zoo_energy", "tendons_energy"])
def test_golden_energy(name):
    """mjENBL_ENERGY: d->energy = (potential, kinetic) as mj_inverse leaves it (mj_energyPos, mj_energyVel:
    engine_sensor.c:920, 1011; called from engine_inverse.c:210-223), 1e-9 relative / 1e-12 absolute."""
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, name, True, mjb.OUT_COUNTS)
    assert nbad == 0
    np.testing.assert_array_equal(bd.counts()["ncon"], ref["ncon"])
    np.testing.assert_allclose(bd.energy(), ref["energy"], rtol=1e-9, atol=1e-12)
    nviol, worst = util.qfrc_violations(bd.qfrc_inverse(), ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)


def test_hundred_humanoids_scene():
    """model/humanoid/100_humanoids.xml of the reference (nv = 2700, 1901 geoms, 1.8 M candidate pairs,
    ~4,300 contacts per state): counters, contact lists and efc ordering bit-exact against the
    reference's dump, qfrc_inverse element-wise within 1e-9 * |ref| + 1e-12."""
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, "humanoids100_4", True, mjb.OUT_COUNTS | mjb.OUT_CONTACT | mjb.OUT_EFC)
    assert nbad == 0
    cnt = bd.counts()
    for k in ("ncon", "ne", "nf", "nl", "nefc"):
        np.testing.assert_array_equal(cnt[k], ref[k], err_msg=k)
    np.testing.assert_array_equal(bd.contacts()["geom"], ref["contact_geom"])
    efc = bd.efc()
    for k in ("type", "id", "state"):
        np.testing.assert_array_equal(efc[k], ref["efc_" + k], err_msg=k)
    # forces of 1e7 summed over thousands of contacts: the same documented allowance as humanoids22_256
    nviol, worst = util.qfrc_violations(bd.qfrc_inverse(), ref["qfrc_inverse"])
    max_viol, max_ratio = util.STRICT_EXCEPTIONS["humanoids22_256"]
    assert nviol <= max_viol and worst <= max_ratio, (nviol, worst)
    nviol, worst = util.qfrc_violations_scaled(bd.qfrc_inverse(), ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)


@pytest.mark.parametrize("name", ["humanoid_fd", "zoo_fd", "humanoid_nocontact_fd"])
def test_inverse_fd_matches_reference(name):
    """mjb_inverseFD against the reference's mjd_inverseFD (engine_derivative_fd.c:611) on the same
    states and eps. A forward difference amplifies the 1e-9-relative agreement of the two
    evaluations by 1/eps, hence the tolerance: 2e-3 of the largest entry of the state's Jacobian
    (contacts switching between the base and a perturbed state are the same in both engines)."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    z = np.load(os.path.join(util.GOLDEN, name + ".npz"))
    base = str(z["base"])
    model = mjb.Model.from_mjb(util.golden(base)[0])
    n = int(z["nstate"])
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(z["z_range"]))
    bd = mjb.BatchData(model, n)
    bd.set_state(qpos, qvel, qacc)
    dq, dv, da, dm = bd.inverse_fd(eps=float(z["eps"]), mass=True)
    for got, key in ((dq, "DfDq"), (dv, "DfDv"), (da, "DfDa"), (dm, "DmDq")):
        ref = z[key]
        scale = np.abs(ref).reshape(n, -1).max(axis=1)[:, None, None]
        err = np.abs(got - ref) / np.maximum(scale, 1e-6)
        assert err.max() < 2e-3, (key, float(err.max()))


def test_inverse_fd_sensor_jacobians_match_reference():
    """mjb_inverseFDSensor: DsDq / DsDv / DsDa of mjd_inverseFD (engine_derivative_fd.c:611-730) on the
    62-sensor fixture, together with the force Jacobians of the same call; same tolerance reading as
    test_inverse_fd_matches_reference (2e-3 of the largest entry of the state's Jacobian)."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    z = np.load(os.path.join(util.GOLDEN, "sensors_fd.npz"))
    model = mjb.Model.from_mjb(util.golden(str(z["base"]))[0])
    n = int(z["nstate"])
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(z["z_range"]))
    bd = mjb.BatchData(model, n)
    bd.set_state(qpos, qvel, qacc)
    got = bd.inverse_fd_sensor(eps=float(z["eps"]))
    for g, key in zip(got, ("DfDq", "DfDv", "DfDa", "DsDq", "DsDv", "DsDa")):
        ref = z[key]
        scale = np.abs(ref).reshape(n, -1).max(axis=1)[:, None, None]
        err = np.abs(g - ref) / np.maximum(scale, 1e-6)
        assert err.max() < 2e-3, (key, float(err.max()))


def test_contact_record_list_overflow_falls_back_on_the_device():
    """The CONTACT list (not the item list) overflows: the survivors fit, the contacts they yield do
    not (capacity forced to 1 contact per state with MJB_CONTACTS_PER_STATE). contact_narrow_kernel
    raises its own flag -- which it never tests itself, so all its CTAs run to the end -- and the
    pooled kernel produces the reference's results."""
    import ctypes
    import os
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200._lib import lib
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref = util.golden("humanoid")
    model = mjb.Model.from_mjb(path)
    n = 8192
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(ref["z_range"]))
    os.environ["MJB_CONTACTS_PER_STATE"] = "1"
    try:
        small = mjb.BatchData(model, n, outmask=mjb.OUT_COUNTS | mjb.OUT_CONTACT, nconmax=64)
    finally:
        os.environ.pop("MJB_CONTACTS_PER_STATE", None)
    full = mjb.BatchData(model, n, outmask=mjb.OUT_COUNTS | mjb.OUT_CONTACT, nconmax=64)
    for bd in (small, full):
        bd.set_state(qpos, qvel, qacc)
        assert bd.inverse() == 0
    out = (ctypes.c_int * 4)()
    assert lib().mjb_debugQueue(small._d, out) == 0
    assert out[2] == 2, list(out)                       # contact list overflowed, item list did not
    assert lib().mjb_debugQueue(full._d, out) == 0 and out[2] == 0
    np.testing.assert_array_equal(small.counts()["ncon"], full.counts()["ncon"])
    np.testing.assert_array_equal(small.contacts()["geom"], full.contacts()["geom"])
    np.testing.assert_array_equal(small.qfrc_inverse(), full.qfrc_inverse())
    np.testing.assert_array_equal(small.counts()["ncon"][:256], ref["ncon"])


def test_get_is_sized_by_the_last_evaluation():
    """mjb_get copies the batch of the LAST evaluation; set_state afterwards must not shrink the
    buffer the Python mirror hands to it, and an undersized `out` is refused."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref = util.golden("humanoid_nocontact")
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, 1000, z_range=tuple(ref["z_range"]))
    bd = mjb.BatchData(model, 1000)
    bd.set_state(qpos, qvel, qacc)
    bd.inverse()
    first = bd.qfrc_inverse().copy()
    bd.set_state(qpos[:10], qvel[:10], qacc[:10])
    assert bd.last_batch() == 1000
    again = bd.qfrc_inverse()
    assert again.shape == (1000, model.int("nv"))
    np.testing.assert_array_equal(first, again)
    with pytest.raises(mjb.MjbError):
        bd.get(mjb.F_QFRC_INVERSE, out=np.empty((10, model.int("nv"))))
