"""CPU check of the per-state pipeline (the exact functions the CUDA kernel runs, compiled as plain
C++ for this purpose only, tests/hostemu) against the golden dumps of the reference.

This is what lets the algorithmic restructuring -- Jacobian-free constraint rows, static candidate
pairs instead of broadphase/midphase, wrench accumulation inside RNE -- be verified in a container
without a GPU. Built with -ffp-contract=off, so even the kinematics are bit-identical here."""
import os
import sys

import numpy as np
import pytest

import util

sys.path.insert(0, os.path.join(util.ROOT, "tests", "hostemu"))
import emu  # noqa: E402

pytestmark = pytest.mark.skipif(not emu.available(), reason="host emulation library not buildable")

CASES = ["humanoid", "humanoid_elliptic", "humanoid_nocontact", "humanoids22",
         "slider_crank_nocontact", "inverse_test", "arm26", "weld", "connect", "zoo", "zoo_elliptic",
         "gravcomp", "humanoid_invdiscrete", "capsbox", "capsbox_elliptic", "boxes", "boxes_elliptic", "tendons",
         "sensors", "mocap", "touch", "touch_elliptic", "camlight", "transmission", "sensors2",
         "humanoid_invdiscrete_fast", "implicitfast", "humanoid_invdiscrete_implicit", "implicit", "adhesion",
         "adhesion_elliptic", "fluid", "fluid_box", "tendon_eq", "slider_crank", "convex"]


def _run(name):
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref = util.golden(name)
    model = mjb.Model.from_mjb(path)
    n = int(ref["nstate"])
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(ref["z_range"]))
    out = emu.run(model, qpos, qvel, qacc, nconmax=int(ref["nconmax"]), njmax=int(ref["njmax"]))
    return model, out, ref


@pytest.mark.parametrize("name", CASES)
def test_discrete_outputs_bit_exact(name):
    model, out, ref = _run(name)
    assert (out["status"] == 0).all()
    for k in ("ncon", "ne", "nf", "nl", "nefc", "contact_geom", "contact_dim", "contact_exclude",
              "contact_efc_address", "efc_type", "efc_id", "efc_state"):
        np.testing.assert_array_equal(out[k], ref[k], err_msg=k)


@pytest.mark.parametrize("name", CASES)
def test_continuous_outputs(name):
    model, out, ref = _run(name)
    nviol, worst = util.qfrc_violations_scaled(out["qfrc_inverse"], ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)
    n = out["qfrc_inverse"].shape[0]
    # position stage is bit-identical on the CPU build
    np.testing.assert_array_equal(emu.slot(model, out, "xpos").reshape(n, -1, 3), ref["xpos"])
    # spatial vectors are kept about the tree origin instead of the tree's centre of mass: the
    # rotational part is bit-identical, the translational part agrees after the change of origin
    nb = model.int("nbody")
    xp = emu.slot(model, out, "xpos").reshape(n, nb, 3)
    cdof = util.to_com_frame(model, xp, emu.slot(model, out, "xquat").reshape(n, nb, 4),
                             emu.slot(model, out, "origin").reshape(n, nb, 3),
                             emu.slot(model, out, "cdof").reshape(n, -1, 6), "dof")
    np.testing.assert_array_equal(cdof[..., :3], ref["cdof"][..., :3])
    np.testing.assert_allclose(cdof, ref["cdof"], rtol=1e-12, atol=1e-13)
    np.testing.assert_array_equal(out["contact_dist"], ref["contact_dist"])
    np.testing.assert_array_equal(out["contact_pos"], ref["contact_pos"])
    np.testing.assert_array_equal(out["contact_frame"], ref["contact_frame"])
    # qM = cdof' crb cdof is independent of the frame origin: equal to rounding
    np.testing.assert_allclose(out["qM"], ref["qM"], rtol=1e-9, atol=1e-12)
    # qLD comes from the articulated-body recursion, not from eliminating M: equal to rounding
    np.testing.assert_allclose(out["qLD"], ref["qLD"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out["qLDiagInv"], ref["qLDiagInv"], rtol=1e-9, atol=1e-12)
    # tendon spring and damper terms are summed in one accumulator here, in two in the reference
    np.testing.assert_allclose(out["qfrc_passive"], ref["qfrc_passive"], rtol=1e-11, atol=1e-13)   # gravcomp / spatial-tendon wrenches are projected with the fused dot product
    np.testing.assert_array_equal(out["efc_pos"], ref["efc_pos"])
    np.testing.assert_allclose(out["efc_D"], ref["efc_D"], rtol=1e-14)
    scale = max(1.0, np.abs(ref["efc_force"]).max())
    np.testing.assert_allclose(out["efc_force"], ref["efc_force"], rtol=1e-9, atol=1e-15 * scale)
    np.testing.assert_allclose(out["efc_aref"], ref["efc_aref"], rtol=1e-9, atol=1e-15 * scale)
    cs = np.abs(ref["qfrc_constraint"]).max(axis=1, keepdims=True)
    assert (np.abs(out["qfrc_constraint"] - ref["qfrc_constraint"]) <= 1e-12 + 1e-12 * cs).all()


@pytest.mark.parametrize("name", util.POST_CASES)
def test_rne_post_constraint_outputs(name):
    """cacc, cfrc_int, cfrc_ext against the reference's mj_rnePostConstraint run after mj_inverse
    (engine_core_smooth.c:2027-2181), incl. its raw-torque convention for weld rows."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref, n, zr = util.post_fixture(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=zr)
    out = emu.run(model, qpos, qvel, qacc, nconmax=640, njmax=1200, post=True)
    nb = model.int("nbody")
    for k in ("cacc", "cfrc_int", "cfrc_ext"):
        nviol, worst = util.spatial_violations(out[k].reshape(n, nb, 6), ref[k])
        assert nviol == 0, (k, nviol, worst)
    # bodies without contacts or equality constraints carry exactly zero external force
    zero = ref["cfrc_ext"] == 0
    assert (out["cfrc_ext"].reshape(n, nb, 6)[zero] == 0).all()


@pytest.mark.parametrize("name", ["humanoid", "zoo", "sensors", "sensors2", "touch", "implicit", "tendons", "weld", "arm26", "transmission",
                                  "humanoid_invdiscrete", "mocap", "fluid", "fluid_box", "tendon_eq", "geomdist", "geomdist_ccd", "slider_crank", "convex"])
def test_outputs_do_not_depend_on_the_debug_dump(name):
    """The product stores a scratch row only where a later stage reads it; the debug dump (mjbOUT_INTERNAL,
    used by the other tests of this file) stores everything. Same results either way, bit for bit."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref = util.golden(name)
    model = mjb.Model.from_mjb(path)
    n = min(32, int(ref["nstate"]))
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(ref["z_range"]))
    kw = dict(nconmax=int(ref["nconmax"]), njmax=int(ref["njmax"]))
    a = emu.run(model, qpos, qvel, qacc, **kw)
    b = emu.run(model, qpos, qvel, qacc, dump=False, **kw)
    for k in b:
        if k != "scratch_dump":
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.parametrize("name", util.KNOWN_ANSWER_CASES)
def test_reference_known_answers(name):
    """Known-answer tests the reference holds, restated on the inverse path: force / torque sensors of bodies
    hanging on connect / weld constraints equal the values written in the model files
    (test/engine/engine_core_smooth_test.cc:160-300, 1e-6), potential energy (engine_sensor_test.cc:400-455,
    exact), camera projection (:595-634, 1e-4), ray distances (engine_ray_test.cc:79-165)."""
    import mujoco_inversedynamicstest_b200 as mjb
    path, z = util.known_answer_fixture(name)
    model = mjb.Model.from_mjb(path)
    out = emu.run(model, z["qpos"], z["qvel"], z["qacc"], nconmax=16, njmax=64, post=name.startswith(("ka_connect", "ka_weld")),
                  dump=False)
    util.check_known_answer(name, z, out.get("sensordata"), out.get("energy"))


@pytest.mark.parametrize("name", util.EQACTIVE_CASES)
def test_per_state_eq_active(name):
    """Per-state d->eq_active: inactive equality constraints emit no rows and every later row (friction loss,
    limits, contacts) moves up; discrete outputs bit-identical, forces to rounding, against the reference."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref, n, zr = util.eqactive_fixture(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=zr)
    ea = util.eq_active_samples(model, n)
    weld_post = name != "mocap_eqactive"      # mocap.xml has a massless tree: no mjbOUT_RNEPOST
    out = emu.run(model, qpos, qvel, qacc, nconmax=int(ref["nconmax"]), njmax=int(ref["njmax"]), post=weld_post,
                  eq_active=ea, dump=False)
    for k in ("ncon", "ne", "nf", "nl", "nefc", "contact_geom", "contact_efc_address", "efc_type", "efc_id", "efc_state"):
        np.testing.assert_array_equal(out[k], ref[k], err_msg=k)
    assert ref["ne"].min() == 0 and ref["ne"].max() > 0
    nviol, worst = util.qfrc_violations_scaled(out["qfrc_inverse"], ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)
    scale = max(1.0, np.abs(ref["efc_force"]).max())
    np.testing.assert_allclose(out["efc_force"], ref["efc_force"], rtol=1e-9, atol=1e-13 * scale)
    if weld_post:
        nb = model.int("nbody")
        for k in ("cacc", "cfrc_int", "cfrc_ext"):
            nviol, worst = util.spatial_violations(out[k].reshape(n, nb, 6), ref[k])
            assert nviol == 0, (k, nviol, worst)


@pytest.mark.parametrize("name", util.XFRC_CASES)
def test_rne_post_constraint_with_xfrc_applied(name):
    """Per-state d->xfrc_applied enters cfrc_ext / cfrc_int (engine_core_smooth.c:2039-2049) and the force /
    torque sensors, not qfrc_inverse; against the reference's mj_inverse + mj_rnePostConstraint."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref, n, zr = util.xfrc_fixture(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=zr)
    x = util.xfrc_samples(model, n)
    out = emu.run(model, qpos, qvel, qacc, nconmax=640, njmax=1200, post=True, xfrc=x, dump=False)
    nb = model.int("nbody")
    for k in ("cacc", "cfrc_int", "cfrc_ext"):
        nviol, worst = util.spatial_violations(out[k].reshape(n, nb, 6), ref[k])
        assert nviol == 0, (k, nviol, worst)
    nviol, worst = util.qfrc_violations_scaled(out["qfrc_inverse"], ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)
    if "sensordata" in ref:
        nviol, worst = util.sensor_violations(model, out["sensordata"], ref["sensordata"])
        assert nviol == 0, (nviol, worst)
    # and the applied wrenches do change the outputs
    plain = emu.run(model, qpos, qvel, qacc, nconmax=640, njmax=1200, post=True)
    assert np.abs(plain["cfrc_ext"] - out["cfrc_ext"]).max() > 1


@pytest.mark.parametrize("name", util.BIAS_CASES)
def test_qfrc_bias(name):
    """qfrc_bias = mj_rne without accelerations (mj_fwdVelocity, engine_forward.c:228), strict
    element-wise 1e-9 rel / 1e-12 abs against the reference."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref, n, zr = util.post_fixture(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=zr)
    out = emu.run(model, qpos, qvel, qacc, nconmax=640, njmax=1200)
    nviol, worst = util.qfrc_violations(out["qfrc_bias"], ref["qfrc_bias"].reshape(n, -1))
    assert nviol == 0, (nviol, worst)


@pytest.mark.parametrize("name", util.FWDINV_CASES)
def test_compare_fwdinv(name):
    """solver_fwdinv of the reference's mj_forward + mj_compareFwdInv (engine_inverse.c:275-316), at
    the forward solution and away from it, with ctrl, qfrc_applied and xfrc_applied in play."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, z = util.fwdinv_fixture(name)
    model = mjb.Model.from_mjb(path)
    n = int(z["nstate"])
    qpos, qvel, _ = generate_states(model, n, z_range=tuple(z["z_range"]))
    fwd = {"qforce": z["qfrc_applied"] + z["qfrc_actuator"], "qfrc_constraint": z["qfrc_constraint"],
           "xfrc": z["xfrc_applied"]}
    out = emu.run(model, qpos, qvel, z["qacc"], nconmax=64, njmax=300, fwd=fwd)
    nviol, worst = util.fwdinv_violations(out["fwdinv"], z)
    assert nviol == 0, (nviol, worst)
    assert z["fwdinv"][1::2].min() > 0.5          # the perturbed half is a non-trivial comparison


def test_per_state_mocap_poses():
    """d->mocap_pos / d->mocap_quat as per-state inputs (mj_kinematics, engine_core_smooth.c:70-86;
    quaternions arrive unnormalised), welds / connects / contacts against the moved bodies."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    z = np.load(os.path.join(util.GOLDEN, "mocap_moved.npz"))
    model = mjb.Model.from_mjb(os.path.join(util.GOLDEN, "mocap.mjb.gz"))
    n = int(z["nstate"])
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(z["z_range"]))
    out = emu.run(model, qpos, qvel, qacc, nconmax=int(z["nconmax"]), njmax=int(z["njmax"]),
                  mocap=(z["mocap_pos"], z["mocap_quat"]))
    for k in ("ncon", "ne", "nf", "nl", "nefc", "contact_geom", "efc_type", "efc_id", "efc_state"):
        np.testing.assert_array_equal(out[k], z[k], err_msg=k)
    np.testing.assert_array_equal(emu.slot(model, out, "xpos").reshape(n, -1, 3), z["xpos"])
    nviol, worst = util.qfrc_violations_scaled(out["qfrc_inverse"], z["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)


@pytest.mark.parametrize("case", ["sensors", "sensors2", "touch", "touch_elliptic", "geomdist", "geomdist_ccd", "actfrc"])
def test_sensordata(case):
    """d->sensordata of the reference's mj_inverse (mj_sensorPos / Vel / Acc, engine_sensor.c) for
    every sensor type evaluated on the device, cutoffs included (tests/golden/models/sensors.xml;
    touch.xml: touch sensors with site volumes of every shape, both cones)."""
    model, out, ref = _run(case)
    if case.startswith("touch"):
        # a touch reading is zero exactly where the reference's is (same contacts inside the zone)
        np.testing.assert_array_equal(out["sensordata"] > 0, ref["sensordata"] > 0)
    nviol, worst = util.sensor_violations(model, out["sensordata"], ref["sensordata"])
    assert nviol == 0, (nviol, worst)
    # position-stage sensors repeat the bit-identical kinematics of the CPU build
    stage = model.array("sensor_needstage").ravel()
    adr, dim = model.array("sensor_adr").ravel(), model.array("sensor_dim").ravel()
    typ = model.array("sensor_type").ravel()
    for a, d, s, t in zip(adr, dim, stage, typ):
        if s == 1 and t not in (34, 13, 40, 41):      # subtreecom, energies, projected moments: differently ordered sums
            np.testing.assert_array_equal(out["sensordata"][:, a:a + d], ref["sensordata"][:, a:a + d])


@pytest.mark.parametrize("name", ["camlight_cl", "humanoid_cl"])
def test_camlight(name):
    """cam_xpos / cam_xmat / light_xpos / light_xdir as mj_camlight leaves them inside mj_invPosition
    (engine_core_smooth.c:275-389), every mjtCamLight mode; bit-identical in the CPU build."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref, n, zr = util.camlight_fixture(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=zr)
    out = emu.run(model, qpos, qvel, qacc, camlight=True, dump=False)
    for k in ("cam_xpos", "cam_xmat", "light_xpos", "light_xdir"):
        np.testing.assert_array_equal(out[k].reshape(ref[k].shape), ref[k], err_msg=k)


@pytest.mark.parametrize("name", ["transmission_trn", "humanoid_trn", "arm26_trn", "slider_crank_trn", "adhesion_trn",
                                  "adhesion_elliptic_trn"])
def test_transmission(name):
    """actuator_length / actuator_moment (dense) / actuator_velocity as mj_transmission and mj_fwdVelocity
    leave them inside mj_inverse (engine_core_smooth.c:865-1346, engine_forward.c:216): every
    transmission type but the adhesion one, against the reference's compressed rows expanded."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref, n, zr = util.transmission_fixture(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=zr)
    out = emu.run(model, qpos, qvel, qacc, transmission=True, dump=False)
    nu, nv = model.int("nu"), model.int("nv")
    np.testing.assert_allclose(out["actuator_length"], ref["actuator_length"], rtol=1e-12, atol=1e-14)
    mom = out["actuator_moment"].reshape(n, nu, nv)
    np.testing.assert_allclose(mom, ref["actuator_moment"], rtol=1e-9, atol=1e-12)
    if name == "humanoid_trn":      # joint transmissions: the rows' structural zeros are exact zeros
        assert (mom[ref["actuator_moment"] == 0] == 0).all()
    np.testing.assert_allclose(out["actuator_velocity"], ref["actuator_velocity"], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("case", ["humanoid_energy", "zoo_energy", "tendons_energy"])
def test_energy(case):
    """d->energy of models with mjENBL_ENERGY (mj_energyPos / mj_energyVel inside mj_inverse,
    engine_inverse.c:210-223) against the reference's dump."""
    model, out, ref = _run(case)
    np.testing.assert_allclose(out["energy"], ref["energy"], rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(out["ncon"], ref["ncon"])


def test_candidate_list_contains_reference_contacts_in_order():
    """Every contact the reference reports (after its broadphase/midphase) appears in the static
    candidate list, in the same relative order (engine_collision_driver.c:265-484)."""
    for name in ("humanoid", "humanoids22"):
        import mujoco_inversedynamicstest_b200 as mjb
        path, ref = util.golden(name)
        model = mjb.Model.from_mjb(path)
        cand = emu.candidates(model)
        index = {(a, b): i for i, (a, b, f) in enumerate(cand)}
        assert len(index) == len(cand)          # no duplicate candidate
        for s in range(ref["contact_geom"].shape[0]):
            pairs = [tuple(p) for p in ref["contact_geom"][s] if p[0] >= 0]
            pos = [index[p] for p in pairs]
            assert pos == sorted(pos), (name, s)


def test_upload_rejections():
    import mujoco_inversedynamicstest_b200 as mjb
    model = mjb.Model.from_mjb(util.golden("slider_crank_nocontact")[0])
    assert len(emu.candidates(model)) == 0
    model.set_opt_int("disableflags", 0)          # contacts on: capsule-cylinder pairs -> mjc_Convex (GJK / EPA)
    assert len(emu.candidates(model)) > 0
    model.set_opt_int("enableflags", 1 << 4)      # mjENBL_MULTICCD: several contacts per convex pair
    with pytest.raises(RuntimeError, match="MULTICCD"):
        emu.candidates(model)
    model.set_opt_int("enableflags", 0)
    model.set_opt_int("disableflags", 1 << 16)    # mjDSBL_NATIVECCD: libccd's MPR
    with pytest.raises(RuntimeError, match="NATIVECCD"):
        emu.candidates(model)
    model.set_opt_int("disableflags", 0)
    model.array("geom_type")[1] = 7               # a mesh geom: no collision function here
    with pytest.raises(RuntimeError, match="collision function"):
        emu.candidates(model)
    # a sensor type that is not evaluated on the device is refused unless sensors are disabled
    sm = mjb.Model.from_mjb(util.golden("sensors")[0])
    assert len(emu.candidates(sm)) > 0
    sm.array("sensor_type")[0] = 44             # mjSENS_USER: filled by the mjcb_sensor callback, which does not exist on the device
    with pytest.raises(RuntimeError, match="sensor 0"):
        emu.candidates(sm)
    sm.set_opt_int("disableflags", 1 << 12)     # mjDSBL_SENSOR
    assert len(emu.candidates(sm)) > 0
    h = mjb.Model.from_mjb(util.golden("humanoid")[0])
    h.set_opt_int("enableflags", 1 << 3)        # mjENBL_INVDISCRETE: Euler, implicit, implicitfast
    for integrator in (0, 2, 3):
        h.set_opt_int("integrator", integrator)
        assert len(emu.candidates(h)) > 0
    h.set_opt_int("integrator", 1)              # RK4: an error in the reference too (engine_inverse.c:91-94)
    with pytest.raises(RuntimeError, match="INVDISCRETE"):
        emu.candidates(h)
    # velocity-dependent gains need ctrl / act, which are not inputs of the batched inverse
    h.set_opt_int("integrator", 3)
    h.array("actuator_gaintype")[0] = 1         # mjGAIN_AFFINE
    h.array("actuator_gainprm").reshape(-1, 10)[0, 2] = -0.5
    with pytest.raises(RuntimeError, match="velocity-dependent gain"):
        emu.candidates(h)


import edge_cases  # noqa: E402


@pytest.mark.parametrize("name", edge_cases.NAMES)
def test_reference_edge_case_models(name):
    """collisions.xml / ContactCount / FilterParent / collision_box / core_constraint (dof-less
    models, dense and sparse) / inertia.xml / tendon wrapping / sparse-forced humanoid: the
    reference's own edge-case models through the per-state pipeline, against the reference's dump
    and against the values the reference's tests hold for the default state."""
    import mujoco_inversedynamicstest_b200 as mjb
    path, ref = util.golden(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = edge_cases.states(model, ref)
    out = emu.run(model, qpos, qvel, qacc, nconmax=int(ref["nconmax"]), njmax=int(ref["njmax"]))
    assert (out["status"] == 0).all()
    edge_cases.check(name, out, ref)
    if "qM" in ref and out["qM"].size:
        # entries that cancel to zero carry the rounding of the largest entry of the matrix
        atol = 1e-12 * max(1.0, float(np.abs(ref["qM"]).max()))
        np.testing.assert_allclose(out["qM"], ref["qM"], rtol=1e-9, atol=atol)
        np.testing.assert_allclose(out["qLD"], ref["qLD"], rtol=1e-9, atol=atol)


def test_dense_and_sparse_jacobian_give_the_same_result():
    """pipeline_test.cc:39-71 (SparseDenseEquivalent): the humanoid with a sparse-forced Jacobian
    yields the reference's dense results; both dumps come from the reference."""
    _, dense = util.golden("humanoid")
    _, sparse = util.golden("ref_humanoid_sparse")
    n = int(sparse["nstate"])
    np.testing.assert_array_equal(dense["ncon"][:n], sparse["ncon"])
    np.testing.assert_array_equal(dense["nefc"][:n], sparse["nefc"])
    np.testing.assert_allclose(dense["qfrc_inverse"][:n], sparse["qfrc_inverse"], rtol=1e-9, atol=1e-9)


def test_fluid_forces_follow_the_passive_flag_and_refuse_implicit_invdiscrete():
    """mj_fluid runs inside mj_passive (engine_passive.c:436-462): with mjDSBL_PASSIVE the medium exerts
    nothing; mjENBL_INVDISCRETE with an implicit integrator would need the fluid velocity derivatives
    (mjd_passive_vel), which are not formed: refused at upload."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref = util.golden("fluid")
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, 16, z_range=tuple(ref["z_range"]))
    on = emu.run(model, qpos, qvel, qacc, nconmax=8, njmax=32)
    damping = -model.array("dof_damping").ravel() * qvel
    assert np.abs(on["qfrc_passive"] - damping).max() > 1.0          # the fish in water: tens of newtons
    model.set_opt_int("disableflags", model.get_opt_int("disableflags") | (1 << 5))    # mjDSBL_PASSIVE
    off = emu.run(model, qpos, qvel, qacc, nconmax=8, njmax=32)
    assert np.abs(off["qfrc_passive"]).max() == 0.0
    model.set_opt_int("disableflags", 0)
    model.set_opt_int("enableflags", 1 << 3)                          # mjENBL_INVDISCRETE
    model.set_opt_int("integrator", 3)                                # implicitfast
    with pytest.raises(RuntimeError, match="fluid"):
        emu.run(model, qpos, qvel, qacc, nconmax=8, njmax=32)
    model.set_opt_int("integrator", 0)                                # Euler: dof damping only, accepted
    emu.run(model, qpos, qvel, qacc, nconmax=8, njmax=32)


def _property_fixture(name):
    z = np.load(os.path.join(util.GOLDEN, name + ".npz"))
    return os.path.join(util.GOLDEN, name + ".mjb.gz"), {k: z[k] for k in z.files}


def test_reference_property_fluid_geoms_equivalent_to_bodies():
    """test/engine/engine_passive_test.cc:42-106: two fluid-interacting boxes as geoms of one floating body or on
    two child bodies of it give the same qfrc_passive (1e-14 in the reference's test); each model also against the
    reference's own output."""
    import mujoco_inversedynamicstest_b200 as mjb
    got = {}
    for name in ("ka_fluid_two_bodies", "ka_fluid_one_body"):
        path, z = _property_fixture(name)
        model = mjb.Model.from_mjb(path)
        out = emu.run(model, z["qpos"], z["qvel"], z["qacc"], nconmax=4, njmax=8)
        got[name] = out["qfrc_passive"][0]
        np.testing.assert_allclose(got[name], z["ref_qfrc_passive"][0], rtol=0, atol=1e-14)
    assert np.abs(got["ka_fluid_one_body"]).max() > 1.0
    np.testing.assert_allclose(got["ka_fluid_two_bodies"], got["ka_fluid_one_body"], rtol=0, atol=1e-14)


def test_reference_property_tendon_spring_deadband():
    """test/engine/engine_passive_test.cc:143-165: outside the deadband the spring force of the spatial tendon is
    stiffness * (springlength[1] - length), inside it is exactly zero."""
    import mujoco_inversedynamicstest_b200 as mjb
    path, z = _property_fixture("ka_tendon_deadband")
    model = mjb.Model.from_mjb(path)
    out = emu.run(model, z["qpos"], z["qvel"], z["qacc"], nconmax=4, njmax=8)
    length = out["sensordata"][0, 0]
    expected = model.array("tendon_stiffness").ravel()[0] * (model.array("tendon_lengthspring").ravel()[1] - length)
    assert expected == -5.0
    assert out["qfrc_passive"][0, 0] == expected
    assert out["qfrc_passive"][1, 0] == 0.0


@pytest.mark.parametrize("name", ["fluid", "fluid_box", "tendon_eq", "geomdist", "geomdist_ccd", "slider_crank", "convex"])
def test_live_reference_on_fresh_states(name):
    """4096 states that are in no fixture, against the reference run live (oracle/_ref): counters and equality /
    limit rows bit-identical, qfrc_inverse inside the element-wise bound, qfrc_passive and sensordata to rounding."""
    if not util.ref_available():
        pytest.skip("reference library not built")
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    from bench import load_reference_model
    path, gold = util.golden(name)
    model = mjb.Model.from_mjb(path)
    n = 4096
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(gold["z_range"]), first=3_000_000)
    ncm, njm = int(gold["nconmax"]), int(gold["njmax"])
    fields = {"ncon": 1, "nefc": 1, "efc_type": njm, "efc_state": njm, "qfrc_passive": None, "contact_geom": ncm,
              "contact_dist": ncm}
    if model.int("nsensordata") > 0:
        fields["sensordata"] = None
    ref, _ = load_reference_model(name).inverse_batch(qpos, qvel, qacc, fields=fields, nthread=4)
    out = emu.run(model, qpos, qvel, qacc, nconmax=ncm, njmax=njm)
    np.testing.assert_array_equal(out["contact_geom"], ref["contact_geom"])
    np.testing.assert_array_equal(out["contact_dist"], ref["contact_dist"][..., 0])     # bit-identical in the CPU build
    assert (out["status"] == 0).all()
    np.testing.assert_array_equal(out["ncon"], ref["ncon"])
    np.testing.assert_array_equal(out["nefc"], ref["nefc"])
    np.testing.assert_array_equal(out["efc_type"], ref["efc_type"][..., 0])
    np.testing.assert_array_equal(out["efc_state"], ref["efc_state"][..., 0])
    nviol, worst = util.qfrc_violations(out["qfrc_inverse"], ref["qfrc_inverse"])
    assert nviol == 0, (nviol, worst)
    np.testing.assert_allclose(out["qfrc_passive"], ref["qfrc_passive"][..., 0], rtol=1e-11, atol=1e-12)
    if "sensordata" in fields:
        nviol, worst = util.sensor_violations(model, out["sensordata"], ref["sensordata"][..., 0])
        assert nviol == 0, (nviol, worst)


def test_convex_pairs_in_degenerate_configurations():
    """GJK / EPA on their degenerate branches (coincident centres, coaxial cylinders, parallel faces, 1e-9 offsets; up
    to 20 contacts per state, penetrations of 0.15): contact sets, distances, positions and frames bit-identical to
    the reference run live."""
    if not util.ref_available():
        pytest.skip("reference library not built")
    import mujoco_inversedynamicstest_b200 as mjb
    from bench import load_reference_model
    model = mjb.Model.from_mjb(util.golden("convex")[0])
    qpos, qvel, qacc = util.convex_degenerate_states(model)
    ncm = 64
    ref, _ = load_reference_model("convex").inverse_batch(qpos, qvel, qacc, nthread=4, fields={
        "ncon": 1, "contact_geom": ncm, "contact_dist": ncm, "contact_pos": ncm, "contact_frame": ncm})
    out = emu.run(model, qpos, qvel, qacc, nconmax=ncm, njmax=400)
    assert (out["status"] == 0).all() and ref["ncon"].max() >= 15
    np.testing.assert_array_equal(out["ncon"], ref["ncon"])
    np.testing.assert_array_equal(out["contact_geom"], ref["contact_geom"])
    np.testing.assert_array_equal(out["contact_dist"], ref["contact_dist"][..., 0])
    np.testing.assert_array_equal(out["contact_pos"], ref["contact_pos"])
    np.testing.assert_array_equal(out["contact_frame"], ref["contact_frame"])
