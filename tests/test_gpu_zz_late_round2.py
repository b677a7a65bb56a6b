"""GPU parity tests of the paths added after the last GPU run of round 2 (fluid forces, equality constraints on
spatial tendons, geom-distance and actuator-force sensors, convex geom pairs through GJK / EPA). Every comparison here
passes in the CPU build of the same pipeline (tests/test_hostemu_pipeline.py); the file sorts after the other GPU test
files so that `pytest -x` reaches it last."""
import os

import numpy as np
import pytest

import util
from test_gpu_parity import _run
from test_gpu_parity_scale import _reference_model, _report

pytestmark = pytest.mark.gpu

CASES = ["fluid", "fluid_box", "tendon_eq", "slider_crank", "convex"]


@pytest.mark.parametrize("name", CASES)
def test_golden_discrete_outputs_bit_exact(name):
    import test_gpu_parity
    test_gpu_parity.test_golden_discrete_outputs_bit_exact(name)


@pytest.mark.parametrize("name", CASES)
def test_golden_qfrc_inverse_within_tolerance(name):
    import test_gpu_parity
    test_gpu_parity.test_golden_qfrc_inverse_within_tolerance(name)


@pytest.mark.parametrize("kernels", ["generic", "specialised"])
@pytest.mark.parametrize("name", CASES)
def test_strict_elementwise_bound(name, kernels):
    import test_gpu_parity_scale
    test_gpu_parity_scale.test_strict_elementwise_bound(name, kernels)


def _property_batch(mjb, name, outmask):
    z = np.load(os.path.join(util.GOLDEN, name + ".npz"))
    model = mjb.Model.from_mjb(os.path.join(util.GOLDEN, name + ".mjb.gz"))
    n = z["qpos"].shape[0]
    bd = mjb.BatchData(model, n, outmask=outmask, nconmax=4, njmax=8)
    bd.set_state(z["qpos"], z["qvel"], z["qacc"])
    assert bd.inverse() == 0
    return model, bd, z


def test_reference_property_fluid_geoms_equivalent_to_bodies():
    """test/engine/engine_passive_test.cc:42-106: the two fluid-interacting boxes as geoms of the floating body or on
    two child bodies of it give the same qfrc_passive (1e-14 there, and here); each also against the reference."""
    import mujoco_inversedynamicstest_b200 as mjb
    got = {}
    for name in ("ka_fluid_two_bodies", "ka_fluid_one_body"):
        model, bd, z = _property_batch(mjb, name, mjb.OUT_QFRC)
        got[name] = bd.get(mjb.F_QFRC_PASSIVE)[0]
        np.testing.assert_allclose(got[name], z["ref_qfrc_passive"][0], rtol=0, atol=1e-14)
    np.testing.assert_allclose(got["ka_fluid_two_bodies"], got["ka_fluid_one_body"], rtol=0, atol=1e-14)


def test_reference_property_tendon_spring_deadband():
    """test/engine/engine_passive_test.cc:143-165: stiffness * (springlength[1] - length) outside the deadband of
    the spatial tendon's spring, exactly zero inside."""
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, z = _property_batch(mjb, "ka_tendon_deadband", mjb.OUT_QFRC)
    length = bd.sensordata()[0, 0]
    expected = model.array("tendon_stiffness").ravel()[0] * (model.array("tendon_lengthspring").ravel()[1] - length)
    qp = bd.get(mjb.F_QFRC_PASSIVE)
    assert abs(qp[0, 0] - expected) <= 1e-14 * abs(expected) and expected == -5.0
    assert qp[1, 0] == 0.0


def test_golden_actuator_force_sensors_read_zero():
    """actuatorfrc / jointactuatorfrc under mj_inverse copy d->actuator_force / d->qfrc_actuator, which the inverse
    path never computes: zeros of a fresh mjData, with the neighbouring readings in their places."""
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, "actfrc", True, 0)
    assert nbad == 0
    got = bd.sensordata()
    nviol, worst = util.sensor_violations(model, got, ref["sensordata"])
    assert nviol == 0, (nviol, worst)
    assert (got[:, [1, 3, 4]] == 0).all() and (got[:, [0, 2, 5]] != 0).all()


@pytest.mark.parametrize("name", ["geomdist", "geomdist_ccd"])
def test_golden_geom_distance_sensors(name):
    """distance / normal / fromto sensors (engine_sensor.c:378-463, mj_geomDistance engine_support.c:1406-1452)
    over primitive geom pairs, geom-geom and body-body, cutoffs reached and not (geomdist), and over the pairs
    the reference measures with mjc_ccd: box-box and the convex pairs (geomdist_ccd; GJK with the cutoff, EPA
    when penetrating)."""
    import mujoco_inversedynamicstest_b200 as mjb
    model, bd, ref, nbad, _ = _run(mjb, name, True, 0)
    assert nbad == 0
    nviol, worst = util.sensor_violations(model, bd.sensordata(), ref["sensordata"])
    assert nviol == 0, (nviol, worst)
    assert (ref["sensordata"] < 0).any() and (ref["sensordata"] == 2.0).any()      # penetrations and cutoffs occur


@pytest.mark.skipif(not util.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kernels", ["specialised", "generic"])
@pytest.mark.parametrize("name", ["fluid", "fluid_box", "tendon_eq", "geomdist", "geomdist_ccd", "slider_crank", "convex"])
def test_live_reference_passive_and_sensor_paths(name, kernels):
    """Fluid forces, equality constraints on spatial tendons and geom-distance sensors on 2^16 states that are in
    no fixture, against the reference run live: counters and row types / states bit-identical, qfrc_inverse inside
    the element-wise bound, qfrc_passive and sensordata within 1e-9 rel / 1e-12 abs."""
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, gold = util.golden(name)
    rm = _reference_model(path)
    model = mjb.Model.from_mjb(path)
    n, nconmax, njmax = 1 << 16, int(gold["nconmax"]), int(gold["njmax"])
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(gold["z_range"]), first=5_000_000)
    fields = {"ncon": 1, "nefc": 1, "efc_type": njmax, "efc_state": njmax, "qfrc_passive": None, "contact_geom": nconmax}
    if model.int("nsensordata") > 0:
        fields["sensordata"] = None
    ref, _ = rm.inverse_batch(qpos, qvel, qacc, fields=fields, nthread=max(1, len(os.sched_getaffinity(0))))
    bd = mjb.BatchData(model, n, outmask=mjb.OUT_QFRC | mjb.OUT_COUNTS | mjb.OUT_EFC | mjb.OUT_CONTACT, nconmax=nconmax,
                       njmax=njmax)
    if kernels == "specialised":
        try:
            bd.specialize()
        except mjb.MjbError as exc:
            pytest.skip(f"not specialised: {exc}")
    bd.set_state(qpos, qvel, qacc)
    assert bd.inverse() == 0
    cnt = bd.counts()
    np.testing.assert_array_equal(cnt["ncon"], ref["ncon"])
    np.testing.assert_array_equal(bd.contacts()["geom"], ref["contact_geom"])
    np.testing.assert_array_equal(cnt["nefc"], ref["nefc"])
    efc = bd.efc()
    np.testing.assert_array_equal(efc["type"], ref["efc_type"][..., 0])
    np.testing.assert_array_equal(efc["state"], ref["efc_state"][..., 0])
    got = bd.qfrc_inverse()
    nviol, worst = util.qfrc_violations(got, ref["qfrc_inverse"])
    _report(name + "_live_64k", {"kernels": kernels, "states": n, "entries": int(got.size), "strict_viol": nviol,
                                 "strict_worst_ratio": worst})
    assert nviol == 0, (nviol, worst)
    np.testing.assert_allclose(bd.get(mjb.F_QFRC_PASSIVE), ref["qfrc_passive"][..., 0], rtol=1e-9, atol=1e-12)
    if "sensordata" in fields:
        nviol, worst = util.sensor_violations(model, bd.sensordata(), ref["sensordata"][..., 0])
        assert nviol == 0, (nviol, worst)


@pytest.mark.skipif(not util.ref_available(), reason="oracle/_ref not built")
def test_convex_pairs_in_degenerate_configurations():
    """GJK / EPA on their degenerate branches (coincident centres, coaxial cylinders, parallel faces, 1e-9 offsets; up
    to 20 contacts per state): contact sets bit-identical to the reference run live, distances / positions / frames
    within 1e-9 rel / 1e-12 abs."""
    import mujoco_inversedynamicstest_b200 as mjb
    path, _ = util.golden("convex")
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = util.convex_degenerate_states(model)
    ncm = 64
    ref, _ = _reference_model(path).inverse_batch(qpos, qvel, qacc, nthread=4, fields={
        "ncon": 1, "contact_geom": ncm, "contact_dist": ncm, "contact_pos": ncm, "contact_frame": ncm})
    bd = mjb.BatchData(model, len(qpos), outmask=mjb.OUT_COUNTS | mjb.OUT_CONTACT, nconmax=ncm, njmax=400)
    bd.set_state(qpos, qvel, qacc)
    assert bd.inverse() == 0
    np.testing.assert_array_equal(bd.counts()["ncon"], ref["ncon"])
    con = bd.contacts()
    np.testing.assert_array_equal(con["geom"], ref["contact_geom"])
    np.testing.assert_allclose(con["dist"], ref["contact_dist"][..., 0], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(con["pos"], ref["contact_pos"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(con["frame"], ref["contact_frame"], rtol=1e-9, atol=1e-11)
