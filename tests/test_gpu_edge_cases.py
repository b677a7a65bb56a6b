"""The reference's edge-case models on the CUDA path, through the C-ABI (see tests/edge_cases.py):
engine_collision_driver_test.cc:52-203, engine_collision_box_test.cc, engine_core_constraint_test.cc:
231-251, engine_core_smooth_test.cc:466-511, pipeline_test.cc:39-71."""
import numpy as np
import pytest

import edge_cases
import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernels", ["generic", "specialised"])
@pytest.mark.parametrize("name", edge_cases.NAMES)
def test_reference_edge_case_models(name, kernels):
    import mujoco_inversedynamicstest_b200 as mjb
    path, ref = util.golden(name)
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = edge_cases.states(model, ref)
    n = int(ref["nstate"])
    bd = mjb.BatchData(model, n, outmask=mjb.OUT_QFRC | mjb.OUT_COUNTS | mjb.OUT_CONTACT | mjb.OUT_EFC | mjb.OUT_INERTIA,
                       nconmax=int(ref["nconmax"]), njmax=int(ref["njmax"]))
    if kernels == "specialised":
        try:
            bd.specialize()
        except mjb.MjbError as exc:
            pytest.skip(f"not specialised: {exc}")
    bd.set_state(qpos, qvel, qacc)
    assert bd.inverse() == 0
    cnt, con, efc = bd.counts(), bd.contacts(), bd.efc()
    out = {k: cnt[k] for k in ("ncon", "ne", "nf", "nl", "nefc")}
    out.update({"contact_geom": con["geom"], "contact_dim": con["dim"], "contact_exclude": con["exclude"],
                "contact_efc_address": con["efc_address"], "efc_type": efc["type"], "efc_id": efc["id"],
                "efc_state": efc["state"], "qfrc_inverse": bd.qfrc_inverse()})
    edge_cases.check(name, out, ref)
    if model.int("nv"):
        atol = 1e-12 * max(1.0, float(np.abs(ref["qM"]).max()))
        np.testing.assert_allclose(bd.get(mjb.F_QM), ref["qM"], rtol=1e-9, atol=atol)
        np.testing.assert_allclose(bd.get(mjb.F_QLD), ref["qLD"], rtol=1e-9, atol=atol)
