"""Host logic: the deterministic state generator and the MJB model ingestion."""
import os

import numpy as np
import pytest

import util


def _model(name="humanoid"):
    import mujoco_inversedynamicstest_b200 as mjb
    return mjb.Model.from_mjb(util.golden(name)[0])


def test_generator_is_counter_based():
    from mujoco_inversedynamicstest_b200.states import generate_states
    m = _model()
    qpos, qvel, qacc = generate_states(m, 1000)
    a, b, c = generate_states(m, 300, first=500)
    np.testing.assert_array_equal(qpos[500:800], a)
    np.testing.assert_array_equal(qvel[500:800], b)
    np.testing.assert_array_equal(qacc[500:800], c)
    q2, _, _ = generate_states(m, 1000, seed=7)
    assert not np.array_equal(qpos, q2)


def test_generator_ranges():
    from mujoco_inversedynamicstest_b200.states import generate_states
    m = _model()
    qpos, qvel, qacc = generate_states(m, 4096)
    assert np.abs(qvel).max() <= 1 and np.abs(qacc).max() <= 10
    np.testing.assert_allclose(np.linalg.norm(qpos[:, 3:7], axis=1), 1, atol=1e-14)
    assert qpos[:, 2].min() >= 0 and qpos[:, 2].max() <= 1.5
    rng = m.array("jnt_range").reshape(-1, 2)
    adr = m.array("jnt_qposadr").ravel()
    lim = m.array("jnt_limited").ravel()
    for j in range(1, m.int("njnt")):
        assert lim[j]
        lo, hi = rng[j]
        q = qpos[:, adr[j]]
        assert q.min() >= lo - 0.1 * (hi - lo) - 1e-12 and q.max() <= hi + 0.1 * (hi - lo) + 1e-12
        assert (q < lo).any() and (q > hi).any()     # both limit sides get exercised


def test_mjb_loader_sizes():
    m = _model()
    assert (m.int("nq"), m.int("nv"), m.int("nbody"), m.int("njnt"), m.int("ngeom")) == (28, 27, 17, 22, 20)
    assert m.int("nM") == 243 and m.int("nC") == 243 and m.int("ntendon") == 2
    m22 = _model("humanoids22")
    assert (m22.int("nv"), m22.int("nbody"), m22.int("ngeom")) == (594, 353, 419)
    assert _model("humanoid_elliptic").get_opt_int("cone") == 1
    assert _model("humanoid_nocontact").get_opt_int("disableflags") & 16


def test_mjb_loader_rejects_garbage(tmp_path):
    import mujoco_inversedynamicstest_b200 as mjb
    p = tmp_path / "bad.mjb"
    p.write_bytes(b"\x00" * 100)
    with pytest.raises(mjb.MjbError, match="header"):
        mjb.Model.from_mjb(str(p))
    with pytest.raises(mjb.MjbError, match="cannot open"):
        mjb.Model.from_mjb(str(tmp_path / "missing.mjb"))


@pytest.mark.skipif(not util.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["humanoid", "humanoids22", "slider_crank_nocontact"])
def test_mjb_loader_matches_reference_loader(name, tmp_path):
    """Every array the upload reads is identical whether the MJB is read by libmjb or by the
    reference's own mj_loadModel."""
    import gzip
    from oracle import reflib
    raw = tmp_path / "m.mjb"
    raw.write_bytes(gzip.open(util.golden(name)[0], "rb").read())
    rm = reflib.Model.from_mjb(str(raw))
    m = _model(name)
    for arr in ("body_parentid", "body_pos", "body_quat", "body_inertia", "jnt_type", "jnt_axis",
                "jnt_range", "dof_parentid", "dof_Madr", "geom_size", "geom_type", "geom_friction",
                "geom_solref", "tendon_range", "wrap_prm", "body_invweight0", "qpos0", "geom_rbound"):
        np.testing.assert_array_equal(m.array(arr), rm.array(arr), err_msg=arr)
    for k in ("nq", "nv", "nbody", "nM", "nC", "npair", "nexclude", "nbvh"):
        assert m.int(k) == rm.int(k)
