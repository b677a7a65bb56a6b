"""One mjbData sharded over several devices (mjb_makeDataMulti). The same device may
be listed twice, which exercises the sharding logic on a single GPU; with two GPUs present the
second shard really lives on device 1."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]


@pytest.mark.parametrize("n", [4096, 4097, 1])
def test_sharded_batch_equals_single_device(n):
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states
    path, ref = util.golden("humanoid")
    model = mjb.Model.from_mjb(path)
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(ref["z_range"]))
    one = mjb.BatchData(model, n, outmask=mjb.OUT_COUNTS | mjb.OUT_CONTACT, nconmax=64)
    one.set_state(qpos, qvel, qacc)
    assert one.inverse() == 0
    multi = mjb.BatchData(model, n, outmask=mjb.OUT_COUNTS | mjb.OUT_CONTACT, nconmax=64, devices=_devices())
    multi.set_state(qpos, qvel, qacc)
    assert multi.inverse() == 0
    np.testing.assert_array_equal(multi.qfrc_inverse(), one.qfrc_inverse())
    np.testing.assert_array_equal(multi.counts()["ncon"], one.counts()["ncon"])
    np.testing.assert_array_equal(multi.contacts()["geom"], one.contacts()["geom"])
    # host-to-host entry point: one host thread per device
    out = multi.inverse_host_arrays(qpos, qvel, qacc)
    np.testing.assert_array_equal(out, one.qfrc_inverse())
    # status flags are counted over all shards
    bad = qpos.copy()
    bad[n - 1, 0] = np.nan
    multi.set_state(bad, qvel, qacc)
    assert multi.inverse() == 1
    assert multi.status()[n - 1] & mjb.STATUS_BADQPOS


def test_single_device_only_calls_are_refused_on_a_sharded_batch():
    import mujoco_inversedynamicstest_b200 as mjb
    path, _ = util.golden("humanoid_nocontact")
    model = mjb.Model.from_mjb(path)
    multi = mjb.BatchData(model, 64, devices=_devices())
    assert multi.device_ptr(mjb.F_QFRC_INVERSE) is None
    with pytest.raises(mjb.MjbError):
        multi.inverse_fd()
