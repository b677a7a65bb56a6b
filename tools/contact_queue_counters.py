import sys, ctypes, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import util
import mujoco_inversedynamicstest_b200 as mjb
from mujoco_inversedynamicstest_b200._lib import lib
from mujoco_inversedynamicstest_b200.states import generate_states
for name in ("humanoids22", "humanoid"):
    path, ref = util.golden(name)
    model = mjb.Model.from_mjb(path)
    n = int(ref["nstate"])
    qpos, qvel, qacc = generate_states(model, n, z_range=tuple(ref["z_range"]))
    bd = mjb.BatchData(model, n, outmask=mjb.OUT_COUNTS, nconmax=int(ref["nconmax"]), njmax=int(ref["njmax"]))
    bd.set_state(qpos, qvel, qacc); bd.inverse()
    out = (ctypes.c_int * 4)(); lib().mjb_debugQueue(bd._d, out)
    print(name, "nstate", n, "items,contacts,overflow,slots =", list(out), "mean ncon", float(np.mean(bd.counts()["ncon"])))
