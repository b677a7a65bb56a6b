"""The reference's own edge-case models, dumped by tests/golden/make_golden.py
(EDGE_CASES) from the reference library: state 0 is the model's default state -- what the reference's
tests evaluate -- the others are perturbed around it (states.near_default_states)."""
import glob
import os

import numpy as np

import util

NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(util.GOLDEN, "ref_*.npz")))

# Values the reference's tests hold for the default state (state 0 of the fixture):
#   engine_collision_driver_test.cc:52-68   AllCollisions: box-sphere_collides, box-sphere_predefined
#   :99-132 ContactCount: ncon == 8;  :134-171 FilterParent: 0 contacts, 1 with the filter disabled
#   :173-203 FilterParentDoesntAffectWorldBody: the pair collides
#   engine_collision_box_test.cc: deep penetration / duplicate removal leave 4 contacts
REFERENCE_HELD = {
    "ref_collisions": {"ncon": 2, "pairs": [(0, 1), (0, 2)]},
    "ref_contact_count": {"ncon": 8},
    "ref_filter_parent": {"ncon": 0},
    "ref_filter_parent_off": {"ncon": 1, "pairs": [(0, 3)]},
    "ref_filter_parent_world": {"ncon": 1, "pairs": [(0, 1)]},
    "ref_boxbox_deep": {"ncon": 4},
}

# fixtures with entries outside the element-wise 1e-9*|ref| + 1e-12 (count, worst ratio accepted):
# components that cancel between contact forces of 1e7-1e9 (eight spheres / deep box-box
# penetration / 650 contacts per state); every entry is within 1e-12 of its state's largest force
STRICT_EXCEPTIONS = {"ref_contact_count": (2, 1e5), "ref_boxbox_deep": (4, 1e3), "ref_planks": (32, 1e3)}


def states(model, ref):
    from mujoco_inversedynamicstest_b200.states import generate_states, near_default_states
    gen = near_default_states if bool(ref["near_default"]) else generate_states
    return gen(model, int(ref["nstate"]))


def check(name, out, ref):
    """out: dict with the discrete outputs and qfrc_inverse of the path under test."""
    for k in ("ncon", "ne", "nf", "nl", "nefc", "contact_geom", "contact_dim", "contact_exclude",
              "contact_efc_address", "efc_type", "efc_id", "efc_state"):
        np.testing.assert_array_equal(out[k], ref[k], err_msg=f"{name}: {k}")
    got, want = out["qfrc_inverse"], ref["qfrc_inverse"]
    if got.size:
        nviol, worst = util.qfrc_violations(got, want)
        max_viol, max_ratio = STRICT_EXCEPTIONS.get(name, (0, 1.0))
        assert nviol <= max_viol and worst <= max_ratio, (name, nviol, worst)
        smax = np.maximum(np.abs(want).max(axis=1, keepdims=True), 1.0)
        assert float((np.abs(got - want) / smax).max()) < 1e-12, name
    held = REFERENCE_HELD.get(name)
    if held:
        assert int(out["ncon"][0]) == held["ncon"], name
        if "pairs" in held:
            pairs = sorted(tuple(sorted(p)) for p in out["contact_geom"][0][: held["ncon"]].tolist())
            assert pairs == held["pairs"], (name, pairs)
