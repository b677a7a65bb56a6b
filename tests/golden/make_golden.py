"""Generate the committed parity fixtures from the reference itself (run where /root/reference and
oracle/_ref/libmujoco_ref.so exist):

    python tests/golden/make_golden.py

For every case: the compiled mjModel as a gzip'd MJB (mj_saveModel) and, for NSTATE seeded states
of mujoco_inversedynamicstest_b200.states.generate_states, the reference's own mj_inverse outputs
(qfrc_inverse, counters, contact geom pairs, efc_type/efc_id, efc_force, qM/qLD/qLDiagInv ...).
The reference publishes no numeric golden vector for qfrc_inverse, so these dumps
of the unmodified CPU engine are the pin; this script is how they were made.
"""
import gzip
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reflib  # noqa: E402
from mujoco_inversedynamicstest_b200.states import generate_states, near_default_states  # noqa: E402

DSBL_CONTACT, DSBL_EQUALITY = 1 << 4, 1 << 1

# name -> (reference xml, {opt int overrides}, nstate, z_range, nconmax, njmax)
CASES = {
    "humanoid": ("model/humanoid/humanoid.xml", {}, 256, (0.0, 1.5), 64, 256),
    "humanoid_elliptic": ("model/humanoid/humanoid.xml", {"cone": 1}, 256, (0.0, 1.5), 64, 256),
    "humanoid_nocontact": ("model/humanoid/humanoid.xml", {"disableflags": DSBL_CONTACT}, 256,
                           (2.0, 3.0), 64, 256),
    "humanoids22": ("model/humanoid/22_humanoids.xml", {}, 8, (0.0, 1.5), 640, 1200),
    "slider_crank_nocontact": ("model/slider_crank/slider_crank.xml", {"disableflags": DSBL_CONTACT},
                               64, (0.0, 1.5), 8, 16),
    "inverse_test": ("src/inverse/test.xml", {}, 64, (0.0, 1.5), 8, 16),
    # BASELINE config 4: tendons, equality constraints, joint limits
    "arm26": ("model/tendon_arm/arm26.xml", {}, 128, (0.0, 1.5), 8, 16),
    "weld": ("test/engine/testdata/weld.xml", {}, 64, (0.0, 1.5), 8, 64),
    "connect": ("test/engine/testdata/connect.xml", {}, 64, (0.0, 1.5), 8, 32),
    # this repository's own coverage scene (tests/golden/models/zoo.xml)
    "zoo": ("repo:tests/golden/models/zoo.xml", {}, 256, (0.0, 0.6), 32, 128),
    "zoo_elliptic": ("repo:tests/golden/models/zoo.xml", {"cone": 1}, 256, (0.0, 0.6), 32, 128),
    # mjENBL_INVDISCRETE with the Euler integrator (mj_discreteAcc, engine_inverse.c:81-164)
    "humanoid_invdiscrete": ("model/humanoid/humanoid.xml", {"enableflags": 1 << 3}, 128, (0.0, 1.5), 64, 256),
    "capsbox": ("repo:tests/golden/models/capsbox.xml", {}, 512, (0.0, 0.7), 48, 200),
    "capsbox_elliptic": ("repo:tests/golden/models/capsbox.xml", {"cone": 1}, 256, (0.0, 0.7), 48, 200),
    "boxes": ("repo:tests/golden/models/boxes.xml", {}, 1024, (0.0, 0.7), 64, 300),
    "boxes_elliptic": ("repo:tests/golden/models/boxes.xml", {"cone": 1}, 256, (0.0, 0.7), 64, 300),
    "tendons": ("repo:tests/golden/models/tendons.xml", {}, 512, (0.3, 1.3), 8, 32),
    "gravcomp": ("repo:tests/golden/models/gravcomp.xml", {}, 128, (0.5, 1.5), 8, 16),
    # mocap bodies at their model pose (what mj_makeData leaves in mjData); moved per state: MOCAP_CASES
    "mocap": ("repo:tests/golden/models/mocap.xml", {}, 128, (0.0, 0.6), 16, 96),
    # every sensor type evaluated on the device (mj_sensorPos / Vel / Acc); sensordata is dumped too
    "sensors": ("repo:tests/golden/models/sensors.xml", {}, 256, (0.0, 0.6), 16, 96),
    # touch sensors (contact list + contact-row forces + ray / site-volume tests), both cones
    "touch": ("repo:tests/golden/models/touch.xml", {}, 256, (0.0, 0.35), 24, 160),
    "touch_elliptic": ("repo:tests/golden/models/touch.xml", {"cone": 1}, 256, (0.0, 0.35), 24, 160),
    # mjENBL_ENERGY: d->energy from mj_energyPos / mj_energyVel inside mj_inverse (engine_inverse.c:210-223)
    "humanoid_energy": ("model/humanoid/humanoid.xml", {"enableflags": 1 << 1}, 128, (0.0, 1.5), 64, 256),
    "zoo_energy": ("repo:tests/golden/models/zoo.xml", {"enableflags": 1 << 1}, 128, (0.0, 0.6), 32, 128),
    "tendons_energy": ("repo:tests/golden/models/tendons.xml", {"enableflags": 1 << 1}, 128, (0.3, 1.3), 8, 32),
    # cameras and lights in every mjtCamLight mode (outputs of mj_camlight: CAMLIGHT_CASES)
    "camlight": ("repo:tests/golden/models/camlight.xml", {}, 128, (0.2, 1.2), 16, 64),
    # mjENBL_INVDISCRETE with the implicitfast integrator (mj_discreteAcc + mjd_actuator_vel / mjd_passive_vel)
    "humanoid_invdiscrete_fast": ("model/humanoid/humanoid.xml", {"enableflags": 1 << 3, "integrator": 3}, 128,
                                  (0.0, 1.5), 64, 256),
    "implicitfast": ("repo:tests/golden/models/implicitfast.xml", {}, 128, (0.2, 1.2), 16, 64),
    # ... and with the implicit integrator (adds mjd_rne_vel, the full dof-dof pattern)
    "humanoid_invdiscrete_implicit": ("model/humanoid/humanoid.xml", {"enableflags": 1 << 3, "integrator": 2}, 128,
                                      (0.0, 1.5), 64, 256),
    "implicit": ("repo:tests/golden/models/implicitfast.xml", {"integrator": 2}, 128, (0.2, 1.2), 16, 64),
    # sensors that read mj_camlight / mj_transmission outputs, magnetometer, clock
    "sensors2": ("repo:tests/golden/models/sensors2.xml", {}, 256, (0.2, 1.2), 16, 64),
    # adhesion actuators (mjTRN_BODY): moments from the contact normals, both cones
    "adhesion": ("repo:tests/golden/models/adhesion.xml", {}, 256, (0.0, 0.3), 48, 200),
    "adhesion_elliptic": ("repo:tests/golden/models/adhesion.xml", {"cone": 1}, 128, (0.0, 0.3), 48, 200),
    # one actuator per branch of mj_transmission (outputs: TRANSMISSION_CASES)
    "transmission": ("repo:tests/golden/models/transmission.xml", {}, 128, (0.2, 1.2), 16, 64),
    # mj_fluid: inertia-box and ellipsoid models in a dense, viscous medium with wind; Stokes terms alone
    "fluid": ("repo:tests/golden/models/fluid.xml", {}, 256, (0.0, 1.2), 8, 32),
    "fluid_box": ("repo:tests/golden/models/fluid_box.xml", {}, 128, (0.5, 1.5), 8, 16),
    # equality constraints on spatial tendons (offset, quartic coupling, coupling with a fixed tendon)
    "tendon_eq": ("repo:tests/golden/models/tendon_eq.xml", {}, 256, (0.3, 1.3), 8, 32),
    # geom-distance sensors (distance / normal / fromto) over primitive pairs, geom-geom and body-body
    "geomdist": ("repo:tests/golden/models/geomdist.xml", {}, 256, (0.0, 0.8), 8, 16),
    "geomdist_ccd": ("repo:tests/golden/models/geomdist_ccd.xml", {}, 256, (0.0, 0.8), 8, 16),
    # actuator-force sensors read what a fresh mjData holds (mj_inverse computes no actuation)
    "actfrc": ("repo:tests/golden/models/actfrc.xml", {}, 32, (0.0, 1.0), 8, 16),
    # geom pairs of mjc_Convex (GJK / EPA): the reference's slider-crank model with contacts on, and a scene with
    # every convex pair type
    "slider_crank": ("model/slider_crank/slider_crank.xml", {}, 256, (0.0, 1.5), 16, 64),
    "convex": ("repo:tests/golden/models/convex.xml", {}, 1024, (0.1, 0.6), 48, 200),
}


def make_case(name):
    xml, opts, nstate, zr, nconmax, njmax = CASES[name]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]) if xml.startswith("repo:")
                              else reflib.reference_path(xml))
    for k, v in opts.items():
        m.set_opt_int(k, m.get_opt_int(k) | v if k == "disableflags" else v)
    raw = os.path.join(HERE, name + ".mjb")
    m.save_mjb(raw)
    with open(raw, "rb") as f, gzip.GzipFile(os.path.join(HERE, name + ".mjb.gz"), "wb",
                                             compresslevel=9, mtime=0) as g:
        g.write(f.read())
    os.remove(raw)

    qpos, qvel, qacc = generate_states(m, nstate, z_range=zr)
    fields = {"ncon": 1, "ne": 1, "nf": 1, "nl": 1, "nefc": 1,
              "contact_geom": nconmax, "contact_dist": nconmax, "contact_dim": nconmax,
              "contact_exclude": nconmax, "contact_efc_address": nconmax, "contact_pos": nconmax,
              "contact_frame": nconmax,
              "efc_type": njmax, "efc_id": njmax, "efc_state": njmax, "efc_force": njmax,
              "efc_pos": njmax, "efc_D": njmax, "efc_aref": njmax,
              "efc_margin": njmax, "efc_R": njmax, "efc_vel": njmax, "efc_diagApprox": njmax,
              "qfrc_passive": None, "qfrc_constraint": None, "qM": None, "qLD": None,
              "qLDiagInv": None, "xpos": None, "cvel": None, "cdof": None}
    if m.int("nsensordata") > 0:
        fields["sensordata"] = None
    if m.get_opt_int("enableflags") & (1 << 1):
        fields["energy"] = None
    out, _ = m.inverse_batch(qpos, qvel, qacc, fields=fields)
    assert out["ncon"].max() <= nconmax and out["nefc"].max() <= njmax, (
        name, out["ncon"].max(), out["nefc"].max())
    save = {"z_range": np.array(zr), "nstate": np.array(nstate), "nconmax": np.array(nconmax),
            "njmax": np.array(njmax)}
    for k, v in out.items():
        save[k] = v[..., 0] if (v.ndim == 3 and v.shape[2] == 1) else v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
    print(f"{name}: nv={m.int('nv')} states={nstate} mean ncon={out['ncon'].mean():.2f} "
          f"mean nefc={out['nefc'].mean():.2f} max={out['ncon'].max()}/{out['nefc'].max()}")


# Large multi-tree scenes (BASELINE config 5) with enough states for statistics: only the discrete
# outputs and qfrc_inverse are kept (the full dump of every contact frame would be tens of MB).
# name -> (reference xml, nstate, z_range, nconmax, njmax)
REDUCED_CASES = {
    "humanoids22_256": ("model/humanoid/22_humanoids.xml", 256, (0.0, 1.5), 704, 1408),
    # 100 humanoids: nv = 2700, 1901 geoms, 1.8 M candidate pairs, ~4,400 contacts per state
    "humanoids100_4": ("model/humanoid/100_humanoids.xml", 4, (0.0, 1.5), 5632, 7168),
}


def make_reduced_case(name):
    xml, nstate, zr, nconmax, njmax = REDUCED_CASES[name]
    m = reflib.Model.from_xml(reflib.reference_path(xml))
    raw = os.path.join(HERE, name + ".mjb")
    m.save_mjb(raw)
    with open(raw, "rb") as f, gzip.GzipFile(os.path.join(HERE, name + ".mjb.gz"), "wb",
                                             compresslevel=9, mtime=0) as g:
        g.write(f.read())
    os.remove(raw)
    qpos, qvel, qacc = generate_states(m, nstate, z_range=zr)
    fields = {"ncon": 1, "ne": 1, "nf": 1, "nl": 1, "nefc": 1, "contact_geom": nconmax,
              "efc_type": njmax, "efc_id": njmax, "efc_state": njmax}
    out, _ = m.inverse_batch(qpos, qvel, qacc, fields=fields, nthread=os.cpu_count() or 1)
    assert out["ncon"].max() <= nconmax and out["nefc"].max() <= njmax, (name, out["ncon"].max(), out["nefc"].max())
    save = {"z_range": np.array(zr), "nstate": np.array(nstate), "nconmax": np.array(nconmax), "njmax": np.array(njmax)}
    for k, v in out.items():
        v = v[..., 0] if (v.ndim == 3 and v.shape[2] == 1) else v
        save[k] = v.astype(np.int16) if (v.dtype == np.int32 and k.startswith(("contact_", "efc_"))) else v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
    print(f"{name}: nv={m.int('nv')} states={nstate} mean ncon={out['ncon'].mean():.1f} "
          f"mean nefc={out['nefc'].mean():.1f} max={out['ncon'].max()}/{out['nefc'].max()}")


# The reference's own edge-case models (what pins results in the reference's tests), run
# through mj_inverse at the default state (state 0, what those tests evaluate) and at perturbed
# states around it. name -> (xml, {opt overrides}, nstate, nconmax, njmax)
T = "test/engine/testdata/"
DSBL_FILTERPARENT, JAC_DENSE, JAC_SPARSE = 1 << 9, 0, 1
EDGE_CASES = {
    # engine_collision_driver_test.cc:52-68 (pairs), :99-132 (ncon == 8), :134-203 (parent filter)
    "ref_collisions": (T + "collisions.xml", {}, 64, 8, 32),
    "ref_contact_count": ("repo:tests/golden/models/contact_count.xml", {}, 64, 16, 64),
    "ref_filter_parent": ("repo:tests/golden/models/filter_parent.xml", {}, 32, 8, 32),
    "ref_filter_parent_off": ("repo:tests/golden/models/filter_parent.xml", {"disableflags": DSBL_FILTERPARENT}, 32, 8, 32),
    "ref_filter_parent_world": ("repo:tests/golden/models/filter_parent_world.xml", {}, 32, 8, 32),
    "ref_midphase": (T + "collision_driver/midphase.xml", {}, 16, 256, 1024),
    "ref_planks": (T + "collision_driver/planks.xml", {}, 8, 1536, 6144),
    # engine_collision_box_test.cc:40-277 (bad / duplicate contact removal, deep penetration: ncon == 4)
    "ref_boxbox_bad0": (T + "collision_box/boxbox_bad0.xml", {}, 64, 64, 256),
    "ref_boxbox_deep": (T + "collision_box/boxbox_deep.xml", {}, 64, 16, 64),
    "ref_boxbox_duplicate": (T + "collision_box/boxbox_duplicate.xml", {}, 64, 16, 64),
    "ref_sphere_cylinder": (T + "collision_primitive/sphere_cylinder.xml", {}, 64, 16, 64),
    # engine_core_constraint_test.cc:231-251 (dof-less / bilateral-margin models, dense and sparse)
    "ref_dofless_contact": (T + "core_constraint/dofless_contact.xml", {"jacobian": JAC_DENSE}, 4, 8, 32),
    "ref_dofless_contact_sparse": (T + "core_constraint/dofless_contact.xml", {"jacobian": JAC_SPARSE}, 4, 8, 32),
    "ref_dofless_tendon_frictional": (T + "core_constraint/dofless_tendon_frictional.xml", {}, 4, 8, 32),
    "ref_dofless_tendon_limited": (T + "core_constraint/dofless_tendon_limited.xml", {}, 4, 8, 32),
    "ref_dofless_tendon_limitedmargin": (T + "core_constraint/dofless_tendon_limitedmargin.xml", {}, 4, 8, 32),
    "ref_dofless_weld": (T + "core_constraint/dofless_weld.xml", {"jacobian": JAC_DENSE}, 32, 8, 64),
    "ref_dofless_weld_sparse": (T + "core_constraint/dofless_weld.xml", {"jacobian": JAC_SPARSE}, 32, 8, 64),
    "ref_joint_limited_bilateral_margin": (T + "core_constraint/joint_limited_bilateral_margin.xml", {}, 64, 8, 32),
    "ref_tendon_limited_bilateral_margin": (T + "core_constraint/tendon_limited_bilateral_margin.xml", {}, 64, 8, 32),
    "ref_soft_weld": (T + "core_constraint/soft_weld.xml", {}, 64, 8, 32),
    # engine_core_smooth_test.cc:466-511 (L'DL == M on inertia.xml), tendon wrapping
    "ref_inertia": (T + "inertia.xml", {}, 64, 16, 64),
    "ref_tendon_wrap_cylinder": (T + "core_smooth/tendon_wrap_cylinder.xml", {}, 64, 16, 64),
    "ref_tendon_wrap_sphere": (T + "core_smooth/tendon_wrap_sphere.xml", {}, 64, 16, 64),
    # pipeline_test.cc:39-71: dense == sparse Jacobian semantics on the humanoid
    "ref_humanoid_sparse": ("model/humanoid/humanoid.xml", {"jacobian": JAC_SPARSE}, 64, 64, 256),
}


def make_edge_case(name):
    xml, opts, nstate, nconmax, njmax = EDGE_CASES[name]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]) if xml.startswith("repo:")
                              else reflib.reference_path(xml))
    for k, v in opts.items():
        m.set_opt_int(k, m.get_opt_int(k) | v if k == "disableflags" else v)
    raw = os.path.join(HERE, name + ".mjb")
    m.save_mjb(raw)
    with open(raw, "rb") as f, gzip.GzipFile(os.path.join(HERE, name + ".mjb.gz"), "wb",
                                             compresslevel=9, mtime=0) as g:
        g.write(f.read())
    os.remove(raw)
    if name == "ref_humanoid_sparse":
        qpos, qvel, qacc = generate_states(m, nstate)
    else:
        qpos, qvel, qacc = near_default_states(m, nstate)
    fields = {"ncon": 1, "ne": 1, "nf": 1, "nl": 1, "nefc": 1, "contact_geom": nconmax, "contact_dist": nconmax,
              "contact_dim": nconmax, "contact_exclude": nconmax, "contact_efc_address": nconmax,
              "efc_type": njmax, "efc_id": njmax, "efc_state": njmax, "efc_force": njmax, "efc_pos": njmax,
              "qfrc_constraint": None, "qM": None, "qLD": None, "qLDiagInv": None}
    out, _ = m.inverse_batch(qpos, qvel, qacc, fields=fields)
    assert out["ncon"].max() <= nconmax and out["nefc"].max() <= njmax, (name, out["ncon"].max(), out["nefc"].max())
    save = {"nstate": np.array(nstate), "nconmax": np.array(nconmax), "njmax": np.array(njmax),
            "near_default": np.array(name != "ref_humanoid_sparse")}
    for k, v in out.items():
        v = v[..., 0] if (v.ndim == 3 and v.shape[2] == 1) else v
        save[k] = v.astype(np.int16) if (v.dtype == np.int32 and k.startswith(("contact_", "efc_"))) else v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
    print(f"{name}: nv={m.int('nv')} ngeom={m.int('ngeom')} states={nstate} ncon at default={out['ncon'][0]} "
          f"mean ncon={out['ncon'].mean():.2f} mean nefc={out['nefc'].mean():.2f} max={out['ncon'].max()}/{out['nefc'].max()}")


# finite-difference Jacobians of the reference (mjd_inverseFD, engine_derivative_fd.c:611) on the
# first states of a case's stream: name -> (case whose model / state stream is used, nstate)
FD_CASES = {"humanoid_fd": ("humanoid", 6), "zoo_fd": ("zoo", 6), "humanoid_nocontact_fd": ("humanoid_nocontact", 6),
            # with the sensor Jacobians DsDq / DsDv / DsDa (62 sensors of every stage)
            "sensors_fd": ("sensors", 6)}
FD_EPS = 1e-6


def make_fd_case(name):
    base, nstate = FD_CASES[name]
    xml, opts, _, zr, _, _ = CASES[base]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]) if xml.startswith("repo:")
                              else reflib.reference_path(xml))
    for k, v in opts.items():
        m.set_opt_int(k, m.get_opt_int(k) | v if k == "disableflags" else v)
    qpos, qvel, qacc = generate_states(m, nstate, z_range=zr)
    if m.int("nsensordata") > 0:
        dq, dv, da, sq, sv, sa = m.inverse_fd_sensor_batch(qpos, qvel, qacc, FD_EPS)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), base=np.array(base), nstate=np.array(nstate),
                            z_range=np.array(zr), eps=np.array(FD_EPS), DfDq=dq, DfDv=dv, DfDa=da,
                            DsDq=sq, DsDv=sv, DsDa=sa)
        print(f"{name}: nv={m.int('nv')} states={nstate} max|DsDq|={np.abs(sq).max():.3g} max|DsDa|={np.abs(sa).max():.3g}")
        return
    dq, dv, da, dm = m.inverse_fd_batch(qpos, qvel, qacc, FD_EPS, mass=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), base=np.array(base), nstate=np.array(nstate),
                        z_range=np.array(zr), eps=np.array(FD_EPS), DfDq=dq, DfDv=dv, DfDa=da, DmDq=dm)
    print(f"{name}: nv={m.int('nv')} states={nstate} max|DfDq|={np.abs(dq).max():.3g}")


# mj_rnePostConstraint after mj_inverse (engine_core_smooth.c:2027-2181; what mj_sensorAcc runs for
# accelerometer / force / torque sensors): cacc, cfrc_int, cfrc_ext -- and qfrc_bias of mj_fwdVelocity
# (engine_forward.c:228) -- on the first states of a case's stream: name -> (case whose model / state stream is used, nstate)
POST_CASES = {"humanoid_post": ("humanoid", 128), "humanoid_elliptic_post": ("humanoid_elliptic", 64),
              "humanoids22_post": ("humanoids22", 4), "weld_post": ("weld", 64),
              "connect_post": ("connect", 64), "zoo_post": ("zoo", 128), "capsbox_post": ("capsbox", 64),
              "boxes_post": ("boxes", 64), "gravcomp_post": ("gravcomp", 32),
              "humanoid_nocontact_post": ("humanoid_nocontact", 64), "tendons_post": ("tendons", 32),
              "arm26_post": ("arm26", 32)}


def make_post_case(name):
    base, nstate = POST_CASES[name]
    xml, opts, _, zr, _, _ = CASES[base]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]) if xml.startswith("repo:")
                              else reflib.reference_path(xml))
    for k, v in opts.items():
        m.set_opt_int(k, m.get_opt_int(k) | v if k == "disableflags" else v)
    qpos, qvel, qacc = generate_states(m, nstate, z_range=zr)
    out, _ = m.inverse_batch(qpos, qvel, qacc, fields={"cacc": None, "cfrc_int": None, "cfrc_ext": None,
                                                          "qfrc_bias": None})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), base=np.array(base), nstate=np.array(nstate),
                        z_range=np.array(zr), **out)
    print(f"{name}: nbody={m.int('nbody')} states={nstate} max|cfrc_ext|={np.abs(out['cfrc_ext']).max():.3g}")


# mj_forward + mj_compareFwdInv (engine_inverse.c:275-316) on the first states of a case's stream,
# with random ctrl, qfrc_applied and xfrc_applied; for every second state qacc is moved away from
# the forward solution so that the two norms are also compared where they are large
FWDINV_CASES = {"humanoid_fwdinv": ("humanoid", 48), "zoo_fwdinv": ("zoo", 48), "weld_fwdinv": ("weld", 32),
                "humanoid_nocontact_fwdinv": ("humanoid_nocontact", 16)}


def make_fwdinv_case(name):
    base, nstate = FWDINV_CASES[name]
    xml, opts, _, zr, _, _ = CASES[base]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]) if xml.startswith("repo:")
                              else reflib.reference_path(xml))
    for k, v in opts.items():
        m.set_opt_int(k, m.get_opt_int(k) | v if k == "disableflags" else v)
    qpos, qvel, qacc = generate_states(m, nstate, z_range=zr)
    rng = np.random.RandomState(20250331)
    nv, nu, nb = m.int("nv"), m.int("nu"), m.int("nbody")
    ctrl = rng.uniform(-1, 1, (nstate, nu)) if nu else None
    qfrc_applied = rng.normal(0, 1, (nstate, nv))
    xfrc = rng.normal(0, 5, (nstate, nb, 6)) * (rng.uniform(0, 1, (nstate, nb, 1)) < 0.5)
    xfrc[:, 0] = 0
    dqacc = 0.1 * qacc * (np.arange(nstate) % 2)[:, None]
    out = m.fwdinv_batch(qpos, qvel, ctrl, qfrc_applied, xfrc, dqacc)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), base=np.array(base), nstate=np.array(nstate),
                        z_range=np.array(zr), qfrc_applied=qfrc_applied, xfrc_applied=xfrc, **out)
    print(f"{name}: states={nstate} fwdinv at the solution max {out['fwdinv'][0::2].max(axis=0)}, "
          f"away from it max {out['fwdinv'][1::2].max(axis=0)}")


# per-state mocap poses (d->mocap_pos / mocap_quat as inputs of mj_kinematics): name -> (case, nstate)
MOCAP_CASES = {"mocap_moved": ("mocap", 128)}


def mocap_poses(m, nstate):
    """Seeded per-state mocap poses around the model pose: +-0.15 m, ~+-0.5 rad, quaternions NOT normalised."""
    rng = np.random.RandomState(20250331)
    nb = m.int("nbody")
    ids = [b for b in range(nb) if m.array("body_mocapid").ravel()[b] >= 0]
    ids.sort(key=lambda b: m.array("body_mocapid").ravel()[b])
    pos0 = m.array("body_pos").reshape(nb, 3)[ids]
    quat0 = m.array("body_quat").reshape(nb, 4)[ids]
    pos = pos0[None] + rng.uniform(-0.15, 0.15, (nstate, len(ids), 3))
    quat = quat0[None] * rng.uniform(0.8, 1.3, (nstate, len(ids), 1)) + rng.uniform(-0.25, 0.25, (nstate, len(ids), 4))
    return pos, quat


def make_mocap_case(name):
    base, nstate = MOCAP_CASES[name]
    xml, opts, _, zr, nconmax, njmax = CASES[base]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]))
    qpos, qvel, qacc = generate_states(m, nstate, z_range=zr)
    pos, quat = mocap_poses(m, nstate)
    fields = {"ncon": 1, "ne": 1, "nf": 1, "nl": 1, "nefc": 1, "contact_geom": nconmax, "efc_type": njmax,
              "efc_id": njmax, "efc_state": njmax, "efc_force": njmax, "xpos": None, "xquat": None,
              "qfrc_constraint": None}
    out, _ = m.inverse_batch(qpos, qvel, qacc, fields=fields, mocap=(pos, quat))
    assert out["ncon"].max() <= nconmax and out["nefc"].max() <= njmax
    np.savez_compressed(os.path.join(HERE, name + ".npz"), base=np.array(base), nstate=np.array(nstate),
                        z_range=np.array(zr), nconmax=np.array(nconmax), njmax=np.array(njmax),
                        mocap_pos=pos, mocap_quat=quat,
                        **{k: (v[..., 0] if (v.ndim == 3 and v.shape[2] == 1) else v) for k, v in out.items()})
    print(f"{name}: states={nstate} mean ncon={out['ncon'].mean():.2f} mean nefc={out['nefc'].mean():.2f}")


# mj_camlight inside mj_invPosition (engine_core_smooth.c:275-389): cam_xpos, cam_xmat, light_xpos,
# light_xdir on the first states of a case's stream: name -> (case whose model / stream is used, nstate)
CAMLIGHT_CASES = {"camlight_cl": ("camlight", 128), "humanoid_cl": ("humanoid", 64)}


def make_camlight_case(name):
    base, nstate = CAMLIGHT_CASES[name]
    xml, opts, _, zr, _, _ = CASES[base]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]) if xml.startswith("repo:")
                              else reflib.reference_path(xml))
    qpos, qvel, qacc = generate_states(m, nstate, z_range=zr)
    out, _ = m.inverse_batch(qpos, qvel, qacc, fields={"cam_xpos": None, "cam_xmat": None, "light_xpos": None,
                                                          "light_xdir": None})
    out.pop("qfrc_inverse")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), base=np.array(base), nstate=np.array(nstate),
                        z_range=np.array(zr), **out)
    print(f"{name}: ncam={m.int('ncam')} nlight={m.int('nlight')} states={nstate}")


# mj_transmission inside mj_invPosition (engine_core_smooth.c:865-1346) and actuator_velocity of
# mj_fwdVelocity (engine_forward.c:216): actuator_length, the compressed actuator_moment expanded to the
# dense nu x nv matrix, actuator_velocity: name -> (case whose model / stream is used, nstate)
TRANSMISSION_CASES = {"transmission_trn": ("transmission", 128), "humanoid_trn": ("humanoid", 32),
                      "arm26_trn": ("arm26", 64), "slider_crank_trn": ("slider_crank_nocontact", 64),
                      "adhesion_trn": ("adhesion", 256), "adhesion_elliptic_trn": ("adhesion_elliptic", 128)}


def make_transmission_case(name):
    base, nstate = TRANSMISSION_CASES[name]
    xml, opts, _, zr, _, _ = CASES[base]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]) if xml.startswith("repo:")
                              else reflib.reference_path(xml))
    for k, v in opts.items():
        m.set_opt_int(k, m.get_opt_int(k) | v if k == "disableflags" else v)
    qpos, qvel, qacc = generate_states(m, nstate, z_range=zr)
    out, _ = m.inverse_batch(qpos, qvel, qacc, fields={"actuator_length": None, "actuator_velocity": None,
                                                          "actuator_moment": None, "moment_rownnz": None,
                                                          "moment_rowadr": None, "moment_colind": None})
    nu, nv = m.int("nu"), m.int("nv")
    dense = np.zeros((nstate, nu, nv))
    for s in range(nstate):
        for i in range(nu):
            adr, nnz = int(out["moment_rowadr"][s, i, 0]), int(out["moment_rownnz"][s, i, 0])
            dense[s, i, out["moment_colind"][s, adr:adr + nnz, 0]] = out["actuator_moment"][s, adr:adr + nnz, 0]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), base=np.array(base), nstate=np.array(nstate),
                        z_range=np.array(zr), actuator_length=out["actuator_length"][..., 0],
                        actuator_velocity=out["actuator_velocity"][..., 0], actuator_moment=dense)
    print(f"{name}: nu={nu} nv={nv} states={nstate} max|moment|={np.abs(dense).max():.3g} "
          f"mean nnz={out['moment_rownnz'].mean():.2f}")


# per-state d->xfrc_applied in the mj_rnePostConstraint outputs (engine_core_smooth.c:2039-2049) and in the
# force / torque sensors that read cfrc_int: name -> (case whose model / state stream is used, nstate)
XFRC_CASES = {"humanoid_xfrc": ("humanoid", 64), "sensors_xfrc": ("sensors", 64), "weld_xfrc": ("weld", 32),
              "humanoids22_xfrc": ("humanoids22", 4)}


def xfrc_samples(m, nstate):
    """Seeded applied wrenches [nstate, nbody, 6]: about half of the bodies loaded, the world body never."""
    rng = np.random.RandomState(20250331)
    nb = m.int("nbody")
    x = rng.normal(0, 5, (nstate, nb, 6)) * (rng.uniform(0, 1, (nstate, nb, 1)) < 0.5)
    x[:, 0] = 0
    return x


def make_xfrc_case(name):
    base, nstate = XFRC_CASES[name]
    xml, opts, _, zr, _, _ = CASES[base]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]) if xml.startswith("repo:")
                              else reflib.reference_path(xml))
    for k, v in opts.items():
        m.set_opt_int(k, m.get_opt_int(k) | v if k == "disableflags" else v)
    qpos, qvel, qacc = generate_states(m, nstate, z_range=zr)
    fields = {"cacc": None, "cfrc_int": None, "cfrc_ext": None}
    if m.int("nsensordata") > 0:
        fields["sensordata"] = None
    out, _ = m.inverse_batch(qpos, qvel, qacc, fields=fields, xfrc=xfrc_samples(m, nstate))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), base=np.array(base), nstate=np.array(nstate),
                        z_range=np.array(zr), **{k: (v[..., 0] if (v.ndim == 3 and v.shape[2] == 1) else v)
                                                 for k, v in out.items()})
    print(f"{name}: nbody={m.int('nbody')} states={nstate} max|cfrc_ext|={np.abs(out['cfrc_ext']).max():.3g}")


# per-state d->eq_active (mj_instantiateEquality skips inactive constraints; every later row moves up):
# name -> (case whose model / state stream is used, nstate)
EQACTIVE_CASES = {"zoo_eqactive": ("zoo", 128), "weld_eqactive": ("weld", 64), "connect_eqactive": ("connect", 64),
                  "mocap_eqactive": ("mocap", 64)}


def eq_active_samples(m, nstate):
    """Seeded per-state flags [nstate, neq]: each constraint on with probability 0.6; state 0 all off, state 1 all on."""
    rng = np.random.RandomState(20250331)
    e = (rng.uniform(0, 1, (nstate, m.int("neq"))) < 0.6).astype(np.uint8)
    e[0] = 0
    e[1] = 1
    return e


def make_eqactive_case(name):
    base, nstate = EQACTIVE_CASES[name]
    xml, opts, _, zr, nconmax, njmax = CASES[base]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]) if xml.startswith("repo:")
                              else reflib.reference_path(xml))
    for k, v in opts.items():
        m.set_opt_int(k, m.get_opt_int(k) | v if k == "disableflags" else v)
    qpos, qvel, qacc = generate_states(m, nstate, z_range=zr)
    fields = {"ncon": 1, "ne": 1, "nf": 1, "nl": 1, "nefc": 1, "contact_geom": nconmax, "contact_efc_address": nconmax,
              "efc_type": njmax, "efc_id": njmax, "efc_state": njmax, "efc_force": njmax, "efc_pos": njmax,
              "qfrc_constraint": None, "cacc": None, "cfrc_int": None, "cfrc_ext": None}
    out, _ = m.inverse_batch(qpos, qvel, qacc, fields=fields, eq_active=eq_active_samples(m, nstate))
    assert out["ncon"].max() <= nconmax and out["nefc"].max() <= njmax
    np.savez_compressed(os.path.join(HERE, name + ".npz"), base=np.array(base), nstate=np.array(nstate),
                        z_range=np.array(zr), nconmax=np.array(nconmax), njmax=np.array(njmax),
                        **{k: (v[..., 0] if (v.ndim == 3 and v.shape[2] == 1) else v) for k, v in out.items()})
    print(f"{name}: neq={m.int('neq')} states={nstate} ne min/mean/max={out['ne'].min()}/{out['ne'].mean():.2f}/{out['ne'].max()}")


# Known-answer tests the reference itself holds, restated on the inverse path: the model is settled /
# evaluated by the reference as its test does, the state (qpos, qvel, qacc) is kept, and the EXPECTED values are the
# ones written in the reference's test or model file -- not an output of the reference.
#   rne_post/*: test/engine/engine_core_smooth_test.cc:160-300 (1000 steps, force / torque sensors == sensor_user, 1e-6)
#   the others: tests/golden/models/ref_*.xml (each cites its test)
RNE_POST_DIR = "test/engine/testdata/core_smooth/rne_post/"
KNOWN_ANSWER_CASES = {
    **{"ka_connect_" + n: (RNE_POST_DIR + "connect/" + n + ".xml", 1000, "sensor_user", 1e-6)
       for n in ("force_free", "force_slide", "force_slide_rotated", "multiple_constraints", "torque_free")},
    **{"ka_weld_" + n: (RNE_POST_DIR + "weld/" + n + ".xml", 1000, "sensor_user", 1e-6)
       for n in ("force_free", "force_free_rotated", "force_torque_free", "force_torque_free_rotated",
                 "force_torque_free_rotated_tendon", "tfratio0_force_free", "tfratio0_force_slide",
                 "tfratio0_force_slide_rotated", "tfratio0_multiple_constraints", "tfratio0_torque_free")},
    # engine_sensor_test.cc:428-455: 2*3*5 at the default state, 7*3*5 at qpos[2] = 7 (EXPECT_EQ)
    "ka_potential_energy": ("repo:tests/golden/models/ref_sensor_potential_energy.xml", 0, [[30.0], [105.0]], 0.0),
    # engine_sensor_test.cc:400-426: data->energy[0] == 2*3*5 with mjENBL_ENERGY (EXPECT_EQ)
    "ka_enable_energy": ("repo:tests/golden/models/ref_sensor_enable_energy.xml", 0, "energy0=30", 0.0),
    # engine_sensor_test.cc:595-634: pixels within 1e-4
    "ka_camprojection": ("repo:tests/golden/models/ref_sensor_camprojection.xml", 0,
                         [[0.0, 0.0, 1920.0, 1200.0, 960.0, 600.0]], 1e-4),
    # engine_sensor_test.cc:221-322 (hand-picked velocities, DoubleNear(1e-14)); the third element sets qvel
    "ka_framevel_linear": ("repo:tests/golden/models/ref_sensor_framevel_linear.xml", [2 ** 0.5, 1.0],
                           [[-(0.5 ** 0.5), 0.5 ** 0.5, 0.0]], 1e-14),
    "ka_framevel_angfixed": ("repo:tests/golden/models/ref_sensor_framevel_angfixed.xml", [1.0], [[0.0, 0.0, 0.0]], 1e-14),
    "ka_framevel_angopposing": ("repo:tests/golden/models/ref_sensor_framevel_angopposing.xml", [-1.0, 1.0],
                                [[0.0, 2.0, 0.0]], 1e-14),
    # engine_ray_test.cc:79-165 (0.9 / 2.9, EXPECT_FLOAT_EQ) through rangefinder sites
    "ka_ray": ("repo:tests/golden/models/ref_ray.xml", 0, [[2.9, 0.9, 0.9]], 1e-6),
}


def make_known_answer_case(name):
    xml, nstep, expected, tol = KNOWN_ANSWER_CASES[name]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]) if xml.startswith("repo:") else reflib.reference_path(xml))
    raw = os.path.join(HERE, name + ".mjb")
    m.save_mjb(raw)
    with open(raw, "rb") as f, gzip.GzipFile(os.path.join(HERE, name + ".mjb.gz"), "wb", compresslevel=9, mtime=0) as g:
        g.write(f.read())
    os.remove(raw)
    set_qvel = nstep if isinstance(nstep, list) else None      # a list in place of the step count: qvel of the test
    qpos, qvel, qacc, sd = m.settle(0 if set_qvel else nstep)
    if set_qvel:
        qvel = np.array(set_qvel)
        qacc = np.zeros_like(qvel)      # velocity-stage sensors do not read it
        out, _ = m.inverse_batch(qpos[None], qvel[None], qacc[None], fields={"sensordata": None})
        sd = out["sensordata"][0, :, 0]
    qpos, qvel, qacc = qpos[None], qvel[None], qacc[None]
    if name == "ka_potential_energy":           # second state of the reference's test: the body lifted to z = 7
        q2 = qpos.copy(); q2[0, 2] = 7
        qpos, qvel, qacc = np.vstack([qpos, q2]), np.vstack([qvel, qvel]), np.vstack([qacc, qacc])
    save = {"qpos": qpos, "qvel": qvel, "qacc": qacc, "tol": np.array(tol), "ref_sensordata": sd}
    if expected == "sensor_user":
        ns, nus = m.int("nsensor"), m.int("nuser_sensor")
        save["expected"] = m.array("sensor_user").reshape(ns, nus)[:, :3].copy()[None]    # [1, nsensor, 3]
        save["sensor_adr"] = m.array("sensor_adr").ravel().copy()
    elif isinstance(expected, str):
        save["expected_energy0"] = np.array(30.0)
    else:
        save["expected"] = np.array(expected)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
    print(f"{name}: nq={m.int('nq')} nv={m.int('nv')} nsensordata={m.int('nsensordata')} reference reads {np.round(sd, 6)}")


# Properties the reference's tests hold for mj_passive (test/engine/engine_passive_test.cc), restated on the inverse
# path: the state of the test, the reference's own qfrc_passive next to it, the property checked by the tests here.
#   fluid_*: :42-106, two models whose qfrc_passive must agree to 1e-14 (qvel = 1..6, quaternion 0.5 0.5 0.5 0.5)
#   tendon_deadband: :143-165, spring force == stiffness * (springlength[1] - length) outside the deadband
#     (EXPECT_EQ), exactly 0 inside (qpos[0] = -1)
PROPERTY_CASES = {
    "ka_fluid_two_bodies": "repo:tests/golden/models/ref_fluid_two_bodies.xml",
    "ka_fluid_one_body": "repo:tests/golden/models/ref_fluid_one_body.xml",
    "ka_tendon_deadband": "test/engine/testdata/tendon_springlength.xml",
}


def make_property_case(name):
    xml = PROPERTY_CASES[name]
    m = reflib.Model.from_xml(os.path.join(ROOT, xml[5:]) if xml.startswith("repo:") else reflib.reference_path(xml))
    raw = os.path.join(HERE, name + ".mjb")
    m.save_mjb(raw)
    with open(raw, "rb") as f, gzip.GzipFile(os.path.join(HERE, name + ".mjb.gz"), "wb", compresslevel=9, mtime=0) as g:
        g.write(f.read())
    os.remove(raw)
    nq, nv = m.int("nq"), m.int("nv")
    qpos = m.array("qpos0").reshape(1, nq).copy()
    if name.startswith("ka_fluid"):
        qpos[0, 3:7] = 0.5
        qvel = np.arange(1.0, 7.0)[None]
    else:
        qpos = np.vstack([qpos, qpos]); qpos[1, 0] = -1
        qvel = np.zeros((2, nv))
    qacc = np.zeros_like(qvel)
    fields = {"qfrc_passive": None}
    if m.int("nsensordata") > 0:
        fields["sensordata"] = None
    out, _ = m.inverse_batch(qpos, qvel, qacc, fields=fields)
    save = {"qpos": qpos, "qvel": qvel, "qacc": qacc}
    for k, v in out.items():
        save["ref_" + k] = v[..., 0] if (v.ndim == 3 and v.shape[2] == 1) else v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
    print(f"{name}: nq={nq} nv={nv} reference qfrc_passive {np.round(save['ref_qfrc_passive'], 6)}")


if __name__ == "__main__":
    for case in (sys.argv[1:] or list(CASES) + list(FD_CASES) + list(POST_CASES) + list(FWDINV_CASES) +
                 list(MOCAP_CASES) + list(REDUCED_CASES) + list(EDGE_CASES) + list(CAMLIGHT_CASES) +
                 list(TRANSMISSION_CASES) + list(XFRC_CASES) + list(EQACTIVE_CASES) +
                 list(KNOWN_ANSWER_CASES) + list(PROPERTY_CASES)):
        (make_property_case if case in PROPERTY_CASES else make_known_answer_case if case in KNOWN_ANSWER_CASES else make_eqactive_case if case in EQACTIVE_CASES else make_xfrc_case if case in XFRC_CASES else make_transmission_case if case in TRANSMISSION_CASES else make_camlight_case if case in CAMLIGHT_CASES else make_edge_case if case in EDGE_CASES else make_reduced_case if case in REDUCED_CASES else make_fd_case if case in FD_CASES else make_post_case if case in POST_CASES else
         make_fwdinv_case if case in FWDINV_CASES else make_mocap_case if case in MOCAP_CASES else
         make_case)(case)
