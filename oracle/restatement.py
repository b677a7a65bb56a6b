"""ctypes driver of oracle/mjinv_oracle.c, the plain-C restatement of the reference's mj_inverse
path (TEST INFRASTRUCTURE: see the header of that file)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mjinv_oracle.c")
LIB = os.path.join(HERE, "lib", "libmjinv_oracle.so")

_INT_SCALARS = ["nq", "nv", "nbody", "njnt", "ngeom", "ntendon", "nwrap", "nexclude", "nM", "nC"]
_INT_ARRAYS = ["body_parentid", "body_rootid", "body_weldid", "body_jntnum", "body_jntadr",
               "body_dofnum", "body_dofadr", "body_geomnum", "body_geomadr", "body_sameframe",
               "body_contype", "body_conaffinity", "body_bvhadr",
               "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_limited",
               "dof_bodyid", "dof_jntid", "dof_parentid", "dof_Madr", "dof_simplenum",
               "geom_type", "geom_bodyid", "geom_contype", "geom_conaffinity", "geom_condim",
               "geom_priority", "geom_sameframe",
               "tendon_adr", "tendon_num", "tendon_limited", "wrap_objid", "exclude_signature"]
_NUM_ARRAYS = ["qpos0", "qpos_spring", "body_pos", "body_quat", "body_ipos", "body_iquat",
               "body_mass", "body_inertia", "body_invweight0",
               "jnt_pos", "jnt_axis", "jnt_stiffness", "jnt_range", "jnt_margin", "jnt_solref",
               "jnt_solimp",
               "dof_armature", "dof_damping", "dof_frictionloss", "dof_invweight0", "dof_solref",
               "dof_solimp", "dof_M0",
               "geom_size", "geom_rbound", "geom_pos", "geom_quat", "geom_friction", "geom_margin",
               "geom_gap", "geom_solmix", "geom_solref", "geom_solimp",
               "tendon_range", "tendon_margin", "tendon_stiffness", "tendon_damping",
               "tendon_lengthspring", "tendon_invweight0", "tendon_solref_lim", "tendon_solimp_lim",
               "wrap_prm"]


class OrcModel(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int) for n in _INT_SCALARS] +
                [("disableflags", ctypes.c_int), ("cone", ctypes.c_int),
                 ("timestep", ctypes.c_double), ("impratio", ctypes.c_double),
                 ("gravity", ctypes.c_double * 3)] +
                [(n, ctypes.c_void_p) for n in _INT_ARRAYS] +
                [(n, ctypes.c_void_p) for n in _NUM_ARRAYS])


class OrcOut(ctypes.Structure):
    _fields_ = [("qfrc_inverse", ctypes.c_void_p), ("qM", ctypes.c_void_p), ("qLD", ctypes.c_void_p),
                ("qLDiagInv", ctypes.c_void_p), ("counts", ctypes.c_void_p),
                ("contact_geom", ctypes.c_void_p), ("maxcon", ctypes.c_int),
                ("efc_type", ctypes.c_void_p), ("efc_id", ctypes.c_void_p),
                ("efc_force", ctypes.c_void_p), ("maxefc", ctypes.c_int)]


def build(force=False):
    if os.path.exists(LIB) and not force and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                    "-fvisibility=hidden", "-D_GNU_SOURCE", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.orc_inverse.argtypes = [ctypes.c_void_p] * 5
        assert L.orc_sizeof_model() == ctypes.sizeof(OrcModel), "struct layout mismatch"
        _lib = L
    return _lib


class Restatement:
    """Binds a model (any object with .int/.array and option getters) to the C restatement."""

    def __init__(self, model, opt):
        """opt: dict with disableflags, cone, timestep, impratio, gravity."""
        self.keep = [model]
        m = OrcModel()
        for n in _INT_SCALARS:
            setattr(m, n, model.int(n))
        m.disableflags, m.cone = int(opt["disableflags"]), int(opt["cone"])
        m.timestep, m.impratio = float(opt["timestep"]), float(opt["impratio"])
        for i in range(3):
            m.gravity[i] = float(opt["gravity"][i])
        for n in _INT_ARRAYS:
            a = np.array(model.array(n), dtype=np.int32, order='C', copy=True)
            self.keep.append(a)
            setattr(m, n, a.ctypes.data)
        for n in _NUM_ARRAYS:
            a = np.array(model.array(n), dtype=np.float64, order='C', copy=True)
            self.keep.append(a)
            setattr(m, n, a.ctypes.data)
        self.m = m
        self.nv, self.nM, self.nC = model.int("nv"), model.int("nM"), model.int("nC")

    def inverse_batch(self, qpos, qvel, qacc, maxcon=64, maxefc=256, inertia=True):
        n = qpos.shape[0]
        out = {"qfrc_inverse": np.zeros((n, self.nv)), "qM": np.zeros((n, self.nM)),
               "qLD": np.zeros((n, self.nC)), "qLDiagInv": np.zeros((n, self.nv)),
               "counts": np.zeros((n, 5), np.int32), "contact_geom": np.zeros((n, maxcon, 2), np.int32),
               "efc_type": np.zeros((n, maxefc), np.int32), "efc_id": np.zeros((n, maxefc), np.int32),
               "efc_force": np.zeros((n, maxefc))}
        qpos, qvel, qacc = (np.ascontiguousarray(x, dtype=np.float64) for x in (qpos, qvel, qacc))
        L = lib()
        for i in range(n):
            o = OrcOut(out["qfrc_inverse"][i].ctypes.data,
                       out["qM"][i].ctypes.data if inertia else None,
                       out["qLD"][i].ctypes.data if inertia else None,
                       out["qLDiagInv"][i].ctypes.data if inertia else None,
                       out["counts"][i].ctypes.data, out["contact_geom"][i].ctypes.data, maxcon,
                       out["efc_type"][i].ctypes.data, out["efc_id"][i].ctypes.data,
                       out["efc_force"][i].ctypes.data, maxefc)
            rc = L.orc_inverse(ctypes.byref(self.m), qpos[i].ctypes.data, qvel[i].ctypes.data,
                               qacc[i].ctypes.data, ctypes.byref(o))
            if rc:
                raise RuntimeError(f"orc_inverse failed on state {i}: rc={rc}")
        c = out["counts"]
        out.update(ncon=c[:, 0], ne=c[:, 1], nf=c[:, 2], nl=c[:, 3], nefc=c[:, 4])
        return out
