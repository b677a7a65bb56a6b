"""ctypes access to the UNMODIFIED reference engine built by oracle/Makefile (oracle/_ref/libmujoco_ref.so).

TEST INFRASTRUCTURE ONLY. Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports this module.

The library is the reference's own src/engine + src/user + src/xml + src/thread compiled from
/root/reference where they lie (see oracle/Makefile), plus oracle/ref_harness.c which provides
name-based array access through the reference's X-macro lists and the threaded batch loop
`for i: copy state -> mj_inverse -> copy outputs` that mjb_inverse() replaces.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmujoco_ref.so")

_CODES = {0: np.float64, 1: np.int32, 2: np.uint8, 3: np.float32}


class _Request(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("out", ctypes.c_void_p), ("maxrows", ctypes.c_int)]


_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `make -C oracle` where /root/reference exists")
        L = ctypes.CDLL(LIB_PATH)
        L.mj_loadXML.restype = ctypes.c_void_p
        L.mj_loadXML.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        L.mj_loadModel.restype = ctypes.c_void_p
        L.mj_loadModel.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
        L.mj_saveModel.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]
        L.mj_deleteModel.argtypes = [ctypes.c_void_p]
        L.mj_copyModel.restype = ctypes.c_void_p
        L.mj_copyModel.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.refh_model_int.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_longlong)]
        L.refh_model_array.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p),
                                       ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                       ctypes.POINTER(ctypes.c_int)]
        L.refh_opt_num.argtypes = [ctypes.c_void_p, ctypes.c_char_p,
                                   ctypes.POINTER(ctypes.POINTER(ctypes.c_double)),
                                   ctypes.POINTER(ctypes.c_int)]
        L.refh_opt_int.argtypes = [ctypes.c_void_p, ctypes.c_char_p,
                                   ctypes.POINTER(ctypes.POINTER(ctypes.c_int))]
        L.refh_inverse_batch.restype = ctypes.c_double
        L.refh_inverse_batch.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.refh_inverse_timers.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.refh_compare_fwdinv.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _lib = L
    return _lib


# contact_* pseudo-fields: columns and dtype (ref_harness.c contact_field)
_CONTACT_FIELDS = {
    "contact_geom": (2, np.int32), "contact_dist": (1, np.float64), "contact_pos": (3, np.float64),
    "contact_frame": (9, np.float64), "contact_dim": (1, np.int32), "contact_exclude": (1, np.int32),
    "contact_efc_address": (1, np.int32), "contact_includemargin": (1, np.float64),
    "contact_friction": (5, np.float64), "contact_solref": (2, np.float64),
    "contact_solreffriction": (2, np.float64), "contact_solimp": (5, np.float64),
    "contact_mu": (1, np.float64),
}
_SCALARS = ("ncon", "ne", "nf", "nl", "nefc", "nJ")
_INT_EFC = ("efc_type", "efc_id", "efc_state", "efc_J_rownnz", "efc_J_rowadr")


class Model:
    """An mjModel owned by the reference library."""

    def __init__(self, ptr, own=True):
        if not ptr:
            raise RuntimeError("null mjModel")
        self.ptr = ctypes.c_void_p(ptr)
        self._own = own

    @classmethod
    def from_xml(cls, path):
        err = ctypes.create_string_buffer(1000)
        p = lib().mj_loadXML(os.fsencode(path), None, err, 1000)
        if not p:
            raise RuntimeError(f"mj_loadXML({path}): {err.value.decode()}")
        return cls(p)

    @classmethod
    def from_mjb(cls, path):
        p = lib().mj_loadModel(os.fsencode(path), None)
        if not p:
            raise RuntimeError(f"mj_loadModel({path}) failed")
        return cls(p)

    def copy(self):
        return Model(lib().mj_copyModel(None, self.ptr))

    def save_mjb(self, path):
        lib().mj_saveModel(self.ptr, os.fsencode(path), None, 0)

    def __del__(self):
        try:
            if self._own and self.ptr:
                lib().mj_deleteModel(self.ptr)
                self.ptr = None
        except Exception:
            pass

    def int(self, name):
        v = ctypes.c_longlong()
        if lib().refh_model_int(self.ptr, name.encode(), ctypes.byref(v)):
            raise KeyError(name)
        return int(v.value)

    def array(self, name):
        """numpy VIEW (nr x nc) of an mjModel array; writes go to the model."""
        ptr = ctypes.c_void_p()
        nr, nc, code = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        if lib().refh_model_array(self.ptr, name.encode(), ctypes.byref(ptr), ctypes.byref(nr),
                                  ctypes.byref(nc), ctypes.byref(code)):
            raise KeyError(name)
        dt = _CODES[code.value]
        n = nr.value * nc.value
        if n == 0 or not ptr.value:
            return np.zeros((nr.value, nc.value), dtype=dt)
        buf = (ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dt).reshape(nr.value, nc.value)

    def opt_int(self, name):
        p = ctypes.POINTER(ctypes.c_int)()
        if lib().refh_opt_int(self.ptr, name.encode(), ctypes.byref(p)):
            raise KeyError(name)
        return p

    def get_opt_int(self, name):
        return int(self.opt_int(name)[0])

    def set_opt_int(self, name, value):
        self.opt_int(name)[0] = int(value)

    def opt_num(self, name):
        p = ctypes.POINTER(ctypes.c_double)()
        n = ctypes.c_int()
        if lib().refh_opt_num(self.ptr, name.encode(), ctypes.byref(p), ctypes.byref(n)):
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    # ------------------------------------------------------------------ batch mj_inverse
    def inverse_fd_batch(self, qpos, qvel, qacc, eps=1e-6, mass=False):
        """The reference's mjd_inverseFD over a batch: (DfDq, DfDv, DfDa[, DmDq]), each
        [nbatch, nv, nv] ([nbatch, nv, nM]); row i = derivative w.r.t. coordinate i."""
        L = lib()
        L.refh_inverse_fd_batch.restype = None
        L.refh_inverse_fd_batch.argtypes = [ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_void_p] * 3 + \
            [ctypes.c_double] + [ctypes.c_void_p] * 4
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        qacc = np.ascontiguousarray(qacc, dtype=np.float64)
        n, nv, nM = qpos.shape[0], self.int("nv"), self.int("nM")
        dq, dv, da = (np.zeros((n, nv, nv)) for _ in range(3))
        dm = np.zeros((n, nv, nM)) if mass else None
        L.refh_inverse_fd_batch(self.ptr, n, qpos.ctypes.data, qvel.ctypes.data, qacc.ctypes.data, float(eps),
                                dq.ctypes.data, dv.ctypes.data, da.ctypes.data,
                                dm.ctypes.data if mass else None)
        return (dq, dv, da, dm) if mass else (dq, dv, da)

    def inverse_fd_sensor_batch(self, qpos, qvel, qacc, eps=1e-6):
        """mjd_inverseFD with sensor Jacobians: (DfDq, DfDv, DfDa, DsDq, DsDv, DsDa)."""
        L = lib()
        L.refh_inverse_fd_sensor_batch.restype = None
        L.refh_inverse_fd_sensor_batch.argtypes = [ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_void_p] * 3 + \
            [ctypes.c_double] + [ctypes.c_void_p] * 6
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        qacc = np.ascontiguousarray(qacc, dtype=np.float64)
        n, nv, ns = qpos.shape[0], self.int("nv"), self.int("nsensordata")
        df = [np.zeros((n, nv, nv)) for _ in range(3)]
        ds = [np.zeros((n, nv, ns)) for _ in range(3)]
        L.refh_inverse_fd_sensor_batch(self.ptr, n, qpos.ctypes.data, qvel.ctypes.data, qacc.ctypes.data, float(eps),
                                       *(a.ctypes.data for a in df), *(a.ctypes.data for a in ds))
        return tuple(df + ds)

    def fwdinv_batch(self, qpos, qvel, ctrl=None, qfrc_applied=None, xfrc_applied=None, dqacc=None):
        """mj_forward + mj_compareFwdInv per state: dict(qacc, qfrc_actuator, qfrc_constraint [n, nv],
        fwdinv [n, 2]); dqacc shifts qacc away from the forward solution before the comparison."""
        L = lib()
        L.refh_fwdinv_batch.restype = None
        L.refh_fwdinv_batch.argtypes = [ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_void_p] * 10
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        qpos, qvel, ctrl, qfrc_applied, xfrc_applied, dqacc = map(c, (qpos, qvel, ctrl, qfrc_applied, xfrc_applied, dqacc))
        n, nv = qpos.shape[0], self.int("nv")
        out = {"qacc": np.zeros((n, nv)), "qfrc_actuator": np.zeros((n, nv)),
               "qfrc_constraint": np.zeros((n, nv)), "fwdinv": np.zeros((n, 2))}
        p = lambda a: None if a is None else a.ctypes.data
        L.refh_fwdinv_batch(self.ptr, n, p(qpos), p(qvel), p(ctrl), p(qfrc_applied), p(xfrc_applied), p(dqacc),
                            p(out["qacc"]), p(out["qfrc_actuator"]), p(out["qfrc_constraint"]), p(out["fwdinv"]))
        return out

    def inverse_batch(self, qpos, qvel, qacc, fields=None, nthread=1, mocap=None, xfrc=None, eq_active=None):
        """Loop the reference's mj_inverse over the batch.

        fields: {name: maxrows} of extra mjData arrays / contact_* pseudo fields / scalar counters
        to collect per state (maxrows ignored for scalars and fixed-size arrays when None).
        Returns (dict of arrays, wall_seconds). Arrays are [nbatch, maxrows, nc] (squeezed if nc==1
        for scalars)."""
        L = lib()
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        qacc = np.ascontiguousarray(qacc, dtype=np.float64)
        n = qpos.shape[0]
        nv = self.int("nv")
        out = {"qfrc_inverse": np.zeros((n, nv))}
        reqs = []
        fields = fields or {}
        for name, maxrows in fields.items():
            if name in _SCALARS:
                arr = np.zeros((n,), dtype=np.int32)
                maxrows = 1
            elif name in _CONTACT_FIELDS:
                nc, dt = _CONTACT_FIELDS[name]
                arr = np.zeros((n, maxrows, nc), dtype=dt)
            else:
                nr, nc, dt = self._data_shape(name)
                if maxrows is None:
                    maxrows = nr
                arr = np.zeros((n, maxrows, nc), dtype=dt)
            out[name] = arr
            reqs.append((name.encode(), arr, maxrows))
        # cacc / cfrc_int / cfrc_ext are outputs of mj_rnePostConstraint, run after mj_inverse on request
        L.refh_set_post_constraint(int(any(k in fields for k in ("cacc", "cfrc_int", "cfrc_ext"))))
        L.refh_set_mocap.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        if mocap is not None:      # (mocap_pos [n, nmocap, 3], mocap_quat [n, nmocap, 4]) per state
            mp = np.ascontiguousarray(mocap[0], dtype=np.float64)
            mq = np.ascontiguousarray(mocap[1], dtype=np.float64)
            L.refh_set_mocap(mp.ctypes.data, mq.ctypes.data)
        else:
            L.refh_set_mocap(None, None)
        L.refh_set_xfrc.argtypes = [ctypes.c_void_p]
        xf = None
        if xfrc is not None:       # d->xfrc_applied per state [n, nbody, 6] (read by mj_rnePostConstraint)
            xf = np.ascontiguousarray(xfrc, dtype=np.float64)
            L.refh_set_xfrc(xf.ctypes.data)
        else:
            L.refh_set_xfrc(None)
        L.refh_set_eq_active.argtypes = [ctypes.c_void_p]
        ea = None
        if eq_active is not None:  # d->eq_active per state [n, neq] bytes
            ea = np.ascontiguousarray(eq_active, dtype=np.uint8)
            L.refh_set_eq_active(ea.ctypes.data)
        else:
            L.refh_set_eq_active(None)
        rq = (_Request * max(1, len(reqs)))()
        for i, (nm, arr, mr) in enumerate(reqs):
            rq[i].name = nm
            rq[i].out = arr.ctypes.data
            rq[i].maxrows = mr
        t = L.refh_inverse_batch(self.ptr, n, qpos.ctypes.data, qvel.ctypes.data, qacc.ctypes.data,
                                 out["qfrc_inverse"].ctypes.data, rq, len(reqs), int(nthread))
        L.refh_set_mocap(None, None)
        L.refh_set_xfrc(None)
        L.refh_set_eq_active(None)
        if t < 0:
            raise RuntimeError("refh_inverse_batch: unknown field requested")
        return out, t

    def _data_shape(self, name):
        # fixed-size mjData arrays: sizes follow the model; arena arrays need an explicit maxrows
        known = {
            "xpos": ("nbody", 3), "xquat": ("nbody", 4), "xmat": ("nbody", 9), "xipos": ("nbody", 3),
            "ximat": ("nbody", 9), "xanchor": ("njnt", 3), "xaxis": ("njnt", 3),
            "geom_xpos": ("ngeom", 3), "geom_xmat": ("ngeom", 9), "subtree_com": ("nbody", 3),
            "cinert": ("nbody", 10), "cdof": ("nv", 6), "cvel": ("nbody", 6), "cdof_dot": ("nv", 6),
            "crb": ("nbody", 10), "qM": ("nM", 1), "qLD": ("nC", 1), "qLDiagInv": ("nv", 1),
            "qfrc_bias": ("nv", 1), "qfrc_passive": ("nv", 1), "qfrc_constraint": ("nv", 1),
            "qfrc_spring": ("nv", 1), "qfrc_damper": ("nv", 1), "ten_length": ("ntendon", 1),
            "ten_velocity": ("ntendon", 1), "ten_J": ("ntendon", "nv"),
            "sensordata": ("nsensordata", 1), "cacc": ("nbody", 6), "cfrc_int": ("nbody", 6), "cfrc_ext": ("nbody", 6),
            "cam_xpos": ("ncam", 3), "cam_xmat": ("ncam", 9), "light_xpos": ("nlight", 3), "light_xdir": ("nlight", 3),
            "actuator_length": ("nu", 1), "actuator_velocity": ("nu", 1),
        }
        if name == "energy":        # member array of mjData: potential, kinetic
            return 2, 1, np.float64
        if name in known:
            r, c = known[name]
            return self.int(r), (self.int(c) if isinstance(c, str) else c), np.float64
        if name in _INT_EFC:
            return 0, 1, np.int32
        if name in ("moment_rownnz", "moment_rowadr"):      # compressed rows of actuator_moment
            return self.int("nu"), 1, np.int32
        if name == "moment_colind":
            return self.int("nJmom"), 1, np.int32
        if name == "actuator_moment":
            return self.int("nJmom"), 1, np.float64
        if name.startswith("efc_"):
            return 0, (4 if name == "efc_KBIP" else 1), np.float64
        raise KeyError(name)

    def inverse_timers(self, qpos, qvel, qacc):
        """Per-stage seconds of single-threaded mj_inverse (d->timer[], mjdata.h:67-92)."""
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        qacc = np.ascontiguousarray(qacc, dtype=np.float64)
        out = np.zeros(32)
        lib().refh_inverse_timers(self.ptr, qpos.shape[0], qpos.ctypes.data, qvel.ctypes.data,
                                  qacc.ctypes.data, out.ctypes.data)
        return out

    def settle(self, nstep):
        """mj_step x nstep from the default state, then mj_forward: (qpos, qvel, qacc, sensordata)."""
        L = lib()
        L.refh_settle.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4
        L.refh_settle.restype = None
        qpos, qvel, qacc = np.zeros(self.int("nq")), np.zeros(self.int("nv")), np.zeros(self.int("nv"))
        sd = np.zeros(max(1, self.int("nsensordata")))
        L.refh_settle(self.ptr, int(nstep), qpos.ctypes.data, qvel.ctypes.data, qacc.ctypes.data, sd.ctypes.data)
        return qpos, qvel, qacc, sd[:self.int("nsensordata")]

    def compare_fwdinv(self, qpos, qvel, ctrl=None, nstep=0):
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        res = np.zeros(2)
        c = None if ctrl is None else np.ascontiguousarray(ctrl, dtype=np.float64)
        lib().refh_compare_fwdinv(self.ptr, qpos.ctypes.data, qvel.ctypes.data,
                                  None if c is None else c.ctypes.data, int(nstep), res.ctypes.data)
        return res


REFERENCE_ROOT = "/root/reference"


def reference_path(rel):
    return os.path.join(REFERENCE_ROOT, rel)
