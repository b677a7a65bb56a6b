"""ctypes driver of tests/hostemu/libhostemu.so (TEST-ONLY host build of the device pipeline)."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_hostemu  # noqa: E402

_FIELDS = ["qfrc_inverse", "qfrc_constraint", "qfrc_passive", "counts", "status", "contact_geom",
           "contact_info", "contact_num", "efc_int", "efc_num", "qM", "qLD", "qLDiagInv",
           "scratch_dump", "cacc", "cfrc_int", "cfrc_ext", "sensordata", "qfrc_bias", "fwd_qforce", "fwd_xfrc", "fwd_qfrc_constraint", "fwdinv", "mocap_pos", "mocap_quat", "energy",
           "cam_xpos", "cam_xmat", "light_xpos", "light_xdir",
           "actuator_length", "actuator_moment", "actuator_velocity", "xfrc_applied", "eq_active"]


class Outputs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in _FIELDS]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build_hostemu.build())
        L.hostemu_inverse.argtypes = ([ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3 +
                                      [ctypes.c_int] * 2 + [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int])
        L.hostemu_nscratch.argtypes = [ctypes.c_void_p]
        L.hostemu_slot.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int),
                                   ctypes.POINTER(ctypes.c_int)]
        L.hostemu_candidates.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_char_p, ctypes.c_int]
        _lib = L
    return _lib


def available():
    return os.path.exists(build_hostemu.LIB) or build_hostemu.include_dir() is not None


def run(model, qpos, qvel, qacc, nconmax=64, njmax=256, post=False, fwd=None, mocap=None, camlight=False,
        transmission=False, xfrc=None, dump=True, eq_active=None):
    """model: object with .ptr (mjModel*) and .int(). Returns dict of arrays shaped like the
    Python host mirror's getters ([n, rows])."""
    L = lib()
    n, nv = qpos.shape[0], model.int("nv")
    nsc = L.hostemu_nscratch(model.ptr)
    if nsc < 0:
        err = ctypes.create_string_buffer(1000)
        L.hostemu_candidates(model.ptr, None, 0, err, 1000)
        raise RuntimeError(err.value.decode())
    a = dict(qfrc_inverse=np.zeros((nv, n)), qfrc_constraint=np.zeros((nv, n)),
             qfrc_passive=np.zeros((nv, n)), counts=np.zeros((5, n), np.int32),
             status=np.zeros((n,), np.int32),
             contact_geom=np.zeros((nconmax * 2, n), np.int32),
             contact_info=np.zeros((nconmax * 3, n), np.int32),
             contact_num=np.zeros((nconmax * 13, n)), efc_int=np.zeros((njmax * 3, n), np.int32),
             efc_num=np.zeros((njmax * 8, n)), qM=np.zeros((model.int("nM"), n)),
             qLD=np.zeros((model.int("nC"), n)), qLDiagInv=np.zeros((nv, n)),
             scratch_dump=np.zeros((nsc, n)), qfrc_bias=np.zeros((nv, n)))
    has_sensors = model.int("nsensordata") > 0 and not (model.get_opt_int("disableflags") & (1 << 12))
    if post or has_sensors:      # mj_rnePostConstraint outputs (mjb_makeData adds them for acceleration-stage sensors)
        nb = model.int("nbody")
        a.update(cacc=np.zeros((6 * nb, n)), cfrc_int=np.zeros((6 * nb, n)), cfrc_ext=np.zeros((6 * nb, n)))
    if fwd is not None:     # mj_compareFwdInv inputs: dict(qforce, qfrc_constraint[, xfrc]) per state
        a["fwd_qforce"] = np.ascontiguousarray(fwd["qforce"].T, dtype=np.float64)
        a["fwd_qfrc_constraint"] = np.ascontiguousarray(fwd["qfrc_constraint"].T, dtype=np.float64)
        if fwd.get("xfrc") is not None:
            a["fwd_xfrc"] = np.ascontiguousarray(fwd["xfrc"].reshape(n, -1).T, dtype=np.float64)
        a["fwdinv"] = np.zeros((2, n))
    if eq_active is not None and model.int("neq") > 0:   # d->eq_active per state [n, neq]
        a["eq_active"] = np.ascontiguousarray((np.asarray(eq_active) != 0).astype(np.float64).T)
    if xfrc is not None:    # d->xfrc_applied per state [n, nbody, 6]
        a["xfrc_applied"] = np.ascontiguousarray(xfrc.reshape(n, -1).T, dtype=np.float64)
    if mocap is not None:   # (mocap_pos [n, nmocap, 3], mocap_quat [n, nmocap, 4])
        a["mocap_pos"] = np.ascontiguousarray(mocap[0].reshape(n, -1).T, dtype=np.float64)
        a["mocap_quat"] = np.ascontiguousarray(mocap[1].reshape(n, -1).T, dtype=np.float64)
    if has_sensors:
        a["sensordata"] = np.zeros((model.int("nsensordata"), n))
    stypes = set(model.array("sensor_type").ravel().tolist()) if has_sensors else set()
    if (model.get_opt_int("enableflags") & (1 << 1)) or (stypes & {40, 41}):     # mjENBL_ENERGY / energy sensors
        a["energy"] = np.zeros((2, n))
    # what mjb_makeData adds for sensors that read mj_camlight / mj_transmission outputs
    camlight = camlight or 8 in stypes
    transmission = transmission or bool(stypes & {13, 14})
    # implicitfast mjENBL_INVDISCRETE with velocity-biased actuators reads the moment rows
    transmission = transmission or bool((model.get_opt_int("enableflags") & (1 << 3)) and
                                        model.get_opt_int("integrator") in (2, 3) and model.int("nu") > 0)
    if camlight:            # mj_camlight outputs (mjbOUT_CAMLIGHT)
        nc, nl = max(1, model.int("ncam")), max(1, model.int("nlight"))
        a.update(cam_xpos=np.zeros((3 * nc, n)), cam_xmat=np.zeros((9 * nc, n)),
                 light_xpos=np.zeros((3 * nl, n)), light_xdir=np.zeros((3 * nl, n)))
    if transmission:        # mj_transmission outputs (mjbOUT_TRANSMISSION)
        nu = max(1, model.int("nu"))
        a.update(actuator_length=np.zeros((nu, n)), actuator_moment=np.zeros((nu * max(1, nv), n)),
                 actuator_velocity=np.zeros((nu, n)))
    if not dump:            # like the product: rows of the scratch are stored only where a later stage reads them
        del a["scratch_dump"]
    o = Outputs(**{k: v.ctypes.data for k, v in a.items()})   # absent members stay NULL
    err = ctypes.create_string_buffer(1000)
    qp, qv, qa = (np.ascontiguousarray(x.T, dtype=np.float64) for x in (qpos, qvel, qacc))
    if L.hostemu_inverse(model.ptr, n, qp.ctypes.data, qv.ctypes.data, qa.ctypes.data, nconmax,
                         njmax, ctypes.byref(o), err, 1000):
        raise RuntimeError(err.value.decode())
    out = {k: v.T.copy() if v.ndim == 2 else v for k, v in a.items()}
    c = out["counts"]
    out.update(ncon=c[:, 0], ne=c[:, 1], nf=c[:, 2], nl=c[:, 3], nefc=c[:, 4])
    out["contact_geom"] = out["contact_geom"].reshape(n, nconmax, 2)
    info = out["contact_info"].reshape(n, nconmax, 3)
    num = out["contact_num"].reshape(n, nconmax, 13)
    out.update(contact_dim=info[:, :, 0], contact_exclude=info[:, :, 1],
               contact_efc_address=info[:, :, 2], contact_dist=num[:, :, 0],
               contact_pos=num[:, :, 1:4], contact_frame=num[:, :, 4:13])
    ei = out["efc_int"].reshape(n, njmax, 3)
    en = out["efc_num"].reshape(n, njmax, 8)
    out.update(efc_type=ei[:, :, 0], efc_id=ei[:, :, 1], efc_state=ei[:, :, 2], efc_pos=en[:, :, 0],
               efc_D=en[:, :, 2], efc_aref=en[:, :, 5], efc_force=en[:, :, 6])
    return out


def slot(model, out, name):
    off, sz = ctypes.c_int(), ctypes.c_int()
    if lib().hostemu_slot(model.ptr, name.encode(), ctypes.byref(off), ctypes.byref(sz)):
        raise KeyError(name)
    return out["scratch_dump"][:, off.value:off.value + sz.value]


def candidates(model, max_pairs=200000):
    buf = np.zeros((max_pairs, 3), np.int32)
    err = ctypes.create_string_buffer(1000)
    n = lib().hostemu_candidates(model.ptr, buf.ctypes.data, max_pairs, err, 1000)
    if n < 0:
        raise RuntimeError(err.value.decode())
    return buf[:n]
