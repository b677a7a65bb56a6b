"""Python host mirror of the batched inverse-dynamics C-ABI (include/mjb.h).

    model = Model.from_mjb("humanoid.mjb")            # or Model(ptr) around the caller's mjModel*
    bd = BatchData(model, nbatch_max, outmask=OUT_COUNTS)
    bd.set_state(qpos, qvel, qacc)                     # host arrays nbatch x nq / nv / nv
    bd.inverse()                                       # == looping mj_inverse over the batch
    qfrc = bd.qfrc_inverse()                           # nbatch x nv

Method names follow the reference workflow they replace: python/mujoco/rollout.cc (batched host
arrays nbatch x nstate) and src/inverse/inverse_test.cpp (mj_inverse per state).
"""
import ctypes

import numpy as np

from ._lib import lib

(OUT_QFRC, OUT_COUNTS, OUT_CONTACT, OUT_EFC, OUT_INERTIA, OUT_INTERNAL, OUT_RNEPOST, OUT_CAMLIGHT,
 OUT_TRANSMISSION) = (1 << i for i in range(9))

STATUS_BADQPOS, STATUS_BADQVEL, STATUS_BADQACC, STATUS_CONTACTFULL, STATUS_CNSTRFULL = (
    1 << i for i in range(5))

(F_QFRC_INVERSE, F_QFRC_CONSTRAINT, F_QFRC_PASSIVE, F_COUNTS, F_STATUS, F_CONTACT_GEOM,
 F_CONTACT_INFO, F_CONTACT_NUM, F_EFC_INT, F_EFC_NUM, F_QM, F_QLD, F_QLDIAGINV, F_INTERNAL,
 F_CACC, F_CFRC_INT, F_CFRC_EXT, F_SENSORDATA, F_QFRC_BIAS, F_ENERGY, F_CAM_XPOS, F_CAM_XMAT, F_LIGHT_XPOS,
 F_LIGHT_XDIR, F_ACTUATOR_LENGTH, F_ACTUATOR_MOMENT, F_ACTUATOR_VELOCITY) = range(27)

_INT_FIELDS = {F_COUNTS, F_STATUS, F_CONTACT_GEOM, F_CONTACT_INFO, F_EFC_INT}
_CODES = {0: np.float64, 1: np.int32, 2: np.uint8, 3: np.float32}


class MjbError(RuntimeError):
    pass


class Model:
    """A `const mjModel*`: either the caller's (from their libmujoco) or one read from an MJB file."""

    def __init__(self, ptr, owned=False, keepalive=None):
        if not ptr:
            raise MjbError("null mjModel pointer")
        self.ptr = ctypes.c_void_p(ptr if isinstance(ptr, int) else ptr.value)
        self._owned = owned
        self._keepalive = keepalive

    @classmethod
    def from_mjb(cls, path):
        err = ctypes.create_string_buffer(512)
        if str(path).endswith(".gz"):
            import gzip
            with gzip.open(path, "rb") as f:
                data = f.read()
            p = lib().mjb_loadModelBuffer(data, len(data), err, 512)
        else:
            p = lib().mjb_loadModel(str(path).encode(), err, 512)
        if not p:
            raise MjbError(f"mjb_loadModel({path}): {err.value.decode()}")
        return cls(p, owned=True)

    def __del__(self):
        try:
            if self._owned and self.ptr:
                lib().mjb_freeModel(self.ptr)
                self.ptr = None
        except Exception:
            pass

    def int(self, name):
        v = ctypes.c_longlong()
        if lib().mjb_modelInt(self.ptr, name.encode(), ctypes.byref(v)):
            raise KeyError(name)
        return int(v.value)

    def array(self, name):
        ptr = ctypes.c_void_p()
        nr, nc, code = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        if lib().mjb_modelArray(self.ptr, name.encode(), ctypes.byref(ptr), ctypes.byref(nr),
                                ctypes.byref(nc), ctypes.byref(code)):
            raise KeyError(name)
        dt = _CODES[code.value]
        n = nr.value * nc.value
        if n == 0 or not ptr.value:
            return np.zeros((nr.value, nc.value), dtype=dt)
        buf = (ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr.value)
        buf._owner = self        # the view keeps the model (and with it the memory) alive
        return np.frombuffer(buf, dtype=dt).reshape(nr.value, nc.value)

    def get_opt_int(self, name):
        p = lib().mjb_modelOptInt(self.ptr, name.encode())
        if not p:
            raise KeyError(name)
        return int(p[0])

    def set_opt_int(self, name, value):
        p = lib().mjb_modelOptInt(self.ptr, name.encode())
        if not p:
            raise KeyError(name)
        p[0] = int(value)


class BatchData:
    """mjbData: the batched counterpart of mjData on one CUDA device."""

    def __init__(self, model, nbatch_max, device=0, outmask=0, nconmax=0, njmax=0, stream=None, devices=None):
        """devices: list of CUDA device indices -> one mjbData sharded over them (mjb_makeDataMulti)."""
        self.model = model
        self.nq, self.nv = model.int("nq"), model.int("nv")
        err = ctypes.create_string_buffer(1024)
        if devices is not None:
            arr = (ctypes.c_int * len(devices))(*[int(x) for x in devices])
            self._d = lib().mjb_makeDataMulti(model.ptr, int(nbatch_max), arr, len(devices), int(outmask),
                                              int(nconmax), int(njmax), err, 1024)
            device = devices[0] if devices else 0
        else:
            self._d = lib().mjb_makeData(model.ptr, int(nbatch_max), int(device), int(outmask),
                                         int(nconmax), int(njmax), err, 1024)
        if not self._d:
            raise MjbError(err.value.decode())
        self._d = ctypes.c_void_p(self._d)
        self.nbatch_max = int(nbatch_max)
        self.nbatch = 0
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "_d", None):
            lib().mjb_deleteData(self._d)
            self._d = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise MjbError(f"{what}: {lib().mjb_lastError(self._d).decode()}")
        return rc

    def set_stream(self, cuda_stream):
        """cuda_stream: integer handle (e.g. torch.cuda.current_stream().cuda_stream)."""
        lib().mjb_setStream(self._d, ctypes.c_void_p(int(cuda_stream)))

    @property
    def stride(self):
        return int(lib().mjb_stride(self._d))

    def set_state(self, qpos, qvel, qacc):
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        qacc = np.ascontiguousarray(qacc, dtype=np.float64)
        n = qpos.shape[0]
        if qpos.shape != (n, self.nq) or qvel.shape != (n, self.nv) or qacc.shape != (n, self.nv):
            raise ValueError("expected qpos [n, nq], qvel [n, nv], qacc [n, nv]")
        self._check(lib().mjb_setState(self._d, n, qpos.ctypes.data, qvel.ctypes.data,
                                       qacc.ctypes.data), "mjb_setState")
        self.nbatch = n

    def set_state_ptr(self, n, qpos_ptr, qvel_ptr, qacc_ptr):
        """Host pointers (e.g. pinned torch tensors) laid out n x nq / n x nv / n x nv."""
        self._check(lib().mjb_setState(self._d, int(n), ctypes.c_void_p(qpos_ptr),
                                       ctypes.c_void_p(qvel_ptr), ctypes.c_void_p(qacc_ptr)),
                    "mjb_setState")
        self.nbatch = int(n)

    def set_state_device(self, n, qpos_ptr, qvel_ptr, qacc_ptr, stride):
        """Device structure-of-arrays pointers ((nq|nv) x stride); no copy."""
        self._check(lib().mjb_setStateDevice(self._d, ctypes.c_void_p(qpos_ptr),
                                             ctypes.c_void_p(qvel_ptr), ctypes.c_void_p(qacc_ptr),
                                             int(stride)), "mjb_setStateDevice")
        self.nbatch = int(n)

    def set_xfrc_applied(self, xfrc_applied):
        """Per-state d->xfrc_applied [nbatch, nbody, 6] (force, torque per body); None returns to zero.
        Enters cfrc_ext / cfrc_int of mj_rnePostConstraint (OUT_RNEPOST), not qfrc_inverse."""
        if xfrc_applied is None:
            self._check(lib().mjb_setXfrcApplied(self._d, 0, None), "mjb_setXfrcApplied")
            return
        x = np.ascontiguousarray(xfrc_applied, dtype=np.float64)
        self._check(lib().mjb_setXfrcApplied(self._d, x.shape[0], x.ctypes.data), "mjb_setXfrcApplied")

    def set_eq_active(self, eq_active):
        """Per-state d->eq_active [nbatch, neq] (0 / 1); None returns to the model's eq_active0."""
        if eq_active is None:
            self._check(lib().mjb_setEqActive(self._d, 0, None), "mjb_setEqActive")
            return
        e = np.ascontiguousarray(eq_active, dtype=np.uint8)
        self._check(lib().mjb_setEqActive(self._d, e.shape[0], e.ctypes.data), "mjb_setEqActive")

    def set_mocap(self, mocap_pos=None, mocap_quat=None):
        """Per-state mocap poses [nbatch, nmocap, 3] / [nbatch, nmocap, 4]; None: the model pose."""
        if mocap_pos is None or mocap_quat is None:
            self._check(lib().mjb_setMocap(self._d, 0, None, None), "mjb_setMocap")
            return
        p = np.ascontiguousarray(mocap_pos, dtype=np.float64)
        q = np.ascontiguousarray(mocap_quat, dtype=np.float64)
        self._check(lib().mjb_setMocap(self._d, p.shape[0], p.ctypes.data, q.ctypes.data), "mjb_setMocap")

    def inverse(self, nbatch=None, sync=True):
        n = self.nbatch if nbatch is None else int(nbatch)
        fn = lib().mjb_inverse if sync else lib().mjb_inverseAsync
        rc = self._check(fn(self.model.ptr, self._d, n), "mjb_inverse")
        self.nbatch = n
        return rc

    def inverse_fd(self, eps=1e-6, mass=False, nbatch=None):
        """mjd_inverseFD over the batch (mjb_inverseFD): (DfDq, DfDv, DfDa[, DmDq]) as
        [n, nv, nv] ([n, nv, nM]) arrays, row i = derivative w.r.t. coordinate i."""
        n = self.nbatch if nbatch is None else int(nbatch)
        nv, nM = self.model.int("nv"), self.model.int("nM")
        dq, dv, da = (np.zeros((n, nv, nv)) for _ in range(3))
        dm = np.zeros((n, nv, nM)) if mass else None
        self._check(lib().mjb_inverseFD(self.model.ptr, self._d, n, float(eps), dq.ctypes.data,
                                        dv.ctypes.data, da.ctypes.data,
                                        dm.ctypes.data if mass else None), "mjb_inverseFD")
        return (dq, dv, da, dm) if mass else (dq, dv, da)

    def inverse_fd_sensor(self, eps=1e-6, nbatch=None):
        """mjd_inverseFD with its sensor Jacobians (mjb_inverseFDSensor): (DfDq, DfDv, DfDa, DsDq, DsDv,
        DsDa), the last three [n, nv, nsensordata]."""
        n = self.nbatch if nbatch is None else int(nbatch)
        nv, ns = self.model.int("nv"), self.model.int("nsensordata")
        df = [np.zeros((n, nv, nv)) for _ in range(3)]
        ds = [np.zeros((n, nv, ns)) for _ in range(3)]
        self._check(lib().mjb_inverseFDSensor(self.model.ptr, self._d, n, float(eps), 0,
                                              *(a.ctypes.data for a in df), *(a.ctypes.data for a in ds), None),
                    "mjb_inverseFDSensor")
        return tuple(df + ds)

    def compare_fwdinv(self, qfrc_constraint, qfrc_applied=None, qfrc_actuator=None, xfrc_applied=None,
                       nbatch=None):
        """mj_compareFwdInv over the batch: [nbatch, 2] = solver_fwdinv of every state
        (engine_inverse.c:275-316). The states set last carry the forward pass's qacc."""
        n = self.nbatch if nbatch is None else int(nbatch)
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64)
                for a in (qfrc_applied, qfrc_actuator, xfrc_applied, qfrc_constraint)]
        out = np.zeros((n, 2))
        ptr = [None if a is None else ctypes.c_void_p(a.ctypes.data) for a in arrs]
        self._check(lib().mjb_compareFwdInv(self.model.ptr, self._d, n, *ptr, ctypes.c_void_p(out.ctypes.data)),
                    "mjb_compareFwdInv")
        return out

    def inverse_skip(self, skipstage=0, skipsensor=1, nbatch=None):
        """mj_inverseSkip over the batch (mjb_inverseSkip)."""
        n = self.nbatch if nbatch is None else int(nbatch)
        rc = self._check(lib().mjb_inverseSkip(self.model.ptr, self._d, n, int(skipstage), int(skipsensor)),
                         "mjb_inverseSkip")
        self.nbatch = n
        return rc

    def specialize(self):
        """mjb_specialize: compile (NVRTC) and switch to the kernels specialised for this model.
        Returns a dict with the cache key, whether the cubin came from the cache and the compile time;
        raises MjbError when the model cannot be specialised (the generic kernels stay in use)."""
        err = ctypes.create_string_buffer(8192)
        if lib().mjb_specialize(self._d, err, 8192):
            raise MjbError(err.value.decode())
        key = ctypes.create_string_buffer(64)
        cached, secs = ctypes.c_int(), ctypes.c_double()
        lib().mjb_specializeInfo(self._d, key, 64, ctypes.byref(cached), ctypes.byref(secs))
        return {"key": key.value.decode(), "from_cache": bool(cached.value), "compile_seconds": secs.value}

    @property
    def specialized(self):
        return bool(lib().mjb_specialized(self._d))

    def kernel_launches(self):
        return int(lib().mjb_kernelLaunches(self._d))

    PHASES = ("smooth", "inertia", "contact_scan", "contact", "backward", "discrete_acc", "tree")

    def phase_timing(self, enable=True):
        """Bracket every phase-kernel launch with CUDA events (mjb_phaseTiming)."""
        lib().mjb_phaseTiming(self._d, 1 if enable else 0)

    def phase_times(self):
        """{kernel: milliseconds} accumulated since the last call (mjb_phaseTimes; synchronises)."""
        ms = (ctypes.c_double * len(self.PHASES))()
        if lib().mjb_phaseTimes(self._d, ms, len(self.PHASES)):
            raise MjbError(f"mjb_phaseTimes: {lib().mjb_lastError(self._d).decode()}")
        return dict(zip(self.PHASES, [float(x) for x in ms]))

    def inverse_host(self, n, qpos_ptr, qvel_ptr, qacc_ptr, out_ptr):
        """Pipelined host-to-host pass (mjb_inverseHost): pointers to HOST arrays n x nq|nv|nv and
        the n x nv output; asynchronous, call synchronize() before reading the output."""
        self._check(lib().mjb_inverseHost(self.model.ptr, self._d, int(n), ctypes.c_void_p(qpos_ptr),
                                          ctypes.c_void_p(qvel_ptr), ctypes.c_void_p(qacc_ptr),
                                          ctypes.c_void_p(out_ptr)), "mjb_inverseHost")
        self.nbatch = int(n)

    def inverse_host_arrays(self, qpos, qvel, qacc):
        """numpy convenience around inverse_host (synchronous)."""
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        qacc = np.ascontiguousarray(qacc, dtype=np.float64)
        out = np.empty((qpos.shape[0], self.nv))
        self.inverse_host(qpos.shape[0], qpos.ctypes.data, qvel.ctypes.data, qacc.ctypes.data,
                          out.ctypes.data)
        self.synchronize()
        return out

    def synchronize(self):
        self._check(lib().mjb_synchronize(self._d), "mjb_synchronize")

    def rows(self, field):
        return int(lib().mjb_fieldRows(self._d, field))

    def device_ptr(self, field):
        return lib().mjb_devicePtr(self._d, field)

    def last_batch(self):
        """Batch size of the last evaluation (what mjb_get copies), independent of later set_state calls."""
        return int(lib().mjb_lastBatch(self._d))

    def get(self, field, out=None):
        rows = self.rows(field)
        dt = np.int32 if field in _INT_FIELDS else np.float64
        n = self.last_batch()
        if out is None:
            out = np.empty((n, rows), dtype=dt)
        elif (out.dtype != dt or out.size < n * rows or not out.flags["C_CONTIGUOUS"]):
            raise MjbError(f"mjb_get: `out` must be a C-contiguous {np.dtype(dt).name} array of at least "
                           f"{n} x {rows} elements")
        self._check(lib().mjb_get(self._d, field, out.ctypes.data), "mjb_get")
        return out

    def get_ptr(self, field, host_ptr):
        self._check(lib().mjb_get(self._d, field, ctypes.c_void_p(host_ptr)), "mjb_get")

    def qfrc_inverse(self):
        return self.get(F_QFRC_INVERSE)

    def status(self):
        return self.get(F_STATUS).ravel()

    def counts(self):
        """dict of ncon, ne, nf, nl, nefc arrays [nbatch]."""
        c = self.get(F_COUNTS)
        return {k: c[:, i] for i, k in enumerate(("ncon", "ne", "nf", "nl", "nefc"))}

    def contacts(self):
        nc = self.rows(F_CONTACT_GEOM) // 2
        geom = self.get(F_CONTACT_GEOM).reshape(-1, nc, 2)
        info = self.get(F_CONTACT_INFO).reshape(-1, nc, 3)
        num = self.get(F_CONTACT_NUM).reshape(-1, nc, 13)
        return {"geom": geom, "dim": info[:, :, 0], "exclude": info[:, :, 1],
                "efc_address": info[:, :, 2], "dist": num[:, :, 0], "pos": num[:, :, 1:4],
                "frame": num[:, :, 4:13]}

    def efc(self):
        nj = self.rows(F_EFC_INT) // 3
        ei = self.get(F_EFC_INT).reshape(-1, nj, 3)
        en = self.get(F_EFC_NUM).reshape(-1, nj, 8)
        names = ("pos", "margin", "D", "R", "vel", "aref", "force", "diagApprox")
        out = {"type": ei[:, :, 0], "id": ei[:, :, 1], "state": ei[:, :, 2]}
        out.update({k: en[:, :, i] for i, k in enumerate(names)})
        return out

    def rne_post_constraint(self):
        """cacc, cfrc_int, cfrc_ext [nbatch, nbody, 6] as mj_rnePostConstraint leaves them
        (src/engine/engine_core_smooth.c:2027-2181); needs OUT_RNEPOST."""
        return {k: self.get(f).reshape(self.last_batch(), -1, 6)
                for k, f in (("cacc", F_CACC), ("cfrc_int", F_CFRC_INT), ("cfrc_ext", F_CFRC_EXT))}

    def sensordata(self):
        """d->sensordata [nbatch, nsensordata] (mj_sensorPos / Vel / Acc; models with sensors)."""
        return self.get(F_SENSORDATA)

    def energy(self):
        """d->energy (potential, kinetic) per state, for models with mjENBL_ENERGY."""
        return self.get(F_ENERGY)

    def camlight(self):
        """cam_xpos [nbatch, ncam, 3], cam_xmat [.., 9], light_xpos / light_xdir [nbatch, nlight, 3] as
        mj_camlight leaves them (src/engine/engine_core_smooth.c:275-389); needs OUT_CAMLIGHT."""
        n = self.last_batch()
        ncam, nlight = self.model.int("ncam"), self.model.int("nlight")
        return {k: self.get(f).reshape(n, cnt, c) for k, f, cnt, c in (
            ("cam_xpos", F_CAM_XPOS, ncam, 3), ("cam_xmat", F_CAM_XMAT, ncam, 9),
            ("light_xpos", F_LIGHT_XPOS, nlight, 3), ("light_xdir", F_LIGHT_XDIR, nlight, 3))}

    def transmission(self):
        """actuator_length [nbatch, nu], actuator_moment [nbatch, nu, nv] (dense), actuator_velocity
        [nbatch, nu] (mj_transmission, src/engine/engine_core_smooth.c:865-1346); needs OUT_TRANSMISSION."""
        n = self.last_batch()
        nu = self.model.int("nu")
        return {"actuator_length": self.get(F_ACTUATOR_LENGTH),
                "actuator_moment": self.get(F_ACTUATOR_MOMENT).reshape(n, nu, self.nv),
                "actuator_velocity": self.get(F_ACTUATOR_VELOCITY)}

    def internal(self, name):
        off, size = ctypes.c_int(), ctypes.c_int()
        if lib().mjb_internalSlot(self._d, name.encode(), ctypes.byref(off), ctypes.byref(size)):
            raise KeyError(name)
        full = self.get(F_INTERNAL)
        return full[:, off.value:off.value + size.value]

    def candidates(self):
        n = lib().mjb_ncandidate(self._d)
        out = np.zeros((n, 3), dtype=np.int32)
        a, b, f = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        for i in range(n):
            lib().mjb_candidate(self._d, i, ctypes.byref(a), ctypes.byref(b), ctypes.byref(f))
            out[i] = (a.value, b.value, f.value)
        return out


def fp64_peak_tflops(device=0):
    return float(lib().mjb_fp64PeakTflops(int(device)))
