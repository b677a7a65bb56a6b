/* mjb.h -- batched inverse dynamics on NVIDIA B200 behind the MuJoCo 3.3.1 C API.
 *
 * Drop-in boundary for the reference fork's inverse-dynamics entry points
 *   void mj_inverse(const mjModel* m, mjData* d);              include/mujoco/mujoco.h:131
 *   void mj_inverseSkip(const mjModel*, mjData*, int, int);    include/mujoco/mujoco.h:137
 * evaluated over `nbatch` independent (qpos, qvel, qacc) states:
 *
 *   for (i = 0; i < nbatch; i++) {            // what the caller does today, cf.
 *     mju_copy(d->qpos, qpos + i*nq, nq);     //   src/inverse/inverse_test.cpp:43-112
 *     mju_copy(d->qvel, qvel + i*nv, nv);     //   python/mujoco/rollout.cc:73-178
 *     mju_copy(d->qacc, qacc + i*nv, nv);
 *     mj_inverse(m, d);
 *     mju_copy(out + i*nv, d->qfrc_inverse, nv);
 *   }
 *
 * becomes mjb_setState(...); mjb_inverse(m, bd, nbatch); mjb_getQfrcInverse(bd, out).
 *
 * The model stays the caller's mjModel (same struct, same headers); one mjbData is the batched
 * counterpart of mjData and owns the device copies. Like mjData it is not re-entrant: one
 * mjbData per calling thread. All entry points are plain C (pointers and sizes only).
 *
 * Host arrays use the reference's row-major "array of states" layout (nbatch x nq etc., as in
 * python/mujoco/rollout.cc:41-66). Device arrays are structure-of-arrays: element (row r, state i)
 * of a field lives at ptr[r*stride + i], stride = mjb_stride(d).
 */
#ifndef MJB_H_
#define MJB_H_

#include <mujoco/mujoco.h>

#if defined(__cplusplus)
extern "C" {
#endif

#define MJB_API __attribute__((visibility("default")))

typedef struct mjbData_ mjbData;

/* which optional per-state outputs mjb_inverse fills (qfrc_inverse is always produced) */
typedef enum mjbOut_ {
  mjbOUT_QFRC      = 1 << 0,  /* qfrc_constraint, qfrc_passive, qfrc_bias           mjdata.h qfrc_*   */
  mjbOUT_COUNTS    = 1 << 1,  /* ncon, ne, nf, nl, nefc                                               */
  mjbOUT_CONTACT   = 1 << 2,  /* contact[].geom, dim, exclude, efc_address, dist, pos, frame          */
  mjbOUT_EFC       = 1 << 3,  /* efc_type, efc_id, efc_state, efc_pos, margin, D, R, vel, aref, force */
  mjbOUT_INERTIA   = 1 << 4,  /* qM, qLD, qLDiagInv (mj_crb, mj_factorM)                              */
  mjbOUT_INTERNAL  = 1 << 5,  /* every position/velocity-stage intermediate (xpos ... cfrc), debug    */
  mjbOUT_RNEPOST   = 1 << 6,  /* cacc, cfrc_int, cfrc_ext as left by mj_rnePostConstraint             *
                               * (include/mujoco/mujoco.h mj_rnePostConstraint,                       *
                               * src/engine/engine_core_smooth.c:2027-2181; mjdata.h cacc/cfrc_*)     */
  mjbOUT_CAMLIGHT  = 1 << 7,  /* cam_xpos, cam_xmat, light_xpos, light_xdir as mj_camlight leaves    *
                               * them inside mj_invPosition (engine_core_smooth.c:275-389)            */
  mjbOUT_TRANSMISSION = 1 << 8 /* actuator_length, actuator_moment (dense nu x nv), actuator_velocity:*
                               * mj_transmission inside mj_invPosition (engine_core_smooth.c:865-1346)*
                               * and mj_fwdVelocity (engine_forward.c:216). Adhesion actuators        *
                               * (mjTRN_BODY) read the contact list: COUNTS and CONTACT come with it. */
} mjbOut;

/* per-state status bits, the batched form of d->warning[] (engine_forward.c:53-102,
 * engine_core_constraint.c:68,237,252) */
typedef enum mjbStatus_ {
  mjbSTATUS_BADQPOS     = 1 << 0,
  mjbSTATUS_BADQVEL     = 1 << 1,
  mjbSTATUS_BADQACC     = 1 << 2,
  mjbSTATUS_CONTACTFULL = 1 << 3,  /* more contacts than nconmax: contact outputs truncated */
  mjbSTATUS_CNSTRFULL   = 1 << 4   /* more rows than njmax: efc outputs truncated           */
} mjbStatus;

/* output fields addressable through mjb_get / mjb_devicePtr */
typedef enum mjbField_ {
  mjbF_QFRC_INVERSE = 0,  /* double nv                       */
  mjbF_QFRC_CONSTRAINT,   /* double nv            (QFRC)     */
  mjbF_QFRC_PASSIVE,      /* double nv            (QFRC)     */
  mjbF_COUNTS,            /* int    5: ncon ne nf nl nefc (COUNTS) */
  mjbF_STATUS,            /* int    1                        */
  mjbF_CONTACT_GEOM,      /* int    nconmax*2     (CONTACT)  */
  mjbF_CONTACT_INFO,      /* int    nconmax*3: dim exclude efc_address (CONTACT) */
  mjbF_CONTACT_NUM,       /* double nconmax*13: dist pos[3] frame[9]   (CONTACT) */
  mjbF_EFC_INT,           /* int    njmax*3: type id state             (EFC)     */
  mjbF_EFC_NUM,           /* double njmax*8: pos margin D R vel aref force diagApprox (EFC) */
  mjbF_QM,                /* double nM            (INERTIA)  */
  mjbF_QLD,               /* double nC            (INERTIA)  */
  mjbF_QLDIAGINV,         /* double nv            (INERTIA)  */
  mjbF_INTERNAL,          /* double mjb_internalSize() (INTERNAL) */
  mjbF_CACC,              /* double nbody*6: [angular, linear] acceleration, com frame (RNEPOST) */
  mjbF_CFRC_INT,          /* double nbody*6: [torque, force] body <- parent, com frame (RNEPOST) */
  mjbF_CFRC_EXT,          /* double nbody*6: [torque, force] of contacts and connect/weld rows (RNEPOST) */
  mjbF_SENSORDATA,        /* double nsensordata: d->sensordata (models with sensors, no mask bit needed) */
  mjbF_QFRC_BIAS,         /* double nv: mj_rne without accelerations, engine_forward.c:228 (QFRC) */
  mjbF_ENERGY,            /* double 2: d->energy = potential, kinetic (mj_energyPos / mj_energyVel, engine_sensor.c:920,
                             1011); produced, without a mask bit, for models with mjENBL_ENERGY -- engine_inverse.c:210-223 */
  mjbF_CAM_XPOS,          /* double ncam*3        (CAMLIGHT) */
  mjbF_CAM_XMAT,          /* double ncam*9        (CAMLIGHT) */
  mjbF_LIGHT_XPOS,        /* double nlight*3      (CAMLIGHT) */
  mjbF_LIGHT_XDIR,        /* double nlight*3      (CAMLIGHT) */
  mjbF_ACTUATOR_LENGTH,   /* double nu            (TRANSMISSION) */
  mjbF_ACTUATOR_MOMENT,   /* double nu*nv, row-major dense; the reference keeps the same rows compressed
                             (moment_rownnz / moment_rowadr / moment_colind, mjdata.h)   (TRANSMISSION) */
  mjbF_ACTUATOR_VELOCITY, /* double nu            (TRANSMISSION) */
  mjbF_COUNT
} mjbField;

/* Validate `m`, flatten its constant tables, upload them to CUDA device `device` and allocate
 * batch buffers for up to nbatch_max states. Returns NULL and writes a message into err (if not
 * NULL) when the model uses a feature outside the supported path (mesh/hfield/SDF geom
 * pairs that survive the static collision filters, flex, plugins, sensor types that
 * are not evaluated on the device (user / plugin) without mjDSBL_SENSOR, INVDISCRETE
 * with RK4; full list in DESIGN.md section 5) or when CUDA fails.
 * Models with sensors get d->sensordata (mjbF_SENSORDATA) from mj_sensorPos / Vel / Acc
 * (src/engine/engine_sensor.c:222,527,708) on every mjb_inverse, like mj_inverse.
 * nconmax / njmax bound the per-state contact / constraint-row OUTPUT arrays (they do not limit
 * the physics); pass 0 for defaults. */
MJB_API mjbData* mjb_makeData(const mjModel* m, int nbatch_max, int device, unsigned outmask,
                              int nconmax, int njmax, char* err, int err_sz);
MJB_API void mjb_deleteData(mjbData* d);

/* run subsequent work on this CUDA stream (a cudaStream_t passed as void*); default stream 0 */
MJB_API void mjb_setStream(mjbData* d, void* cuda_stream);

/* copy nbatch states from HOST arrays (nbatch x nq, nbatch x nv, nbatch x nv) to the device */
MJB_API int mjb_setState(mjbData* d, int nbatch, const mjtNum* qpos, const mjtNum* qvel,
                         const mjtNum* qacc);
/* per-state poses of the mocap bodies, d->mocap_pos / d->mocap_quat (include/mujoco/mjdata.h; read by
 * mj_kinematics, src/engine/engine_core_smooth.c:70-86): HOST arrays nbatch x nmocap x 3 and
 * nbatch x nmocap x 4, used by every following mjb_inverse until replaced. NULL pointers return to
 * the model pose (what mj_makeData / mj_resetData leave in mjData). */
MJB_API int mjb_setMocap(mjbData* d, int nbatch, const mjtNum* mocap_pos, const mjtNum* mocap_quat);
/* per-state d->xfrc_applied (include/mujoco/mjdata.h): HOST array nbatch x nbody x 6, (force, torque) per
 * body in world axes at the body's centre of mass, used by every following mjb_inverse until replaced;
 * NULL returns to zero. mj_inverse ignores applied wrenches in qfrc_inverse; they enter the
 * mj_rnePostConstraint outputs cfrc_ext / cfrc_int (src/engine/engine_core_smooth.c:2039-2049) and the
 * force / torque sensors that read cfrc_int. Needs mjbOUT_RNEPOST (or such sensors) to have any effect. */
MJB_API int mjb_setXfrcApplied(mjbData* d, int nbatch, const mjtNum* xfrc_applied);
/* per-state d->eq_active (include/mujoco/mjdata.h; read by mj_instantiateEquality,
 * src/engine/engine_core_constraint.c:493-763): HOST array nbatch x neq of mjtByte, used by every following
 * mjb_inverse until replaced; NULL returns to the model's eq_active0 (what mj_makeData / mj_resetData
 * leave in mjData). Row numbers of everything after the equality block follow per state. */
MJB_API int mjb_setEqActive(mjbData* d, int nbatch, const unsigned char* eq_active);
/* adopt DEVICE structure-of-arrays inputs without a copy: (nq|nv) x stride, stride >= nbatch.
 * Pass NULL pointers to return to the internal buffers. */
MJB_API int mjb_setStateDevice(mjbData* d, const mjtNum* qpos, const mjtNum* qvel,
                               const mjtNum* qacc, long long stride);

/* mj_inverse over states [0, nbatch). Returns the number of states with a non-zero status word,
 * or a negative value on a CUDA error (message via mjb_lastError). Synchronous with respect to
 * the host only through the getters; the launch itself is asynchronous on the stream. */
MJB_API int mjb_inverse(const mjModel* m, mjbData* d, int nbatch);
/* mj_inverseSkip over the batch (include/mujoco/mujoco.h:137; used by src/inverse/inverse_test.cpp:93
 * with mjSTAGE_VEL, skipsensor = 1). skipstage only permits the CPU engine to reuse earlier stages;
 * the batched engine recomputes everything, which yields the same results. skipsensor != 0 leaves
 * sensordata untouched. Returns like mjb_inverse. */
MJB_API int mjb_inverseSkip(const mjModel* m, mjbData* d, int nbatch, int skipstage, int skipsensor);
/* mjd_inverseFD over the batch (include/mujoco/mujoco.h mjd_inverseFD, src/engine/
 * engine_derivative_fd.c:611; flg_actuation = 0, sensors not evaluated): forward-difference
 * Jacobians of qfrc_inverse with respect to qpos (through mj_integratePos), qvel and qacc of the
 * states last given to mjb_setState / mjb_setStateDevice. Outputs are HOST arrays, any of them may
 * be NULL: DfDq, DfDv, DfDa nbatch x nv x nv, DmDq nbatch x nv x nM, row i = derivative with
 * respect to coordinate i (the reference's transposed layout). 1 + 3 nv evaluations per state,
 * generated and differenced on the device. Synchronous. */
MJB_API int mjb_inverseFD(const mjModel* m, mjbData* d, int nbatch, mjtNum eps, mjtNum* DfDq,
                          mjtNum* DfDv, mjtNum* DfDa, mjtNum* DmDq);
/* The same with the sensor Jacobians of mjd_inverseFD (engine_derivative_fd.c:611-730): DsDq, DsDv,
 * DsDa, each nbatch x nv x nsensordata (row i = derivative of sensordata with respect to coordinate
 * i), any of them may be NULL; the force / mass Jacobians as in mjb_inverseFD. flg_actuation must be
 * 0: the inverse path evaluates no actuation (the reference subtracts qfrc_actuator when it is set).
 * With all three sensor outputs NULL the perturbed evaluations skip the sensors, as the reference
 * does (skipsensor, :629). */
MJB_API int mjb_inverseFDSensor(const mjModel* m, mjbData* d, int nbatch, mjtNum eps, int flg_actuation,
                                mjtNum* DfDq, mjtNum* DfDv, mjtNum* DfDa,
                                mjtNum* DsDq, mjtNum* DsDv, mjtNum* DsDa, mjtNum* DmDq);
/* mj_compareFwdInv over the batch (include/mujoco/mujoco.h mj_compareFwdInv, src/engine/
 * engine_inverse.c:275-316): how well the inverse dynamics reproduce what a forward pass applied.
 * The states last given to mjb_setState / mjb_setStateDevice carry the forward pass's qacc. HOST
 * arrays, one row per state: qfrc_applied, qfrc_actuator nbatch x nv (either may be NULL = zero),
 * xfrc_applied nbatch x nbody x 6 (force then torque per body, mjData layout; may be NULL; projected
 * with J' on the device like mj_xfrcAccumulate, engine_support.c:1247-1260), qfrc_constraint
 * nbatch x nv of the forward pass. Output fwdinv nbatch x 2 = d->solver_fwdinv:
 *   [0] |qfrc_constraint(forward) - qfrc_constraint(inverse)|,
 *   [1] |qfrc_applied + qfrc_actuator + J'xfrc_applied - qfrc_inverse|  (both 0 for states without
 * constraint rows, as in the reference). Synchronous; returns 0 or a negative value on error. */
MJB_API int mjb_compareFwdInv(const mjModel* m, mjbData* d, int nbatch, const mjtNum* qfrc_applied,
                              const mjtNum* qfrc_actuator, const mjtNum* xfrc_applied,
                              const mjtNum* qfrc_constraint, mjtNum* fwdinv);
/* same, without reading back the status count (fully asynchronous) */
MJB_API int mjb_inverseAsync(const mjModel* m, mjbData* d, int nbatch);

/* host-to-host form of the loop: qpos/qvel/qacc in (nbatch x nq|nv, HOST, pinned for best speed),
 * qfrc_inverse out (nbatch x nv, HOST). The batch is pipelined in pieces over three CUDA streams
 * (H2D copy | kernels | D2H copy) so transfers overlap compute. The input arrays are read from the
 * moment of the call (they must hold the states when it is made and stay unchanged until the results
 * are there); consecutive calls overlap: the copy-in of one call runs under the kernels of the
 * previous one. Asynchronous: the results are valid after mjb_synchronize(d) (or once the stream
 * given to mjb_setStream has drained). Returns 0, or a negative value on a CUDA error. */
MJB_API int mjb_inverseHost(const mjModel* m, mjbData* d, int nbatch, const mjtNum* qpos,
                            const mjtNum* qvel, const mjtNum* qacc, mjtNum* qfrc_inverse);

/* copy a field of the last mjb_inverse to a HOST array laid out nbatch x rows (row-major), where
 * nbatch is the batch size of the last evaluation: mjb_lastBatch(d) */
MJB_API int mjb_get(mjbData* d, int field, void* host_out);
MJB_API int mjb_lastBatch(const mjbData* d);
MJB_API int mjb_getQfrcInverse(mjbData* d, mjtNum* qfrc_inverse);
/* DEVICE structure-of-arrays view of a field (rows x stride), valid until mjb_deleteData */
MJB_API const void* mjb_devicePtr(mjbData* d, int field);
MJB_API int mjb_fieldRows(const mjbData* d, int field);
MJB_API long long mjb_stride(const mjbData* d);

/* layout of mjbF_INTERNAL: offset (in doubles) and length of a named intermediate, e.g. "xpos",
 * "cdof", "cvel"; returns -1 for an unknown name */
MJB_API int mjb_internalSlot(const mjbData* d, const char* name, int* offset, int* size);
MJB_API int mjb_internalSize(const mjbData* d);

/* candidate geom pairs that survived the static filters, in contact order (for tests/tools) */
MJB_API int mjb_ncandidate(const mjbData* d);
MJB_API void mjb_candidate(const mjbData* d, int i, int* geom1, int* geom2, int* func);

MJB_API const char* mjb_lastError(const mjbData* d);

/* number of mj_inverse phase kernels this mjbData has launched so far (diagnostics / benchmark) */
MJB_API long long mjb_kernelLaunches(const mjbData* d);

/* Per-kernel timing of the phase kernels (diagnostics / benchmark): while enabled, every launch is
 * bracketed by CUDA events on the launching stream. mjb_phaseTimes synchronises and returns the
 * milliseconds accumulated since the last call in ms[0..n): smooth, inertia, contact_scan, contact,
 * backward, discrete_acc, tree (the fused forward + inertia stages of a specialised model). */
MJB_API void mjb_phaseTiming(mjbData* d, int enable);
MJB_API int mjb_phaseTimes(mjbData* d, double* ms, int n);

/* diagnostics: counters of the item-parallel contact phase of the last chunk
 * (items, contacts, overflow flags -- bit0 item list, bit1 contact list --, slots); -1 if that path is
 * not in use */
MJB_API int mjb_debugQueue(mjbData* d, int* out4);

/* Several devices behind one mjbData: shard g lives on devices[g] and evaluates the
 * contiguous range [g*per, (g+1)*per) of every batch, per = ceil(nbatch / ndevice); the model is
 * replicated on each device, there is no collective. mjb_setState / mjb_inverse / mjb_inverseAsync /
 * mjb_inverseSkip / mjb_inverseHost / mjb_get / mjb_setMocap / mjb_synchronize / mjb_specialize work on
 * the whole batch with the same host arrays as on one device; mjb_inverseHost drives every device from
 * its own host thread (the reference's batched API chunks over a thread pool the same way,
 * python/mujoco/rollout.cc:180-210). Device views (mjb_devicePtr, mjb_setStateDevice), mjb_inverseFD
 * and mjb_compareFwdInv need a single-device mjbData. A device may be listed more than once. */
MJB_API mjbData* mjb_makeDataMulti(const mjModel* m, int nbatch_max, const int* devices, int ndevice,
                                   unsigned outmask, int nconmax, int njmax, char* err, int err_sz);
MJB_API int mjb_ndevice(const mjbData* d);

/* Model specialisation. mjb_specialize compiles the phase kernels FOR THIS MODEL (NVRTC, sm_100a; the
 * model's tables become compile-time constants and the loops over bodies / dofs / candidate pairs are
 * expanded) and switches this mjbData to them; the compiled module is cached on disk
 * (<directory of libmjb.so>/jitcache or $MJB_JIT_CACHE). Results are those of the generic kernels up to
 * floating-point contraction. Returns 0, or -1 with a message when the model is too large to expand or
 * NVRTC / the driver API cannot be loaded -- the generic kernels then stay in use. The environment
 * variable MJB_JIT=1 makes mjb_makeData call it for every model.
 * mjb_precompile fills the cache without a GPU (build step); info receives "<key> compiled|cached <s>". */
MJB_API int mjb_specialize(mjbData* d, char* err, int err_sz);
MJB_API int mjb_specialized(const mjbData* d);
MJB_API int mjb_specializeInfo(const mjbData* d, char* key, int key_sz, int* from_cache, double* compile_seconds);
MJB_API int mjb_precompile(const mjModel* m, char* info, int info_sz);

/* wait for the stream */
MJB_API int mjb_synchronize(mjbData* d);

/* measured FP64 FMA peak of the device in TFLOP/s (roofline denominator; not on the data path) */
MJB_API double mjb_fp64PeakTflops(int device);

#if defined(__cplusplus)
}
#endif

#endif  /* MJB_H_ */
