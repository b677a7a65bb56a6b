/* Test/bench harness around the UNMODIFIED reference engine (oracle/_ref/libmujoco_ref.so).
 *
 * TEST INFRASTRUCTURE ONLY: nothing in the product path (mujoco_inversedynamicstest_b200/)
 * may link or call this file. It is compiled against the reference's own headers
 * (/root/reference/include) and linked into libmujoco_ref.so by oracle/Makefile.
 *
 * It provides
 *   - name-based access to every mjModel / mjData array through the reference's X-macro
 *     lists (include/mujoco/mjxmacro.h:69,183,594,703-760), so the Python tests never
 *     restate a struct layout;
 *   - refh_inverse_batch(): the CPU loop that mjb_inverse() replaces, i.e.
 *         for i in batch: copy (qpos,qvel,qacc)_i -> d ; mj_inverse(m, d) ; copy outputs
 *     spread over host cores with the reference's own thread pool exactly as SURVEY.md
 *     §8(d) prescribes (mju_threadPoolCreate src/thread/thread_pool.cc:163, one mjData per
 *     worker, chunk = max(1, nbatch/(10*P)) following python/mujoco/rollout.cc:308-312).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <mujoco/mujoco.h>
#include <mujoco/mjxmacro.h>
#include "thread/thread_pool.h"

#define REFH_API __attribute__((visibility("default")))

enum { REFH_DOUBLE = 0, REFH_INT = 1, REFH_BYTE = 2, REFH_FLOAT = 3, REFH_OTHER = 4 };

#define refh_code_mjtNum REFH_DOUBLE
#define refh_code_int REFH_INT
#define refh_code_mjtByte REFH_BYTE
#define refh_code_float REFH_FLOAT
#define refh_code_char REFH_BYTE
#define refh_code_uintptr_t REFH_OTHER
#define refh_code_mjContact REFH_OTHER
#define refh_code_mjWarningStat REFH_OTHER
#define refh_code_mjTimerStat REFH_OTHER
#define refh_code_mjSolverStat REFH_OTHER
#define refh_code_size_t REFH_OTHER
#define refh_code_mjtSize REFH_OTHER

/* ------------------------------------------------------------------ model access */

REFH_API int refh_model_int(const mjModel* m, const char* name, long long* out) {
#define X(field) if (!strcmp(name, #field)) { *out = (long long)m->field; return 0; }
  MJMODEL_INTS
#undef X
  return -1;
}

REFH_API int refh_model_array(const mjModel* m, const char* name, const void** ptr,
                              int* nr, int* nc, int* code) {
  MJMODEL_POINTERS_PREAMBLE(m)
  (void)nuser_body; (void)nuser_jnt; (void)nuser_geom; (void)nuser_site; (void)nuser_cam;
  (void)nuser_tendon; (void)nuser_actuator; (void)nuser_sensor; (void)nq; (void)nv; (void)na;
  (void)nu; (void)nmocap;
#define X(type, field, r, c) \
  if (!strcmp(name, #field)) { *ptr = m->field; *nr = (int)m->r; *nc = (int)(c); *code = refh_code_##type; return 0; }
  MJMODEL_POINTERS
#undef X
  return -1;
}

REFH_API int refh_opt_num(mjModel* m, const char* name, double** ptr, int* n) {
#define X(type, field) if (!strcmp(name, #field)) { *ptr = &m->opt.field; *n = 1; return 0; }
  MJOPTION_FLOATS
#undef X
#define X(field, cnt) if (!strcmp(name, #field)) { *ptr = m->opt.field; *n = (cnt); return 0; }
  MJOPTION_VECTORS
#undef X
  return -1;
}

REFH_API int refh_opt_int(mjModel* m, const char* name, int** ptr) {
#define X(type, field) if (!strcmp(name, #field)) { *ptr = &m->opt.field; return 0; }
  MJOPTION_INTS
#undef X
  return -1;
}

/* ------------------------------------------------------------------ data access */

REFH_API int refh_data_int(const mjData* d, const char* name, long long* out) {
#define X(type, field) if (!strcmp(name, #field)) { *out = (long long)d->field; return 0; }
  MJDATA_SCALAR
#undef X
  return -1;
}

/* Fixed-size (buffer) arrays and arena arrays of mjData. */
REFH_API int refh_data_array(const mjModel* m, const mjData* d, const char* name,
                             const void** ptr, int* nr, int* nc, int* code) {
  MJDATA_POINTERS_PREAMBLE(m)
  (void)nv;
  /* d->energy is a member array of mjData (potential, kinetic), not one of the X-macro pointers */
  if (!strcmp(name, "energy")) { *ptr = d->energy; *nr = 2; *nc = 1; *code = refh_code_mjtNum; return 0; }
#define X(type, field, r, c) \
  if (!strcmp(name, #field)) { *ptr = d->field; *nr = (int)m->r; *nc = (int)(c); *code = refh_code_##type; return 0; }
  MJDATA_POINTERS
#undef X
#undef MJ_M
#undef MJ_D
#define MJ_M(n) m->n
#define MJ_D(n) d->n
#define X(type, field, r, c) \
  if (!strcmp(name, #field)) { *ptr = d->field; *nr = (int)(r); *nc = (int)(c); *code = refh_code_##type; return 0; }
  MJDATA_ARENA_POINTERS_SOLVER
#undef X
#undef MJ_M
#undef MJ_D
#define MJ_M(n) n
#define MJ_D(n) n
  return -1;
}

/* ------------------------------------------------------------------ batch outputs
 * A request names an mjData field (or a contact_* pseudo-field) and gives an output
 * buffer with room for `maxrows` rows per state; rows beyond the state's actual row count
 * are filled with 0 (numbers) / -1 (ints). */

typedef struct {
  const char* name;
  void* out;       /* [nbatch][maxrows*nc] of double or int */
  int maxrows;
} refhRequest;

enum {
  CF_NONE = 0, CF_GEOM, CF_DIST, CF_POS, CF_FRAME, CF_DIM, CF_EXCLUDE, CF_EFCADR,
  CF_INCLUDEMARGIN, CF_FRICTION, CF_SOLREF, CF_SOLREFFRICTION, CF_SOLIMP, CF_MU
};

static int contact_field(const char* name, int* nc, int* code) {
  static const struct { const char* n; int id, nc, code; } tab[] = {
    {"contact_geom", CF_GEOM, 2, REFH_INT}, {"contact_dist", CF_DIST, 1, REFH_DOUBLE},
    {"contact_pos", CF_POS, 3, REFH_DOUBLE}, {"contact_frame", CF_FRAME, 9, REFH_DOUBLE},
    {"contact_dim", CF_DIM, 1, REFH_INT}, {"contact_exclude", CF_EXCLUDE, 1, REFH_INT},
    {"contact_efc_address", CF_EFCADR, 1, REFH_INT},
    {"contact_includemargin", CF_INCLUDEMARGIN, 1, REFH_DOUBLE},
    {"contact_friction", CF_FRICTION, 5, REFH_DOUBLE}, {"contact_solref", CF_SOLREF, 2, REFH_DOUBLE},
    {"contact_solreffriction", CF_SOLREFFRICTION, 2, REFH_DOUBLE},
    {"contact_solimp", CF_SOLIMP, 5, REFH_DOUBLE}, {"contact_mu", CF_MU, 1, REFH_DOUBLE},
  };
  for (size_t i = 0; i < sizeof(tab)/sizeof(tab[0]); i++) {
    if (!strcmp(name, tab[i].n)) { *nc = tab[i].nc; *code = tab[i].code; return tab[i].id; }
  }
  return CF_NONE;
}

static void copy_contact_field(const mjData* d, int id, int nc, int code, void* out, int maxrows) {
  for (int r = 0; r < maxrows; r++) {
    double* od = (double*)out + (size_t)r*nc;
    int* oi = (int*)out + (size_t)r*nc;
    if (r >= d->ncon) {
      for (int c = 0; c < nc; c++) { if (code == REFH_INT) oi[c] = -1; else od[c] = 0; }
      continue;
    }
    const mjContact* con = d->contact + r;
    switch (id) {
      case CF_GEOM: oi[0] = con->geom[0]; oi[1] = con->geom[1]; break;
      case CF_DIST: od[0] = con->dist; break;
      case CF_POS: memcpy(od, con->pos, 3*sizeof(double)); break;
      case CF_FRAME: memcpy(od, con->frame, 9*sizeof(double)); break;
      case CF_DIM: oi[0] = con->dim; break;
      case CF_EXCLUDE: oi[0] = con->exclude; break;
      case CF_EFCADR: oi[0] = con->efc_address; break;
      case CF_INCLUDEMARGIN: od[0] = con->includemargin; break;
      case CF_FRICTION: memcpy(od, con->friction, 5*sizeof(double)); break;
      case CF_SOLREF: memcpy(od, con->solref, 2*sizeof(double)); break;
      case CF_SOLREFFRICTION: memcpy(od, con->solreffriction, 2*sizeof(double)); break;
      case CF_SOLIMP: memcpy(od, con->solimp, 5*sizeof(double)); break;
      case CF_MU: od[0] = con->mu; break;
    }
  }
}

/* Returns -1 for an unknown name. */
static int copy_request(const mjModel* m, const mjData* d, const refhRequest* rq, size_t i) {
  int nc, code;
  int cf = contact_field(rq->name, &nc, &code);
  if (cf != CF_NONE) {
    size_t esz = code == REFH_INT ? sizeof(int) : sizeof(double);
    copy_contact_field(d, cf, nc, code, (char*)rq->out + i*(size_t)rq->maxrows*nc*esz, rq->maxrows);
    return 0;
  }
  long long sc;
  if (!refh_data_int(d, rq->name, &sc)) {   /* scalar ints: ncon, nefc, ne, nf, nl, nJ ... */
    ((int*)rq->out)[i] = (int)sc;
    return 0;
  }
  const void* ptr; int nr;
  if (refh_data_array(m, d, rq->name, &ptr, &nr, &nc, &code)) return -1;
  if (code != REFH_DOUBLE && code != REFH_INT) return -1;
  size_t esz = code == REFH_INT ? sizeof(int) : sizeof(double);
  char* out = (char*)rq->out + i*(size_t)rq->maxrows*nc*esz;
  int rows = nr < rq->maxrows ? nr : rq->maxrows;
  if (rows > 0) memcpy(out, ptr, (size_t)rows*nc*esz);
  for (size_t k = (size_t)rows*nc; k < (size_t)rq->maxrows*nc; k++) {
    if (code == REFH_INT) ((int*)out)[k] = -1; else ((double*)out)[k] = 0;
  }
  return 0;
}

/* ------------------------------------------------------------------ the batch loop */

typedef struct {
  const mjModel* m;
  mjData** d;             /* one per worker (index = worker id, 0 = calling thread) */
  mjThreadPool* pool;
  const double *qpos, *qvel, *qacc;
  double* qfrc_inverse;   /* may be NULL */
  const refhRequest* req;
  int nreq;
  size_t begin, end;
  int err;
} refhChunk;

/* when set, mj_rnePostConstraint runs after mj_inverse (mj_sensorAcc calls it only for models with
 * acceleration / force / torque sensors, engine_sensor.c; cacc, cfrc_int, cfrc_ext are its outputs) */
static int refh_post_constraint = 0;
REFH_API void refh_set_post_constraint(int on) { refh_post_constraint = on; }

/* per-state mocap poses for the batch loop (nbatch x nmocap x 3 | 4), NULL: model pose */
static const double* refh_mocap_pos = NULL;
static const double* refh_mocap_quat = NULL;
REFH_API void refh_set_mocap(const double* pos, const double* quat) { refh_mocap_pos = pos; refh_mocap_quat = quat; }
/* per-state d->xfrc_applied for the batch loop (nbatch x nbody x 6), NULL: zero */
static const double* refh_xfrc = NULL;
REFH_API void refh_set_xfrc(const double* xfrc) { refh_xfrc = xfrc; }
/* per-state d->eq_active for the batch loop (nbatch x neq bytes), NULL: eq_active0 */
static const unsigned char* refh_eq_active = NULL;
REFH_API void refh_set_eq_active(const unsigned char* a) { refh_eq_active = a; }

static void* run_chunk(void* arg) {
  refhChunk* c = (refhChunk*)arg;
  const mjModel* m = c->m;
  size_t wid = c->pool ? mju_threadPoolCurrentWorkerId(c->pool) : 0;
  mjData* d = c->d[wid];
  for (size_t i = c->begin; i < c->end; i++) {
    mju_copy(d->qpos, c->qpos + i*m->nq, m->nq);
    mju_copy(d->qvel, c->qvel + i*m->nv, m->nv);
    mju_copy(d->qacc, c->qacc + i*m->nv, m->nv);
    if (refh_mocap_pos && refh_mocap_quat) {
      mju_copy(d->mocap_pos, refh_mocap_pos + i*3*m->nmocap, 3*m->nmocap);
      mju_copy(d->mocap_quat, refh_mocap_quat + i*4*m->nmocap, 4*m->nmocap);
    }
    if (refh_xfrc) mju_copy(d->xfrc_applied, refh_xfrc + (size_t)i*6*m->nbody, 6*m->nbody);
    if (refh_eq_active) memcpy(d->eq_active, refh_eq_active + (size_t)i*m->neq, (size_t)m->neq);
    mj_inverse(m, d);
    if (refh_post_constraint) mj_rnePostConstraint(m, d);
    if (c->qfrc_inverse) mju_copy(c->qfrc_inverse + i*m->nv, d->qfrc_inverse, m->nv);
    for (int r = 0; r < c->nreq; r++) {
      if (copy_request(m, d, c->req + r, i)) c->err = 1;
    }
  }
  return NULL;
}

/* Loop mj_inverse over nbatch states with nthread workers; returns wall seconds, <0 on error. */
REFH_API double refh_inverse_batch(const mjModel* m, long long nbatch, const double* qpos,
                                   const double* qvel, const double* qacc, double* qfrc_inverse,
                                   const refhRequest* req, int nreq, int nthread) {
  if (nthread < 1) nthread = 1;
  if (nthread > mjMAXTHREAD - 1) nthread = mjMAXTHREAD - 1;
  mjThreadPool* pool = nthread > 1 ? mju_threadPoolCreate((size_t)nthread) : NULL;
  int nd = nthread + 1;
  mjData** d = (mjData**)calloc((size_t)nd, sizeof(mjData*));
  for (int i = 0; i < nd; i++) d[i] = mj_makeData(m);

  size_t chunk = (size_t)(nbatch/(10*(long long)nthread));
  if (chunk < 1) chunk = 1;
  size_t nchunk = ((size_t)nbatch + chunk - 1)/chunk;
  refhChunk* chunks = (refhChunk*)calloc(nchunk ? nchunk : 1, sizeof(refhChunk));
  mjTask* tasks = (mjTask*)calloc(nchunk ? nchunk : 1, sizeof(mjTask));

  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (size_t k = 0; k < nchunk; k++) {
    refhChunk* c = chunks + k;
    c->m = m; c->d = d; c->pool = pool;
    c->qpos = qpos; c->qvel = qvel; c->qacc = qacc; c->qfrc_inverse = qfrc_inverse;
    c->req = req; c->nreq = nreq;
    c->begin = k*chunk;
    c->end = (k + 1)*chunk < (size_t)nbatch ? (k + 1)*chunk : (size_t)nbatch;
    if (pool) {
      mju_defaultTask(tasks + k);
      tasks[k].func = run_chunk;
      tasks[k].args = c;
      mju_threadPoolEnqueue(pool, tasks + k);   /* spins when the 640-slot queue is full */
    } else {
      run_chunk(c);
    }
  }
  if (pool) for (size_t k = 0; k < nchunk; k++) mju_taskJoin(tasks + k);
  clock_gettime(CLOCK_MONOTONIC, &t1);

  int err = 0;
  for (size_t k = 0; k < nchunk; k++) err |= chunks[k].err;
  free(tasks); free(chunks);
  for (int i = 0; i < nd; i++) mj_deleteData(d[i]);
  free(d);
  if (pool) mju_threadPoolDestroy(pool);
  if (err) return -1.0;
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9*(double)(t1.tv_nsec - t0.tv_nsec);
}

/* Per-stage split of one thread's mj_inverse using the reference's own TM_START/TM_END timers
 * (src/engine/engine_macro.h:35-40, include/mujoco/mjdata.h:67-92). out[mjNTIMER] in seconds. */
static mjtNum refh_clock(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (mjtNum)t.tv_sec*1e6 + (mjtNum)t.tv_nsec*1e-3;   /* microseconds */
}

REFH_API void refh_inverse_timers(const mjModel* m, long long nbatch, const double* qpos,
                                  const double* qvel, const double* qacc, double* out) {
  mjfTime old = mjcb_time;
  mjcb_time = refh_clock;
  mjData* d = mj_makeData(m);
  for (long long i = 0; i < nbatch; i++) {
    mju_copy(d->qpos, qpos + i*m->nq, m->nq);
    mju_copy(d->qvel, qvel + i*m->nv, m->nv);
    mju_copy(d->qacc, qacc + i*m->nv, m->nv);
    mj_inverse(m, d);
  }
  for (int t = 0; t < mjNTIMER; t++) out[t] = 1e-6*d->timer[t].duration;
  mj_deleteData(d);
  mjcb_time = old;
}

/* forward dynamics + compareFwdInv on one state (restates test/engine/engine_inverse_test.cc:35-56):
 * returns solver_fwdinv[0..1]. */
REFH_API void refh_compare_fwdinv(const mjModel* m, const double* qpos, const double* qvel,
                                  const double* ctrl, int nstep, double fwdinv[2]) {
  mjData* d = mj_makeData(m);
  mju_copy(d->qpos, qpos, m->nq);
  mju_copy(d->qvel, qvel, m->nv);
  if (ctrl) mju_copy(d->ctrl, ctrl, m->nu);
  for (int i = 0; i < nstep; i++) mj_step(m, d);
  mj_forward(m, d);
  mj_compareFwdInv(m, d);
  fwdinv[0] = d->solver_fwdinv[0];
  fwdinv[1] = d->solver_fwdinv[1];
  mj_deleteData(d);
}

/* Settle a model the way the reference's rne_post tests do (test/engine/engine_core_smooth_test.cc:160-240:
 * mj_step x nstep from the default state), then mj_forward: the state (qpos, qvel, qacc) and the
 * reference's sensordata at it. Used to restate those known-answer tests on the inverse path. */
REFH_API void refh_settle(const mjModel* m, int nstep, double* qpos, double* qvel, double* qacc,
                          double* sensordata) {
  mjData* d = mj_makeData(m);
  for (int i = 0; i < nstep; i++) mj_step(m, d);
  mj_forward(m, d);
  mju_copy(qpos, d->qpos, m->nq);
  mju_copy(qvel, d->qvel, m->nv);
  mju_copy(qacc, d->qacc, m->nv);
  if (sensordata) mju_copy(sensordata, d->sensordata, m->nsensordata);
  mj_deleteData(d);
}

/* mj_forward + mj_compareFwdInv over a batch (engine_inverse.c:275-316): for every state the
 * forward dynamics run on (qpos, qvel, ctrl, qfrc_applied, xfrc_applied); d->qacc is then shifted by
 * dqacc (NULL: left as solved, so that the comparison is also exercised away from the solution),
 * and the quantities mjb_compareFwdInv takes as inputs plus the reference's solver_fwdinv are
 * returned. Arrays are row-major per state; ctrl / qfrc_applied / xfrc_applied / dqacc may be NULL. */
REFH_API void refh_fwdinv_batch(const mjModel* m, long long nbatch, const double* qpos,
                                const double* qvel, const double* ctrl, const double* qfrc_applied,
                                const double* xfrc_applied, const double* dqacc, double* qacc,
                                double* qfrc_actuator, double* qfrc_constraint, double* fwdinv) {
  mjData* d = mj_makeData(m);
  const size_t nv = (size_t)m->nv;
  for (long long i = 0; i < nbatch; i++) {
    mju_copy(d->qpos, qpos + i*m->nq, m->nq);
    mju_copy(d->qvel, qvel + i*m->nv, m->nv);
    if (ctrl) mju_copy(d->ctrl, ctrl + i*m->nu, m->nu);
    if (qfrc_applied) mju_copy(d->qfrc_applied, qfrc_applied + i*nv, m->nv);
    if (xfrc_applied) mju_copy(d->xfrc_applied, xfrc_applied + i*6*m->nbody, 6*m->nbody);
    mju_zero(d->qacc_warmstart, m->nv);
    mj_forward(m, d);
    if (dqacc) mju_addTo(d->qacc, dqacc + i*nv, m->nv);
    mju_copy(qacc + i*nv, d->qacc, m->nv);
    mju_copy(qfrc_actuator + i*nv, d->qfrc_actuator, m->nv);
    mju_copy(qfrc_constraint + i*nv, d->qfrc_constraint, m->nv);
    mj_compareFwdInv(m, d);
    fwdinv[2*i] = d->solver_fwdinv[0];
    fwdinv[2*i + 1] = d->solver_fwdinv[1];
  }
  mj_deleteData(d);
}

/* the reference's mjd_inverseFD (src/engine/engine_derivative_fd.c:611) looped over a batch, no
 * actuation, no sensors: DfDq / DfDv / DfDa are nbatch x nv x nv (row i = derivative with respect
 * to coordinate i), DmDq nbatch x nv x nM (may be NULL). */
REFH_API void refh_inverse_fd_batch(const mjModel* m, long long nbatch, const double* qpos,
                                    const double* qvel, const double* qacc, double eps,
                                    double* DfDq, double* DfDv, double* DfDa, double* DmDq) {
  mjData* d = mj_makeData(m);
  const size_t nv = (size_t)m->nv, nM = (size_t)m->nM;
  for (long long i = 0; i < nbatch; i++) {
    mju_copy(d->qpos, qpos + i*m->nq, m->nq);
    mju_copy(d->qvel, qvel + i*m->nv, m->nv);
    mju_copy(d->qacc, qacc + i*m->nv, m->nv);
    mjd_inverseFD(m, d, eps, 0, DfDq + i*nv*nv, DfDv + i*nv*nv, DfDa + i*nv*nv, NULL, NULL, NULL,
                  DmDq ? DmDq + i*nv*nM : NULL);
  }
  mj_deleteData(d);
}

/* the same with the sensor Jacobians DsDq / DsDv / DsDa, nbatch x nv x nsensordata each */
REFH_API void refh_inverse_fd_sensor_batch(const mjModel* m, long long nbatch, const double* qpos,
                                           const double* qvel, const double* qacc, double eps,
                                           double* DfDq, double* DfDv, double* DfDa,
                                           double* DsDq, double* DsDv, double* DsDa) {
  mjData* d = mj_makeData(m);
  const size_t nv = (size_t)m->nv, ns = (size_t)m->nsensordata;
  for (long long i = 0; i < nbatch; i++) {
    mju_copy(d->qpos, qpos + i*m->nq, m->nq);
    mju_copy(d->qvel, qvel + i*m->nv, m->nv);
    mju_copy(d->qacc, qacc + i*m->nv, m->nv);
    mjd_inverseFD(m, d, eps, 0, DfDq + i*nv*nv, DfDv + i*nv*nv, DfDa + i*nv*nv,
                  DsDq + i*nv*ns, DsDv + i*nv*ns, DsDa + i*nv*ns, NULL);
  }
  mj_deleteData(d);
}

REFH_API int refh_sizeof_request(void) { return (int)sizeof(refhRequest); }
