/* Drop-in test of the C boundary, compiled as plain C against the reference's own headers
 * (/root/reference/include/mujoco/mujoco.h) and include/mjb.h, linked with the reference library
 * (oracle/_ref/libmujoco_ref.so) and libmjb.so in ONE process:
 *
 *     m = mj_loadModel(...)                       a model struct built by the reference's loader
 *     for i: copy state i -> mj_inverse(m, d)     the loop of src/inverse/inverse_test.cpp:43-112
 *     mjb_makeData(m) -> mjb_setState -> mjb_inverse(m, bd, n) -> mjb_getQfrcInverse
 *
 * and compares qfrc_inverse element-wise at 1e-9 relative + 1e-12 absolute, then changes m->opt
 * between two calls (mj_inverse honours m->opt on every call) and compares again, and finally checks
 * that a process with mjcb_passive set is refused. Exit code 0 = all good.
 *
 * usage: dropin <model.mjb> [nstate] [entries allowed outside the element-wise bound, default 0]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mujoco/mujoco.h>

#include "mjb.h"

static unsigned long long rng_state = 20250331ull;
static double uniform(double lo, double hi) {             /* splitmix64 */
  unsigned long long z = (rng_state += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  z ^= z >> 31;
  return lo + (hi - lo) * ((double)(z >> 11) / 9007199254740992.0);
}

static void make_states(const mjModel* m, int n, mjtNum* qpos, mjtNum* qvel, mjtNum* qacc) {
  for (int s = 0; s < n; s++) {
    mjtNum* q = qpos + (size_t)s*m->nq;
    for (int j = 0; j < m->njnt; j++) {
      int a = m->jnt_qposadr[j];
      if (m->jnt_type[j] == mjJNT_FREE) {
        q[a] = uniform(-1, 1); q[a+1] = uniform(-1, 1); q[a+2] = uniform(0.0, 1.5);
        a += 3;
      }
      if (m->jnt_type[j] == mjJNT_FREE || m->jnt_type[j] == mjJNT_BALL) {
        double nrm = 0;
        for (int k = 0; k < 4; k++) { q[a+k] = uniform(-1, 1); nrm += q[a+k]*q[a+k]; }
        nrm = sqrt(nrm);
        for (int k = 0; k < 4; k++) q[a+k] /= nrm;
      } else if (m->jnt_limited[j]) {
        double lo = m->jnt_range[2*j], hi = m->jnt_range[2*j+1], w = hi - lo;
        q[a] = uniform(lo - 0.1*w, hi + 0.1*w);
      } else {
        q[a] = m->qpos0[a] + uniform(-1, 1);
      }
    }
    for (int i = 0; i < m->nv; i++) {
      qvel[(size_t)s*m->nv + i] = uniform(-1, 1);
      qacc[(size_t)s*m->nv + i] = uniform(-10, 10);
    }
  }
}

static void cpu_loop(const mjModel* m, mjData* d, int n, const mjtNum* qpos, const mjtNum* qvel,
                     const mjtNum* qacc, mjtNum* out) {
  for (int s = 0; s < n; s++) {
    mju_copy(d->qpos, qpos + (size_t)s*m->nq, m->nq);
    mju_copy(d->qvel, qvel + (size_t)s*m->nv, m->nv);
    mju_copy(d->qacc, qacc + (size_t)s*m->nv, m->nv);
    mj_inverse(m, d);
    mju_copy(out + (size_t)s*m->nv, d->qfrc_inverse, m->nv);
  }
}

static size_t allowed_bad = 0;

static int compare(const char* what, const mjtNum* got, const mjtNum* ref, size_t count) {
  size_t bad = 0;
  double worst = 0;
  for (size_t i = 0; i < count; i++) {
    double tol = 1e-12 + 1e-9*fabs(ref[i]);
    double r = fabs(got[i] - ref[i]) / tol;
    if (r > worst) worst = r;
    if (r > 1) bad++;
  }
  printf("%s: %zu entries, %zu outside 1e-9*|ref| + 1e-12, worst ratio %.3g\n", what, count, bad, worst);
  return bad > allowed_bad;
}

static void passive_cb(const mjModel* m, mjData* d) { (void)m; (void)d; }

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: dropin <model.mjb> [nstate] [allowed]\n"); return 2; }
  const int n = argc > 2 ? atoi(argv[2]) : 512;
  if (argc > 3) allowed_bad = (size_t)atoi(argv[3]);
  mjModel* m = mj_loadModel(argv[1], NULL);
  if (!m) { fprintf(stderr, "mj_loadModel(%s) failed\n", argv[1]); return 2; }
  mjData* d = mj_makeData(m);
  mjtNum* qpos = malloc(sizeof(mjtNum)*(size_t)n*m->nq);
  mjtNum* qvel = malloc(sizeof(mjtNum)*(size_t)n*m->nv);
  mjtNum* qacc = malloc(sizeof(mjtNum)*(size_t)n*m->nv);
  mjtNum* ref = malloc(sizeof(mjtNum)*(size_t)n*m->nv);
  mjtNum* got = malloc(sizeof(mjtNum)*(size_t)n*m->nv);
  make_states(m, n, qpos, qvel, qacc);
  int fail = 0;

  char err[1024] = "";
  mjbData* bd = mjb_makeData(m, n, 0, 0, 0, 0, err, sizeof err);
  if (!bd) { fprintf(stderr, "mjb_makeData: %s\n", err); return 1; }

  /* 1. as loaded */
  cpu_loop(m, d, n, qpos, qvel, qacc, ref);
  if (mjb_setState(bd, n, qpos, qvel, qacc) || mjb_inverse(m, bd, n) < 0 || mjb_getQfrcInverse(bd, got)) {
    fprintf(stderr, "mjb_inverse: %s\n", mjb_lastError(bd)); return 1;
  }
  fail |= compare("model as loaded", got, ref, (size_t)n*m->nv);

  /* 2. the caller toggles m->opt between calls, as with mj_inverse: gravity off, elliptic cones */
  m->opt.disableflags |= mjDSBL_GRAVITY;
  m->opt.cone = mjCONE_ELLIPTIC;
  cpu_loop(m, d, n, qpos, qvel, qacc, ref);
  if (mjb_inverse(m, bd, n) < 0 || mjb_getQfrcInverse(bd, got)) {
    fprintf(stderr, "mjb_inverse after m->opt change: %s\n", mjb_lastError(bd)); return 1;
  }
  fail |= compare("after m->opt change (no gravity, elliptic)", got, ref, (size_t)n*m->nv);
  m->opt.disableflags &= ~mjDSBL_GRAVITY;
  m->opt.cone = mjCONE_PYRAMIDAL;
  mjb_deleteData(bd);

  /* 3. a global callback on the path: refused at upload */
  mjcb_passive = passive_cb;
  bd = mjb_makeData(m, n, 0, 0, 0, 0, err, sizeof err);
  mjcb_passive = NULL;
  if (bd) { printf("mjb_makeData accepted a model while mjcb_passive was set\n"); fail = 1; mjb_deleteData(bd); }
  else printf("with mjcb_passive set: refused (%s)\n", err);

  free(qpos); free(qvel); free(qacc); free(ref); free(got);
  mj_deleteData(d);
  mj_deleteModel(m);
  printf(fail ? "DROPIN FAILED\n" : "DROPIN OK\n");
  return fail;
}
