/* mjinv_oracle.c -- plain-C CPU restatement of the reference's mj_inverse path.
 *
 * TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * anything under oracle/; the product never links or calls this file.
 *
 * It restates the algorithm in the reference's OWN formulation -- dense constraint Jacobian efc_J
 * built from point Jacobians, J*qvel / J*qacc / J'*force as matrix products, in-place elimination
 * for L'DL -- which is deliberately different from the product's restructured GPU pipeline
 * (Jacobian-free rows, articulated-body factorisation, static candidate pairs). Agreement of the
 * two with the reference library (oracle/_ref) on the golden fixtures is therefore a three-way
 * check. Parity is pinned: tests/test_oracle_restatement.py compares every output below with the
 * dumps of the unmodified reference (tests/golden, made by tests/golden/make_golden.py).
 *
 * Scope: the humanoid-class path -- free/ball/slide/hinge joints, fixed tendons, joint/tendon
 * limits, dof friction loss, plane/sphere/capsule contacts with pyramidal or elliptic cones,
 * springs/dampers, dense Jacobian semantics. One state per call; no threading.
 *
 * Each function cites the reference file:line it follows (relative to /root/reference).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define MINVAL 1E-15
#define MAXCON 1024
#define MAXEFC 4096

/* model arrays by reference name (include/mujoco/mjmodel.h:593-1155); bytes widened to int */
typedef struct {
  int nq, nv, nbody, njnt, ngeom, ntendon, nwrap, nexclude, nM, nC;
  int disableflags, cone;
  double timestep, impratio, gravity[3];
  const int *body_parentid, *body_rootid, *body_weldid, *body_jntnum, *body_jntadr, *body_dofnum,
      *body_dofadr, *body_geomnum, *body_geomadr, *body_sameframe, *body_contype,
      *body_conaffinity, *body_bvhadr;
  const int *jnt_type, *jnt_qposadr, *jnt_dofadr, *jnt_bodyid, *jnt_limited;
  const int *dof_bodyid, *dof_jntid, *dof_parentid, *dof_Madr, *dof_simplenum;
  const int *geom_type, *geom_bodyid, *geom_contype, *geom_conaffinity, *geom_condim,
      *geom_priority, *geom_sameframe;
  const int *tendon_adr, *tendon_num, *tendon_limited, *wrap_objid, *exclude_signature;
  const double *qpos0, *qpos_spring, *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass,
      *body_inertia, *body_invweight0;
  const double *jnt_pos, *jnt_axis, *jnt_stiffness, *jnt_range, *jnt_margin, *jnt_solref,
      *jnt_solimp;
  const double *dof_armature, *dof_damping, *dof_frictionloss, *dof_invweight0, *dof_solref,
      *dof_solimp, *dof_M0;
  const double *geom_size, *geom_rbound, *geom_pos, *geom_quat, *geom_friction, *geom_margin,
      *geom_gap, *geom_solmix, *geom_solref, *geom_solimp;
  const double *tendon_range, *tendon_margin, *tendon_stiffness, *tendon_damping,
      *tendon_lengthspring, *tendon_invweight0, *tendon_solref_lim, *tendon_solimp_lim, *wrap_prm;
} OrcModel;

typedef struct {
  double* qfrc_inverse;   /* nv */
  double* qM;             /* nM   (may be NULL) */
  double* qLD;            /* nM   (may be NULL) */
  double* qLDiagInv;      /* nv   (may be NULL) */
  int* counts;            /* ncon, ne, nf, nl, nefc */
  int* contact_geom;      /* 2*maxcon, -1 padded (may be NULL) */
  int maxcon;
  int* efc_type;          /* maxefc, -1 padded (may be NULL) */
  int* efc_id;
  double* efc_force;
  int maxefc;
} OrcOut;

/* ---------------------------------------------------------------- small math (engine_util_*) */
static double dot3(const double* a, const double* b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
static void cross(double* r, const double* a, const double* b) {
  double t[3] = {a[1]*b[2] - a[2]*b[1], a[2]*b[0] - a[0]*b[2], a[0]*b[1] - a[1]*b[0]};
  memcpy(r, t, sizeof(t));
}
static double normalize3(double* v) {                       /* engine_util_blas.c:123 */
  double n = sqrt(v[0]*v[0] + v[1]*v[1] + v[2]*v[2]);
  if (n < MINVAL) { v[0] = 1; v[1] = 0; v[2] = 0; }
  else { double inv = 1/n; v[0] *= inv; v[1] *= inv; v[2] *= inv; }
  return n;
}
static void normalize4(double* v) {                         /* engine_util_blas.c:269 */
  double n = sqrt(v[0]*v[0] + v[1]*v[1] + v[2]*v[2] + v[3]*v[3]);
  if (n < MINVAL) { v[0] = 1; v[1] = v[2] = v[3] = 0; }
  else if (fabs(n - 1) > MINVAL) { double inv = 1/n; for (int i = 0; i < 4; i++) v[i] *= inv; }
}
static void rot_vec_quat(double* r, const double* v, const double* q) {   /* util_spatial.c:25 */
  if (v[0] == 0 && v[1] == 0 && v[2] == 0) { r[0] = r[1] = r[2] = 0; return; }
  if (q[0] == 1 && q[1] == 0 && q[2] == 0 && q[3] == 0) { memcpy(r, v, 24); return; }
  double t[3] = {q[0]*v[0] + q[2]*v[2] - q[3]*v[1], q[0]*v[1] + q[3]*v[0] - q[1]*v[2],
                 q[0]*v[2] + q[1]*v[1] - q[2]*v[0]};
  double o[3] = {v[0] + 2*(q[2]*t[2] - q[3]*t[1]), v[1] + 2*(q[3]*t[0] - q[1]*t[2]),
                 v[2] + 2*(q[1]*t[1] - q[2]*t[0])};
  memcpy(r, o, 24);
}
static void mul_quat(double* r, const double* a, const double* b) {       /* util_spatial.c:65 */
  double t[4] = {a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3],
                 a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2],
                 a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1],
                 a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0]};
  memcpy(r, t, 32);
}
static void quat2mat(double* m, const double* q) {                        /* util_spatial.c:149 */
  if (q[0] == 1 && q[1] == 0 && q[2] == 0 && q[3] == 0) {
    memset(m, 0, 72); m[0] = m[4] = m[8] = 1; return;
  }
  double q00 = q[0]*q[0], q01 = q[0]*q[1], q02 = q[0]*q[2], q03 = q[0]*q[3], q11 = q[1]*q[1],
         q12 = q[1]*q[2], q13 = q[1]*q[3], q22 = q[2]*q[2], q23 = q[2]*q[3], q33 = q[3]*q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2*(q12 - q03); m[2] = 2*(q13 + q02); m[3] = 2*(q12 + q03);
  m[5] = 2*(q23 - q01); m[6] = 2*(q13 - q02); m[7] = 2*(q23 + q01);
}
static void mat_vec3(double* r, const double* m, const double* v) {
  double t[3] = {m[0]*v[0] + m[1]*v[1] + m[2]*v[2], m[3]*v[0] + m[4]*v[1] + m[5]*v[2],
                 m[6]*v[0] + m[7]*v[1] + m[8]*v[2]};
  memcpy(r, t, 24);
}
static double dotn(const double* a, const double* b, int n) {             /* util_blas.c:677 */
  double r0 = 0, r1 = 0, r2 = 0, r3 = 0, res;
  int i = 0;
  for (; i <= n - 4; i += 4) { r0 += a[i]*b[i]; r1 += a[i+1]*b[i+1]; r2 += a[i+2]*b[i+2]; r3 += a[i+3]*b[i+3]; }
  res = (r0 + r2) + (r1 + r3);
  int left = n - i;
  if (left == 3) res += a[i]*b[i] + a[i+1]*b[i+1] + a[i+2]*b[i+2];
  else if (left == 2) res += a[i]*b[i] + a[i+1]*b[i+1];
  else if (left == 1) res += a[i]*b[i];
  return res;
}
static void quat2vel(double* r, const double* q) {                        /* util_spatial.c:119 */
  double ax[3] = {q[1], q[2], q[3]};
  double s = normalize3(ax);
  double speed = 2*atan2(s, q[0]);
  if (speed > M_PI) speed -= 2*M_PI;
  r[0] = ax[0]*speed; r[1] = ax[1]*speed; r[2] = ax[2]*speed;
}
static void mul_inert(double* r, const double* i, const double* v) {      /* util_spatial.c:452 */
  r[0] = i[0]*v[0] + i[3]*v[1] + i[4]*v[2] - i[8]*v[4] + i[7]*v[5];
  r[1] = i[3]*v[0] + i[1]*v[1] + i[5]*v[2] + i[8]*v[3] - i[6]*v[5];
  r[2] = i[4]*v[0] + i[5]*v[1] + i[2]*v[2] - i[7]*v[3] + i[6]*v[4];
  r[3] = i[8]*v[1] - i[7]*v[2] + i[9]*v[3];
  r[4] = i[6]*v[2] - i[8]*v[0] + i[9]*v[4];
  r[5] = i[7]*v[0] - i[6]*v[1] + i[9]*v[5];
}
static void cross_motion(double* r, const double* w, const double* v) {   /* util_spatial.c:385 */
  r[0] = -w[2]*v[1] + w[1]*v[2]; r[1] = w[2]*v[0] - w[0]*v[2]; r[2] = -w[1]*v[0] + w[0]*v[1];
  r[3] = -w[2]*v[4] + w[1]*v[5]; r[4] = w[2]*v[3] - w[0]*v[5]; r[5] = -w[1]*v[3] + w[0]*v[4];
  r[3] += -w[5]*v[1] + w[4]*v[2]; r[4] += w[5]*v[0] - w[3]*v[2]; r[5] += -w[4]*v[0] + w[3]*v[1];
}
static void cross_force(double* r, const double* w, const double* f) {    /* util_spatial.c:401 */
  r[0] = -w[2]*f[1] + w[1]*f[2]; r[1] = w[2]*f[0] - w[0]*f[2]; r[2] = -w[1]*f[0] + w[0]*f[1];
  r[3] = -w[2]*f[4] + w[1]*f[5]; r[4] = w[2]*f[3] - w[0]*f[5]; r[5] = -w[1]*f[3] + w[0]*f[4];
  r[0] += -w[5]*f[4] + w[4]*f[5]; r[1] += w[5]*f[3] - w[3]*f[5]; r[2] += -w[4]*f[3] + w[3]*f[4];
}

/* ---------------------------------------------------------------- per-state workspace */
typedef struct {
  double *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *gpos, *gmat, *com, *cinert,
      *cdof, *cvel, *cdof_dot, *ten_length, *ten_J, *qfrc_passive;
  /* contacts */
  int ncon;
  int con_geom[MAXCON][2], con_dim[MAXCON], con_exclude[MAXCON];
  double con_dist[MAXCON], con_pos[MAXCON][3], con_frame[MAXCON][9], con_includemargin[MAXCON],
      con_friction[MAXCON][5], con_solref[MAXCON][2], con_solimp[MAXCON][5], con_mu[MAXCON];
  /* constraints (dense) */
  int nefc, ne, nf, nl;
  int efc_type[MAXEFC], efc_id[MAXEFC];
  double *J;   /* MAXEFC x nv */
  double efc_pos[MAXEFC], efc_margin[MAXEFC], efc_floss[MAXEFC], efc_dA[MAXEFC], efc_R[MAXEFC],
      efc_D[MAXEFC], efc_K[MAXEFC], efc_B[MAXEFC], efc_imp[MAXEFC], efc_vel[MAXEFC],
      efc_aref[MAXEFC], efc_force[MAXEFC];
} Work;

/* mj_kinematics (engine_core_smooth.c:38-178) */
static void kinematics(const OrcModel* m, Work* w, const double* qpos) {
  memset(w->xpos, 0, 24); memset(w->xipos, 0, 24);
  w->xquat[0] = 1; w->xquat[1] = w->xquat[2] = w->xquat[3] = 0;
  memset(w->xmat, 0, 72); memset(w->ximat, 0, 72);
  w->xmat[0] = w->xmat[4] = w->xmat[8] = 1; w->ximat[0] = w->ximat[4] = w->ximat[8] = 1;
  for (int i = 1; i < m->nbody; i++) {
    double pos[3], quat[4];
    int ja = m->body_jntadr[i], jn = m->body_jntnum[i];
    if (jn == 1 && m->jnt_type[ja] == 0) {
      int qa = m->jnt_qposadr[ja];
      memcpy(pos, qpos + qa, 24); memcpy(quat, qpos + qa + 3, 32);
      normalize4(quat);
      memcpy(w->xanchor + 3*ja, pos, 24); memcpy(w->xaxis + 3*ja, m->jnt_axis + 3*ja, 24);
    } else {
      int pid = m->body_parentid[i];
      if (pid) {
        mat_vec3(pos, w->xmat + 9*pid, m->body_pos + 3*i);
        for (int k = 0; k < 3; k++) pos[k] += w->xpos[3*pid + k];
        mul_quat(quat, w->xquat + 4*pid, m->body_quat + 4*i);
      } else {
        memcpy(pos, m->body_pos + 3*i, 24); memcpy(quat, m->body_quat + 4*i, 32);
      }
      for (int j = 0; j < jn; j++) {
        int jid = ja + j, qa = m->jnt_qposadr[jid], jt = m->jnt_type[jid];
        double ax[3], an[3];
        rot_vec_quat(ax, m->jnt_axis + 3*jid, quat);
        rot_vec_quat(an, m->jnt_pos + 3*jid, quat);
        for (int k = 0; k < 3; k++) an[k] += pos[k];
        if (jt == 2) {
          double d = qpos[qa] - m->qpos0[qa];
          for (int k = 0; k < 3; k++) pos[k] += ax[k]*d;
        } else {
          double ql[4], vec[3];
          if (jt == 1) { memcpy(ql, qpos + qa, 32); normalize4(ql); }
          else {
            double ang = qpos[qa] - m->qpos0[qa];
            if (ang == 0) { ql[0] = 1; ql[1] = ql[2] = ql[3] = 0; }
            else {
              double s = sin(ang*0.5);
              ql[0] = cos(ang*0.5);
              for (int k = 0; k < 3; k++) ql[1 + k] = m->jnt_axis[3*jid + k]*s;
            }
          }
          mul_quat(quat, quat, ql);
          rot_vec_quat(vec, m->jnt_pos + 3*jid, quat);
          for (int k = 0; k < 3; k++) pos[k] = an[k] - vec[k];
        }
        memcpy(w->xanchor + 3*jid, an, 24); memcpy(w->xaxis + 3*jid, ax, 24);
      }
    }
    normalize4(quat);
    memcpy(w->xquat + 4*i, quat, 32); memcpy(w->xpos + 3*i, pos, 24);
    quat2mat(w->xmat + 9*i, quat);
  }
  /* mj_local2Global (engine_support.c:1565-1607) for inertial frames and geoms */
  for (int i = 1; i < m->nbody; i++) {
    int sf = m->body_sameframe[i];
    if (sf == 1) memcpy(w->xipos + 3*i, w->xpos + 3*i, 24);
    else {
      mat_vec3(w->xipos + 3*i, w->xmat + 9*i, m->body_ipos + 3*i);
      for (int k = 0; k < 3; k++) w->xipos[3*i + k] += w->xpos[3*i + k];
    }
    if (sf == 0) { double t[4]; mul_quat(t, w->xquat + 4*i, m->body_iquat + 4*i); quat2mat(w->ximat + 9*i, t); }
    else memcpy(w->ximat + 9*i, w->xmat + 9*i, 72);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g], sf = m->geom_sameframe[g];
    if (sf == 1) memcpy(w->gpos + 3*g, w->xpos + 3*b, 24);
    else if (sf == 2) memcpy(w->gpos + 3*g, w->xipos + 3*b, 24);
    else {
      mat_vec3(w->gpos + 3*g, w->xmat + 9*b, m->geom_pos + 3*g);
      for (int k = 0; k < 3; k++) w->gpos[3*g + k] += w->xpos[3*b + k];
    }
    if (sf == 0) { double t[4]; mul_quat(t, w->xquat + 4*b, m->geom_quat + 4*g); quat2mat(w->gmat + 9*g, t); }
    else if (sf == 1 || sf == 3) memcpy(w->gmat + 9*g, w->xmat + 9*b, 72);
    else memcpy(w->gmat + 9*g, w->ximat + 9*b, 72);
  }
}

/* mj_comPos (engine_core_smooth.c:183-270) */
static void com_pos(const OrcModel* m, Work* w) {
  int nb = m->nbody;
  double* ms = (double*)calloc((size_t)nb, sizeof(double));
  memset(w->com, 0, sizeof(double)*3*nb);
  for (int i = nb - 1; i >= 0; i--) {
    for (int k = 0; k < 3; k++) w->com[3*i + k] += w->xipos[3*i + k]*m->body_mass[i];
    ms[i] += m->body_mass[i];
    if (i) {
      int j = m->body_parentid[i];
      for (int k = 0; k < 3; k++) w->com[3*j + k] += w->com[3*i + k];
      ms[j] += ms[i];
    }
    if (ms[i] < MINVAL) memcpy(w->com + 3*i, w->xipos + 3*i, 24);
    else { double inv = 1.0/fmax(MINVAL, ms[i]); for (int k = 0; k < 3; k++) w->com[3*i + k] *= inv; }
  }
  free(ms);
  memset(w->cinert, 0, 80);
  for (int i = 1; i < nb; i++) {                         /* mju_inertCom util_spatial.c:417 */
    const double* in = m->body_inertia + 3*i; const double* mat = w->ximat + 9*i;
    double mass = m->body_mass[i], d[3], *r = w->cinert + 10*i;
    for (int k = 0; k < 3; k++) d[k] = w->xipos[3*i + k] - w->com[3*m->body_rootid[i] + k];
    double t[9] = {mat[0]*in[0], mat[3]*in[0], mat[6]*in[0], mat[1]*in[1], mat[4]*in[1], mat[7]*in[1],
                   mat[2]*in[2], mat[5]*in[2], mat[8]*in[2]};
    r[0] = mat[0]*t[0] + mat[1]*t[3] + mat[2]*t[6]; r[1] = mat[3]*t[1] + mat[4]*t[4] + mat[5]*t[7];
    r[2] = mat[6]*t[2] + mat[7]*t[5] + mat[8]*t[8]; r[3] = mat[0]*t[1] + mat[1]*t[4] + mat[2]*t[7];
    r[4] = mat[0]*t[2] + mat[1]*t[5] + mat[2]*t[8]; r[5] = mat[3]*t[2] + mat[4]*t[5] + mat[5]*t[8];
    r[0] += mass*(d[1]*d[1] + d[2]*d[2]); r[1] += mass*(d[0]*d[0] + d[2]*d[2]);
    r[2] += mass*(d[0]*d[0] + d[1]*d[1]); r[3] -= mass*d[0]*d[1]; r[4] -= mass*d[0]*d[2];
    r[5] -= mass*d[1]*d[2]; r[6] = mass*d[0]; r[7] = mass*d[1]; r[8] = mass*d[2]; r[9] = mass;
  }
  for (int j = 0; j < m->njnt; j++) {
    int da = 6*m->jnt_dofadr[j], bi = m->jnt_bodyid[j], jt = m->jnt_type[j], skip = 0;
    double off[3];
    for (int k = 0; k < 3; k++) off[k] = w->com[3*m->body_rootid[bi] + k] - w->xanchor[3*j + k];
    if (jt == 0) {
      memset(w->cdof + da, 0, 18*8);
      for (int i = 0; i < 3; i++) w->cdof[da + 3 + 7*i] = 1;
      skip = 18;
    }
    if (jt == 0 || jt == 1) {
      for (int i = 0; i < 3; i++) {
        double ax[3] = {w->xmat[9*bi + i], w->xmat[9*bi + i + 3], w->xmat[9*bi + i + 6]};
        memcpy(w->cdof + da + skip + 6*i, ax, 24);
        cross(w->cdof + da + skip + 6*i + 3, ax, off);
      }
    } else if (jt == 2) {
      memset(w->cdof + da, 0, 24); memcpy(w->cdof + da + 3, w->xaxis + 3*j, 24);
    } else {
      memcpy(w->cdof + da, w->xaxis + 3*j, 24); cross(w->cdof + da + 3, w->xaxis + 3*j, off);
    }
  }
}

/* fixed tendons of mj_tendon (engine_core_smooth.c:699-723), dense ten_J */
static void tendons(const OrcModel* m, Work* w, const double* qpos) {
  memset(w->ten_J, 0, sizeof(double)*m->ntendon*m->nv);
  for (int t = 0; t < m->ntendon; t++) {
    w->ten_length[t] = 0;
    for (int j = 0; j < m->tendon_num[t]; j++) {
      int k = m->wrap_objid[m->tendon_adr[t] + j];
      w->ten_length[t] += m->wrap_prm[m->tendon_adr[t] + j]*qpos[m->jnt_qposadr[k]];
      w->ten_J[t*m->nv + m->jnt_dofadr[k]] = m->wrap_prm[m->tendon_adr[t] + j];
    }
  }
}

/* mj_comVel (engine_core_smooth.c:1833-1896) */
static void com_vel(const OrcModel* m, Work* w, const double* qvel) {
  memset(w->cvel, 0, 48);
  for (int i = 1; i < m->nbody; i++) {
    int bda = m->body_dofadr[i], n = m->body_dofnum[i];
    double v[6];
    memcpy(v, w->cvel + 6*m->body_parentid[i], 48);
    for (int j = 0; j < n; j++) {
      int jt = m->jnt_type[m->dof_jntid[bda + j]];
      if (jt == 0) {
        memset(w->cdof_dot + 6*bda, 0, 18*8);
        for (int r = 0; r < 3; r++) for (int k = 0; k < 6; k++) v[k] += w->cdof[6*(bda + r) + k]*qvel[bda + r];
        j += 3;
      }
      if (jt == 0 || jt == 1) {
        for (int r = 0; r < 3; r++) cross_motion(w->cdof_dot + 6*(bda + j + r), v, w->cdof + 6*(bda + j + r));
        for (int r = 0; r < 3; r++) for (int k = 0; k < 6; k++) v[k] += w->cdof[6*(bda + j + r) + k]*qvel[bda + j + r];
        j += 2;
      } else {
        cross_motion(w->cdof_dot + 6*(bda + j), v, w->cdof + 6*(bda + j));
        for (int k = 0; k < 6; k++) v[k] += w->cdof[6*(bda + j) + k]*qvel[bda + j];
      }
    }
    memcpy(w->cvel + 6*i, v, 48);
  }
}

/* mj_passive: springs and dampers (engine_passive.c:57-115, 336-375, 436-462) */
static void passive(const OrcModel* m, Work* w, const double* qpos, const double* qvel) {
  int nv = m->nv;
  memset(w->qfrc_passive, 0, sizeof(double)*nv);
  if (m->disableflags & (1 << 5)) return;
  for (int i = 0; i < m->njnt; i++) {
    double k = m->jnt_stiffness[i];
    if (k == 0) continue;
    int pa = m->jnt_qposadr[i], da = m->jnt_dofadr[i], jt = m->jnt_type[i];
    if (jt == 0) { for (int r = 0; r < 3; r++) w->qfrc_passive[da + r] = -k*(qpos[pa + r] - m->qpos_spring[pa + r]); da += 3; pa += 3; }
    if (jt == 0 || jt == 1) {
      double q[4], qn[4], qd[4], dif[3];
      memcpy(q, qpos + pa, 32); normalize4(q);
      qn[0] = m->qpos_spring[pa]; for (int r = 1; r < 4; r++) qn[r] = -m->qpos_spring[pa + r];
      mul_quat(qd, qn, q); quat2vel(dif, qd);
      for (int r = 0; r < 3; r++) w->qfrc_passive[da + r] = -k*dif[r];
    } else w->qfrc_passive[da] = -k*(qpos[pa] - m->qpos_spring[pa]);
  }
  for (int i = 0; i < nv; i++) if (m->dof_damping[i] != 0) w->qfrc_passive[i] += -m->dof_damping[i]*qvel[i];
  for (int t = 0; t < m->ntendon; t++) {
    double ks = m->tendon_stiffness[t], kd = m->tendon_damping[t];
    if (ks == 0 && kd == 0) continue;
    double len = w->ten_length[t], fs = 0, vel = dotn(w->ten_J + t*nv, qvel, nv);
    if (len > m->tendon_lengthspring[2*t + 1]) fs = ks*(m->tendon_lengthspring[2*t + 1] - len);
    else if (len < m->tendon_lengthspring[2*t]) fs = ks*(m->tendon_lengthspring[2*t] - len);
    for (int i = 0; i < nv; i++) w->qfrc_passive[i] += w->ten_J[t*nv + i]*fs + w->ten_J[t*nv + i]*(-kd*vel);
  }
}

/* ---------------------------------------------------------------- collision */
static int filter_bitmask(int ct1, int ca1, int ct2, int ca2) { return !(ct1 & ca2) && !(ct2 & ca1); }

static void make_frame(double* f) {                               /* util_spatial.c:526 */
  normalize3(f);
  if (sqrt(f[3]*f[3] + f[4]*f[4] + f[5]*f[5]) < 0.5) {
    f[3] = f[4] = f[5] = 0;
    if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  }
  double d = dot3(f, f + 3);
  for (int k = 0; k < 3; k++) f[3 + k] -= f[k]*d;
  normalize3(f + 3);
  cross(f + 6, f, f + 3);
}

typedef struct { double dist, pos[3], frame[9]; } Hit;

static int plane_sphere(Hit* h, double margin, const double* p1, const double* m1, const double* p2, double r) {
  double n[3] = {m1[2], m1[5], m1[8]}, t[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double cd = dot3(t, n);
  if (cd > margin + r) return 0;
  h->dist = cd - r;
  memcpy(h->frame, n, 24); memset(h->frame + 3, 0, 24);
  for (int k = 0; k < 3; k++) h->pos[k] = p2[k] + n[k]*(-h->dist/2 - r);
  return 1;
}
static int sphere_sphere(Hit* h, double margin, const double* p1, const double* m1, double r1,
                         const double* p2, const double* m2, double r2) {
  double d[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
  double c2 = dot3(d, d), md = margin + r1 + r2;
  if (c2 > md*md) return 0;
  h->dist = sqrt(c2) - r1 - r2;
  for (int k = 0; k < 3; k++) h->frame[k] = p2[k] - p1[k];
  if (normalize3(h->frame) < MINVAL) {
    double a1[3] = {m1[2], m1[5], m1[8]}, a2[3] = {m2[2], m2[5], m2[8]};
    cross(h->frame, a1, a2); normalize3(h->frame);
  }
  for (int k = 0; k < 3; k++) h->pos[k] = h->frame[k]*(r1 + h->dist/2) + p1[k];
  memset(h->frame + 3, 0, 24);
  return 1;
}
static double clipd(double x, double lo, double hi) { return fmax(lo, fmin(hi, x)); }

/* narrow phase dispatch for plane/sphere/capsule (engine_collision_primitive.c:28-460) */
static int narrow(const OrcModel* m, const Work* w, int g1, int g2, double margin, Hit* h) {
  int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
  const double *p1 = w->gpos + 3*g1, *p2 = w->gpos + 3*g2, *m1 = w->gmat + 9*g1, *m2 = w->gmat + 9*g2;
  const double *s1 = m->geom_size + 3*g1, *s2 = m->geom_size + 3*g2;
  if (t1 == 0 && t2 == 2) return plane_sphere(h, margin, p1, m1, p2, s2[0]);
  if (t1 == 0 && t2 == 3) {
    double ax[3] = {m2[2], m2[5], m2[8]}, a[3], b[3];
    for (int k = 0; k < 3; k++) { a[k] = p2[k] + s2[1]*ax[k]; b[k] = p2[k] - s2[1]*ax[k]; }
    int n1 = plane_sphere(h, margin, p1, m1, a, s2[0]);
    int n2 = plane_sphere(h + n1, margin, p1, m1, b, s2[0]);
    for (int i = 0; i < n1 + n2; i++) memcpy(h[i].frame + 3, ax, 24);
    return n1 + n2;
  }
  if (t1 == 2 && t2 == 2) return sphere_sphere(h, margin, p1, m1, s1[0], p2, m2, s2[0]);
  if (t1 == 2 && t2 == 3) {
    double ax[3] = {m2[2], m2[5], m2[8]}, v[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    double x = clipd(dot3(ax, v), -s2[1], s2[1]);
    for (int k = 0; k < 3; k++) v[k] = ax[k]*x + p2[k];
    return sphere_sphere(h, margin, p1, m1, s1[0], v, m2, s2[0]);
  }
  if (t1 == 3 && t2 == 3) {
    double a1[3], a2[3], d[3], v1[3], v2[3];
    for (int k = 0; k < 3; k++) { a1[k] = m1[2 + 3*k]*s1[1]; a2[k] = m2[2 + 3*k]*s2[1]; d[k] = p1[k] - p2[k]; }
    double ma = dot3(a1, a1), mb = -dot3(a1, a2), mc = dot3(a2, a2), u = -dot3(a1, d), v = dot3(a2, d);
    double det = ma*mc - mb*mb;
    if (fabs(det) >= MINVAL) {
      double x1 = (mc*u - mb*v)/det, x2 = (ma*v - mb*u)/det;
      if (x1 > 1) { x1 = 1; x2 = (v - mb)/mc; } else if (x1 < -1) { x1 = -1; x2 = (v + mb)/mc; }
      if (x2 > 1) { x2 = 1; x1 = clipd((u - mb)/ma, -1, 1); }
      else if (x2 < -1) { x2 = -1; x1 = clipd((u + mb)/ma, -1, 1); }
      for (int k = 0; k < 3; k++) { v1[k] = a1[k]*x1 + p1[k]; v2[k] = a2[k]*x2 + p2[k]; }
      return sphere_sphere(h, margin, v1, m1, s1[0], v2, m2, s2[0]);
    }
    int n = 0;
    double x2, x1;
    for (int k = 0; k < 3; k++) v1[k] = p1[k] + a1[k];
    x2 = clipd((v - mb)/mc, -1, 1);
    for (int k = 0; k < 3; k++) v2[k] = a2[k]*x2 + p2[k];
    n += sphere_sphere(h + n, margin, v1, m1, s1[0], v2, m2, s2[0]);
    for (int k = 0; k < 3; k++) v1[k] = p1[k] - a1[k];
    x2 = clipd((v + mb)/mc, -1, 1);
    for (int k = 0; k < 3; k++) v2[k] = a2[k]*x2 + p2[k];
    n += sphere_sphere(h + n, margin, v1, m1, s1[0], v2, m2, s2[0]);
    if (n >= 2) return n;
    for (int k = 0; k < 3; k++) v2[k] = p2[k] + a2[k];
    x1 = clipd((u - mb)/ma, -1, 1);
    for (int k = 0; k < 3; k++) v1[k] = a1[k]*x1 + p1[k];
    n += sphere_sphere(h + n, margin, v1, m1, s1[0], v2, m2, s2[0]);
    if (n >= 2) return n;
    for (int k = 0; k < 3; k++) v2[k] = p2[k] - a2[k];
    x1 = clipd((u + mb)/ma, -1, 1);
    for (int k = 0; k < 3; k++) v1[k] = a1[k]*x1 + p1[k];
    n += sphere_sphere(h + n, margin, v1, m1, s1[0], v2, m2, s2[0]);
    return n;
  }
  return -1;   /* outside the restated set */
}

/* mj_collideGeoms (engine_collision_driver.c:1440-1632) incl. mj_contactParam (:1289-1382) */
static int collide_geoms(const OrcModel* m, Work* w, int g1, int g2) {
  if (m->geom_type[g1] > m->geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
  int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
  if (t1 == 0 && (t2 == 0 || t2 == 1)) return 0;                 /* no collision function */
  if (filter_bitmask(m->geom_contype[g1], m->geom_conaffinity[g1], m->geom_contype[g2], m->geom_conaffinity[g2])) return 0;
  double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
  /* mj_filterSphere (:146-163) */
  double rb1 = m->geom_rbound[g1], rb2 = m->geom_rbound[g2];
  const double *p1 = w->gpos + 3*g1, *p2 = w->gpos + 3*g2;
  if (rb1 > 0 && rb2 > 0) {
    double d[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]}, b = rb1 + rb2 + margin;
    if (d[0]*d[0] + d[1]*d[1] + d[2]*d[2] > b*b) return 0;
  } else if (t1 == 0 && rb2 > 0) {
    double n[3] = {w->gmat[9*g1 + 2], w->gmat[9*g1 + 5], w->gmat[9*g1 + 8]};
    double d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    if (dot3(d, n) > margin + rb2) return 0;
  }
  Hit h[4];
  int num = narrow(m, w, g1, g2, margin, h);
  if (num < 0) return -1;
  if (!num) return 0;
  /* parameters */
  int condim; double gap = fmax(m->geom_gap[g1], m->geom_gap[g2]), solref[2], solimp[5], fri[3];
  int pr1 = m->geom_priority[g1], pr2 = m->geom_priority[g2];
  if (pr1 != pr2) {
    int g = pr1 > pr2 ? g1 : g2;
    condim = m->geom_condim[g];
    memcpy(solref, m->geom_solref + 2*g, 16); memcpy(solimp, m->geom_solimp + 5*g, 40);
    memcpy(fri, m->geom_friction + 3*g, 24);
  } else {
    condim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
    double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2], mix;
    if (s1 >= MINVAL && s2 >= MINVAL) mix = s1/(s1 + s2);
    else if (s1 < MINVAL && s2 < MINVAL) mix = 0.5;
    else if (s1 < MINVAL) mix = 0.0; else mix = 1.0;
    const double *r1 = m->geom_solref + 2*g1, *r2 = m->geom_solref + 2*g2;
    for (int i = 0; i < 2; i++) solref[i] = (r1[0] > 0 && r2[0] > 0) ? mix*r1[i] + (1 - mix)*r2[i] : fmin(r1[i], r2[i]);
    for (int i = 0; i < 5; i++) solimp[i] = mix*m->geom_solimp[5*g1 + i] + (1 - mix)*m->geom_solimp[5*g2 + i];
    for (int i = 0; i < 3; i++) fri[i] = fmax(m->geom_friction[3*g1 + i], m->geom_friction[3*g2 + i]);
  }
  for (int i = 0; i < num; i++) {
    if (w->ncon >= MAXCON) return -2;
    int c = w->ncon++;
    w->con_geom[c][0] = g1; w->con_geom[c][1] = g2;
    w->con_dim[c] = condim; w->con_dist[c] = h[i].dist;
    w->con_includemargin[c] = margin - gap;
    memcpy(w->con_pos[c], h[i].pos, 24); memcpy(w->con_frame[c], h[i].frame, 72);
    make_frame(w->con_frame[c]);
    w->con_exclude[c] = h[i].dist >= margin - gap;
    double f5[5] = {fri[0], fri[0], fri[1], fri[2], fri[2]};
    for (int k = 0; k < 5; k++) w->con_friction[c][k] = fmax(1e-5, f5[k]);
    memcpy(w->con_solref[c], solref, 16); memcpy(w->con_solimp[c], solimp, 40);
    w->con_mu[c] = 0;
  }
  return num;
}

/* mj_collision with the broadphase / midphase replaced by an exhaustive body-pair loop in signature
 * order (engine_collision_driver.c:265-484; body filters :168-183, :330-346, :937-990) */
static int collision(const OrcModel* m, Work* w) {
  w->ncon = 0;
  if (m->disableflags & ((1 << 0) | (1 << 4))) return 0;
  int dsbl_fp = m->disableflags & (1 << 9), dsbl_mid = m->disableflags & (1 << 13);
  for (int b1 = 0; b1 < m->nbody; b1++) {
    if (!(m->body_contype[b1] || m->body_conaffinity[b1])) continue;
    for (int b2 = b1 + 1; b2 < m->nbody; b2++) {
      if (!(m->body_contype[b2] || m->body_conaffinity[b2])) continue;
      int w1 = m->body_weldid[b1], w2 = m->body_weldid[b2];
      int pw1 = m->body_weldid[m->body_parentid[w1]], pw2 = m->body_weldid[m->body_parentid[w2]];
      if (w1 == w2) continue;
      if (!dsbl_fp && w1 != 0 && w2 != 0 && (w1 == pw2 || w2 == pw1)) continue;
      /* a static body reaches the narrow phase only through SAP or the plane/world rule: both are
       * conservative, so no pair is lost by testing it here */
      int ct1 = 0, ca1 = 0, ct2 = 0, ca2 = 0;
      for (int g = m->body_geomadr[b1]; g < m->body_geomadr[b1] + m->body_geomnum[b1]; g++) { ct1 |= m->geom_contype[g]; ca1 |= m->geom_conaffinity[g]; }
      for (int g = m->body_geomadr[b2]; g < m->body_geomadr[b2] + m->body_geomnum[b2]; g++) { ct2 |= m->geom_contype[g]; ca2 |= m->geom_conaffinity[g]; }
      if (!(ct1 & ca2) && !(ct2 & ca1)) continue;
      if (filter_bitmask(m->body_contype[b1], m->body_conaffinity[b1], m->body_contype[b2], m->body_conaffinity[b2])) continue;
      int sig = (b1 << 16) + b2, excl = 0;
      for (int e = 0; e < m->nexclude; e++) if (m->exclude_signature[e] == sig) excl = 1;
      if (excl) continue;
      int first = w->ncon;
      for (int g1 = m->body_geomadr[b1]; g1 < m->body_geomadr[b1] + m->body_geomnum[b1]; g1++)
        for (int g2 = m->body_geomadr[b2]; g2 < m->body_geomadr[b2] + m->body_geomnum[b2]; g2++) {
          int r = collide_geoms(m, w, g1, g2);
          if (r < 0) return r;
        }
      /* midphase pairs: contacts stably sorted by stored (geom0, geom1) (:227-257, :362-375) */
      int multi = !(m->body_geomnum[b1] == 1 && m->body_geomnum[b2] == 1);
      if (multi && !dsbl_mid && m->body_bvhadr[b1] >= 0 && m->body_bvhadr[b2] >= 0) {
        for (int i = first + 1; i < w->ncon; i++) {          /* insertion sort = stable */
          for (int j = i; j > first; j--) {
            int a0 = w->con_geom[j - 1][0], a1 = w->con_geom[j - 1][1], c0 = w->con_geom[j][0], c1 = w->con_geom[j][1];
            if (a0 < c0 || (a0 == c0 && a1 <= c1)) break;
#define SWAPV(arr, sz) { char tmp_[sz]; memcpy(tmp_, &arr[j - 1], sz); memcpy(&arr[j - 1], &arr[j], sz); memcpy(&arr[j], tmp_, sz); }
            SWAPV(w->con_geom, 8) SWAPV(w->con_dim, 4) SWAPV(w->con_exclude, 4) SWAPV(w->con_dist, 8)
            SWAPV(w->con_pos, 24) SWAPV(w->con_frame, 72) SWAPV(w->con_includemargin, 8)
            SWAPV(w->con_friction, 40) SWAPV(w->con_solref, 16) SWAPV(w->con_solimp, 40)
          }
        }
      }
    }
  }
  return 0;
}

/* ---------------------------------------------------------------- constraints */
/* mj_jac (engine_support.c:389-439): translational and rotational Jacobian of a point on a body */
static void jac(const OrcModel* m, const Work* w, double* jp, double* jr, const double* pt, int body) {
  int nv = m->nv;
  double off[3];
  memset(jp, 0, sizeof(double)*3*nv); memset(jr, 0, sizeof(double)*3*nv);
  for (int k = 0; k < 3; k++) off[k] = pt[k] - w->com[3*m->body_rootid[body] + k];
  while (body && !m->body_dofnum[body]) body = m->body_parentid[body];
  if (!body) return;
  for (int i = m->body_dofadr[body] + m->body_dofnum[body] - 1; i >= 0; i = m->dof_parentid[i]) {
    const double* cd = w->cdof + 6*i;
    double t[3];
    cross(t, cd, off);
    for (int k = 0; k < 3; k++) { jr[i + k*nv] = cd[k]; jp[i + k*nv] = cd[3 + k] + t[k]; }
  }
}

static void add_row(const OrcModel* m, Work* w, const double* row, double pos, double margin, double floss,
                    int type, int id) {
  int r = w->nefc++;
  memcpy(w->J + (size_t)r*m->nv, row, sizeof(double)*m->nv);
  w->efc_pos[r] = pos; w->efc_margin[r] = margin; w->efc_floss[r] = floss;
  w->efc_type[r] = type; w->efc_id[r] = id;
}

/* mj_instantiateFriction / Limit / Contact (engine_core_constraint.c:768-1131), dense rows */
static int make_constraint(const OrcModel* m, Work* w, const double* qpos) {
  int nv = m->nv;
  w->nefc = w->ne = w->nf = w->nl = 0;
  if (m->disableflags & 1) return 0;
  double* row = (double*)calloc((size_t)nv*16, sizeof(double));
  if (!(m->disableflags & (1 << 2))) {
    for (int i = 0; i < nv; i++) if (m->dof_frictionloss[i] > 0) {
      memset(row, 0, sizeof(double)*nv); row[i] = 1;
      add_row(m, w, row, 0, 0, m->dof_frictionloss[i], 1, i); w->nf++;
    }
  }
  if (!(m->disableflags & (1 << 3))) {
    for (int i = 0; i < m->njnt; i++) {
      if (!m->jnt_limited[i]) continue;
      double margin = m->jnt_margin[i];
      int jt = m->jnt_type[i];
      if (jt == 2 || jt == 3) {
        double value = qpos[m->jnt_qposadr[i]];
        for (int side = -1; side <= 1; side += 2) {
          double dist = side*(m->jnt_range[2*i + (side + 1)/2] - value);
          if (dist < margin) {
            memset(row, 0, sizeof(double)*nv); row[m->jnt_dofadr[i]] = -(double)side;
            add_row(m, w, row, dist, margin, 0, 3, i); w->nl++;
          }
        }
      } else if (jt == 1) {
        double q[4], aa[3];
        memcpy(q, qpos + m->jnt_qposadr[i], 32); normalize4(q); quat2vel(aa, q);
        double value = normalize3(aa), dist = fmax(m->jnt_range[2*i], m->jnt_range[2*i + 1]) - value;
        if (dist < margin) {
          memset(row, 0, sizeof(double)*nv);
          for (int k = 0; k < 3; k++) row[m->jnt_dofadr[i] + k] = -aa[k];
          add_row(m, w, row, dist, margin, 0, 3, i); w->nl++;
        }
      }
    }
    for (int t = 0; t < m->ntendon; t++) {
      if (!m->tendon_limited[t]) continue;
      double value = w->ten_length[t], margin = m->tendon_margin[t];
      for (int side = -1; side <= 1; side += 2) {
        double dist = side*(m->tendon_range[2*t + (side + 1)/2] - value);
        if (dist < margin) {
          for (int i = 0; i < nv; i++) row[i] = w->ten_J[t*nv + i]*(-side);
          add_row(m, w, row, dist, margin, 0, 4, t); w->nl++;
        }
      }
    }
  }
  /* contacts */
  double *j1p = row + nv, *j1r = row + 4*nv, *j2p = row + 7*nv, *j2r = row + 10*nv, *jc = row + 13*nv;
  double* jac6 = (double*)calloc((size_t)6*nv, sizeof(double));
  for (int c = 0; c < w->ncon && nv > 0 && !(m->disableflags & (1 << 4)); c++) {
    if (w->con_exclude[c]) continue;
    int dim = w->con_dim[c];
    if (w->nefc + 2*dim >= MAXEFC) { free(row); free(jac6); return -2; }
    int b1 = m->geom_bodyid[w->con_geom[c][0]], b2 = m->geom_bodyid[w->con_geom[c][1]];
    jac(m, w, j1p, j1r, w->con_pos[c], b1);
    jac(m, w, j2p, j2r, w->con_pos[c], b2);
    /* jac = frame * (jac2 - jac1): rows 0..2 translation, 3..5 rotation */
    for (int r = 0; r < 3; r++) for (int i = 0; i < nv; i++) {
      double sp = 0, sr = 0;
      for (int k = 0; k < 3; k++) {
        sp += w->con_frame[c][3*r + k]*(j2p[i + k*nv] - j1p[i + k*nv]);
        sr += w->con_frame[c][3*r + k]*(j2r[i + k*nv] - j1r[i + k*nv]);
      }
      jac6[r*nv + i] = sp; jac6[(3 + r)*nv + i] = sr;
    }
    if (dim == 1) add_row(m, w, jac6, w->con_dist[c], w->con_includemargin[c], 0, 5, c);
    else if (m->cone == 0) {
      for (int k = 1; k < dim; k++) {
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
          for (int i = 0; i < nv; i++) jc[i] = jac6[i] + sgn*w->con_friction[c][k - 1]*jac6[k*nv + i];
          add_row(m, w, jc, w->con_dist[c], w->con_includemargin[c], 0, 6, c);
        }
      }
    } else {
      for (int k = 0; k < dim; k++) add_row(m, w, jac6 + k*nv, k ? 0 : w->con_dist[c], k ? 0 : w->con_includemargin[c], 0, 7, c);
    }
  }
  free(jac6);
  free(row);
  return 0;
}

static void sol_params(const OrcModel* m, const double* ref_in, const double* imp_in, double* ref, double* imp) {
  ref[0] = ref_in[0]; ref[1] = ref_in[1];                         /* getsolparam :1316-1387 */
  if ((ref[0] > 0) ^ (ref[1] > 0)) { ref[0] = 0.02; ref[1] = 1; }
  if (!(m->disableflags & (1 << 11)) && ref[0] > 0) ref[0] = fmax(ref[0], 2*m->timestep);
  imp[0] = fmin(0.9999, fmax(0.0001, imp_in[0])); imp[1] = fmin(0.9999, fmax(0.0001, imp_in[1]));
  imp[2] = fmax(0, imp_in[2]); imp[3] = fmin(0.9999, fmax(0.0001, imp_in[3])); imp[4] = fmax(1, imp_in[4]);
}
static double powr(double a, double b) { return b == 1 ? a : (b == 2 ? a*a : pow(a, b)); }
static double impedance(const double* s, double pos, double margin) { /* getimpedance :1441 */
  if (s[0] == s[1] || s[2] <= MINVAL) return 0.5*(s[0] + s[1]);
  double x = (pos - margin)/s[2];
  if (x < 0) x = -x;
  if (x >= 1 || x <= 0) return x >= 1 ? s[1] : s[0];
  double y;
  if (s[4] == 1) y = x;
  else if (x <= s[3]) y = (1/powr(s[3], s[4] - 1))*powr(x, s[4]);
  else y = 1 - (1/powr(1 - s[3], s[4] - 1))*powr(1 - x, s[4]);
  return s[0] + y*(s[1] - s[0]);
}

/* mj_diagApprox (:1138-1311) + mj_makeImpedance (:1494-1608) */
static void impedances(const OrcModel* m, Work* w) {
  for (int i = 0; i < w->nefc; i++) {
    int id = w->efc_id[i], tp = w->efc_type[i], dim = 1;
    const double *sr, *si;
    double ref[2], imp5[5];
    if (tp == 1) { sr = m->dof_solref + 2*id; si = m->dof_solimp + 5*id; w->efc_dA[i] = m->dof_invweight0[id]; }
    else if (tp == 3) { sr = m->jnt_solref + 2*id; si = m->jnt_solimp + 5*id; w->efc_dA[i] = m->dof_invweight0[m->jnt_dofadr[id]]; }
    else if (tp == 4) { sr = m->tendon_solref_lim + 2*id; si = m->tendon_solimp_lim + 5*id; w->efc_dA[i] = m->tendon_invweight0[id]; }
    else {
      sr = w->con_solref[id]; si = w->con_solimp[id];
      int b1 = m->geom_bodyid[w->con_geom[id][0]], b2 = m->geom_bodyid[w->con_geom[id][1]], cd = w->con_dim[id];
      double tran = m->body_invweight0[2*b1] + m->body_invweight0[2*b2];
      double rot = m->body_invweight0[2*b1 + 1] + m->body_invweight0[2*b2 + 1];
      if (tp == 5) w->efc_dA[i] = tran;
      else if (tp == 7) { dim = cd; for (int j = 0; j < cd; j++) w->efc_dA[i + j] = j < 3 ? tran : rot; }
      else { dim = 2*(cd - 1); for (int j = 0; j < cd - 1; j++) { double f = w->con_friction[id][j]; w->efc_dA[i + 2*j] = w->efc_dA[i + 2*j + 1] = tran + f*f*(j < 2 ? tran : rot); } }
    }
    sol_params(m, sr, si, ref, imp5);
    double im = impedance(imp5, w->efc_pos[i], w->efc_margin[i]);
    for (int j = 0; j < dim; j++) {
      int r = i + j, fric = (tp == 1) || (tp == 7 && j > 0);
      w->efc_R[r] = fmax(MINVAL, (1 - im)*w->efc_dA[r]/im);
      w->efc_K[r] = fric ? 0 : (ref[0] > 0 ? 1/fmax(MINVAL, imp5[1]*imp5[1]*ref[0]*ref[0]*ref[1]*ref[1]) : -ref[0]/fmax(MINVAL, imp5[1]*imp5[1]));
      w->efc_B[r] = ref[1] > 0 ? 2/fmax(MINVAL, imp5[1]*ref[0]) : -ref[1]/fmax(MINVAL, imp5[1]);
      w->efc_imp[r] = im;
    }
    i += dim - 1;
  }
  for (int i = w->ne + w->nf; i < w->nefc; i++) {
    int tp = w->efc_type[i];
    if (tp != 6 && tp != 7) continue;
    int id = w->efc_id[i], dim = w->con_dim[id];
    const double* fr = w->con_friction[id];
    w->efc_R[i + 1] = w->efc_R[i]/fmax(MINVAL, m->impratio);
    w->con_mu[id] = fr[0]*sqrt(w->efc_R[i + 1]/w->efc_R[i]);
    if (tp == 7) {
      for (int j = 1; j < dim - 1; j++) w->efc_R[i + j + 1] = w->efc_R[i + 1]*fr[0]*fr[0]/(fr[j]*fr[j]);
      i += dim - 1;
    } else {
      double Rpy = 2*w->con_mu[id]*w->con_mu[id]*w->efc_R[i];
      for (int j = 0; j < 2*(dim - 1); j++) w->efc_R[i + j] = Rpy;
      i += 2*(dim - 1) - 1;
    }
  }
  for (int i = 0; i < w->nefc; i++) w->efc_D[i] = 1/w->efc_R[i];
}

/* mj_referenceConstraint (:2362) + mj_invConstraint (engine_inverse.c:169) + mj_constraintUpdate (:2387) */
static void constraint_forces(const OrcModel* m, Work* w, const double* qvel, const double* qacc, double* qfrc_constraint) {
  int nv = m->nv, ne = w->ne, nf = w->nf;
  memset(qfrc_constraint, 0, sizeof(double)*nv);
  for (int i = 0; i < w->nefc; i++) {
    w->efc_vel[i] = dotn(w->J + (size_t)i*nv, qvel, nv);
    w->efc_aref[i] = -w->efc_B[i]*w->efc_vel[i] - w->efc_K[i]*w->efc_imp[i]*(w->efc_pos[i] - w->efc_margin[i]);
  }
  for (int i = 0; i < w->nefc; i++) {
    double jar = dotn(w->J + (size_t)i*nv, qacc, nv) - w->efc_aref[i];
    w->efc_force[i] = -w->efc_D[i]*jar;
    if (i < ne) continue;
    if (i < ne + nf) {
      double fl = w->efc_floss[i];
      if (jar <= -w->efc_R[i]*fl) w->efc_force[i] = fl;
      else if (jar >= w->efc_R[i]*fl) w->efc_force[i] = -fl;
      continue;
    }
    if (w->efc_type[i] != 7) { if (jar >= 0) w->efc_force[i] = 0; continue; }
    int id = w->efc_id[i], dim = w->con_dim[id];
    double mu = w->con_mu[id], U[6], jr[6], tt = 0;
    const double* fr = w->con_friction[id];
    for (int j = 0; j < dim; j++) jr[j] = dotn(w->J + (size_t)(i + j)*nv, qacc, nv) - w->efc_aref[i + j];
    U[0] = jr[0]*mu;
    for (int j = 1; j < dim; j++) { U[j] = jr[j]*fr[j - 1]; tt += U[j]*U[j]; }
    double N = U[0], T = sqrt(tt);
    if (N >= mu*T || (T <= 0 && N >= 0)) { for (int j = 0; j < dim; j++) w->efc_force[i + j] = 0; }
    else if (mu*N + T <= 0 || (T <= 0 && N < 0)) { for (int j = 0; j < dim; j++) w->efc_force[i + j] = -w->efc_D[i + j]*jr[j]; }
    else {
      double Dm = w->efc_D[i]/(mu*mu*(1 + mu*mu)), NmT = N - mu*T;
      w->efc_force[i] = -Dm*NmT*mu;
      for (int j = 1; j < dim; j++) w->efc_force[i + j] = -w->efc_force[i]/T*U[j]*fr[j - 1];
    }
    i += dim - 1;
  }
  for (int i = 0; i < w->nefc; i++) if (w->efc_force[i]) for (int k = 0; k < nv; k++) qfrc_constraint[k] += w->J[(size_t)i*nv + k]*w->efc_force[i];
}

/* mj_rne(flg_acc=1) (engine_core_smooth.c:1969-2023) */
static void rne(const OrcModel* m, const Work* w, const double* qvel, const double* qacc, double* result) {
  int nb = m->nbody;
  double* cacc = (double*)calloc((size_t)nb*12, sizeof(double));
  double* cfrc = cacc + 6*nb;
  if (!(m->disableflags & (1 << 6))) for (int k = 0; k < 3; k++) cacc[3 + k] = -m->gravity[k];
  for (int i = 1; i < nb; i++) {
    int bda = m->body_dofadr[i], n = m->body_dofnum[i];
    double t[6] = {0}, t1[6], t2[6];
    for (int j = 0; j < n; j++) for (int k = 0; k < 6; k++) t[k] += w->cdof_dot[6*(bda + j) + k]*qvel[bda + j];
    for (int k = 0; k < 6; k++) cacc[6*i + k] = cacc[6*m->body_parentid[i] + k] + t[k];
    memset(t, 0, 48);
    for (int j = 0; j < n; j++) for (int k = 0; k < 6; k++) t[k] += w->cdof[6*(bda + j) + k]*qacc[bda + j];
    for (int k = 0; k < 6; k++) cacc[6*i + k] += t[k];
    mul_inert(cfrc + 6*i, w->cinert + 10*i, cacc + 6*i);
    mul_inert(t1, w->cinert + 10*i, w->cvel + 6*i);
    cross_force(t2, w->cvel + 6*i, t1);
    for (int k = 0; k < 6; k++) cfrc[6*i + k] += t2[k];
  }
  for (int i = nb - 1; i > 0; i--) if (m->body_parentid[i]) for (int k = 0; k < 6; k++) cfrc[6*m->body_parentid[i] + k] += cfrc[6*i + k];
  for (int i = 0; i < m->nv; i++) result[i] = dotn(w->cdof + 6*i, cfrc + 6*m->dof_bodyid[i], 6);
  free(cacc);
}

/* mj_crb (:1353-1401) and mj_factorI by in-place elimination (:1483-1511), legacy M layout */
static void inertia(const OrcModel* m, const Work* w, double* qM, double* qLD, double* diaginv) {
  int nb = m->nbody, nv = m->nv;
  double* crb = (double*)malloc(sizeof(double)*10*nb);
  memcpy(crb, w->cinert, sizeof(double)*10*nb);
  for (int i = nb - 1; i > 0; i--) if (m->body_parentid[i] > 0) for (int k = 0; k < 10; k++) crb[10*m->body_parentid[i] + k] += crb[10*i + k];
  memset(qM, 0, sizeof(double)*m->nM);
  for (int i = 0; i < nv; i++) {
    double buf[6];
    int adr = m->dof_Madr[i];
    if (m->dof_simplenum[i]) { qM[adr] = m->dof_M0[i]; continue; }   /* :1375-1385 */
    qM[adr] = m->dof_armature[i];
    mul_inert(buf, crb + 10*m->dof_bodyid[i], w->cdof + 6*i);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) qM[adr++] += dotn(w->cdof + 6*j, buf, 6);
  }
  free(crb);
  if (!qLD) return;
  /* dense copy, eliminate rows from the last to the first, read back in C order (ancestors ascending, diag last) */
  double* A = (double*)calloc((size_t)nv*nv, sizeof(double));
  for (int i = 0; i < nv; i++) { int adr = m->dof_Madr[i]; for (int j = i; j >= 0; j = m->dof_parentid[j]) A[i*nv + j] = qM[adr++]; }
  for (int k = nv - 1; k >= 0; k--) {
    double invD = 1/A[k*nv + k];
    diaginv[k] = invD;
    if (m->dof_simplenum[k]) continue;                                  /* :1498 */
    for (int i = m->dof_parentid[k]; i >= 0; i = m->dof_parentid[i]) {
      double tmp = A[k*nv + i]*invD;
      for (int j = i; j >= 0; j = m->dof_parentid[j]) A[i*nv + j] -= tmp*A[k*nv + j];
      A[k*nv + i] = tmp;
    }
  }
  int adr = 0;
  for (int i = 0; i < nv; i++) {
    int chain[4096], n = 0;
    for (int j = i; j >= 0; j = m->dof_parentid[j]) { chain[n++] = j; if (m->dof_simplenum[i]) break; }
    for (int s = n - 1; s >= 0; s--) qLD[adr++] = A[i*nv + chain[s]];   /* reduced C layout, engine_io.c:952 */
  }
  free(A);
}

/* mj_inverse for one state (engine_inverse.c:197-261); returns 0, -1 (geom pair outside the restated
 * set) or -2 (capacity) */
ORC_API int orc_inverse(const OrcModel* m, const double* qpos, const double* qvel, const double* qacc, OrcOut* out) {
  int nb = m->nbody, nv = m->nv, nj = m->njnt, ng = m->ngeom;
  Work* w = (Work*)calloc(1, sizeof(Work));
  size_t nd = (size_t)nb*(3 + 4 + 9 + 3 + 9 + 3 + 10 + 6) + (size_t)nj*6 + (size_t)ng*12 + (size_t)nv*(6 + 6 + 1) +
              (size_t)m->ntendon*(1 + nv) + 16;
  double* buf = (double*)calloc(nd, sizeof(double));
  double* p = buf;
#define TAKE(field, n) w->field = p; p += (n);
  TAKE(xpos, 3*nb) TAKE(xquat, 4*nb) TAKE(xmat, 9*nb) TAKE(xipos, 3*nb) TAKE(ximat, 9*nb) TAKE(com, 3*nb)
  TAKE(cinert, 10*nb) TAKE(cvel, 6*nb) TAKE(xanchor, 3*nj) TAKE(xaxis, 3*nj) TAKE(gpos, 3*ng) TAKE(gmat, 9*ng)
  TAKE(cdof, 6*nv) TAKE(cdof_dot, 6*nv) TAKE(qfrc_passive, nv) TAKE(ten_length, m->ntendon) TAKE(ten_J, m->ntendon*nv)
  w->J = (double*)calloc((size_t)MAXEFC*(nv ? nv : 1), sizeof(double));
  double* qfrc_constraint = (double*)calloc((size_t)nv + 1, sizeof(double));
  int rc = 0;

  kinematics(m, w, qpos);
  com_pos(m, w);
  tendons(m, w, qpos);
  if (out->qM) inertia(m, w, out->qM, out->qLD, out->qLDiagInv);
  rc = collision(m, w);
  if (!rc) rc = make_constraint(m, w, qpos);
  if (!rc) {
    impedances(m, w);
    com_vel(m, w, qvel);
    passive(m, w, qpos, qvel);
    constraint_forces(m, w, qvel, qacc, qfrc_constraint);
    rne(m, w, qvel, qacc, out->qfrc_inverse);
    for (int i = 0; i < nv; i++) out->qfrc_inverse[i] += m->dof_armature[i]*qacc[i] - w->qfrc_passive[i] - qfrc_constraint[i];
    out->counts[0] = w->ncon; out->counts[1] = w->ne; out->counts[2] = w->nf; out->counts[3] = w->nl; out->counts[4] = w->nefc;
    if (out->contact_geom) for (int c = 0; c < out->maxcon; c++) {
      out->contact_geom[2*c] = c < w->ncon ? w->con_geom[c][0] : -1;
      out->contact_geom[2*c + 1] = c < w->ncon ? w->con_geom[c][1] : -1;
    }
    if (out->efc_type) for (int r = 0; r < out->maxefc; r++) {
      out->efc_type[r] = r < w->nefc ? w->efc_type[r] : -1;
      out->efc_id[r] = r < w->nefc ? w->efc_id[r] : -1;
      out->efc_force[r] = r < w->nefc ? w->efc_force[r] : 0;
    }
  }
  free(qfrc_constraint); free(w->J); free(buf); free(w);
  return rc;
}

ORC_API int orc_sizeof_model(void) { return (int)sizeof(OrcModel); }
