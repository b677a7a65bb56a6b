// Part of the per-state mj_inverse pipeline (mjb_pipeline.h includes it inside namespace mjb, after the
// context and accessor macros; not a stand-alone header).
// mj_rnePostConstraint outputs (incl. applied wrenches), backward half of mj_rne with the final combine, qfrc_bias.
#ifndef MJB_BACKWARD_H_
#define MJB_BACKWARD_H_

// ------------------------------------------------------------------------------------------
// mj_rnePostConstraint (engine_core_smooth.c:2027-2181): cacc, cfrc_ext, cfrc_int per body, in
// the reference's frame (origin at subtree_com of the body's tree). Everything it needs already
// exists in the backward sweep: cacc from the forward sweep, the per-body constraint wrenches in
// the two carriers (contacts, connect / weld rows; xfrc_applied is zero on this path exactly as
// in an mjData fresh from mj_makeData), the subtree sums of the main loop. What is left is the
// change of origin O -> C = subtree_com[root]:  motion  lin_C = lin_O + ang x d,
//                                              force   trq_C = trq_O - d x F,   d = C - O,
// with d = sum(mass*(xipos - O)) / sum(mass) over the tree, read from cinert[6..9] (mju_inertCom).

// before the main loop (which folds the children into the carriers): per tree, d, then cacc and
// the per-body cfrc_ext; d is left in the first three rows of the ROOT's cfrc_int output, which
// the main loop overwrites last within the tree (bodies of a tree are contiguous, root first)
MJB_HD inline void post_constraint_begin(Ctx& c) {
  const mjbHdr& H = *c.H;
  const int nbody = H.nbody;
  const size_t N = (size_t)c.N;
  const int* body_parentid = MI(body_parentid);
  double* cinert = SC(cinert); double* cacc = SC(cacc);
  double* fext = SC(cfrc_ext); double* fext1 = SC(cfrc_ext1);
  for (int k = 0; k < 6; k++) {
    double a = 0;
    if (k >= 3 && !(H.disableflags & MJB_DSBL_GRAVITY)) a = -H.gravity[k - 3];
    c.out.cacc[(size_t)k*N + c.s] = a;
    c.out.cfrc_ext[(size_t)k*N + c.s] = 0;
  }
  int r = 1;
  while (r < nbody) {
    int e = r + 1;
    while (e < nbody && body_parentid[e] != 0) e++;
    double ms[4] = {0, 0, 0, 0};
    for (int b = r; b < e; b++) {
      double t[4];
      ldn(t, cinert, 10*b + 6, 4);
      for (int k = 0; k < 4; k++) ms[k] += t[k];
    }
    double d[3] = {0, 0, 0};
    if (ms[3] >= MJB_MINVAL) { d[0] = ms[0]/ms[3]; d[1] = ms[1]/ms[3]; d[2] = ms[2]/ms[3]; }
    for (int k = 0; k < 3; k++) c.out.cfrc_int[(size_t)(6*r + k)*N + c.s] = d[k];
    for (int b = r; b < e; b++) {
      double a[6], w[6] = {0, 0, 0, 0, 0, 0}, w1[6] = {0, 0, 0, 0, 0, 0}, cr[3];
      ldn(a, cacc, 6*b, 6);
      if (wmask_test(c, b, true)) ldn(w, fext, 6*b, 6);
      if (wmask_test(c, b, false)) ldn(w1, fext1, 6*b, 6);
      cross3(cr, a, d);
      for (int k = 0; k < 3; k++) a[3 + k] += cr[k];
      for (int k = 0; k < 6; k++) w[k] -= w1[k];
      cross3(cr, d, w + 3);
      for (int k = 0; k < 3; k++) w[k] -= cr[k];
      for (int k = 0; k < 6; k++) {
        c.out.cacc[(size_t)(6*b + k)*N + c.s] = a[k];
        c.out.cfrc_ext[(size_t)(6*b + k)*N + c.s] = w[k];
      }
    }
    r = e;
  }
}

// after the main loop: the weld torque convention of the reference and the world body's row
MJB_HD inline void post_constraint_end(Ctx& c) {
  const mjbHdr& H = *c.H;
  const size_t N = (size_t)c.N;
  const int* body_parentid = MI(body_parentid);
  if (H.neq && rows_enabled(H) && !(H.disableflags & MJB_DSBL_EQUALITY)) {
    const int* eq_int = MI(eq_int);
    for (int i = 0; i < H.neq; i++) {
      const int* ei = eq_int + MJB_EQ_NI*i;
      if (ei[MJB_EQI_TYPE] != 1 || !eq_enabled(c, ei, i) || ei[MJB_EQI_SKIP]) continue;
      double dT[3];
      ldn(dT, SC(weld_dt), 3*i, 3);
      for (int side = 0; side < 2; side++) {
        const int body = ei[side == 0 ? MJB_EQI_B0 : MJB_EQI_B1];
        const double sg = side == 0 ? 1.0 : -1.0;
        if (!body) continue;
        for (int k = 0; k < 3; k++) c.out.cfrc_ext[(size_t)(6*body + k)*N + c.s] += sg*dT[k];
        for (int a = body; a > 0; a = body_parentid[a]) {
          for (int k = 0; k < 3; k++) c.out.cfrc_int[(size_t)(6*a + k)*N + c.s] -= sg*dT[k];
        }
      }
    }
  }
  // the reference adds the trees' root rows into the world body's row as they are (:2178-2180)
  double s0[6] = {0, 0, 0, 0, 0, 0};
  for (int b = H.nbody - 1; b > 0; b--) {
    if (body_parentid[b] == 0) {
      for (int k = 0; k < 6; k++) s0[k] += c.out.cfrc_int[(size_t)(6*b + k)*N + c.s];
    }
  }
  for (int k = 0; k < 6; k++) c.out.cfrc_int[(size_t)k*N + c.s] = s0[k];
}

// d->xfrc_applied in the mj_rnePostConstraint outputs (engine_core_smooth.c:2039-2049, 2171-2181), as a
// pass of its own after the backward sweep (only when the caller has set per-state applied wrenches):
// X_b = the body's applied (force, torque) at xipos re-expressed about the tree's centre of mass;
//   cfrc_ext[b] += X_b,   cfrc_int[b] -= sum of X over the subtree of b,
// and the world body's row, the plain sum of the root rows, loses the roots' sums. The subtree sums
// live in the ia rows of the scratch (free after the inertia kernel).
MJB_HD inline int sensor_object(Ctx& c, int objtype, int objid, double* pos, double* quat);   // defined with the sensors
MJB_HD inline void post_xfrc(Ctx& c) {
  const mjbHdr& H = *c.H;
  const int nbody = H.nbody;
  const size_t N = (size_t)c.N;
  const int* body_parentid = MI(body_parentid); const int* rootid = MI(body_rootid);
  double* t = SC(ia);
  const double zero[6] = {0, 0, 0, 0, 0, 0};
  for (int b = 0; b < nbody; b++) stn(t, 6*b, zero, 6);
  double world[6] = {0, 0, 0, 0, 0, 0};
  int tree = -1;
  double com[3] = {0, 0, 0};
  for (int b = nbody - 1; b > 0; b--) {
    double x[6], X[6] = {0, 0, 0, 0, 0, 0}, acc[6];
    bool any = false;
    for (int k = 0; k < 6; k++) { x[k] = c.out.xfrc_applied[(size_t)(6*b + k)*N + c.s]; any = any || x[k] != 0; }
    if (any) {
      const int r = rootid[b];
      if (r != tree) {
        // subtree_com of the tree's root = O + sum m (xipos - O) / sum m, from cinert[6..9]
        double ms[4] = {0, 0, 0, 0}, o[3];
        int e = r;
        do {
          double q[4];
          ldn(q, SC(cinert), 10*e + 6, 4);
          for (int k = 0; k < 4; k++) ms[k] += q[k];
          e++;
        } while (e < nbody && body_parentid[e] != 0);
        ldn(o, SC(origin), 3*r, 3);
        for (int k = 0; k < 3; k++) com[k] = o[k] + (ms[3] >= MJB_MINVAL ? ms[k]/ms[3] : 0.0);
        tree = r;
      }
      double xi[3], q4[4], cr[3];
      sensor_object(c, MJB_OBJ_BODY, b, xi, q4);
      const double dif[3] = {com[0] - xi[0], com[1] - xi[1], com[2] - xi[2]};
      cross3(cr, dif, x);                                    // (newpos - oldpos) x force
      for (int k = 0; k < 3; k++) { X[k] = x[3 + k] - cr[k]; X[3 + k] = x[k]; }
      for (int k = 0; k < 6; k++) c.out.cfrc_ext[(size_t)(6*b + k)*N + c.s] += X[k];
    }
    ldn(acc, t, 6*b, 6);
    bool nz = any;
    for (int k = 0; k < 6; k++) { acc[k] += X[k]; nz = nz || acc[k] != 0; }
    if (!nz) continue;
    for (int k = 0; k < 6; k++) c.out.cfrc_int[(size_t)(6*b + k)*N + c.s] -= acc[k];
    const int p = body_parentid[b];
    if (p) {
      double pa[6];
      ldn(pa, t, 6*p, 6);
      for (int k = 0; k < 6; k++) pa[k] += acc[k];
      stn(t, 6*p, pa, 6);
    } else {
      for (int k = 0; k < 6; k++) world[k] += acc[k];
    }
  }
  for (int k = 0; k < 6; k++) c.out.cfrc_int[(size_t)k*N + c.s] -= world[k];
}

// ------------------------------------------------------------------------------------------
// backward half of mj_rne(flg_acc=1) (engine_core_smooth.c:2008-2020) fused with the last loop of
// mj_inverseSkip (engine_inverse.c:249-252). Constraint wrenches are accumulated up the tree
// separately and projected with the same cdof, which is J'*efc_force for the point constraints.
template <bool kGravcomp>
MJB_HD inline void rne_and_output(Ctx& c) {
  const mjbHdr& H = *c.H;
  const int nbody = H.nbody;
  double* cfrc = SC(cfrc); double* fext = SC(cfrc_ext); double* fext1 = SC(cfrc_ext1);
  double* cdof = SC(cdof);
  const int* body_parentid = MI(body_parentid);
  const int* body_dofadr = MI(body_dofadr);
  const int* body_dofnum = MI(body_dofnum);
  const double* armature = MD(dof_armature);
  double* qc = SC(qfrc_c); double* qp = SC(qfrc_passive);
  const size_t N = (size_t)c.N;

  // Leaves-to-root: when body b is visited all its children have already pushed their sums into
  // it, so its inertial force f and net constraint wrench w ('+' minus '-' side) are final: push
  // them to the parent and project them on the body's own dofs right away. Every body does its
  // loads first and its stores last (one memory round trip per body).
  // gravity compensation (only models that have it): a third wrench carrier, projected into
  // qfrc_passive except on joints whose gravcomp is routed through actuators (engine_passive.c:459-489)
  constexpr bool gcomp = kGravcomp;    // compile-time: the carrier costs registers only where used
  double* fgc = SC(cfrc_gc);
  const int* dof_jntid = MI(dof_jntid);
  const int* jnt_actgravcomp = MI(jnt_actgravcomp);
  const bool post = c.out.cfrc_int != nullptr;
  if (post) post_constraint_begin(c);
  int post_tree = -1;
  double post_d[3] = {0, 0, 0};
  int carry_for = -1;
  bool carry_w = false;
  double cf[6], cw[6], cg[6] = {0, 0, 0, 0, 0, 0};
  auto rne_body = [&](const int b) MJB_BODY_LAMBDA {
    const int p = body_parentid[b];
    const bool push = p && b != p + 1;       // child p+1 hands over in registers (depth-first order)
    double f[6], w[6] = {0, 0, 0, 0, 0, 0}, w1[6] = {0, 0, 0, 0, 0, 0}, pf[6];
    double pw1[6] = {0, 0, 0, 0, 0, 0}, g[6] = {0, 0, 0, 0, 0, 0}, pg[6];
    ldn(f, cfrc, 6*b, 6);
    // constraint wrenches: only rows some constraint has written for this state (wrench masks);
    // a contact-free state reads none of them
    const bool has_w = wmask_test(c, b, true), has_w1 = wmask_test(c, b, false);
    if (has_w) ldn_ro(w, fext, 6*b, 6);
    if (has_w1) ldn(w1, fext1, 6*b, 6);
    if (gcomp) ldn(g, fgc, 6*b, 6);
    bool carried_w = false;
    if (push) {
      ldn(pf, cfrc, 6*p, 6);
      if (wmask_test(c, p, false)) ldn(pw1, fext1, 6*p, 6);
      if (gcomp) ldn(pg, fgc, 6*p, 6);
    }
    for (int k = 0; k < 6; k++) w[k] -= w1[k];
    if (carry_for == b) {
      for (int k = 0; k < 6; k++) { f[k] += cf[k]; w[k] += cw[k]; g[k] += cg[k]; }
      carried_w = carry_w;
    }
    const bool any_w = has_w || has_w1 || carried_w;    // this subtree carries a constraint wrench
    if (post) {
      // cfrc_int = sum over the subtree of (inertial force - external force), re-expressed about
      // the tree's centre of mass; the shift was left in the root's row by post_constraint_begin
      const int r = MI(body_rootid)[b];
      if (r != post_tree) {
        for (int k = 0; k < 3; k++) post_d[k] = c.out.cfrc_int[(size_t)(6*r + k)*N + c.s];
        post_tree = r;
      }
      double o[6], cr[3];
      for (int k = 0; k < 6; k++) o[k] = f[k] - w[k];
      cross3(cr, post_d, o + 3);
      for (int k = 0; k < 3; k++) o[k] -= cr[k];
      for (int k = 0; k < 6; k++) c.out.cfrc_int[(size_t)(6*b + k)*N + c.s] = o[k];
    }
    const int d0 = body_dofadr[b], dn = body_dofnum[b];
    MJB_UNROLL
    for (int i = d0; i < d0 + dn; i++) {
      double cd[6];
      ldn_ro(cd, cdof, 6*i, 6);
      const double qfrc_constraint = AT(qc, i) + dot6f(cd, w);
      double passive_i = AT(qp, i);
      if (gcomp && !(H.has_gravcomp && jnt_actgravcomp[dof_jntid[i]])) passive_i += dot6f(cd, g);
      double res = dot6f(cd, f);
      res += armature[i]*QACC(i) - passive_i - qfrc_constraint;
      c.out.qfrc_inverse[(size_t)i*N + c.s] = res;
      if (c.out.qfrc_constraint) c.out.qfrc_constraint[(size_t)i*N + c.s] = qfrc_constraint;
      if (c.out.qfrc_passive) c.out.qfrc_passive[(size_t)i*N + c.s] = passive_i;
    }
    if (push) {
      // parent's net = own '+' - own '-' + children's nets: children are folded into its '-' side
      for (int k = 0; k < 6; k++) { pf[k] += f[k]; pw1[k] -= w[k]; }
      stn(cfrc, 6*p, pf, 6);
      if (any_w) { stn(fext1, 6*p, pw1, 6); wmask_test_and_set(c, p, false, false); }
      if (gcomp) {
        for (int k = 0; k < 6; k++) pg[k] += g[k];
        stn(fgc, 6*p, pg, 6);
      }
    } else if (p) {
      for (int k = 0; k < 6; k++) { cf[k] = f[k]; cw[k] = w[k]; cg[k] = g[k]; }
      carry_for = p;
      carry_w = any_w;
    }
  };
  MJB_BODY_LOOP_DOWN(rne_body, 1, nbody, 1, MJB_SPEC_NBODY);
  if (post) post_constraint_end(c);
}

// ------------------------------------------------------------------------------------------
// qfrc_bias = mj_rne(m, d, 0, qfrc_bias) of mj_fwdVelocity (engine_forward.c:228,
// engine_core_smooth.c:1969-2023): Coriolis, centrifugal and gravitational forces. The forward
// sweep carries the full acceleration A = A_bias + sum cdof*qacc and the qacc part alone
// (cacc_lin), so the bias acceleration is their difference and the body force
// cinert*A_bias + cvel x* (cinert*cvel) is accumulated up the tree in a row block that is free
// after the inertia kernel (ia) and projected on the dofs. Runs after the backward sweep.
MJB_HD inline void bias_forces(Ctx& c) {
  const mjbHdr& H = *c.H;
  const size_t N = (size_t)c.N;
  const int* body_parentid = MI(body_parentid);
  const int* dof_bodyid = MI(dof_bodyid);
  double* tmp = SC(ia);
  for (int b = 1; b < H.nbody; b++) {
    double ci[10], a[6], al[6], v[6], f[6], u1[6], u2[6];
    ldn(ci, SC(cinert), 10*b, 10); ldn(a, SC(cacc), 6*b, 6); ldn(al, SC(cacc_lin), 6*b, 6);
    ldn(v, SC(cvel), 6*b, 6);
    for (int k = 0; k < 6; k++) a[k] -= al[k];
    mulInertVec(f, ci, a);
    mulInertVec(u1, ci, v);
    crossForce(u2, v, u1);
    for (int k = 0; k < 6; k++) f[k] += u2[k];
    stn(tmp, 6*b, f, 6);
  }
  for (int b = H.nbody - 1; b > 0; b--) {
    const int p = body_parentid[b];
    if (!p) continue;
    double f[6], pf[6];
    ldn(f, tmp, 6*b, 6); ldn(pf, tmp, 6*p, 6);
    for (int k = 0; k < 6; k++) pf[k] += f[k];
    stn(tmp, 6*p, pf, 6);
  }
  for (int i = 0; i < H.nv; i++) {
    double cd[6], f[6];
    ldn(cd, SC(cdof), 6*i, 6); ldn(f, tmp, 6*dof_bodyid[i], 6);
    c.out.qfrc_bias[(size_t)i*N + c.s] = dot6(cd, f);
  }
}


#endif  // MJB_BACKWARD_H_
