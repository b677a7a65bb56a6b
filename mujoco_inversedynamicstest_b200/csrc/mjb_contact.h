// Part of the per-state mj_inverse pipeline (mjb_pipeline.h includes it inside namespace mjb, after the
// context and accessor macros; not a stand-alone header).
// Contact rows: relative motion and wrenches at a contact point, pyramidal and elliptic rows (mj_instantiateContact .. mj_constraintUpdate).
#ifndef MJB_CONTACT_H_
#define MJB_CONTACT_H_

// ------------------------------------------------------------------------------------------
// contacts

struct Con { double dist; double pos[3]; double frame[9]; };
#define MJB_MAXCON_PAIR 24   // most contacts one geom pair can yield before clean-up (box-box)

// relative spatial motion of body b2 minus body b1 at point p, from a per-body carrier array
// (cvel or cacc_lin): lin = (lin2 + ang2 x (p - O2)) - (lin1 + ang1 x (p - O1)), ang = ang2 - ang1
MJB_HD inline void rel_motion(Ctx& c, const double* carrier, int b1, int b2, const double* p,
                              double* lin, double* ang) {
  const int* rootid = MI(body_rootid);
  double* com = SC(origin);
  double v1[6], v2[6], o1[3], o2[3], r[3], cr1[3], cr2[3];
  ldn(v1, carrier, 6*b1, 6); ldn(v2, carrier, 6*b2, 6);
  ldn(o1, com, 3*rootid[b1], 3); ldn(o2, com, 3*rootid[b2], 3);
  r[0] = p[0] - o1[0]; r[1] = p[1] - o1[1]; r[2] = p[2] - o1[2];
  cross3(cr1, v1, r);
  r[0] = p[0] - o2[0]; r[1] = p[1] - o2[1]; r[2] = p[2] - o2[2];
  cross3(cr2, v2, r);
  for (int k = 0; k < 3; k++) {
    lin[k] = (v2[3+k] + cr2[k]) - (v1[3+k] + cr1[k]);
    ang[k] = v2[k] - v1[k];
  }
}

// the same for the two bodies of a contact, velocity and acceleration carriers at once, from their
// carrier records (MJB_SC_crec): one 128-byte line per moving body
MJB_HD inline void contact_rel_motion(Ctx& c, int b1, int b2, const double* p, double* vlin, double* vang,
                                      double* alin, double* aang) {
  const int* body_static = MI(body_static);
  double r1[16], r2[16], r[3], c1[3], c2[3];
  load_crec(c, b1, body_static[b1] != 0, r1);
  load_crec(c, b2, body_static[b2] != 0, r2);
  r[0] = p[0] - r1[12]; r[1] = p[1] - r1[13]; r[2] = p[2] - r1[14];
  double a1[3], a2[3];
  cross3(c1, r1, r);
  cross3(a1, r1 + 6, r);
  r[0] = p[0] - r2[12]; r[1] = p[1] - r2[13]; r[2] = p[2] - r2[14];
  cross3(c2, r2, r);
  cross3(a2, r2 + 6, r);
  for (int k = 0; k < 3; k++) {
    vlin[k] = (r2[3+k] + c2[k]) - (r1[3+k] + c1[k]);
    vang[k] = r2[k] - r1[k];
    alin[k] = (r2[9+k] + a2[k]) - (r1[9+k] + a1[k]);
    aang[k] = r2[6+k] - r1[6+k];
  }
}

// add the wrench (torque T about point p, force F at p) to body b2 and its opposite to body b1
MJB_HD inline void apply_wrench(Ctx& c, int b1, int b2, const double* p, const double* F,
                                const double* T) {
  add_wrench(c, b2, p, F, T, true);
  add_wrench(c, b1, p, F, T, false);
}

// rows a contact will occupy and its exclude flag (mj_setContact :1387-1413 exclude-in-gap rule,
// mj_instantiateContact :1072-1076 NV == 0 rule, :1084-1126 row counts)
MJB_HD inline int contact_row_count(Ctx& c, int ci, double dist, int* exclude) {
  const mjbHdr& H = *c.H;
  const int* cint = MI(cand_int) + MJB_CAND_NI*ci;
  const double includemargin = MD(cand_num)[MJB_CAND_NN*ci + MJB_CN_INCLUDEMARGIN];
  *exclude = (dist >= includemargin) ? 1 : 0;
  if (*exclude || (H.disableflags & MJB_DSBL_CONSTRAINT) || H.nv == 0) return 0;
  if (cint[MJB_CI_FLAGS] & 1) { *exclude = 3; return 0; }     // no dof on either side (NV == 0)
  const int dim = cint[MJB_CI_DIM];
  return dim == 1 ? 1 : (H.cone == 0 ? 2*(dim - 1) : dim);
}

// Contact k of the state bound to c (frame already completed by mju_makeFrame): writes the contact
// outputs, evaluates its rows starting at row efc_address (< 0: none) and returns J'f as the world
// force F and torque T3 at con.pos (+ on body 2, - on body 1). c.nefc is left after the last row.
MJB_HD inline void contact_rows(Ctx& c, int ci, const Con& con, int k, int exclude, int efc_address,
                                double* F, double* T3) {
  const mjbHdr& H = *c.H;
  const int* cint = MI(cand_int) + MJB_CAND_NI*ci;
  const double* cn = MD(cand_num) + MJB_CAND_NN*ci;
  const int dim = cint[MJB_CI_DIM];
  const int b1 = cint[MJB_CI_B1], b2 = cint[MJB_CI_B2];
  const double includemargin = cn[MJB_CN_INCLUDEMARGIN];
  F[0] = F[1] = F[2] = 0; T3[0] = T3[1] = T3[2] = 0;

  if (c.out.contact_geom) {
    if (k < c.nconmax) {
      const size_t N = (size_t)c.N;
      int* cg = c.out.contact_geom + c.s;
      int* cinfo = c.out.contact_info + c.s;
      double* cnum = c.out.contact_num + c.s;
      cg[(size_t)(2*k)*N] = cint[MJB_CI_G1];
      cg[(size_t)(2*k + 1)*N] = cint[MJB_CI_G2];
      cinfo[(size_t)(3*k)*N] = dim;
      cinfo[(size_t)(3*k + 1)*N] = exclude;
      cinfo[(size_t)(3*k + 2)*N] = efc_address;
      cnum[(size_t)(13*k)*N] = con.dist;
      for (int j = 0; j < 3; j++) cnum[(size_t)(13*k + 1 + j)*N] = con.pos[j];
      for (int j = 0; j < 9; j++) cnum[(size_t)(13*k + 4 + j)*N] = con.frame[j];
    } else {
      c.status |= kStatusContactFull;
    }
  }
  if (efc_address < 0) return;
  int row = efc_address;

  const double* sp = cn + MJB_CN_SP;
  const double* friction = cn + MJB_CN_FRICTION;
  const double tran = cn[MJB_CN_DA_TRAN], rot = cn[MJB_CN_DA_ROT];
  const double imp = impedance(sp, con.dist, includemargin);
  const double K = sp[MJB_SP_K], B = sp[MJB_SP_B];
  const double pen = con.dist - includemargin;

  // relative motion in the contact frame: index 0..2 translation, 3..5 rotation
  double lin[3], ang[3], alin[3], aang[3], vel[6], acc[6];
  contact_rel_motion(c, b1, b2, con.pos, lin, ang, alin, aang);
  for (int j = 0; j < 3; j++) {
    vel[j] = dot3(con.frame + 3*j, lin);
    vel[3 + j] = dot3(con.frame + 3*j, ang);
  }
  for (int j = 0; j < 3; j++) {
    acc[j] = dot3(con.frame + 3*j, alin);
    acc[3 + j] = dot3(con.frame + 3*j, aang);
  }

  // force coefficients along the 6 contact-frame directions (J' f)
  double fc[6] = {0, 0, 0, 0, 0, 0};

  if (dim == 1) {
    const double R = fmax(MJB_MINVAL, (1 - imp)*tran/imp);
    const double D = 1/R;
    const double aref = -B*vel[0] - K*imp*pen;
    const double jar = acc[0] - aref;
    double force = -D*jar;
    int state = MJB_STATE_QUADRATIC;
    if (jar >= 0) { force = 0; state = MJB_STATE_SATISFIED; }
    emit_row(c, row++, MJB_CNSTR_CONTACT_FRICTIONLESS, k, con.dist, includemargin, D, R, vel[0], aref,
             force, state, imp);
    fc[0] = force;
  } else if (H.cone == 0) {
    // pyramidal: R of all 2(dim-1) rows = Rpy (engine_core_constraint.c:1557-1597)
    const double dA0 = tran + friction[0]*friction[0]*tran;
    const double R0 = fmax(MJB_MINVAL, (1 - imp)*dA0/imp);
    const double R1 = R0/fmax(MJB_MINVAL, H.impratio);
    const double mu = friction[0]*sqrt(R1/R0);
    const double Rpy = 2*mu*mu*R0;
    const double D = 1/Rpy;
    for (int j = 1; j < dim; j++) {
      const double fr = friction[j - 1];
      for (int sgn = 1; sgn >= -1; sgn -= 2) {
        const double v = vel[0] + sgn*fr*vel[j];
        const double a = acc[0] + sgn*fr*acc[j];
        const double aref = -B*v - K*imp*pen;
        const double jar = a - aref;
        double force = -D*jar;
        int state = MJB_STATE_QUADRATIC;
        if (jar >= 0) { force = 0; state = MJB_STATE_SATISFIED; }
        emit_row(c, row++, MJB_CNSTR_CONTACT_PYRAMIDAL, k, con.dist, includemargin, D, Rpy, v, aref, force,
                 state, imp);
        fc[0] += force;
        fc[j] += sgn*fr*force;
      }
    }
  } else {
    // elliptic
    double R[6], jar[6], aref[6], force[6];
    const double Bf = cn[MJB_CN_BFRIC];
    R[0] = fmax(MJB_MINVAL, (1 - imp)*tran/imp);
    R[1] = R[0]/fmax(MJB_MINVAL, H.impratio);
    const double mu = friction[0]*sqrt(R[1]/R[0]);
    for (int j = 1; j < dim - 1; j++) {
      R[j + 1] = R[1]*friction[0]*friction[0]/(friction[j]*friction[j]);
    }
    aref[0] = -B*vel[0] - K*imp*pen;
    jar[0] = acc[0] - aref[0];
    for (int j = 1; j < dim; j++) {
      aref[j] = -Bf*vel[j];   // K = 0, pos = margin = 0 on friction rows
      jar[j] = acc[j] - aref[j];
    }
    for (int j = 0; j < dim; j++) force[j] = -(1/R[j])*jar[j];

    // mj_constraintUpdate elliptic branch (:2459-2540)
    double U[6];
    U[0] = jar[0]*mu;
    double tt = 0;
    for (int j = 1; j < dim; j++) { U[j] = jar[j]*friction[j - 1]; tt += U[j]*U[j]; }
    const double Nn = U[0];
    const double T = sqrt(tt);
    int state;
    if (Nn >= mu*T || (T <= 0 && Nn >= 0)) {
      for (int j = 0; j < dim; j++) force[j] = 0;
      state = MJB_STATE_SATISFIED;
    } else if (mu*Nn + T <= 0 || (T <= 0 && Nn < 0)) {
      state = MJB_STATE_QUADRATIC;
    } else {
      const double Dm = (1/R[0]) / (mu*mu*(1 + mu*mu));
      const double NmT = Nn - mu*T;
      force[0] = -Dm*NmT*mu;
      for (int j = 1; j < dim; j++) force[j] = -force[0]/T*U[j]*friction[j - 1];
      state = MJB_STATE_CONE;
    }
    for (int j = 0; j < dim; j++) {
      emit_row(c, row++, MJB_CNSTR_CONTACT_ELLIPTIC, k, j == 0 ? con.dist : 0.0,
               j == 0 ? includemargin : 0.0, 1/R[j], R[j], vel[j], aref[j], force[j], state, imp);
      fc[j] = force[j];
    }
  }

  // J' f : world-frame force and torque at the contact point
  for (int a = 0; a < 3; a++) {
    F[a] = con.frame[a]*fc[0] + con.frame[3 + a]*fc[1] + con.frame[6 + a]*fc[2];
    T3[a] = con.frame[a]*fc[3] + con.frame[3 + a]*fc[4] + con.frame[6 + a]*fc[5];
  }
  c.nefc = row;
}

// One detected contact handled entirely by the thread that owns the state: mj_setContact
// (engine_collision_driver.c:1387), mj_instantiateContact (engine_core_constraint.c:964-1131),
// mj_diagApprox (:1245-1306), mj_makeImpedance (:1494-1608), mj_referenceConstraint,
// mj_invConstraint and the contact part of mj_constraintUpdate (:2446-2540), then J'*force as
// body wrenches. (The warp-pooled contact kernel calls the pieces separately.)
MJB_HD inline void process_contact(Ctx& c, int ci, Con& con) {
  const int* cint = MI(cand_int) + MJB_CAND_NI*ci;
  int exclude;
  const int rows = contact_row_count(c, ci, con.dist, &exclude);
  const int k = c.ncon++;
  double F[3], T3[3];
  contact_rows(c, ci, con, k, exclude, rows ? c.nefc : -1, F, T3);
  if (rows) apply_wrench(c, cint[MJB_CI_B1], cint[MJB_CI_B2], con.pos, F, T3);
}


#endif  // MJB_CONTACT_H_
