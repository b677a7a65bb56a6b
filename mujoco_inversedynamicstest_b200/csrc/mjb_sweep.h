// Part of the per-state mj_inverse pipeline (mjb_pipeline.h includes it inside namespace mjb, after the
// context and accessor macros; not a stand-alone header).
// The fused forward sweep (mj_kinematics, mj_comPos, mj_comVel, forward half of mj_rne, per-dof rows, geom frames, carrier records).
#ifndef MJB_SWEEP_H_
#define MJB_SWEEP_H_

// ------------------------------------------------------------------------------------------
// Forward sweep: ONE root-to-leaves pass that does, per body, the work the reference spreads over
//   mj_kinematics  (engine_core_smooth.c:38-178, mj_local2Global engine_support.c:1565)
//   mj_comPos      (:183-270; cinert via mju_inertCom, cdof via mju_dofCom)
//   mj_comVel      (:1833-1896)
//   mj_rne forward (:1969-2005, flg_acc = 1)
// so that a body's pose, joint axes, velocity and acceleration never leave registers between those
// stages; only what later phases read is written to the per-state scratch.
//
// Frame of the spatial quantities. The reference expresses cdof/cvel/cacc/cinert/cfrc about the
// centre of mass of the kinematic tree (subtree_com[body_rootid]), which is known only after a
// full kinematics pass. Spatial algebra holds about ANY fixed world point, and qfrc_inverse, qM,
// qLD, J*v, J'*f are independent of it, so the sweep uses the tree origin
//     O_tree = position of the tree's root body before its joints act
//            = qpos[0:3] of a free root, body_pos of a jointed or welded root
// which is known when the root is entered (|x - O| stays of the order of the tree's size, like the
// reference's com-based offsets). Bodies of one tree are contiguous in the body order, so O is
// carried in registers; it is also stored per root body for the constraint phases.
//
// Carry. Bodies are in depth-first order, so a body's parent is very often the body just
// processed: its pose/velocity/acceleration are then still in registers (P, Q, V, A, AL) and are
// read from scratch only when the parent is an earlier body (warp-uniform test).
//
//   cvel      spatial velocity about O                              (contact rows, cfrc)
//   cacc_lin  sum of cdof*qacc along the dof chain = carrier of J*qacc for point constraints
//   cacc      rne acceleration incl. -gravity and the cdof_dot*qvel bias (children only)
//   cfrc      cinert*cacc + cvel x* (cinert*cvel)                   (backward pass)
// cdof_dot = cvel x cdof lives only in registers (the reference stores it for its second sweep).

// L1 prefetch of the input rows (qpos/qvel/qacc) the joints of body b will read, issued one body
// ahead so that the HBM latency of the state's inputs overlaps the current body's arithmetic
MJB_DI void prefetch_line(const double* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
MJB_HD inline void prefetch_body_inputs(Ctx& c, int b) {
  const int jntadr = MI(body_jntadr)[b], jntnum = MI(body_jntnum)[b];
  const int* jnt_qposadr = MI(jnt_qposadr); const int* jnt_dofadr = MI(jnt_dofadr);
  const int* jnt_dofnum = MI(jnt_dofnum_tab);
  MJB_UNROLL
  for (int j = jntadr; j < jntadr + jntnum; j++) {
    const int nd = jnt_dofnum[j], qa = jnt_qposadr[j], da = jnt_dofadr[j];
    const int nq = nd == 6 ? 7 : (nd == 3 ? 4 : 1);
    for (int k = 0; k < nq; k++) prefetch_line(&QPOS(qa + k));
    for (int k = 0; k < nd; k++) { prefetch_line(&QVEL(da + k)); prefetch_line(&QACC(da + k)); }
  }
}

// pose of the geoms of body b from the body's frames held in registers (mj_local2Global)
MJB_HD inline void body_geoms(Ctx& c, int b, const double* pos, const double* quat, const double* mat,
                              const double* ip, const double* im) {
  const int* body_geomadr = MI(body_geomadr);
  const int* body_geomnum = MI(body_geomnum);
  const int* geom_sameframe = MI(geom_sameframe);
  const double* geom_pos = MD(geom_pos); const double* geom_quat = MD(geom_quat);
  double* gxmat = SC(geom_xmat);
  // what a later phase reads of this geom (upload, geom_store): nothing when it is in no candidate
  // pair; position + z axis for plane / sphere / capsule pairs; the full frame for the other
  // narrow-phase functions and for tendon wrapping. The debug dump stores everything.
  const int* geom_store = MI(geom_store);
  const bool dump = c.out.scratch_dump != nullptr;
  const int g0 = body_geomadr[b], gn = body_geomnum[b];
  MJB_UNROLL
  for (int g = g0; g < g0 + gn; g++) {
    int store = dump ? 3 : geom_store[g];
    if (store & 4) store = (c.out.actuator_length || c.out.sensordata) ? 3 : (store & 3);
    if (!store) continue;
    const int sf = geom_sameframe[g];
    double gp[3], gm[9];
    if (sf == MJB_SAMEFRAME_BODY) {
      gp[0] = pos[0]; gp[1] = pos[1]; gp[2] = pos[2];
    } else if (sf == MJB_SAMEFRAME_INERTIA) {
      gp[0] = ip[0]; gp[1] = ip[1]; gp[2] = ip[2];
    } else {
      mulMatVec3(gp, mat, geom_pos + 3*g);
      gp[0] += pos[0]; gp[1] += pos[1]; gp[2] += pos[2];
    }
    if (sf == MJB_SAMEFRAME_NONE) {
      double tq[4];
      mulQuat(tq, quat, geom_quat + 4*g);
      quat2Mat(gm, tq);
    } else if (sf == MJB_SAMEFRAME_BODY || sf == MJB_SAMEFRAME_BODYROT) {
      for (int k = 0; k < 9; k++) gm[k] = mat[k];
    } else {
      for (int k = 0; k < 9; k++) gm[k] = im[k];
    }
    st_rec4(geom_vec(c, MJB_SC_geom_xpos, g), gp[0], gp[1], gp[2], 0.0);
    st_rec4(geom_vec(c, MJB_SC_geom_zaxis, g), gm[2], gm[5], gm[8], 0.0);
    if (store & 2) sts(gxmat, 9*g, gm, 9);
  }
}

// ------------------------------------------------------------------------------------------
// mj_fluid (engine_passive.c:403-431): forces of the surrounding medium (opt.density, opt.viscosity,
// opt.wind) on every body with mass, either on the body's equivalent inertia box
// (mj_inertiaBoxFluidModel :527-583) or, when one of its geoms asks for it, on the ellipsoids that
// approximate its geoms (mj_ellipsoidFluidModel :588-646, mj_addedMassForces :650-690,
// mj_viscousForces :705-789). The reference applies each force with mj_applyFT into qfrc_fluid; here
// it is a wrench about the tree origin on the passive-wrench carrier of the forward sweep, which the
// backward sweep projects on the dofs of the body's chain (the same J'f without the Jacobian).
// The force laws are functions of a local velocity and model constants: out-of-line leaf functions.

// local 6D velocity [ang, lin] of the body at world point p in the frame with axes R (columns), relative
// to the wind: mj_objectVelocity(flg_local) followed by the rotated wind (engine_passive.c:539-549)
MJB_DI void fluid_local_velocity(double* lvel, const double* V, const double* O, const double* p,
                                 const double* R, const double* wind) {
  const double dif[3] = {p[0] - O[0], p[1] - O[1], p[2] - O[2]};
  double cr[3];
  cross3(cr, dif, V);
  const double lin[3] = {V[3] - cr[0], V[4] - cr[1], V[5] - cr[2]};
  for (int k = 0; k < 3; k++) {            // mju_mulMatTVec3
    lvel[k] = R[k]*V[0] + R[3 + k]*V[1] + R[6 + k]*V[2];
    lvel[3 + k] = R[k]*lin[0] + R[3 + k]*lin[1] + R[6 + k]*lin[2];
    lvel[3 + k] -= R[k]*wind[0] + R[3 + k]*wind[1] + R[6 + k]*wind[2];
  }
}

// inertia-box model: Stokes drag of the equivalent sphere plus quadratic drag of the box faces (:551-575)
MJB_COLD inline void fluid_box_force(double* lfrc, const double* lvel, const double* box, double density,
                                     double viscosity) {
  for (int k = 0; k < 6; k++) lfrc[k] = 0;
  if (viscosity > 0) {
    const double diam = (box[0] + box[1] + box[2])/3.0;
    const double sa = -MJB_PI*diam*diam*diam*viscosity;
    const double sl = -3.0*MJB_PI*diam*viscosity;
    for (int k = 0; k < 3; k++) { lfrc[k] = lvel[k]*sa; lfrc[3 + k] = lvel[3 + k]*sl; }
  }
  if (density > 0) {
    lfrc[3] -= 0.5*density*box[1]*box[2]*fabs(lvel[3])*lvel[3];
    lfrc[4] -= 0.5*density*box[0]*box[2]*fabs(lvel[4])*lvel[4];
    lfrc[5] -= 0.5*density*box[0]*box[1]*fabs(lvel[5])*lvel[5];
    lfrc[0] -= density*box[0]*(box[1]*box[1]*box[1]*box[1] + box[2]*box[2]*box[2]*box[2])*fabs(lvel[0])*lvel[0]/64.0;
    lfrc[1] -= density*box[1]*(box[0]*box[0]*box[0]*box[0] + box[2]*box[2]*box[2]*box[2])*fabs(lvel[1])*lvel[1]/64.0;
    lfrc[2] -= density*box[2]*(box[0]*box[0]*box[0]*box[0] + box[1]*box[1]*box[1]*box[1])*fabs(lvel[2])*lvel[2]/64.0;
  }
}

MJB_DI double fluid_sq(double x) { return x*x; }
MJB_DI double fluid_p4(double x) { return (x*x)*(x*x); }

// ellipsoid model of one geom: fg = its MJB_FLUID_NG record (interaction, blunt / slender / angular drag,
// Kutta and Magnus lift coefficients, virtual mass[3] and inertia[3], semi-axes[3])
MJB_COLD inline void fluid_ellipsoid_force(double* lfrc, const double* lvel, const double* fg, double density,
                                           double viscosity) {
  const double blunt = fg[1], slender = fg[2], angdrag = fg[3], kutta = fg[4], magnus = fg[5];
  const double* vmass = fg + 6; const double* vinertia = fg + 9; const double* size = fg + 12;
  const double lin[3] = {lvel[3], lvel[4], lvel[5]};
  const double ang[3] = {lvel[0], lvel[1], lvel[2]};
  for (int k = 0; k < 6; k++) lfrc[k] = 0;

  // momentum of the fluid that moves with the body (added mass; the acceleration terms are off in the reference)
  {
    const double plin[3] = {density*vmass[0]*lin[0], density*vmass[1]*lin[1], density*vmass[2]*lin[2]};
    const double pang[3] = {density*vinertia[0]*ang[0], density*vinertia[1]*ang[1], density*vinertia[2]*ang[2]};
    double f[3], t1[3], t2[3];
    cross3(f, plin, ang); cross3(t1, plin, lin); cross3(t2, pang, ang);
    for (int k = 0; k < 3; k++) { lfrc[k] += t1[k]; lfrc[k] += t2[k]; lfrc[3 + k] += f[k]; }
  }

  // lift and drag
  const double volume = 4.0/3.0 * MJB_PI * size[0] * size[1] * size[2];
  const double m01 = size[0] > size[1] ? size[0] : size[1], n01 = size[0] < size[1] ? size[0] : size[1];
  const double d_max = m01 > size[2] ? m01 : size[2];
  const double d_min = n01 < size[2] ? n01 : size[2];
  const double d_mid = size[0] + size[1] + size[2] - d_max - d_min;
  const double A_max = MJB_PI * d_max * d_mid;

  double fmag[3];
  cross3(fmag, ang, lin);
  for (int k = 0; k < 3; k++) fmag[k] *= magnus * density * volume;

  // projection of the ellipsoid along the velocity: area and the cosine to its (unnormalised) normal
  const double proj_denom = fluid_p4(size[1] * size[2]) * fluid_sq(lin[0]) +
                            fluid_p4(size[2] * size[0]) * fluid_sq(lin[1]) +
                            fluid_p4(size[0] * size[1]) * fluid_sq(lin[2]);
  const double proj_num = fluid_sq(size[1] * size[2] * lin[0]) +
                          fluid_sq(size[2] * size[0] * lin[1]) +
                          fluid_sq(size[0] * size[1] * lin[2]);
  const double A_proj = MJB_PI * sqrt(proj_denom / (proj_num > MJB_MINVAL ? proj_num : MJB_MINVAL));
  const double nrm[3] = {fluid_sq(size[1] * size[2]) * lin[0], fluid_sq(size[2] * size[0]) * lin[1],
                         fluid_sq(size[0] * size[1]) * lin[2]};
  const double speed = sqrt(lin[0]*lin[0] + lin[1]*lin[1] + lin[2]*lin[2]);
  const double cden = speed * proj_denom;
  const double cos_alpha = proj_num / (cden > MJB_MINVAL ? cden : MJB_MINVAL);
  double circ[3], fkut[3];
  cross3(circ, nrm, lin);
  for (int k = 0; k < 3; k++) circ[k] *= kutta * density * cos_alpha * A_proj;
  cross3(fkut, circ, lin);

  // Stokes terms of the equivalent sphere and the moments that scale the quadratic angular drag
  const double eqD = 2.0/3.0 * (size[0] + size[1] + size[2]);
  const double cforce = 3.0 * MJB_PI * eqD;
  const double ctorq = MJB_PI * eqD*eqD*eqD;
  const double I_max = 8.0/15.0 * MJB_PI * d_mid * fluid_p4(d_max);
  double II[3];
  for (int k = 0; k < 3; k++) {
    const double d0 = size[k], d1 = size[(k + 1) % 3], d2 = size[(k + 2) % 3];
    II[k] = 8.0/15.0 * MJB_PI * d0 * fluid_p4(d1 > d2 ? d1 : d2);
  }
  const double mom[3] = {ang[0] * (angdrag*II[0] + slender*(I_max - II[0])),
                         ang[1] * (angdrag*II[1] + slender*(I_max - II[1])),
                         ang[2] * (angdrag*II[2] + slender*(I_max - II[2]))};
  const double drag_lin = viscosity*cforce + density*speed*(A_proj*blunt + slender*(A_max - A_proj));
  const double drag_ang = viscosity * ctorq + density * sqrt(mom[0]*mom[0] + mom[1]*mom[1] + mom[2]*mom[2]);
  for (int k = 0; k < 3; k++) {
    lfrc[k] -= drag_ang * ang[k];
    lfrc[3 + k] += fmag[k] + fkut[k] - drag_lin*lin[k];
  }
  for (int k = 0; k < 6; k++) lfrc[k] = lfrc[k]*fg[0];
}

// local wrench lfrc acting at world point p, frame R -> wrench about the tree origin O, added to wg
MJB_DI void fluid_add_wrench(double* wg, const double* lfrc, const double* R, const double* p, const double* O) {
  double T[3], F[3], cr[3];
  mulMatVec3(T, R, lfrc); mulMatVec3(F, R, lfrc + 3);
  const double r[3] = {p[0] - O[0], p[1] - O[1], p[2] - O[2]};
  cross3(cr, r, F);
  for (int k = 0; k < 3; k++) { wg[k] += cr[k] + T[k]; wg[3 + k] += F[k]; }
}

// the fluid wrench of one body. Out of line and free of the per-state context (an out-of-line call that
// takes Ctx& would push the whole context into local memory): the caller packs the body's kinematics
//   kin = [ V 6 (velocity about O) | O 3 | pos 3 | quat 4 | mat 9 | ip 3 | im 9 ]
// inside its `has_fluid` branch, so models without a medium pay one warp-uniform branch and no registers.
// fb: the body's fluid_body record; geom tables as in the model blob; [g0, g0 + gn) the body's geoms.
#define MJB_FLUID_KIN 37
MJB_COLD inline void fluid_wrench(double* wg, const double* kin, const double* fb, const mjbHdr* hdr,
                                  const int* geom_sameframe, const double* geom_pos, const double* geom_quat,
                                  const double* fluid_geom, int g0, int gn) {
  const double* V = kin; const double* O = kin + 6; const double* pos = kin + 9; const double* quat = kin + 12;
  const double* mat = kin + 16; const double* ip = kin + 25; const double* im = kin + 28;
  const int kind = (int)fb[0];
  double lvel[6], lfrc[6];
  if (kind == MJB_FLUID_BOX) {
    fluid_local_velocity(lvel, V, O, ip, im, hdr->wind);
    fluid_box_force(lfrc, lvel, fb + 1, hdr->density, hdr->viscosity);
    fluid_add_wrench(wg, lfrc, im, ip, O);
    return;
  }
  if (kind != MJB_FLUID_ELLIPSOID) return;
  for (int g = g0; g < g0 + gn; g++) {
    const double* fg = fluid_geom + MJB_FLUID_NG*g;
    if (fg[0] == 0.0) continue;
    // the geom's pose as mj_kinematics forms it (engine_core_smooth.c:139-157, sameframe shortcuts)
    const int sf = geom_sameframe[g];
    double gp[3], gm[9];
    if (sf == MJB_SAMEFRAME_BODY) {
      gp[0] = pos[0]; gp[1] = pos[1]; gp[2] = pos[2];
    } else if (sf == MJB_SAMEFRAME_INERTIA) {
      gp[0] = ip[0]; gp[1] = ip[1]; gp[2] = ip[2];
    } else {
      mulMatVec3(gp, mat, geom_pos + 3*g);
      gp[0] += pos[0]; gp[1] += pos[1]; gp[2] += pos[2];
    }
    if (sf == MJB_SAMEFRAME_NONE) {
      double tq[4];
      mulQuat(tq, quat, geom_quat + 4*g);
      quat2Mat(gm, tq);
    } else if (sf == MJB_SAMEFRAME_BODY || sf == MJB_SAMEFRAME_BODYROT) {
      for (int k = 0; k < 9; k++) gm[k] = mat[k];
    } else {
      for (int k = 0; k < 9; k++) gm[k] = im[k];
    }
    fluid_local_velocity(lvel, V, O, gp, gm, hdr->wind);
    fluid_ellipsoid_force(lfrc, lvel, fg, hdr->density, hdr->viscosity);
    fluid_add_wrench(wg, lfrc, gm, gp, O);
  }
}

// Body range [kLo, kHi) (kHi = 0: up to nbody). The generic kernels run the whole tree in one call;
// the model-specialised build cuts the expanded sweep into stages of a few thousand instructions,
// one kernel each, so that the code of a kernel stays resident in the instruction cache (measured:
// ONE expanded kernel of 24 K instructions ran 1.5x slower than the generic loop although it
// executes 2.6x fewer instructions -- every warp streams 390 KB of cold code per state). A stage
// that does not start at body 1 reloads the tree origin and lets its first body fetch the parent
// from scratch; a stage that does not end at the last body stores the carry of its last body.
// a finished cdof row: to the scratch (backward sweep, constraint rows) and, in a fused subtree
// stage, to the thread-local rows the inertia sweep of the same kernel reads
MJB_HD inline void store_cdof(Ctx& c, double* cdof, int dof, const double* cd) {
  sts(cdof, 6*dof, cd, 6);
  if (c.lcd) { for (int k = 0; k < 6; k++) c.lcd[6*(dof - c.ldof0) + k] = cd[k]; }
}

// kFluid: mj_fluid is compiled in. The generic CUDA kernels keep it out of the plain smooth-kernel instantiation
// (phase_smooth), whose register budget is tuned on models without a medium; everywhere else the flag of the model
// decides (a constant in the specialised build).
template <int kLo = 1, int kHi = 0, bool kHandOver = true, bool kFluid = true>
MJB_HD inline void forward_sweep(Ctx& c) {
  const mjbHdr& H = *c.H;
  const int nbody = H.nbody;
  const int lo = kLo, hi = kHi ? kHi : nbody;
  double* xpos = SC(xpos); double* xquat = SC(xquat); double* org = SC(origin);
  double* cvel = SC(cvel); double* cal = SC(cacc_lin); double* cacc = SC(cacc);
  double* cfrc = SC(cfrc); double* cinert = SC(cinert); double* cdof = SC(cdof);
  const int* body_parentid = MI(body_parentid);
  const int* body_jntadr = MI(body_jntadr);
  const int* body_jntnum = MI(body_jntnum);
  const int* body_dofadr = MI(body_dofadr);
  const int* body_dofnum = MI(body_dofnum);
  const int* body_mocapid = MI(body_mocapid);
  const int* body_sameframe = MI(body_sameframe);
  const int* tree_flags = MI(body_tree_flags);
  const int* jnt_type = MI(jnt_type);
  const int* jnt_qposadr = MI(jnt_qposadr);
  const int* jnt_dofadr = MI(jnt_dofadr);
  const int* dof_jntid = MI(dof_jntid);
  const double* body_pos = MD(body_pos); const double* body_quat = MD(body_quat);
  const double* body_ipos = MD(body_ipos); const double* body_iquat = MD(body_iquat);
  const double* body_mass = MD(body_mass); const double* body_inertia = MD(body_inertia);
  const double* jnt_pos = MD(jnt_pos); const double* jnt_axis = MD(jnt_axis);
  const double* qpos0 = MD(qpos0);

  // Carry of the body just processed, in per-thread shared-memory slots (on chip, and out of the
  // register budget of the joint loop): 0..2 pos, 3..6 quat, 7..12 cvel, 13..18 cacc, 19..24 cacc_lin
#define CS(k) c.sm[(k) * MJB_SMS]
  // world body: identity pose, zero velocity, acceleration = -gravity (mj_rne :1979-1982)
  double O[3] = {0, 0, 0};
  int carry = 0;          // body whose pose / velocity / acceleration are in the carry slots
  if (lo == 1) {
    double P[3] = {0, 0, 0}, Q[4] = {1, 0, 0, 0};
    double Z[6] = {0, 0, 0, 0, 0, 0}, A[6] = {0, 0, 0, 0, 0, 0};
    if (!(H.disableflags & MJB_DSBL_GRAVITY)) {
      A[3] = -H.gravity[0]; A[4] = -H.gravity[1]; A[5] = -H.gravity[2];
    }
    stc(xpos, 0, P, 3); stc(xquat, 0, Q, 4); stc(org, 0, O, 3);
    stc(cvel, 0, Z, 6); stc(cal, 0, Z, 6); stc(cacc, 0, A, 6);
    for (int k = 0; k < 3; k++) CS(k) = 0;
    CS(3) = 1; CS(4) = 0; CS(5) = 0; CS(6) = 0;
    for (int k = 0; k < 6; k++) { CS(7 + k) = 0; CS(13 + k) = A[k]; CS(19 + k) = 0; }
    const double I9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    body_geoms(c, 0, P, Q, I9, P, I9);
  } else {
    carry = -1;
    ldn(O, org, 3*MI(body_rootid)[lo], 3);
  }

  if (lo < nbody) prefetch_body_inputs(c, lo);
  auto sweep_body = [&](const int b) MJB_BODY_LAMBDA {
    if (b + 1 < nbody) prefetch_body_inputs(c, b + 1);
    const int pid = body_parentid[b];
    const int jntadr = body_jntadr[b], jntnum = body_jntnum[b];
    const int bda = body_dofadr[b], dofnum = body_dofnum[b];
    if (pid != carry) {
      double t[25];
      ldn(t, xpos, 3*pid, 3); ldn(t + 3, xquat, 4*pid, 4);
      ldn(t + 7, cvel, 6*pid, 6); ldn(t + 13, cacc, 6*pid, 6); ldn(t + 19, cal, 6*pid, 6);
      for (int k = 0; k < 25; k++) CS(k) = t[k];
    }
    double V[6];
    for (int k = 0; k < 6; k++) V[k] = CS(7 + k);
    double pos[3], quat[4];
    double t1[6] = {0, 0, 0, 0, 0, 0};   // cdof_dot' * qvel   (mju_mulDofVec, row by row)
    double t2[6] = {0, 0, 0, 0, 0, 0};   // cdof' * qacc
    const bool isfree = jntnum == 1 && jnt_type[jntadr] == MJB_JNT_FREE;
    bool has_ball = false;

    if (isfree) {
      const int qadr = jnt_qposadr[jntadr];
      for (int k = 0; k < 3; k++) pos[k] = QPOS(qadr + k);
      for (int k = 0; k < 4; k++) quat[k] = QPOS(qadr + 3 + k);
      normalize4(quat);
      if (pid == 0) { O[0] = pos[0]; O[1] = pos[1]; O[2] = pos[2]; stc(org, 3*b, O, 3); }
      quat_dof_forces(c, jntadr, qadr, bda, MJB_JNT_FREE, quat);
    } else {
      double bquat[4] = {body_quat[4*b], body_quat[4*b+1], body_quat[4*b+2], body_quat[4*b+3]};
      const int mid = body_mocapid[b];
      if (mid >= 0) {                  // mocap pose: the caller's, else the model pose (mj_resetData default)
        if (c.out.mocap_quat) {
          for (int k = 0; k < 4; k++) bquat[k] = c.out.mocap_quat[(size_t)(4*mid + k)*(size_t)c.N + c.s];
        }
        normalize4(bquat);
      }
      if (pid) {
        double pm[9];
        const double Q[4] = {CS(3), CS(4), CS(5), CS(6)};
        quat2Mat(pm, Q);                 // == the parent's xmat (same function of the same xquat)
        mulMatVec3(pos, pm, body_pos + 3*b);
        pos[0] += CS(0); pos[1] += CS(1); pos[2] += CS(2);
        mulQuat(quat, Q, bquat);
      } else {
        for (int k = 0; k < 3; k++) pos[k] = body_pos[3*b + k];
        if (mid >= 0 && c.out.mocap_pos) {
          for (int k = 0; k < 3; k++) pos[k] = c.out.mocap_pos[(size_t)(3*mid + k)*(size_t)c.N + c.s];
        }
        for (int k = 0; k < 4; k++) quat[k] = bquat[k];
        O[0] = pos[0]; O[1] = pos[1]; O[2] = pos[2];
        stc(org, 3*b, O, 3);
      }
      MJB_UNROLL
      for (int j = 0; j < jntnum; j++) has_ball = has_ball || jnt_type[jntadr + j] == MJB_JNT_BALL;

      MJB_UNROLL
      for (int j = 0; j < jntnum; j++) {
        const int jid = jntadr + j;
        const int qadr = jnt_qposadr[jid];
        const int dadr = jnt_dofadr[jid];
        const int jtype = jnt_type[jid];
        double ax[3], anc[3], cd[6];
        rotVecQuat(ax, jnt_axis + 3*jid, quat);
        rotVecQuat(anc, jnt_pos + 3*jid, quat);
        anc[0] += pos[0]; anc[1] += pos[1]; anc[2] += pos[2];
        const double off[3] = {O[0] - anc[0], O[1] - anc[1], O[2] - anc[2]};

        double jq = 0, jqv = 0, jqa = 0;
        if (jtype != MJB_JNT_BALL) { jq = QPOS(qadr); jqv = QVEL(dadr); jqa = QACC(dadr); }
        if (jtype == MJB_JNT_SLIDE) {
          const double q = jq - qpos0[qadr];
          pos[0] += ax[0]*q; pos[1] += ax[1]*q; pos[2] += ax[2]*q;
          cd[0] = 0; cd[1] = 0; cd[2] = 0; cd[3] = ax[0]; cd[4] = ax[1]; cd[5] = ax[2];
        } else {
          double qloc[4];
          if (jtype == MJB_JNT_BALL) {
            for (int k = 0; k < 4; k++) qloc[k] = QPOS(qadr + k);
            normalize4(qloc);
            // cdof of a ball joint uses the body's FINAL orientation (mj_comPos :243-252): keep
            // the anchor offset in the dof's slots until the pose is complete
            stn(cdof, 6*dadr, off, 3);
            quat_dof_forces(c, jid, qadr, dadr, MJB_JNT_BALL, qloc);
          } else {
            // mju_axisAngle2Quat (engine_util_spatial.c:97)
            const double angle = jq - qpos0[qadr];
            double sn, cs;
            sincos(angle*0.5, &sn, &cs);
            qloc[0] = cs;
            qloc[1] = jnt_axis[3*jid]*sn; qloc[2] = jnt_axis[3*jid+1]*sn; qloc[3] = jnt_axis[3*jid+2]*sn;
            cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
            cross3(cd + 3, ax, off);
          }
          mulQuat(quat, quat, qloc);
          double vec[3];
          rotVecQuat(vec, jnt_pos + 3*jid, quat);
          pos[0] = anc[0] - vec[0]; pos[1] = anc[1] - vec[1]; pos[2] = anc[2] - vec[2];
        }
        if (jtype != MJB_JNT_BALL) {
          store_cdof(c, cdof, dadr, cd);
          if (!has_ball) {
            // mj_comVel / mj_rne for a scalar dof, fused: cdof_dot uses the velocity so far
            double dd[6];
            crossMotion(dd, V, cd);
            for (int k = 0; k < 6; k++) { t1[k] += dd[k]*jqv; V[k] += cd[k]*jqv; t2[k] += cd[k]*jqa; }
          }
          scalar_dof_forces(c, jid, qadr, dadr, jq, jqv, jqa);
        }
      }
    }

    normalize4(quat);
    double mat[9];
    quat2Mat(mat, quat);
    // the pose goes to scratch only where a later consumer reads it: a child that is not b+1
    // (bit 2), an equality constraint or a tendon site on this body (bit 3), or the debug dump
    if ((tree_flags[b] & 12) || c.out.scratch_dump || c.out.sensordata || c.out.fwd_xfrc || c.out.cam_xpos || c.out.actuator_length ||
        c.out.xfrc_applied) {
      stc(xquat, 4*b, quat, 4);
      stc(xpos, 3*b, pos, 3);
    }

    if (isfree) {
      // translational dofs: cdof = [0, e_r], cdof_dot = 0
      for (int r = 0; r < 3; r++) {
        double cd[6] = {0, 0, 0, r == 0 ? 1.0 : 0.0, r == 1 ? 1.0 : 0.0, r == 2 ? 1.0 : 0.0};
        store_cdof(c, cdof, bda + r, cd);
        V[3 + r] += QVEL(bda + r);
        t2[3 + r] += QACC(bda + r);
      }
      // rotational dofs: body axes; the anchor is the body origin, O - anchor = (O - pos)
      const double off[3] = {O[0] - pos[0], O[1] - pos[1], O[2] - pos[2]};
      double cd[3][6];
      for (int r = 0; r < 3; r++) {
        cd[r][0] = mat[r]; cd[r][1] = mat[r + 3]; cd[r][2] = mat[r + 6];
        cross3(cd[r] + 3, cd[r], off);
        store_cdof(c, cdof, bda + 3 + r, cd[r]);
      }
      // all three use the velocity BEFORE this joint's rotation (mj_comVel :1855-1876)
      for (int r = 0; r < 3; r++) {
        double dd[6];
        crossMotion(dd, V, cd[r]);
        const double qv = QVEL(bda + 3 + r);
        for (int k = 0; k < 6; k++) t1[k] += dd[k]*qv;
      }
      for (int r = 0; r < 3; r++) {
        const double qv = QVEL(bda + 3 + r), qa = QACC(bda + 3 + r);
        for (int k = 0; k < 6; k++) { V[k] += cd[r][k]*qv; t2[k] += cd[r][k]*qa; }
      }
    } else if (has_ball) {
      // general path: finish the ball-joint cdofs with the final orientation, then run the dof
      // loop of mj_comVel over the body's dofs from scratch
      MJB_UNROLL
      for (int j = 0; j < jntnum; j++) {
        const int jid = jntadr + j;
        if (jnt_type[jid] != MJB_JNT_BALL) continue;
        const int dadr = jnt_dofadr[jid];
        double off[3];
        ldn(off, cdof, 6*dadr, 3);
        for (int r = 0; r < 3; r++) {
          double cd[6] = {mat[r], mat[r + 3], mat[r + 6], 0, 0, 0};
          cross3(cd + 3, cd, off);
          store_cdof(c, cdof, dadr + r, cd);
        }
      }
      MJB_UNROLL
      for (int j = 0; j < dofnum; j++) {
        const int jt = jnt_type[dof_jntid[bda + j]];
        if (jt == MJB_JNT_BALL) {
          double cd[3][6];
          for (int r = 0; r < 3; r++) {
            double dd[6];
            ldn(cd[r], cdof, 6*(bda + j + r), 6);
            crossMotion(dd, V, cd[r]);
            const double qv = QVEL(bda + j + r);
            for (int k = 0; k < 6; k++) t1[k] += dd[k]*qv;
          }
          for (int r = 0; r < 3; r++) {
            const double qv = QVEL(bda + j + r), qa = QACC(bda + j + r);
            for (int k = 0; k < 6; k++) { V[k] += cd[r][k]*qv; t2[k] += cd[r][k]*qa; }
          }
          j += 2;
        } else {
          double cd[6], dd[6];
          ldn(cd, cdof, 6*(bda + j), 6);
          crossMotion(dd, V, cd);
          const double qv = QVEL(bda + j), qa = QACC(bda + j);
          for (int k = 0; k < 6; k++) { t1[k] += dd[k]*qv; V[k] += cd[k]*qv; t2[k] += cd[k]*qa; }
        }
      }
    }

    double A[6], AL[6];
    for (int k = 0; k < 6; k++) {
      A[k] = CS(13 + k); A[k] += t1[k]; A[k] += t2[k];
      AL[k] = CS(19 + k) + t2[k];
      CS(7 + k) = V[k]; CS(13 + k) = A[k]; CS(19 + k) = AL[k];
    }
    // read back only by a child that is not b+1 (bit 2), by constraint rows on this body (bit 4:
    // candidate pairs, equality constraints, tendon sites), or by the optional per-body outputs
    if ((tree_flags[b] & (4 | 16)) || c.out.scratch_dump || c.out.sensordata || c.out.fwd_xfrc || c.out.fwdinv ||
        c.out.cacc || c.out.qfrc_bias || c.out.energy) {
      stc(cvel, 6*b, V, 6);
      stc(cal, 6*b, AL, 6);
    }
    // body of a candidate pair: the carrier record the contact rows gather (one 128-byte line)
    if (tree_flags[b] & 32) store_crec(c, b, V, AL, O);
    // read back only by a child that is not b+1, or by the mj_rnePostConstraint outputs
    if ((tree_flags[b] & 4) || c.out.cacc || c.out.qfrc_bias) stc(cacc, 6*b, A, 6);

    // inertial frame (mj_kinematics :159-165), cinert (mju_inertCom) and the rne body force
    const int sf = body_sameframe[b];
    double ip[3], im[9];
    if (sf == MJB_SAMEFRAME_BODY) {
      ip[0] = pos[0]; ip[1] = pos[1]; ip[2] = pos[2];
    } else {
      mulMatVec3(ip, mat, body_ipos + 3*b);
      ip[0] += pos[0]; ip[1] += pos[1]; ip[2] += pos[2];
    }
    if (sf == MJB_SAMEFRAME_NONE) {
      double tq[4];
      mulQuat(tq, quat, body_iquat + 4*b);
      quat2Mat(im, tq);
    } else {
      for (int k = 0; k < 9; k++) im[k] = mat[k];
    }
    {
      const double off[3] = {ip[0] - O[0], ip[1] - O[1], ip[2] - O[2]};
      double ci[10], f[6], u1[6], u2[6];
      inertCom(ci, body_inertia + 3*b, im, off, body_mass[b]);
      if (c.lci) { for (int k = 0; k < 10; k++) c.lci[10*(b - c.lbody0) + k] = ci[k]; }
      if (!c.lci || c.out.cfrc_int || c.out.qfrc_bias || c.out.sensordata || c.out.scratch_dump || c.out.energy) sts(cinert, 10*b, ci, 10);
      mulInertVec(f, ci, A);
      mulInertVec(u1, ci, V);
      crossForce(u2, V, u1);
      for (int k = 0; k < 6; k++) f[k] += u2[k];
      sts(cfrc, 6*b, f, 6);
      if (H.passive_wrench) {
        // passive-wrench carrier (projected into qfrc_passive by the backward sweep): starts with
        // mj_gravcomp (engine_passive.c:381-401), force -gravity*mass*gravcomp at the body's centre
        // of mass as a wrench about O; spatial-tendon springs and dampers are added later
        double wg[6] = {0, 0, 0, 0, 0, 0};
        if (H.has_gravcomp) {
          const double sgc = -(body_mass[b] * MD(body_gravcomp)[b]);
          const double F[3] = {H.gravity[0]*sgc, H.gravity[1]*sgc, H.gravity[2]*sgc};
          cross3(wg, off, F);
          wg[3] = F[0]; wg[4] = F[1]; wg[5] = F[2];
        }
        if (kFluid && H.has_fluid) {
          const double* fb = MD(fluid_body) + 4*b;
          if (fb[0] != 0.0) {
            double kin[MJB_FLUID_KIN], wf[6] = {0, 0, 0, 0, 0, 0};
            for (int k = 0; k < 6; k++) kin[k] = V[k];
            for (int k = 0; k < 3; k++) { kin[6 + k] = O[k]; kin[9 + k] = pos[k]; kin[25 + k] = ip[k]; }
            for (int k = 0; k < 4; k++) kin[12 + k] = quat[k];
            for (int k = 0; k < 9; k++) { kin[16 + k] = mat[k]; kin[28 + k] = im[k]; }
            fluid_wrench(wf, kin, fb, c.H, MI(geom_sameframe), MD(geom_pos), MD(geom_quat), MD(fluid_geom),
                         MI(body_geomadr)[b], MI(body_geomnum)[b]);
            for (int k = 0; k < 6; k++) wg[k] += wf[k];
          }
        }
        stn(SC(cfrc_gc), 6*b, wg, 6);
      }
    }

    body_geoms(c, b, pos, quat, mat, ip, im);

    CS(0) = pos[0]; CS(1) = pos[1]; CS(2) = pos[2];
    CS(3) = quat[0]; CS(4) = quat[1]; CS(5) = quat[2]; CS(6) = quat[3];
    carry = b;
  };
  MJB_BODY_LOOP_UP(sweep_body, lo, hi, kLo, (kHi ? kHi : MJB_SPEC_NBODY));
  if (kHandOver && hi < nbody && carry > 0) {
    // hand-over to the next stage: everything a child may read of the last body of this one
    double t[25];
    for (int k = 0; k < 25; k++) t[k] = CS(k);
    stc(xpos, 3*carry, t, 3); stc(xquat, 4*carry, t + 3, 4);
    stc(cvel, 6*carry, t + 7, 6); stc(cacc, 6*carry, t + 13, 6); stc(cal, 6*carry, t + 19, 6);
  }
#undef CS
}


#endif  // MJB_SWEEP_H_
