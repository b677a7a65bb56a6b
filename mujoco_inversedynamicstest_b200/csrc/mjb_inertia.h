// Part of the per-state mj_inverse pipeline (mjb_pipeline.h includes it inside namespace mjb, after the
// context and accessor macros; not a stand-alone header).
// mj_crb + mj_factorM as one articulated-body sweep, mj_discreteAcc (Euler, implicitfast, implicit).
#ifndef MJB_INERTIA_H_
#define MJB_INERTIA_H_

// ------------------------------------------------------------------------------------------
// mj_crb (engine_core_smooth.c:1353-1401) and mj_factorM / mj_factorI (:1470-1511) in one
// leaves-to-root sweep.
//
// qM is the reference's composite-rigid-body form: M(k,i) = cdof_i . (crb_body(k) cdof_k) for i on
// the ancestor chain of k. For qLD the reference eliminates rows of M in place, O(sum depth^2)
// read-modify-writes of a matrix that would have to live in HBM here. The same unique L'DL factors
// are obtained without touching M from the articulated-body recursion (Featherstone, RBDA ch. 6/7):
// with every spatial quantity already expressed in the common com-based frame,
//     IA_b   = cinert_b + sum_children IA_c          (6x6 symmetric, 21 numbers)
//     U_k    = IA_b cdof_k ,  D_k = cdof_k.U_k + armature_k        (dofs of b, last to first)
//     L(k,i) = cdof_i.U_k / D_k  for every ancestor dof i ;  IA_b -= U_k U_k' / D_k
// so each entry of qLD is computed once (one dot product) and written once, and the ancestor walk
// is shared with the qM entries. Values agree with mj_factorI to rounding.

// y = A x for a symmetric 6x6 stored as its 21 upper-triangular entries (row-major)
MJB_DI void sym6_mul(double* y, const double* A, const double* x) {
  // explicitly fused (see mjb_math.h): used only by the inertia sweep
#define MJB_ROW6(a0, a1, a2, a3, a4, a5) \
  fma(A[a5], x[5], fma(A[a4], x[4], fma(A[a3], x[3], fma(A[a2], x[2], fma(A[a1], x[1], A[a0]*x[0])))))
  y[0] = MJB_ROW6(0, 1, 2, 3, 4, 5);
  y[1] = MJB_ROW6(1, 6, 7, 8, 9, 10);
  y[2] = MJB_ROW6(2, 7, 11, 12, 13, 14);
  y[3] = MJB_ROW6(3, 8, 12, 15, 16, 17);
  y[4] = MJB_ROW6(4, 9, 13, 16, 18, 19);
  y[5] = MJB_ROW6(5, 10, 14, 17, 19, 20);
#undef MJB_ROW6
}

// the 10-number rigid inertia of mju_inertCom as a symmetric 6x6 (layout of mju_mulInertVec)
MJB_DI void inert_to_sym6(double* A, const double* i) {
  A[0] = i[0];  A[1] = i[3];  A[2] = i[4];  A[3] = 0;     A[4] = -i[8]; A[5] = i[7];
  A[6] = i[1];  A[7] = i[5];  A[8] = i[8];  A[9] = 0;     A[10] = -i[6];
  A[11] = i[2]; A[12] = -i[7]; A[13] = i[6]; A[14] = 0;
  A[15] = i[9]; A[16] = 0;    A[17] = 0;
  A[18] = i[9]; A[19] = 0;
  A[20] = i[9];
}

// A += the symmetric 6x6 of a 10-number rigid inertia (same layout as inert_to_sym6)
MJB_DI void inert_add_sym6(double* A, const double* i) {
  A[0] += i[0];  A[1] += i[3];  A[2] += i[4];  A[4] -= i[8]; A[5] += i[7];
  A[6] += i[1];  A[7] += i[5];  A[8] += i[8];  A[10] -= i[6];
  A[11] += i[2]; A[12] -= i[7]; A[13] += i[6];
  A[15] += i[9];
  A[18] += i[9];
  A[20] += i[9];
}

// Body range [kLo, kHi) as in forward_sweep, visited downwards. At a stage boundary the register
// hand-over from body p+1 to its parent p becomes a push through the parent's scratch accumulators.
template <int kLo = 1, int kHi = 0>
MJB_HD inline void inertia(Ctx& c) {
  const mjbHdr& H = *c.H;
  const int nbody = H.nbody;
  const int lo = kLo, hi = kHi ? kHi : nbody;
  double* crb = SC(crb); double* ia = SC(ia); double* cinert = SC(cinert); double* cdof = SC(cdof);
  const int* body_parentid = MI(body_parentid);
  const int* body_dofadr = MI(body_dofadr);
  const int* body_dofnum = MI(body_dofnum);
  const int* dof_parentid = MI(dof_parentid);
  const int* dof_Madr = MI(dof_Madr);
  const int* dof_simplenum = MI(dof_simplenum);
  const int* rownnz = MI(C_rownnz); const int* rowadr = MI(C_rowadr);
  const double* armature = MD(dof_armature);
  const double* dof_M0 = MD(dof_M0);
  const size_t N = (size_t)c.N;
  double* qM = c.out.qM + c.s; double* qLD = c.out.qLD + c.s; double* qLDiagInv = c.out.qLDiagInv + c.s;

  // Children are folded into their parent leaves-to-root. Bodies are in depth-first order, so the
  // child visited last before a body p is p+1: it hands its sums over in registers (carry). Only
  // the other children go through the parent's scratch accumulators, the first of them (highest
  // index, bit1) by a plain store, the rest by a batched read-modify-write; no clearing pass.
  const int* tree_flags = MI(body_tree_flags);   // bit1: highest-index child, bit2: has a child != body+1
  bool carried = false;          // cr, A hold the sums handed over by body b+1
  double cr[10], A[21];
  auto inertia_body = [&](const int b) MJB_BODY_LAMBDA {
    const int flags = tree_flags[b];
#if defined(__CUDA_ARCH__) && !defined(MJB_NO_INERTIA_PREFETCH)
    // the rows the NEXT body of the sweep (b - 1) will read -- its cinert, the cdofs of its dofs and,
    // where other children pushed into it, its accumulators -- are requested now, so that their HBM
    // latency runs under this body's arithmetic (the kernel holds 12 warps per SM: nothing else hides it)
    if (b - 1 >= lo && !c.lci) {
      const int nb = b - 1;
      for (int j = 0; j < 10; j++) prefetch_line(cinert + (size_t)(10*nb + j) * MJB_LS);
      const int a0 = body_dofadr[nb], an = body_dofnum[nb];
      MJB_UNROLL
      for (int k = a0; k < a0 + an; k++) {
        for (int j = 0; j < 6; j++) prefetch_line(cdof + (size_t)(6*k + j) * MJB_LS);
      }
      if (tree_flags[nb] & 4) {
        for (int j = 0; j < 10; j++) prefetch_line(crb + (size_t)(10*nb + j) * MJB_LS);
        for (int j = 0; j < 21; j++) prefetch_line(ia + (size_t)(21*nb + j) * MJB_LS);
      }
    }
#endif
    {
      double ci[10];
      if (c.lci) { for (int j = 0; j < 10; j++) ci[j] = c.lci[10*(b - c.lbody0) + j]; }
      else ldn_ro(ci, cinert, 10*b, 10);
      if (carried) {
        for (int j = 0; j < 10; j++) cr[j] += ci[j];
        inert_add_sym6(A, ci);
      } else {
        for (int j = 0; j < 10; j++) cr[j] = ci[j];
        inert_to_sym6(A, ci);
      }
    }
    // sums pushed through scratch: by children other than b+1, and by b+1 itself when it was the
    // last body of the previous stage
    if ((flags & 4) || (b + 1 == hi && hi < nbody && body_parentid[b + 1] == b)) {
      double pc[10], pA[21];
      ldn(pc, crb, 10*b, 10); ldn(pA, ia, 21*b, 21);
      for (int j = 0; j < 10; j++) cr[j] += pc[j];
      for (int j = 0; j < 21; j++) A[j] += pA[j];
    }
    const int adr0 = body_dofadr[b], num = body_dofnum[b];
    MJB_UNROLL
    for (int k = adr0 + num - 1; k >= adr0; k--) {
      const int madr = dof_Madr[k];
      const int diag = rowadr[k] + rownnz[k] - 1;
      if (dof_simplenum[k]) {
        // simple body: M is diagonal and constant (engine_core_smooth.c:1375-1385, :1498)
        const double m0 = dof_M0[k];
        qM[(size_t)madr*N] = m0;
        qLD[(size_t)diag*N] = m0;
        qLDiagInv[(size_t)k*N] = 1/m0;
        int t = 1;    // legacy qM keeps the (zero) ancestor entries, the reduced qLD row does not
        for (int i = dof_parentid[k]; i >= 0; i = dof_parentid[i], t++) qM[(size_t)(madr + t)*N] = 0;
        continue;
      }
      double S[6], buf[6], U[6];
      if (c.lcd) { for (int j = 0; j < 6; j++) S[j] = c.lcd[6*(k - c.ldof0) + j]; }
      else ldn_ro(S, cdof, 6*k, 6);
      mulInertVecF(buf, cr, S);
      sym6_mul(U, A, S);
      const double Mkk = armature[k] + dot6f(S, buf);
      const double D = armature[k] + dot6f(S, U);
      const double invD = 1/D;
      qM[(size_t)madr*N] = Mkk;
      qLD[(size_t)diag*N] = D;
      qLDiagInv[(size_t)k*N] = invD;
      // ancestor walk, MJB_ANC ancestors at a time: their cdofs are loaded together before any of
      // the results is stored (one memory round trip per group instead of one per ancestor)
      int t = 1;
      for (int i = dof_parentid[k]; i >= 0;) {
        double Si[MJB_ANC][6];
        int n = 0;
#pragma unroll
        for (int g = 0; g < MJB_ANC; g++) {
          if (i >= 0) {
            if (c.lcd && i >= c.ldof0) { for (int j = 0; j < 6; j++) Si[g][j] = c.lcd[6*(i - c.ldof0) + j]; }
            else ldn_ro(Si[g], cdof, 6*i, 6);
            i = dof_parentid[i]; n = g + 1;
          }
        }
#pragma unroll
        for (int g = 0; g < MJB_ANC; g++) {
          if (g < n) {
            qM[(size_t)(madr + t + g)*N] = dot6f(Si[g], buf);
            qLD[(size_t)(diag - t - g)*N] = dot6f(Si[g], U) * invD;
          }
        }
        t += n;
      }
      // IA -= U U' / D
      int e = 0;
      for (int r = 0; r < 6; r++) {
        const double ur = U[r]*invD;
        for (int q = r; q < 6; q++, e++) A[e] = fma(-ur, U[q], A[e]);
      }
    }
    const int p = body_parentid[b];
    if (p > 0) {
      if (b == p + 1 && b != lo) {
        carried = true;          // cr, A stay in registers for the parent, which is visited next
        return;
      }
      carried = false;
      if (flags & 2) {
        stn(crb, 10*p, cr, 10);
        stn(ia, 21*p, A, 21);
      } else {
        double pc[10], pA[21];
        ldn(pc, crb, 10*p, 10); ldn(pA, ia, 21*p, 21);
        for (int j = 0; j < 10; j++) pc[j] += cr[j];
        for (int j = 0; j < 21; j++) pA[j] += A[j];
        stn(crb, 10*p, pc, 10); stn(ia, 21*p, pA, 21);
      }
    } else {
      carried = false;
    }
  };
  MJB_BODY_LOOP_DOWN(inertia_body, lo, hi, kLo, (kHi ? kHi : MJB_SPEC_NBODY));
}

// ------------------------------------------------------------------------------------------
// mj_discreteAcc (engine_inverse.c:81-164): the discrete-time qacc is converted to the
// continuous-time one the rest of mj_inverse works with. Euler:
//     qacc' = M^-1 (M + h diag(B)) qacc = qacc + h M^-1 (B .* qacc),
// implicitfast / implicit: qacc' = qacc - h M^-1 (qDeriv qacc) with the analytic qDeriv (below),
// using the L'DL factors of the inertia kernel. mj_solveLD (engine_core_smooth.c:1629-1707) on the
// reduced row layout of qLD (row i: ancestors ascending, diagonal last). The reference forms
// (M + hB) qacc with mj_mulM and solves; the two agree to rounding (cond(M) * eps).
// Gravity-free bias force of the velocity field w = sv*qvel + sa*qacc: mj_comVel + mj_rne(flg_acc = 0)
// (engine_core_smooth.c:1833-1895, 1969-2023) restated on the cdof / cinert rows of the scratch, with
// the per-body cvel, cacc and force in the ia rows (18 of the 21 doubles per body). The result is
// ADDED to dst with the factor scale. The function is exactly quadratic in w, which is what the
// implicit integrator's mjd_rne_vel term is built from (discrete_acc below).
MJB_HD inline void coriolis_add(Ctx& c, double sv, double sa, double scale, double* dst) {
  const mjbHdr& H = *c.H;
  const int* body_parentid = MI(body_parentid); const int* body_dofadr = MI(body_dofadr);
  const int* body_dofnum = MI(body_dofnum); const int* dof_jntid = MI(dof_jntid); const int* jnt_type = MI(jnt_type);
  const int* dof_bodyid = MI(dof_bodyid);
  double* t = SC(ia);
  const double zero[18] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  stn(t, 0, zero, 18);
  for (int i = 1; i < H.nbody; i++) {
    double w[18], ci[10], tmp[6], tmp1[6];
    ldn(w, t, 18*body_parentid[i], 12);
    double* cvel = w; double* cacc = w + 6; double* f = w + 12;
    const int bda = body_dofadr[i], dofnum = body_dofnum[i];
    for (int j = 0; j < dofnum; j++) {
      const int jt = jnt_type[dof_jntid[bda + j]];
      int first = j, count = 1;
      if (jt == MJB_JNT_FREE) {
        // translational dofs: no cdof_dot, the velocity is added first
        for (int k = 0; k < 3; k++) {
          double cd[6];
          ldn(cd, SC(cdof), 6*(bda + k), 6);
          const double wj = sv*QVEL(bda + k) + sa*QACC(bda + k);
          for (int r = 0; r < 6; r++) cvel[r] += cd[r]*wj;
        }
        first = j + 3; count = 3;
      } else if (jt == MJB_JNT_BALL) {
        count = 3;
      }
      // cdof_dot of the group from the velocity before the group, then the group's velocity
      double add[6] = {0, 0, 0, 0, 0, 0};
      for (int k = 0; k < count; k++) {
        double cd[6], cdd[6];
        ldn(cd, SC(cdof), 6*(bda + first + k), 6);
        const double wj = sv*QVEL(bda + first + k) + sa*QACC(bda + first + k);
        crossMotion(cdd, cvel, cd);
        for (int r = 0; r < 6; r++) { cacc[r] += cdd[r]*wj; add[r] += cd[r]*wj; }
      }
      for (int r = 0; r < 6; r++) cvel[r] += add[r];
      j = first + count - 1;
    }
    ldn(ci, SC(cinert), 10*i, 10);
    mulInertVec(f, ci, cacc);
    mulInertVec(tmp, ci, cvel);
    crossForce(tmp1, cvel, tmp);
    for (int r = 0; r < 6; r++) f[r] += tmp1[r];
    stn(t, 18*i, w, 18);
  }
  for (int i = H.nbody - 1; i > 0; i--) {
    const int p = body_parentid[i];
    if (!p) continue;
    double f[6], pf[6];
    ldn(f, t, 18*i + 12, 6); ldn(pf, t, 18*p + 12, 6);
    for (int r = 0; r < 6; r++) pf[r] += f[r];
    stn(t, 18*p + 12, pf, 6);
  }
  for (int j = 0; j < H.nv; j++) {
    double cd[6], f[6];
    ldn(cd, SC(cdof), 6*j, 6); ldn(f, t, 18*dof_bodyid[j] + 12, 6);
    AT(dst, j) += scale*dot6(cd, f);
  }
}

MJB_HD inline void discrete_acc(Ctx& c, double* qacc_out) {
  const mjbHdr& H = *c.H;
  const int nv = H.nv;
  const size_t N = (size_t)c.N;
  const int* rownnz = MI(C_rownnz); const int* rowadr = MI(C_rowadr); const int* colind = MI(C_colind);
  const int* simplenum = MI(dof_simplenum);
  const double* damping = MD(dof_damping);
  const double* qLD = c.out.qLD + c.s; const double* qLDiagInv = c.out.qLDiagInv + c.s;
  double* x = SC(qfrc_c);                 // free at this point: the smooth phase is rerun afterwards
  const bool fast = H.discrete_acc >= 2;        // implicitfast (2) or implicit (3)
  const bool full = H.discrete_acc == 3;
  // x = -h * qDeriv * qacc. Euler: qDeriv = -diag(damping) (engine_inverse.c:111-116). implicitfast
  // (:133-152): qDeriv = sum_actuators bias_vel * m'm  -  diag(damping)  -  sum_tendons damping * J'J
  // (mjd_actuator_vel, mjd_passive_vel) on M's sparsity pattern, i.e. an entry (a, b) of a product
  // J'J exists only when one of the two dofs is an ancestor of the other (or a == b)
  const bool damp = !fast || !(H.disableflags & MJB_DSBL_PASSIVE);
  for (int i = 0; i < nv; i++) AT(x, i) = damp ? H.timestep * damping[i] * QACC(i) : 0.0;
  if (fast) {
    const int* dof_parentid = MI(dof_parentid);
    // mj_mulM multiplies with the modified M (engine_support.c:966-1020): the off-diagonal entries of
    // a "simple" dof's row are not visited
    auto related = [&](int a, int b) {
      int lo = a < b ? a : b, hi = a < b ? b : a;
      if (!full && hi != lo && simplenum[hi]) return false;
      while (hi > lo) hi = dof_parentid[hi];
      return hi == lo;
    };
    if (!(H.disableflags & MJB_DSBL_PASSIVE)) {
      const int* tendon_adr = MI(tendon_adr); const int* tendon_num = MI(tendon_num);
      const int* wrap_objid = MI(wrap_objid); const int* jnt_dofadr = MI(jnt_dofadr);
      const double* wrap_prm = MD(wrap_prm); const double* tdamp = MD(tendon_damping);
      for (int t = 0; t < H.ntendon; t++) {
        if (!(tdamp[t] > 0)) continue;
        const int adr = tendon_adr[t], num = tendon_num[t];
        for (int k = 0; k < num; k++) {
          const int a = jnt_dofadr[wrap_objid[adr + k]];
          double s = 0;
          for (int l = 0; l < num; l++) {
            const int b = jnt_dofadr[wrap_objid[adr + l]];
            if (related(a, b)) s += wrap_prm[adr + l] * QACC(b);
          }
          AT(x, a) += H.timestep * tdamp[t] * wrap_prm[adr + k] * s;
        }
      }
    }
    if (full) {
      // implicit (engine_inverse.c:120-131): qDeriv also carries -d qfrc_bias / d qvel (mjd_rne_vel,
      // engine_derivative.c:604-686) and the product runs over the full dof-dof pattern. The bias force
      // c(v) is exactly quadratic in v, so its derivative along qacc is the polarisation
      //   (dc/dv) a = c(v + a) - c(v) - c(a)      (gravity excluded: constant in v)
      double* r = SC(qfrc_passive);        // free here as well: recomputed by the second smooth pass
      for (int i = 0; i < nv; i++) AT(r, i) = 0;
      coriolis_add(c, 1.0, 1.0, 1.0, r);
      coriolis_add(c, 1.0, 0.0, -1.0, r);
      coriolis_add(c, 0.0, 1.0, -1.0, r);
      for (int i = 0; i < nv; i++) AT(x, i) += H.timestep * AT(r, i);
    }
    if (H.discrete_trn) {
      const double* bv = MD(act_biasvel);
      for (int u = 0; u < H.nu; u++) {
        if (bv[u] == 0) continue;
        const double* row = c.out.actuator_moment + (size_t)u*nv*N + c.s;
        for (int a = 0; a < nv; a++) {
          const double ra = row[(size_t)a*N];
          if (ra == 0) continue;
          double s = 0;
          for (int b = 0; b < nv; b++) {
            const double rb = row[(size_t)b*N];
            if (rb != 0 && related(a, b)) s += rb * QACC(b);
          }
          AT(x, a) -= H.timestep * bv[u] * ra * s;
        }
      }
    }
  }
  // x <- L^-T x
  for (int i = nv - 1; i > 0; i--) {
    if (simplenum[i]) continue;
    const double xi = AT(x, i);
    if (xi != 0) {
      const int start = rowadr[i], end = start + rownnz[i] - 1;
      for (int adr = start; adr < end; adr++) AT(x, colind[adr]) -= qLD[(size_t)adr*N] * xi;
    }
  }
  // x <- D^-1 x
  for (int i = 0; i < nv; i++) AT(x, i) *= qLDiagInv[(size_t)i*N];
  // x <- L^-1 x
  for (int i = 1; i < nv; i++) {
    if (simplenum[i]) continue;
    const int d = rownnz[i] - 1;
    if (d > 0) {
      const int adr = rowadr[i];
      double acc = 0;
      for (int k = 0; k < d; k++) acc += qLD[(size_t)(adr + k)*N] * AT(x, colind[adr + k]);
      AT(x, i) -= acc;
    }
  }
  for (int i = 0; i < nv; i++) qacc_out[(size_t)i*N] = QACC(i) + AT(x, i);
}


#endif  // MJB_INERTIA_H_
