// Part of the per-state mj_inverse pipeline (mjb_pipeline.h includes it inside namespace mjb, after the
// context and accessor macros; not a stand-alone header).
// Constraint-row arithmetic, equality rows, tendons (fixed and spatial, mju_wrap), passive forces, the per-dof pieces of friction loss / limits.
#ifndef MJB_ROWS_H_
#define MJB_ROWS_H_

// ------------------------------------------------------------------------------------------
// constraint-row arithmetic shared by all row types

// getimpedance (engine_core_constraint.c:1441-1489) on pre-clamped solimp
// general power: a = 1/power(mid, p-1) ; y = a*power(x, p)  /  mirrored above the midpoint
MJB_COLD inline double impedance_power(double x, double mid, double p) {
  if (x <= mid) {
    const double a = 1/pow(mid, p - 1);
    return a*pow(x, p);
  }
  const double b = 1/pow(1 - mid, p - 1);
  return 1 - b*pow(1 - x, p);
}

MJB_HD inline double impedance(const double* sp, double pos, double margin) {
  const double d0 = sp[MJB_SP_D0], d1 = sp[MJB_SP_D1], width = sp[MJB_SP_WIDTH];
  if (d0 == d1 || width <= MJB_MINVAL) return 0.5*(d0 + d1);
  double x = (pos - margin) / width;
  if (x < 0) x = -x;
  if (x >= 1 || x <= 0) return x >= 1 ? d1 : d0;
  const double mid = sp[MJB_SP_MID], p = sp[MJB_SP_POWER];
  double y;
  if (p == 1) {
    y = x;
  } else if (p == 2) {
    y = x <= mid ? (1/mid)*(x*x) : 1 - (1/(1 - mid))*((1 - x)*(1 - x));
  } else {
    y = impedance_power(x, mid, p);
  }
  return d0 + y*(d1 - d0);
}

// write one constraint row to the optional efc outputs; returns its row index
MJB_HD inline int emit_row(Ctx& c, int row, int type, int id, double pos, double margin, double D,
                           double R, double vel, double aref, double force, int state, double imp) {
  if (c.out.efc_int) {
    if (row < c.njmax) {
      const size_t N = (size_t)c.N;
      int* ei = c.out.efc_int + c.s;
      double* en = c.out.efc_num + c.s;
      ei[(size_t)(3*row + 0)*N] = type;
      ei[(size_t)(3*row + 1)*N] = id;
      ei[(size_t)(3*row + 2)*N] = state;
      en[(size_t)(8*row + 0)*N] = pos;
      en[(size_t)(8*row + 1)*N] = margin;
      en[(size_t)(8*row + 2)*N] = D;
      en[(size_t)(8*row + 3)*N] = R;
      en[(size_t)(8*row + 4)*N] = vel;
      en[(size_t)(8*row + 5)*N] = aref;
      en[(size_t)(8*row + 6)*N] = force;
      en[(size_t)(8*row + 7)*N] = R*imp/(1 - imp);   // efc_diagApprox after mj_makeImpedance:1605
    } else {
      c.status |= kStatusCnstrFull;
    }
  }
  return row;
}

// one scalar row of type friction / limit: returns the constraint force.
//   pos, margin -> impedance ; dA = diagApprox ; vel = J*qvel ; jacc = J*qacc
// mj_makeImpedance (engine_core_constraint.c:1494-1608), mj_referenceConstraint (:2362),
// mj_invConstraint (engine_inverse.c:169) and mj_constraintUpdate (:2387-2457) for this row.
MJB_HD inline double scalar_row(Ctx& c, int row, int type, int id, const double* sp, double pos,
                                double margin, double dA, double vel, double jacc, double floss) {
  const double imp = impedance(sp, pos, margin);
  const double R = fmax(MJB_MINVAL, (1 - imp)*dA/imp);
  const double D = 1 / R;
  const bool isfriction = (type == MJB_CNSTR_FRICTION_DOF || type == MJB_CNSTR_FRICTION_TENDON);
  const double K = isfriction ? 0.0 : sp[MJB_SP_K];
  const double aref = -sp[MJB_SP_B]*vel - K*imp*(pos - margin);
  const double jar = jacc - aref;
  double force = -D*jar;
  int state = MJB_STATE_QUADRATIC;
  if (isfriction) {
    if (jar <= -R*floss) { force = floss; state = MJB_STATE_LINEARNEG; }
    else if (jar >= R*floss) { force = -floss; state = MJB_STATE_LINEARPOS; }
  } else if (type != MJB_CNSTR_EQUALITY) {
    if (jar >= 0) { force = 0; state = MJB_STATE_SATISFIED; }
  }
  emit_row(c, row, type, id, pos, margin, D, R, vel, aref, force, state, imp);
  return force;
}

// ------------------------------------------------------------------------------------------
// mj_instantiateEquality (engine_core_constraint.c:493-763): connect, weld, joint and (fixed)
// tendon couplings, evaluated without forming the Jacobian.
//   connect / weld translation rows: J = jacp(body0, pos0) - jacp(body1, pos1), world axes
//   weld rotation rows: J = torquescale * 0.5 * vec( neg(q1) (0, jacr0 - jacr1) q0 relpose ), a
//     linear map L of the angular-velocity difference; J*v = L(w0 - w1), J'f = torque L'f
//   joint / tendon rows: single dofs (tendons: their joint list) with the polynomial's derivative

// spatial motion of body b at world point p from a carrier array (cvel or cacc_lin)
MJB_HD inline void point_motion(Ctx& c, const double* carrier, int b, const double* p, double* lin,
                                double* ang) {
  const int* rootid = MI(body_rootid);
  double* com = SC(origin);
  double v[6], o[3], r[3], cr[3];
  ldn(v, carrier, 6*b, 6); ldn(o, com, 3*rootid[b], 3);
  r[0] = p[0] - o[0]; r[1] = p[1] - o[1]; r[2] = p[2] - o[2];
  cross3(cr, v, r);
  for (int k = 0; k < 3; k++) { lin[k] = v[3 + k] + cr[k]; ang[k] = v[k]; }
}

// wrench [ (p - O_b) x F + T ; F ] on body b: added to cfrc_ext (positive side) or to cfrc_ext1
// (negative side; subtracted in the backward pass). Two accumulators keep every running sum in
// contact order whatever the interleaving of the two sides (see the pooled contact kernel).
MJB_HD inline void add_wrench(Ctx& c, int b, const double* p, const double* F, const double* T,
                              bool positive) {
  const int* rootid = MI(body_rootid);
  double* com = SC(origin);
  double* fe = positive ? SC(cfrc_ext) : SC(cfrc_ext1);
  double o[3], r[3], cr[3];
  ldn(o, com, 3*rootid[b], 3);
  r[0] = p[0] - o[0]; r[1] = p[1] - o[1]; r[2] = p[2] - o[2];
  cross3(cr, r, F);
  // one batched read-modify-write (loads first, stores last): a single memory round trip; the
  // first wrench on this (body, side) of the state starts from zero (wrench-accumulator masks)
  double w[6] = {0, 0, 0, 0, 0, 0};
  if (wmask_test_and_set(c, b, positive)) ldn(w, fe, 6*b, 6);
  for (int k = 0; k < 3; k++) { w[k] += cr[k] + T[k]; w[3 + k] += F[k]; }
  stn(fe, 6*b, w, 6);
}

// ------------------------------------------------------------------------------------------
// Spatial tendons (mj_tendon, engine_core_smooth.c:726-856; mju_wrap and its 2D helpers,
// engine_util_misc.c:34-420). The path is a sequence of sites, optionally wrapping around a sphere
// or cylinder between two sites, with pulleys scaling the branches. The reference builds the row
// ten_J with mj_jacDifPair for every straight segment whose end points sit on different bodies;
// here the segment itself is handed to a callback, which turns it into J*qvel / J*qacc (relative
// point motion along the segment) or into J'*f (opposite forces along the segment on the two
// bodies) without forming the row.

MJB_DI double norm2(const double* v) { return sqrt(v[0]*v[0] + v[1]*v[1]); }
MJB_DI double normalize2(double* v) {                    // mju_normalize with n = 2
  const double norm = sqrt(v[0]*v[0] + v[1]*v[1]);
  if (norm < MJB_MINVAL) { v[0] = 1; v[1] = 0; }
  else { const double inv = 1/norm; v[0] *= inv; v[1] *= inv; }
  return norm;
}

// do the 2D segments p1-p2 and p3-p4 intersect (engine_util_misc.c:34)
MJB_DI bool wrap_intersect(const double* p1, const double* p2, const double* p3, const double* p4) {
  const double det = (p4[1]-p3[1])*(p2[0]-p1[0]) - (p4[0]-p3[0])*(p2[1]-p1[1]);
  if (fabs(det) < MJB_MINVAL) return false;
  const double a = ((p4[0]-p3[0])*(p1[1]-p3[1]) - (p4[1]-p3[1])*(p1[0]-p3[0])) / det;
  const double b = ((p2[0]-p1[0])*(p1[1]-p3[1]) - (p2[1]-p1[1])*(p1[0]-p3[0])) / det;
  return a >= 0 && a <= 1 && b >= 0 && b <= 1;
}

// arc length between two points of the circle (:54)
MJB_DI double wrap_arc(const double* p0, const double* p1, int ind, double radius) {
  double p0n[2] = {p0[0], p0[1]}, p1n[2] = {p1[0], p1[1]};
  normalize2(p0n); normalize2(p1n);
  double angle = acos(p0n[0]*p1n[0] + p0n[1]*p1n[1]);
  const double cross = p0[1]*p1[0] - p0[0]*p1[1];
  if ((cross > 0 && ind) || (cross < 0 && !ind)) angle = 2*MJB_PI - angle;
  return radius*angle;
}

// 2D wrap around a circle centred at the origin (:79-153): tangent points in pnt, arc length or -1
MJB_HD inline double wrap_circle(double* pnt, const double* end, const double* side, double radius) {
  const double sqlen0 = end[0]*end[0] + end[1]*end[1];
  const double sqlen1 = end[2]*end[2] + end[3]*end[3];
  const double sqrad = radius*radius;
  if (sqlen0 < sqrad || sqlen1 < sqrad || radius < MJB_MINVAL) return -1;
  const double dif[2] = {end[2] - end[0], end[3] - end[1]};
  const double dd = dif[0]*dif[0] + dif[1]*dif[1];
  if (dd < MJB_MINVAL) return -1;
  double a = -(dif[0]*end[0] + dif[1]*end[1])/dd;
  if (a < 0) a = 0; else if (a > 1) a = 1;
  double tmp[2] = {a*dif[0] + end[0], a*dif[1] + end[1]};
  if (tmp[0]*tmp[0] + tmp[1]*tmp[1] > sqrad && (!side || side[0]*tmp[0] + side[1]*tmp[1] >= 0)) return -1;
  double sol[2][2][2], good[2];
  for (int i = 0; i < 2; i++) {
    const double sqrt0 = sqrt(sqlen0 - sqrad), sqrt1 = sqrt(sqlen1 - sqrad);
    const int sgn = i == 0 ? 1 : -1;
    sol[i][0][0] = (end[0]*sqrad + sgn*radius*end[1]*sqrt0)/sqlen0;
    sol[i][0][1] = (end[1]*sqrad - sgn*radius*end[0]*sqrt0)/sqlen0;
    sol[i][1][0] = (end[2]*sqrad - sgn*radius*end[3]*sqrt1)/sqlen1;
    sol[i][1][1] = (end[3]*sqrad + sgn*radius*end[2]*sqrt1)/sqlen1;
    if (side) {
      tmp[0] = sol[i][0][0] + sol[i][1][0]; tmp[1] = sol[i][0][1] + sol[i][1][1];
      normalize2(tmp);
      good[i] = tmp[0]*side[0] + tmp[1]*side[1];
    } else {
      tmp[0] = sol[i][0][0] - sol[i][1][0]; tmp[1] = sol[i][0][1] - sol[i][1][1];
      good[i] = -(tmp[0]*tmp[0] + tmp[1]*tmp[1]);
    }
    if (wrap_intersect(end, sol[i][0], end + 2, sol[i][1])) good[i] = -10000;
  }
  const int i = good[0] > good[1] ? 0 : 1;
  pnt[0] = sol[i][0][0]; pnt[1] = sol[i][0][1]; pnt[2] = sol[i][1][0]; pnt[3] = sol[i][1][1];
  if (wrap_intersect(end, pnt, end + 2, pnt + 2)) return -1;
  return wrap_arc(sol[i][0], sol[i][1], i, radius);
}

// 2D wrap on the inside of the circle (:160-283): one touching point (both pnt pairs), 0 or -1
MJB_HD inline double wrap_inside(double* pnt, const double* end, double radius) {
  const int maxiter = 20;
  const double zinit = 1 - 1e-7, tolerance = 1e-6;
  const double len0 = norm2(end), len1 = norm2(end + 2);
  const double dif[2] = {end[2] - end[0], end[3] - end[1]};
  const double dd = dif[0]*dif[0] + dif[1]*dif[1];
  if (len0 <= radius || len1 <= radius || radius < MJB_MINVAL || len0 < MJB_MINVAL || len1 < MJB_MINVAL) return -1;
  if (dd > MJB_MINVAL) {
    const double a = -(dif[0]*end[0] + dif[1]*end[1]) / dd;
    if (a > 0 && a < 1) {
      const double tmp[2] = {end[0] + dif[0]*a, end[1] + dif[1]*a};
      if (norm2(tmp) <= radius) return -1;
    }
  }
  pnt[0] = 0.5*(end[0] + end[2]); pnt[1] = 0.5*(end[1] + end[3]);
  normalize2(pnt);
  pnt[0] *= radius; pnt[1] *= radius;
  pnt[2] = pnt[0]; pnt[3] = pnt[1];
  const double A = radius/len0, B = radius/len1;
  const double cosG = (len0*len0 + len1*len1 - dd) / (2*len0*len1);
  if (cosG < -1 + MJB_MINVAL) return -1;
  else if (cosG > 1 - MJB_MINVAL) return 0;
  const double G = acos(cosG);
  double z = zinit;
  double f = asin(A*z) + asin(B*z) - 2*asin(z) + G;
  if (f > 0) return 0;
  int iter;
  for (iter = 0; iter < maxiter && fabs(f) > tolerance; iter++) {
    const double df = A/fmax(MJB_MINVAL, sqrt(1 - z*z*A*A)) + B/fmax(MJB_MINVAL, sqrt(1 - z*z*B*B)) -
                      2/fmax(MJB_MINVAL, sqrt(1 - z*z));
    if (df > -MJB_MINVAL) return 0;
    const double z1 = z - f/df;
    if (z1 > z) return 0;
    z = z1;
    f = asin(A*z) + asin(B*z) - 2*asin(z) + G;
    if (f > tolerance) return 0;
  }
  if (iter >= maxiter) return 0;
  double vec[2], ang;
  if (end[0]*end[3] - end[1]*end[2] > 0) { vec[0] = end[0]; vec[1] = end[1]; ang = asin(z) - asin(A*z); }
  else { vec[0] = end[2]; vec[1] = end[3]; ang = asin(z) - asin(B*z); }
  normalize2(vec);
  pnt[0] = radius*(cos(ang)*vec[0] - sin(ang)*vec[1]);
  pnt[1] = radius*(sin(ang)*vec[0] + cos(ang)*vec[1]);
  pnt[2] = pnt[0]; pnt[3] = pnt[1];
  return 0;
}

// mju_wrap (:293-420): wrap the segment x0-x1 around a sphere (cylinder = false) or a cylinder;
// returns the arc length and the two 3D tangent points in wpnt, or -1 when the straight segment is kept
MJB_HD inline double wrap_geom(double* wpnt, const double* x0, const double* x1, const double* xpos,
                               const double* xmat, double radius, bool cylinder, const double* side) {
  double p[2][3], t[3];
  t[0] = x0[0] - xpos[0]; t[1] = x0[1] - xpos[1]; t[2] = x0[2] - xpos[2];
  for (int i = 0; i < 3; i++) p[0][i] = xmat[i]*t[0] + xmat[3 + i]*t[1] + xmat[6 + i]*t[2];
  t[0] = x1[0] - xpos[0]; t[1] = x1[1] - xpos[1]; t[2] = x1[2] - xpos[2];
  for (int i = 0; i < 3; i++) p[1][i] = xmat[i]*t[0] + xmat[3 + i]*t[1] + xmat[6 + i]*t[2];
  if (sqrt(dot3(p[0], p[0])) < MJB_MINVAL || sqrt(dot3(p[1], p[1])) < MJB_MINVAL) return -1;
  double axis[2][3];
  if (!cylinder) {
    axis[0][0] = p[0][0]; axis[0][1] = p[0][1]; axis[0][2] = p[0][2];
    normalize3(axis[0]);
    double normal[3];
    cross3(normal, p[0], p[1]);
    const double nrm = normalize3(normal);
    if (nrm < MJB_MINVAL) {
      int i = 0;
      if (fabs(axis[0][1]) > fabs(axis[0][0]) && fabs(axis[0][1]) > fabs(axis[0][2])) i = 1;
      if (fabs(axis[0][2]) > fabs(axis[0][0]) && fabs(axis[0][2]) > fabs(axis[0][1])) i = 2;
      axis[1][0] = 1; axis[1][1] = 1; axis[1][2] = 1;
      axis[1][i] = 0;
      cross3(normal, axis[0], axis[1]);
      normalize3(normal);
    }
    cross3(axis[1], normal, axis[0]);
    normalize3(axis[1]);
  } else {
    axis[0][0] = 1; axis[0][1] = 0; axis[0][2] = 0;
    axis[1][0] = 0; axis[1][1] = 1; axis[1][2] = 0;
  }
  double s[3] = {0, 0, 0}, d[4], sd[2] = {0, 0};
  d[0] = dot3(p[0], axis[0]); d[1] = dot3(p[0], axis[1]);
  d[2] = dot3(p[1], axis[0]); d[3] = dot3(p[1], axis[1]);
  if (side) {
    t[0] = side[0] - xpos[0]; t[1] = side[1] - xpos[1]; t[2] = side[2] - xpos[2];
    for (int i = 0; i < 3; i++) s[i] = xmat[i]*t[0] + xmat[3 + i]*t[1] + xmat[6 + i]*t[2];
    sd[0] = dot3(s, axis[0]); sd[1] = dot3(s, axis[1]);
    normalize2(sd);
    sd[0] *= radius; sd[1] *= radius;
  }
  double wlen, pnt[4];
  if (side && sqrt(dot3(s, s)) < radius) wlen = wrap_inside(pnt, d, radius);
  else wlen = wrap_circle(pnt, d, side ? sd : (const double*)0, radius);
  if (wlen < 0) return -1;
  double res[6];
  for (int i = 0; i < 2; i++) {
    for (int k = 0; k < 3; k++) res[3*i + k] = axis[0][k]*pnt[2*i];
    for (int k = 0; k < 3; k++) res[3*i + k] += axis[1][k]*pnt[2*i + 1];
  }
  if (cylinder) {
    const double L0 = sqrt((p[0][0]-res[0])*(p[0][0]-res[0]) + (p[0][1]-res[1])*(p[0][1]-res[1]));
    const double L1 = sqrt((p[1][0]-res[3])*(p[1][0]-res[3]) + (p[1][1]-res[4])*(p[1][1]-res[4]));
    res[2] = p[0][2] + (p[1][2] - p[0][2])*L0 / (L0 + wlen + L1);
    res[5] = p[0][2] + (p[1][2] - p[0][2])*(L0 + wlen) / (L0 + wlen + L1);
    const double height = fabs(res[5] - res[2]);
    wlen = sqrt(wlen*wlen + height*height);
  }
  mulMatVec3(wpnt, xmat, res);
  mulMatVec3(wpnt + 3, xmat, res + 3);
  for (int k = 0; k < 3; k++) { wpnt[k] += xpos[k]; wpnt[3 + k] += xpos[k]; }
  return wlen;
}

// world position of a site (mj_local2Global for sites, engine_core_smooth.c:172-177)
MJB_HD inline void site_world_pos(Ctx& c, int sid, double* out) {
  const int b = MI(site_bodyid)[sid];
  const int sf = MI(site_sameframe)[sid];
  double bp[3], bq[4], bm[9];
  ldn(bp, SC(xpos), 3*b, 3);
  if (sf == MJB_SAMEFRAME_BODY) { out[0] = bp[0]; out[1] = bp[1]; out[2] = bp[2]; return; }
  ldn(bq, SC(xquat), 4*b, 4);
  quat2Mat(bm, bq);
  if (sf == MJB_SAMEFRAME_INERTIA) {
    mulMatVec3(out, bm, MD(body_ipos) + 3*b);      // == xipos of the body
  } else {
    mulMatVec3(out, bm, MD(site_pos) + 3*sid);
  }
  out[0] += bp[0]; out[1] += bp[1]; out[2] += bp[2];
}

// Walk the path of spatial tendon t (engine_core_smooth.c:726-856). Returns its length; calls
// seg(body_a, point_a, body_b, point_b, dir, divisor) for every straight segment between different
// bodies, dir = unit vector from a to b.
template <typename F>
MJB_HD inline double spatial_tendon_walk(Ctx& c, int t, F seg) {
  const int* wrap_type = MI(wrap_type); const int* wrap_objid = MI(wrap_objid);
  const double* wrap_prm = MD(wrap_prm);
  const int* site_bodyid = MI(site_bodyid); const int* geom_bodyid = MI(geom_bodyid);
  const double* geom_size = MD(geom_size);
  const int adr = MI(tendon_adr)[t], num = MI(tendon_num)[t];
  double divisor = 1, L = 0;
  int j = 0;
  while (j < num - 1) {
    int type0 = wrap_type[adr + j], type1 = wrap_type[adr + j + 1];
    int id0 = wrap_objid[adr + j], id1 = wrap_objid[adr + j + 1];
    if (type0 == MJB_WRAP_PULLEY || type1 == MJB_WRAP_PULLEY) {
      if (type0 == MJB_WRAP_PULLEY) divisor = wrap_prm[adr + j];
      j++;
      continue;
    }
    double wlen = -1, wpnt[12];
    int wbody[4], wrapid = -1;
    bool wrapping = false;
    site_world_pos(c, id0, wpnt);
    wbody[0] = site_bodyid[id0];
    if (type1 == MJB_WRAP_SPHERE || type1 == MJB_WRAP_CYLINDER) {
      const bool cylinder = type1 == MJB_WRAP_CYLINDER;
      wrapping = true;
      wrapid = id1;
      id1 = wrap_objid[adr + j + 2];
      const double prm = wrap_prm[adr + j + 1];
      const int sideid = (int)(prm + (prm > 0 ? 0.5 : -0.5));      // mju_round
      double x1[3], gp[3], gm[9], sidep[3];
      site_world_pos(c, id1, x1);
      load_geom_pos(c, wrapid, gp); ldn(gm, SC(geom_xmat), 9*wrapid, 9);
      if (sideid >= 0) site_world_pos(c, sideid, sidep);
      wlen = wrap_geom(wpnt + 3, wpnt, x1, gp, gm, geom_size[3*wrapid], cylinder,
                       sideid >= 0 ? sidep : (const double*)0);
    }
    if (wlen < 0) {
      site_world_pos(c, id1, wpnt + 3);
      wbody[1] = site_bodyid[id1];
      const double d[3] = {wpnt[0] - wpnt[3], wpnt[1] - wpnt[4], wpnt[2] - wpnt[5]};
      L += sqrt(d[0]*d[0] + d[1]*d[1] + d[2]*d[2]) / divisor;
    } else {
      site_world_pos(c, id1, wpnt + 9);
      wbody[1] = wbody[2] = geom_bodyid[wrapid];
      wbody[3] = site_bodyid[id1];
      const double d0[3] = {wpnt[0] - wpnt[3], wpnt[1] - wpnt[4], wpnt[2] - wpnt[5]};
      const double d1[3] = {wpnt[6] - wpnt[9], wpnt[7] - wpnt[10], wpnt[8] - wpnt[11]};
      L += (sqrt(d0[0]*d0[0] + d0[1]*d0[1] + d0[2]*d0[2]) + wlen +
            sqrt(d1[0]*d1[0] + d1[1]*d1[1] + d1[2]*d1[2])) / divisor;
    }
    for (int k = 0; k < (wlen < 0 ? 1 : 3); k++) {
      if (wbody[k] != wbody[k + 1]) {
        double dif[3] = {wpnt[3*k + 3] - wpnt[3*k], wpnt[3*k + 4] - wpnt[3*k + 1], wpnt[3*k + 5] - wpnt[3*k + 2]};
        normalize3(dif);
        seg(wbody[k], wpnt + 3*k, wbody[k + 1], wpnt + 3*k + 3, dif, divisor);
      }
    }
    j += wrapping ? 2 : 1;
  }
  return L;
}

// J*qvel and J*qacc of spatial tendon t from the body carriers
MJB_HD inline double spatial_tendon_kinematics(Ctx& c, int t, double* vel, double* acc) {
  double v = 0, a = 0;
  const double L = spatial_tendon_walk(c, t, [&](int ba, const double* pa, int bb, const double* pb,
                                                 const double* dif, double divisor) {
    double la[3], lb[3], ang[3];
    point_motion(c, SC(cvel), ba, pa, la, ang);
    point_motion(c, SC(cvel), bb, pb, lb, ang);
    const double dv[3] = {lb[0] - la[0], lb[1] - la[1], lb[2] - la[2]};
    v += dot3(dif, dv) / divisor;
    point_motion(c, SC(cacc_lin), ba, pa, la, ang);
    point_motion(c, SC(cacc_lin), bb, pb, lb, ang);
    const double da[3] = {lb[0] - la[0], lb[1] - la[1], lb[2] - la[2]};
    a += dot3(dif, da) / divisor;
  });
  *vel = v; *acc = a;
  return L;
}

// wrench [ (p - O_b) x F ; F ] on body b, added (sign +1) or subtracted (-1) in a carrier array
// masked: the carrier is the '+' constraint-wrench accumulator, whose rows are valid only under
// the state's wrench mask (the passive carrier is initialised for every body by the forward sweep)
MJB_HD inline void add_force_to(Ctx& c, double* carrier, int b, const double* p, const double* F, double sign,
                                bool masked) {
  if (MI(body_static)[b]) return;
  double o[3], r[3], cr[3], w[6] = {0, 0, 0, 0, 0, 0};
  ldn(o, SC(origin), 3*MI(body_rootid)[b], 3);
  r[0] = p[0] - o[0]; r[1] = p[1] - o[1]; r[2] = p[2] - o[2];
  cross3(cr, r, F);
  if (!masked || wmask_test_and_set(c, b, true)) ldn(w, carrier, 6*b, 6);
  for (int k = 0; k < 3; k++) { w[k] += sign*cr[k]; w[3 + k] += sign*F[k]; }
  stn(carrier, 6*b, w, 6);
}

// J'*f of spatial tendon t: constraint forces go to the constraint-wrench carrier, passive forces
// (spring, damper) to the passive-wrench carrier that the backward sweep projects into qfrc_passive
MJB_HD inline void spatial_tendon_apply(Ctx& c, int t, double f, bool passive) {
  double* carrier = passive ? SC(cfrc_gc) : SC(cfrc_ext);
  spatial_tendon_walk(c, t, [&](int ba, const double* pa, int bb, const double* pb, const double* dif,
                                double divisor) {
    const double s = f / divisor;
    const double F[3] = {dif[0]*s, dif[1]*s, dif[2]*s};
    add_force_to(c, carrier, bb, pb, F, 1.0, !passive);
    add_force_to(c, carrier, ba, pa, F, -1.0, !passive);
  });
}

// ------------------------------------------------------------------------------------------
// mj_tendon (engine_core_smooth.c:651-860) without the Jacobian rows: length, J*qvel
// (ten_velocity, engine_forward.c:205-210) and J*qacc of every tendon. Fixed tendons: coefficients
// on scalar joints (:699-723); spatial tendons: the path walk above.
template <bool kSpatial>
MJB_HD inline void tendon_kinematics(Ctx& c) {
  const mjbHdr& H = *c.H;
  if (!H.ntendon) return;
  double* L = SC(ten_length); double* V = SC(ten_velocity); double* A = SC(ten_acc);
  const int* tendon_adr = MI(tendon_adr); const int* tendon_num = MI(tendon_num);
  const int* wrap_objid = MI(wrap_objid); const int* wrap_type = MI(wrap_type);
  const int* jnt_qposadr = MI(jnt_qposadr); const int* jnt_dofadr = MI(jnt_dofadr);
  const int* tendon_active = MI(tendon_active);
  const double* wrap_prm = MD(wrap_prm);
  MJB_UNROLL
  for (int t = 0; t < H.ntendon; t++) {
    const int adr = tendon_adr[t], num = tendon_num[t];
    double len = 0, vel = 0, acc = 0;
    if (wrap_type[adr] == MJB_WRAP_JOINT) {
      MJB_UNROLL
      for (int j = 0; j < num; j++) {
        const int k = wrap_objid[adr + j];
        len += wrap_prm[adr + j] * QPOS(jnt_qposadr[k]);
        vel += wrap_prm[adr + j] * QVEL(jnt_dofadr[k]);
        acc += wrap_prm[adr + j] * QACC(jnt_dofadr[k]);
      }
    } else if (kSpatial && tendon_active[t]) {
      // a spatial tendon that carries no force is output-only in the reference too: skipped
      // (kSpatial: the path walk is compiled only into the kernel instantiation that needs it)
      len = spatial_tendon_kinematics(c, t, &vel, &acc);
    }
    AT(L, t) = len; AT(V, t) = vel; AT(A, t) = acc;
  }
}

// J'*f of tendon t into the joint-space accumulator (fixed) or the body-wrench carriers (spatial)
template <bool kSpatial>
MJB_HD inline void tendon_apply(Ctx& c, int t, double f, double* qdst, bool passive) {
  const int adr = MI(tendon_adr)[t], num = MI(tendon_num)[t];
  if (MI(wrap_type)[adr] == MJB_WRAP_JOINT) {
    const int* wrap_objid = MI(wrap_objid); const int* jnt_dofadr = MI(jnt_dofadr);
    const double* wrap_prm = MD(wrap_prm);
    MJB_UNROLL
    for (int j = 0; j < num; j++) AT(qdst, jnt_dofadr[wrap_objid[adr + j]]) += wrap_prm[adr + j]*f;
  } else if (kSpatial) {
    spatial_tendon_apply(c, t, f, passive);
  }
}

// ------------------------------------------------------------------------------------------
// mj_passive (engine_passive.c:57-379,436-497): joint springs and dof dampers are evaluated per dof
// inside the forward sweep (scalar_dof_forces / quat_dof_forces); this adds the tendon
// spring-dampers; gravity compensation is a body wrench handled by the sweeps. Fluid, flex,
// callbacks and plugins are rejected at upload.
template <bool kSpatial>
MJB_HD inline void passive_tendons(Ctx& c) {
  const mjbHdr& H = *c.H;
  if ((H.disableflags & MJB_DSBL_PASSIVE) || !H.ntendon) return;
  const double* stiff = MD(tendon_stiffness); const double* damp = MD(tendon_damping);
  const double* ls = MD(tendon_lengthspring);
  double* L = SC(ten_length); double* V = SC(ten_velocity);
  MJB_UNROLL
  for (int t = 0; t < H.ntendon; t++) {
    const double ks = stiff[t], kd = damp[t];
    if (ks == 0 && kd == 0) continue;
    const double len = AT(L, t), lower = ls[2*t], upper = ls[2*t+1];
    double fs = 0;
    if (len > upper) fs = ks*(upper - len);
    else if (len < lower) fs = ks*(lower - len);
    const double fd = -kd*AT(V, t);
    // spring and damper are accumulated separately in the reference, then added
    tendon_apply<kSpatial>(c, t, fs + fd, SC(qfrc_passive), true);
  }
}

// 0.5 * vec( quat1 * (0, a) * quat )   (engine_core_constraint.c:617-635)
MJB_DI void weld_rot_map(double* res, const double* quat1, const double* quat, const double* a) {
  double q2[4] = {-quat1[1]*a[0] - quat1[2]*a[1] - quat1[3]*a[2],
                  quat1[0]*a[0] + quat1[2]*a[2] - quat1[3]*a[1],
                  quat1[0]*a[1] + quat1[3]*a[0] - quat1[1]*a[2],
                  quat1[0]*a[2] + quat1[1]*a[1] - quat1[2]*a[0]};   // mju_mulQuatAxis
  double q3[4];
  mulQuat(q3, q2, quat);
  res[0] = 0.5*q3[1]; res[1] = 0.5*q3[2]; res[2] = 0.5*q3[3];
}

// Equality constraint i enabled for this state: d->eq_active[i] when the caller has set per-state flags
// (mjb_setEqActive), else the model's eq_active0 (what mj_makeData / mj_resetData leave in mjData).
MJB_HD inline bool eq_enabled(Ctx& c, const int* ei, int i) {
  return c.out.eq_active ? c.out.eq_active[(size_t)i*(size_t)c.N + c.s] != 0 : ei[MJB_EQI_ACTIVE] != 0;
}
// Number of equality rows of this state = first friction-loss row (mj_makeConstraint order). A model
// constant (H.ne_rows, folded into the specialised kernels) unless per-state flags are set.
MJB_HD inline int ne_base(Ctx& c) {
  const mjbHdr& H = *c.H;
  if (H.neq == 0 || !c.out.eq_active) return H.ne_rows;
  if ((H.disableflags & (MJB_DSBL_EQUALITY | MJB_DSBL_CONSTRAINT))) return 0;
  const int* eq_int = MI(eq_int);
  int n = 0;
  for (int i = 0; i < H.neq; i++) {
    const int* ei = eq_int + MJB_EQ_NI*i;
    if (!eq_enabled(c, ei, i)) continue;
    const int type = ei[MJB_EQI_TYPE];
    if (type == 0 || type == 1) { if (!ei[MJB_EQI_SKIP]) n += type == 0 ? 3 : 6; }
    else n++;
  }
  return n;
}

template <bool kSpatial>
MJB_HD inline void equality_rows(Ctx& c) {
  const mjbHdr& H = *c.H;
  if (H.neq == 0 || (H.disableflags & MJB_DSBL_EQUALITY)) return;
  const int* eq_int = MI(eq_int);
  const double* eq_num = MD(eq_num);
  const double* eq_data = MD(eq_data);
  const double* sp_eq = MD(sp_eq);
  double* xpos = SC(xpos); double* xquat = SC(xquat);
  double* qc = SC(qfrc_c);
  for (int i = 0; i < H.neq; i++) {
    const int* ei = eq_int + MJB_EQ_NI*i;
    const double* en = eq_num + MJB_EQ_NN*i;
    const double* sp = sp_eq + MJB_SP_N*i;
    if (!eq_enabled(c, ei, i)) continue;
    const int type = ei[MJB_EQI_TYPE];
    if (type == 0 || type == 1) {
      if (ei[MJB_EQI_SKIP]) continue;
      const int b0 = ei[MJB_EQI_B0], b1 = ei[MJB_EQI_B1];
      double pos0[3], pos1[3], m0[9], m1[9], p[3], bq0[4], bq1[4];
      ldn(bq0, xquat, 4*b0, 4); ldn(bq1, xquat, 4*b1, 4);
      quat2Mat(m0, bq0); quat2Mat(m1, bq1);            // == xmat of the two bodies
      mulMatVec3(pos0, m0, en + MJB_EQN_ANCHOR0); ldn(p, xpos, 3*b0, 3);
      pos0[0] += p[0]; pos0[1] += p[1]; pos0[2] += p[2];
      mulMatVec3(pos1, m1, en + MJB_EQN_ANCHOR1); ldn(p, xpos, 3*b1, 3);
      pos1[0] += p[0]; pos1[1] += p[1]; pos1[2] += p[2];
      const int nrow = type == 0 ? 3 : 6;
      double cpos[6], vel[6], acc[6], l0[3], a0[3], l1[3], a1[3];
      for (int k = 0; k < 3; k++) cpos[k] = pos0[k] - pos1[k];
      point_motion(c, SC(cvel), b0, pos0, l0, a0);
      point_motion(c, SC(cvel), b1, pos1, l1, a1);
      double wv[3] = {a0[0] - a1[0], a0[1] - a1[1], a0[2] - a1[2]};
      for (int k = 0; k < 3; k++) vel[k] = l0[k] - l1[k];
      point_motion(c, SC(cacc_lin), b0, pos0, l0, a0);
      point_motion(c, SC(cacc_lin), b1, pos1, l1, a1);
      double wa[3] = {a0[0] - a1[0], a0[1] - a1[1], a0[2] - a1[2]};
      for (int k = 0; k < 3; k++) acc[k] = l0[k] - l1[k];
      double quat[4] = {1, 0, 0, 0}, quat1[4] = {1, 0, 0, 0};
      const double ts = en[MJB_EQN_TORQUESCALE];
      if (type == 1) {
        double t[4], q2[4];
        const double* q0 = bq0; const double* q1 = bq1;
        mulQuat(quat, q0, en + MJB_EQN_Q0);            // q0 * relpose   (or body0 * site_quat0)
        if (ei[MJB_EQI_SITE]) { mulQuat(t, q1, en + MJB_EQN_Q1); }
        else { t[0] = q1[0]; t[1] = q1[1]; t[2] = q1[2]; t[3] = q1[3]; }
        quat1[0] = t[0]; quat1[1] = -t[1]; quat1[2] = -t[2]; quat1[3] = -t[3];
        mulQuat(q2, quat1, quat);
        cpos[3] = q2[1]*ts; cpos[4] = q2[2]*ts; cpos[5] = q2[3]*ts;
        double r[3];
        weld_rot_map(r, quat1, quat, wv);
        vel[3] = r[0]*ts; vel[4] = r[1]*ts; vel[5] = r[2]*ts;
        weld_rot_map(r, quat1, quat, wa);
        acc[3] = r[0]*ts; acc[4] = r[1]*ts; acc[5] = r[2]*ts;
      }
      // getposdim (:1392-1425): all rows share the impedance of the norm of the residual
      double nn = 0;
      for (int k = 0; k < nrow; k++) nn += cpos[k]*cpos[k];
      const double imp = impedance(sp, sqrt(nn), 0);
      double f[6] = {0, 0, 0, 0, 0, 0};
      for (int r = 0; r < nrow; r++) {
        const double dA = r < 3 ? en[MJB_EQN_DA_TRAN] : en[MJB_EQN_DA_ROT];
        const double R = fmax(MJB_MINVAL, (1 - imp)*dA/imp);
        const double D = 1/R;
        const double aref = -sp[MJB_SP_B]*vel[r] - sp[MJB_SP_K]*imp*cpos[r];
        const double jar = acc[r] - aref;
        f[r] = -D*jar;
        emit_row(c, c.ne + r, MJB_CNSTR_EQUALITY, i, cpos[r], 0, D, R, vel[r], aref, f[r], MJB_STATE_QUADRATIC, imp);
      }
      c.ne += nrow;
      double T[3] = {0, 0, 0};
      if (type == 1) {
        // torque = L' f_rot, with the columns of L obtained by mapping the unit vectors
        double fr[3] = {f[3]*ts, f[4]*ts, f[5]*ts};
        for (int k = 0; k < 3; k++) {
          double e[3] = {k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0}, col[3];
          weld_rot_map(col, quat1, quat, e);
          T[k] = col[0]*fr[0] + col[1]*fr[1] + col[2]*fr[2];
        }
        // mj_rnePostConstraint reports the RAW rotational row forces as the weld's torque
        // (engine_core_smooth.c:2092-2095), not J'f: keep the difference for the cfrc_ext output
        if (c.out.cfrc_ext) {
          const double dT[3] = {f[3] - T[0], f[4] - T[1], f[5] - T[2]};
          stn(SC(weld_dt), 3*i, dT, 3);
        }
      }
      add_wrench(c, b0, pos0, f, T, true);
      add_wrench(c, b1, pos1, f, T, false);
    } else {
      // joint / tendon coupling (:640-719)
      const double* data = eq_data + 11*i;
      const int id0 = ei[MJB_EQI_B0], id1 = ei[MJB_EQI_B1];
      const int* jnt_qposadr = MI(jnt_qposadr); const int* jnt_dofadr = MI(jnt_dofadr);
      const int* tendon_adr = MI(tendon_adr); const int* tendon_num = MI(tendon_num);
      const int* wrap_objid = MI(wrap_objid);
      const double* wrap_prm = MD(wrap_prm);
      double pos[2] = {0, 0}, ref[2] = {0, 0}, v[2] = {0, 0}, a[2] = {0, 0};
      for (int j = 0; j < 1 + (id1 >= 0); j++) {
        const int id = j == 0 ? id0 : id1;
        if (type == 2) {
          pos[j] = QPOS(jnt_qposadr[id]); ref[j] = MD(qpos0)[jnt_qposadr[id]];
          v[j] = QVEL(jnt_dofadr[id]); a[j] = QACC(jnt_dofadr[id]);
        } else {
          pos[j] = AT(SC(ten_length), id); ref[j] = MD(tendon_length0)[id];
          if (kSpatial && MI(wrap_type)[tendon_adr[id]] != MJB_WRAP_JOINT) {
            // spatial tendon: J*qvel, J*qacc of its path (tendon_kinematics, relative point motion)
            v[j] = AT(SC(ten_velocity), id); a[j] = AT(SC(ten_acc), id);
            continue;
          }
          for (int w = 0; w < tendon_num[id]; w++) {
            const int dof = jnt_dofadr[wrap_objid[tendon_adr[id] + w]];
            v[j] += wrap_prm[tendon_adr[id] + w]*QVEL(dof);
            a[j] += wrap_prm[tendon_adr[id] + w]*QACC(dof);
          }
        }
      }
      double cpos, deriv = 0;
      if (id1 >= 0) {
        const double dif = pos[1] - ref[1];
        cpos = pos[0] - ref[0] - data[0] -
               (data[1]*dif + data[2]*dif*dif + data[3]*dif*dif*dif + data[4]*dif*dif*dif*dif);
        deriv = data[1] + 2*data[2]*dif + 3*data[3]*dif*dif + 4*data[4]*dif*dif*dif;
      } else {
        cpos = pos[0] - ref[0] - data[0];
      }
      const double vel = v[0] - deriv*v[1], acc = a[0] - deriv*a[1];
      const double f = scalar_row(c, c.ne, MJB_CNSTR_EQUALITY, i, sp, cpos, 0, en[MJB_EQN_DA_TRAN], vel, acc, 0);
      c.ne++;
      for (int j = 0; j < 1 + (id1 >= 0); j++) {
        const int id = j == 0 ? id0 : id1;
        const double fj = j == 0 ? f : -deriv*f;
        if (type == 2) {
          AT(qc, jnt_dofadr[id]) += fj;
        } else if (kSpatial && MI(wrap_type)[tendon_adr[id]] != MJB_WRAP_JOINT) {
          spatial_tendon_apply(c, id, fj, false);      // opposite forces along every segment of the path
        } else {
          for (int w = 0; w < tendon_num[id]; w++) {
            AT(qc, jnt_dofadr[wrap_objid[tendon_adr[id] + w]]) += wrap_prm[tendon_adr[id] + w]*fj;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Per-dof pieces of mj_passive / mj_instantiateFriction / mj_instantiateLimit, called from the
// forward sweep while the dof's qpos/qvel/qacc are in registers. Row numbers are explicit: the
// numbers of equality rows (ne_rows) and friction-loss rows (nf_rows, dofs first then tendons) are
// model constants, so friction row of dof i is ne_rows + dof_frow[i] and limit rows follow at
// ne_rows + nf_rows + (running count), in joint order like the reference.

// mj_checkPos / mj_checkVel / mj_checkAcc (engine_forward.c:53-102): flag, do not reset
MJB_DI int bad_value(double v) { return !(v == v) || v > MJB_MAXVAL || v < -MJB_MAXVAL; }

MJB_HD inline bool rows_enabled(const mjbHdr& H) { return !(H.disableflags & MJB_DSBL_CONSTRAINT); }

// friction-loss row of dof i (engine_core_constraint.c:768-790); returns the row's force
MJB_HD inline double dof_friction_row(Ctx& c, int i, double qv, double qa) {
  const mjbHdr& H = *c.H;
  const int frow = MI(dof_frow)[i];
  if (frow < 0) return 0;
  return scalar_row(c, ne_base(c) + frow, MJB_CNSTR_FRICTION_DOF, i, MD(sp_dof_friction) + MJB_SP_N*i,
                    0, 0, MD(dof_invweight0)[i], qv, qa, MD(dof_frictionloss)[i]);
}

// limit rows of a slide/hinge joint (engine_core_constraint.c:851-871); returns J'f on its dof
MJB_HD inline double joint_limit_rows(Ctx& c, int jid, int dof, double q, double qv, double qa) {
  const mjbHdr& H = *c.H;
  if (!MI(jnt_limited)[jid] || (H.disableflags & MJB_DSBL_LIMIT) || !rows_enabled(H)) return 0;
  const double margin = MD(jnt_margin)[jid];
  const double* range = MD(jnt_range) + 2*jid;
  double acc = 0;
  for (int side = -1; side <= 1; side += 2) {
    const double dist = side * (range[(side + 1)/2] - q);
    if (dist < margin) {
      // J = -side at this dof
      const double f = scalar_row(c, ne_base(c) + H.nf_rows + c.nl, MJB_CNSTR_LIMIT_JOINT, jid,
                                  MD(sp_jnt_limit) + MJB_SP_N*jid, dist, margin,
                                  MD(dof_invweight0)[dof], -side*qv, -side*qa, 0);
      acc += -side*f;
      c.nl++;
    }
  }
  return acc;
}

// everything a scalar (slide/hinge) dof contributes besides the rigid-body terms: input checks,
// spring + damper -> qfrc_passive, friction-loss and limit rows -> qfrc_c
MJB_HD inline void scalar_dof_forces(Ctx& c, int jid, int qadr, int dof, double q, double qv, double qa) {
  const mjbHdr& H = *c.H;
  if (bad_value(q)) c.status |= kStatusBadQpos;
  if (bad_value(qv)) c.status |= kStatusBadQvel;
  if (bad_value(qa)) c.status |= kStatusBadQacc;
  double passive = 0;
  if (!(H.disableflags & MJB_DSBL_PASSIVE)) {
    const double k = MD(jnt_stiffness)[jid], dmp = MD(dof_damping)[dof];
    if (k != 0) passive = -k*(q - MD(qpos_spring)[qadr]);
    if (dmp != 0) passive += -dmp*qv;
  }
  AT(SC(qfrc_passive), dof) = passive;
  double qc = 0;
  if (rows_enabled(H)) {
    qc = dof_friction_row(c, dof, qv, qa);
    qc += joint_limit_rows(c, jid, dof, q, qv, qa);
  }
  AT(SC(qfrc_c), dof) = qc;
}

// the same for the 3 rotational dofs of a ball or free joint whose (normalised) quaternion is quat
// (engine_passive.c:84-98, engine_core_constraint.c:875-918); free joints: ntrans = 3 translations
// in front, handled here too
MJB_HD inline void quat_dof_forces(Ctx& c, int jid, int qadr, int dof, int jt, const double* quat) {
  const mjbHdr& H = *c.H;
  const double k = MD(jnt_stiffness)[jid];
  const bool passive_on = !(H.disableflags & MJB_DSBL_PASSIVE);
  const double* dof_damping = MD(dof_damping);
  double* qp = SC(qfrc_passive); double* qcs = SC(qfrc_c);
  int padr = qadr, d = dof;
  if (jt == MJB_JNT_FREE) {
    for (int r = 0; r < 3; r++) {
      const double q = QPOS(padr + r), qv = QVEL(d + r), qa = QACC(d + r);
      if (bad_value(q)) c.status |= kStatusBadQpos;
      if (bad_value(qv)) c.status |= kStatusBadQvel;
      if (bad_value(qa)) c.status |= kStatusBadQacc;
      double passive = 0;
      if (passive_on) {
        if (k != 0) passive = -k*(q - MD(qpos_spring)[padr + r]);
        const double dmp = dof_damping[d + r];
        if (dmp != 0) passive += -dmp*qv;
      }
      AT(qp, d + r) = passive;
      AT(qcs, d + r) = rows_enabled(H) ? dof_friction_row(c, d + r, qv, qa) : 0.0;
    }
    padr += 3; d += 3;
  }
  double qv[3], qa[3];
  for (int r = 0; r < 3; r++) {
    qv[r] = QVEL(d + r); qa[r] = QACC(d + r);
    if (bad_value(qv[r])) c.status |= kStatusBadQvel;
    if (bad_value(qa[r])) c.status |= kStatusBadQacc;
  }
  for (int r = 0; r < 4; r++) if (bad_value(QPOS(padr + r))) c.status |= kStatusBadQpos;
  double spring[3] = {0, 0, 0};
  if (passive_on && k != 0) {
    double dif[3];
    subQuat(dif, quat, MD(qpos_spring) + padr);
    for (int r = 0; r < 3; r++) spring[r] = -k*dif[r];
  }
  double qc[3] = {0, 0, 0};
  if (rows_enabled(H)) {
    for (int r = 0; r < 3; r++) qc[r] = dof_friction_row(c, d + r, qv[r], qa[r]);
    if (jt == MJB_JNT_BALL && MI(jnt_limited)[jid] && !(H.disableflags & MJB_DSBL_LIMIT)) {
      const double* range = MD(jnt_range) + 2*jid;
      const double margin = MD(jnt_margin)[jid];
      double aa[3];
      quat2Vel(aa, quat, 1);
      const double value = normalize3(aa);
      const double dist = fmax(range[0], range[1]) - value;
      if (dist < margin) {
        // J = -angleAxis on the three dofs
        double vel = 0, jacc = 0;
        for (int r = 0; r < 3; r++) { vel += -aa[r]*qv[r]; jacc += -aa[r]*qa[r]; }
        const double f = scalar_row(c, ne_base(c) + H.nf_rows + c.nl, MJB_CNSTR_LIMIT_JOINT, jid,
                                    MD(sp_jnt_limit) + MJB_SP_N*jid, dist, margin,
                                    MD(dof_invweight0)[d], vel, jacc, 0);
        for (int r = 0; r < 3; r++) qc[r] += -aa[r]*f;
        c.nl++;
      }
    }
  }
  for (int r = 0; r < 3; r++) {
    double passive = spring[r];
    if (passive_on) {
      const double dmp = dof_damping[d + r];
      if (dmp != 0) passive += -dmp*qv[r];
    }
    AT(qp, d + r) = passive;
    AT(qcs, d + r) = qc[r];
  }
}

// friction-loss rows of fixed tendons (engine_core_constraint.c:793-816), after the dof rows
template <bool kSpatial>
MJB_HD inline void tendon_friction_rows(Ctx& c) {
  const mjbHdr& H = *c.H;
  if ((H.disableflags & MJB_DSBL_FRICTIONLOSS) || !H.ntendon) return;
  double* qc = SC(qfrc_c);
  const double* tfl = MD(tendon_frictionloss);
  const double* tiw = MD(tendon_invweight0);
  const double* tsp = MD(sp_tendon_friction);
  const int* tendon_adr = MI(tendon_adr); const int* tendon_num = MI(tendon_num);
  const int* wrap_objid = MI(wrap_objid); const int* jnt_dofadr = MI(jnt_dofadr);
  const double* wrap_prm = MD(wrap_prm);
  double* V = SC(ten_velocity); double* A = SC(ten_acc);
  int row = ne_base(c) + H.nf_dof_rows;
  MJB_UNROLL
  for (int t = 0; t < H.ntendon; t++) {
    if (tfl[t] > 0) {
      const double f = scalar_row(c, row++, MJB_CNSTR_FRICTION_TENDON, t, tsp + MJB_SP_N*t, 0, 0, tiw[t],
                                  AT(V, t), AT(A, t), tfl[t]);
      tendon_apply<kSpatial>(c, t, f, qc, false);
    }
  }
}

// limit rows of fixed tendons (engine_core_constraint.c:923-955), after the joint limit rows
template <bool kSpatial>
MJB_HD inline void tendon_limit_rows(Ctx& c) {
  const mjbHdr& H = *c.H;
  if ((H.disableflags & MJB_DSBL_LIMIT) || !H.ntendon) return;
  double* qc = SC(qfrc_c);
  const int* jnt_dofadr = MI(jnt_dofadr);
  const int* tendon_limited = MI(tendon_limited);
  const double* tendon_range = MD(tendon_range);
  const double* tendon_margin = MD(tendon_margin);
  const double* tiw = MD(tendon_invweight0);
  const double* tsp = MD(sp_tendon_limit);
  const int* tendon_adr = MI(tendon_adr); const int* tendon_num = MI(tendon_num);
  const int* wrap_objid = MI(wrap_objid);
  const double* wrap_prm = MD(wrap_prm);
  double* L = SC(ten_length); double* V = SC(ten_velocity); double* A = SC(ten_acc);
  MJB_UNROLL
  for (int t = 0; t < H.ntendon; t++) {
    if (!tendon_limited[t]) continue;
    const double value = AT(L, t), margin = tendon_margin[t];
    for (int side = -1; side <= 1; side += 2) {
      const double dist = side * (tendon_range[2*t + (side + 1)/2] - value);
      if (dist < margin) {
        // J = -side * ten_J
        const double f = scalar_row(c, ne_base(c) + H.nf_rows + c.nl, MJB_CNSTR_LIMIT_TENDON, t,
                                    tsp + MJB_SP_N*t, dist, margin, tiw[t], -side*AT(V, t), -side*AT(A, t), 0);
        tendon_apply<kSpatial>(c, t, -side*f, qc, false);
        c.nl++;
      }
    }
  }
}


#endif  // MJB_ROWS_H_
