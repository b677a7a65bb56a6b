// Part of the per-state mj_inverse pipeline (mjb_pipeline.h includes it inside namespace mjb, after the
// context and accessor macros; not a stand-alone header).
// Output-only stages: sensors (mj_sensorPos / Vel / Acc, mju_rayGeom, cam_project), energies, mj_camlight, mj_transmission, mj_compareFwdInv.
#ifndef MJB_OUTPUTS_H_
#define MJB_OUTPUTS_H_

// ------------------------------------------------------------------------------------------
// Sensors: mj_sensorPos, mj_sensorVel, mj_sensorAcc (engine_sensor.c:222-520, 527-704, 708-913) for
// the sensor types whose inputs exist on this path (the others are refused at upload). Runs after
// the backward sweep: body poses and cvel / cacc come from the scratch (about the tree origin O,
// so points are offset from O instead of subtree_com), cfrc_int from the mj_rnePostConstraint
// output (about the tree's centre of mass C = O + d).

// world pose of a sensor object (get_xpos_xmat / get_xquat, engine_sensor.c:73-123; the frames
// are the ones mj_kinematics builds with mj_local2Global, engine_core_smooth.c:159-200)
MJB_HD inline int sensor_object(Ctx& c, int objtype, int objid, double* pos, double* quat) {
  int body = objid;
  const double* lpos = nullptr; const double* lquat = nullptr;
  if (objtype == MJB_OBJ_BODY) { lpos = MD(body_ipos) + 3*objid; lquat = MD(body_iquat) + 4*objid; }
  else if (objtype == MJB_OBJ_GEOM) {
    body = MI(geom_bodyid)[objid]; lpos = MD(geom_pos) + 3*objid; lquat = MD(geom_quat) + 4*objid;
  } else if (objtype == MJB_OBJ_SITE) {
    body = MI(site_bodyid)[objid]; lpos = MD(site_pos) + 3*objid; lquat = MD(site_quat) + 4*objid;
  }
  double bq[4];
  ldn(pos, SC(xpos), 3*body, 3); ldn(bq, SC(xquat), 4*body, 4);
  if (lpos) {
    double m[9], r[3];
    quat2Mat(m, bq);
    mulMatVec3(r, m, lpos);
    pos[0] += r[0]; pos[1] += r[1]; pos[2] += r[2];
    mulQuat(quat, bq, lquat);
  } else {
    for (int k = 0; k < 4; k++) quat[k] = bq[k];
  }
  return body;
}

// mj_objectVelocity / mj_objectAcceleration in world axes (engine_support.c:1265-1370): motion of
// the body-fixed point `pos` from a carrier about O; acc adds the correction omega x v
MJB_HD inline void sensor_point_motion(Ctx& c, const double* carrier, int body, const double* pos, double* res) {
  double lin[3], ang[3];
  point_motion(c, carrier, body, pos, lin, ang);
  for (int k = 0; k < 3; k++) { res[k] = ang[k]; res[3 + k] = lin[k]; }
}

// mj_subtreeVel (engine_core_smooth.c:1900-1960): subtree_linvel and subtree_angmom of every body,
// left in the ia rows of the scratch (free after the inertia kernel), 21 doubles per body:
// [0..5] body velocity at xipos (world axes), [6..8] subtree_linvel, [9..11] subtree_angmom,
// [12..14] subtree_com, [15..17] xipos. Only for models with subtreelinvel / subtreeangmom sensors.
MJB_HD inline void subtree_velocities(Ctx& c) {
  const mjbHdr& H = *c.H;
  const int nbody = H.nbody;
  const int* body_parentid = MI(body_parentid);
  const double* mass = MD(body_mass); const double* stm = MD(body_subtreemass);
  const double* inertia = MD(body_inertia);
  double* t = SC(ia);
  for (int i = 0; i < nbody; i++) {
    double pos[3], quat[4], m9[9], bv[6], w[21], dv[3], lw[3];
    sensor_object(c, MJB_OBJ_BODY, i, pos, quat);
    quat2Mat(m9, quat);
    sensor_point_motion(c, SC(cvel), i, pos, bv);
    mulMatTVec3(lw, m9, bv);
    lw[0] *= inertia[3*i]; lw[1] *= inertia[3*i + 1]; lw[2] *= inertia[3*i + 2];
    mulMatVec3(dv, m9, lw);
    for (int k = 0; k < 6; k++) w[k] = bv[k];
    for (int k = 0; k < 3; k++) {
      w[6 + k] = bv[3 + k]*mass[i]; w[9 + k] = dv[k]; w[12 + k] = pos[k]*mass[i]; w[15 + k] = pos[k];
    }
    w[18] = w[19] = w[20] = 0;
    stn(t, 21*i, w, 21);
  }
  // subtree_com (mj_comPos :194-213) and subtree_linvel: momenta up the tree, then the means
  for (int i = nbody - 1; i >= 0; i--) {
    double a[9];
    ldn(a, t, 21*i + 6, 9);
    if (i) {
      const int p = body_parentid[i];
      double pa[9];
      ldn(pa, t, 21*p + 6, 9);
      for (int k = 0; k < 3; k++) { pa[k] += a[k]; pa[6 + k] += a[6 + k]; }
      stn(t, 21*p + 6, pa, 9);
    }
    const double inv = 1/fmax(MJB_MINVAL, stm[i]);
    for (int k = 0; k < 3; k++) a[k] *= inv;
    if (stm[i] < MJB_MINVAL) {
      ldn(a + 6, t, 21*i + 15, 3);
    } else {
      for (int k = 0; k < 3; k++) a[6 + k] /= stm[i];
    }
    stn(t, 21*i + 6, a, 9);
  }
  for (int i = nbody - 1; i > 0; i--) {
    const int p = body_parentid[i];
    double w[21], pw[21], dx[3], dv[3], dL[3];
    ldn(w, t, 21*i, 21); ldn(pw, t, 21*p, 21);
    for (int k = 0; k < 3; k++) { dx[k] = w[15 + k] - w[12 + k]; dv[k] = (w[3 + k] - w[6 + k])*mass[i]; }
    cross3(dL, dx, dv);
    for (int k = 0; k < 3; k++) { w[9 + k] += dL[k]; pw[9 + k] += w[9 + k]; }
    for (int k = 0; k < 3; k++) { dx[k] = w[12 + k] - pw[12 + k]; dv[k] = (w[6 + k] - pw[6 + k])*stm[i]; }
    cross3(dL, dx, dv);
    for (int k = 0; k < 3; k++) pw[9 + k] += dL[k];
    stn(t, 21*i + 9, w + 9, 3);
    stn(t, 21*p + 9, pw + 9, 3);
  }
}

// mju_rayGeom for the site shapes of touch sensors (engine_ray.c:37-52, 105-128, 222-440, 818-843):
// distance along the ray pnt + x*vec to a sphere / capsule / ellipsoid / cylinder / box at
// (pos, mat) with the given size, -1 without intersection.
MJB_DI double ray_quad(double a, double b, double cc, double* x) {
  double det = b*b - a*cc;
  if (det < MJB_MINVAL) { x[0] = -1; x[1] = -1; return -1; }
  det = sqrt(det);
  x[0] = (-b - det)/a;
  x[1] = (-b + det)/a;
  return x[0] >= 0 ? x[0] : (x[1] >= 0 ? x[1] : -1.0);
}
MJB_DI double ray_sphere(const double* pos, double dist_sqr, const double* pnt, const double* vec) {
  const double dif[3] = {pnt[0] - pos[0], pnt[1] - pos[1], pnt[2] - pos[2]};
  const double a = vec[0]*vec[0] + vec[1]*vec[1] + vec[2]*vec[2];
  const double b = vec[0]*dif[0] + vec[1]*dif[1] + vec[2]*dif[2];
  const double cc = dif[0]*dif[0] + dif[1]*dif[1] + dif[2]*dif[2] - dist_sqr;
  double xx[2];
  return ray_quad(a, b, cc, xx);
}
MJB_HD inline double ray_geom(const double* pos, const double* mat, const double* size, const double* pnt,
                              const double* vec, int type) {
  if (type == MJB_GEOM_SPHERE) return ray_sphere(pos, size[0]*size[0], pnt, vec);
  // ray_map: point and direction in the shape's frame
  const double dif[3] = {pnt[0] - pos[0], pnt[1] - pos[1], pnt[2] - pos[2]};
  double lp[3], lv[3], xx[2];
  mulMatTVec3(lp, mat, dif);
  mulMatTVec3(lv, mat, vec);
  if (type == MJB_GEOM_PLANE) {
    // ray_plane (engine_ray.c:191-217): front face only, inside the rendered rectangle when it has one
    if (lv[2] > -MJB_MINVAL) return -1;
    const double xp = -lp[2]/lv[2];
    if (xp < 0) return -1;
    const double p0 = lp[0] + xp*lv[0], p1 = lp[1] + xp*lv[1];
    return ((size[0] <= 0 || fabs(p0) <= size[0]) && (size[1] <= 0 || fabs(p1) <= size[1])) ? xp : -1.0;
  }
  if (type == MJB_GEOM_ELLIPSOID) {
    const double s[3] = {1/(size[0]*size[0]), 1/(size[1]*size[1]), 1/(size[2]*size[2])};
    const double a = s[0]*lv[0]*lv[0] + s[1]*lv[1]*lv[1] + s[2]*lv[2]*lv[2];
    const double b = s[0]*lv[0]*lp[0] + s[1]*lv[1]*lp[1] + s[2]*lv[2]*lp[2];
    const double cc = s[0]*lp[0]*lp[0] + s[1]*lp[1]*lp[1] + s[2]*lp[2]*lp[2] - 1;
    return ray_quad(a, b, cc, xx);
  }
  double x = -1, sol;
  if (type == MJB_GEOM_CAPSULE) {
    const double ssz = size[0] + size[1];
    if (ray_sphere(pos, ssz*ssz, pnt, vec) < 0) return -1;
    double a = lv[0]*lv[0] + lv[1]*lv[1];
    double b = lv[0]*lp[0] + lv[1]*lp[1];
    double cc = lp[0]*lp[0] + lp[1]*lp[1] - size[0]*size[0];
    sol = ray_quad(a, b, cc, xx);
    if (sol >= 0 && fabs(lp[2] + sol*lv[2]) <= size[1]) x = sol;
    a = lv[0]*lv[0] + lv[1]*lv[1] + lv[2]*lv[2];
    for (int cap = 1; cap >= -1; cap -= 2) {          // top cap, then bottom cap
      const double ld[3] = {lp[0], lp[1], lp[2] - cap*size[1]};
      b = lv[0]*ld[0] + lv[1]*ld[1] + lv[2]*ld[2];
      cc = ld[0]*ld[0] + ld[1]*ld[1] + ld[2]*ld[2] - size[0]*size[0];
      ray_quad(a, b, cc, xx);
      for (int i = 0; i < 2; i++) {
        const double z = lp[2] + xx[i]*lv[2];
        if (xx[i] >= 0 && (cap > 0 ? z >= size[1] : z <= -size[1]) && (x < 0 || xx[i] < x)) x = xx[i];
      }
    }
    return x;
  }
  if (type == MJB_GEOM_CYLINDER) {
    if (ray_sphere(pos, size[0]*size[0] + size[1]*size[1], pnt, vec) < 0) return -1;
    if (fabs(lv[2]) > MJB_MINVAL) {
      for (int side = -1; side <= 1; side += 2) {
        sol = (side*size[1] - lp[2])/lv[2];
        if (sol >= 0) {
          const double p0 = lp[0] + sol*lv[0], p1 = lp[1] + sol*lv[1];
          if (p0*p0 + p1*p1 <= size[0]*size[0] && (x < 0 || sol < x)) x = sol;
        }
      }
    }
    const double a = lv[0]*lv[0] + lv[1]*lv[1];
    const double b = lv[0]*lp[0] + lv[1]*lp[1];
    const double cc = lp[0]*lp[0] + lp[1]*lp[1] - size[0]*size[0];
    sol = ray_quad(a, b, cc, xx);
    if (sol >= 0 && fabs(lp[2] + sol*lv[2]) <= size[1] && (x < 0 || sol < x)) x = sol;
    return x;
  }
  if (type == MJB_GEOM_BOX) {
    if (ray_sphere(pos, size[0]*size[0] + size[1]*size[1] + size[2]*size[2], pnt, vec) < 0) return -1;
    for (int i = 0; i < 3; i++) {
      if (fabs(lv[i]) > MJB_MINVAL) {
        const int f0 = i == 0 ? 1 : 0, f1 = i == 2 ? 1 : 2;
        for (int side = -1; side <= 1; side += 2) {
          sol = (side*size[i] - lp[i])/lv[i];
          if (sol >= 0) {
            const double p0 = lp[f0] + sol*lv[f0], p1 = lp[f1] + sol*lv[f1];
            if (fabs(p0) <= size[f0] && fabs(p1) <= size[f1] && (x < 0 || sol < x)) x = sol;
          }
        }
      }
    }
    return x;
  }
  return -1;
}

// the first limit row of a joint / tendon as the limit sensors see it (engine_sensor.c:293-313,
// 600-617, 837-855): value and velocity of the coordinate in, (pos - margin, vel, force) of the row
// out; false when neither side is active. Same arithmetic as scalar_row, nothing is emitted.
MJB_HD inline bool limit_row_readings(const double* sp, const double* range, double margin, double dA,
                                      double value, double vel, double acc, double* out3) {
  for (int side = -1; side <= 1; side += 2) {
    const double dist = side * (range[(side + 1)/2] - value);
    if (dist < margin) {
      const double imp = impedance(sp, dist, margin);
      const double R = fmax(MJB_MINVAL, (1 - imp)*dA/imp);
      const double rv = -side*vel;
      const double aref = -sp[MJB_SP_B]*rv - sp[MJB_SP_K]*imp*(dist - margin);
      const double jar = -side*acc - aref;
      out3[0] = dist - margin; out3[1] = rv; out3[2] = jar >= 0 ? 0.0 : -(1/R)*jar;
      return true;
    }
  }
  return false;
}

// cam_project (engine_sensor.c:126-215): pixel coordinates of a world point, through the product of
// the image, focal, rotation and translation matrices accumulated in the reference's loop order
MJB_HD inline void cam_project(double* px, const double* target, const double* cam_xpos, const double* cam_xmat,
                               const double* prj) {
  double translation[4][4] = {{1, 0, 0, -cam_xpos[0]}, {0, 1, 0, -cam_xpos[1]}, {0, 0, 1, -cam_xpos[2]}, {0, 0, 0, 1}};
  double rotation[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 1}};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rotation[i][j] = cam_xmat[j*3 + i];
  const double focal[3][4] = {{-prj[0], 0, 0, 0}, {0, prj[1], 0, 0}, {0, 0, 1.0, 0}};
  const double image[3][3] = {{1, 0, prj[2]}, {0, 1, prj[3]}, {0, 0, 1}};
  double proj[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 4; k++)
        for (int l = 0; l < 4; l++)
          for (int n = 0; n < 4; n++) proj[i][n] += image[i][j] * focal[j][k] * rotation[k][l] * translation[l][n];
  const double hom[4] = {target[0], target[1], target[2], 1};
  double pix[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) pix[i] += proj[i][j] * hom[j];
  double denom = pix[2];
  if (fabs(denom) < MJB_MINVAL) denom = denom < 0 ? fmin(denom, -MJB_MINVAL) : fmax(denom, MJB_MINVAL);
  px[0] = pix[0] / denom;
  px[1] = pix[1] / denom;
}

// one geom pair of mj_geomDistance (engine_support.c:1434-1449): the pair's primitive collision function with
// the sensor cutoff as margin; returns the number of contacts written to con (normal in frame[0..2]).
// Free of the per-state context, out of line: the sensor kernel carries one copy of the primitives.
MJB_COLD inline int geom_pair_contacts(Con* con, int fn, double margin, const double* pos1, const double* mat1,
                                       const double* size1, const double* pos2, const double* mat2,
                                       const double* size2) {
  switch (fn) {
    case MJB_FN_PLANE_SPHERE: return plane_sphere(con, margin, pos1, mat1, pos2, size2[0]);
    case MJB_FN_PLANE_CAPSULE: return plane_capsule(con, margin, pos1, mat1, pos2, mat2, size2);
    case MJB_FN_PLANE_CYLINDER: return plane_cylinder(con, margin, pos1, mat1, pos2, mat2, size2);
    case MJB_FN_PLANE_BOX: return plane_box(con, margin, pos1, mat1, pos2, mat2, size2);
    case MJB_FN_PLANE_ELLIPSOID: return plane_ellipsoid(con, margin, pos1, mat1, pos2, mat2, size2);
    case MJB_FN_SPHERE_SPHERE: return sphere_sphere(con, margin, pos1, mat1, size1[0], pos2, mat2, size2[0]);
    case MJB_FN_SPHERE_CAPSULE: return sphere_capsule(con, margin, pos1, mat1, size1, pos2, mat2, size2);
    case MJB_FN_SPHERE_CYLINDER: return sphere_cylinder(con, margin, pos1, mat1, size1, pos2, mat2, size2);
    case MJB_FN_SPHERE_BOX: return sphere_box(con, margin, pos1, size1, pos2, mat2, size2);
    case MJB_FN_CAPSULE_CAPSULE: return capsule_capsule(con, margin, pos1, mat1, size1, pos2, mat2, size2);
    case MJB_FN_CAPSULE_BOX: return capsule_box(con, margin, pos1, mat1, size1, pos2, mat2, size2);
    default: return 0;
  }
}

// kCcd: geom-distance sensors over pairs that the reference measures with mjc_ccd (box-box and the convex pairs,
// mj_geomDistanceCCD engine_support.c:1376-1402) are compiled in -- the kernel instantiation for models with such sensors
template <bool kCcd = MJB_CONVEX_DEFAULT>
MJB_HD inline void sensors(Ctx& c) {
  const mjbHdr& H = *c.H;
  const size_t N = (size_t)c.N;
  const int* sen = MI(sensor_int);
  const double* cutoff = MD(sensor_cutoff);
  const int* body_parentid = MI(body_parentid);
  const int* rootid = MI(body_rootid);
  if (H.sensor_subtreevel) subtree_velocities(c);
  for (int i = 0; i < H.nsensor; i++) {
    const int* si = sen + MJB_SEN_NI*i;
    const int type = si[MJB_SEN_TYPE], objtype = si[MJB_SEN_OBJTYPE], objid = si[MJB_SEN_OBJID];
    const int reftype = si[MJB_SEN_REFTYPE], refid = si[MJB_SEN_REFID];
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (type == MJB_SENS_GEOMDIST || type == MJB_SENS_GEOMNORMAL || type == MJB_SENS_GEOMFROMTO) {
      // smallest signed distance between the geoms of two objects, its direction, its end points
      // (engine_sensor.c:378-463); the pairs and their collision functions are listed at upload
      const double margin = cutoff[i];
      const int* pr = MI(sensor_pairs) + 4*objid;
      double dist = margin, fromto[6] = {0, 0, 0, 0, 0, 0};
      for (int p = 0; p < refid; p++, pr += 4) {
        double p1[3], q1[4], m1[9], p2[3], q2[4], m2[9];
        sensor_object(c, MJB_OBJ_GEOM, pr[0], p1, q1); quat2Mat(m1, q1);
        sensor_object(c, MJB_OBJ_GEOM, pr[1], p2, q2); quat2Mat(m2, q2);
        if (pr[2] == MJB_FN_CONVEX || pr[2] == MJB_FN_BOX_BOX) {
          // GJK distance with the cutoff, EPA depth when penetrating; the end points stay in the pair's type order
          if (kCcd) {
            CvxRun r;
            cvx_ccd(r, 0, margin, MI(geom_type)[pr[0]], p1, m1, MD(geom_size) + 3*pr[0], MI(geom_type)[pr[1]], p2, m2,
                    MD(geom_size) + 3*pr[1], H.ccd_tolerance, H.ccd_iterations);
            if (r.dist < dist) {
              dist = r.dist;
              for (int k = 0; k < 3; k++) { fromto[k] = r.nx > 0 ? r.xa[k] : 0; fromto[3 + k] = r.nx > 0 ? r.xb[k] : 0; }
            }
          }
          continue;
        }
        Con con[8];
        const int num = geom_pair_contacts(con, pr[2], margin, p1, m1, MD(geom_size) + 3*pr[0], p2, m2,
                                           MD(geom_size) + 3*pr[1]);
        double dnew = margin;
        int smallest = -1;
        for (int k = 0; k < num; k++) {
          if (con[k].dist < dnew) { dnew = con[k].dist; smallest = k; }
        }
        if (smallest >= 0 && dnew < dist) {
          dist = dnew;
          const double sign = pr[3] ? -1 : 1;
          for (int k = 0; k < 3; k++) {
            fromto[k] = con[smallest].pos[k] + con[smallest].frame[k]*(-0.5*sign*dnew);
            fromto[3 + k] = con[smallest].pos[k] + con[smallest].frame[k]*(0.5*sign*dnew);
          }
        }
      }
      if (type == MJB_SENS_GEOMDIST) {
        v[0] = dist;
      } else if (type == MJB_SENS_GEOMNORMAL) {
        double nrm[3] = {fromto[3] - fromto[0], fromto[4] - fromto[1], fromto[5] - fromto[2]};
        if (nrm[0] != 0 || nrm[1] != 0 || nrm[2] != 0) normalize3(nrm);
        v[0] = nrm[0]; v[1] = nrm[1]; v[2] = nrm[2];
      } else {
        for (int k = 0; k < 6; k++) v[k] = fromto[k];
      }
    } else if (type == MJB_SENS_TOUCH) {
      // sum of the normal forces of the contacts of the site's body whose normal ray meets the site
      // volume (engine_sensor.c:750-793); mj_contactForce's normal component is the row force
      // (frictionless, elliptic) or the sum of the pyramid's row forces (engine_support.c:1459-1490)
      double pos[3], quat[4], m[9];
      const int body = sensor_object(c, MJB_OBJ_SITE, objid, pos, quat);
      quat2Mat(m, quat);
      const int* geom_bodyid = MI(geom_bodyid);
      int ncon = c.isc[MJB_ISC_NCON * MJB_LS];
      if (ncon > c.nconmax) ncon = c.nconmax;
      double total = 0;
      for (int k = 0; k < ncon; k++) {
        const int adr = c.out.contact_info[(size_t)(3*k + 2)*N + c.s];
        if (adr < 0) continue;
        const int b1 = geom_bodyid[c.out.contact_geom[(size_t)(2*k)*N + c.s]];
        const int b2 = geom_bodyid[c.out.contact_geom[(size_t)(2*k + 1)*N + c.s]];
        if (body != b1 && body != b2) continue;
        const int dim = c.out.contact_info[(size_t)(3*k)*N + c.s];
        const int nrow = (dim > 1 && H.cone == 0) ? 2*(dim - 1) : 1;
        if (adr + nrow > c.njmax) continue;
        double fn = 0;
        for (int r = 0; r < nrow; r++) fn += c.out.efc_num[(size_t)(8*(adr + r) + 6)*N + c.s];
        if (fn <= 0) continue;
        double ray[3], p[3];
        for (int j = 0; j < 3; j++) {
          ray[j] = c.out.contact_num[(size_t)(13*k + 4 + j)*N + c.s]*fn;
          p[j] = c.out.contact_num[(size_t)(13*k + 1 + j)*N + c.s];
        }
        normalize3(ray);
        if (body == b2) { ray[0] = -ray[0]; ray[1] = -ray[1]; ray[2] = -ray[2]; }
        if (ray_geom(pos, m, MD(site_size) + 3*objid, p, ray, MI(site_type)[objid]) >= 0) total += fn;
      }
      v[0] = total;
    } else if (type == MJB_SENS_MAGNETOMETER) {
      // opt.magnetic in the site frame (engine_sensor.c:254-257)
      double pos[3], quat[4], m[9];
      sensor_object(c, MJB_OBJ_SITE, objid, pos, quat);
      quat2Mat(m, quat);
      mulMatTVec3(v, m, H.magnetic);
    } else if (type == MJB_SENS_RANGEFINDER) {
      // mj_ray from the site along its z axis over every geom that is not eliminated (engine_sensor.c:266-275,
      // engine_ray.c:69-100, 1145-1185: the site's own body, invisible geoms -- the static part of the
      // test is the ray_geom table); geom frames are rebuilt from the body poses (mj_local2Global)
      double pos[3], quat[4], m[9], gp[3], gq[4], gm[9];
      const int body = sensor_object(c, MJB_OBJ_SITE, objid, pos, quat);
      quat2Mat(m, quat);
      const double rvec[3] = {m[2], m[5], m[8]};
      const int* ray_ok = MI(ray_geom); const int* geom_bodyid = MI(geom_bodyid); const int* geom_type = MI(geom_type);
      double dist = -1;
      for (int g = 0; g < H.ngeom; g++) {
        if (!ray_ok[g] || geom_bodyid[g] == body) continue;
        sensor_object(c, MJB_OBJ_GEOM, g, gp, gq);
        quat2Mat(gm, gq);
        const double nd = ray_geom(gp, gm, MD(geom_size) + 3*g, pos, rvec, geom_type[g]);
        if (nd >= 0 && (nd < dist || dist < 0)) dist = nd;
      }
      v[0] = dist;
    } else if (type == MJB_SENS_CAMPROJECTION) {
      // site position in the image of camera refid (engine_sensor.c:259-264); the camera pose is
      // mj_camlight's output (mjb_makeData adds mjbOUT_CAMLIGHT for these sensors)
      double pos[3], quat[4], cp[3], cm[9];
      sensor_object(c, MJB_OBJ_SITE, objid, pos, quat);
      for (int k = 0; k < 3; k++) cp[k] = c.out.cam_xpos[(size_t)(3*refid + k)*N + c.s];
      for (int k = 0; k < 9; k++) cm[k] = c.out.cam_xmat[(size_t)(9*refid + k)*N + c.s];
      cam_project(v, pos, cp, cm, MD(cam_proj) + 4*refid);
    } else if (type == MJB_SENS_ACTUATORPOS) {
      v[0] = c.out.actuator_length[(size_t)objid*N + c.s];
    } else if (type == MJB_SENS_ACTUATORVEL) {
      v[0] = c.out.actuator_velocity[(size_t)objid*N + c.s];
    } else if (type == MJB_SENS_E_POTENTIAL) {
      v[0] = c.out.energy[c.s];
    } else if (type == MJB_SENS_E_KINETIC) {
      v[0] = c.out.energy[N + c.s];
    } else if (type == MJB_SENS_CLOCK) {
      v[0] = 0;       // d->time is not part of the batched state: 0, as after mj_resetData
    } else if (type == MJB_SENS_JOINTPOS) {
      v[0] = QPOS(MI(jnt_qposadr)[objid]);
    } else if (type == MJB_SENS_JOINTVEL) {
      v[0] = QVEL(MI(jnt_dofadr)[objid]);
    } else if (type == MJB_SENS_TENDONPOS || type == MJB_SENS_TENDONVEL) {
      if (MI(wrap_type)[MI(tendon_adr)[objid]] != MJB_WRAP_JOINT && !MI(tendon_active)[objid]) {
        // a spatial tendon that carries no force is not walked by the smooth phase: walk it here
        double vel, acc;
        const double len = spatial_tendon_kinematics(c, objid, &vel, &acc);
        v[0] = type == MJB_SENS_TENDONPOS ? len : vel;
      } else {
        v[0] = type == MJB_SENS_TENDONPOS ? AT(SC(ten_length), objid) : AT(SC(ten_velocity), objid);
      }
    } else if (type == MJB_SENS_BALLQUAT) {
      const int a = MI(jnt_qposadr)[objid];
      for (int k = 0; k < 4; k++) v[k] = QPOS(a + k);
      normalize4(v);
    } else if (type == MJB_SENS_BALLANGVEL) {
      const int a = MI(jnt_dofadr)[objid];
      for (int k = 0; k < 3; k++) v[k] = QVEL(a + k);
    } else if (type >= MJB_SENS_JOINTLIMITPOS && type <= MJB_SENS_TENDONLIMITFRC) {
      // limit sensors: the readings of the object's first limit row, 0 while the limit is inactive
      double r3[3] = {0, 0, 0};
      if (!(H.disableflags & MJB_DSBL_LIMIT) && rows_enabled(H)) {
        if (type <= MJB_SENS_JOINTLIMITFRC) {
          if (MI(jnt_limited)[objid]) {
            const int dof = MI(jnt_dofadr)[objid];
            limit_row_readings(MD(sp_jnt_limit) + MJB_SP_N*objid, MD(jnt_range) + 2*objid, MD(jnt_margin)[objid],
                               MD(dof_invweight0)[dof], QPOS(MI(jnt_qposadr)[objid]), QVEL(dof), QACC(dof), r3);
          }
        } else if (MI(tendon_limited)[objid]) {
          limit_row_readings(MD(sp_tendon_limit) + MJB_SP_N*objid, MD(tendon_range) + 2*objid,
                             MD(tendon_margin)[objid], MD(tendon_invweight0)[objid], AT(SC(ten_length), objid),
                             AT(SC(ten_velocity), objid), AT(SC(ten_acc), objid), r3);
        }
      }
      v[0] = r3[(type - MJB_SENS_JOINTLIMITPOS) % 3];
    } else if (type == MJB_SENS_SUBTREELINVEL) {
      ldn(v, SC(ia), 21*objid + 6, 3);
    } else if (type == MJB_SENS_SUBTREEANGMOM) {
      ldn(v, SC(ia), 21*objid + 9, 3);
    } else if (type == MJB_SENS_SUBTREECOM) {
      // mj_comPos (engine_core_smooth.c:183-225): mass-weighted mean of xipos over the subtree,
      // whose bodies are contiguous; mass*(xipos - O) and mass are cinert[6..9]
      double ms[4] = {0, 0, 0, 0}, o[3];
      int e = objid;
      do {
        double t[4];
        ldn(t, SC(cinert), 10*e + 6, 4);
        for (int k = 0; k < 4; k++) ms[k] += t[k];
        e++;
      } while (e < H.nbody && body_parentid[e] >= objid && objid > 0);
      if (objid == 0) {
        // the world's subtree is every body, each tree about its own origin
        ms[0] = ms[1] = ms[2] = ms[3] = 0;
        for (int b = 1; b < H.nbody; b++) {
          double t[4];
          ldn(t, SC(cinert), 10*b + 6, 4); ldn(o, SC(origin), 3*rootid[b], 3);
          for (int k = 0; k < 3; k++) ms[k] += t[k] + t[3]*o[k];
          ms[3] += t[3];
        }
        o[0] = o[1] = o[2] = 0;
      } else {
        ldn(o, SC(origin), 3*rootid[objid], 3);
      }
      if (ms[3] >= MJB_MINVAL) {
        for (int k = 0; k < 3; k++) v[k] = o[k] + ms[k]/ms[3];
      } else {
        double q[4];
        sensor_object(c, MJB_OBJ_BODY, objid, v, q);
      }
    } else if (type >= MJB_SENS_FRAMEPOS && type <= MJB_SENS_FRAMEZAXIS) {
      double pos[3], quat[4], rpos[3], rquat[4], rmat[9];
      sensor_object(c, objtype, objid, pos, quat);
      if (refid >= 0) { sensor_object(c, reftype, refid, rpos, rquat); quat2Mat(rmat, rquat); }
      if (type == MJB_SENS_FRAMEQUAT) {
        if (refid >= 0) {
          const double nq[4] = {rquat[0], -rquat[1], -rquat[2], -rquat[3]};
          mulQuat(v, nq, quat);
        } else {
          for (int k = 0; k < 4; k++) v[k] = quat[k];
        }
      } else {
        double w[3];
        if (type == MJB_SENS_FRAMEPOS) {
          for (int k = 0; k < 3; k++) w[k] = refid >= 0 ? pos[k] - rpos[k] : pos[k];
        } else {
          double m[9];
          quat2Mat(m, quat);
          const int off = type - MJB_SENS_FRAMEXAXIS;
          w[0] = m[off]; w[1] = m[off + 3]; w[2] = m[off + 6];
        }
        if (refid >= 0) mulMatTVec3(v, rmat, w); else { v[0] = w[0]; v[1] = w[1]; v[2] = w[2]; }
      }
    } else if (type == MJB_SENS_VELOCIMETER || type == MJB_SENS_GYRO || type == MJB_SENS_ACCELEROMETER) {
      // site velocity / acceleration in the site frame
      double pos[3], quat[4], m[9], x[6], a[6];
      const int body = sensor_object(c, MJB_OBJ_SITE, objid, pos, quat);
      quat2Mat(m, quat);
      sensor_point_motion(c, SC(cvel), body, pos, x);
      if (type == MJB_SENS_ACCELEROMETER) {
        double cr[3];
        sensor_point_motion(c, SC(cacc), body, pos, a);
        cross3(cr, x, x + 3);
        for (int k = 0; k < 3; k++) a[3 + k] += cr[k];
        mulMatTVec3(v, m, a + 3);
      } else {
        mulMatTVec3(v, m, type == MJB_SENS_GYRO ? x : x + 3);
      }
    } else if (type == MJB_SENS_FORCE || type == MJB_SENS_TORQUE) {
      // cfrc_int of the site's body moved from C to the site and rotated into the site frame
      double pos[3], quat[4], m[9], f[6], o[3], ms[4] = {0, 0, 0, 0};
      const int body = sensor_object(c, MJB_OBJ_SITE, objid, pos, quat);
      quat2Mat(m, quat);
      const int r = rootid[body];
      int e = r;
      do {
        double t[4];
        ldn(t, SC(cinert), 10*e + 6, 4);
        for (int k = 0; k < 4; k++) ms[k] += t[k];
        e++;
      } while (e < H.nbody && body_parentid[e] != 0);
      ldn(o, SC(origin), 3*r, 3);
      for (int k = 0; k < 6; k++) f[k] = c.out.cfrc_int[(size_t)(6*body + k)*N + c.s];
      if (type == MJB_SENS_FORCE) {
        mulMatTVec3(v, m, f + 3);
      } else {
        double dif[3], cr[3];
        for (int k = 0; k < 3; k++) dif[k] = pos[k] - (o[k] + (ms[3] >= MJB_MINVAL ? ms[k]/ms[3] : 0.0));
        cross3(cr, dif, f + 3);
        for (int k = 0; k < 3; k++) f[k] -= cr[k];
        mulMatTVec3(v, m, f);
      }
    } else if (type == MJB_SENS_FRAMELINVEL || type == MJB_SENS_FRAMEANGVEL) {
      double pos[3], quat[4], x[6];
      const int body = sensor_object(c, objtype, objid, pos, quat);
      sensor_point_motion(c, SC(cvel), body, pos, x);
      if (refid >= 0) {
        // relative to a moving reference frame (engine_sensor.c:625-647)
        double rpos[3], rquat[4], rmat[9], xr[6], rel[6], rvec[3], cr[3];
        const int rbody = sensor_object(c, reftype, refid, rpos, rquat);
        quat2Mat(rmat, rquat);
        sensor_point_motion(c, SC(cvel), rbody, rpos, xr);
        for (int k = 0; k < 6; k++) rel[k] = x[k] - xr[k];
        for (int k = 0; k < 3; k++) rvec[k] = pos[k] - rpos[k];
        cross3(cr, rvec, xr);
        for (int k = 0; k < 3; k++) rel[3 + k] += cr[k];
        mulMatTVec3(x, rmat, rel);
        mulMatTVec3(x + 3, rmat, rel + 3);
      }
      for (int k = 0; k < 3; k++) v[k] = type == MJB_SENS_FRAMELINVEL ? x[3 + k] : x[k];
    } else if (type == MJB_SENS_FRAMELINACC || type == MJB_SENS_FRAMEANGACC) {
      double pos[3], quat[4], x[6], a[6], cr[3];
      const int body = sensor_object(c, objtype, objid, pos, quat);
      sensor_point_motion(c, SC(cvel), body, pos, x);
      sensor_point_motion(c, SC(cacc), body, pos, a);
      cross3(cr, x, x + 3);
      for (int k = 0; k < 3; k++) v[k] = type == MJB_SENS_FRAMELINACC ? a[3 + k] + cr[k] : a[k];
    }
    // apply_cutoff (engine_sensor.c:40-68): real values on both sides, positive ones from above
    const double cut = cutoff[i];
    const int dt = si[MJB_SEN_DATATYPE];
    for (int k = 0; k < si[MJB_SEN_DIM] && k < 6; k++) {
      double x = v[k];
      if (cut > 0 && type != MJB_SENS_GEOMFROMTO) {
        if (dt == MJB_DATATYPE_REAL) x = x < -cut ? -cut : (x > cut ? cut : x);
        else if (dt == MJB_DATATYPE_POSITIVE) x = cut < x ? cut : x;
      }
      c.out.sensordata[(size_t)(si[MJB_SEN_ADR] + k)*N + c.s] = x;
    }
  }
}

// ------------------------------------------------------------------------------------------
// mj_energyPos / mj_energyVel (engine_sensor.c:920-1008, 1011-1020) for models with mjENBL_ENERGY,
// after the sweeps, from the scratch:
//   energy[0] = -sum_b m_b g.xipos_b + joint springs + tendon springs   (flex models are refused)
//   energy[1] = 0.5 qvel' M qvel
// m_b (xipos_b - O) and m_b are cinert[6..9] of the body (about the tree origin O), so the gravity
// term needs no pose; the kinetic energy is summed per body, 0.5 cvel.(cinert cvel), plus the
// armature terms -- the same quadratic form as qvel' M qvel (M = sum_b J_b' I_b J_b + armature).
MJB_HD inline void energy(Ctx& c) {
  const mjbHdr& H = *c.H;
  const size_t N = (size_t)c.N;
  const int* rootid = MI(body_rootid);
  double e0 = 0;
  if (!(H.disableflags & MJB_DSBL_GRAVITY)) {
    for (int b = 1; b < H.nbody; b++) {
      double t[4], o[3];
      ldn(t, SC(cinert), 10*b + 6, 4); ldn(o, SC(origin), 3*rootid[b], 3);
      const double mx[3] = {t[0] + t[3]*o[0], t[1] + t[3]*o[1], t[2] + t[3]*o[2]};     // m * xipos
      e0 -= dot3(H.gravity, mx);
    }
  }
  if (!(H.disableflags & MJB_DSBL_PASSIVE)) {
    const int* jnt_type = MI(jnt_type); const int* jnt_qposadr = MI(jnt_qposadr);
    const double* stiff = MD(jnt_stiffness); const double* qs = MD(qpos_spring);
    for (int j = 0; j < H.njnt; j++) {
      const double k = stiff[j];
      int padr = jnt_qposadr[j];
      const int jt = jnt_type[j];
      if (jt == MJB_JNT_FREE) {
        // as the reference has it (:940-944): the first FOUR coordinates normalised as a unit, then the
        // first three of them against the spring position
        double q4[4] = {QPOS(padr), QPOS(padr + 1), QPOS(padr + 2), QPOS(padr + 3)};
        normalize4(q4);
        const double dif[3] = {q4[0] - qs[padr], q4[1] - qs[padr + 1], q4[2] - qs[padr + 2]};
        e0 += 0.5*k*dot3(dif, dif);
        padr += 3;
      }
      if (jt == MJB_JNT_FREE || jt == MJB_JNT_BALL) {
        // mju_subQuat on the quaternion as stored (:953 passes d->qpos, not the normalised copy)
        const double q4[4] = {QPOS(padr), QPOS(padr + 1), QPOS(padr + 2), QPOS(padr + 3)};
        double dif[3];
        subQuat(dif, q4, qs + padr);
        e0 += 0.5*k*dot3(dif, dif);
      } else {
        const double d = QPOS(padr) - qs[padr];
        e0 += 0.5*k*d*d;
      }
    }
    const double* tstiff = MD(tendon_stiffness); const double* ls = MD(tendon_lengthspring);
    for (int t = 0; t < H.ntendon; t++) {
      const double length = AT(SC(ten_length), t);
      double disp = 0;
      if (length > ls[2*t + 1]) disp = ls[2*t + 1] - length;
      else if (length < ls[2*t]) disp = ls[2*t] - length;
      e0 += 0.5*tstiff[t]*disp*disp;
    }
  }
  double e1 = 0;
  for (int b = 1; b < H.nbody; b++) {
    double ci[10], v[6], iv[6];
    ldn(ci, SC(cinert), 10*b, 10); ldn(v, SC(cvel), 6*b, 6);
    mulInertVec(iv, ci, v);
    e1 += dot6(v, iv);
  }
  const double* arm = MD(dof_armature);
  for (int i = 0; i < H.nv; i++) { const double qv = QVEL(i); e1 += arm[i]*qv*qv; }
  c.out.energy[c.s] = e0;
  c.out.energy[N + c.s] = 0.5*e1;
}

// ------------------------------------------------------------------------------------------
// mj_camlight (engine_core_smooth.c:275-389): world poses of cameras and lights from the body poses
// of the sweep. subtree_com of every body (mj_comPos :190-213, the reference's accumulation order)
// is rebuilt in the ia rows of the scratch (free after the inertia kernel) only when a camera or
// light tracks or targets a subtree's centre of mass.
MJB_HD inline void camlight_point(Ctx& c, int body, const double* lpos, double* pos, double* bq) {
  ldn(pos, SC(xpos), 3*body, 3); ldn(bq, SC(xquat), 4*body, 4);
  double m[9], r[3];
  quat2Mat(m, bq);
  mulMatVec3(r, m, lpos);
  pos[0] = r[0] + pos[0]; pos[1] = r[1] + pos[1]; pos[2] = r[2] + pos[2];
}
MJB_HD inline void camlight(Ctx& c) {
  const mjbHdr& H = *c.H;
  const size_t N = (size_t)c.N;
  const int* cam_mode = MI(cam_mode); const int* light_mode = MI(light_mode);
  bool need_com = false;
  for (int i = 0; i < H.ncam; i++) need_com |= cam_mode[i] == MJB_CAMLIGHT_TRACKCOM || cam_mode[i] == MJB_CAMLIGHT_TARGETBODYCOM;
  for (int i = 0; i < H.nlight; i++) need_com |= light_mode[i] == MJB_CAMLIGHT_TRACKCOM || light_mode[i] == MJB_CAMLIGHT_TARGETBODYCOM;
  double* com = SC(ia);      // [0..2] subtree_com, [3] subtree mass, [4..6] xipos per body (stride 8)
  if (need_com) {
    const int* body_parentid = MI(body_parentid);
    const double* mass = MD(body_mass);
    const double zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < H.nbody; i++) stn(com, 8*i, zero, 8);
    for (int i = H.nbody - 1; i >= 0; i--) {
      double w[8], xi[3], bq[4];
      ldn(w, com, 8*i, 8);
      camlight_point(c, i, MD(body_ipos) + 3*i, xi, bq);
      for (int k = 0; k < 3; k++) w[k] += xi[k]*mass[i];
      w[3] += mass[i];
      if (i) {
        const int p = body_parentid[i];
        double pw[4];
        ldn(pw, com, 8*p, 4);
        for (int k = 0; k < 4; k++) pw[k] += w[k];
        stn(com, 8*p, pw, 4);
      }
      if (w[3] < MJB_MINVAL) {
        for (int k = 0; k < 3; k++) w[k] = xi[k];
      } else {
        const double inv = 1.0/fmax(MJB_MINVAL, w[3]);
        for (int k = 0; k < 3; k++) w[k] *= inv;
      }
      stn(com, 8*i, w, 4);
    }
  }
  double tgt[3];
  for (int i = 0; i < H.ncam; i++) {
    const int id = MI(cam_bodyid)[i], id1 = MI(cam_targetbodyid)[i], mode = cam_mode[i];
    double pos[3], bq[4], q[4], mat[9];
    camlight_point(c, id, MD(cam_pos) + 3*i, pos, bq);
    mulQuat(q, bq, MD(cam_quat) + 4*i);
    quat2Mat(mat, q);
    if (mode == MJB_CAMLIGHT_TRACK || mode == MJB_CAMLIGHT_TRACKCOM) {
      for (int k = 0; k < 9; k++) mat[k] = MD(cam_mat0)[9*i + k];
      if (mode == MJB_CAMLIGHT_TRACK) {
        ldn(pos, SC(xpos), 3*id, 3);
        for (int k = 0; k < 3; k++) pos[k] += MD(cam_pos0)[3*i + k];
      } else {
        ldn(pos, com, 8*id, 3);
        for (int k = 0; k < 3; k++) pos[k] += MD(cam_poscom0)[3*i + k];
      }
    } else if ((mode == MJB_CAMLIGHT_TARGETBODY || mode == MJB_CAMLIGHT_TARGETBODYCOM) && id1 >= 0) {
      if (mode == MJB_CAMLIGHT_TARGETBODY) ldn(tgt, SC(xpos), 3*id1, 3); else ldn(tgt, com, 8*id1, 3);
      double T[9];
      for (int k = 0; k < 3; k++) T[6 + k] = pos[k] - tgt[k];   // z axis = -viewing direction
      normalize3(T + 6);
      T[3] = 0; T[4] = 0; T[5] = 1;
      cross3(T, T + 3, T + 6);
      normalize3(T);
      cross3(T + 3, T + 6, T);
      normalize3(T + 3);
      for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) mat[3*r + k] = T[3*k + r];
    }
    for (int k = 0; k < 3; k++) c.out.cam_xpos[(size_t)(3*i + k)*N + c.s] = pos[k];
    for (int k = 0; k < 9; k++) c.out.cam_xmat[(size_t)(9*i + k)*N + c.s] = mat[k];
  }
  for (int i = 0; i < H.nlight; i++) {
    const int id = MI(light_bodyid)[i], id1 = MI(light_targetbodyid)[i], mode = light_mode[i];
    double pos[3], bq[4], dir[3];
    camlight_point(c, id, MD(light_pos) + 3*i, pos, bq);
    rotVecQuat(dir, MD(light_dir) + 3*i, bq);
    if (mode == MJB_CAMLIGHT_TRACK || mode == MJB_CAMLIGHT_TRACKCOM) {
      for (int k = 0; k < 3; k++) dir[k] = MD(light_dir0)[3*i + k];
      if (mode == MJB_CAMLIGHT_TRACK) {
        ldn(pos, SC(xpos), 3*id, 3);
        for (int k = 0; k < 3; k++) pos[k] += MD(light_pos0)[3*i + k];
      } else {
        ldn(pos, com, 8*id, 3);
        for (int k = 0; k < 3; k++) pos[k] += MD(light_poscom0)[3*i + k];
      }
    } else if ((mode == MJB_CAMLIGHT_TARGETBODY || mode == MJB_CAMLIGHT_TARGETBODYCOM) && id1 >= 0) {
      if (mode == MJB_CAMLIGHT_TARGETBODY) ldn(tgt, SC(xpos), 3*id1, 3); else ldn(tgt, com, 8*id1, 3);
      for (int k = 0; k < 3; k++) dir[k] = tgt[k] - pos[k];
    }
    normalize3(dir);
    for (int k = 0; k < 3; k++) c.out.light_xpos[(size_t)(3*i + k)*N + c.s] = pos[k];
    for (int k = 0; k < 3; k++) c.out.light_xdir[(size_t)(3*i + k)*N + c.s] = dir[k];
  }
}

// ------------------------------------------------------------------------------------------
// mj_transmission (engine_core_smooth.c:865-1346) and actuator_velocity (mj_fwdVelocity,
// engine_forward.c:216): actuator_length [nu], actuator_moment as the DENSE nu x nv matrix the
// reference's compressed rows (moment_rownnz / rowadr / colind) expand to, actuator_velocity [nu].
// The Jacobians of the reference (mj_jacSite, mj_jacPointAxis, ten_J) are never formed: a moment row
// is the projection of a wrench on the dof chain of a body,
//   row[j] += s * ( F . (ang_j x (p - O) + lin_j) + T . ang_j ),   cdof_j = (ang_j, lin_j) about O,
// the column of (jacp' F + jacr' T).
MJB_HD inline void trn_project(Ctx& c, double* row, int body, int stop_dof, const double* p, const double* F,
                               const double* T, double s) {
  const int wb = MI(body_weldid)[body];
  if (!MI(body_dofnum)[wb]) return;
  const int* dof_parentid = MI(dof_parentid);
  const size_t N = (size_t)c.N;
  double o[3];
  ldn(o, SC(origin), 3*MI(body_rootid)[wb], 3);
  const double r[3] = {p[0] - o[0], p[1] - o[1], p[2] - o[2]};
  for (int j = MI(body_dofadr)[wb] + MI(body_dofnum)[wb] - 1; j >= 0 && j != stop_dof; j = dof_parentid[j]) {
    double cd[6], jp[3], v = 0;
    ldn(cd, SC(cdof), 6*j, 6);
    if (F) {
      cross3(jp, cd, r);
      jp[0] += cd[3]; jp[1] += cd[4]; jp[2] += cd[5];
      v = jp[0]*F[0] + jp[1]*F[1] + jp[2]*F[2];
    }
    if (T) v += cd[0]*T[0] + cd[1]*T[1] + cd[2]*T[2];
    row[(size_t)j*N] += s*v;
  }
}
MJB_HD inline void transmission(Ctx& c) {
  const mjbHdr& H = *c.H;
  const size_t N = (size_t)c.N;
  const int nv = H.nv;
  const int* trntype = MI(actuator_trntype); const int* trn = MI(actuator_trn);
  const int* jnt_type = MI(jnt_type); const int* jnt_qposadr = MI(jnt_qposadr); const int* jnt_dofadr = MI(jnt_dofadr);
  const int* site_bodyid = MI(site_bodyid);
  for (int i = 0; i < H.nu; i++) {
    const int id = trn[2*i], type = trntype[i];
    const double* gear = MD(actuator_gear) + 6*i;
    double* row = c.out.actuator_moment + (size_t)i*nv*N + c.s;
    for (int j = 0; j < nv; j++) row[(size_t)j*N] = 0;
    double length = 0;
    if (type == MJB_TRN_JOINT || type == MJB_TRN_JOINTINPARENT) {
      const int jt = jnt_type[id], qadr = jnt_qposadr[id], dadr = jnt_dofadr[id];
      if (jt == MJB_JNT_SLIDE || jt == MJB_JNT_HINGE) {
        length = QPOS(qadr)*gear[0];
        row[(size_t)dadr*N] = gear[0];
      } else {
        // ball: gear axis against the joint's expmap; free: the last three dofs take the rotational gear
        const int q0 = jt == MJB_JNT_BALL ? qadr : qadr + 3;
        const double* g = jt == MJB_JNT_BALL ? gear : gear + 3;
        double quat[4] = {QPOS(q0), QPOS(q0 + 1), QPOS(q0 + 2), QPOS(q0 + 3)};
        double axis[3], ga[3] = {g[0], g[1], g[2]};
        normalize4(quat);
        if (jt == MJB_JNT_BALL) quat2Vel(axis, quat, 1);
        if (type == MJB_TRN_JOINTINPARENT) {
          const double nq[4] = {quat[0], -quat[1], -quat[2], -quat[3]};
          rotVecQuat(ga, g, nq);
        }
        if (jt == MJB_JNT_BALL) {
          length = dot3(axis, ga);
          for (int k = 0; k < 3; k++) row[(size_t)(dadr + k)*N] = ga[k];
        } else {
          for (int k = 0; k < 3; k++) { row[(size_t)(dadr + k)*N] = gear[k]; row[(size_t)(dadr + 3 + k)*N] = ga[k]; }
        }
      }
    } else if (type == MJB_TRN_SLIDERCRANK) {
      const int ids = trn[2*i + 1];
      const double rod = MD(actuator_cranklength)[i];
      double p[3], ps[3], q[4], qs[4], ms[9];
      sensor_object(c, MJB_OBJ_SITE, id, p, q);
      sensor_object(c, MJB_OBJ_SITE, ids, ps, qs);
      quat2Mat(ms, qs);
      const double axis[3] = {ms[2], ms[5], ms[8]};
      const double vec[3] = {p[0] - ps[0], p[1] - ps[1], p[2] - ps[2]};
      const double av = dot3(vec, axis);
      const double det = av*av + rod*rod - dot3(vec, vec);
      double dlda[3], dldv[3];
      if (det <= 0) {
        length = av;
        for (int k = 0; k < 3; k++) { dlda[k] = vec[k]; dldv[k] = axis[k]; }
      } else {
        const double sdet = sqrt(det);
        length = av - sdet;
        for (int k = 0; k < 3; k++) {
          dldv[k] = axis[k]*(1 - av/sdet) + vec[k]*(1/sdet);
          dlda[k] = vec[k]*(1 - av/sdet);
        }
      }
      // dl/dq = dlda . (jacr_slider x axis) + dldv . (jacp_crank - jacp_slider): the first term is the
      // torque axis x dlda on the slider's body
      double tq[3];
      cross3(tq, axis, dlda);
      trn_project(c, row, site_bodyid[id], -1, p, dldv, nullptr, gear[0]);
      trn_project(c, row, site_bodyid[ids], -1, ps, dldv, nullptr, -gear[0]);
      trn_project(c, row, site_bodyid[ids], -1, ps, nullptr, tq, gear[0]);
      length *= gear[0];
    } else if (type == MJB_TRN_TENDON) {
      const int adr = MI(tendon_adr)[id], num = MI(tendon_num)[id];
      if (MI(wrap_type)[adr] == MJB_WRAP_JOINT) {
        const int* wrap_objid = MI(wrap_objid); const double* wrap_prm = MD(wrap_prm);
        for (int j = 0; j < num; j++) {
          const int k = wrap_objid[adr + j];
          length += wrap_prm[adr + j]*QPOS(jnt_qposadr[k]);
          row[(size_t)jnt_dofadr[k]*N] += wrap_prm[adr + j]*gear[0];
        }
        length *= gear[0];
      } else {
        length = gear[0]*spatial_tendon_walk(c, id, [&](int ba, const double* pa, int bb, const double* pb,
                                                          const double* dif, double divisor) {
          trn_project(c, row, bb, -1, pb, dif, nullptr, gear[0]/divisor);
          trn_project(c, row, ba, -1, pa, dif, nullptr, -gear[0]/divisor);
        });
      }
    } else if (type == MJB_TRN_SITE) {
      const int refid = trn[2*i + 1];
      double p[3], q[4], m9[9];
      sensor_object(c, MJB_OBJ_SITE, id, p, q);
      if (refid < 0) {
        double w[6];
        quat2Mat(m9, q);
        mulMatVec3(w, m9, gear); mulMatVec3(w + 3, m9, gear + 3);
        trn_project(c, row, site_bodyid[id], -1, p, w, w + 3, 1.0);
      } else {
        // difference of the two sites' Jacobians with the columns of their common ancestors cleared
        // (:1112-1160): each chain is walked only down to the first common dof
        const int* dof_parentid = MI(dof_parentid);
        const int b0 = MI(body_weldid)[site_bodyid[id]], b1 = MI(body_weldid)[site_bodyid[refid]];
        int d0 = MI(body_dofadr)[b0] + MI(body_dofnum)[b0] - 1, d1 = MI(body_dofadr)[b1] + MI(body_dofnum)[b1] - 1;
        int common = -1;
        if (d0 >= 0 && d1 >= 0) {
          while (d0 != d1) {
            if (d0 < d1) d1 = dof_parentid[d1]; else d0 = dof_parentid[d0];
            if (d0 == -1 || d1 == -1) break;
          }
          if (d0 == d1) common = d0;
        }
        double pr[3], qr[4], mr[9], w[3];
        sensor_object(c, MJB_OBJ_SITE, refid, pr, qr);
        quat2Mat(mr, qr);
        if (gear[0] != 0 || gear[1] != 0 || gear[2] != 0) {
          double vec[3] = {p[0] - pr[0], p[1] - pr[1], p[2] - pr[2]}, lv[3];
          mulMatTVec3(lv, mr, vec);
          length += dot3(lv, gear);
          mulMatVec3(w, mr, gear);
          trn_project(c, row, site_bodyid[id], common, p, w, nullptr, 1.0);
          trn_project(c, row, site_bodyid[refid], common, pr, w, nullptr, -1.0);
        }
        if (gear[3] != 0 || gear[4] != 0 || gear[5] != 0) {
          // the reference composes the quaternions as site_quat * xquat here (:1174-1176), kept as is
          double bq[4], sq[4], rq[4], vec[3];
          ldn(bq, SC(xquat), 4*site_bodyid[id], 4);
          mulQuat(sq, MD(site_quat) + 4*id, bq);
          ldn(bq, SC(xquat), 4*site_bodyid[refid], 4);
          mulQuat(rq, MD(site_quat) + 4*refid, bq);
          subQuat(vec, sq, rq);
          length += dot3(vec, gear + 3);
          mulMatVec3(w, mr, gear + 3);
          trn_project(c, row, site_bodyid[id], common, p, nullptr, w, 1.0);
          trn_project(c, row, site_bodyid[refid], common, pr, nullptr, w, -1.0);
        }
      }
    } else if (type == MJB_TRN_BODY) {
      // adhesion (:1222-1330): minus the mean over the body's contacts (active, or excluded in the gap) of
      // the contact normal's Jacobian row, normal . (jacp(body 2) - jacp(body 1)) at the contact point. For
      // active contacts the reference forms it from the efc rows (the normal row, or the pyramid's rows
      // with equal weights, whose tangential parts cancel); the contact list is this stage's input here
      // (mjb_makeData switches the contact outputs on for models with such actuators)
      const int* geom_bodyid = MI(geom_bodyid);
      const int ncon = c.out.counts[c.s];
      int counter = 0;
      for (int k = 0; k < ncon && k < c.nconmax; k++) {
        const int b1 = geom_bodyid[c.out.contact_geom[(size_t)(2*k)*N + c.s]];
        const int b2 = geom_bodyid[c.out.contact_geom[(size_t)(2*k + 1)*N + c.s]];
        if (b1 != id && b2 != id) continue;
        const int excl = c.out.contact_info[(size_t)(3*k + 1)*N + c.s];
        if (excl != 0 && excl != 1) continue;
        counter++;
        double pos[3], nrm[3];
        for (int j = 0; j < 3; j++) {
          pos[j] = c.out.contact_num[(size_t)(13*k + 1 + j)*N + c.s];
          nrm[j] = c.out.contact_num[(size_t)(13*k + 4 + j)*N + c.s];
        }
        trn_project(c, row, b2, -1, pos, nrm, nullptr, 1.0);
        trn_project(c, row, b1, -1, pos, nrm, nullptr, -1.0);
      }
      if (counter) {
        const double sc = -1.0/counter;
        for (int j = 0; j < nv; j++) row[(size_t)j*N] *= sc;
      }
    }
    c.out.actuator_length[(size_t)i*N + c.s] = length;
    // actuator_velocity = moment . qvel over the row's non-zeros, in mju_dotSparse's order (four
    // interleaved partial sums over whole groups of four, then the tail; engine_util_sparse.h:115-157)
    int nnz = 0;
    for (int j = 0; j < nv; j++) nnz += row[(size_t)j*N] != 0;
    double r4[4] = {0, 0, 0, 0}, tail = 0;
    int k = 0;
    const int whole = nnz & ~3;
    for (int j = 0; j < nv; j++) {
      const double v = row[(size_t)j*N];
      if (v == 0) continue;
      if (k < whole) r4[k & 3] += v*QVEL(j); else { if (k == whole) tail = (r4[0] + r4[2]) + (r4[1] + r4[3]); tail += v*QVEL(j); }
      k++;
    }
    if (nnz == whole) tail = (r4[0] + r4[2]) + (r4[1] + r4[3]);
    c.out.actuator_velocity[(size_t)i*N + c.s] = tail;
  }
}

// ------------------------------------------------------------------------------------------
// mj_compareFwdInv (engine_inverse.c:275-316) for one state, after the backward sweep:
//   fwdinv[0] = | qfrc_constraint(forward) - qfrc_constraint(inverse) |
//   fwdinv[1] = | qfrc_applied + qfrc_actuator + J'*xfrc_applied - qfrc_inverse |
// J'*xfrc_applied (mj_xfrcAccumulate / mj_applyFT at xipos, engine_support.c:1194-1260) is formed
// like every other J'f on this path: the wrench about the tree origin is summed up the tree (in
// the ia rows, free by now) and projected with cdof.
MJB_HD inline void compare_fwdinv(Ctx& c) {
  const mjbHdr& H = *c.H;
  const size_t N = (size_t)c.N;
  const int* body_parentid = MI(body_parentid);
  const int* dof_bodyid = MI(dof_bodyid);
  const int* rootid = MI(body_rootid);
  double* tmp = SC(ia);
  const bool xf = c.out.fwd_xfrc != nullptr;
  if (xf) {
    for (int b = 1; b < H.nbody; b++) {
      double F[3], T[3], p[3], q[4], o[3], r[3], w[6];
      for (int k = 0; k < 3; k++) {
        F[k] = c.out.fwd_xfrc[(size_t)(6*b + k)*N + c.s];
        T[k] = c.out.fwd_xfrc[(size_t)(6*b + 3 + k)*N + c.s];
      }
      sensor_object(c, MJB_OBJ_BODY, b, p, q);
      ldn(o, SC(origin), 3*rootid[b], 3);
      r[0] = p[0] - o[0]; r[1] = p[1] - o[1]; r[2] = p[2] - o[2];
      cross3(w, r, F);
      for (int k = 0; k < 3; k++) { w[k] += T[k]; w[3 + k] = F[k]; }
      stn(tmp, 6*b, w, 6);
    }
    for (int b = H.nbody - 1; b > 0; b--) {
      const int p = body_parentid[b];
      if (!p) continue;
      double f[6], pf[6];
      ldn(f, tmp, 6*b, 6); ldn(pf, tmp, 6*p, 6);
      for (int k = 0; k < 6; k++) pf[k] += f[k];
      stn(tmp, 6*p, pf, 6);
    }
  }
  double s0 = 0, s1 = 0;
  for (int i = 0; i < H.nv; i++) {
    double qf = c.out.fwd_qforce[(size_t)i*N + c.s];
    if (xf) {
      double cd[6], f[6];
      ldn(cd, SC(cdof), 6*i, 6); ldn(f, tmp, 6*dof_bodyid[i], 6);
      qf += dot6(cd, f);
    }
    const double d1 = qf - c.out.qfrc_inverse[(size_t)i*N + c.s];
    const double d0 = c.out.fwd_qfrc_constraint[(size_t)i*N + c.s] - c.out.qfrc_constraint[(size_t)i*N + c.s];
    s0 += d0*d0; s1 += d1*d1;
  }
  // no constraint rows: the reference returns zeros without running the inverse (:283-286)
  const bool none = c.isc[MJB_ISC_NEFC * MJB_LS] == 0;
  c.out.fwdinv[c.s] = none ? 0.0 : sqrt(s0);
  c.out.fwdinv[N + c.s] = none ? 0.0 : sqrt(s1);
}


#endif  // MJB_OUTPUTS_H_
