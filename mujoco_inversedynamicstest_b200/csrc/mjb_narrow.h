// Part of the per-state mj_inverse pipeline (mjb_pipeline.h includes it inside namespace mjb, after the
// context and accessor macros; not a stand-alone header).
// Narrow phase (engine_collision_primitive.c, engine_collision_box.c), candidate scan (mj_filterSphere, tree-level culling), per-state contact processing.
#ifndef MJB_NARROW_H_
#define MJB_NARROW_H_

// ---- narrow phase: primitives of engine_collision_primitive.c -----------------------------

// mjraw_PlaneSphere (:28)
MJB_NP inline int plane_sphere(Con* con, double margin, const double* pos1, const double* mat1,
                               const double* pos2, double radius) {
  con->frame[0] = mat1[2]; con->frame[1] = mat1[5]; con->frame[2] = mat1[8];
  double tmp[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  const double cdist = dot3(tmp, con->frame);
  if (cdist > margin + radius) return 0;
  con->dist = cdist - radius;
  const double s = -con->dist/2 - radius;
  con->pos[0] = pos2[0] + con->frame[0]*s;
  con->pos[1] = pos2[1] + con->frame[1]*s;
  con->pos[2] = pos2[2] + con->frame[2]*s;
  con->frame[3] = 0; con->frame[4] = 0; con->frame[5] = 0;
  return 1;
}

// mjc_PlaneCapsule (:64)
MJB_HD inline int plane_capsule(Con* con, double margin, const double* pos1, const double* mat1,
                                const double* pos2, const double* mat2, const double* size2) {
  const double axis[3] = {mat2[2], mat2[5], mat2[8]};
  const double seg[3] = {size2[1]*axis[0], size2[1]*axis[1], size2[1]*axis[2]};
  double p[3] = {pos2[0] + seg[0], pos2[1] + seg[1], pos2[2] + seg[2]};
  const int n1 = plane_sphere(con, margin, pos1, mat1, p, size2[0]);
  p[0] = pos2[0] - seg[0]; p[1] = pos2[1] - seg[1]; p[2] = pos2[2] - seg[2];
  const int n2 = plane_sphere(con + n1, margin, pos1, mat1, p, size2[0]);
  if (n1) { con[0].frame[3] = axis[0]; con[0].frame[4] = axis[1]; con[0].frame[5] = axis[2]; }
  if (n2) { con[n1].frame[3] = axis[0]; con[n1].frame[4] = axis[1]; con[n1].frame[5] = axis[2]; }
  return n1 + n2;
}

// mjc_PlaneCylinder (:95)
MJB_HD inline int plane_cylinder(Con* con, double margin, const double* pos1, const double* mat1,
                                 const double* pos2, const double* mat2, const double* size2) {
  const double normal[3] = {mat1[2], mat1[5], mat1[8]};
  double axis[3] = {mat2[2], mat2[5], mat2[8]};
  double prjaxis = dot3(normal, axis);
  if (prjaxis > 0) { axis[0] = -axis[0]; axis[1] = -axis[1]; axis[2] = -axis[2]; prjaxis = -prjaxis; }
  double vec[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  const double dist0 = dot3(vec, normal);
  vec[0] = axis[0]*prjaxis - normal[0]; vec[1] = axis[1]*prjaxis - normal[1];
  vec[2] = axis[2]*prjaxis - normal[2];
  const double len_sqr = dot3(vec, vec);
  if (len_sqr >= MJB_MINVAL*MJB_MINVAL) {
    const double scl = size2[0]/sqrt(len_sqr);
    vec[0] *= scl; vec[1] *= scl; vec[2] *= scl;
  } else {
    vec[0] = mat2[0]*size2[0]; vec[1] = mat2[3]*size2[0]; vec[2] = mat2[6]*size2[0];
  }
  const double prjvec = dot3(vec, normal);
  axis[0] *= size2[1]; axis[1] *= size2[1]; axis[2] *= size2[1];
  prjaxis *= size2[1];

  int cnt = 0;
  if (dist0 + prjaxis + prjvec <= margin) {
    Con& q = con[cnt];
    q.dist = dist0 + prjaxis + prjvec;
    for (int k = 0; k < 3; k++) {
      q.pos[k] = pos2[k] + vec[k]; q.pos[k] += axis[k]; q.pos[k] += normal[k]*(-q.dist*0.5);
      q.frame[k] = normal[k]; q.frame[3 + k] = 0;
    }
    cnt++;
  } else {
    return 0;
  }
  if (dist0 - prjaxis + prjvec <= margin) {
    Con& q = con[cnt];
    q.dist = dist0 - prjaxis + prjvec;
    for (int k = 0; k < 3; k++) {
      q.pos[k] = pos2[k] + vec[k]; q.pos[k] -= axis[k]; q.pos[k] += normal[k]*(-q.dist*0.5);
      q.frame[k] = normal[k]; q.frame[3 + k] = 0;
    }
    cnt++;
  }
  const double prjvec1 = -prjvec*0.5;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    double vec1[3];
    cross3(vec1, vec, axis);
    normalize3(vec1);
    const double sc = size2[0]*sqrt(3.0)/2;
    vec1[0] *= sc; vec1[1] *= sc; vec1[2] *= sc;
    for (int pt = 0; pt < 2; pt++) {
      Con& q = con[cnt];
      q.dist = dist0 + prjaxis + prjvec1;
      for (int k = 0; k < 3; k++) {
        q.pos[k] = pt == 0 ? pos2[k] + vec1[k] : pos2[k] - vec1[k];
        q.pos[k] += axis[k];
        q.pos[k] += vec[k]*(-0.5);
        q.pos[k] += normal[k]*(-q.dist*0.5);
        q.frame[k] = normal[k]; q.frame[3 + k] = 0;
      }
      cnt++;
    }
  }
  return cnt;
}

// mjc_PlaneBox (:200)
MJB_HD inline int plane_box(Con* con, double margin, const double* pos1, const double* mat1,
                            const double* pos2, const double* mat2, const double* size2) {
  const double norm[3] = {mat1[2], mat1[5], mat1[8]};
  const double dif[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  const double dist = dot3(dif, norm);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    double vec[3], corner[3];
    vec[0] = (i & 1 ? size2[0] : -size2[0]);
    vec[1] = (i & 2 ? size2[1] : -size2[1]);
    vec[2] = (i & 4 ? size2[2] : -size2[2]);
    mulMatVec3(corner, mat2, vec);
    const double ldist = dot3(norm, corner);
    if (dist + ldist > margin || ldist > 0) continue;
    Con& q = con[cnt];
    q.dist = dist + ldist;
    for (int k = 0; k < 3; k++) {
      q.frame[k] = norm[k]; q.frame[3 + k] = 0;
      corner[k] += pos2[k];
      q.pos[k] = corner[k] + norm[k]*(-q.dist/2);
    }
    if (++cnt >= 4) return 4;
  }
  return cnt;
}

// mjc_PlaneConvex for an ellipsoid (engine_collision_convex.c:1045-1080; support function :570-581
// with zero margin, local direction :553, back to the global frame :700-705)
MJB_HD inline int plane_ellipsoid(Con* con, double margin, const double* pos1, const double* mat1,
                                  const double* pos2, const double* mat2, const double* size2) {
  const double normal[3] = {mat1[2], mat1[5], mat1[8]};
  const double dir[3] = {-mat1[2], -mat1[5], -mat1[8]};
  double res[3];
  for (int i = 0; i < 3; i++) {
    const double local = mat2[i]*dir[0] + mat2[3 + i]*dir[1] + mat2[6 + i]*dir[2];   // mat2' * dir
    res[i] = local * size2[i];
  }
  normalize3(res);
  for (int i = 0; i < 3; i++) res[i] *= size2[i];
  double vec[3];
  mulMatVec3(vec, mat2, res);
  vec[0] += pos2[0]; vec[1] += pos2[1]; vec[2] += pos2[2];
  const double dif[3] = {vec[0] - pos1[0], vec[1] - pos1[1], vec[2] - pos1[2]};
  const double dist = dot3(normal, dif);
  if (dist > margin) return 0;
  con->dist = dist;
  for (int k = 0; k < 3; k++) {
    con->pos[k] = vec[k] + normal[k]*(-0.5*dist);
    con->frame[k] = normal[k];
    con->frame[3 + k] = 0;
  }
  return 1;
}

// mjraw_SphereBox (engine_collision_box.c:39-106)
MJB_HD inline int sphere_box(Con* con, double margin, const double* pos1, const double* size1,
                             const double* pos2, const double* mat2, const double* size2) {
  double tmp[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  double center[3], clamped[3], deepest[3], pos[3];
  for (int i = 0; i < 3; i++) center[i] = mat2[i]*tmp[0] + mat2[3 + i]*tmp[1] + mat2[6 + i]*tmp[2];
  for (int i = 0; i < 3; i++) {
    clamped[i] = center[i];
    if (size2[i] > 0) {                                   // mju_clampVec (:22-35)
      if (clamped[i] < -size2[i]) clamped[i] = -size2[i];
      else if (clamped[i] > size2[i]) clamped[i] = size2[i];
    }
    deepest[i] = center[i];
    tmp[i] = clamped[i] - center[i];
  }
  double dist = normalize3(tmp);
  if (dist - size1[0] > margin) return 0;

  if (dist <= MJB_MINVAL) {                               // sphere centre inside the box
    double closest = (size2[0] + size2[1] + size2[2]) * 2;
    int k = 0;
    for (int i = 0; i < 6; i++) {
      const double face = fabs((i % 2 ? 1 : -1)*size2[i / 2] - center[i / 2]);
      if (closest > face) { closest = face; k = i; }
    }
    double nearest[3] = {0, 0, 0};
    nearest[k / 2] = (k % 2 ? -1 : 1);
    for (int i = 0; i < 3; i++) pos[i] = center[i] + nearest[i]*((size1[0] - closest) / 2);
    mulMatVec3(con->frame, mat2, nearest);
    dist = -closest;
  } else {
    for (int i = 0; i < 3; i++) {
      deepest[i] += tmp[i]*size1[0];
      pos[i] = 0;
      pos[i] += clamped[i]*0.5;
      pos[i] += deepest[i]*0.5;
    }
    mulMatVec3(con->frame, mat2, tmp);
  }
  double g[3];
  mulMatVec3(g, mat2, pos);
  con->pos[0] = g[0] + pos2[0]; con->pos[1] = g[1] + pos2[1]; con->pos[2] = g[2] + pos2[2];
  con->dist = dist - size1[0];
  con->frame[3] = 0; con->frame[4] = 0; con->frame[5] = 0;
  return 1;
}

// mjraw_CapsuleBox (engine_collision_box.c:121-595): the capsule's segment is brought into the box
// frame; the closest feature of the box (a face under one of the two end points, or one of the 12
// edges against the segment) gives the first contact sphere, and the relative orientation of the
// segment and that feature decides whether and where a second sphere is placed along the segment.
// Both spheres then go through sphere_box. Arithmetic follows the reference expression by
// expression (the predicates dist < bestdist etc. decide contact counts).
MJB_HD inline int capsule_box(Con* con, double margin, const double* pos1, const double* mat1,
                              const double* size1, const double* pos2, const double* mat2,
                              const double* size2) {
  const double halflength = size1[1];
  double pos[3], axis[3], halfaxis[3];
  {
    const double d[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
    const double a[3] = {mat1[2], mat1[5], mat1[8]};
    for (int i = 0; i < 3; i++) {
      pos[i] = mat2[i]*d[0] + mat2[3 + i]*d[1] + mat2[6 + i]*d[2];     // mat2' * d
      axis[i] = mat2[i]*a[0] + mat2[3 + i]*a[1] + mat2[6 + i]*a[2];
      halfaxis[i] = axis[i]*halflength;
    }
  }
  const int axisdir = (halfaxis[0] > 0 ? 1 : 0) + (halfaxis[1] > 0 ? 2 : 0) + (halfaxis[2] > 0 ? 4 : 0);

  double bestdist = margin + 2*(size1[0] + halflength + size2[0] + size2[1] + size2[2]);
  double bestsegmentpos = 0, bestboxpos = 0, secondpos = -4;
  int cltype = -4, clface = -1, clcorner = 0, cledge = 0;

  // a face of the box under one of the segment's end points
  for (int e = -1; e <= 1; e += 2) {
    double q[3], orig[3];
    int nclamp = 0, last = -1;
    for (int j = 0; j < 3; j++) {
      orig[j] = pos[j] + halfaxis[j]*e;
      q[j] = orig[j];
      if (q[j] < -size2[j]) { nclamp++; last = j; q[j] = -size2[j]; }
      else if (q[j] > size2[j]) { nclamp++; last = j; q[j] = size2[j]; }
    }
    if (nclamp > 1) continue;
    const double d[3] = {q[0] - orig[0], q[1] - orig[1], q[2] - orig[2]};
    const double dist = dot3(d, d);
    if (dist < bestdist) { bestdist = dist; bestsegmentpos = e; cltype = -2 + e; clface = last; }
  }

  // the 12 edges: edge j-direction through corner i (bit j of i clear), against the segment
  for (int j = 0; j < 3; j++) {
    for (int i = 0; i < 8; i++) {
      if (i & (1 << j)) continue;
      double start[3] = {(i & 1 ? 1 : -1)*size2[0], (i & 2 ? 1 : -1)*size2[1], (i & 4 ? 1 : -1)*size2[2]};
      start[j] = 0;
      double dif[3] = {start[0] - pos[0], start[1] - pos[1], start[2] - pos[2]};
      const double ma = size2[j]*size2[j];
      const double mb = -size2[j]*halfaxis[j];
      const double mc = size1[1]*size1[1];
      const double u = -size2[j]*dif[j];
      const double v = dot3(halfaxis, dif);
      const double det = ma*mc - mb*mb;
      if (fabs(det) < MJB_MINVAL) continue;
      const double idet = 1/det;
      double x1 = (mc*u - mb*v)*idet;      // along the edge, -1..1
      double x2 = (ma*v - mb*u)*idet;      // along the segment, -1..1
      int s1 = 1, s2 = 1;                  // 1: interior, 0 / 2: clamped to the lower / upper end
      if (x1 > 1) { x1 = 1; s1 = 2; x2 = (v - mb)*(1/mc); }
      else if (x1 < -1) { x1 = -1; s1 = 0; x2 = (v + mb)*(1/mc); }
      if (x2 > 1) {
        x2 = 1; s2 = 2; x1 = (u - mb)*(1/ma);
        if (x1 > 1) { x1 = 1; s1 = 2; } else if (x1 < -1) { x1 = -1; s1 = 0; }
      } else if (x2 < -1) {
        x2 = -1; s2 = 0; x1 = (u + mb)*(1/ma);
        if (x1 > 1) { x1 = 1; s1 = 2; } else if (x1 < -1) { x1 = -1; s1 = 0; }
      }
      for (int k = 0; k < 3; k++) dif[k] += halfaxis[k]*(-x2);
      dif[j] += size2[j]*x1;
      const double d2 = dot3(dif, dif);
      if (d2 < bestdist - MJB_MINVAL) {
        const int code = s1*3 + s2;
        bestdist = d2; bestsegmentpos = x2; bestboxpos = x1;
        clcorner = i + (1 << j)*(code / 6);
        cledge = j;
        cltype = code;
      }
    }
  }
  if (cltype == -4) return 0;

  // second sphere: how far along the segment from the first one
  bool second = true;
  double mul = 1;
  if (cltype >= 0 && cltype / 3 != 1) {
    // closest to a corner of the box
    int c1 = axisdir ^ clcorner;
    if (c1 == 0 || c1 == 7) {
      second = false;                       // pointing at / away from the corner
    } else {
      double de, dp;
      if (c1 == 1 || c1 == 2 || c1 == 4) {
        mul = 1; de = 1 - bestsegmentpos; dp = 1 + bestsegmentpos;
      } else {
        mul = -1; c1 = 7 - c1; dp = 1 - bestsegmentpos; de = 1 + bestsegmentpos;
      }
      const int ax = c1 == 1 ? 0 : (c1 == 2 ? 1 : 2);
      const int ax1 = (ax + 1) % 3, ax2 = (ax + 2) % 3;
      if (axis[ax]*axis[ax] > 0.5) {        // along the edge
        secondpos = de;
        const double e1 = 2*size2[ax] / fabs(halfaxis[ax]);
        if (e1 < secondpos) secondpos = e1;
        secondpos *= mul;
      } else {                              // along a face
        secondpos = dp;
        double e1 = 2*size2[ax1] / fabs(halfaxis[ax1]);
        if (e1 < secondpos) secondpos = e1;
        e1 = 2*size2[ax2] / fabs(halfaxis[ax2]);
        if (e1 < secondpos) secondpos = e1;
        secondpos *= -mul;
      }
    }
  } else if (cltype >= 0) {
    // closest to the interior of an edge: T configuration (no second point) or a cross
    int c1 = (axisdir ^ clcorner) & (7 - (1 << cledge));
    if (c1 != 1 && c1 != 2 && c1 != 4) {
      second = false;
    } else {
      const int ax = cledge;
      int ax1 = (ax + 1) % 3, ax2 = (ax + 2) % 3;
      if (fabs(axis[ax1]) > fabs(axis[ax2])) ax1 = ax2;
      ax2 = 3 - ax - ax1;
      if (c1 & (1 << ax2)) { mul = 1; secondpos = 1 - bestsegmentpos; }
      else { mul = -1; secondpos = 1 + bestsegmentpos; }
      double e1 = 2*size2[ax2] / fabs(halfaxis[ax2]);
      if (e1 < secondpos) secondpos = e1;
      const double e2 = (((axisdir & (1 << ax)) != 0) == ((c1 & (1 << ax2)) != 0)) ? 1 - bestboxpos
                                                                                  : 1 + bestboxpos;
      e1 = size2[ax]*e2 / fabs(halfaxis[ax]);
      if (e1 < secondpos) secondpos = e1;
      secondpos *= mul;
    }
  } else {
    // an end point above a face: walk towards the other end while still above the box
    if (clface == -1) {
      second = false;                       // the end point is inside the box
    } else {
      mul = cltype == -3 ? 1 : -1;
      secondpos = 2;
      const double t[3] = {pos[0] + halfaxis[0]*(-mul), pos[1] + halfaxis[1]*(-mul), pos[2] + halfaxis[2]*(-mul)};
      for (int i = 0; i < 3; i++) {
        if (i == clface) continue;
        double e1 = (size2[i] - t[i]) / halfaxis[i] * mul;
        if (e1 > 0 && e1 < secondpos) secondpos = e1;
        e1 = (-size2[i] - t[i]) / halfaxis[i] * mul;
        if (e1 > 0 && e1 < secondpos) secondpos = e1;
      }
      secondpos *= mul;
    }
  }
  (void)second;   // the reference tests secondpos itself (> -3 once assigned)

  double loc[3], cen[3];
  for (int k = 0; k < 3; k++) loc[k] = pos[k] + halfaxis[k]*bestsegmentpos;
  mulMatVec3(cen, mat2, loc);
  cen[0] += pos2[0]; cen[1] += pos2[1]; cen[2] += pos2[2];
  int n = sphere_box(con, margin, cen, size1, pos2, mat2, size2);
  if (secondpos > -3) {
    for (int k = 0; k < 3; k++) loc[k] = pos[k] + halfaxis[k]*(secondpos + bestsegmentpos);
    mulMatVec3(cen, mat2, loc);
    cen[0] += pos2[0]; cen[1] += pos2[1]; cen[2] += pos2[2];
    n += sphere_box(con + n, margin, cen, size1, pos2, mat2, size2);
  }
  return n;
}

// mjraw_SphereSphere (:250)
MJB_NP inline int sphere_sphere(Con* con, double margin, const double* pos1, const double* mat1,
                                double r1, const double* pos2, const double* mat2, double r2) {
  const double dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  const double cdist_sqr = dot3(dif, dif);
  const double min_dist = margin + r1 + r2;
  if (cdist_sqr > min_dist*min_dist) return 0;
  con->dist = sqrt(cdist_sqr) - r1 - r2;
  con->frame[0] = pos2[0] - pos1[0]; con->frame[1] = pos2[1] - pos1[1]; con->frame[2] = pos2[2] - pos1[2];
  const double len = normalize3(con->frame);
  if (len < MJB_MINVAL) {
    const double axis1[3] = {mat1[2], mat1[5], mat1[8]};
    const double axis2[3] = {mat2[2], mat2[5], mat2[8]};
    cross3(con->frame, axis1, axis2);
    normalize3(con->frame);
  }
  const double s = r1 + con->dist/2;
  con->pos[0] = con->frame[0]*s + pos1[0];
  con->pos[1] = con->frame[1]*s + pos1[1];
  con->pos[2] = con->frame[2]*s + pos1[2];
  con->frame[3] = 0; con->frame[4] = 0; con->frame[5] = 0;
  return 1;
}

MJB_DI double clip(double x, double lo, double hi) {  // mju_clip
  return fmax(lo, fmin(hi, x));
}

// mjraw_SphereCapsule (:295)
MJB_HD inline int sphere_capsule(Con* con, double margin, const double* pos1, const double* mat1,
                                 const double* size1, const double* pos2, const double* mat2,
                                 const double* size2) {
  const double len = size2[1];
  const double axis[3] = {mat2[2], mat2[5], mat2[8]};
  double vec[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  const double x = clip(dot3(axis, vec), -len, len);
  vec[0] = axis[0]*x + pos2[0]; vec[1] = axis[1]*x + pos2[1]; vec[2] = axis[2]*x + pos2[2];
  return sphere_sphere(con, margin, pos1, mat1, size1[0], vec, mat2, size2[0]);
}

// mjc_SphereCylinder (:324)
MJB_HD inline int sphere_cylinder(Con* con, double margin, const double* pos1, const double* mat1,
                                  const double* size1, const double* pos2, const double* mat2,
                                  const double* size2) {
  const double radius = size2[0], height = size2[1];
  const double axis[3] = {mat2[2], mat2[5], mat2[8]};
  double vec[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  const double x = dot3(axis, vec);
  double a_proj[3] = {axis[0]*x, axis[1]*x, axis[2]*x};
  double p_proj[3] = {vec[0] - a_proj[0], vec[1] - a_proj[1], vec[2] - a_proj[2]};
  const double p_proj_sqr = dot3(p_proj, p_proj);
  int collide_side = fabs(x) < height;
  int collide_cap = p_proj_sqr < radius*radius;
  if (collide_side && collide_cap) {
    const double dist_cap = height - fabs(x);
    const double dist_radius = radius - sqrt(p_proj_sqr);
    if (dist_cap < dist_radius) collide_side = 0; else collide_cap = 0;
  }
  if (collide_side) {
    a_proj[0] += pos2[0]; a_proj[1] += pos2[1]; a_proj[2] += pos2[2];
    return sphere_sphere(con, margin, pos1, mat1, size1[0], a_proj, mat2, size2[0]);
  }
  if (collide_cap) {
    double flipmat[9] = {-mat2[0], mat2[1], -mat2[2], -mat2[3], mat2[4], -mat2[5],
                         -mat2[6], mat2[7], -mat2[8]};
    double pos_cap[3];
    const double hs = x > 0 ? height : -height;
    pos_cap[0] = pos2[0] + axis[0]*hs; pos_cap[1] = pos2[1] + axis[1]*hs; pos_cap[2] = pos2[2] + axis[2]*hs;
    const int n = plane_sphere(con, margin, pos_cap, x > 0 ? mat2 : flipmat, pos1, size1[0]);
    if (n) { con->frame[0] = -con->frame[0]; con->frame[1] = -con->frame[1]; con->frame[2] = -con->frame[2]; }
    return n;
  }
  const double scl = size2[0] / sqrt(p_proj_sqr);
  const double hs = x > 0 ? height : -height;
  for (int k = 0; k < 3; k++) {
    p_proj[k] *= scl;
    vec[k] = axis[k]*hs;
    vec[k] += p_proj[k];
    vec[k] += pos2[k];
  }
  return sphere_sphere(con, margin, pos1, mat1, size1[0], vec, mat2, 0.0);
}

// mjraw_CapsuleCapsule (:398)
MJB_HD inline int capsule_capsule(Con* con, double margin, const double* pos1, const double* mat1,
                                  const double* size1, const double* pos2, const double* mat2,
                                  const double* size2) {
  const double axis1[3] = {mat1[2]*size1[1], mat1[5]*size1[1], mat1[8]*size1[1]};
  const double axis2[3] = {mat2[2]*size2[1], mat2[5]*size2[1], mat2[8]*size2[1]};
  const double dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  const double ma = dot3(axis1, axis1);
  const double mb = -dot3(axis1, axis2);
  const double mc = dot3(axis2, axis2);
  const double u = -dot3(axis1, dif);
  const double v = dot3(axis2, dif);
  const double det = ma*mc - mb*mb;
  double vec1[3], vec2[3];

  if (fabs(det) >= MJB_MINVAL) {
    double x1 = (mc*u - mb*v) / det;
    double x2 = (ma*v - mb*u) / det;
    if (x1 > 1) { x1 = 1; x2 = (v - mb) / mc; }
    else if (x1 < -1) { x1 = -1; x2 = (v + mb) / mc; }
    if (x2 > 1) { x2 = 1; x1 = clip((u - mb) / ma, -1, 1); }
    else if (x2 < -1) { x2 = -1; x1 = clip((u + mb) / ma, -1, 1); }
    for (int k = 0; k < 3; k++) {
      vec1[k] = axis1[k]*x1 + pos1[k];
      vec2[k] = axis2[k]*x2 + pos2[k];
    }
    return sphere_sphere(con, margin, vec1, mat1, size1[0], vec2, mat2, size2[0]);
  }

  // parallel axes
  for (int k = 0; k < 3; k++) vec1[k] = pos1[k] + axis1[k];
  double x2 = clip((v - mb) / mc, -1, 1);
  for (int k = 0; k < 3; k++) vec2[k] = axis2[k]*x2 + pos2[k];
  const int n1 = sphere_sphere(con, margin, vec1, mat1, size1[0], vec2, mat2, size2[0]);

  for (int k = 0; k < 3; k++) vec1[k] = pos1[k] - axis1[k];
  x2 = clip((v + mb) / mc, -1, 1);
  for (int k = 0; k < 3; k++) vec2[k] = axis2[k]*x2 + pos2[k];
  const int n2 = sphere_sphere(con + n1, margin, vec1, mat1, size1[0], vec2, mat2, size2[0]);
  if (n1 + n2 >= 2) return n1 + n2;

  for (int k = 0; k < 3; k++) vec2[k] = pos2[k] + axis2[k];
  double x1 = clip((u - mb) / ma, -1, 1);
  for (int k = 0; k < 3; k++) vec1[k] = axis1[k]*x1 + pos1[k];
  const int n3 = sphere_sphere(con + n1 + n2, margin, vec1, mat1, size1[0], vec2, mat2, size2[0]);
  if (n1 + n2 + n3 >= 2) return n1 + n2 + n3;

  for (int k = 0; k < 3; k++) vec2[k] = pos2[k] - axis2[k];
  x1 = clip((u + mb) / ma, -1, 1);
  for (int k = 0; k < 3; k++) vec1[k] = axis1[k]*x1 + pos1[k];
  const int n4 = sphere_sphere(con + n1 + n2 + n3, margin, vec1, mat1, size1[0], vec2, mat2, size2[0]);
  return n1 + n2 + n3 + n4;
}

// mjc_BoxBox (engine_collision_box.c:607-1343) followed by the driver's removal of bad and repeated
// contacts (engine_collision_driver.c:1522-1590, mju_outsideBox engine_util_misc.c:911).
//
// Structure of the reference: (A) separating-axis search over the 6 face normals and the 9 edge x
// edge directions, keeping the axis of least penetration; (B) face case: the other box's closest
// face polygon is clipped against the reference face (edge/edge intersections, reference corners
// inside the polygon, polygon corners inside the face) and points above the margin are dropped;
// (C) edge case: the same clipping for the quadrilateral spanned by the two closest edges of box 2,
// projected along the separating direction. The predicates decide how many contacts a pair yields,
// so every expression keeps the reference's operand order.
#define MJB_BOXBOX_MAXCON 24   // 16 edge clips + 4 + 4 corner points

struct BoxAxes { int i0, i1, i2; double f0, f1, f2; };

// axis permutation / signs that turn face `face` (0..2: +x,+y,+z of the frame, 3..5: the negatives)
// into the local +z direction (the reference's rotmore matrices and rotaxis / rotmatx macros)
MJB_DI BoxAxes box_face_axes(int face) {
  BoxAxes a = {0, 1, 2, 1, 1, 1};
  if (face == 0) { a.i0 = 2; a.f0 = -1; a.i2 = 0; }
  else if (face == 1) { a.i1 = 2; a.f1 = -1; a.i2 = 1; }
  else if (face == 3) { a.i0 = 2; a.i2 = 0; a.f2 = -1; }
  else if (face == 4) { a.i1 = 2; a.i2 = 1; a.f2 = -1; }
  else if (face == 5) { a.f0 = -1; a.f2 = -1; }
  return a;
}
MJB_DI void box_rotmore(double* m, int face) {
  for (int k = 0; k < 9; k++) m[k] = 0;
  if (face == 0) { m[2] = -1; m[4] = 1; m[6] = 1; }
  else if (face == 1) { m[0] = 1; m[5] = -1; m[7] = 1; }
  else if (face == 2) { m[0] = 1; m[4] = 1; m[8] = 1; }
  else if (face == 3) { m[2] = 1; m[4] = 1; m[6] = -1; }
  else if (face == 4) { m[0] = 1; m[5] = 1; m[7] = -1; }
  else { m[0] = -1; m[4] = 1; m[8] = -1; }
}
MJB_DI void box_rotaxis(double* res, const double* v, const BoxAxes& a) {
  const double r0 = v[a.i0]*a.f0, r1 = v[a.i1]*a.f1, r2 = v[a.i2]*a.f2;
  res[0] = r0; res[1] = r1; res[2] = r2;
}
MJB_DI void box_rotmatx(double* res, const double* m, const BoxAxes& a) {
  for (int k = 0; k < 3; k++) {
    res[k] = m[3*a.i0 + k]*a.f0; res[3 + k] = m[3*a.i1 + k]*a.f1; res[6 + k] = m[3*a.i2 + k]*a.f2;
  }
}
MJB_DI void mulMatTVec3(double* res, const double* m, const double* v) {  // engine_util_blas.c:179
  const double t0 = m[0]*v[0] + m[3]*v[1] + m[6]*v[2];
  const double t1 = m[1]*v[0] + m[4]*v[1] + m[7]*v[2];
  const double t2 = m[2]*v[0] + m[5]*v[1] + m[8]*v[2];
  res[0] = t0; res[1] = t1; res[2] = t2;
}
MJB_DI void mulMatMatT3(double* res, const double* a, const double* b) {  // engine_util_blas.c:223
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) res[3*i + j] = a[3*i]*b[3*j] + a[3*i + 1]*b[3*j + 1] + a[3*i + 2]*b[3*j + 2];
}

// clip the segment (o, o + d) of the plane z = const against the rectangle |x| <= lim[0],
// |y| <= lim[1]: parameters c1 in [0, 1] where it crosses the four border lines (the reference's
// "lines" loop); calls emit(c1, q, l, c2) for every crossing inside the border segment
template <typename F>
MJB_DI void box_clip_line(const double* line, const double* lim, F emit) {
  for (int q = 0; q < 2; q++) {
    const double a = line[q], b = line[3 + q], c = line[1 - q], d = line[4 - q];
    if (fabs(b) > MJB_MINVAL) {
      for (int j = -1; j <= 1; j += 2) {
        const double l = lim[q]*j;
        const double c1 = (l - a)*(1/b);
        if (c1 < 0 || c1 > 1) continue;
        const double c2 = c + d*c1;
        if (fabs(c2) > lim[1 - q]) continue;
        emit(c1, q, l, c2);
      }
    }
  }
}

// mju_outsideBox (engine_util_misc.c:911) with inflate = 1.01
MJB_DI int outside_box(const double* point, const double* pos, const double* mat, const double* size) {
  const double inflate = 1.01;
  double vec[3] = {point[0] - pos[0], point[1] - pos[1], point[2] - pos[2]};
  mulMatTVec3(vec, mat, vec);
  const double big[3] = {size[0]*inflate, size[1]*inflate, size[2]*inflate};
  if (vec[0] > big[0] || vec[0] < -big[0] || vec[1] > big[1] || vec[1] < -big[1] ||
      vec[2] > big[2] || vec[2] < -big[2]) return 1;
  const double small[3] = {size[0]/inflate, size[1]/inflate, size[2]/inflate};
  if (vec[0] < small[0] && vec[0] > -small[0] && vec[1] < small[1] && vec[1] > -small[1] &&
      vec[2] < small[2] && vec[2] > -small[2]) return -1;
  return 0;
}

MJB_HD inline int box_box_raw(Con* con, double margin, const double* pos1, const double* mat1,
                              const double* size1, const double* pos2, const double* mat2,
                              const double* size2) {
  double pos21[3], pos12[3], rot[9], rott[9], rotabs[9], rottabs[9], plen1[3], plen2[3];
  double points[MJB_BOXBOX_MAXCON][3], depth[MJB_BOXBOX_MAXCON];
  double clnorm[3] = {0, 0, 0};
  int n = 0, code = -1, cle1 = 0, cle2 = 0, in = 0;
  const double margin2 = margin*margin;
  {
    double t[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
    mulMatTVec3(pos21, mat1, t);
    t[0] = pos1[0] - pos2[0]; t[1] = pos1[1] - pos2[1]; t[2] = pos1[2] - pos2[2];
    mulMatTVec3(pos12, mat2, t);
  }
  for (int i = 0; i < 3; i++)        // rot = mat1' * mat2  (engine_util_blas.c:208)
    for (int j = 0; j < 3; j++)
      rot[3*i + j] = mat1[i]*mat2[j] + mat1[3 + i]*mat2[3 + j] + mat1[6 + i]*mat2[6 + j];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rott[3*j + i] = rot[3*i + j];
  for (int i = 0; i < 9; i++) { rotabs[i] = fabs(rot[i]); rottabs[i] = fabs(rott[i]); }
  mulMatVec3(plen2, rotabs, size2);
  mulMatTVec3(plen1, rotabs, size1);

  // (A) least-penetration axis: face normals of box 1 (code 0..5) and of box 2 (6..11)
  double penetration = margin;
  for (int i = 0; i < 3; i++) penetration += size1[i]*3 + size2[i]*3;
  for (int i = 0; i < 3; i++) {
    const double c1 = -fabs(pos21[i]) + size1[i] + plen2[i];
    const double c2 = -fabs(pos12[i]) + size2[i] + plen1[i];
    if (c1 < -margin || c2 < -margin) return 0;
    if (c1 < penetration) { penetration = c1; code = i + 3*(pos21[i] < 0) + 0; }
    if (c2 < penetration) { penetration = c2; code = i + 3*(pos12[i] < 0) + 6; }
  }
  // edge i of box 1 x edge j of box 2 (code 12 + 3i + j)
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) {
      double ax[3] = {0, 0, 0};
      if (i == 0) { ax[1] = -rott[3*j + 2]; ax[2] = +rott[3*j + 1]; }
      else if (i == 1) { ax[0] = +rott[3*j + 2]; ax[2] = -rott[3*j + 0]; }
      else { ax[0] = -rott[3*j + 1]; ax[1] = +rott[3*j + 0]; }
      const double c1 = normalize3(ax);
      if (c1 < MJB_MINVAL) continue;
      const double c2 = dot3(pos21, ax);
      double c3 = 0;
      for (int k = 0; k < 3; k++) if (k != i) c3 += size1[k]*fabs(ax[k]);
      for (int k = 0; k < 3; k++) if (k != j) c3 += size2[k]*rotabs[3*i + 3 - k - j]/c1;
      c3 -= fabs(c2);
      if (c3 < -margin) return 0;
      if (c3 < penetration*(1 - 1e-12)) {
        penetration = c3;
        cle1 = 0;
        for (int k = 0; k < 3; k++) if (k != i) if ((ax[k] > 0) ^ (c2 < 0)) cle1 += 1 << k;
        cle2 = 0;
        for (int k = 0; k < 3; k++)
          if (k != j) if ((rot[3*i + 3 - k - j] > 0) ^ (c2 < 0) ^ ((k - j + 3) % 3 == 1)) cle2 += 1 << k;
        code = 12 + i*3 + j;
        clnorm[0] = ax[0]; clnorm[1] = ax[1]; clnorm[2] = ax[2];
        in = c2 < 0;
      }
    }
  }
  if (code == -1) return 0;

  if (code < 12) {
    // (B) a face of box (q2 ? 2 : 1) is the reference face, turned to local +z
    const int q1 = code % 6, q2 = code / 6;
    const BoxAxes A = box_face_axes(q1);
    double rotmore[9], r[9], rt[9], p[3], tmp1[3], s[3];
    box_rotmore(rotmore, q1);
    if (q2) {
      mulMatMatT3(r, rotmore, rot);
      box_rotaxis(p, pos12, A); box_rotaxis(tmp1, size2, A);
      s[0] = size1[0]; s[1] = size1[1]; s[2] = size1[2];
    } else {
      box_rotmatx(r, rot, A);
      box_rotaxis(p, pos21, A); box_rotaxis(tmp1, size1, A);
      s[0] = size2[0]; s[1] = size2[1]; s[2] = size2[2];
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rt[3*j + i] = r[3*i + j];
    const double ss[3] = {fabs(tmp1[0]), fabs(tmp1[1]), fabs(tmp1[2])};
    const double lx = ss[0], ly = ss[1], hz = ss[2];
    p[2] -= hz;
    int clcorner = 0;
    for (int i = 0; i < 3; i++) if (r[6 + i] < 0) clcorner += 1 << i;
    double pts[6][3];
    for (int a = 0; a < 6; a++) for (int k = 0; k < 3; k++) pts[a][k] = 0;
    for (int k = 0; k < 3; k++) pts[0][k] = p[k];
    for (int a = 0; a < 3; a++) {
      const double sc = s[a]*((clcorner & (1 << a)) ? 1 : -1);
      for (int k = 0; k < 3; k++) pts[0][k] += rt[3*a + k]*sc;
    }
    int m = 1;
    for (int i = 0; i < 3; i++) {
      if (fabs(r[6 + i]) < 0.5) {
        const double sc = s[i]*((clcorner & (1 << i)) ? -2 : 2);
        for (int k = 0; k < 3; k++) pts[m][k] = rt[3*i + k]*sc;
        m++;
      }
    }
    for (int k = 0; k < 3; k++) {
      pts[3][k] = pts[0][k] + pts[1][k];
      pts[4][k] = pts[0][k] + pts[2][k];
      pts[5][k] = pts[3][k] + pts[2][k];
    }
    double lines[4][6];
    int nlines = 0;
    auto set_line = [&](const double* o, const double* d) {
      for (int k = 0; k < 3; k++) { lines[nlines][k] = o[k]; lines[nlines][3 + k] = d[k]; }
      nlines++;
    };
    if (m > 1) set_line(pts[0], pts[1]);
    if (m > 2) { set_line(pts[0], pts[2]); set_line(pts[3], pts[2]); set_line(pts[4], pts[1]); }
    for (int i = 0; i < nlines; i++) {
      const double* L = lines[i];
      box_clip_line(L, ss, [&](double c1, int, double, double) {
        for (int k = 0; k < 3; k++) points[n][k] = L[k] + L[3 + k]*c1;
        n++;
      });
    }
    {
      const double a = pts[1][0], b = pts[2][0], c = pts[1][1], d = pts[2][1];
      const double c1 = a*d - b*c;
      if (m > 2) {
        for (int i = 0; i < 4; i++) {
          const double llx = i / 2 ? lx : -lx, lly = i % 2 ? ly : -ly;
          const double x = llx - pts[0][0], y = lly - pts[0][1];
          const double u = (x*d - y*b)*(1/c1), v = (y*a - x*c)*(1/c1);
          if (u <= 0 || v <= 0 || u >= 1 || v >= 1) continue;
          points[n][0] = llx; points[n][1] = lly;
          points[n][2] = (pts[0][2] + u*pts[1][2] + v*pts[2][2]);
          n++;
        }
      }
    }
    for (int i = 0; i < (1 << (m - 1)); i++) {
      const double* t = pts[i == 0 ? 0 : i + 2];
      if (i) if (t[0] <= -lx || t[0] >= lx) continue;
      if (i) if (t[1] <= -ly || t[1] >= ly) continue;
      for (int k = 0; k < 3; k++) points[n][k] = t[k];
      n++;
    }
    const int cand = n;
    n = 0;
    for (int i = 0; i < cand; i++) {
      if (points[i][2] > margin) continue;
      for (int k = 0; k < 3; k++) points[n][k] = points[i][k];
      depth[n] = points[n][2];
      points[n][2] *= 0.5;
      n++;
    }
    mulMatMatT3(r, q2 ? mat2 : mat1, rotmore);
    const double* pc = q2 ? pos2 : pos1;
    const double sg = q2 ? -1 : 1;
    const double nrm[3] = {sg*r[2], sg*r[5], sg*r[8]};
    for (int i = 0; i < n; i++) {
      con[i].dist = points[i][2];        // as the reference: the halved coordinate, not depth[i]
      points[i][2] += hz;
      double g[3];
      mulMatVec3(g, r, points[i]);
      for (int k = 0; k < 3; k++) { con[i].pos[k] = g[k] + pc[k]; con[i].frame[k] = nrm[k]; con[i].frame[3 + k] = 0; }
    }
    (void)depth;
    return n;
  }

  // (C) edge i of box 1 against edge j of box 2
  code -= 12;
  const int q1 = code / 3, q2 = code % 3;
  int ax1 = q2 == 0 ? 1 : (q2 == 1 ? 0 : 1), ax2 = q2 == 0 ? 2 : (q2 == 1 ? 2 : 0);
  int pax1 = q1 == 0 ? 1 : (q1 == 1 ? 0 : 1), pax2 = q1 == 0 ? 2 : (q1 == 1 ? 2 : 0);
  if (rotabs[3*q1 + ax1] < rotabs[3*q1 + ax2]) { ax1 = ax2; ax2 = 3 - q2 - ax1; }
  if (rottabs[3*q2 + pax1] < rottabs[3*q2 + pax2]) { pax1 = pax2; pax2 = 3 - q1 - pax1; }
  const int clface = (cle1 & (1 << pax2)) ? pax2 : pax2 + 3;
  const BoxAxes A = box_face_axes(clface);
  double rotmore[9], r[9], rt[9], p[3], rnorm[3], s[3];
  box_rotmore(rotmore, clface);
  box_rotaxis(p, pos21, A);
  box_rotaxis(rnorm, clnorm, A);
  box_rotmatx(r, rot, A);
  {
    double t[3];
    mulMatTVec3(t, rotmore, size1);
    s[0] = fabs(t[0]); s[1] = fabs(t[1]); s[2] = fabs(t[2]);
  }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rt[3*j + i] = r[3*i + j];
  const double lx = s[0], ly = s[1], hz = s[2];
  p[2] -= hz;

  // the four end points of the two closest edges of box 2 (direction q2)
  for (int e = 0; e < 2; e++) {
    const double s1 = size2[ax1]*(((cle2 & (1 << ax1)) != 0) == (e == 0) ? 1 : -1);
    const double s2 = size2[ax2]*((cle2 & (1 << ax2)) ? 1 : -1);
    double base[3];
    for (int k = 0; k < 3; k++) base[k] = p[k];
    for (int k = 0; k < 3; k++) base[k] += rt[3*ax1 + k]*s1;
    for (int k = 0; k < 3; k++) base[k] += rt[3*ax2 + k]*s2;
    for (int k = 0; k < 3; k++) {
      points[2*e][k] = base[k] + rt[3*q2 + k]*size2[q2];
      points[2*e + 1][k] = base[k] + rt[3*q2 + k]*(-size2[q2]);
    }
  }
  double axi[3][3], pu[4][3], ppts2[4][2], pts[3][3];
  for (int k = 0; k < 3; k++) {
    axi[0][k] = points[0][k];
    axi[1][k] = points[1][k] - points[0][k];
    axi[2][k] = points[2][k] - points[0][k];
  }
  if (fabs(rnorm[2]) < MJB_MINVAL) return 0;
  const double innorm = (1/rnorm[2])*(in ? -1 : 1);
  for (int i = 0; i < 4; i++) {
    const double c1 = -points[i][2]*(1/rnorm[2]);
    for (int k = 0; k < 3; k++) pu[i][k] = points[i][k];
    for (int k = 0; k < 3; k++) points[i][k] += rnorm[k]*c1;
    ppts2[i][0] = points[i][0]; ppts2[i][1] = points[i][1];
  }
  for (int k = 0; k < 3; k++) {
    pts[0][k] = points[0][k];
    pts[1][k] = points[1][k] - points[0][k];
    pts[2][k] = points[2][k] - points[0][k];
  }
  double lines[4][6], linesu[4][6];
  for (int k = 0; k < 3; k++) {
    lines[0][k] = pts[0][k]; lines[0][3 + k] = pts[1][k];
    linesu[0][k] = axi[0][k]; linesu[0][3 + k] = axi[1][k];
    lines[1][k] = pts[0][k]; lines[1][3 + k] = pts[2][k];
    linesu[1][k] = axi[0][k]; linesu[1][3 + k] = axi[2][k];
    lines[2][k] = pts[0][k] + pts[1][k]; lines[2][3 + k] = pts[2][k];
    linesu[2][k] = axi[0][k] + axi[1][k]; linesu[2][3 + k] = axi[2][k];
    lines[3][k] = pts[0][k] + pts[2][k]; lines[3][3 + k] = pts[1][k];
    linesu[3][k] = axi[0][k] + axi[2][k]; linesu[3][3 + k] = axi[1][k];
  }
  n = 0;
  for (int i = 0; i < 4; i++) {
    const double* LU = linesu[i];
    box_clip_line(lines[i], s, [&](double c1, int q, double l, double c2) {
      if ((LU[2] + LU[5]*c1)*innorm > margin) return;
      for (int k = 0; k < 3; k++) points[n][k] = LU[k]*0.5;
      for (int k = 0; k < 3; k++) points[n][k] += LU[3 + k]*(0.5*c1);
      points[n][0 + q] += 0.5*l;
      points[n][1 - q] += 0.5*c2;
      depth[n] = points[n][2]*innorm*2;
      n++;
    });
  }
  const int nl = n;
  {
    const double a = pts[1][0], b = pts[2][0], c = pts[1][1], d = pts[2][1];
    // c1 starts as the determinant of the quadrilateral's edge vectors and is REUSED below for the
    // squared distance, exactly like the reference (:1229-1277): once a corner gets that far, the
    // following corners are tested with the overwritten value. Kept for identical contact sets.
    double c1 = a*d - b*c;
    for (int i = 0; i < 4; i++) {
      const double llx = i / 2 ? lx : -lx, lly = i % 2 ? ly : -ly;
      const double x = llx - pts[0][0], y = lly - pts[0][1];
      double u = (x*d - y*b)*(1/c1), v = (y*a - x*c)*(1/c1);
      if (nl == 0) {
        if ((u < 0 || u > 1) && (v < 0 || v > 1)) continue;
      } else {
        if (u < 0 || u > 1 || v < 0 || v > 1) continue;
      }
      if (u < 0) u = 0;
      if (u > 1) u = 1;
      if (v < 0) v = 0;
      if (v > 1) v = 1;
      double t[3];
      for (int k = 0; k < 3; k++) t[k] = pu[0][k]*(1 - u - v);
      for (int k = 0; k < 3; k++) t[k] += pu[1][k]*u;
      for (int k = 0; k < 3; k++) t[k] += pu[2][k]*v;
      points[n][0] = llx; points[n][1] = lly; points[n][2] = 0;
      const double df[3] = {points[n][0] - t[0], points[n][1] - t[1], points[n][2] - t[2]};
      c1 = dot3(df, df);
      if (t[2] > 0) if (c1 > margin2) continue;
      for (int k = 0; k < 3; k++) points[n][k] = (points[n][k] + t[k])*0.5;
      depth[n] = sqrt(c1)*(t[2] < 0 ? -1 : 1);
      n++;
    }
  }
  const int nf = n;
  for (int i = 0; i < 4; i++) {
    const double x = ppts2[i][0], y = ppts2[i][1];
    if (nl == 0) {
      if (nf != 0) if (x < -lx || x > lx) if (y < -ly || y > ly) continue;
    } else {
      if (x < -lx || x > lx || y < -ly || y > ly) continue;
    }
    double c1 = 0;
    for (int j = 0; j < 2; j++) {
      if (ppts2[i][j] < -s[j]) c1 += (ppts2[i][j] + s[j])*(ppts2[i][j] + s[j]);
      else if (ppts2[i][j] > s[j]) c1 += (ppts2[i][j] - s[j])*(ppts2[i][j] - s[j]);
    }
    c1 += pu[i][2]*innorm*pu[i][2]*innorm;
    if (pu[i][2] > 0) if (c1 > margin2) continue;
    double t[3] = {ppts2[i][0]*0.5, ppts2[i][1]*0.5, 0};
    for (int j = 0; j < 2; j++) {
      if (ppts2[i][j] < -s[j]) t[j] = -s[j]*0.5;
      else if (ppts2[i][j] > s[j]) t[j] = +s[j]*0.5;
    }
    for (int k = 0; k < 3; k++) points[n][k] = t[k] + pu[i][k]*0.5;
    depth[n] = sqrt(c1)*(pu[i][2] < 0 ? -1 : 1);
    n++;
  }
  mulMatMatT3(r, mat1, rotmore);
  double wn[3];
  mulMatVec3(wn, r, rnorm);
  const double sg = in ? -1 : 1;
  for (int i = 0; i < n; i++) {
    con[i].dist = depth[i];
    points[i][2] += hz;
    double g[3];
    mulMatVec3(g, r, points[i]);
    for (int k = 0; k < 3; k++) { con[i].pos[k] = g[k] + pos1[k]; con[i].frame[k] = wn[k]*sg; con[i].frame[3 + k] = 0; }
  }
  return n;
}

// box-box with the driver's clean-up: contacts outside one box and not inside the other are bad,
// of two contacts at exactly the same position the earlier one is dropped
MJB_COLD inline int box_box(Con* con, double margin, const double* pos1, const double* mat1,
                          const double* size1, const double* pos2, const double* mat2,
                          const double* size2) {
  const int num = box_box_raw(con, margin, pos1, mat1, size1, pos2, mat2, size2);
  const double sz1[3] = {size1[0] + margin, size1[1] + margin, size1[2] + margin};
  const double sz2[3] = {size2[0] + margin, size2[1] + margin, size2[2] + margin};
  unsigned bad = 0;
  for (int i = 0; i < num; i++) {
    const int out1 = outside_box(con[i].pos, pos1, mat1, sz1);
    const int out2 = outside_box(con[i].pos, pos2, mat2, sz2);
    if ((out1 == 1 && out2 != -1) || (out2 == 1 && out1 != -1)) bad |= 1u << i;
  }
  for (int i = 0; i < num - 1; i++) {
    if (bad & (1u << i)) continue;
    for (int j = i + 1; j < num; j++) {
      if (bad & (1u << j)) continue;
      if (con[i].pos[0] == con[j].pos[0] && con[i].pos[1] == con[j].pos[1] && con[i].pos[2] == con[j].pos[2]) {
        bad |= 1u << i;
        break;
      }
    }
  }
  int k = 0;
  for (int j = 0; j < num; j++) {
    if (bad & (1u << j)) continue;
    if (k < j) con[k] = con[j];
    k++;
  }
  return k;
}

// narrow phase of candidate pair ci on the state bound to c; contact frames are completed
// (mju_makeFrame) before returning. Returns the number of contacts (<= MJB_MAXCON_PAIR).
// kSimple: the model's pairs are all plane / sphere / capsule against sphere / capsule (mjbHdr::simple_pairs),
// so that at most two contacts come back and `con` can stay in registers (no dynamically indexed primitive
// is compiled in).
// kConvex: the GJK / EPA pairs are compiled in (their polytope takes ~60 KB of the thread's stack): only the kernel
// instantiations that models with such pairs launch, and the CPU build
#if defined(__CUDACC__)
#define MJB_CONVEX_DEFAULT false
#else
#define MJB_CONVEX_DEFAULT true
#endif
template <bool kSimple = false, bool kConvex = MJB_CONVEX_DEFAULT>
MJB_HD inline int narrow_pair(Ctx& c, int ci, Con* con) {
  const int* cint = MI(cand_int) + MJB_CAND_NI*ci;
  const double* cn = MD(cand_num) + MJB_CAND_NN*ci;
  const double* geom_size = MD(geom_size);
  double* gxmat = SC(geom_xmat);
  const int g1 = cint[MJB_CI_G1], g2 = cint[MJB_CI_G2];
  const double margin = cn[MJB_CN_MARGIN];
  const int func = cint[MJB_CI_FUNC];
  double pos1[3], pos2[3], mat1[9], mat2[9];
  load_geom_pos(c, g1, pos1); load_geom_pos(c, g2, pos2);
  if (kSimple || func == MJB_FN_PLANE_SPHERE || func == MJB_FN_PLANE_CAPSULE || func == MJB_FN_SPHERE_SPHERE ||
      func == MJB_FN_SPHERE_CAPSULE || func == MJB_FN_CAPSULE_CAPSULE) {
    // these read only the z axis of either frame (plane normal, capsule axis): one vector each
    double z1[3], z2[3];
    load_geom_z(c, g1, z1); load_geom_z(c, g2, z2);
    for (int k = 0; k < 9; k++) { mat1[k] = 0; mat2[k] = 0; }
    for (int k = 0; k < 3; k++) { mat1[3*k + 2] = z1[k]; mat2[3*k + 2] = z2[k]; }
  } else {
    ldn(mat1, gxmat, 9*g1, 9); ldn(mat2, gxmat, 9*g2, 9);
  }
  const double* size1 = geom_size + 3*g1;
  const double* size2 = geom_size + 3*g2;
  int num = 0;
  switch (func) {
    case MJB_FN_PLANE_SPHERE: num = plane_sphere(con, margin, pos1, mat1, pos2, size2[0]); break;
    case MJB_FN_PLANE_CAPSULE: num = plane_capsule(con, margin, pos1, mat1, pos2, mat2, size2); break;
    case MJB_FN_PLANE_CYLINDER: if (!kSimple) num = plane_cylinder(con, margin, pos1, mat1, pos2, mat2, size2); break;
    case MJB_FN_PLANE_BOX: if (!kSimple) num = plane_box(con, margin, pos1, mat1, pos2, mat2, size2); break;
    case MJB_FN_PLANE_ELLIPSOID: if (!kSimple) num = plane_ellipsoid(con, margin, pos1, mat1, pos2, mat2, size2); break;
    case MJB_FN_SPHERE_BOX: if (!kSimple) num = sphere_box(con, margin, pos1, size1, pos2, mat2, size2); break;
    case MJB_FN_CAPSULE_BOX: if (!kSimple) num = capsule_box(con, margin, pos1, mat1, size1, pos2, mat2, size2); break;
    case MJB_FN_BOX_BOX: if (!kSimple) num = box_box(con, margin, pos1, mat1, size1, pos2, mat2, size2); break;
    case MJB_FN_SPHERE_SPHERE:
      num = sphere_sphere(con, margin, pos1, mat1, size1[0], pos2, mat2, size2[0]); break;
    case MJB_FN_SPHERE_CAPSULE:
      num = sphere_capsule(con, margin, pos1, mat1, size1, pos2, mat2, size2); break;
    case MJB_FN_SPHERE_CYLINDER:
      if (!kSimple) num = sphere_cylinder(con, margin, pos1, mat1, size1, pos2, mat2, size2); break;
    case MJB_FN_CAPSULE_CAPSULE:
      num = capsule_capsule(con, margin, pos1, mat1, size1, pos2, mat2, size2); break;
    case MJB_FN_CONVEX:
      if (kConvex) num = convex_pair(con, margin, MI(geom_type)[g1], pos1, mat1, size1, MI(geom_type)[g2], pos2, mat2, size2,
                                     c.H->ccd_tolerance, c.H->ccd_iterations);
      break;
    default: break;
  }
  for (int k = 0; k < num; k++) makeFrame(con[k].frame);
  return num;
}

// Exact pre-test of the narrow phase: would candidate ci yield at least one contact on this state?
// For the sphere / capsule / plane primitives the accept decision of the reference is a single
// distance comparison that comes before any square root, normalisation or frame construction
// (mjraw_PlaneSphere :36, mjraw_SphereSphere :262), so the test evaluates exactly those
// expressions and nothing else. The pooled contact kernel runs it on every bounding-sphere
// survivor with all lanes busy and sends only the hits (about one in four for the humanoid) to
// narrow_pair. Other pair types report true and are decided by narrow_pair itself.
MJB_DI bool sphere_pair_hit(double margin, const double* pos1, double r1, const double* pos2, double r2) {
  const double dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  const double cdist_sqr = dot3(dif, dif);
  const double min_dist = margin + r1 + r2;
  return !(cdist_sqr > min_dist*min_dist);
}
MJB_DI bool plane_sphere_hit(double margin, const double* pos1, const double* normal, const double* pos2,
                             double radius) {
  const double tmp[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  return !(dot3(tmp, normal) > margin + radius);
}

MJB_HD inline bool narrow_test(Ctx& c, int ci) {
  const int* cint = MI(cand_int) + MJB_CAND_NI*ci;
  const int func = cint[MJB_CI_FUNC];
  if (!(func == MJB_FN_PLANE_SPHERE || func == MJB_FN_PLANE_CAPSULE || func == MJB_FN_SPHERE_SPHERE ||
        func == MJB_FN_SPHERE_CAPSULE || func == MJB_FN_CAPSULE_CAPSULE)) return true;
  const double margin = MD(cand_num)[MJB_CAND_NN*ci + MJB_CN_MARGIN];
  const double* geom_size = MD(geom_size);
  const int g1 = cint[MJB_CI_G1], g2 = cint[MJB_CI_G2];
  const double* size1 = geom_size + 3*g1;
  const double* size2 = geom_size + 3*g2;
  double pos1[3], pos2[3], z1[3], z2[3];
  load_geom_pos(c, g1, pos1); load_geom_pos(c, g2, pos2);
  load_geom_z(c, g1, z1); load_geom_z(c, g2, z2);
  switch (func) {
    case MJB_FN_PLANE_SPHERE:
      return plane_sphere_hit(margin, pos1, z1, pos2, size2[0]);
    case MJB_FN_PLANE_CAPSULE: {
      const double seg[3] = {size2[1]*z2[0], size2[1]*z2[1], size2[1]*z2[2]};
      double p[3] = {pos2[0] + seg[0], pos2[1] + seg[1], pos2[2] + seg[2]};
      const bool h1 = plane_sphere_hit(margin, pos1, z1, p, size2[0]);
      p[0] = pos2[0] - seg[0]; p[1] = pos2[1] - seg[1]; p[2] = pos2[2] - seg[2];
      return h1 || plane_sphere_hit(margin, pos1, z1, p, size2[0]);
    }
    case MJB_FN_SPHERE_SPHERE:
      return sphere_pair_hit(margin, pos1, size1[0], pos2, size2[0]);
    case MJB_FN_SPHERE_CAPSULE: {
      double vec[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
      const double x = clip(dot3(z2, vec), -size2[1], size2[1]);
      vec[0] = z2[0]*x + pos2[0]; vec[1] = z2[1]*x + pos2[1]; vec[2] = z2[2]*x + pos2[2];
      return sphere_pair_hit(margin, pos1, size1[0], vec, size2[0]);
    }
    default: {   // MJB_FN_CAPSULE_CAPSULE, the closest-point search of mjraw_CapsuleCapsule (:398)
      const double axis1[3] = {z1[0]*size1[1], z1[1]*size1[1], z1[2]*size1[1]};
      const double axis2[3] = {z2[0]*size2[1], z2[1]*size2[1], z2[2]*size2[1]};
      const double dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
      const double ma = dot3(axis1, axis1);
      const double mb = -dot3(axis1, axis2);
      const double mc = dot3(axis2, axis2);
      const double u = -dot3(axis1, dif);
      const double v = dot3(axis2, dif);
      const double det = ma*mc - mb*mb;
      double vec1[3], vec2[3];
      if (fabs(det) >= MJB_MINVAL) {
        double x1 = (mc*u - mb*v) / det;
        double x2 = (ma*v - mb*u) / det;
        if (x1 > 1) { x1 = 1; x2 = (v - mb) / mc; }
        else if (x1 < -1) { x1 = -1; x2 = (v + mb) / mc; }
        if (x2 > 1) { x2 = 1; x1 = clip((u - mb) / ma, -1, 1); }
        else if (x2 < -1) { x2 = -1; x1 = clip((u + mb) / ma, -1, 1); }
        for (int k = 0; k < 3; k++) {
          vec1[k] = axis1[k]*x1 + pos1[k];
          vec2[k] = axis2[k]*x2 + pos2[k];
        }
        return sphere_pair_hit(margin, vec1, size1[0], vec2, size2[0]);
      }
      return true;   // parallel axes (|det| < mjMINVAL): up to four sphere tests, left to narrow_pair
    }
  }
}

// narrow phase of one candidate pair followed by the rows of every contact it yields
MJB_HD inline void collide_pair(Ctx& c, int ci) {
  Con con[MJB_MAXCON_PAIR];
  const int num = narrow_pair(c, ci, con);
  for (int k = 0; k < num; k++) process_contact(c, ci, con[k]);
}

// mj_collision over the static candidate list (engine_collision_driver.c:265-484) followed by the
// contact rows of mj_makeConstraint. The candidate list already encodes the body-pair filters,
// explicit pairs, and the reference's contact ordering (see mjb_upload.cc).
//
// Divergence control, in two kernels:
//  contact_scan   : every lane tests the same candidate on its own state with the cheap
//                   bounding-sphere filter (mj_filterSphere :146-163; uniform control flow) and
//                   records the survivors as a bit mask plus their count.
//  contact_process: each lane expands its mask into a private list and all lanes process their own
//                   k-th survivor together (narrow phase + contact rows), so a lane is busy
//                   whenever it has work instead of idling while another lane's pair is expanded.
//                   Per-lane order is candidate order, so the reference's contact order is kept.
// (A counting sort of the states by survivor count in front of contact_process was measured and
//  dropped: lane occupancy did not improve -- the idle lanes come from the contact / no-contact
//  outcome of the narrow phase -- while the permuted, uncoalesced scratch accesses tripled the
//  DRAM traffic; profiles/r01_launches_sorted_contact_experiment.csv.)
MJB_HD inline int contact_scan(Ctx& c) {
  const mjbHdr& H = *c.H;
  // compact scan rows (mjb_upload.cc): geom 1 | kind << 28, geom 2, and the bound
  const int* scan_int = MI(scan_int);
  const double* scan_bound = MD(scan_bound);
  const int ncand = H.ncand;
  int last_g1 = -1, total = 0;
  double pos1[3] = {0, 0, 0}, nrm[3] = {0, 0, 0};
  // survivor words are written in order; `cw` is the word being assembled in `bits`
  unsigned bits = 0;
  int cw = 0;
  auto advance_to = [&](int ci) {        // candidates up to ci are decided: flush the words in front of ci's
    const int w = ci >> 5;
    if (w != cw) {
      c.isc[(size_t)(MJB_ISC_MASK + cw) * MJB_LS] = (int)bits;
      for (int k = cw + 1; k < w; k++) c.isc[(size_t)(MJB_ISC_MASK + k) * MJB_LS] = 0;
      bits = 0;
      cw = w;
    }
  };
  auto test = [&](int ci) {
    const int g1k = scan_int[2*ci], g2 = scan_int[2*ci + 1];
    const int g1 = g1k & 0xfffffff, planeflag = (int)((unsigned)g1k >> 28);
    const double bound = scan_bound[ci];
    if (g1 != last_g1) {          // candidates are grouped by geom 1: keep it in registers
      load_geom_pos(c, g1, pos1);
      if (planeflag == 1) load_geom_z(c, g1, nrm);
      last_g1 = g1;
    }
    double pos2[3];
    load_geom_pos(c, g2, pos2);
    bool pass = true;
    if (planeflag == 0) {
      const double dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
      pass = !(dif[0]*dif[0] + dif[1]*dif[1] + dif[2]*dif[2] > bound*bound);
    } else if (planeflag == 1) {
      const double dif[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
      pass = !(dot3(dif, nrm) > bound);
    }
    advance_to(ci);
    if (pass) { bits |= 1u << (ci & 31); total++; }
  };
  if (H.nrun == 0) {
    MJB_UNROLL
    for (int ci = 0; ci < ncand; ci++) test(ci);
  } else {
    // tree-level broadphase: bounding sphere of every kinematic tree about its origin, over the
    // geoms a candidate pair reads; a run of candidates between two trees is skipped when the
    // spheres are further apart than the largest contact margin. A warp skips a run only if all of
    // its 32 states do (uniform control flow; the others' tests fail anyway).
    const int* tree_int = MI(tree_int); const int* geom_store = MI(geom_store);
    const double* rbound = MD(geom_rbound);
    double* org = SC(origin); double* ts = SC(tree_sphere);
    const double max_margin = MD(scan_misc)[0];
    for (int t = 0; t < H.ntree; t++) {
      double O[3], r = 0;
      ldn(O, org, 3*tree_int[3*t], 3);
      for (int g = tree_int[3*t + 1]; g < tree_int[3*t + 2]; g++) {
        if (!(geom_store[g] & 3)) continue;     // bit 2 alone: kept only in runs with transmission / sensor outputs
        double p[3];
        load_geom_pos(c, g, p);
        const double d[3] = {p[0] - O[0], p[1] - O[1], p[2] - O[2]};
        r = fmax(r, sqrt(d[0]*d[0] + d[1]*d[1] + d[2]*d[2]) + rbound[g]);
      }
      const double rec[4] = {O[0], O[1], O[2], r};
      stn(ts, 4*t, rec, 4);
    }
    const int* run = MI(scan_run);
    for (int k = 0; k < H.nrun; k++) {
      const int first = run[4*k], count = run[4*k + 1], t1 = run[4*k + 2], t2 = run[4*k + 3];
      bool near = true;
      if (t1 >= 0 && t2 >= 0 && t1 != t2) {
        double a[4], b4[4];
        ldn(a, ts, 4*t1, 4); ldn(b4, ts, 4*t2, 4);
        const double d[3] = {a[0] - b4[0], a[1] - b4[1], a[2] - b4[2]};
        // slack of 1e-9 relative: the skip must never be tighter than the geom-level test it replaces
        const double reach = (a[3] + b4[3] + max_margin) * (1 + 1e-9) + 1e-12;
        near = !(d[0]*d[0] + d[1]*d[1] + d[2]*d[2] > reach*reach);
      }
      if (!MJB_WARP_ANY(near)) continue;           // every candidate of the run fails: bits stay 0
      for (int ci = first; ci < first + count; ci++) test(ci);
    }
  }
  // flush the word in progress and clear the words behind it
  if (ncand > 0) {
    c.isc[(size_t)(MJB_ISC_MASK + cw) * MJB_LS] = (int)bits;
    for (int k = cw + 1; k <= (ncand - 1) >> 5; k++) c.isc[(size_t)(MJB_ISC_MASK + k) * MJB_LS] = 0;
  }
  c.isc[(size_t)MJB_ISC_NSURV * MJB_LS] = total;
  return total;
}

MJB_HD inline void contact_process(Ctx& c, bool valid, int* list, int lstride, int cap) {
  const mjbHdr& H = *c.H;
  const int nwords = (H.ncand + 31) >> 5;
  int w = 0;
  unsigned bits = (valid && nwords > 0) ? (unsigned)c.isc[(size_t)MJB_ISC_MASK * MJB_LS] : 0u;
  while (true) {
    int cnt = 0;
    if (valid) {
      while (cnt < cap) {
        while (bits == 0 && w + 1 < nwords) {
          w++;
          bits = (unsigned)c.isc[(size_t)(MJB_ISC_MASK + w) * MJB_LS];
        }
        if (bits == 0) break;
#if defined(__CUDA_ARCH__)
        const int b = __ffs((int)bits) - 1;
#else
        const int b = __builtin_ctz(bits);
#endif
        bits &= bits - 1;
        list[cnt*lstride] = (w << 5) + b;
        cnt++;
      }
    }
    const int maxcnt = MJB_WARP_MAX(cnt);
    if (maxcnt == 0) break;
    for (int k = 0; k < maxcnt; k++) {
      if (k < cnt) collide_pair(c, list[k*lstride]);
    }
  }
}


#endif  // MJB_NARROW_H_
