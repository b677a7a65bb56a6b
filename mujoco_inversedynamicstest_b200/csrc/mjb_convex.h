// Part of the per-state mj_inverse pipeline (mjb_pipeline.h includes it inside namespace mjb, before mjb_narrow.h).
//
// Convex geom pairs: the pairs of the reference's collision table that go through mjc_Convex
// (engine_collision_driver.c, mjCOLLISIONFUNC: sphere / capsule / ellipsoid / cylinder / box against ellipsoid,
// capsule / cylinder / box against cylinder, ellipsoid against box), with the reference's native pipeline
// (mjDSBL_NATIVECCD clear): GJK on the Minkowski difference for the distance (engine_collision_gjk.c:163-275,
// Montanari's signed-volume sub-algorithm :544-818), a tetrahedron search when only contact matters (:393-450),
// EPA on a polytope grown from the final simplex for the penetration (:892-1160, :1329-1462), driven as mjc_ccd
// does (:2215-2343: spheres and capsules shrunk to points / segments first and inflated afterwards) and turned
// into one contact as mjc_CCDIteration does (engine_collision_convex.c:791-819). Support points per geom type:
// engine_collision_convex.c:146-336.
//
// Whether a pair touches must be the reference's decision bit for bit, and both algorithms are iterations whose
// path depends on every comparison, so each expression keeps the reference's operand order (the library is built
// without floating-point contraction). The organisation is this file's own: one value type per geom, vertices and
// faces in fixed arrays on the thread's stack addressed by 16-bit indices (a face list of MJB_CVX_MAXFACE like the
// reference's lower bound of 1000), the horizon walk with an explicit stack instead of recursion.
#ifndef MJB_CONVEX_H_
#define MJB_CONVEX_H_

#define MJB_CVX_MAXFACE 1000      // faces of the expanding polytope (engine_collision_gjk.c:2297)
#define MJB_CVX_MAXVERT (5 + MJB_CVX_MAXIT)
#define MJB_CVX_MAXHORIZON (6 + MJB_CVX_MAXIT)
#define MJB_CVX_MAXDEPTH 512      // frames of the horizon walk
#define MJB_CVX_BIG 3.40282346638528859811704183484516925e+38   // FLT_MAX as a double

// one geom as the support mapping sees it; `shape` is the geom type, or a point / segment for the shrunk
// sphere / capsule of the first GJK run
enum { MJB_CVX_POINT = 100, MJB_CVX_SEGMENT = 101 };
struct CvxGeom {
  int type;            // geom type (mjtGeom)
  int shape;           // type, MJB_CVX_POINT or MJB_CVX_SEGMENT
  const double* pos;   // geom_xpos
  const double* mat;   // geom_xmat (rows)
  const double* size;  // geom_size
  double margin;       // contact margin carried by the support mapping (half of it on either geom)
};
struct CvxVert { double m[3], a[3], b[3]; };     // point of the Minkowski difference, its points on geom 1 / geom 2
struct CvxFace {
  double w[3];         // projection of the origin on the face's plane
  double dist;         // |w|
  short v[3];          // vertices
  short adj[3];        // faces across the edges (v0 v1), (v1 v2), (v2 v0)
  short slot;          // position in the candidate list; -1 not listed, -2 deleted
};
struct CvxRun {        // state shared by the stages (mjCCDStatus)
  double dist, xa[3], xb[3];
  int nx, iter, nsimplex;
  CvxVert simplex[4];
  double tolerance, cutoff;
  int maxit;
};
struct CvxPoly {
  CvxVert verts[MJB_CVX_MAXVERT];
  CvxFace faces[MJB_CVX_MAXFACE];
  short cand[MJB_CVX_MAXFACE];
  int nverts, nfaces, ncand;
};

MJB_DI void cvx_sub(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
MJB_DI void cvx_scl(double* r, const double* v, double s) { r[0] = s*v[0]; r[1] = s*v[1]; r[2] = s*v[2]; }
MJB_DI void cvx_cpy(double* r, const double* v) { r[0] = v[0]; r[1] = v[1]; r[2] = v[2]; }
MJB_DI double cvx_norm(const double* v) { return sqrt(dot3(v, v)); }
MJB_DI double cvx_det(const double* a, const double* b, const double* c) {
  return a[0]*(b[1]*c[2] - b[2]*c[1]) + a[1]*(b[2]*c[0] - b[0]*c[2]) + a[2]*(b[0]*c[1] - b[1]*c[0]);
}
MJB_DI int cvx_same_sign(double a, double b) { return (a > 0 && b > 0) ? 1 : ((a < 0 && b < 0) ? -1 : 0); }

// ---- support points (engine_collision_convex.c:146-336): the farthest point of the geom along dir
MJB_NP inline void cvx_geom_support(double* res, const CvxGeom& g, const double* dir) {
  const double* mat = g.mat; const double* pos = g.pos; const double* size = g.size;
  if (g.shape == MJB_CVX_POINT) { cvx_cpy(res, pos); return; }
  if (g.shape == MJB_GEOM_SPHERE) {
    res[0] = size[0]*dir[0] + pos[0]; res[1] = size[0]*dir[1] + pos[1]; res[2] = size[0]*dir[2] + pos[2];
    return;
  }
  const double l[3] = {mat[0]*dir[0] + mat[3]*dir[1] + mat[6]*dir[2], mat[1]*dir[0] + mat[4]*dir[1] + mat[7]*dir[2],
                       mat[2]*dir[0] + mat[5]*dir[1] + mat[8]*dir[2]};
  double t[3];
  if (g.shape == MJB_CVX_SEGMENT) {
    t[0] = 0; t[1] = 0; t[2] = (l[2] >= 0 ? size[1] : -size[1]);
  } else if (g.shape == MJB_GEOM_CAPSULE) {
    t[0] = l[0] * size[0]; t[1] = l[1] * size[0]; t[2] = l[2] * size[0];
    t[2] += (l[2] >= 0 ? size[1] : -size[1]);
  } else if (g.shape == MJB_GEOM_ELLIPSOID) {
    t[0] = l[0] * size[0]; t[1] = l[1] * size[1]; t[2] = l[2] * size[2];
    const double n = sqrt(t[0]*t[0] + t[1]*t[1] + t[2]*t[2]);
    if (n < MJB_MINVAL) {
      t[0] = size[0]; t[1] = 0; t[2] = 0;
    } else {
      const double ninv = 1/n;
      t[0] *= ninv * size[0]; t[1] *= ninv * size[1]; t[2] *= ninv * size[2];
    }
  } else if (g.shape == MJB_GEOM_CYLINDER) {
    double n = l[0]*l[0] + l[1]*l[1];
    if (n > MJB_MINVAL*MJB_MINVAL) {
      n = size[0] / sqrt(n);
      t[0] = l[0] * n; t[1] = l[1] * n;
    } else {
      t[0] = t[1] = 0;
    }
    t[2] = (l[2] < 0 ? -1.0 : (l[2] > 0 ? 1.0 : 0.0)) * size[1];
  } else {                                  // box
    t[0] = (l[0] >= 0 ? 1 : -1) * size[0]; t[1] = (l[1] >= 0 ? 1 : -1) * size[1]; t[2] = (l[2] >= 0 ? 1 : -1) * size[2];
  }
  res[0] = mat[0]*t[0] + mat[1]*t[1] + mat[2]*t[2];
  res[1] = mat[3]*t[0] + mat[4]*t[1] + mat[5]*t[2];
  res[2] = mat[6]*t[0] + mat[7]*t[1] + mat[8]*t[2];
  res[0] += pos[0]; res[1] += pos[1]; res[2] += pos[2];
}

// support point of the Minkowski difference along dir (dneg = -dir), margins included (:277-298)
MJB_DI void cvx_support(CvxVert& v, const CvxGeom& A, const CvxGeom& B, const double* dir, const double* dneg) {
  cvx_geom_support(v.a, A, dir);
  if (A.margin > 0) {
    const double h = 0.5 * A.margin;
    v.a[0] += dir[0] * h; v.a[1] += dir[1] * h; v.a[2] += dir[2] * h;
  }
  cvx_geom_support(v.b, B, dneg);
  if (B.margin > 0) {
    const double h = 0.5 * B.margin;
    v.b[0] += dneg[0] * h; v.b[1] += dneg[1] * h; v.b[2] += dneg[2] * h;
  }
  cvx_sub(v.m, v.a, v.b);
}

// ---- closest point of a simplex to the origin in barycentric coordinates (:453-818)
MJB_DI void cvx_comb(double* r, const double* c, int n, const double* p0, const double* p1, const double* p2,
                     const double* p3) {
  if (n == 1) {
    r[0] = c[0]*p0[0]; r[1] = c[0]*p0[1]; r[2] = c[0]*p0[2];
  } else if (n == 2) {
    r[0] = c[0]*p0[0] + c[1]*p1[0]; r[1] = c[0]*p0[1] + c[1]*p1[1]; r[2] = c[0]*p0[2] + c[1]*p1[2];
  } else if (n == 3) {
    r[0] = c[0]*p0[0] + c[1]*p1[0] + c[2]*p2[0]; r[1] = c[0]*p0[1] + c[1]*p1[1] + c[2]*p2[1];
    r[2] = c[0]*p0[2] + c[1]*p1[2] + c[2]*p2[2];
  } else if (n == 4) {
    r[0] = c[0]*p0[0] + c[1]*p1[0] + c[2]*p2[0] + c[3]*p3[0]; r[1] = c[0]*p0[1] + c[1]*p1[1] + c[2]*p2[1] + c[3]*p3[1];
    r[2] = c[0]*p0[2] + c[1]*p1[2] + c[2]*p2[2] + c[3]*p3[2];
  }
}

// origin projected on the plane of three points; 1 if the points are collinear (:482-517)
MJB_NP inline int cvx_project_plane(double* res, const double* p0, const double* p1, const double* p2) {
  double e10[3], e20[3], e21[3], n[3], nv, nn;
  cvx_sub(e10, p1, p0); cvx_sub(e20, p2, p0); cvx_sub(e21, p2, p1);
  cross3(n, e21, e10);
  nv = dot3(n, p1); nn = dot3(n, n);
  if (nn == 0) return 1;
  if (nv != 0 && nn > MJB_MINVAL) { cvx_scl(res, n, nv / nn); return 0; }
  cross3(n, e10, e20);
  nv = dot3(n, p0); nn = dot3(n, n);
  if (nn == 0) return 1;
  if (nv != 0 && nn > MJB_MINVAL) { cvx_scl(res, n, nv / nn); return 0; }
  cross3(n, e20, e21);
  nv = dot3(n, p2); nn = dot3(n, n);
  cvx_scl(res, n, nv / nn);
  return 0;
}

MJB_NP inline void cvx_bary1(double* lam, const double* p0, const double* p1) {
  double d[3], o[3];
  cvx_sub(d, p1, p0);
  const double s = -(dot3(p1, d) / dot3(d, d));
  o[0] = p1[0] + s*d[0]; o[1] = p1[1] + s*d[1]; o[2] = p1[2] + s*d[2];
  double mumax = 0;
  int ax = 0;
  for (int i = 0; i < 3; i++) {
    const double mu = p0[i] - p1[i];
    if (fabs(mu) >= fabs(mumax)) { mumax = mu; ax = i; }
  }
  const double c0 = o[ax] - p1[ax], c1 = p0[ax] - o[ax];
  if (cvx_same_sign(mumax, c0) && cvx_same_sign(mumax, c1)) {
    lam[0] = c0 / mumax; lam[1] = c1 / mumax;
  } else {
    lam[0] = 0; lam[1] = 1;
  }
}

// minors of the 4 x 4 barycentric system with the last row of ones: the triangle's projected areas
MJB_DI void cvx_minors(double* M, const double* p0, const double* p1, const double* p2) {
  M[0] = p1[1]*p2[2] - p1[2]*p2[1] - p0[1]*p2[2] + p0[2]*p2[1] + p0[1]*p1[2] - p0[2]*p1[1];
  M[1] = p1[0]*p2[2] - p1[2]*p2[0] - p0[0]*p2[2] + p0[2]*p2[0] + p0[0]*p1[2] - p0[2]*p1[0];
  M[2] = p1[0]*p2[1] - p1[1]*p2[0] - p0[0]*p2[1] + p0[1]*p2[0] + p0[0]*p1[1] - p0[1]*p1[0];
}
// the two axes kept after dropping the one with the largest projected area, and that area
MJB_DI double cvx_drop_axis(const double* M, int* x, int* y) {
  const double m0 = fabs(M[0]), m1 = fabs(M[1]), m2 = fabs(M[2]);
  if (m0 >= m1 && m0 >= m2) { *x = 1; *y = 2; return M[0]; }
  if (m1 >= m2) { *x = 0; *y = 2; return M[1]; }
  *x = 0; *y = 1; return M[2];
}
// signed areas of the triangles (q, p1, p2), (q, p0, p2), (q, p0, p1) in the kept plane
MJB_DI void cvx_areas(double* C, const double* q, const double* p0, const double* p1, const double* p2, int x, int y) {
  C[0] = q[x]*p1[y] + q[y]*p2[x] + p1[x]*p2[y] - q[x]*p2[y] - q[y]*p1[x] - p2[x]*p1[y];
  C[1] = q[x]*p2[y] + q[y]*p0[x] + p2[x]*p0[y] - q[x]*p0[y] - q[y]*p2[x] - p0[x]*p2[y];
  C[2] = q[x]*p0[y] + q[y]*p1[x] + p0[x]*p1[y] - q[x]*p1[y] - q[y]*p0[x] - p1[x]*p0[y];
}

MJB_NP inline void cvx_bary2(double* lam, const double* p0, const double* p1, const double* p2) {
  double o[3];
  if (cvx_project_plane(o, p0, p1, p2)) {
    cvx_bary1(lam, p0, p1);
    lam[2] = 0;
    return;
  }
  double M[3], C[3];
  int x, y;
  cvx_minors(M, p0, p1, p2);
  const double Mmax = cvx_drop_axis(M, &x, &y);
  cvx_areas(C, o, p0, p1, p2, x, y);
  const int s0 = cvx_same_sign(Mmax, C[0]), s1 = cvx_same_sign(Mmax, C[1]), s2 = cvx_same_sign(Mmax, C[2]);
  if (s0 && s1 && s2) {
    lam[0] = C[0] / Mmax; lam[1] = C[1] / Mmax; lam[2] = C[2] / Mmax;
    return;
  }
  double dmin = MJB_MAXVAL;
  if (!s0) {
    double l2[2], q[3];
    cvx_bary1(l2, p1, p2);
    cvx_comb(q, l2, 2, p1, p2, nullptr, nullptr);
    lam[0] = 0; lam[1] = l2[0]; lam[2] = l2[1];
    dmin = dot3(q, q);
  }
  if (!s1) {
    double l2[2], q[3];
    cvx_bary1(l2, p0, p2);
    cvx_comb(q, l2, 2, p0, p2, nullptr, nullptr);
    const double d = dot3(q, q);
    if (d < dmin) { lam[0] = l2[0]; lam[1] = 0; lam[2] = l2[1]; dmin = d; }
  }
  if (!s2) {
    double l2[2], q[3];
    cvx_bary1(l2, p0, p1);
    cvx_comb(q, l2, 2, p0, p1, nullptr, nullptr);
    const double d = dot3(q, q);
    if (d < dmin) { lam[0] = l2[0]; lam[1] = l2[1]; lam[2] = 0; }
  }
}

MJB_NP inline void cvx_bary3(double* lam, const double* p0, const double* p1, const double* p2, const double* p3) {
  const double c0 = -cvx_det(p1, p2, p3), c1 = cvx_det(p0, p2, p3), c2 = -cvx_det(p0, p1, p3), c3 = cvx_det(p0, p1, p2);
  const double det = c0 + c1 + c2 + c3;
  const int s0 = cvx_same_sign(det, c0), s1 = cvx_same_sign(det, c1), s2 = cvx_same_sign(det, c2),
            s3 = cvx_same_sign(det, c3);
  if (s0 && s1 && s2 && s3) {
    lam[0] = c0 / det; lam[1] = c1 / det; lam[2] = c2 / det; lam[3] = c3 / det;
    return;
  }
  double dmin = MJB_MAXVAL;
  if (!s0) {
    double l3[3], q[3];
    cvx_bary2(l3, p1, p2, p3);
    cvx_comb(q, l3, 3, p1, p2, p3, nullptr);
    lam[0] = 0; lam[1] = l3[0]; lam[2] = l3[1]; lam[3] = l3[2];
    dmin = dot3(q, q);
  }
  if (!s1) {
    double l3[3], q[3];
    cvx_bary2(l3, p0, p2, p3);
    cvx_comb(q, l3, 3, p0, p2, p3, nullptr);
    const double d = dot3(q, q);
    if (d < dmin) { lam[0] = l3[0]; lam[1] = 0; lam[2] = l3[1]; lam[3] = l3[2]; dmin = d; }
  }
  if (!s2) {
    double l3[3], q[3];
    cvx_bary2(l3, p0, p1, p3);
    cvx_comb(q, l3, 3, p0, p1, p3, nullptr);
    const double d = dot3(q, q);
    if (d < dmin) { lam[0] = l3[0]; lam[1] = l3[1]; lam[2] = 0; lam[3] = l3[2]; dmin = d; }
  }
  if (!s3) {
    double l3[3], q[3];
    cvx_bary2(l3, p0, p1, p2);
    cvx_comb(q, l3, 3, p0, p1, p2, nullptr);
    const double d = dot3(q, q);
    if (d < dmin) { lam[0] = l3[0]; lam[1] = l3[1]; lam[2] = l3[2]; lam[3] = 0; }
  }
}

// ---- GJK (:163-275) -------------------------------------------------------------------------------
MJB_DI bool cvx_close(const double* a, const double* b) {
  return fabs(a[0] - b[0]) < MJB_MINVAL && fabs(a[1] - b[1]) < MJB_MINVAL && fabs(a[2] - b[2]) < MJB_MINVAL;
}

// signed distance of the origin to the plane of (p0, p1, p2) with its unit normal; MJB_MAXVAL if degenerate (:375-389)
MJB_DI double cvx_plane_distance(double* nrm, const CvxVert& p0, const CvxVert& p1, const CvxVert& p2) {
  double d1[3], d2[3];
  cvx_sub(d1, p2.m, p0.m); cvx_sub(d2, p1.m, p0.m);
  cross3(nrm, d1, d2);
  double n = dot3(nrm, nrm);
  if (n > MJB_MINVAL*MJB_MINVAL && n < MJB_MAXVAL*MJB_MAXVAL) {
    n = 1/sqrt(n);
    cvx_scl(nrm, nrm, n);
    return dot3(nrm, p0.m);
  }
  return MJB_MAXVAL;
}

// does the tetrahedron of the run's simplex, pushed outwards face by face, reach the origin? 1 yes (simplex
// updated for EPA), 0 no, -1 undecided (:393-450)
MJB_NP inline int cvx_tetra_search(CvxRun& r, const CvxGeom& A, const CvxGeom& B) {
  CvxVert sx[4] = {r.simplex[0], r.simplex[1], r.simplex[2], r.simplex[3]};
  int s[4] = {0, 1, 2, 3};
  int k = r.iter;
  for (; k < r.maxit; k++) {
    double dist[4], nrm[12];
    dist[0] = cvx_plane_distance(nrm + 0, sx[s[2]], sx[s[1]], sx[s[3]]);
    dist[1] = cvx_plane_distance(nrm + 3, sx[s[0]], sx[s[2]], sx[s[3]]);
    dist[2] = cvx_plane_distance(nrm + 6, sx[s[1]], sx[s[0]], sx[s[3]]);
    dist[3] = cvx_plane_distance(nrm + 9, sx[s[0]], sx[s[1]], sx[s[2]]);
    if (!dist[3] || !dist[2] || !dist[1] || !dist[0]) { r.iter = k; return -1; }
    int i = (dist[0] < dist[1]) ? 0 : 1;
    int j = (dist[2] < dist[3]) ? 2 : 3;
    const int worst = (dist[i] < dist[j]) ? i : j;
    if (dist[worst] > 0) {
      r.nsimplex = 4;
      r.simplex[0] = sx[s[0]]; r.simplex[1] = sx[s[1]]; r.simplex[2] = sx[s[2]]; r.simplex[3] = sx[s[3]];
      r.iter = k;
      return 1;
    }
    const double* dir = nrm + 3*worst;
    const double dneg[3] = {-dir[0], -dir[1], -dir[2]};
    cvx_support(sx[s[worst]], A, B, dir, dneg);
    if (dot3(dir, sx[s[worst]].m) < 0) { r.nsimplex = 0; r.iter = k; return 0; }
    i = (worst + 1) & 3; j = (worst + 2) & 3;
    const int t = s[i]; s[i] = s[j]; s[j] = t;
  }
  r.iter = k;
  return -1;
}

MJB_NP inline void cvx_gjk(CvxRun& r, const CvxGeom& A, const CvxGeom& B) {
  const bool want_dist = r.cutoff > 0;
  bool tetra_first = !want_dist;
  CvxVert* sx = r.simplex;
  int n = 0, k = 0;
  double x[3], lam[4] = {1, 0, 0, 0};
  const double cutoff2 = r.cutoff * r.cutoff;
  // two boxes without margin end after finitely many steps: tolerance 0 (:150-159)
  const bool discrete = A.margin == 0 && B.margin == 0 && A.type == MJB_GEOM_BOX && B.type == MJB_GEOM_BOX;
  const double eps = discrete ? 0 : r.tolerance * r.tolerance;
  cvx_sub(x, r.xa, r.xb);
  for (; k < r.maxit; k++) {
    {   // support point along -x, the direction normalised (:301-325)
      double dir[3] = {-1, 0, 0}, dneg[3] = {1, 0, 0};
      double nn = dot3(x, x);
      if (nn > MJB_MINVAL*MJB_MINVAL) {
        nn = 1/sqrt(nn);
        cvx_scl(dneg, x, nn);
        cvx_scl(dir, dneg, -1);
      }
      cvx_support(sx[n], A, B, dir, dneg);
    }
    const double* sk = sx[n].m;
    double diff[3];
    cvx_sub(diff, x, sk);
    if (2*dot3(x, diff) < eps) {
      if (!k) n = 1;
      break;
    }
    if (!want_dist) {
      if (dot3(x, sk) > 0) { r.iter = k; r.nsimplex = 0; r.nx = 0; r.dist = MJB_MAXVAL; return; }
    } else if (r.cutoff < MJB_MAXVAL) {
      const double vs = dot3(x, sk), vv = dot3(x, x);
      if (dot3(x, sk) > 0 && (vs*vs / vv) >= cutoff2) { r.iter = k; r.nsimplex = 0; r.nx = 0; r.dist = MJB_MAXVAL; return; }
    }
    if (n == 3 && tetra_first) {
      r.iter = k;
      const int hit = cvx_tetra_search(r, A, B);
      if (hit != -1) { r.nx = 0; r.dist = hit > 0 ? 0 : MJB_MAXVAL; return; }
      k = r.iter;
      tetra_first = false;
    }
    lam[0] = lam[1] = lam[2] = lam[3] = 0;
    if (n + 1 == 4) cvx_bary3(lam, sx[0].m, sx[1].m, sx[2].m, sx[3].m);
    else if (n + 1 == 3) cvx_bary2(lam, sx[0].m, sx[1].m, sx[2].m);
    else if (n + 1 == 2) cvx_bary1(lam, sx[0].m, sx[1].m);
    else lam[0] = 1;
    n = 0;
    for (int i = 0; i < 4; i++) {
      if (lam[i] == 0) continue;
      sx[n] = sx[i];
      lam[n++] = lam[i];
    }
    double xn[3];
    cvx_comb(xn, lam, n, sx[0].m, sx[1].m, sx[2].m, sx[3].m);
    if (cvx_close(xn, x)) break;
    cvx_cpy(x, xn);
    if (n == 4) break;
  }
  cvx_comb(r.xa, lam, n, sx[0].a, sx[1].a, sx[2].a, sx[3].a);
  cvx_comb(r.xb, lam, n, sx[0].b, sx[1].b, sx[2].b, sx[3].b);
  r.nx = 1; r.iter = k; r.nsimplex = n;
  r.dist = cvx_norm(x);
}

// ---- EPA: the polytope (:820-1226) ------------------------------------------------------------------
MJB_DI int cvx_add_vertex(CvxPoly& P, const CvxVert& v) {
  const int n = P.nverts++;
  CvxVert& q = P.verts[n];
  cvx_cpy(q.a, v.a); cvx_cpy(q.b, v.b);
  cvx_sub(q.m, v.a, v.b);
  return n;
}
// new vertex from the support mapping along d of length dn (:328-354)
MJB_DI int cvx_add_support(CvxPoly& P, const CvxGeom& A, const CvxGeom& B, const double* d, double dn) {
  double dir[3] = {1, 0, 0}, dneg[3] = {-1, 0, 0};
  if (dn > MJB_MINVAL) {
    dir[0] = d[0] / dn; dir[1] = d[1] / dn; dir[2] = d[2] / dn;
    cvx_scl(dneg, dir, -1);
  }
  const int n = P.nverts++;
  cvx_support(P.verts[n], A, B, dir, dneg);
  return n;
}
// face (v0, v1, v2) with its neighbours; returns the distance of its plane from the origin, 0 if degenerate (:1192-1215)
MJB_DI double cvx_add_face(CvxPoly& P, int v0, int v1, int v2, int a0, int a1, int a2) {
  CvxFace& f = P.faces[P.nfaces++];
  f.v[0] = (short)v0; f.v[1] = (short)v1; f.v[2] = (short)v2;
  f.adj[0] = (short)a0; f.adj[1] = (short)a1; f.adj[2] = (short)a2;
  if (cvx_project_plane(f.w, P.verts[v2].m, P.verts[v1].m, P.verts[v0].m)) return 0;
  f.dist = cvx_norm(f.w);
  f.slot = -1;
  return f.dist;
}
MJB_DI void cvx_drop_face(CvxPoly& P, int fi) {
  CvxFace& f = P.faces[fi];
  if (f.slot >= 0) {
    P.cand[f.slot] = P.cand[--P.ncand];
    P.faces[P.cand[f.slot]].slot = f.slot;
  }
  f.slot = -2;
}
MJB_DI void cvx_list_all(CvxPoly& P, int n) {
  for (int i = 0; i < n; i++) { P.cand[i] = (short)i; P.faces[i].slot = (short)i; }
  P.ncand = n;
}
// the run's simplex becomes the triangle (v0, v1, v2) of the polytope, which is emptied (:820-838)
MJB_DI void cvx_restart_from_face(CvxPoly& P, CvxRun& r, int v0, int v1, int v2) {
  r.nsimplex = 3;
  const int vi[3] = {v0, v1, v2};
  for (int k = 0; k < 3; k++) {
    cvx_cpy(r.simplex[k].a, P.verts[vi[k]].a); cvx_cpy(r.simplex[k].b, P.verts[vi[k]].b);
    cvx_cpy(r.simplex[k].m, P.verts[vi[k]].m);
  }
  P.nfaces = 0; P.nverts = 0;
}
MJB_DI bool cvx_same_side(const double* p0, const double* p1, const double* p2, const double* p3) {
  double d1[3], d2[3], d3[3], d4[3], n[3];
  cvx_sub(d1, p1, p0); cvx_sub(d2, p2, p0);
  cross3(n, d1, d2);
  cvx_sub(d3, p3, p0);
  const double s1 = dot3(n, d3);
  cvx_scl(d4, p0, -1);
  const double s2 = dot3(n, d4);
  return (s1 > 0 && s2 > 0) || (s1 < 0 && s2 < 0);
}
MJB_DI bool cvx_tetra_has_origin(const double* p0, const double* p1, const double* p2, const double* p3) {
  return cvx_same_side(p0, p1, p2, p3) && cvx_same_side(p1, p2, p3, p0) && cvx_same_side(p2, p3, p0, p1) &&
         cvx_same_side(p3, p0, p1, p2);
}
// is p on the triangle (p0, p1, p2)? (:976-1036)
MJB_DI void cvx_affine(double* lam, const double* p0, const double* p1, const double* p2, const double* p) {
  double M[3], C[3];
  int x, y;
  cvx_minors(M, p0, p1, p2);
  const double Mmax = cvx_drop_axis(M, &x, &y);
  cvx_areas(C, p, p0, p1, p2, x, y);
  lam[0] = C[0] / Mmax; lam[1] = C[1] / Mmax; lam[2] = C[2] / Mmax;
}
MJB_DI bool cvx_on_triangle(const double* p0, const double* p1, const double* p2, const double* p) {
  double lam[3];
  cvx_affine(lam, p0, p1, p2, p);
  if (lam[0] < 0 || lam[1] < 0 || lam[2] < 0) return false;
  double q[3], d[3];
  q[0] = p0[0]*lam[0] + p1[0]*lam[1] + p2[0]*lam[2];
  q[1] = p0[1]*lam[0] + p1[1]*lam[1] + p2[1]*lam[2];
  q[2] = p0[2]*lam[0] + p1[2]*lam[1] + p2[2]*lam[2];
  cvx_sub(d, q, p);
  return cvx_norm(d) < MJB_MINVAL;
}

// initial polytopes; 0 on success (:892-1157)
MJB_NP inline int cvx_start_triangle(CvxPoly& P, CvxRun& r, const CvxGeom& A, const CvxGeom& B) {
  const double* p0 = r.simplex[0].m; const double* p1 = r.simplex[1].m; const double* p2 = r.simplex[2].m;
  double d1[3], d2[3], n[3], nneg[3];
  cvx_sub(d1, p1, p0); cvx_sub(d2, p2, p0);
  cross3(n, d1, d2);
  const double nn = cvx_norm(n);
  if (nn < MJB_MINVAL) return 1;
  cvx_scl(nneg, n, -1);
  const int i0 = cvx_add_vertex(P, r.simplex[0]), i1 = cvx_add_vertex(P, r.simplex[1]), i2 = cvx_add_vertex(P, r.simplex[2]);
  const int i4 = cvx_add_support(P, A, B, nneg, nn);
  const int i3 = cvx_add_support(P, A, B, n, nn);
  const double* p3 = P.verts[i3].m; const double* p4 = P.verts[i4].m;
  if (cvx_on_triangle(p0, p1, p2, p3)) return 2;
  if (cvx_on_triangle(p0, p1, p2, p4)) return 3;
  if (r.dist > 10*MJB_MINVAL && !cvx_tetra_has_origin(p0, p1, p2, p3) && !cvx_tetra_has_origin(p0, p1, p2, p4)) return 4;
  if (cvx_add_face(P, i3, i0, i1, 1, 3, 2) < MJB_MINVAL) return 5;
  if (cvx_add_face(P, i3, i2, i0, 2, 4, 0) < MJB_MINVAL) return 5;
  if (cvx_add_face(P, i3, i1, i2, 0, 5, 1) < MJB_MINVAL) return 5;
  if (cvx_add_face(P, i4, i1, i0, 5, 0, 4) < MJB_MINVAL) return 5;
  if (cvx_add_face(P, i4, i0, i2, 3, 1, 5) < MJB_MINVAL) return 5;
  if (cvx_add_face(P, i4, i2, i1, 4, 2, 3) < MJB_MINVAL) return 5;
  cvx_list_all(P, 6);
  return 0;
}

MJB_NP inline int cvx_start_segment(CvxPoly& P, CvxRun& r, const CvxGeom& A, const CvxGeom& B) {
  const double* p0 = r.simplex[0].m; const double* p1 = r.simplex[1].m;
  double d[3];
  cvx_sub(d, p1, p0);
  double smallest = MJB_MAXVAL;
  int ax = 0;
  for (int i = 0; i < 3; i++) {
    if (fabs(d[i]) < smallest) { smallest = fabs(d[i]); ax = i; }
  }
  double e[3] = {0, 0, 0};
  e[ax] = 1;
  double t1[3], t2[3], t3[3], R[9];
  cross3(t1, e, d);
  {   // rotation by 120 degrees about d (:873-888)
    const double len = cvx_norm(d);
    const double u1 = d[0] / len, u2 = d[1] / len, u3 = d[2] / len;
    const double sn = 0.86602540378, cs = -0.5;
    R[0] = cs + u1*u1*(1 - cs);    R[1] = u1*u2*(1 - cs) - u3*sn; R[2] = u1*u3*(1 - cs) + u2*sn;
    R[3] = u2*u1*(1 - cs) + u3*sn; R[4] = cs + u2*u2*(1 - cs);    R[5] = u2*u3*(1 - cs) - u1*sn;
    R[6] = u1*u3*(1 - cs) - u2*sn; R[7] = u2*u3*(1 - cs) + u1*sn; R[8] = cs + u3*u3*(1 - cs);
  }
  mulMatVec3(t2, R, t1);
  mulMatVec3(t3, R, t2);
  const int i0 = cvx_add_vertex(P, r.simplex[0]), i1 = cvx_add_vertex(P, r.simplex[1]);
  const int i2 = cvx_add_support(P, A, B, t1, cvx_norm(t1));
  const int i3 = cvx_add_support(P, A, B, t2, cvx_norm(t2));
  const int i4 = cvx_add_support(P, A, B, t3, cvx_norm(t3));
  const double* p2 = P.verts[i2].m; const double* p3 = P.verts[i3].m; const double* p4 = P.verts[i4].m;
  if (cvx_add_face(P, i0, i2, i3, 1, 3, 2) < MJB_MINVAL) { cvx_restart_from_face(P, r, i0, i2, i3); return cvx_start_triangle(P, r, A, B); }
  if (cvx_add_face(P, i0, i4, i2, 2, 4, 0) < MJB_MINVAL) { cvx_restart_from_face(P, r, i0, i4, i2); return cvx_start_triangle(P, r, A, B); }
  if (cvx_add_face(P, i0, i3, i4, 0, 5, 1) < MJB_MINVAL) { cvx_restart_from_face(P, r, i0, i3, i4); return cvx_start_triangle(P, r, A, B); }
  if (cvx_add_face(P, i1, i3, i2, 5, 0, 4) < MJB_MINVAL) { cvx_restart_from_face(P, r, i1, i3, i2); return cvx_start_triangle(P, r, A, B); }
  if (cvx_add_face(P, i1, i2, i4, 3, 1, 5) < MJB_MINVAL) { cvx_restart_from_face(P, r, i1, i2, i4); return cvx_start_triangle(P, r, A, B); }
  if (cvx_add_face(P, i1, i4, i3, 4, 2, 3) < MJB_MINVAL) { cvx_restart_from_face(P, r, i1, i4, i3); return cvx_start_triangle(P, r, A, B); }
  // p0 / p1 point into the run's simplex, which a restart rewrites; here it is untouched
  if (r.dist > 10*MJB_MINVAL && !cvx_tetra_has_origin(p0, p2, p3, p4) && !cvx_tetra_has_origin(p1, p2, p3, p4)) return 6;
  cvx_list_all(P, 6);
  return 0;
}

MJB_NP inline int cvx_start_tetrahedron(CvxPoly& P, CvxRun& r, const CvxGeom& A, const CvxGeom& B) {
  const int i0 = cvx_add_vertex(P, r.simplex[0]), i1 = cvx_add_vertex(P, r.simplex[1]);
  const int i2 = cvx_add_vertex(P, r.simplex[2]), i3 = cvx_add_vertex(P, r.simplex[3]);
  if (cvx_add_face(P, i0, i1, i2, 1, 3, 2) < MJB_MINVAL) { cvx_restart_from_face(P, r, i0, i1, i2); return cvx_start_triangle(P, r, A, B); }
  if (cvx_add_face(P, i0, i3, i1, 2, 3, 0) < MJB_MINVAL) { cvx_restart_from_face(P, r, i0, i3, i1); return cvx_start_triangle(P, r, A, B); }
  if (cvx_add_face(P, i0, i2, i3, 0, 3, 1) < MJB_MINVAL) { cvx_restart_from_face(P, r, i0, i2, i3); return cvx_start_triangle(P, r, A, B); }
  if (cvx_add_face(P, i3, i2, i1, 2, 0, 1) < MJB_MINVAL) { cvx_restart_from_face(P, r, i3, i2, i1); return cvx_start_triangle(P, r, A, B); }
  if (!cvx_tetra_has_origin(P.verts[i0].m, P.verts[i1].m, P.verts[i2].m, P.verts[i3].m)) return 7;
  cvx_list_all(P, 4);
  return 0;
}

// ---- EPA: expansion (:1229-1462) --------------------------------------------------------------------
MJB_DI int cvx_edge_of(const CvxFace& f, int vertex) { return f.v[0] == vertex ? 0 : (f.v[1] == vertex ? 1 : 2); }

// Faces visible from w are dropped and the edges between dropped and kept faces -- the horizon -- collected as
// (kept face, its edge), in the order of the reference's depth-first recursion (horizonRec :1246-1268, horizon
// :1272-1296), which an explicit stack reproduces: a frame is a dropped face entered through edge e that still
// has to look across its edges e+1 and e+2. Returns the number of edges, -1 if a table would overflow.
struct CvxFrame { short face; signed char e, k; short nb; signed char nbe, waiting; };
MJB_NP inline int cvx_horizon(CvxPoly& P, int start, const double* w, short* hface, signed char* hedge) {
  CvxFrame st[MJB_CVX_MAXDEPTH];
  int nh = 0;
  auto visible = [&](int fi) { const CvxFace& f = P.faces[fi]; return dot3(f.w, w) >= f.dist * f.dist; };
  // walk from face fi entered through its edge e; true if fi was visible (and is gone)
  auto walk = [&](int fi, int e) -> int {
    if (!visible(fi)) return 0;
    cvx_drop_face(P, fi);
    int sp = 0, ret = 1;
    st[sp++] = CvxFrame{(short)fi, (signed char)e, 1, 0, 0, 0};
    while (sp) {
      CvxFrame& t = st[sp - 1];
      if (t.waiting) {
        if (!ret) { if (nh >= MJB_CVX_MAXHORIZON) return -1; hface[nh] = t.nb; hedge[nh++] = t.nbe; }
        t.waiting = 0; t.k++;
      }
      if (t.k >= 3) { sp--; ret = 1; continue; }
      const CvxFace& f = P.faces[t.face];
      const int i = (t.e + t.k) % 3;
      const int nb = f.adj[i];
      if (P.faces[nb].slot > -2) {
        const int nbe = cvx_edge_of(P.faces[nb], f.v[(i + 1) % 3]);
        t.waiting = 1; t.nb = (short)nb; t.nbe = (signed char)nbe;
        if (!visible(nb)) { ret = 0; continue; }
        cvx_drop_face(P, nb);
        if (sp >= MJB_CVX_MAXDEPTH) return -1;
        st[sp++] = CvxFrame{(short)nb, (signed char)nbe, 1, 0, 0, 0};
        continue;
      }
      t.k++;
    }
    return 1;
  };
  const CvxFace& f0 = P.faces[start];
  cvx_drop_face(P, start);
  for (int k = 0; k < 3; k++) {
    const int nb = f0.adj[k];
    const int nbe = cvx_edge_of(P.faces[nb], f0.v[(k + 1) % 3]);
    if (k > 0 && !(P.faces[nb].slot > -2)) continue;
    const int vis = walk(nb, nbe);
    if (vis < 0) return -1;
    if (!vis) { if (nh >= MJB_CVX_MAXHORIZON) return -1; hface[nh] = (short)nb; hedge[nh++] = (signed char)nbe; }
  }
  return nh;
}

// grows the polytope until the closest face is within tolerance of the surface; returns that face or -1 (:1329-1462)
MJB_NP inline int cvx_epa(CvxPoly& P, CvxRun& r, const CvxGeom& A, const CvxGeom& B) {
  const double tol = r.tolerance;
  double lower, upper = MJB_CVX_BIG;
  int face = -1, prev = -1, k;
  short hface[MJB_CVX_MAXHORIZON];
  signed char hedge[MJB_CVX_MAXHORIZON];
  for (k = 0; k < r.maxit; k++) {
    prev = face;
    lower = MJB_CVX_BIG;
    for (int i = 0; i < P.ncand; i++) {
      if (P.faces[P.cand[i]].dist < lower) { face = P.cand[i]; lower = P.faces[face].dist; }
    }
    if (lower > upper || face < 0) { face = prev; break; }
    if (lower <= 0) break;
    const int wi = cvx_add_support(P, A, B, P.faces[face].w, lower);
    const double* w = P.verts[wi].m;
    const double upper_k = dot3(P.faces[face].w, w) / lower;
    if (upper_k < upper) upper = upper_k;
    if (upper - lower < tol) break;
    const int ne = cvx_horizon(P, face, w, hface, hedge);
    if (ne < 3) { face = -1; break; }
    const int nf = P.nfaces;
    if (ne > MJB_CVX_MAXFACE - P.nfaces) break;
    bool failed = false;
    for (int i = 0; i < ne; i++) {
      const int cur = nf + i;
      const int before = i == 0 ? nf + ne - 1 : cur - 1;
      const int after = i == 0 ? nf + 1 : nf + (i + 1) % ne;
      CvxFace& hf = P.faces[hface[i]];
      const int e = hedge[i];
      const int v1 = hf.v[e], v2 = hf.v[(e + 1) % 3];
      hf.adj[e] = (short)cur;
      const double dist = cvx_add_face(P, wi, v2, v1, before, hface[i], after);
      if (dist == 0) { failed = true; break; }
      if (dist >= lower && dist <= upper) {
        const int s = P.ncand++;
        P.cand[s] = (short)(P.nfaces - 1);
        P.faces[P.nfaces - 1].slot = (short)s;
      }
    }
    if (failed) { face = -1; break; }
    if (!P.ncand || face < 0) break;
  }
  if (face >= 0) {
    // witness points: the face's affine coordinates of the origin's projection, applied on either geom (:1300-1325)
    const CvxFace& f = P.faces[face];
    const CvxVert& q0 = P.verts[f.v[0]]; const CvxVert& q1 = P.verts[f.v[1]]; const CvxVert& q2 = P.verts[f.v[2]];
    double lam[3];
    cvx_affine(lam, q0.m, q1.m, q2.m, f.w);
    for (int c = 0; c < 3; c++) {
      r.xa[c] = q0.a[c]*lam[0] + q1.a[c]*lam[1] + q2.a[c]*lam[2];
      r.xb[c] = q0.b[c]*lam[0] + q1.b[c]*lam[1] + q2.b[c]*lam[2];
    }
    r.nx = 1;
    r.dist = -f.dist;
  } else {
    r.nx = 0;
    r.dist = 0;
  }
  return face;
}

// ---- the driver (mjc_ccd :2215-2343 with max_contacts = 1): geom 1 / geom 2 in the pair's type order, `margin` the
// contact margin carried by the support mappings, `cutoff` the largest distance worth reporting (0: contact only).
// Leaves the distance (negative: penetration, MJB_MAXVAL: beyond the cut-off) and the witness points in r.
MJB_NP inline void cvx_ccd(CvxRun& r, double margin, double cutoff, int type1, const double* pos1, const double* mat1,
                           const double* size1, int type2, const double* pos2, const double* mat2,
                           const double* size2, double tolerance, int maxit) {
  CvxGeom A = {type1, type1, pos1, mat1, size1, margin};
  CvxGeom B = {type2, type2, pos2, mat2, size2, margin};
  cvx_cpy(r.xa, pos1); cvx_cpy(r.xb, pos2);
  r.iter = 0; r.tolerance = tolerance; r.maxit = maxit; r.cutoff = cutoff;
  r.nx = 0; r.nsimplex = 0; r.dist = 0;
  if (type1 == MJB_GEOM_SPHERE || type2 == MJB_GEOM_SPHERE || type1 == MJB_GEOM_CAPSULE || type2 == MJB_GEOM_CAPSULE) {
    // spheres and capsules as points / segments first: their radius (and their half of the margin) comes back
    // by inflating the witness points, unless the cores themselves come closer than the tolerance
    double full1 = 0, full2 = 0;
    if (type1 == MJB_GEOM_SPHERE || type1 == MJB_GEOM_CAPSULE) {
      full1 = size1[0] + 0.5*margin;
      A.shape = type1 == MJB_GEOM_SPHERE ? MJB_CVX_POINT : MJB_CVX_SEGMENT;
      A.margin = 0;
    }
    if (type2 == MJB_GEOM_SPHERE || type2 == MJB_GEOM_CAPSULE) {
      full2 = size2[0] + 0.5*margin;
      B.shape = type2 == MJB_GEOM_SPHERE ? MJB_CVX_POINT : MJB_CVX_SEGMENT;
      B.margin = 0;
    }
    r.cutoff += full1 + full2;
    cvx_gjk(r, A, B);
    r.cutoff = cutoff;
    A.shape = type1; A.margin = margin;
    B.shape = type2; B.margin = margin;
    if (r.dist > r.tolerance) {
      double n[3];
      cvx_sub(n, r.xb, r.xa);
      normalize3(n);
      if (full1) { r.xa[0] += full1 * n[0]; r.xa[1] += full1 * n[1]; r.xa[2] += full1 * n[2]; }
      if (full2) { r.xb[0] -= full2 * n[0]; r.xb[1] -= full2 * n[1]; r.xb[2] -= full2 * n[2]; }
      r.dist -= (full1 + full2);
      if (r.dist > r.cutoff) r.dist = MJB_MAXVAL;
      return;
    }
    r.iter = 0;
    cvx_cpy(r.xa, pos1); cvx_cpy(r.xb, pos2);
  }
  cvx_gjk(r, A, B);
  if (r.dist <= tolerance && r.nsimplex > 1) {
    r.dist = 0;
    CvxPoly P;
    P.nfaces = P.ncand = P.nverts = 0;
    int bad;
    if (r.nsimplex == 2) bad = cvx_start_segment(P, r, A, B);
    else if (r.nsimplex == 3) bad = cvx_start_triangle(P, r, A, B);
    else bad = cvx_start_tetrahedron(P, r, A, B);
    if (!bad) cvx_epa(P, r, A, B);
  }
}

// the contact of a convex pair (mjc_CCDIteration, engine_collision_convex.c:791-819); returns 0 or 1
MJB_NP inline int convex_pair(Con* con, double margin, int type1, const double* pos1, const double* mat1,
                              const double* size1, int type2, const double* pos2, const double* mat2,
                              const double* size2, double tolerance, int maxit) {
  CvxRun r;
  cvx_ccd(r, margin, 0, type1, pos1, mat1, size1, type2, pos2, mat2, size2, tolerance, maxit);
  if (!(r.dist < 0)) return 0;
  if (r.nx < 1) return 0;
  Con& c = con[0];
  c.dist = margin + r.dist;
  cvx_sub(c.frame, r.xa, r.xb);
  normalize3(c.frame);
  c.pos[0] = 0.5 * (r.xa[0] + r.xb[0]); c.pos[1] = 0.5 * (r.xa[1] + r.xb[1]); c.pos[2] = 0.5 * (r.xa[2] + r.xb[2]);
  c.frame[3] = c.frame[4] = c.frame[5] = 0;
  return 1;
}

#endif  // MJB_CONVEX_H_
