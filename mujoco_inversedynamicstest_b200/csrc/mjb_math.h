// Small fp64 vector / quaternion / spatial-algebra device inlines used by the mj_inverse kernels.
// Each function keeps the operation order of the reference routine it replaces (cited), because
// contact / limit predicates downstream are compared bit-for-bit with the CPU engine.
#ifndef MJB_MATH_H_
#define MJB_MATH_H_

#include "mjb_model.h"

#if defined(__CUDACC__)
#define MJB_DI __host__ __device__ __forceinline__
#else
#include <math.h>
#define MJB_DI static inline
#endif

namespace mjb {

MJB_DI double dot3(const double* a, const double* b) {  // engine_util_blas.c:150
  return a[0]*b[0] + a[1]*b[1] + a[2]*b[2];
}

MJB_DI void cross3(double* r, const double* a, const double* b) {  // engine_util_spatial.c:371
  double t0 = a[1]*b[2] - a[2]*b[1];
  double t1 = a[2]*b[0] - a[0]*b[2];
  double t2 = a[0]*b[1] - a[1]*b[0];
  r[0] = t0; r[1] = t1; r[2] = t2;
}

// engine_util_blas.c:123  (returns the norm before normalisation)
MJB_DI double normalize3(double* v) {
  double norm = sqrt(v[0]*v[0] + v[1]*v[1] + v[2]*v[2]);
  if (norm < MJB_MINVAL) {
    v[0] = 1; v[1] = 0; v[2] = 0;
  } else {
    double inv = 1/norm;
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
  }
  return norm;
}

// engine_util_blas.c:269  (renormalises only when |norm-1| > mjMINVAL)
MJB_DI double normalize4(double* v) {
  double norm = sqrt(v[0]*v[0] + v[1]*v[1] + v[2]*v[2] + v[3]*v[3]);
  if (norm < MJB_MINVAL) {
    v[0] = 1; v[1] = 0; v[2] = 0; v[3] = 0;
  } else if (fabs(norm - 1) > MJB_MINVAL) {
    double inv = 1/norm;
    v[0] *= inv; v[1] *= inv; v[2] *= inv; v[3] *= inv;
  }
  return norm;
}

// engine_util_spatial.c:25. The reference's zero-vector / unit-quaternion shortcuts return exactly
// what the general formula returns for finite inputs, so they are not branched on here.
MJB_DI void rotVecQuat(double* res, const double* vec, const double* q) {
  double t0 = q[0]*vec[0] + q[2]*vec[2] - q[3]*vec[1];
  double t1 = q[0]*vec[1] + q[3]*vec[0] - q[1]*vec[2];
  double t2 = q[0]*vec[2] + q[1]*vec[1] - q[2]*vec[0];
  double r0 = vec[0] + 2 * (q[2]*t2 - q[3]*t1);
  double r1 = vec[1] + 2 * (q[3]*t0 - q[1]*t2);
  double r2 = vec[2] + 2 * (q[1]*t1 - q[2]*t0);
  res[0] = r0; res[1] = r1; res[2] = r2;
}

MJB_DI void mulQuat(double* res, const double* a, const double* b) {  // engine_util_spatial.c:65
  double t0 = a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3];
  double t1 = a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2];
  double t2 = a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1];
  double t3 = a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0];
  res[0] = t0; res[1] = t1; res[2] = t2; res[3] = t3;
}

MJB_DI void quat2Mat(double* res, const double* q) {  // engine_util_spatial.c:149
  const double q00 = q[0]*q[0], q01 = q[0]*q[1], q02 = q[0]*q[2], q03 = q[0]*q[3];
  const double q11 = q[1]*q[1], q12 = q[1]*q[2], q13 = q[1]*q[3];
  const double q22 = q[2]*q[2], q23 = q[2]*q[3], q33 = q[3]*q[3];
  res[0] = q00 + q11 - q22 - q33;
  res[4] = q00 - q11 + q22 - q33;
  res[8] = q00 - q11 - q22 + q33;
  res[1] = 2*(q12 - q03);
  res[2] = 2*(q13 + q02);
  res[3] = 2*(q12 + q03);
  res[5] = 2*(q23 - q01);
  res[6] = 2*(q13 - q02);
  res[7] = 2*(q23 + q01);
}

MJB_DI void mulMatVec3(double* res, const double* m, const double* v) {  // engine_util_blas.c:165
  double t0 = m[0]*v[0] + m[1]*v[1] + m[2]*v[2];
  double t1 = m[3]*v[0] + m[4]*v[1] + m[5]*v[2];
  double t2 = m[6]*v[0] + m[7]*v[1] + m[8]*v[2];
  res[0] = t0; res[1] = t1; res[2] = t2;
}

// engine_util_spatial.c:119 (out of line: atan2 is large and the callers are cold)
#if defined(__CUDACC__)
__host__ __device__ __noinline__ inline
#else
static inline
#endif
void quat2Vel(double* res, const double* q, double dt) {
  double axis[3] = {q[1], q[2], q[3]};
  double sin_a_2 = normalize3(axis);
  double speed = 2 * atan2(sin_a_2, q[0]);
  if (speed > MJB_PI) speed -= 2*MJB_PI;
  speed /= dt;
  res[0] = axis[0]*speed; res[1] = axis[1]*speed; res[2] = axis[2]*speed;
}

// engine_util_spatial.c:136   qb*quat(res) = qa
MJB_DI void subQuat(double* res, const double* qa, const double* qb) {
  double qneg[4] = {qb[0], -qb[1], -qb[2], -qb[3]};
  double qdif[4];
  mulQuat(qdif, qneg, qa);
  quat2Vel(res, qdif, 1);
}

// engine_util_spatial.c:385
MJB_DI void crossMotion(double* res, const double* vel, const double* v) {
  res[0] = -vel[2]*v[1] + vel[1]*v[2];
  res[1] =  vel[2]*v[0] - vel[0]*v[2];
  res[2] = -vel[1]*v[0] + vel[0]*v[1];
  res[3] = -vel[2]*v[4] + vel[1]*v[5];
  res[4] =  vel[2]*v[3] - vel[0]*v[5];
  res[5] = -vel[1]*v[3] + vel[0]*v[4];
  res[3] += -vel[5]*v[1] + vel[4]*v[2];
  res[4] +=  vel[5]*v[0] - vel[3]*v[2];
  res[5] += -vel[4]*v[0] + vel[3]*v[1];
}

// engine_util_spatial.c:401
MJB_DI void crossForce(double* res, const double* vel, const double* f) {
  res[0] = -vel[2]*f[1] + vel[1]*f[2];
  res[1] =  vel[2]*f[0] - vel[0]*f[2];
  res[2] = -vel[1]*f[0] + vel[0]*f[1];
  res[3] = -vel[2]*f[4] + vel[1]*f[5];
  res[4] =  vel[2]*f[3] - vel[0]*f[5];
  res[5] = -vel[1]*f[3] + vel[0]*f[4];
  res[0] += -vel[5]*f[4] + vel[4]*f[5];
  res[1] +=  vel[5]*f[3] - vel[3]*f[5];
  res[2] += -vel[4]*f[3] + vel[3]*f[4];
}

// engine_util_spatial.c:417
MJB_DI void inertCom(double* res, const double* inert, const double* mat, const double* dif,
                     double mass) {
  double tmp[9] = {mat[0]*inert[0], mat[3]*inert[0], mat[6]*inert[0],
                   mat[1]*inert[1], mat[4]*inert[1], mat[7]*inert[1],
                   mat[2]*inert[2], mat[5]*inert[2], mat[8]*inert[2]};
  res[0] = mat[0]*tmp[0] + mat[1]*tmp[3] + mat[2]*tmp[6];
  res[1] = mat[3]*tmp[1] + mat[4]*tmp[4] + mat[5]*tmp[7];
  res[2] = mat[6]*tmp[2] + mat[7]*tmp[5] + mat[8]*tmp[8];
  res[3] = mat[0]*tmp[1] + mat[1]*tmp[4] + mat[2]*tmp[7];
  res[4] = mat[0]*tmp[2] + mat[1]*tmp[5] + mat[2]*tmp[8];
  res[5] = mat[3]*tmp[2] + mat[4]*tmp[5] + mat[5]*tmp[8];
  res[0] += mass*(dif[1]*dif[1] + dif[2]*dif[2]);
  res[1] += mass*(dif[0]*dif[0] + dif[2]*dif[2]);
  res[2] += mass*(dif[0]*dif[0] + dif[1]*dif[1]);
  res[3] -= mass*dif[0]*dif[1];
  res[4] -= mass*dif[0]*dif[2];
  res[5] -= mass*dif[1]*dif[2];
  res[6] = mass*dif[0];
  res[7] = mass*dif[1];
  res[8] = mass*dif[2];
  res[9] = mass;
}

// engine_util_spatial.c:452
MJB_DI void mulInertVec(double* res, const double* i, const double* v) {
  res[0] = i[0]*v[0] + i[3]*v[1] + i[4]*v[2] - i[8]*v[4] + i[7]*v[5];
  res[1] = i[3]*v[0] + i[1]*v[1] + i[5]*v[2] + i[8]*v[3] - i[6]*v[5];
  res[2] = i[4]*v[0] + i[5]*v[1] + i[2]*v[2] - i[7]*v[3] + i[6]*v[4];
  res[3] = i[8]*v[1] - i[7]*v[2] + i[9]*v[3];
  res[4] = i[6]*v[2] - i[8]*v[0] + i[9]*v[4];
  res[5] = i[7]*v[0] - i[6]*v[1] + i[9]*v[5];
}

// ---- explicitly fused variants -------------------------------------------------------------
// libmjb is compiled WITHOUT automatic fp contraction (nvcc/NVRTC -fmad=false), like the reference
// (its x86-64 builds have no FMA): every value that feeds a contact / limit predicate or a stiff
// constraint row (poses, distances, frames, J*v, J*qacc) then carries exactly the reference's
// rounding, which the row arithmetic amplifies by D*K ~ 1e6-1e8. The two leaves-to-root sweeps
// (mj_crb + mj_factorM, mj_rne backward) only form sums that go straight into outputs, so they use
// these fused forms: same values to a few ulp, 30 % fewer fp64 instructions, and -- being explicit
// -- the same bits whichever compiler (nvcc, NVRTC) or kernel variant evaluates them.
MJB_DI double dot6f(const double* a, const double* b) {
  double r = a[0]*b[0];
  r = fma(a[1], b[1], r); r = fma(a[2], b[2], r); r = fma(a[3], b[3], r);
  r = fma(a[4], b[4], r); r = fma(a[5], b[5], r);
  return r;
}

MJB_DI void mulInertVecF(double* res, const double* i, const double* v) {   // mulInertVec, fused
  res[0] = fma(i[7], v[5], fma(-i[8], v[4], fma(i[4], v[2], fma(i[3], v[1], i[0]*v[0]))));
  res[1] = fma(-i[6], v[5], fma(i[8], v[3], fma(i[5], v[2], fma(i[1], v[1], i[3]*v[0]))));
  res[2] = fma(i[6], v[4], fma(-i[7], v[3], fma(i[2], v[2], fma(i[5], v[1], i[4]*v[0]))));
  res[3] = fma(i[9], v[3], fma(-i[7], v[2], i[8]*v[1]));
  res[4] = fma(i[9], v[4], fma(-i[8], v[0], i[6]*v[2]));
  res[5] = fma(i[9], v[5], fma(-i[6], v[1], i[7]*v[0]));
}

// engine_util_spatial.c:526
MJB_DI void makeFrame(double* frame) {
  normalize3(frame);
  double ynorm = sqrt(frame[3]*frame[3] + frame[4]*frame[4] + frame[5]*frame[5]);
  if (ynorm < 0.5) {
    frame[3] = 0; frame[4] = 0; frame[5] = 0;
    if (frame[1] < 0.5 && frame[1] > -0.5) frame[4] = 1; else frame[5] = 1;
  }
  double dt = dot3(frame, frame + 3);
  frame[3] -= frame[0]*dt;
  frame[4] -= frame[1]*dt;
  frame[5] -= frame[2]*dt;
  normalize3(frame + 3);
  cross3(frame + 6, frame, frame + 3);
}

}  // namespace mjb

#endif  // MJB_MATH_H_
