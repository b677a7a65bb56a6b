// The mj_inverse pipeline for ONE state, written for one CUDA thread per state.
//
// Every function below is the per-thread body of a stage of MuJoCo 3.3.1's mj_inverse
// (reference src/engine/engine_inverse.c:197-261); all control flow that depends only on the
// model (tree topology, joint types, candidate geom pairs) is warp-uniform, so the 32 lanes of a
// warp execute the same instruction stream on 32 different states. Intermediates live in a
// per-state scratch laid out [slot][NS] (NS = states per chunk) so that lanes touch consecutive doubles.
//
// Constraint rows are evaluated WITHOUT forming efc_J: a contact row's J*qvel / J*qacc is the
// relative spatial velocity / acceleration of the two bodies at the contact point projected on
// the contact frame (identical to jacdif*q by mj_jac, engine_support.c:389-439), and J'*force is
// applied as a wrench on the two bodies and folded into the backward pass of mj_rne. Joint and
// tendon limit rows touch single dofs directly.
//
// The file compiles as CUDA device code (nvcc) and, for the CPU-side unit tests only, as plain
// C++ (tests/hostemu); libmjb.so never contains or calls the host build.
#ifndef MJB_PIPELINE_H_
#define MJB_PIPELINE_H_

#include <stddef.h>

#include "mjb_math.h"
#include "mjb_model.h"

#if defined(__CUDACC__)
#ifdef MJB_SPECIALIZED
// specialised build: every function that takes the per-state context is inlined, so that the
// context stays in registers and its table pointers stay compile-time constants in the callee
#define MJB_HD __host__ __device__ __attribute__((always_inline))
#else
#define MJB_HD __host__ __device__
#endif
// rarely taken, code-heavy leaf paths (pow, atan2) are kept out of line to keep the phase kernels
// small; only functions of scalars qualify (an out-of-line call taking Ctx& would force the whole
// context into local memory: measured +45% on the smooth kernel)
#define MJB_COLD __host__ __device__ __noinline__
#ifdef MJB_NARROW_NOINLINE
#define MJB_NP __host__ __device__ __noinline__
#else
#define MJB_NP __host__ __device__
#endif
#else
#define MJB_HD
#define MJB_NP
#define MJB_COLD
#endif

// warp votes used to keep compaction loops warp-uniform; identity in the single-lane host build
#if defined(__CUDA_ARCH__)
#define MJB_WARP_ANY(p) (__any_sync(0xffffffffu, (p)) != 0)
#define MJB_WARP_MAX(x) (__reduce_max_sync(0xffffffffu, (x)))
#else
#define MJB_WARP_ANY(p) (p)
#define MJB_WARP_MAX(x) (x)
#endif

// Model-specialised build (csrc/mjb_spec_kernels.h, compiled per model at mjb_makeData by NVRTC): the
// model blob is a `__device__ const` array in the translation unit, so every table look-up whose
// index is a compile-time constant folds to an immediate. The loops over bodies are then expanded
// at compile time (MJB_BODY_LOOP_*: template recursion calling the loop body with a constant
// index; the optimiser's own `#pragma unroll` gives up on bodies of this size) and the small
// model-bounded inner loops (joints / dofs / geoms of one body, tendon terms) carry MJB_UNROLL.
// In the generic build both are plain loops.
#ifdef MJB_SPECIALIZED
#define MJB_UNROLL _Pragma("unroll")
#define MJB_BODY_LAMBDA __attribute__((always_inline))   // every expansion of a loop body is inlined
#else
#define MJB_UNROLL
#define MJB_BODY_LAMBDA
#endif

namespace mjb {

#ifdef MJB_SPECIALIZED
template <int B, int E, typename F>
MJB_DI void static_for_up(F& f) { if constexpr (B < E) { f(B); static_for_up<B + 1, E>(f); } }
template <int B, int E, typename F>
MJB_DI void static_for_down(F& f) { if constexpr (B > E) { f(B); static_for_down<B - 1, E>(f); } }
// body(b) for b = lo .. hi-1 / b = hi-1 .. lo ; kLo, kHi are the compile-time values of lo, hi
#define MJB_BODY_LOOP_UP(body, lo, hi, kLo, kHi) static_for_up<kLo, kHi>(body)
#define MJB_BODY_LOOP_DOWN(body, lo, hi, kLo, kHi) static_for_down<(kHi) - 1, (kLo) - 1>(body)
#else
#define MJB_BODY_LOOP_UP(body, lo, hi, kLo, kHi) for (int b_ = (lo); b_ < (hi); b_++) body(b_)
#define MJB_BODY_LOOP_DOWN(body, lo, hi, kLo, kHi) for (int b_ = (hi) - 1; b_ >= (lo); b_--) body(b_)
#endif

// status bits (mirrored in include/mjb.h)
enum { kStatusBadQpos = 1, kStatusBadQvel = 2, kStatusBadQacc = 4, kStatusContactFull = 8,
       kStatusCnstrFull = 16 };

// optional per-state outputs, all structure-of-arrays [row][stride]
struct Outputs {
  double* qfrc_inverse;     // [nv][stride]            always
  double* qfrc_constraint;  // [nv][stride]            or null
  double* qfrc_passive;     // [nv][stride]            or null
  int* counts;              // [5][stride]: ncon, ne, nf, nl, nefc          or null
  int* status;              // [stride] bit flags
  // contacts, nconmax rows per state (null when not requested)
  int* contact_geom;        // [nconmax*2][stride]
  int* contact_info;        // [nconmax*3][stride]: dim, exclude, efc_address
  double* contact_num;      // [nconmax*13][stride]: dist, pos[3], frame[9]
  // constraint rows, njmax rows per state (null when not requested)
  int* efc_int;             // [njmax*3][stride]: type, id, state
  double* efc_num;          // [njmax*8][stride]: pos, margin, D, R, vel, aref, force, diagApprox
  // inertia (null when not requested)
  double* qM;               // [nM][stride]
  double* qLD;              // [nC][stride]
  double* qLDiagInv;        // [nv][stride]
  // debug dump: every scratch slot copied out as [nscratch][stride] (null when not requested)
  double* scratch_dump;
  // mj_rnePostConstraint (engine_core_smooth.c:2027-2181), all three or none, [nbody*6][stride],
  // in the reference's frame (origin at subtree_com of the body's kinematic tree)
  double* cacc;
  double* cfrc_int;
  double* cfrc_ext;
  // sensordata [nsensordata][stride] (mj_sensorPos / Vel / Acc), null for models without sensors
  double* sensordata;
  // qfrc_bias [nv][stride] = mj_rne(flg_acc = 0) (engine_forward.c:193-231), or null
  double* qfrc_bias;
  // mj_compareFwdInv (engine_inverse.c:275-316): quantities of the forward pass to compare with
  // (inputs) and the two norms (output); all null outside mjb_compareFwdInv
  const double* fwd_qforce;           // [nv][stride]      qfrc_applied + qfrc_actuator
  const double* fwd_xfrc;             // [nbody*6][stride] xfrc_applied: force, torque per body; or null
  const double* fwd_qfrc_constraint;  // [nv][stride]
  double* fwdinv;                     // [2][stride]
  // per-state mocap poses (d->mocap_pos / d->mocap_quat, INPUTS; engine_core_smooth.c:70-86), or null:
  // mocap bodies then sit at their model pose, as after mj_resetData
  const double* mocap_pos;            // [nmocap*3][stride]
  const double* mocap_quat;           // [nmocap*4][stride]
  // d->energy [2][stride] (mj_energyPos, mj_energyVel), only for models with mjENBL_ENERGY; else null
  double* energy;
  // mj_camlight (engine_core_smooth.c:275-389), all four or none (mjbOUT_CAMLIGHT)
  double* cam_xpos;                   // [ncam*3][stride]
  double* cam_xmat;                   // [ncam*9][stride]
  double* light_xpos;                 // [nlight*3][stride]
  double* light_xdir;                 // [nlight*3][stride]
  // mj_transmission + actuator_velocity of mj_fwdVelocity, all three or none (mjbOUT_TRANSMISSION)
  double* actuator_length;            // [nu][stride]
  double* actuator_moment;            // [nu*nv][stride], dense rows
  double* actuator_velocity;          // [nu][stride]
  // d->xfrc_applied per state (INPUT; mjb_setXfrcApplied): [nbody*6][stride], force then torque per body.
  // mj_inverse itself ignores it; mj_rnePostConstraint adds it to cfrc_ext (engine_core_smooth.c:2039-2049)
  const double* xfrc_applied;
  // d->eq_active per state (INPUT; mjb_setEqActive): [neq][stride], 0 / 1 as doubles; null: the model's eq_active0
  const double* eq_active;
};

struct Ctx {
  const mjbHdr* H;
  const int* I;         // int section of the model blob
  const double* D;      // double section of the model blob
  double* sm;           // per-thread on-chip slots (shared memory, stride MJB_SMS): forward-sweep carry
  double* sc;           // double scratch [nscratch][NS], already offset to this state
  int* isc;             // int scratch [MJB_ISC_COUNT][NS], already offset to this state
  const double* qpos;   // already offset to this state
  const double* qvel;
  const double* qacc;
  long long N;          // stride of state-indexed arrays
  long long s;          // state index
  int nconmax, njmax;
  Outputs out;
  int ncon, ne, nf, nl, nefc, status;
  // Fused subtree stages of the specialised build (phase_tree): cinert of bodies >= lbody0 and cdof
  // of dofs >= ldof0 are handed from the forward sweep to the inertia sweep of the SAME kernel in
  // these thread-local rows (registers) instead of the scratch; null otherwise.
  double* lci;
  double* lcd;
  int lbody0, ldof0;
};

// lane stride of the per-state scratch: device scratch is blocked per warp, element (slot, state i)
// at ((i/32)*nslots + slot)*32 + i%32, so the 32 lanes of a warp touch one 256-byte line per slot and
// a warp's whole working set is one contiguous block (slot offsets fold into load/store immediates)
#if defined(__CUDACC__)
#define MJB_LS 32
#else
#define MJB_LS 1
#endif

#ifndef MJB_ANC
#define MJB_ANC 2
#endif

// per-thread shared-memory slots of the forward sweep (parent carry): stride between slots
#if defined(__CUDACC__)
#ifndef MJB_SMOOTH_THREADS
#define MJB_SMOOTH_THREADS 256
#endif
#define MJB_SMS MJB_SMOOTH_THREADS
#else
#define MJB_SMS 1
#endif
#define MJB_SM_SLOTS 25   // P[3] Q[4] V[6] A[6] AL[6]

#define MI(name) (c.I + c.H->ioff[MJB_I_##name])
#define MD(name) (c.D + c.H->noff[MJB_N_##name])
#define SC(name) (c.sc + (size_t)c.H->scoff[MJB_SC_##name] * MJB_LS)
#define AT(p, k) (p)[(size_t)(k) * MJB_LS]
#define QPOS(i) c.qpos[(size_t)(i) * (size_t)c.N]
#define QVEL(i) c.qvel[(size_t)(i) * (size_t)c.N]
#define QACC(i) c.qacc[(size_t)(i) * (size_t)c.N]

MJB_DI void ldn_(double* dst, const double* p, int first, int n, size_t stride) {
  for (int k = 0; k < n; k++) dst[k] = p[(size_t)(first + k) * stride];
}
// rows that the running kernel only reads (written by an earlier kernel): ld.global.nc, which tells
// the compiler that no store of this kernel can alias them, so the loads can be hoisted above the
// output stores and repeated loads of the same row (ancestor cdofs) are merged
MJB_DI void ldn_ro_(double* dst, const double* p, int first, int n, size_t stride) {
#if defined(__CUDA_ARCH__)
  for (int k = 0; k < n; k++) dst[k] = __ldg(p + (size_t)(first + k) * stride);
#else
  for (int k = 0; k < n; k++) dst[k] = p[(size_t)(first + k) * stride];
#endif
}
MJB_DI void stn_(double* p, int first, const double* src, int n, size_t stride) {
  for (int k = 0; k < n; k++) p[(size_t)(first + k) * stride] = src[k];
}
// Stores of rows that only a LATER kernel reads (cdof, cinert, cfrc, geom poses ...): evict-first
// (st.global.cs) so that they do not push the register-spill lines and the inputs out of L1.
// MJB_STREAM_HINTS: 0 off, 1 those rows, 2 every store of the forward sweep.
#ifndef MJB_STREAM_HINTS
#define MJB_STREAM_HINTS 1
#endif
MJB_DI void stn_stream_(double* p, int first, const double* src, int n, size_t stride) {
#if defined(__CUDA_ARCH__) && MJB_STREAM_HINTS
  for (int k = 0; k < n; k++) __stcs(p + (size_t)(first + k) * stride, src[k]);
#else
  for (int k = 0; k < n; k++) p[(size_t)(first + k) * stride] = src[k];
#endif
}
#define sts(p, first, src, n) stn_stream_(p, first, src, n, MJB_LS)
#if MJB_STREAM_HINTS >= 2
#define stc(p, first, src, n) stn_stream_(p, first, src, n, MJB_LS)
#else
#define stc(p, first, src, n) stn_(p, first, src, n, MJB_LS)
#endif
#define ldn(dst, p, first, n) ldn_(dst, p, first, n, MJB_LS)
#define ldn_ro(dst, p, first, n) ldn_ro_(dst, p, first, n, MJB_LS)
#define stn(p, first, src, n) stn_(p, first, src, n, MJB_LS)

// Contact carrier records (MJB_SC_crec, mjb_model.h): 16 doubles per (body, state), state-major
// inside the warp block, written by the forward sweep with four 256-bit stores per body (each lane
// fills whole 32-byte sectors) and gathered by the contact rows with four 256-bit loads from ONE
// 128-byte line (the row layout costs a contact 15 sectors of 8 useful bytes per body).
// [0..5] cvel, [6..11] cacc_lin, [12..14] tree origin. Measured on 2^20 humanoid states: contact rows
// 1.96 -> 1.09 ms, forward sweep 2.28 -> 2.66 ms (strided 32-byte stores); the part-major variant
// (part q of 32 states contiguous: coalesced stores, four lines per gather) gave 2.50 and 1.37 ms.
#define MJB_CREC_PART 4                 // doubles between the parts of one record
MJB_HD inline double* crec_ptr(Ctx& c, int b) {
#if defined(__CUDA_ARCH__)
  // c.sc = (256-byte aligned block base) + lane: recover the lane from the address
  const size_t ln = ((size_t)c.sc >> 3) & 31;
  return c.sc - ln + ((size_t)c.H->scoff[MJB_SC_crec] + 16*(size_t)b) * 32 + ln*16;
#else
  return c.sc + ((size_t)c.H->scoff[MJB_SC_crec] + 16*(size_t)b);
#endif
}
// 256-bit global accesses exist from PTX ISA 8.8 (CUDA 12.9) on; older NVRTCs get 128-bit pairs
#if defined(__CUDA_ARCH__) && (__CUDACC_VER_MAJOR__ > 12 || (__CUDACC_VER_MAJOR__ == 12 && __CUDACC_VER_MINOR__ >= 9))
#define MJB_WIDE256 1
#else
#define MJB_WIDE256 0
#endif
MJB_DI void st_rec4(double* p, double a, double b, double cc, double d) {
#if defined(__CUDA_ARCH__)
#if MJB_WIDE256
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(a), "d"(b), "d"(cc), "d"(d) : "memory");
#else
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" :: "l"(p), "d"(a), "d"(b) : "memory");
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" :: "l"(p + 2), "d"(cc), "d"(d) : "memory");
#endif
#else
  p[0] = a; p[1] = b; p[2] = cc; p[3] = d;
#endif
}
MJB_DI void ld_rec4(double* dst, const double* p) {
#if defined(__CUDA_ARCH__)
#if MJB_WIDE256
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];"
               : "=d"(dst[0]), "=d"(dst[1]), "=d"(dst[2]), "=d"(dst[3]) : "l"(p) : "memory");
#else
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(dst[0]), "=d"(dst[1]) : "l"(p) : "memory");
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(dst[2]), "=d"(dst[3]) : "l"(p + 2) : "memory");
#endif
#else
  dst[0] = p[0]; dst[1] = p[1]; dst[2] = p[2]; dst[3] = p[3];
#endif
}
MJB_HD inline void store_crec(Ctx& c, int b, const double* V, const double* AL, const double* O) {
  double* r = crec_ptr(c, b);
  st_rec4(r, V[0], V[1], V[2], V[3]);
  st_rec4(r + MJB_CREC_PART, V[4], V[5], AL[0], AL[1]);
  st_rec4(r + 2*MJB_CREC_PART, AL[2], AL[3], AL[4], AL[5]);
  st_rec4(r + 3*MJB_CREC_PART, O[0], O[1], O[2], 0.0);
}
// Geom position / z axis as 4-double vectors per (geom, state) (MJB_SC_geom_xpos, MJB_SC_geom_zaxis): the sweep
// writes each with one coalesced 256-bit store, the scans read them with one coalesced 256-bit load per
// geom and the item-parallel narrow phase gathers ONE sector per vector (the row layout: three).
MJB_HD inline double* geom_vec(Ctx& c, int slot, int g) {
#if defined(__CUDA_ARCH__)
  const size_t ln = ((size_t)c.sc >> 3) & 31;
  return c.sc - ln + ((size_t)c.H->scoff[slot] + 4*(size_t)g) * 32 + ln*4;
#else
  return c.sc + ((size_t)c.H->scoff[slot] + 4*(size_t)g);
#endif
}
MJB_HD inline void load_geom_pos(Ctx& c, int g, double* p3) {
  double v[4];
  ld_rec4(v, geom_vec(c, MJB_SC_geom_xpos, g));
  p3[0] = v[0]; p3[1] = v[1]; p3[2] = v[2];
}
MJB_HD inline void load_geom_z(Ctx& c, int g, double* z3) {
  double v[4];
  ld_rec4(v, geom_vec(c, MJB_SC_geom_zaxis, g));
  z3[0] = v[0]; z3[1] = v[1]; z3[2] = v[2];
}
// rec[16] of body b; a body without a dof on its chain to the world has zero carriers and no record
MJB_HD inline void load_crec(Ctx& c, int b, bool is_static, double* rec) {
  if (is_static) {
    for (int k = 0; k < 16; k++) rec[k] = 0;
    return;
  }
  const double* r = crec_ptr(c, b);
  ld_rec4(rec, r); ld_rec4(rec + 4, r + MJB_CREC_PART); ld_rec4(rec + 8, r + 2*MJB_CREC_PART); ld_rec4(rec + 12, r + 3*MJB_CREC_PART);
}

// per-state integer scratch rows: counters carried between the phase kernels
// MJB_ISC_NSURV / MJB_ISC_MASK..: survivors of the contact scan (count, then ceil(ncand/32) words)
// MJB_ISC_ITEMBASE: first entry of the state's survivors in the chunk's global item list (-1: none)
enum { MJB_ISC_NCON = 0, MJB_ISC_NE, MJB_ISC_NF, MJB_ISC_NL, MJB_ISC_NEFC, MJB_ISC_STATUS,
       MJB_ISC_NSURV, MJB_ISC_ITEMBASE, MJB_ISC_MASK, MJB_ISC_COUNT = MJB_ISC_MASK };

// Wrench-accumulator masks: the per-body constraint wrench rows cfrc_ext ('+' side) and cfrc_ext1
// ('-' side) are NOT cleared per state; a body's row is valid only if its bit is set in the state's
// mask (one bit per body and side, after the survivor words). The forward sweep clears the masks
// (a few ints instead of 12*nbody doubles), whoever adds the first wrench to a body starts from zero
// and sets the bit, and the backward sweep reads only rows whose bit is set.
MJB_HD inline int isc_wmask_row(const mjbHdr& H) { return MJB_ISC_MASK + (H.ncand + 31) / 32 + 1; }
MJB_HD inline int isc_wmask_words(const mjbHdr& H) { return (H.nbody + 31) / 32; }
MJB_HD inline int isc_rows(const mjbHdr& H) { return isc_wmask_row(H) + 2 * isc_wmask_words(H); }

// sets the bit of (body, side) and returns whether it was already set. Lanes of one warp may work
// on different bodies of the SAME state (pooled / item-parallel contact kernels): atomic on device.
MJB_HD inline bool wmask_test_and_set(Ctx& c, int body, bool positive, bool shared = true) {
  const mjbHdr& H = *c.H;
  int* w = c.isc + (size_t)(isc_wmask_row(H) + (positive ? 0 : isc_wmask_words(H)) + (body >> 5)) * MJB_LS;
  const int bit = (int)(1u << (body & 31));
#if defined(__CUDA_ARCH__)
  if (shared) return (atomicOr(w, bit) & bit) != 0;
  const bool was = (*w & bit) != 0;      // thread-per-state kernel: the word belongs to this thread
  *w |= bit;
  return was;
#else
  const bool was = (*w & bit) != 0;
  *w |= bit;
  return was;
#endif
}
MJB_HD inline bool wmask_test(Ctx& c, int body, bool positive) {
  const mjbHdr& H = *c.H;
  const int* w = c.isc + (size_t)(isc_wmask_row(H) + (positive ? 0 : isc_wmask_words(H)) + (body >> 5)) * MJB_LS;
  return ((*w >> (body & 31)) & 1) != 0;
}
MJB_HD inline void wmask_clear(Ctx& c) {
  const mjbHdr& H = *c.H;
  const int n = 2 * isc_wmask_words(H), row = isc_wmask_row(H);
  for (int k = 0; k < n; k++) c.isc[(size_t)(row + k) * MJB_LS] = 0;
}

MJB_HD inline void save_counters(Ctx& c) {
  c.isc[MJB_ISC_NCON * MJB_LS] = c.ncon; c.isc[MJB_ISC_NE * MJB_LS] = c.ne;
  c.isc[MJB_ISC_NF * MJB_LS] = c.nf; c.isc[MJB_ISC_NL * MJB_LS] = c.nl;
  c.isc[MJB_ISC_NEFC * MJB_LS] = c.nefc; c.isc[MJB_ISC_STATUS * MJB_LS] = c.status;
}
MJB_HD inline void load_counters(Ctx& c) {
  c.ncon = c.isc[MJB_ISC_NCON * MJB_LS]; c.ne = c.isc[MJB_ISC_NE * MJB_LS];
  c.nf = c.isc[MJB_ISC_NF * MJB_LS]; c.nl = c.isc[MJB_ISC_NL * MJB_LS];
  c.nefc = c.isc[MJB_ISC_NEFC * MJB_LS]; c.status = c.isc[MJB_ISC_STATUS * MJB_LS];
}

// engine_util_blas.c:677 with n == 6 (same association as the reference's 4-lane order)
MJB_DI double dot6(const double* a, const double* b) {
  double res = (a[0]*b[0] + a[2]*b[2]) + (a[1]*b[1] + a[3]*b[3]);
  res += a[4]*b[4] + a[5]*b[5];
  return res;
}

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

template <int kLo = 1, int kHi = 0, bool kHandOver = true>
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
template <bool kSimple = false>
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
    double v[4] = {0, 0, 0, 0};
    if (type == MJB_SENS_TOUCH) {
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
    for (int k = 0; k < si[MJB_SEN_DIM] && k < 4; k++) {
      double x = v[k];
      if (cut > 0) {
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

// ------------------------------------------------------------------------------------------
// mj_inverseSkip(m, d, mjSTAGE_NONE, skipsensor=1) for one state (engine_inverse.c:197-261), cut
// into four phases that run as separate kernels (each with its own register budget / occupancy)
// and hand their intermediates over through the per-state scratch in HBM:
//   smooth   : mj_kinematics, mj_comPos, fixed tendons, mj_comVel, mj_passive,
//              friction-loss and limit rows of mj_makeConstraint .. mj_invConstraint
//   inertia  : mj_crb, mj_factorM                      (only when qM/qLD/qLDiagInv are requested)
//   contact  : mj_collision + contact rows             (only when contacts are enabled)
//   backward : mj_rne(flg_acc=1), J'f, final combine, output bookkeeping

template <bool kSpatial>
MJB_HD inline void smooth_tail(Ctx& c);

template <bool kSpatial, int kLo = 1, int kHi = 0>
MJB_HD inline void phase_smooth(Ctx& c) {
  const mjbHdr& H = *c.H;
  if (kLo == 1) {
    c.ncon = c.ne = c.nf = c.nl = c.nefc = 0;
    c.status = 0;
  } else {
    load_counters(c);      // a later stage of the sweep: running limit-row count and status bits
  }
  forward_sweep<kLo, kHi>(c);   // incl. input checks, joint springs/dampers, dof friction and joint limit rows
  if (kHi != 0 && kHi < H.nbody) { save_counters(c); return; }
  smooth_tail<kSpatial>(c);
}

// what follows the sweep over the last body: tendons, equality rows, the row counters
template <bool kSpatial>
MJB_HD inline void smooth_tail(Ctx& c) {
  const mjbHdr& H = *c.H;
  tendon_kinematics<kSpatial>(c);
  passive_tendons<kSpatial>(c);
  wmask_clear(c);          // no wrench on any body yet (the accumulator rows themselves are not cleared)
  if (rows_enabled(H)) {
    equality_rows(c);      // rows [0, ne)
    tendon_friction_rows<kSpatial>(c);
    tendon_limit_rows<kSpatial>(c);
    c.nf = H.nf_rows;
  }
  c.nefc = c.ne + c.nf + c.nl;
  save_counters(c);
}

template <int kLo = 1, int kHi = 0>
MJB_HD inline void phase_inertia(Ctx& c) { inertia<kLo, kHi>(c); }

// Active joint-limit rows of all joints in front of body `lo`, from the inputs alone (a limit is
// active iff dist < margin, which depends on qpos and the model only: same expressions as
// joint_limit_rows / quat_dof_forces). The tree stages below run body ranges in an order that is
// not the body order, while a limit row's index is its rank in body order.
MJB_HD inline int limit_rows_before(Ctx& c, int lo) {
  const mjbHdr& H = *c.H;
  if ((H.disableflags & MJB_DSBL_LIMIT) || !rows_enabled(H)) return 0;
  const int* jnt_type = MI(jnt_type); const int* jnt_limited = MI(jnt_limited);
  const int* jnt_qposadr = MI(jnt_qposadr); const int* jnt_bodyid = MI(jnt_bodyid);
  const double* jnt_range = MD(jnt_range); const double* jnt_margin = MD(jnt_margin);
  int n = 0;
  MJB_UNROLL
  for (int j = 0; j < H.njnt; j++) {
    if (jnt_bodyid[j] >= lo || !jnt_limited[j]) continue;    // joints are ordered by body
    const int jt = jnt_type[j], qadr = jnt_qposadr[j];
    const double margin = jnt_margin[j];
    if (jt == MJB_JNT_HINGE || jt == MJB_JNT_SLIDE) {
      const double q = QPOS(qadr);
      for (int side = -1; side <= 1; side += 2) {
        const double dist = side * (jnt_range[2*j + (side + 1)/2] - q);
        if (dist < margin) n++;
      }
    } else if (jt == MJB_JNT_BALL) {
      double quat[4], aa[3];
      for (int k = 0; k < 4; k++) quat[k] = QPOS(qadr + k);
      normalize4(quat);
      quat2Vel(aa, quat, 1);
      const double value = normalize3(aa);
      const double dist = fmax(jnt_range[2*j], jnt_range[2*j + 1]) - value;
      if (dist < margin) n++;
    }
  }
  return n;
}

// Tree stages of the model-specialised build (mjb_jit.cu plans them, mjb_spec_kernels.cuh launches
// one kernel per stage). The tree is cut into a TRUNK (bodies whose subtree is too large for one
// kernel) and complete SUBTREES:
//   kTreeFwd   : forward sweep over a range of trunk bodies                 (runs first, ascending)
//   kTreeFused : forward sweep over a range of complete subtrees, then mj_crb + mj_factorM over the
//                same range while cinert / cdof of its bodies are still in registers; the composite
//                inertias are pushed to the (trunk) parent through its scratch accumulators
//   kTreeBwd   : mj_crb + mj_factorM over a range of trunk bodies
// Fused and Bwd stages run in DESCENDING body order, which is the order the accumulator protocol of
// inertia() assumes (highest-index child first). kTail adds what follows the last body of the
// sweep (tendons, equality rows, row counters); it is attached to the stage that runs last.
// cinert of subtree bodies is never written to HBM, and cdof / cinert are not read back from it.
enum { kTreeFwd = 0, kTreeFused = 1, kTreeBwd = 2 };

template <bool kSpatial, int kMode, int kLo, int kHi, bool kFirst, bool kTail, int kBodies, int kDof0, int kDofs>
MJB_HD inline void phase_tree(Ctx& c) {
  const mjbHdr& H = *c.H;
  double lci[kMode == kTreeFused ? 10*kBodies : 1];
  double lcd[kMode == kTreeFused ? 6*(kDofs > 0 ? kDofs : 1) : 1];
  if (kMode != kTreeBwd) {
    if (kFirst) {
      c.ncon = c.ne = c.nf = c.nl = c.nefc = 0;
      c.status = 0;
    } else {
      load_counters(c);
    }
    // limit rows are numbered in body order: number this range's rows from the rank of its first row
    const int total_before = c.nl;
    const int prefix = limit_rows_before(c, kLo);
    c.nl = prefix;
    if (kMode == kTreeFused) {
      c.lci = lci; c.lcd = lcd;
      c.lbody0 = kLo; c.ldof0 = kDof0;
      forward_sweep<kLo, kHi, false>(c);
    } else {
      forward_sweep<kLo, kHi, true>(c);
    }
    c.nl = total_before + (c.nl - prefix);
  }
  if (kMode != kTreeFwd) {
    inertia<kLo, kHi>(c);
    c.lci = nullptr; c.lcd = nullptr;
  }
  if (kTail) {
    if (kMode == kTreeBwd) load_counters(c);
    smooth_tail<kSpatial>(c);      // ends with save_counters
  } else if (kMode != kTreeBwd) {
    save_counters(c);
  }
}

MJB_HD inline bool contacts_enabled(const mjbHdr& H) {
  return !(H.disableflags & (MJB_DSBL_CONSTRAINT | MJB_DSBL_CONTACT)) && H.ncand > 0;
}

MJB_HD inline void phase_contact(Ctx& c, bool valid, int* list, int lstride, int cap) {
  if (valid) load_counters(c);
  contact_process(c, valid, list, lstride, cap);
  if (valid) save_counters(c);
}

template <bool kGravcomp>
MJB_HD inline void phase_backward(Ctx& c) {
  const mjbHdr& H = *c.H;
  load_counters(c);
  rne_and_output<kGravcomp>(c);

  const size_t N = (size_t)c.N;
  if (c.out.counts) {
    c.out.counts[0*N + c.s] = c.ncon;
    c.out.counts[1*N + c.s] = c.ne;
    c.out.counts[2*N + c.s] = c.nf;
    c.out.counts[3*N + c.s] = c.nl;
    c.out.counts[4*N + c.s] = c.nefc;
  }
  // unused rows of the fixed-capacity outputs: -1 for ids, 0 for numbers
  if (c.out.contact_geom) {
    for (int k = c.ncon; k < c.nconmax; k++) {
      for (int j = 0; j < 2; j++) c.out.contact_geom[(size_t)(2*k + j)*N + c.s] = -1;
      for (int j = 0; j < 3; j++) c.out.contact_info[(size_t)(3*k + j)*N + c.s] = -1;
      for (int j = 0; j < 13; j++) c.out.contact_num[(size_t)(13*k + j)*N + c.s] = 0;
    }
  }
  if (c.out.efc_int) {
    for (int k = c.nefc; k < c.njmax; k++) {
      for (int j = 0; j < 3; j++) c.out.efc_int[(size_t)(3*k + j)*N + c.s] = -1;
      for (int j = 0; j < 8; j++) c.out.efc_num[(size_t)(8*k + j)*N + c.s] = 0;
    }
  }
  if (c.out.status) c.out.status[c.s] = c.status;
  if (c.out.scratch_dump) {
    for (int k = 0; k < H.nscratch; k++) c.out.scratch_dump[(size_t)k*N + c.s] = AT(c.sc, k);
  }
}

// all phases for one state in sequence (single-lane host build of the tests)
MJB_HD inline void inverse_one_state(Ctx& c, double* qacc_discrete = nullptr) {
  int list[64];
  phase_smooth<true>(c);
  if (c.out.qM || c.out.qLD || c.out.qLDiagInv) phase_inertia(c);
  if (c.H->discrete_acc) {
    // converted accelerations replace qacc for everything that follows (engine_inverse.c:227-252)
    if (c.H->discrete_trn) transmission(c);
    discrete_acc(c, qacc_discrete + c.s);
    c.qacc = qacc_discrete + c.s;
    phase_smooth<true>(c);
  }
  if (contacts_enabled(*c.H)) {
    contact_scan(c);
    phase_contact(c, true, list, 1, 64);
  }
  if (c.H->passive_wrench) phase_backward<true>(c); else phase_backward<false>(c);
  if (c.out.xfrc_applied && c.out.cfrc_ext) post_xfrc(c);
  if (c.out.qfrc_bias) bias_forces(c);
  if (c.out.energy) energy(c);
  if (c.out.cam_xpos) camlight(c);
  if (c.out.actuator_length) transmission(c);
  if (c.out.sensordata) sensors(c);
  if (c.out.fwdinv) compare_fwdinv(c);
}

#undef MI
#undef MD
#undef SC
#undef AT
#undef ldn
#undef ldn_ro
#undef sts
#undef stc
#undef stn
#undef QPOS
#undef QVEL
#undef QACC

}  // namespace mjb

#endif  // MJB_PIPELINE_H_
