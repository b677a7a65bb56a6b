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

// The stages, in pipeline order. Each part is a fragment of this header (inside namespace mjb, using the
// accessor macros above); csrc/mjb_kernels.cu, the NVRTC translation unit of a specialised model and the
// test-only host build all see one and the same text.
#include "mjb_rows.h"
#include "mjb_sweep.h"
#include "mjb_contact.h"
#include "mjb_convex.h"
#include "mjb_narrow.h"
#include "mjb_backward.h"
#include "mjb_outputs.h"
#include "mjb_inertia.h"

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
  // generic CUDA kernels: fluid forces live in the instantiation for the rarer features (kSpatial), which
  // mjb_inverse launches for models with force-carrying spatial tendons or a medium
#if defined(__CUDACC__) && !defined(MJB_SPECIALIZED)
  constexpr bool kFluid = kSpatial;
#else
  constexpr bool kFluid = true;
#endif
  forward_sweep<kLo, kHi, true, kFluid>(c);   // incl. input checks, joint springs/dampers, dof friction and joint limit rows
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
    equality_rows<kSpatial>(c);      // rows [0, ne)
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
  if (c.out.sensordata) sensors<>(c);
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
