// Model-specialised sm_100a kernels of libmjb, compiled PER MODEL at mjb_makeData by NVRTC
// (csrc/mjb_jit.cc). The translation unit NVRTC sees is
//
//     #define MJB_SPECIALIZED 1
//     #define MJB_SPEC_NBODY <nbody> ...            (sizes needed as template arguments)
//     alignas(16) __device__ const unsigned long long kSpecBlob[] = { <the model blob> };
//     #include "mjb_spec_kernels.cuh"
//
// so the model blob (csrc/mjb_model.h) is a compile-time constant: with the loops over bodies /
// dofs / candidates expanded (MJB_BODY_LOOP_*, MJB_UNROLL in mjb_pipeline.h) every table look-up
// folds into an immediate or a constant-bank operand of the fp64 instruction that uses it, joint-type
// and topology branches disappear, and scratch rows are addressed with immediates. Measured on the
// humanoid: 2.6-2.9x fewer executed instructions than the generic kernels, 56 % of them fp64
// (generic: 21 %). Same per-state functions, same scratch layout, same results as the generic
// kernels of mjb_kernels.cu, which stay the path for models too large to expand and for hosts
// without NVRTC.
#ifndef MJB_SPEC_KERNELS_CUH_
#define MJB_SPEC_KERNELS_CUH_

#ifndef MJB_SPECIALIZED
#error "mjb_spec_kernels.cuh is compiled only through the model-specialising prelude (mjb_jit.cc)"
#endif

#include "mjb_launch.h"

namespace mjb {

__device__ __forceinline__ void spec_ctx(Ctx& c, const LaunchArgs& a) {
  const unsigned char* model = reinterpret_cast<const unsigned char*>(kSpecBlob);
  const mjbHdr* H = reinterpret_cast<const mjbHdr*>(model);
  c.H = H;
  c.I = reinterpret_cast<const int*>(model + H->int_section);
  c.D = reinterpret_cast<const double*>(model + H->num_section);
  c.N = a.stride;
  c.nconmax = a.nconmax;
  c.njmax = a.njmax;
  c.out = a.out;
  c.ncon = c.ne = c.nf = c.nl = c.nefc = c.status = 0;
  c.sm = nullptr; c.sc = nullptr; c.isc = nullptr;
  c.qpos = nullptr; c.qvel = nullptr; c.qacc = nullptr;
  c.s = 0;
  c.lci = nullptr; c.lcd = nullptr; c.lbody0 = 0; c.ldof0 = 0;
}

__device__ __forceinline__ void spec_bind(Ctx& c, const LaunchArgs& a, long long local) {
  const long long s = a.chunk_start + local;
  c.s = s;
  const long long blk = local >> 5, ln = local & 31;
  c.sc = a.scratch + (blk * a.nscratch << 5) + ln;
  c.isc = a.iscratch + (blk * a.niscratch << 5) + ln;
  c.qpos = a.qpos + s;
  c.qvel = a.qvel + s;
  c.qacc = a.qacc + s;
}

}  // namespace mjb

// Stages. The expanded forward sweep and the expanded inertia sweep are cut into body ranges of a few
// thousand instructions each (MJBS_SMOOTH_STAGES / MJBS_INERTIA_STAGES, chosen by the host from a
// per-body cost estimate), one kernel per range, so that each kernel's code stays resident in the
// instruction cache; stages hand over through the same scratch rows the generic kernels use.
// measured on the humanoid (GPU calls X, Y): resident CTAs 1 or 2: 2.75 ms, 3: 2.66 ms, 4: 2.62 ms
// (arm26, 3 bodies: 0.517 ms with 2, 0.548 ms with 4 -- small models keep 2)
#ifndef MJBS_SMOOTH_CTAS
#define MJBS_SMOOTH_CTAS (MJB_SPEC_NBODY >= 8 ? 4 : 2)
#endif
#ifndef MJBS_INERTIA_CTAS
#define MJBS_INERTIA_CTAS 3
#endif
#ifndef MJBS_BACKWARD_CTAS
#define MJBS_BACKWARD_CTAS 4
#endif
#ifndef MJBS_SCAN_CTAS
#define MJBS_SCAN_CTAS 4
#endif

#define MJBS_STATE_LOOP(i)                                                                        \
  for (long long i = (long long)blockIdx.x * mjb::kThreads + threadIdx.x; i < a.chunk_n;          \
       i += (long long)gridDim.x * mjb::kThreads)

// per-thread carry slots of the forward sweep in shared memory: stride = CTA size
#if MJB_SMS != 128
#error "the specialised kernels are compiled with MJB_SMOOTH_THREADS = 128 (mjb_jit.cu)"
#endif

#define MJBS_DEFINE_SMOOTH(idx, lo, hi)                                                            \
  extern "C" __global__ void __launch_bounds__(mjb::kThreads, MJBS_SMOOTH_CTAS)                    \
  mjbs_smooth_##idx(mjb::LaunchArgs a) {                                                           \
    __shared__ double carry_slots[MJB_SM_SLOTS * mjb::kThreads];                                   \
    mjb::Ctx c;                                                                                    \
    mjb::spec_ctx(c, a);                                                                           \
    c.sm = carry_slots + threadIdx.x;                                                              \
    MJBS_STATE_LOOP(i) {                                                                           \
      mjb::spec_bind(c, a, i);                                                                     \
      mjb::phase_smooth<MJB_SPEC_SPATIAL != 0, lo, hi>(c);                                         \
    }                                                                                              \
  }
MJBS_SMOOTH_STAGES(MJBS_DEFINE_SMOOTH)

#define MJBS_DEFINE_INERTIA(idx, lo, hi)                                                           \
  extern "C" __global__ void __launch_bounds__(mjb::kThreads, MJBS_INERTIA_CTAS)                   \
  mjbs_inertia_##idx(mjb::LaunchArgs a) {                                                          \
    mjb::Ctx c;                                                                                    \
    mjb::spec_ctx(c, a);                                                                           \
    MJBS_STATE_LOOP(i) {                                                                           \
      mjb::spec_bind(c, a, i);                                                                     \
      mjb::phase_inertia<lo, hi>(c);                                                               \
    }                                                                                              \
  }
MJBS_INERTIA_STAGES(MJBS_DEFINE_INERTIA)

// tree stages (trunk forward / fused subtree / trunk backward, mjb_pipeline.h phase_tree), used when
// the inertia outputs are requested; launched in the order of the list
#ifndef MJBS_TREE_CTAS
#define MJBS_TREE_CTAS 2
#endif
#define MJBS_DEFINE_TREE(idx, mode, lo, hi, first, tail, nbodies, dof0, ndofs)                     \
  extern "C" __global__ void __launch_bounds__(mjb::kThreads, MJBS_TREE_CTAS)                      \
  mjbs_tree_##idx(mjb::LaunchArgs a) {                                                             \
    __shared__ double carry_slots[MJB_SM_SLOTS * mjb::kThreads];                                   \
    mjb::Ctx c;                                                                                    \
    mjb::spec_ctx(c, a);                                                                           \
    c.sm = carry_slots + threadIdx.x;                                                              \
    MJBS_STATE_LOOP(i) {                                                                           \
      mjb::spec_bind(c, a, i);                                                                     \
      mjb::phase_tree<MJB_SPEC_SPATIAL != 0, mode, lo, hi, first, tail, nbodies, dof0, ndofs>(c);  \
    }                                                                                              \
  }
MJBS_TREE_STAGES(MJBS_DEFINE_TREE)

extern "C" __global__ void __launch_bounds__(mjb::kThreads, MJBS_SCAN_CTAS) mjbs_contact_scan(mjb::LaunchArgs a) {
  mjb::Ctx c;
  mjb::spec_ctx(c, a);
  const int nwords = (c.H->ncand + 31) >> 5;
  for (long long i0 = (long long)blockIdx.x * mjb::kThreads; i0 < a.chunk_n; i0 += (long long)gridDim.x * mjb::kThreads) {
    const long long i = i0 + threadIdx.x;
    const bool valid = i < a.chunk_n;
    mjb::spec_bind(c, a, valid ? i : 0);
    if (valid) mjb::contact_scan(c);
    if (a.cq) mjb::scan_append_items(a, c.isc, valid, i, nwords);   // item-parallel path: survivors -> item list
  }
}

extern "C" __global__ void __launch_bounds__(mjb::kThreads, MJBS_BACKWARD_CTAS) mjbs_backward(mjb::LaunchArgs a) {
  mjb::Ctx c;
  mjb::spec_ctx(c, a);
  MJBS_STATE_LOOP(i) {
    mjb::spec_bind(c, a, i);
    mjb::phase_backward<MJB_SPEC_PASSIVE_WRENCH != 0>(c);
  }
}

#endif  // MJB_SPEC_KERNELS_CUH_
