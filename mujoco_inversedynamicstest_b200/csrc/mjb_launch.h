// Launch arguments shared by the precompiled sm_100a kernels (mjb_kernels.cu), the model-specialised
// kernels compiled at run time (mjb_spec_kernels.cuh) and the C-ABI host code. Plain pointers and
// sizes only; no dependency on the CUDA runtime headers, so NVRTC can compile it as is.
#ifndef MJB_LAUNCH_H_
#define MJB_LAUNCH_H_

#include "mjb_model.h"
#include "mjb_pipeline.h"

namespace mjb {

// launch geometry: 128-thread CTAs (4 warps), grids capped at a multiple of the 148 SMs; every
// kernel walks its chunk of states with a block-stride loop
constexpr int kThreads = 128;
constexpr int kSMs = 148;
#ifndef MJB_LISTCAP
#define MJB_LISTCAP 24   // measured: 24 -> 5.16 ms, 16 -> 5.45, 32 -> 5.96 (2^20 humanoid states)
#endif
constexpr int kListCap = MJB_LISTCAP;                  // per-lane survivor list of the contact kernel

// Item-parallel contact phase: global lists of one chunk. Survivors of the bounding-sphere scan
// ("items") and the contacts they yield are appended with warp-aggregated atomics; a state finds its
// items through MJB_ISC_ITEMBASE / NSURV and an item its contacts through ItemCon, so placement
// order does not matter. If a list would overflow, `overflow` is raised and the chunk is handled
// by the pooled contact kernel instead (every kernel of either path checks the flag first).
// `overflow` is raised only by the scan kernels (item list full) and `overflow_contacts` only by
// contact_narrow_kernel (contact list full), so the flag a kernel tests on entry cannot change while
// that kernel runs: every thread of every CTA takes the same decision.
struct ContactQueue { int nitems; int ncontacts; int overflow; int nslots; int overflow_contacts; };
struct ContactItem { int state; int ci; };          // chunk-local state, candidate pair
// the item's contacts: contacts[base .. base+count); code0 / code1: rows of its first two contacts (> 0) or
// -1 - exclude (no rows), set by the narrow kernel so that the numbering pass need not read the contact
// records (further contacts of an item carry their code in ContactRec::efc_address)
struct ItemCon { int base; int count; int code0; int code1; };
// slot (state's first slot + k) -> contact record, contact index k in its state, first efc row (or -1 - exclude)
struct SlotRec { int rec; int k; int efc_address; int pad; };
struct ContactRec {
  int state, ci, k, efc_address;                    // efc_address: the contact's row code (see ItemCon); k unused
  double dist, pos[3], frame[6];                    // normal, tangent (third axis is their cross product)
};


struct LaunchArgs {
  const unsigned char* model;   // device blob (mjbHdr + sections)
  int model_bytes;              // the staged part of the blob (mjbHdr.staged_bytes)
  int sensor_cold;              // the sensor kernel reads tables behind the staged part (cam_project): model from global memory
  int model_in_smem;            // 1: stage the blob into shared memory with a TMA bulk copy
  const double* qpos;           // [nq][stride]
  const double* qvel;           // [nv][stride]
  const double* qacc;           // [nv][stride]
  double* qacc_discrete;        // [nv][stride] continuous-time qacc (mjENBL_INVDISCRETE), or null
  double* scratch;              // [chunk_stride/32][nscratch][32]  intermediates of one chunk of states
  int* iscratch;                // [chunk_stride/32][niscratch][32]
  int nscratch, niscratch;      // slots per state: mjbHdr::nscratch, MJB_ISC_MASK + ceil(ncand/32) + 1
  long long chunk_stride;       // states per chunk (multiple of 32)
  long long chunk_start;        // first state of the chunk
  int chunk_n;                  // states in the chunk
  long long stride;             // row stride of every state-indexed input/output array
  int nconmax, njmax;
  ContactQueue* cq;             // item-parallel contact path (null: pooled kernel only)
  ContactItem* items; ItemCon* item_con; ContactRec* contacts;
  SlotRec* slot_rec;            // slot (state's first slot + k) -> contact record and its numbering
  int items_cap, contacts_cap;
  int has_contacts;             // run the contact kernel (ncand > 0 and contacts enabled)
  int has_spatial;              // mjbHdr::has_spatial || has_fluid (force-carrying spatial tendons, fluid forces: smooth kernel variant)
  int has_gravcomp;             // mjbHdr::has_gravcomp (selects the backward kernel instantiation)
  int max_pair_contacts;        // mjbHdr::max_pair_contacts (sizes the per-warp contact pool)
  int simple_pairs;             // mjbHdr::simple_pairs (selects the narrow-phase kernel instantiation)
  int sensor_ccd;               // mjbHdr::sensor_ccd (the sensor kernel instantiation that carries GJK / EPA)
  int has_convex;               // mjbHdr::has_convex (the instantiations that carry GJK / EPA and its polytope)
  int skip_sensors;             // mj_inverseSkip(skipsensor = 1): leave sensordata as it is
  int inertia_subwarp;          // 1: mj_crb + mj_factorM by inertia_subwarp_kernel (8 lanes per state, on-chip intermediates)
  int sub_nv, sub_nbody, sub_nC; //    its sizes (mjbHdr::nv, nbody, nC)
  int scan_wide;                // > 0: warp-per-state candidate scan with this many states per CTA
                                //      (scenes with long candidate lists, mjb_kernels.cu)
  int* cmask;                   // survivor masks of the warp-per-state scans, [chunk_stride][ceil(ncand/32)], one state's
                                //   words contiguous (the thread-per-state scan keeps them interleaved in iscratch), or null
  const int* pair_ci;           // [ngeom][ngeom] candidate index of a geom pair, -1: none (contact_scan_pairs_kernel), or null
  int* scan_buf;                // per-warp candidate buffers of the wide scan (scan_wide_buf_ints), or null
  int scan_buf_cap;             //   ints per warp
  int scan_ngeom;               // mjbHdr::ngeom (sizes the wide scan's shared memory)
  Outputs out;
};

#if defined(__CUDACC__)
// Tail of the candidate scan (thread per state): the warp appends the bounding-sphere survivors of its
// 32 states to the chunk's global item list with ONE atomicAdd; every state gets a contiguous range
// in candidate order (MJB_ISC_ITEMBASE, -1 if the list is full). All 32 lanes must call it.
__device__ __forceinline__ void scan_append_items(const LaunchArgs& a, int* isc, bool valid, long long local_state,
                                                  int nwords) {
  const int lane = threadIdx.x & 31;
  const int nsurv = valid ? isc[MJB_ISC_NSURV * MJB_LS] : 0;
  int incl = nsurv;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (lane == 0 && total) base = atomicAdd(&a.cq->nitems, total);
  base = __shfl_sync(0xffffffffu, base, 0);
  const bool fits = base + total <= a.items_cap;
  if (!fits && lane == 0) a.cq->overflow = 1;
  const int mybase = fits ? base + incl - nsurv : -1;
  if (valid) isc[MJB_ISC_ITEMBASE * MJB_LS] = mybase;
  if (valid && fits && nsurv) {
    int k = 0;
    for (int w = 0; w < nwords; w++) {
      unsigned bits = (unsigned)isc[(size_t)(MJB_ISC_MASK + w) * MJB_LS];
      while (bits) {
        const int b = __ffs((int)bits) - 1;
        bits &= bits - 1;
        a.items[mybase + k] = ContactItem{(int)local_state, (w << 5) + b};
        k++;
      }
    }
  }
}
#endif

}  // namespace mjb

#endif  // MJB_LAUNCH_H_
