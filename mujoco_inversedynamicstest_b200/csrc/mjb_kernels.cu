// sm_100a kernels of libmjb: the mj_inverse phase kernels (one thread per state), the AoS<->SoA
// transposes at the host boundary, and an FP64 FMA peak probe for the roofline denominator.
//
// Thread mapping. The per-state matrices of mj_inverse are tiny, sparse and tree-structured
// (nv = 27 for the humanoid): there is no contraction to feed tensor cores, and a warp-per-state
// mapping would leave most fp64 lanes idle during the depth-serial tree sweeps (1-4 bodies per
// level). So each THREAD owns one state and all model-driven control flow (topology, joint types,
// candidate geom pairs) is uniform across the warp; only contact/limit activity diverges.
// The pipeline is cut into four phase kernels so that each gets its own register budget and
// occupancy; intermediates are handed over through a per-state scratch laid out [slot][chunk] in
// HBM (consecutive lanes -> consecutive doubles, fully coalesced 256-byte warp accesses). The model
// blob is staged once per CTA into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier)
// and read through warp-uniform shared-memory broadcasts.
#include "mjb_kernels.cuh"
#include "mjb_jit.h"

#include <cstdio>
#include <vector>
#include <cstdlib>

// resident CTAs per SM each phase kernel is compiled for (register budget = 65536 / (threads * CTAS)).
// Measured on B200 (profiles/README.md): the smooth kernel is fastest WITHOUT register spills at low
// occupancy -- one 256-thread CTA per SM, 246 registers -- because spilled values miss the L1 that the
// streamed scratch rows keep flushing (2.99 ms vs 4.0 ms at 128 registers for 2^20 humanoid states).
#ifndef MJB_CTAS_SMOOTH
#define MJB_CTAS_SMOOTH 1
#endif
#ifndef MJB_CTAS_INERTIA
#define MJB_CTAS_INERTIA 3
#endif
#ifndef MJB_CTAS_BACKWARD
#define MJB_CTAS_BACKWARD 4
#endif
#ifndef MJB_ROWS_CTAS
#define MJB_ROWS_CTAS 4
#endif
#ifndef MJB_CTAS_CONTACT
#define MJB_CTAS_CONTACT 2
#endif

namespace mjb {

// ------------------------------------------------------------------------------------------
// TMA bulk copy of the model blob global -> shared (SASS: UBLKCP), completion on an mbarrier

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void stage_model_tma(unsigned char* smem_dst, const unsigned char* gsrc,
                                                uint32_t bytes, uint64_t* mbar) {
  const uint32_t bar = smem_u32(mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
    // one bulk copy moves at most what the tx-count can express; split large blobs in 64 KB pieces
    uint32_t done = 0;
    while (done < bytes) {
      const uint32_t chunk = (bytes - done) < 65536u ? (bytes - done) : 65536u;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"(smem_u32(smem_dst + done)), "l"(gsrc + done), "r"(chunk), "r"(bar)
          : "memory");
      done += chunk;
    }
  }
  // every thread waits for phase 0 of the barrier
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(0)
      : "memory");
}

// ------------------------------------------------------------------------------------------
// phase kernels (see mjb_pipeline.h: phase_smooth / phase_inertia / phase_contact / phase_backward)

struct Prologue {
  const unsigned char* model;
};

template <bool kModelInSmem>
__device__ __forceinline__ void make_ctx(Ctx& c, const LaunchArgs& a, unsigned char* smem,
                                         uint64_t* mbar) {
  const unsigned char* model = a.model;
  if (kModelInSmem) {
    stage_model_tma(smem, a.model, static_cast<uint32_t>(a.model_bytes), mbar);
    model = smem;
  }
  const mjbHdr* H = reinterpret_cast<const mjbHdr*>(model);
  c.H = H;
  c.I = reinterpret_cast<const int*>(model + H->int_section);
  c.D = reinterpret_cast<const double*>(model + H->num_section);
  c.N = a.stride;
  c.nconmax = a.nconmax;
  c.njmax = a.njmax;
  c.out = a.out;
  c.ncon = c.ne = c.nf = c.nl = c.nefc = c.status = 0;
  c.sm = nullptr;
  // every field is defined before a Ctx is copied: copying a context with indeterminate members is
  // undefined behaviour and did produce a corrupted table pointer in one kernel instantiation
  c.sc = nullptr; c.isc = nullptr;
  c.qpos = nullptr; c.qvel = nullptr; c.qacc = nullptr;
  c.s = 0;
  c.lci = nullptr; c.lcd = nullptr; c.lbody0 = 0; c.ldof0 = 0;
}

__device__ __forceinline__ void bind_state(Ctx& c, const LaunchArgs& a, long long local) {
  const long long s = a.chunk_start + local;
  c.s = s;
  // warp-blocked scratch: block (local/32) holds every slot of 32 consecutive states
  const long long blk = local >> 5, ln = local & 31;
  c.sc = a.scratch + (blk * a.nscratch << 5) + ln;
  c.isc = a.iscratch + (blk * a.niscratch << 5) + ln;
  c.qpos = a.qpos + s;
  c.qvel = a.qvel + s;
  c.qacc = a.qacc + s;
}

// The smooth kernel runs with its own CTA size (MJB_SMOOTH_THREADS, = the stride MJB_SMS of the
// per-thread carry slots): fewer, larger CTAs per SM mean fewer copies of the model blob in shared
// memory and therefore more L1 for the streamed scratch rows.
constexpr int kSmoothThreads = MJB_SMS;

template <bool kModelInSmem, bool kSpatial = false>
__global__ void __launch_bounds__(kSmoothThreads, MJB_CTAS_SMOOTH) smooth_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  // per-thread carry slots of the forward sweep, after the model blob
  c.sm = reinterpret_cast<double*>(smem + (kModelInSmem ? ((a.model_bytes + 127) & ~127) : 0)) + threadIdx.x;
  for (long long i = (long long)blockIdx.x * kSmoothThreads + threadIdx.x; i < a.chunk_n;
       i += (long long)gridDim.x * kSmoothThreads) {
    bind_state(c, a, i);
    phase_smooth<kSpatial>(c);
  }
}

template <bool kModelInSmem>
__global__ void __launch_bounds__(kThreads, MJB_CTAS_INERTIA) inertia_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.chunk_n;
       i += (long long)gridDim.x * kThreads) {
    bind_state(c, a, i);
    phase_inertia(c);
  }
}

// ------------------------------------------------------------------------------------------
// mj_crb + mj_factorM with a SUB-WARP of 8 lanes per state and every intermediate on chip (the
// mapping BASELINE.json's north_star names: engine_core_smooth.c:1353-1401, :1470-1511).
//
// One CTA owns one warp block of the scratch (32 consecutive states) at a time: 256 threads = 32
// states x 8 workers, the 8 workers of a state in one warp (4 states per warp). Per state the
// kernel keeps cdof (6 nv), the composite inertias (10 nbody) and the sparse matrix (nC entries,
// the reference's C layout: row i = ancestors ascending, self last) in shared memory, state-major
// with odd strides so that the coalesced transfers (lane = state) and the worker accesses (8
// entries of 4 states) spread over the banks:
//   load   : cdof / cinert rows of the block, 256-byte lines, straight from the scratch
//   crb    : every worker owns components of the 10-number inertias and walks the bodies leaves to
//            root on its own (no exchange between workers)
//   M      : worker g takes dofs g, g+8, ...: buf = crb[body] cdof_i, then one dot product per
//            ancestor (the walk of mj_crb), written into the matrix
//   qM out : the finished entries leave as whole 256-byte rows (legacy layout through mapM2C)
//   L'DL   : mj_factorI's elimination in place, rows k = nv-1 .. 0 in lock step (__syncwarp), the
//            ancestor rows of k dealt out to the 8 workers in pairs (p, na-1-p) of equal work
//   out    : qLD / qLDiagInv as whole rows
// HBM sees the inputs once (cinert + cdof) and the outputs once; nothing is re-read and there are
// no accumulator pushes. Selected per model when the three arrays fit in shared memory.
constexpr int kSubG = 8;            // workers per state
constexpr int kSubS = 32;           // states per CTA (one warp block of the scratch)

struct SubwarpLayout { int cd, cr, ws, tab_ints; size_t bytes; };
SubwarpLayout subwarp_layout(int nv, int nbody, int nC) {
  SubwarpLayout L;
  L.cd = (6 * nv) | 1; L.cr = (10 * nbody) | 1; L.ws = nC | 1;
  L.tab_ints = ((nbody + 7 * nv + 3 * nC) + 1) & ~1;          // index tables, kept 8-byte aligned
  L.bytes = (size_t)kSubS * (size_t)(L.cd + L.cr + L.ws) * sizeof(double) + (size_t)2 * nv * sizeof(double) +
            (size_t)L.tab_ints * sizeof(int);
  return L;
}
bool inertia_subwarp_fits(int nv, int nbody, int nC) {
  return nv > 0 && subwarp_layout(nv, nbody, nC).bytes <= 220 * 1024;
}

__global__ void __launch_bounds__(kSubS * kSubG, 1) inertia_subwarp_kernel(LaunchArgs a) {
  extern __shared__ __align__(16) double sw[];
  const mjbHdr* H = reinterpret_cast<const mjbHdr*>(a.model);
  const int* I = reinterpret_cast<const int*>(a.model + H->int_section);
  const double* D = reinterpret_cast<const double*>(a.model + H->num_section);
  const int nv = H->nv, nbody = H->nbody, nC = H->nC;
  const int CD = (6 * nv) | 1, CR = (10 * nbody) | 1, WS = nC | 1;
  double* cd_all = sw;
  double* cr_all = cd_all + (size_t)kSubS * CD;
  double* ws_all = cr_all + (size_t)kSubS * CR;
  // the model's index tables, once per CTA, in shared memory: every look-up of the sweeps below is an
  // LDS with at most 8 distinct addresses per warp instead of a dependent global load
  double* armature = ws_all + (size_t)kSubS * WS;             // [nv]
  double* dof_M0 = armature + nv;                              // [nv]
  int* body_parentid = reinterpret_cast<int*>(dof_M0 + nv);    // [nbody]
  int* dof_bodyid = body_parentid + nbody;                     // [nv] ...
  int* dof_parentid = dof_bodyid + nv;
  int* dof_Madr = dof_parentid + nv;
  int* dof_simplenum = dof_Madr + nv;
  int* rownnz = dof_simplenum + nv;
  int* rowadr = rownnz + nv;
  int* diagadr = rowadr + nv;                                  // rowadr + rownnz - 1
  int* colind = diagadr + nv;                                  // [nC]
  int* mapM2C = colind + nC;                                   // [nC]
  int* ancrow = mapM2C + nC;                                   // [nC]: rowadr[colind[e]], the row a C entry's column owns
  const int tid = threadIdx.x;
  {
    const int* g_rownnz = I + H->ioff[MJB_I_C_rownnz];
    const int* g_rowadr = I + H->ioff[MJB_I_C_rowadr];
    const int* g_colind = I + H->ioff[MJB_I_C_colind];
    for (int i = tid; i < nbody; i += blockDim.x) body_parentid[i] = (I + H->ioff[MJB_I_body_parentid])[i];
    for (int i = tid; i < nv; i += blockDim.x) {
      dof_bodyid[i] = (I + H->ioff[MJB_I_dof_bodyid])[i];
      dof_parentid[i] = (I + H->ioff[MJB_I_dof_parentid])[i];
      dof_Madr[i] = (I + H->ioff[MJB_I_dof_Madr])[i];
      dof_simplenum[i] = (I + H->ioff[MJB_I_dof_simplenum])[i];
      rownnz[i] = g_rownnz[i]; rowadr[i] = g_rowadr[i]; diagadr[i] = g_rowadr[i] + g_rownnz[i] - 1;
      armature[i] = (D + H->noff[MJB_N_dof_armature])[i];
      dof_M0[i] = (D + H->noff[MJB_N_dof_M0])[i];
    }
    for (int e = tid; e < nC; e += blockDim.x) {
      colind[e] = g_colind[e];
      mapM2C[e] = (I + H->ioff[MJB_I_mapM2C])[e];
      ancrow[e] = g_rowadr[g_colind[e]];
    }
  }
  const int s = tid / kSubG, g = tid % kSubG;          // compute phases: state, worker
  const int sl = tid & 31, rl = tid >> 5;              // transfer phases: state = lane, row slice
  constexpr int kRows = kSubS * kSubG / 32;            // rows moved per pass
  double* cd = cd_all + (size_t)s * CD;
  double* cr = cr_all + (size_t)s * CR;
  double* ws = ws_all + (size_t)s * WS;
  double* cd_t = cd_all + (size_t)sl * CD;             // transfer views: this lane's state
  double* cr_t = cr_all + (size_t)sl * CR;
  double* ws_t = ws_all + (size_t)sl * WS;
  const size_t off_cdof = (size_t)H->scoff[MJB_SC_cdof], off_cinert = (size_t)H->scoff[MJB_SC_cinert];
  const size_t N = (size_t)a.stride;
  const long long nblocks = (a.chunk_n + 31) >> 5;
  const int ncd = 6 * nv, ncr = 10 * nbody;

  for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const double* sc = a.scratch + (blk * a.nscratch << 5);
    const long long sg = a.chunk_start + (blk << 5) + sl;          // global state of the transfer lane
    const bool live = (blk << 5) + sl < a.chunk_n;
    // ---- load
    {
      const double* src = sc + (off_cdof << 5) + sl;
      for (int r = rl; r < ncd; r += kRows) cd_t[r] = __ldcs(src + ((size_t)r << 5));
      src = sc + (off_cinert << 5) + sl;
      for (int r = rl; r < ncr; r += kRows) cr_t[r] = __ldcs(src + ((size_t)r << 5));
    }
    __syncthreads();
    // ---- crb: components g and g + 8 of every body, leaves to root
    for (int comp = g; comp < 10; comp += kSubG) {
      for (int b = nbody - 1; b > 0; b--) {
        const int p = body_parentid[b];
        if (p > 0) cr[10 * p + comp] += cr[10 * b + comp];
      }
    }
    __syncwarp();
    // ---- M: the entries of rows g, g + 8, ...
    for (int i = g; i < nv; i += kSubG) {
      const int diag = diagadr[i];
      if (dof_simplenum[i]) {
        // simple body: M is diagonal and constant (engine_core_smooth.c:1375-1385); the legacy qM
        // keeps the zero ancestor entries, which the reduced rows do not hold
        ws[diag] = dof_M0[i];
        const long long st = a.chunk_start + (blk << 5) + s;
        if ((blk << 5) + s < a.chunk_n) {
          int t = 1;
          for (int j = dof_parentid[i]; j >= 0; j = dof_parentid[j], t++) a.out.qM[(size_t)(dof_Madr[i] + t) * N + st] = 0;
        }
        continue;
      }
      double S[6], inert[10], buf[6];
      const double* ci = cd + 6 * i;
#pragma unroll
      for (int k = 0; k < 6; k++) S[k] = ci[k];
      const double* cb = cr + 10 * dof_bodyid[i];
#pragma unroll
      for (int k = 0; k < 10; k++) inert[k] = cb[k];
      mulInertVecF(buf, inert, S);
      ws[diag] = armature[i] + dot6f(S, buf);
      double* we = ws + diag - 1;
      for (int j = dof_parentid[i]; j >= 0; j = dof_parentid[j], we--) {
        const double* cj = cd + 6 * j;
        double Sj[6];
#pragma unroll
        for (int k = 0; k < 6; k++) Sj[k] = cj[k];
        *we = dot6f(Sj, buf);
      }
    }
    __syncthreads();
    // ---- qM out (legacy layout), whole rows
    if (live) {
      double* dst = a.out.qM + sg;
      for (int k = rl; k < nC; k += kRows) __stcs(dst + (size_t)mapM2C[k] * N, ws_t[k]);
    }
    __syncthreads();
    // ---- L'DL in place (mj_factorI): 1/D of row k lands in the dead cdof row of the state
    for (int k = nv - 1; k >= 0; k--) {
      const int na = rownnz[k] - 1;                         // ancestors of k
      double* rowk = ws + rowadr[k];
      const int* anck = ancrow + rowadr[k];
      const double invD = 1 / rowk[na];
      if (g == 0) cd[k] = invD;
      if (dof_simplenum[k] || na == 0) continue;
      // ancestor rows p (length p + 1) in pairs (p, na - 1 - p): equal work per worker
      double tmp[2];
      int pp[2];
      pp[0] = g < na ? g : -1;
      pp[1] = (na - 1 - g >= kSubG) ? na - 1 - g : -1;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        tmp[q] = 0;
        if (pp[q] < 0) continue;
        const int p = pp[q];
        tmp[q] = rowk[p] * invD;
        double* ri = ws + anck[p];
        const double scl = -tmp[q];
#pragma unroll 4
        for (int t = 0; t <= p; t++) ri[t] = fma(scl, rowk[t], ri[t]);
      }
      for (int p = kSubG + g; p < na - kSubG; p += kSubG) {      // deep chains only (na > 16)
        const double tp = rowk[p] * invD;
        double* ri = ws + anck[p];
        for (int t = 0; t <= p; t++) ri[t] = fma(-tp, rowk[t], ri[t]);
      }
      __syncwarp();
      // row k itself is scaled after every worker has read its unscaled entries
#pragma unroll
      for (int q = 0; q < 2; q++) if (pp[q] >= 0) rowk[pp[q]] = tmp[q];
      for (int p = kSubG + g; p < na - kSubG; p += kSubG) rowk[p] *= invD;
      __syncwarp();
    }
    __syncthreads();
    // ---- qLD, qLDiagInv out
    if (live) {
      double* dst = a.out.qLD + sg;
      for (int k = rl; k < nC; k += kRows) __stcs(dst + (size_t)k * N, ws_t[k]);
      dst = a.out.qLDiagInv + sg;
      for (int k = rl; k < nv; k += kRows) __stcs(dst + (size_t)k * N, cd_t[k]);
    }
    __syncthreads();
  }
}

// mjENBL_INVDISCRETE: convert the discrete-time qacc with the factors the inertia kernel just wrote
template <bool kModelInSmem>
__global__ void __launch_bounds__(kThreads, 4) discrete_acc_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.chunk_n;
       i += (long long)gridDim.x * kThreads) {
    bind_state(c, a, i);
    discrete_acc(c, a.qacc_discrete + c.s);
  }
}

// contact scan: bounding-sphere survivors of every state as a bit mask + count
template <bool kModelInSmem>
__global__ void __launch_bounds__(kThreads, 4) contact_scan_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  const int nwords = (c.H->ncand + 31) >> 5;
  for (long long i0 = (long long)blockIdx.x * kThreads; i0 < a.chunk_n; i0 += (long long)gridDim.x * kThreads) {
    const long long i = i0 + threadIdx.x;
    const bool valid = i < a.chunk_n;
    bind_state(c, a, valid ? i : 0);
    if (valid) contact_scan(c);
    if (a.cq) scan_append_items(a, c.isc, valid, i, nwords);     // item-parallel path: survivors -> item list
  }
}

// ------------------------------------------------------------------------------------------
// Warp-per-state candidate scan for scenes with long candidate lists (BASELINE config 5: 22
// humanoids, ~90 K candidate pairs, ~1,200 survivors per state).
//
// With one thread per state every lane walks the whole candidate list and gathers two geom
// positions per candidate from its own scratch column: 2^15 states x 90 K candidates x 48 bytes of
// L1/L2 sector traffic (49.5 ms, 69 % of that step). Here ONE WARP owns one state and its 32 lanes
// test 32 consecutive candidates per step, in two stages:
//   1. a CONSERVATIVE single-precision filter over all candidates. The state's geom positions are
//      staged once in shared memory as floats relative to the state's first geom (structure of
//      arrays: lanes testing consecutive geoms read consecutive words), the compact candidate rows
//      (geom ids, filter kind: 8 bytes, mjb_upload.cc) and the bounds rounded UP to float are
//      streamed through shared memory in tiles shared by the CTA's warps. A candidate is dropped
//      only if its float distance exceeds the float bound by more than `slack`, which covers the
//      rounding of the float path many times over (|coordinate| <= M after centring: error of the
//      distance <= ~7e-7 M, slack = max(1e-4, 8e-6 M)); everything else ("maybe": the true
//      survivors plus a 0.1 mm band, plus all plane candidates) is buffered in candidate order;
//   2. the EXACT test -- mj_filterSphere's double-precision arithmetic, identical to contact_scan
//      (mjb_pipeline.h) -- over the buffered candidates only, 32 at a time with all lanes busy,
//      compacting the buffer in place.
// The state's survivors are then appended to the chunk's global item list with one atomicAdd,
// which makes contact_items_kernel unnecessary on this path. The fp64 pipe (64 lanes per SM and
// clock) sees ~1.5 % of the candidates instead of all of them. The ballot of stage 1 is kept as
// the survivor mask word of those 32 candidates for the pooled fallback, which applies the exact
// tests itself; stage 2 clears the bits of the candidates it rejects.
constexpr int kWideTile = 2048;     // candidate rows per shared-memory tile
constexpr int kWideCtas = 2;        // resident CTAs per SM

size_t scan_wide_smem_bytes(int ngeom, int states_per_cta) {
  const size_t gp = (size_t)((ngeom + 3) & ~3);
  return (size_t)states_per_cta * gp * sizeof(float4) + (size_t)kWideTile * (sizeof(float) + 2 * sizeof(int));
}

// states per CTA (= warps) of the wide scan for this model, 0 if the flat thread-per-state scan is used
int scan_wide_states(int ncand, int ngeom) {
  if (ncand < 4096) return 0;
  for (int w = 16; w >= 2; w--) if (scan_wide_smem_bytes(ngeom, w) <= (size_t)(220 * 1024) / kWideCtas) return w;
  return 0;
}
// per-warp buffer (ints, global memory) that carries the stage-1 candidates to stage 2, and how many
// warps the launch has: the grid is fixed (kWideCtas CTAs per SM)
int scan_wide_buf_cap(int ncand) { return ncand < 16384 ? ((ncand + 31) & ~31) : 16384; }
long long scan_wide_buf_ints(int ncand, int ngeom) {
  return (long long)kSMs * kWideCtas * scan_wide_states(ncand, ngeom) * scan_wide_buf_cap(ncand);
}

// mj_filterSphere (engine_collision_driver.c:146-163) on candidate ci of the state whose scratch
// column is sc, in double precision (same expressions as contact_scan, mjb_pipeline.h)
__device__ __forceinline__ bool scan_exact_test(const double* sc, size_t off_gxpos, size_t off_gz,
                                                const int* scan_int, const double* scan_bound, int ci) {
  const int g1k = scan_int[2 * ci], g2 = scan_int[2 * ci + 1];
  const int g1 = g1k & 0xfffffff, planeflag = (int)((unsigned)g1k >> 28);
  if (planeflag > 1) return true;
  const double bound = scan_bound[ci];
  // geom positions / z axes are 4-double vectors per (geom, state) (mjb_pipeline.h geom_vec): sc is
  // the state's column (block base + lane), the vector of geom g sits at block + (off + 4 g) * 32 + 4 * lane
  const size_t ln = ((size_t)sc >> 3) & 31;
  const double* blk = sc - ln;
  double p1[4], p2[4];
  ld_rec4(p1, blk + ((off_gxpos + 4 * (size_t)g1) << 5) + 4 * ln);
  ld_rec4(p2, blk + ((off_gxpos + 4 * (size_t)g2) << 5) + 4 * ln);
  if (planeflag == 0) {
    const double dif[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    return !(dif[0]*dif[0] + dif[1]*dif[1] + dif[2]*dif[2] > bound*bound);
  }
  double nrm[4];
  ld_rec4(nrm, blk + ((off_gz + 4 * (size_t)g1) << 5) + 4 * ln);
  const double dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  return !(dot3(dif, nrm) > bound);
}

__global__ void __launch_bounds__(512, kWideCtas) contact_scan_wide_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const mjbHdr* H = reinterpret_cast<const mjbHdr*>(a.model);
  const int* I = reinterpret_cast<const int*>(a.model + H->int_section);
  const double* D = reinterpret_cast<const double*>(a.model + H->num_section);
  const int ngeom = H->ngeom, ncand = H->ncand;
  const int* scan_int = I + H->ioff[MJB_I_scan_int];
  const double* scan_bound = D + H->noff[MJB_N_scan_bound];
  const int* geom_store = I + H->ioff[MJB_I_geom_store];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int gp = (ngeom + 3) & ~3;
  float4* gx_all = reinterpret_cast<float4*>(smem);                         // [W][gp]: x, y, z of a geom in one 16-byte word
  float* tile_bound = reinterpret_cast<float*>(gx_all + (size_t)W * gp);    // [kWideTile]
  int2* tile_int = reinterpret_cast<int2*>(tile_bound + kWideTile);         // [kWideTile]
  float4* gx = gx_all + (size_t)warp * gp;
  const int cap = a.scan_buf_cap;
  int* buf = a.scan_buf + ((size_t)blockIdx.x * W + warp) * cap;            // this warp's candidate buffer
  const size_t off_gxpos = (size_t)H->scoff[MJB_SC_geom_xpos], off_gxmat = (size_t)H->scoff[MJB_SC_geom_zaxis];
  const int mask_row = MJB_ISC_MASK;
  const unsigned below = (1u << lane) - 1u;

  for (long long s0 = (long long)blockIdx.x * W; s0 < a.chunk_n; s0 += (long long)gridDim.x * W) {
    const long long s = s0 + warp;
    const bool valid = s < a.chunk_n;
    const long long sb = valid ? s : 0;
    const double* sc = a.scratch + ((sb >> 5) * a.nscratch << 5) + (sb & 31);
    int* isc = a.iscratch + ((sb >> 5) * a.niscratch << 5) + (sb & 31);
    int* maskw = a.cmask + (size_t)sb * ((ncand + 31) >> 5);    // this state's survivor mask, contiguous words
    __syncthreads();                      // the previous round's tiles and positions are consumed
    float slack = 1e-4f;
    if (valid) {
      // positions relative to a stored geom (geom 1 of the first candidate), as floats
      const int gref = scan_int[0] & 0xfffffff;
      const double* gblk = sc - (sb & 31) + 4 * (sb & 31);     // this state's 4-double geom vectors: + (off + 4 g) * 32
      const double r0 = gblk[(off_gxpos + 4 * gref) << 5], r1 = gblk[((off_gxpos + 4 * gref) << 5) + 1],
                   r2 = gblk[((off_gxpos + 4 * gref) << 5) + 2];
      float m = 0;
      for (int g = lane; g < ngeom; g += 32) {
        // geoms outside every candidate pair are not stored by the sweep (geom_store, mjb_upload.cc)
        const bool stored = (geom_store[g] & 3) != 0;
        const double* gv = gblk + ((off_gxpos + 4 * (size_t)g) << 5);
        const float x = stored ? (float)(gv[0] - r0) : 0.f;
        const float y = stored ? (float)(gv[1] - r1) : 0.f;
        const float z = stored ? (float)(gv[2] - r2) : 0.f;
        gx[g] = make_float4(x, y, z, 0.f);
        m = fmaxf(m, fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      slack = fmaxf(1e-4f, 8e-6f * m);
      if (!(m < 1e30f)) slack = 3e38f;    // non-finite positions: everything goes to the exact test
    }
    // stage 1: conservative float filter over all candidates. Plane and unfiltered candidates carry an
    // infinite float bound and the padding of the last tile a NaN (tile staging), so the loop is ONE
    // comparison per candidate: no branch on the kind, no bounds check. A state with non-finite
    // positions skips it and takes the exact path below.
    int count = 0;
    const bool finite = slack < 1e38f;
    for (int t0 = 0; t0 < ncand; t0 += kWideTile) {
      const int nt = ncand - t0 < kWideTile ? ncand - t0 : kWideTile;
      const int ntp = (nt + 31) & ~31;
      __syncthreads();
      for (int i = threadIdx.x; i < ntp; i += blockDim.x) {
        if (i < nt) {
          const int g1k = scan_int[2 * (t0 + i)];
          tile_bound[i] = ((unsigned)g1k >> 28) ? __int_as_float(0x7f800000) : __double2float_ru(scan_bound[t0 + i]);
          tile_int[i] = make_int2(g1k & 0xfffffff, scan_int[2 * (t0 + i) + 1]);
        } else {
          tile_bound[i] = __int_as_float(0x7fc00000);
          tile_int[i] = make_int2(0, 0);
        }
      }
      __syncthreads();
      if (!valid || !finite) continue;
      int* mw = maskw + (t0 >> 5);
      const int2* ti = tile_int + lane;
      const float* tb = tile_bound + lane;
#pragma unroll 4
      for (int c0 = 0; c0 < ntp; c0 += 32) {
        const int2 gi = ti[c0];
        const float t = tb[c0] + slack;
        const float4 p1 = gx[gi.x], p2 = gx[gi.y];
        const float dx = p1.x - p2.x, dy = p1.y - p2.y, dz = p1.z - p2.z;
        const bool maybe = fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= t * t;
        const unsigned m = __ballot_sync(0xffffffffu, maybe);
        if (lane == 0) mw[c0 >> 5] = (int)m;
        if (maybe) {
          const int k = count + __popc(m & below);
          if (k < cap) buf[k] = t0 + c0 + lane;
        }
        count += __popc(m);
      }
    }
    if (valid && !finite) count = cap + 1;      // everything through the exact path
    if (!valid) continue;
    __syncwarp();
    int total = 0, base = 0;
    bool fits = true;
    if (count <= cap) {
      // stage 2: exact test of the buffered candidates, compacted in place (writes trail the reads)
      for (int k0 = 0; k0 < count; k0 += 32) {
        const int k = k0 + lane;
        const int ci = k < count ? buf[k] : -1;
        const bool ok = ci >= 0 && scan_exact_test(sc, off_gxpos, off_gxmat, scan_int, scan_bound, ci);
        __syncwarp();
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) buf[total + __popc(m & ((1u << lane) - 1u))] = ci;
        else if (ci >= 0) atomicAnd(&maskw[ci >> 5], ~(1 << (ci & 31)));
        total += __popc(m);
        __syncwarp();
      }
      if (lane == 0) isc[(size_t)MJB_ISC_NSURV * MJB_LS] = total;
      if (!a.cq) continue;                 // pooled path only: the masks are all it needs
      if (lane == 0 && total) base = atomicAdd(&a.cq->nitems, total);
      base = __shfl_sync(0xffffffffu, base, 0);
      fits = base + total <= a.items_cap;
      if (!fits && lane == 0) a.cq->overflow = 1;
      if (lane == 0) isc[(size_t)MJB_ISC_ITEMBASE * MJB_LS] = fits ? base : -1;
      if (fits) for (int k = lane; k < total; k += 32) a.items[base + k] = ContactItem{(int)s, buf[k]};
    } else {
      // more candidates than the buffer holds (a very dense state): exact test of the whole list
      // straight from the tables, first to count and fix the masks, then to write the items
      for (int pass = 0; pass < 2; pass++) {
        int n = 0;
        for (int c0 = 0; c0 < ncand; c0 += 32) {
          const int ci = c0 + lane;
          const bool ok = ci < ncand && scan_exact_test(sc, off_gxpos, off_gxmat, scan_int, scan_bound, ci);
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (pass == 0) { if (lane == 0) maskw[c0 >> 5] = (int)m; }
          else if (ok && fits) a.items[base + n + __popc(m & ((1u << lane) - 1u))] = ContactItem{(int)s, ci};
          n += __popc(m);
        }
        if (pass == 1) break;
        total = n;
        if (lane == 0) isc[(size_t)MJB_ISC_NSURV * MJB_LS] = total;
        if (!a.cq) break;
        if (lane == 0 && total) base = atomicAdd(&a.cq->nitems, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        fits = base + total <= a.items_cap;
        if (!fits && lane == 0) a.cq->overflow = 1;
        if (lane == 0) isc[(size_t)MJB_ISC_ITEMBASE * MJB_LS] = fits ? base : -1;
        if (!fits) break;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// The same scan organised over GEOM PAIRS instead of the candidate list. The list-driven filter
// above is bound by shared memory: two gathered 16-byte position reads per candidate whose geom
// ids are scattered over the banks (ncu: 40 % issue-active with 44 % fewer instructions than the
// version before it, same 7.25 ms). In a scene where nearly every geom pair is a candidate
// (22 humanoids: 86,889 candidates of 87,571 pairs) the float filter needs no list at all:
//   1. every lane keeps ONE geom (position, bounding radius + largest margin + slack) in registers
//      and the warp walks the other geoms, one broadcast shared-memory read per 32 pair tests;
//      pairs that may touch (the survivors, a 0.1 mm band, everything involving a plane) are buffered
//      as (g1, g2);
//   2. the buffered pairs are looked up in the pair -> candidate table (mjb_makeData), pairs that
//      are no candidates drop out, the others take the EXACT test of their candidate row (same
//      function as above) and set their bit in the state's survivor mask;
//   3. the mask words, walked in order, yield the state's items in candidate order.
// The float filter is conservative for every candidate: its bound rbound1 + rbound2 + (largest
// margin of any candidate) + slack is not smaller than the candidate's own; geoms without a
// bounding radius (planes) are always kept, geoms outside every candidate pair never.
__global__ void __launch_bounds__(512, kWideCtas) contact_scan_pairs_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const mjbHdr* H = reinterpret_cast<const mjbHdr*>(a.model);
  const int* I = reinterpret_cast<const int*>(a.model + H->int_section);
  const double* D = reinterpret_cast<const double*>(a.model + H->num_section);
  const int ngeom = H->ngeom, ncand = H->ncand;
  const int* scan_int = I + H->ioff[MJB_I_scan_int];
  const double* scan_bound = D + H->noff[MJB_N_scan_bound];
  const int* geom_store = I + H->ioff[MJB_I_geom_store];
  const double* geom_rbound = D + H->noff[MJB_N_geom_rbound];
  const float max_margin = __double2float_ru((D + H->noff[MJB_N_scan_misc])[0]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int gp = (ngeom + 3) & ~3;
  float4* gx = reinterpret_cast<float4*>(smem) + (size_t)warp * gp;         // x, y, z, bounding radius of every geom
  const int cap = a.scan_buf_cap;
  int* buf = a.scan_buf + ((size_t)blockIdx.x * W + warp) * cap;            // this warp's pair buffer
  const size_t off_gxpos = (size_t)H->scoff[MJB_SC_geom_xpos], off_gxmat = (size_t)H->scoff[MJB_SC_geom_zaxis];
  const int nwords = (ncand + 31) >> 5;
  const unsigned below = (1u << lane) - 1u;
  const float kInf = __int_as_float(0x7f800000), kNaN = __int_as_float(0x7fc00000);

  for (long long s0 = (long long)blockIdx.x * W; s0 < a.chunk_n; s0 += (long long)gridDim.x * W) {
    const long long s = s0 + warp;
    if (s >= a.chunk_n) continue;                     // warp-uniform; no CTA-wide barrier in this kernel
    const double* sc = a.scratch + ((s >> 5) * a.nscratch << 5) + (s & 31);
    int* isc = a.iscratch + ((s >> 5) * a.niscratch << 5) + (s & 31);
    int* cm = a.cmask + (size_t)s * nwords;           // this state's survivor mask, contiguous words
    __syncwarp();
    // positions relative to a stored geom (geom 1 of the first candidate), as floats; radius: NaN = never
    // (geom outside every candidate pair), inf = always (no bounding radius: planes)
    const int gref = scan_int[0] & 0xfffffff;
    const double* gblk = sc - (s & 31) + 4 * (s & 31);       // this state's 4-double geom vectors: + (off + 4 g) * 32
    const double r0 = gblk[(off_gxpos + 4 * gref) << 5], r1 = gblk[((off_gxpos + 4 * gref) << 5) + 1],
                 r2 = gblk[((off_gxpos + 4 * gref) << 5) + 2];
    float m = 0;
    for (int g = lane; g < gp; g += 32) {
      const bool stored = g < ngeom && (geom_store[g] & 3) != 0;
      const double* gv = gblk + ((off_gxpos + 4 * (size_t)(g < ngeom ? g : 0)) << 5);
      const float x = stored ? (float)(gv[0] - r0) : 0.f;
      const float y = stored ? (float)(gv[1] - r1) : 0.f;
      const float z = stored ? (float)(gv[2] - r2) : 0.f;
      float rb = kNaN;
      if (stored) { const double r = geom_rbound[g]; rb = r > 0 ? __double2float_ru(r) : kInf; }
      gx[g] = make_float4(x, y, z, rb);
      m = fmaxf(m, fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))));
    }
    for (int w = lane; w < nwords; w += 32) cm[w] = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const bool finite = m < 1e30f;
    const float extra = max_margin + fmaxf(1e-4f, 8e-6f * m);
    __syncwarp();
    // stage 1: all pairs g1 < g2, 32 values of g2 per pass
    int count = 0;
    if (finite) {
      for (int tb = 0; tb < ngeom; tb += 32) {
        const int g2 = tb + lane;
        const float4 p2 = gx[g2 < gp ? g2 : gp - 1];
        const float r2f = (g2 < ngeom ? p2.w : kNaN) + extra;
        const int g1end = tb + 32 < ngeom ? tb + 32 : ngeom;
#pragma unroll 4
        for (int g1 = 0; g1 < g1end; g1++) {
          const float4 p1 = gx[g1];
          const float t = p1.w + r2f;
          const float dx = p1.x - p2.x, dy = p1.y - p2.y, dz = p1.z - p2.z;
          const bool maybe = (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= t * t) && g1 < g2;
          const unsigned mm = __ballot_sync(0xffffffffu, maybe);
          if (mm) {
            if (maybe) {
              const int k = count + __popc(mm & below);
              if (k < cap) buf[k] = (g1 << 16) | g2;
            }
            count += __popc(mm);
          }
        }
      }
    }
    __syncwarp();
    int total = 0;
    if (finite && count <= cap) {
      // stage 2: candidate of each buffered pair, exact test, survivor bit
      for (int k0 = 0; k0 < count; k0 += 32) {
        const int k = k0 + lane;
        bool ok = false;
        int ci = -1;
        if (k < count) {
          const int pr = buf[k];
          ci = a.pair_ci[(size_t)(pr >> 16) * ngeom + (pr & 0xffff)];
          ok = ci >= 0 && scan_exact_test(sc, off_gxpos, off_gxmat, scan_int, scan_bound, ci);
        }
        if (ok) atomicOr(&cm[ci >> 5], 1 << (ci & 31));
        total += __popc(__ballot_sync(0xffffffffu, ok));
      }
    } else {
      // non-finite positions, or more pairs than the buffer holds: exact test of the whole list
      for (int c0 = 0; c0 < ncand; c0 += 32) {
        const int ci = c0 + lane;
        const bool ok = ci < ncand && scan_exact_test(sc, off_gxpos, off_gxmat, scan_int, scan_bound, ci);
        const unsigned mm = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) cm[c0 >> 5] = (int)mm;
        total += __popc(mm);
      }
    }
    __syncwarp();
    if (lane == 0) isc[(size_t)MJB_ISC_NSURV * MJB_LS] = total;
    if (!a.cq) continue;                   // pooled path only: the masks are all it needs
    // stage 3: the mask words in order -> the state's items in candidate order
    int base = 0;
    if (lane == 0 && total) base = atomicAdd(&a.cq->nitems, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    const bool fits = base + total <= a.items_cap;
    if (!fits && lane == 0) a.cq->overflow = 1;
    if (lane == 0) isc[(size_t)MJB_ISC_ITEMBASE * MJB_LS] = fits ? base : -1;
    if (!fits || !total) continue;
    int done = 0;
    for (int w0 = 0; w0 < nwords; w0 += 32) {
      const int w = w0 + lane;
      // (read at L2: the bits were set with atomics after this warp's own zeroing stores went through L1)
      unsigned bits = w < nwords ? (unsigned)__ldcg(cm + w) : 0u;
      const int n = __popc(bits);
      int incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      int k = base + done + incl - n;
      while (bits) {
        const int b = __ffs((int)bits) - 1;
        bits &= bits - 1;
        a.items[k++] = ContactItem{(int)s, (w << 5) + b};
      }
      done += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

// ------------------------------------------------------------------------------------------
// contact kernel: warp-pooled narrow phase and contact rows.
//
// A warp owns 32 consecutive states. Work per state is very uneven (0..60 bounding-sphere
// survivors, 0..36 contacts for the humanoid), so instead of each lane expanding only its own
// state, the warp pools the work items of its 32 states and deals them out 32 at a time:
//   1. every lane expands its survivor bit mask into a private list (shared memory);
//   2. NARROW PHASE over the pooled survivors: item j belongs to (owner lane, k-th survivor),
//      found by a shuffle binary search over the exclusive prefix of the list lengths; the lane
//      binds its context to the owner's state (scratch column, outputs) and runs the narrow phase;
//      the resulting contacts are appended to a per-warp pool in shared memory at positions given
//      by a warp prefix sum, which keeps them in (owner, candidate) order;
//   3. CONTACT ROWS over the pooled contacts: contact index and first efc row of every record come
//      from a segmented warp scan on top of the owner lane's running counters (read by shuffle),
//      so the numbering equals the sequential order of the reference; the lane evaluates the rows
//      on the owner's state and adds the resulting wrench to the two bodies of the owner's state.
//      Records of one round that hit the same (owner, body) are found with __match_any_sync and
//      applied one after the other in pool order, so the sums are conflict-free and bitwise
//      deterministic (same order as a sequential sweep over the state's contacts).
// Lanes are ~fully occupied in 2 and 3, and all memory traffic stays inside the 32 scratch
// columns of the warp (same 256-byte lines), unlike a global sort of states.

// records per warp: 32 + 64 for pairs with <= 2 contacts (the humanoid), 32 + 4*maxper otherwise
__host__ __device__ inline int contact_pool_cap(int maxper) {
  return 32 + (4*maxper > 64 ? 4*maxper : 64);
}

struct PoolRec {
  int owner;
  int ci;
  double dist;
  double pos[3];
  double frame[6];     // normal, tangent (third axis is their cross product)
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// inclusive scan of v restricted to runs of equal key (keys are non-decreasing across lanes)
__device__ __forceinline__ int warp_seg_incl_scan(int v, int key, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, d);
    const int k = __shfl_up_sync(0xffffffffu, key, d);
    if (lane >= d && k == key) v += t;
  }
  return v;
}

// kConvex: the instantiation for models with mjc_Convex pairs (GJK / EPA, ~60 KB of stack per thread)
template <bool kModelInSmem, bool kConvex = false>
__global__ void __launch_bounds__(kThreads, kConvex ? 1 : MJB_CTAS_CONTACT) contact_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  // fallback of the item-parallel path: runs only when that path overflowed its lists
  if (a.cq && !(*(volatile int*)&a.cq->overflow | *(volatile int*)&a.cq->overflow_contacts)) return;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  const mjbHdr& H = *c.H;

  // dynamic shared memory after the model blob: survivor lists, then per-warp pools and counters
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // pool of contact records per warp: 31 left over from the last full-round drain plus one round
  // of narrow-phase output. Pairs that can yield many contacts (box-box: 24 before clean-up) get
  // fewer lanes per round (per_round) instead of a larger pool.
  const int maxper = H.max_pair_contacts;
  const int pool_cap = contact_pool_cap(maxper);
  const int per_round = (pool_cap - 32) / maxper < 32 ? (pool_cap - 32) / maxper : 32;
  size_t off = kModelInSmem ? (size_t)((a.model_bytes + 127) & ~127) : 0;
  int* lists = reinterpret_cast<int*>(smem + off);
  off += sizeof(int) * kListCap * kThreads;
  int* wcnt = reinterpret_cast<int*>(smem + off) + warp * 128;   // [upd_rows | upd_cons | pcount | wstatus]
  off += sizeof(int) * 128 * (kThreads / 32);
  int* hits = reinterpret_cast<int*>(smem + off) + warp * (32 * kListCap);   // (owner << 27) | candidate
  off += sizeof(int) * 32 * kListCap * (kThreads / 32);
  off = (off + 15) & ~(size_t)15;
  PoolRec* pool = reinterpret_cast<PoolRec*>(smem + off) + (size_t)warp * pool_cap;
  int* upd_rows = wcnt; int* upd_cons = wcnt + 32; int* wstatus = wcnt + 96;
  int* mylist = lists + threadIdx.x;
  upd_rows[lane] = 0; upd_cons[lane] = 0; wstatus[lane] = 0;
  __syncwarp();

  const int nwords = (H.ncand + 31) >> 5;
  for (long long i0 = (long long)blockIdx.x * kThreads; i0 < a.chunk_n;
       i0 += (long long)gridDim.x * kThreads) {
    const long long wbase = i0 + warp * 32;          // first state (chunk-local) of this warp
    const long long i = wbase + lane;
    const bool valid = i < a.chunk_n;
    Ctx own = c;
    bind_state(own, a, valid ? i : 0LL);
    if (valid) load_counters(own);
    int ncon = own.ncon, nefc = own.nefc;            // running counters of the state this lane owns
    int w = 0;
    // survivor mask of this lane's state: interleaved rows of iscratch (thread-per-state scan) or the
    // contiguous per-state words the warp-per-state scans write
    const int* cmw = (a.cmask && valid) ? a.cmask + (size_t)i * nwords : nullptr;
    auto mask_word = [&](int ww) -> unsigned {
      return cmw ? (unsigned)cmw[ww] : (unsigned)own.isc[(size_t)(MJB_ISC_MASK + ww) * MJB_LS];
    };
    unsigned bits = (valid && nwords > 0) ? mask_word(0) : 0u;
    int pool_n = 0;

    // step 3 on the current pool
    // only full rounds of 32 records are processed until the warp's last call (final), so that the
    // row arithmetic always runs with all lanes busy; the remainder moves to the pool's front
    auto drain = [&](bool final) {
      const int* body_static = c.I + H.ioff[MJB_I_body_static];
      const int nproc = final ? pool_n : (pool_n & ~31);
      for (int r0 = 0; r0 < nproc; r0 += 32) {
        const int r = r0 + lane;
        const bool has = r < pool_n;
        int owner = 64 + lane, ci = 0, rows = 0, exclude = 0;
        Ctx co = c;
        if (has) {
          owner = pool[r].owner; ci = pool[r].ci;
          bind_state(co, a, wbase + owner);
          rows = contact_row_count(co, ci, pool[r].dist, &exclude);
        }
        const int rows_incl = warp_seg_incl_scan(rows, owner, lane);
        const int cons_incl = warp_seg_incl_scan(has ? 1 : 0, owner, lane);
        const int src = has ? owner : lane;
        const int base_ncon = __shfl_sync(0xffffffffu, ncon, src);
        const int base_nefc = __shfl_sync(0xffffffffu, nefc, src);
        const int next_owner = __shfl_down_sync(0xffffffffu, owner, 1);
        double F[3] = {0, 0, 0}, T3[3] = {0, 0, 0}, p[3] = {0, 0, 0};
        int b1 = 0, b2 = 0;
        if (has) {
          Con con;
          con.dist = pool[r].dist;
          for (int k = 0; k < 3; k++) { con.pos[k] = pool[r].pos[k]; p[k] = con.pos[k]; }
          for (int k = 0; k < 6; k++) con.frame[k] = pool[r].frame[k];
          cross3(con.frame + 6, con.frame, con.frame + 3);         // as mju_makeFrame's last step
          co.status = 0;
          contact_rows(co, ci, con, base_ncon + cons_incl - 1, exclude,
                       rows ? base_nefc + rows_incl - rows : -1, F, T3);
          const int* cint = c.I + H.ioff[MJB_I_cand_int] + MJB_CAND_NI*ci;
          b1 = cint[MJB_CI_B1]; b2 = cint[MJB_CI_B2];
          if (co.status) atomicOr(&wstatus[owner], co.status);
          if (lane == 31 || next_owner != owner) { upd_rows[owner] = rows_incl; upd_cons[owner] = cons_incl; }
        }
        // J'f: + wrench on body 2, - wrench on body 1 of the owner's state (static bodies never
        // reach a dof and are skipped); same-target records go in pool order
#pragma unroll
        for (int side = 0; side < 2; side++) {
          const int body = side == 0 ? b2 : b1;
          const bool act = has && rows > 0 && !body_static[body];
          const unsigned key = act ? (((unsigned)owner << 20) | (unsigned)body) : (0x80000000u | lane);
          const unsigned grp = __match_any_sync(0xffffffffu, key);
          const int rank = __popc(grp & ((1u << lane) - 1u));
          const int maxrank = __reduce_max_sync(0xffffffffu, rank);
          for (int t = 0; t <= maxrank; t++) {
            if (act && rank == t) add_wrench(co, body, p, F, T3, side == 0);
            __syncwarp();
          }
        }
        __syncwarp();
        ncon += upd_cons[lane]; nefc += upd_rows[lane];
        upd_cons[lane] = 0; upd_rows[lane] = 0;
        __syncwarp();
      }
      const int rem = pool_n - nproc;
      if (rem > 0 && nproc > 0) {
        PoolRec keep;
        if (lane < rem) keep = pool[nproc + lane];
        __syncwarp();
        if (lane < rem) pool[lane] = keep;
        __syncwarp();
      }
      pool_n = rem;
    };

    while (true) {
      // step 1: expand the mask into the private list
      int cnt = 0;
      if (valid) {
        while (cnt < kListCap) {
          while (bits == 0 && w + 1 < nwords) {
            w++;
            bits = mask_word(w);
          }
          if (bits == 0) break;
          const int b = __ffs((int)bits) - 1;
          bits &= bits - 1;
          mylist[cnt * kThreads] = (w << 5) + b;
          cnt++;
        }
      }
      const int incl = warp_incl_scan(cnt, lane);
      const int excl = incl - cnt;
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total == 0) break;
      __syncwarp();

      // step 2a: exact hit test of every pooled survivor, all lanes busy; hits are compacted in
      // order into the warp's hit list as (owner, candidate)
      int nhit = 0;
      for (int j0 = 0; j0 < total; j0 += 32) {
        const int j = j0 + lane;
        const bool has = j < total;
        int o = 0;                                   // largest lane with excl <= j
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const int t = o + step;
          const int e = __shfl_sync(0xffffffffu, excl, t & 31);
          if (t < 32 && e <= j) o = t;
        }
        const int eo = __shfl_sync(0xffffffffu, excl, o);
        bool hit = false;
        int ci = 0;
        if (has) {
          ci = lists[(j - eo) * kThreads + warp * 32 + o];
          Ctx co = c;
          bind_state(co, a, wbase + o);
          hit = narrow_test(co, ci);
        }
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hit) {
          const int dst = nhit + __popc(hm & ((1u << lane) - 1u));
          hits[dst] = (o << 27) | ci;
        }
        nhit += __popc(hm);
      }
      __syncwarp();

      // step 2b: narrow phase of the hits
      for (int j0 = 0; j0 < nhit; j0 += per_round) {
        const int j = j0 + lane;
        const bool has = lane < per_round && j < nhit;
        Con con[MJB_MAXCON_PAIR];
        int num = 0, ci = 0, o = 0;
        if (has) {
          o = (unsigned)hits[j] >> 27; ci = hits[j] & 0x7ffffff;
          Ctx co = c;
          bind_state(co, a, wbase + o);
          num = narrow_pair<false, kConvex>(co, ci, con);
        }
        const int nincl = warp_incl_scan(num, lane);
        const int dst = pool_n + nincl - num;
        for (int k = 0; k < num; k++) {
          PoolRec& rec = pool[dst + k];
          rec.owner = o; rec.ci = ci; rec.dist = con[k].dist;
          for (int q = 0; q < 3; q++) rec.pos[q] = con[k].pos[q];
          for (int q = 0; q < 6; q++) rec.frame[q] = con[k].frame[q];
        }
        pool_n += __shfl_sync(0xffffffffu, nincl, 31);
        __syncwarp();
        if (pool_n + per_round*maxper > pool_cap) drain(false);
      }
      // a pass ends with an empty pool: the next pass restarts at owner 0, and the row numbering
      // of a round relies on every owner forming ONE run of consecutive records
      if (pool_n) drain(true);
    }

    own.ncon = ncon; own.nefc = nefc;
    own.status |= wstatus[lane];
    wstatus[lane] = 0;
    if (valid) save_counters(own);
    __syncwarp();
  }
}


// ------------------------------------------------------------------------------------------
// Item-parallel contact phase (default path). The pooled kernel above serialises the work of 32
// states inside one warp at 8 warps per SM; here every stage is its own kernel with one thread
// per work item and no shared-memory pools, so each runs at full occupancy with dense lanes:
//   (scan tail)   : state  -> its bounding-sphere survivors appended to the global item list
//                             (scan_append_items, mjb_launch.h; the warp-per-state scan does it itself)
//   contact_narrow: item   -> exact hit test; hits are compacted per CTA and run the narrow phase;
//                             contacts are appended to the global contact list
//   contact_index : state  -> walks its items in candidate order: contact index k and first efc row
//                             of every contact (the reference's sequential numbering)
//   contact_rows  : contact-> constraint rows; J'f kept in the record
// and the backward kernel adds the records' wrenches to its own state in contact order.

constexpr int kNarrowTiles = 4;      // tiles of 256 items tested per dense narrow-phase round

// four resident CTAs (64 registers, 156 B of spills) beat three (80 registers): 1.31 -> 1.22 ms (GPU call Y)
#ifndef MJB_NARROW_SIMPLE_CTAS
#define MJB_NARROW_SIMPLE_CTAS 4
#endif
template <bool kModelInSmem, bool kSimple, bool kConvex = false>
__global__ void __launch_bounds__(256, kSimple ? MJB_NARROW_SIMPLE_CTAS : (kConvex ? 1 : 2)) contact_narrow_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ int hitlist[256 * kNarrowTiles];
  __shared__ int nhit;
  if (*(volatile int*)&a.cq->overflow) return;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  const int lane = threadIdx.x & 31;
  const int n = a.cq->nitems;
  const int span = 256 * kNarrowTiles;
  for (int tile = blockIdx.x * span; tile < n; tile += gridDim.x * span) {
    if (threadIdx.x == 0) nhit = 0;
    __syncthreads();
    // exact hit test of `span` items, all lanes busy; hits are collected per CTA
    for (int t = 0; t < kNarrowTiles; t++) {
      const int i = tile + t * 256 + threadIdx.x;
      bool hit = false;
      if (i < n) {
        const ContactItem it = a.items[i];
        Ctx& co = c;      // rebinding c is enough (a by-value copy of the context faulted here for
        bind_state(co, a, it.state);   //  models read from global memory; cause not established)
        hit = narrow_test(co, it.ci);
        a.item_con[i] = ItemCon{0, 0, 0, 0};
      }
      const unsigned hm = __ballot_sync(0xffffffffu, hit);
      int wbase = 0;
      if (lane == 0 && hm) wbase = atomicAdd(&nhit, __popc(hm));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (hit) hitlist[wbase + __popc(hm & ((1u << lane) - 1u))] = i;
    }
    __syncthreads();
    const int nh = nhit;
    // dense narrow phase over the CTA's hits (order inside the list is irrelevant)
    for (int h0 = 0; h0 < nh; h0 += 256) {
      Con con[kSimple ? 2 : MJB_MAXCON_PAIR];
      int num = 0, item = -1;
      ContactItem it = {0, 0};
      if (h0 + (int)threadIdx.x < nh) {
        item = hitlist[h0 + threadIdx.x];
        it = a.items[item];
        Ctx& co = c;      // (a by-value copy of the context is what faulted here; rebinding c is enough)
        bind_state(co, a, it.state);
        num = narrow_pair<kSimple, kConvex>(co, it.ci, con);
      }
      if (__any_sync(0xffffffffu, num > 0)) {
        const int incl = warp_incl_scan(num, lane);
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == 0) base = atomicAdd(&a.cq->ncontacts, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + total > a.contacts_cap) {
          if (lane == 0) a.cq->overflow_contacts = 1;
        } else if (num > 0) {
          const int first = base + incl - num;
          int code[2] = {0, 0};
          for (int k = 0; k < num; k++) {
            // rows this contact will occupy (mj_instantiateContact's counts), or -1 - exclude
            int exclude;
            const int rows = contact_row_count(c, it.ci, con[k].dist, &exclude);
            const int cd = rows ? rows : -1 - exclude;
            if (k < 2) code[k] = cd;
            ContactRec& r = a.contacts[first + k];
            r.state = it.state; r.ci = it.ci; r.k = -1; r.efc_address = cd;
            r.dist = con[k].dist;
            for (int q = 0; q < 3; q++) r.pos[q] = con[k].pos[q];
            for (int q = 0; q < 6; q++) r.frame[q] = con[k].frame[q];
          }
          a.item_con[item] = ItemCon{first, num, code[0], code[1]};
        }
      }
    }
    __syncthreads();
  }
}

template <bool kModelInSmem>
__global__ void __launch_bounds__(kThreads, 4) contact_index_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  if (*(volatile int*)&a.cq->overflow | *(volatile int*)&a.cq->overflow_contacts) return;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  const int lane = threadIdx.x & 31;
  for (long long i0 = (long long)blockIdx.x * kThreads; i0 < a.chunk_n; i0 += (long long)gridDim.x * kThreads) {
    const long long i = i0 + threadIdx.x;
    const bool valid = i < a.chunk_n;
    bind_state(c, a, valid ? i : 0);
    const int nsurv = valid ? c.isc[MJB_ISC_NSURV * MJB_LS] : 0;
    const int ibase = valid ? c.isc[MJB_ISC_ITEMBASE * MJB_LS] : 0;
    // contacts of this state: a contiguous slot range, allocated warp-wide (order across warps is
    // irrelevant, inside a warp it follows the states)
    int mine = 0;
    for (int j = 0; j < nsurv; j++) mine += a.item_con[ibase + j].count;
    const int incl = warp_incl_scan(mine, lane);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 0 && total) base = atomicAdd(&a.cq->nslots, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    const int sbase = base + incl - mine;
    if (valid) {
      c.isc[MJB_ISC_ITEMBASE * MJB_LS] = sbase;        // from here on: first slot of the state ...
      c.isc[MJB_ISC_NSURV * MJB_LS] = mine;            // ... and its number of contacts
    }
    if (!mine) continue;
    load_counters(c);
    int slot = sbase;
    for (int j = 0; j < nsurv; j++) {
      const ItemCon ic = a.item_con[ibase + j];
      for (int q = 0; q < ic.count; q++) {
        const int cd = q == 0 ? ic.code0 : (q == 1 ? ic.code1 : a.contacts[ic.base + q].efc_address);
        const int rows = cd > 0 ? cd : 0;
        // first efc row, or < 0: no rows, exclude flag = -1 - value
        a.slot_rec[slot++] = SlotRec{ic.base + q, c.ncon++, rows ? c.nefc : cd, 0};
        c.nefc += rows;
      }
    }
    save_counters(c);
  }
}

// The same numbering for scenes with hundreds of items per state (22 humanoids: 1,210 items, 322
// contacts): ONE WARP per state, 32 items at a time -- item records are read coalesced, contact
// indices and row addresses come from warp prefix sums over the items' contact and row counts
// instead of a 1,210-step walk by a single thread (3.2 ms per 2^15 scene states, 4 % issue-active).
// The state's slots are its own contiguous range; contact_rows_kernel<.., true> walks it per state.
template <bool kModelInSmem>
__global__ void __launch_bounds__(kThreads, 4) contact_index_wide_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  if (*(volatile int*)&a.cq->overflow | *(volatile int*)&a.cq->overflow_contacts) return;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int W = kThreads / 32;
  for (long long s0 = (long long)blockIdx.x * W; s0 < a.chunk_n; s0 += (long long)gridDim.x * W) {
    const long long s = s0 + warp;
    if (s >= a.chunk_n) continue;                          // warp-uniform
    bind_state(c, a, s);
    const int nsurv = c.isc[MJB_ISC_NSURV * MJB_LS];
    const int ibase = c.isc[MJB_ISC_ITEMBASE * MJB_LS];
    int mine = 0;
    for (int j0 = 0; j0 < nsurv; j0 += 32) {
      const int j = j0 + lane;
      if (j < nsurv) mine += a.item_con[ibase + j].count;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    int sbase = 0;
    if (lane == 0 && mine) sbase = atomicAdd(&a.cq->nslots, mine);
    sbase = __shfl_sync(0xffffffffu, sbase, 0);
    __syncwarp();                                          // every lane has read the item range
    if (lane == 0) {
      c.isc[MJB_ISC_ITEMBASE * MJB_LS] = sbase;            // from here on: first slot of the state ...
      c.isc[MJB_ISC_NSURV * MJB_LS] = mine;                // ... and its number of contacts
    }
    if (!mine) continue;
    load_counters(c);
    const int ncon0 = c.ncon;
    int ncon = c.ncon, nefc = c.nefc;
    for (int j0 = 0; j0 < nsurv; j0 += 32) {
      const int j = j0 + lane;
      ItemCon ic = {0, 0, 0, 0};
      if (j < nsurv) ic = a.item_con[ibase + j];
      auto code_of = [&](int q) { return q == 0 ? ic.code0 : (q == 1 ? ic.code1 : a.contacts[ic.base + q].efc_address); };
      int rsum = 0;
      for (int q = 0; q < ic.count; q++) { const int cd = code_of(q); rsum += cd > 0 ? cd : 0; }
      const int kincl = warp_incl_scan(ic.count, lane), rincl = warp_incl_scan(rsum, lane);
      int k = ncon + kincl - ic.count, row = nefc + rincl - rsum;
      for (int q = 0; q < ic.count; q++) {
        const int cd = code_of(q);
        const int rows = cd > 0 ? cd : 0;
        a.slot_rec[sbase + (k - ncon0)] = SlotRec{ic.base + q, k, rows ? row : cd, 0};
        k++; row += rows;
      }
      ncon += __shfl_sync(0xffffffffu, kincl, 31);
      nefc += __shfl_sync(0xffffffffu, rincl, 31);
    }
    c.ncon = ncon; c.nefc = nefc;
    if (lane == 0) save_counters(c);
  }
}

// One warp per 32 consecutive states: their contacts occupy one contiguous, (state, k)-ordered slot
// range (contact_index_kernel), which the warp processes 32 slots at a time. J'f is applied to the
// owner state's body accumulators right here; records of one round that hit the same (state, body)
// are found with __match_any_sync and applied one after the other in slot order, so the sums are
// conflict-free and bitwise deterministic. No other warp touches these states.
// kWarpPerState: the slots were numbered by contact_index_wide_kernel, one warp walks ONE state.
template <bool kModelInSmem, bool kWarpPerState>
__global__ void __launch_bounds__(kThreads, MJB_ROWS_CTAS) contact_rows_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  if (*(volatile int*)&a.cq->overflow | *(volatile int*)&a.cq->overflow_contacts) return;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int* cand_int = c.I + c.H->ioff[MJB_I_cand_int];
  const int* body_static = c.I + c.H->ioff[MJB_I_body_static];
  const long long per_cta = kWarpPerState ? kThreads / 32 : kThreads;      // states per CTA and pass
  for (long long i0 = (long long)blockIdx.x * per_cta; i0 < a.chunk_n; i0 += (long long)gridDim.x * per_cta) {
    const long long w0 = kWarpPerState ? i0 + warp : i0 + warp * 32;      // first (only) state of this warp
    int total, slot0;
    if (kWarpPerState) {
      if (w0 >= a.chunk_n) continue;                      // warp-uniform
      const int* isc = a.iscratch + ((w0 >> 5) * a.niscratch << 5) + (w0 & 31);
      total = isc[MJB_ISC_NSURV * MJB_LS];
      slot0 = isc[MJB_ISC_ITEMBASE * MJB_LS];
    } else {
      const long long i = w0 + lane;
      const bool valid = i < a.chunk_n;
      const int* isc = a.iscratch + (((valid ? i : 0) >> 5) * a.niscratch << 5) + ((valid ? i : 0) & 31);
      const int mine = valid ? isc[MJB_ISC_NSURV * MJB_LS] : 0;
      const int sbase = valid ? isc[MJB_ISC_ITEMBASE * MJB_LS] : 0;
      const int incl = warp_incl_scan(mine, lane);
      total = __shfl_sync(0xffffffffu, incl, 31);
      slot0 = __shfl_sync(0xffffffffu, sbase, 0);         // lane 0's base = first slot of the warp
    }
    if (!total) continue;
    int wstatus = 0;
    for (int r0 = 0; r0 < total; r0 += 32) {
      const bool has = r0 + lane < total;
      double F[3] = {0, 0, 0}, T3[3] = {0, 0, 0}, p[3] = {0, 0, 0};
      int b1 = 0, b2 = 0, owner = 64 + lane;
      bool active = false;
      Ctx& co = c;
      if (has) {
        const SlotRec sr = a.slot_rec[slot0 + r0 + lane];
        const ContactRec& r = a.contacts[sr.rec];
        owner = r.state - (int)w0;
        bind_state(co, a, r.state);
        Con con;
        con.dist = r.dist;
        for (int k = 0; k < 3; k++) { con.pos[k] = r.pos[k]; p[k] = r.pos[k]; }
        for (int k = 0; k < 6; k++) con.frame[k] = r.frame[k];
        cross3(con.frame + 6, con.frame, con.frame + 3);          // as mju_makeFrame's last step
        const int exclude = sr.efc_address >= 0 ? 0 : -1 - sr.efc_address;
        co.status = 0;
        contact_rows(co, r.ci, con, sr.k, exclude, sr.efc_address >= 0 ? sr.efc_address : -1, F, T3);
        active = sr.efc_address >= 0;
        b1 = cand_int[MJB_CAND_NI*r.ci + MJB_CI_B1]; b2 = cand_int[MJB_CAND_NI*r.ci + MJB_CI_B2];
        if (co.status) wstatus |= co.status, atomicOr(&co.isc[MJB_ISC_STATUS * MJB_LS], co.status);
      }
      // J'f: records of this round that hit the same (state, body) form a group (__match_any_sync);
      // the group's wrenches are summed in lane order with shuffles by its first lane, which then
      // does ONE batched read-modify-write of the accumulator row (deterministic, conflict-free)
#pragma unroll
      for (int side = 0; side < 2; side++) {
        const int body = side == 0 ? b2 : b1;
        const bool act = has && active && !body_static[body];
        double W[6] = {0, 0, 0, 0, 0, 0};
        if (act) {
          double o4[4];                 // tree origin from the body's carrier record (same line as its carriers)
          ld_rec4(o4, crec_ptr(co, body) + 3*MJB_CREC_PART);
          const double rr[3] = {p[0] - o4[0], p[1] - o4[1], p[2] - o4[2]};
          double cr[3];
          cross3(cr, rr, F);
          for (int k = 0; k < 3; k++) { W[k] = cr[k] + T3[k]; W[3 + k] = F[k]; }
        }
        const unsigned key = act ? (((unsigned)owner << 20) | (unsigned)body) : (0x80000000u | lane);
        const unsigned grp = __match_any_sync(0xffffffffu, key);
        const int cnt = __popc(grp);
        const int maxcnt = __reduce_max_sync(0xffffffffu, cnt);
        const bool leader = act && (__ffs((int)grp) - 1) == lane;
        // the chain starts from the accumulator row and adds the members one by one in slot order:
        // the same association as one read-modify-write per contact, so the bits do not depend on
        // where a state sits in the batch (which contacts happen to share a round)
        double* fe = co.sc + (size_t)c.H->scoff[side == 0 ? MJB_SC_cfrc_ext : MJB_SC_cfrc_ext1] * MJB_LS;
        double acc[6] = {0, 0, 0, 0, 0, 0};
        // the first wrench on this (state, body, side) starts from zero: the accumulator rows are
        // not cleared per state, their validity is the state's wrench mask (mjb_pipeline.h)
        if (leader && wmask_test_and_set(co, body, side == 0)) {
          for (int k = 0; k < 6; k++) acc[k] = fe[(size_t)(6*body + k) * MJB_LS];
        }
        for (int k = 0; k < 6; k++) acc[k] += W[k];
        for (int t = 1; t < maxcnt; t++) {
          const int src = t < cnt ? (int)__fns(grp, 0, t + 1) : lane;      // t-th further member of my group
#pragma unroll
          for (int k = 0; k < 6; k++) {
            const double v = __shfl_sync(0xffffffffu, W[k], src);
            if (t < cnt) acc[k] += v;
          }
        }
        if (leader) for (int k = 0; k < 6; k++) fe[(size_t)(6*body + k) * MJB_LS] = acc[k];
      }
      __syncwarp();
    }
    (void)wstatus;
  }
}

size_t contact_smem_bytes(int model_bytes, int model_in_smem, int max_pair_contacts) {
  size_t off = model_in_smem ? (size_t)((model_bytes + 127) & ~127) : 0;
  off += sizeof(int) * kListCap * kThreads;
  off += sizeof(int) * 128 * (kThreads / 32);
  off += sizeof(int) * 32 * kListCap * (kThreads / 32);
  off = (off + 15) & ~(size_t)15;
  off += sizeof(PoolRec) * (size_t)contact_pool_cap(max_pair_contacts) * (kThreads / 32);
  return off;
}

template <bool kModelInSmem, bool kGravcomp>
__global__ void __launch_bounds__(kThreads, MJB_CTAS_BACKWARD) backward_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.chunk_n;
       i += (long long)gridDim.x * kThreads) {
    bind_state(c, a, i);
    phase_backward<kGravcomp>(c);
  }
}

// qfrc_bias of the chunk (mj_rne without accelerations), only when requested
template <bool kModelInSmem>
__global__ void __launch_bounds__(kThreads, 4) bias_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.chunk_n;
       i += (long long)gridDim.x * kThreads) {
    bind_state(c, a, i);
    bias_forces(c);
  }
}

// mj_compareFwdInv over the chunk (only from mjb_compareFwdInv)
template <bool kModelInSmem>
__global__ void __launch_bounds__(kThreads, 4) fwdinv_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.chunk_n;
       i += (long long)gridDim.x * kThreads) {
    bind_state(c, a, i);
    compare_fwdinv(c);
  }
}

// sensordata of the chunk (mj_sensorPos / Vel / Acc); launched only for models with sensors, after
// the backward kernel has left cacc / cfrc_int in the outputs
template <bool kModelInSmem, bool kCcd = false>
__global__ void __launch_bounds__(kThreads, kCcd ? 1 : 4) sensor_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.chunk_n;
       i += (long long)gridDim.x * kThreads) {
    bind_state(c, a, i);
    sensors<kCcd>(c);
  }
}

// d->energy of models with mjENBL_ENERGY (mj_energyPos / mj_energyVel)
template <bool kModelInSmem>
__global__ void __launch_bounds__(kThreads, 4) energy_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.chunk_n;
       i += (long long)gridDim.x * kThreads) {
    bind_state(c, a, i);
    energy(c);
  }
}

// cam_xpos / cam_xmat / light_xpos / light_xdir (mj_camlight), only when mjbOUT_CAMLIGHT is requested
template <bool kModelInSmem>
__global__ void __launch_bounds__(kThreads, 4) camlight_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.chunk_n;
       i += (long long)gridDim.x * kThreads) {
    bind_state(c, a, i);
    camlight(c);
  }
}

// actuator_length / actuator_moment / actuator_velocity (mj_transmission), only when mjbOUT_TRANSMISSION is requested
template <bool kModelInSmem>
__global__ void __launch_bounds__(kThreads, 4) transmission_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.chunk_n;
       i += (long long)gridDim.x * kThreads) {
    bind_state(c, a, i);
    transmission(c);
  }
}

// d->xfrc_applied folded into cfrc_ext / cfrc_int (mj_rnePostConstraint), only when the caller has set it
template <bool kModelInSmem>
__global__ void __launch_bounds__(kThreads, 4) post_xfrc_kernel(LaunchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  Ctx c;
  make_ctx<kModelInSmem>(c, a, smem, &mbar);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.chunk_n;
       i += (long long)gridDim.x * kThreads) {
    bind_state(c, a, i);
    post_xfrc(c);
  }
}

size_t inverse_smem_bytes(int model_bytes, int model_in_smem) {
  return model_in_smem ? static_cast<size_t>(model_bytes) : 0;
}

size_t smooth_smem_bytes(int model_bytes, int model_in_smem) {
  return (model_in_smem ? (size_t)((model_bytes + 127) & ~127) : 0) +
         sizeof(double) * MJB_SM_SLOTS * kSmoothThreads;
}

template <typename K>
static cudaError_t launch_phase(K kernel, const LaunchArgs& args, size_t smem, int ctas_per_sm,
                                cudaStream_t stream, int threads = kThreads, int resident = 0,
                                int states_per_cta = 0) {
  const int per = states_per_cta > 0 ? states_per_cta : threads;
  int grid = (int)(((long long)args.chunk_n + per - 1) / per);
  const int cap = kSMs * ctas_per_sm;
  if (grid > cap) grid = cap;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
#ifndef MJB_CARVEOUT_HINT
#define MJB_CARVEOUT_HINT 0   // measured: the explicit carveout costs the contact kernel 13 % (less L1)
#endif
  if (MJB_CARVEOUT_HINT && smem > 0 && resident > 0) {
    // ask for the shared-memory carveout that lets all CTAs this kernel is compiled for be resident
    // (the default heuristic left the contact kernel at about half of its intended occupancy)
    const size_t want = (smem + 1024) * (size_t)resident;
    int pct = (int)((want * 100 + 227 * 1024 - 1) / (227 * 1024));
    if (pct > 100) pct = 100;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  }
  kernel<<<grid, threads, smem, stream>>>(args);
  return cudaGetLastError();
}

namespace {
// the smooth kernel instantiation that contains the spatial-tendon path walk is used only by models
// with force-carrying spatial tendons (it needs more registers than the plain one)
cudaError_t launch_smooth(const LaunchArgs& a, cudaStream_t stream) {
  const size_t sm = smooth_smem_bytes(a.model_bytes, a.model_in_smem);
  const int cap = 2 * MJB_CTAS_SMOOTH;
  if (a.has_spatial) {
    return a.model_in_smem ? launch_phase(smooth_kernel<true, true>, a, sm, cap, stream, kSmoothThreads)
                           : launch_phase(smooth_kernel<false, true>, a, sm, cap, stream, kSmoothThreads);
  }
  return a.model_in_smem ? launch_phase(smooth_kernel<true, false>, a, sm, cap, stream, kSmoothThreads)
                         : launch_phase(smooth_kernel<false, false>, a, sm, cap, stream, kSmoothThreads);
}

cudaError_t generic_inertia(const LaunchArgs& a, cudaStream_t stream) {
  const size_t smem = inverse_smem_bytes(a.model_bytes, a.model_in_smem);
  return a.model_in_smem ? launch_phase(inertia_kernel<true>, a, smem, 8, stream)
                         : launch_phase(inertia_kernel<false>, a, 0, 8, stream);
}
cudaError_t generic_scan(const LaunchArgs& a, cudaStream_t stream) {
  const size_t smem = inverse_smem_bytes(a.model_bytes, a.model_in_smem);
  return a.model_in_smem ? launch_phase(contact_scan_kernel<true>, a, smem, 8, stream)
                         : launch_phase(contact_scan_kernel<false>, a, 0, 8, stream);
}
cudaError_t generic_backward(const LaunchArgs& a, cudaStream_t stream) {
  const size_t smem = inverse_smem_bytes(a.model_bytes, a.model_in_smem);
  if (a.has_gravcomp) {
    return a.model_in_smem ? launch_phase(backward_kernel<true, true>, a, smem, 8, stream)
                           : launch_phase(backward_kernel<false, true>, a, 0, 8, stream);
  }
  return a.model_in_smem ? launch_phase(backward_kernel<true, false>, a, smem, 8, stream)
                         : launch_phase(backward_kernel<false, false>, a, 0, 8, stream);
}

// one specialised kernel over the chunk, same grid rule as launch_phase
cudaError_t launch_spec(void* fn, const LaunchArgs& args, int ctas_per_sm, cudaStream_t stream) {
  int grid = (args.chunk_n + kThreads - 1) / kThreads;
  if (grid > kSMs * ctas_per_sm) grid = kSMs * ctas_per_sm;
  return jitLaunch(fn, grid, kThreads, 0, stream, args);
}

// one phase of a chunk: the specialised kernels (one per stage, in order) when there are any, else
// the generic kernel
template <typename G>
cudaError_t run_phase(const std::vector<void*>* stages, void* single, const LaunchArgs& args,
                      cudaStream_t stream, int* launches, G generic) {
  if (stages && !stages->empty()) {
    for (void* fn : *stages) {
      ++*launches;
      cudaError_t e = launch_spec(fn, args, 8, stream);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  ++*launches;
  return single ? launch_spec(single, args, 8, stream) : generic(args, stream);
}

struct PhaseScope {   // records begin/end events around one kernel launch when timing is on
  const PhaseTimer* t; cudaStream_t s; int phase; cudaEvent_t b;
  PhaseScope(const PhaseTimer* t_, cudaStream_t s_, int phase_) : t(t_), s(s_), phase(phase_), b(nullptr) {
    if (t) { b = t->next_event(t->ctx); cudaEventRecord(b, s); }
  }
  ~PhaseScope() {
    if (t) { cudaEvent_t e = t->next_event(t->ctx); cudaEventRecord(e, s); t->mark(t->ctx, phase, b, e); }
  }
};
}  // namespace

cudaError_t launch_inverse(const LaunchArgs& args, cudaStream_t stream, int* launches,
                           const PhaseTimer* timer, const SpecKernels* spec) {
  *launches = 0;
  if (args.chunk_n <= 0) return cudaSuccess;
  const size_t smem = inverse_smem_bytes(args.model_bytes, args.model_in_smem);
  const bool in_smem = args.model_in_smem != 0;
  const bool want_inertia = args.out.qM || args.out.qLD || args.out.qLDiagInv;
  const std::vector<void*>* const st_smooth = spec ? &spec->smooth : nullptr;
  const std::vector<void*>* const st_inertia = spec ? &spec->inertia : nullptr;
  void* const fn_scan = spec ? spec->contact_scan : nullptr;
  void* const fn_backward = spec ? spec->backward : nullptr;
  cudaError_t e;
  if (args.qacc_discrete) {
    // mjENBL_INVDISCRETE (engine_inverse.c:227-234): position stage + factorisation on the given
    // state, qacc converted, then the whole pipeline on the converted accelerations
    LaunchArgs pre = args;
    pre.qacc_discrete = nullptr;
    pre.has_contacts = 0;
    { PhaseScope ps(timer, stream, kPhaseSmooth);
    e = run_phase(st_smooth, nullptr, pre, stream, launches, launch_smooth); }
    if (e != cudaSuccess) return e;
    { PhaseScope ps(timer, stream, kPhaseInertia);
    e = run_phase(st_inertia, nullptr, pre, stream, launches, generic_inertia); }
    if (e != cudaSuccess) return e;
    if (args.out.actuator_length) {
      // implicitfast with velocity-biased actuators: mjd_actuator_vel reads actuator_moment
      e = launch_phase(transmission_kernel<false>, pre, 0, 8, stream);
      if (e != cudaSuccess) return e;
      *launches += 1;
    }
    { PhaseScope ps(timer, stream, kPhaseDiscrete);
    e = launch_phase(discrete_acc_kernel<false>, args, 0, 8, stream); }     // reads a cold table (act_biasvel)
    if (e != cudaSuccess) return e;
    *launches += 1;
    LaunchArgs post = args;
    post.qacc = args.qacc_discrete;
    post.qacc_discrete = nullptr;
    int n2 = 0;
    e = launch_inverse(post, stream, &n2, timer, spec);
    *launches += n2;
    return e;
  }
  if (want_inertia && spec && !spec->tree.empty()) {
    // specialised tree stages: forward sweep and mj_crb / mj_factorM fused per subtree (phase_tree)
    PhaseScope ps(timer, stream, kPhaseTree);
    e = run_phase(&spec->tree, nullptr, args, stream, launches, launch_smooth);
    if (e != cudaSuccess) return e;
  } else {
    { PhaseScope ps(timer, stream, kPhaseSmooth);
    e = run_phase(st_smooth, nullptr, args, stream, launches, launch_smooth); }
    if (e != cudaSuccess) return e;
    if (want_inertia) {
      PhaseScope ps(timer, stream, kPhaseInertia);
      if (args.inertia_subwarp) {
        // 8 lanes per state, intermediates in shared memory (inertia_subwarp_kernel)
        const size_t wsm = subwarp_layout(args.sub_nv, args.sub_nbody, args.sub_nC).bytes;
        e = cudaFuncSetAttribute(inertia_subwarp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm);
        if (e != cudaSuccess) return e;
        long long grid = ((long long)args.chunk_n + 31) >> 5;
        if (grid > kSMs) grid = kSMs;
        inertia_subwarp_kernel<<<(int)grid, kSubS * kSubG, wsm, stream>>>(args);
        e = cudaGetLastError();
        ++*launches;
      } else {
        e = run_phase(st_inertia, nullptr, args, stream, launches, generic_inertia);
      }
      if (e != cudaSuccess) return e;
    }
  }
  if (args.has_contacts) {
    if (args.cq) {
      e = cudaMemsetAsync(args.cq, 0, sizeof(ContactQueue), stream);
      if (e != cudaSuccess) return e;
    }
    { PhaseScope ps(timer, stream, kPhaseScan);
      if (args.scan_wide > 0) {
        // one warp per state; appends the survivors to the item list itself
        const size_t wsm = scan_wide_smem_bytes(args.scan_ngeom, args.scan_wide);
        e = cudaFuncSetAttribute(contact_scan_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm);
        if (e != cudaSuccess) return e;
        int grid = (args.chunk_n + args.scan_wide - 1) / args.scan_wide;
        if (grid > kSMs * kWideCtas) grid = kSMs * kWideCtas;     // the per-warp buffers are sized for this grid
        if (args.pair_ci) {
          // geom-pair organisation (needs only the per-warp positions in shared memory)
          const size_t psm = (size_t)args.scan_wide * (size_t)((args.scan_ngeom + 3) & ~3) * sizeof(float4);
          e = cudaFuncSetAttribute(contact_scan_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
          if (e != cudaSuccess) return e;
          contact_scan_pairs_kernel<<<grid, 32 * args.scan_wide, psm, stream>>>(args);
        } else {
          contact_scan_wide_kernel<<<grid, 32 * args.scan_wide, wsm, stream>>>(args);
        }
        e = cudaGetLastError();
        ++*launches;
      } else {
        e = run_phase(nullptr, fn_scan, args, stream, launches, generic_scan);
      }
    }
    if (e != cudaSuccess) return e;
// The contact kernel reads the model tables through L1 instead of a shared-memory copy: its lanes
// index the candidate tables with per-lane (non-uniform) indices anyway, and the 25 KB per CTA are
// worth more as L1 for the scattered scratch accesses (measured 5.54 vs 5.96 ms per 2^20 states).
// Item-path kernels read the model from global memory (through L1) instead of a staged copy: a 57 KB
// copy per CTA caps every kernel at four CTAs per SM. Measured on the humanoid (profiles/r02_*): the
// numbering kernel, which reads no model table at all, 0.39 -> 0.19 ms (warps active 17 % -> 69 %);
// the rows kernel 1.09 -> 1.05 ms; the narrow kernel 1.31 -> 1.30 ms. A fifth resident CTA for the
// rows kernel (96 registers, spills) costs more than it gives: 1.29 ms.
#ifndef MJB_INDEX_GLOBAL
#define MJB_INDEX_GLOBAL 1
#endif
#ifndef MJB_ROWS_GLOBAL
#define MJB_ROWS_GLOBAL 1
#endif
#ifndef MJB_NARROW_GLOBAL
#define MJB_NARROW_GLOBAL 1
#endif
#ifndef MJB_CONTACT_MODEL_SMEM
#define MJB_CONTACT_MODEL_SMEM 0
#endif
    if (args.cq) {
      { PhaseScope ps(timer, stream, kPhaseContact);
        // the narrow / rows kernels walk lists whose length is only known on the device: full grids
        // (the grid follows the ITEM list, not the states: a 22-humanoid scene has 1,200 items per state,
        // and a grid of chunk_n / 256 CTAs left most of the chip idle -- one CTA per state, capped)
        if (args.has_convex) {
          // GJK / EPA pairs: one resident CTA (the polytope lives on the threads' stacks), model from global memory
          e = launch_phase(contact_narrow_kernel<false, false, true>, args, 0, 2, stream, 256, 0, 1);
        } else if (args.simple_pairs) {
          e = (in_smem && !MJB_NARROW_GLOBAL) ? launch_phase(contact_narrow_kernel<true, true>, args, smem, 2 * MJB_NARROW_SIMPLE_CTAS, stream, 256, 0, 1)
                      : launch_phase(contact_narrow_kernel<false, true>, args, 0, 2 * MJB_NARROW_SIMPLE_CTAS, stream, 256, 0, 1);
        } else {
          e = in_smem ? launch_phase(contact_narrow_kernel<true, false>, args, smem, 4, stream, 256, 0, 1)
                      : launch_phase(contact_narrow_kernel<false, false>, args, 0, 4, stream, 256, 0, 1);
        }
        if (e != cudaSuccess) return e;
        if (args.scan_wide > 0) {
          // scenes with hundreds of items per state: one warp per state (full grids: the work per
          // state is only known on the device)
          const int spc = kThreads / 32;
          e = in_smem ? launch_phase(contact_index_wide_kernel<true>, args, smem, 8, stream, kThreads, 0, spc)
                      : launch_phase(contact_index_wide_kernel<false>, args, 0, 8, stream, kThreads, 0, spc);
          if (e != cudaSuccess) return e;
          e = in_smem ? launch_phase(contact_rows_kernel<true, true>, args, smem, 8, stream, kThreads, 0, spc)
                      : launch_phase(contact_rows_kernel<false, true>, args, 0, 8, stream, kThreads, 0, spc);
          if (e != cudaSuccess) return e;
        } else {
          // the numbering kernel reads no model table: no staging, occupancy bounded by registers only
          e = (in_smem && !MJB_INDEX_GLOBAL) ? launch_phase(contact_index_kernel<true>, args, smem, 8, stream)
                                             : launch_phase(contact_index_kernel<false>, args, 0, MJB_INDEX_GLOBAL ? 12 : 8, stream);
          if (e != cudaSuccess) return e;
          e = (in_smem && !MJB_ROWS_GLOBAL) ? launch_phase(contact_rows_kernel<true, false>, args, smem, 8, stream)
                                            : launch_phase(contact_rows_kernel<false, false>, args, 0, 8, stream);
          if (e != cudaSuccess) return e;
        }
      }
      *launches += 3;
    }
    const bool csm = in_smem && MJB_CONTACT_MODEL_SMEM;
    const size_t csmem = contact_smem_bytes(args.model_bytes, csm, args.max_pair_contacts);
    { PhaseScope ps(timer, stream, kPhaseContact);
    e = args.has_convex ? launch_phase(contact_kernel<false, true>, args, contact_smem_bytes(args.model_bytes, false, args.max_pair_contacts), 8, stream, kThreads, 1)
        : csm ? launch_phase(contact_kernel<true>, args, csmem, 8, stream, kThreads, MJB_CTAS_CONTACT)
            : launch_phase(contact_kernel<false>, args, csmem, 8, stream, kThreads, MJB_CTAS_CONTACT); }
    if (e != cudaSuccess) return e;
    *launches += 1;
  }
  PhaseScope ps_backward(timer, stream, kPhaseBackward);
  e = run_phase(nullptr, fn_backward, args, stream, launches, generic_backward);
  if (e != cudaSuccess) return e;
  if (args.out.xfrc_applied && args.out.cfrc_ext) {
    e = in_smem ? launch_phase(post_xfrc_kernel<true>, args, smem, 8, stream)
                : launch_phase(post_xfrc_kernel<false>, args, 0, 8, stream);
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  if (args.out.qfrc_bias) {
    e = in_smem ? launch_phase(bias_kernel<true>, args, smem, 8, stream)
                : launch_phase(bias_kernel<false>, args, 0, 8, stream);
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  if (args.out.fwdinv) {
    e = in_smem ? launch_phase(fwdinv_kernel<true>, args, smem, 8, stream)
                : launch_phase(fwdinv_kernel<false>, args, 0, 8, stream);
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  if (args.out.energy) {
    e = in_smem ? launch_phase(energy_kernel<true>, args, smem, 8, stream)
                : launch_phase(energy_kernel<false>, args, 0, 8, stream);
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  if (args.out.actuator_length) {
    // the actuator / camera / light tables are not in the staged part of the blob: model from global memory
    e = launch_phase(transmission_kernel<false>, args, 0, 8, stream);
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  if (args.out.cam_xpos) {
    e = launch_phase(camlight_kernel<false>, args, 0, 8, stream);
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  // sensors last: they read the energies, the camera poses and the transmission outputs
  if (args.out.sensordata && !args.skip_sensors) {
    e = args.sensor_ccd ? launch_phase(sensor_kernel<false, true>, args, 0, 2, stream)
        : (in_smem && !args.sensor_cold) ? launch_phase(sensor_kernel<true>, args, smem, 8, stream)
                                       : launch_phase(sensor_kernel<false>, args, 0, 8, stream);
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  return cudaSuccess;
}

// ------------------------------------------------------------------------------------------
// boundary transposes: host layout is [state][row] (array of states), device layout [row][state].
// 32x32 tiles through shared memory so that both the read and the write are coalesced.

template <typename T>
__global__ void aos_to_soa_kernel(const T* __restrict__ aos, T* __restrict__ soa, int n, int rows,
                                  long long stride) {
  __shared__ T tile[32][33];
  const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int s = s0 + i, r = r0 + threadIdx.x;
    if (s < n && r < rows) tile[i][threadIdx.x] = aos[(size_t)s * rows + r];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, s = s0 + threadIdx.x;
    if (s < n && r < rows) soa[(size_t)r * stride + s] = tile[threadIdx.x][i];
  }
}

template <typename T>
__global__ void soa_to_aos_kernel(const T* __restrict__ soa, T* __restrict__ aos, int n, int rows,
                                  long long stride) {
  __shared__ T tile[32][33];
  const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, s = s0 + threadIdx.x;
    if (s < n && r < rows) tile[i][threadIdx.x] = soa[(size_t)r * stride + s];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int s = s0 + i, r = r0 + threadIdx.x;
    if (s < n && r < rows) aos[(size_t)s * rows + r] = tile[threadIdx.x][i];
  }
}

static dim3 transpose_grid(int n, int rows) { return dim3((n + 31) / 32, (rows + 31) / 32); }

cudaError_t launch_aos_to_soa(const double* aos, double* soa, int n, int rows, long long stride,
                              cudaStream_t stream) {
  if (n <= 0 || rows <= 0) return cudaSuccess;
  aos_to_soa_kernel<double><<<transpose_grid(n, rows), dim3(32, 8), 0, stream>>>(aos, soa, n, rows, stride);
  return cudaGetLastError();
}

cudaError_t launch_soa_to_aos(const double* soa, double* aos, int n, int rows, long long stride,
                              cudaStream_t stream) {
  if (n <= 0 || rows <= 0) return cudaSuccess;
  soa_to_aos_kernel<double><<<transpose_grid(n, rows), dim3(32, 8), 0, stream>>>(soa, aos, n, rows, stride);
  return cudaGetLastError();
}

cudaError_t launch_soa_to_aos_int(const int* soa, int* aos, int n, int rows, long long stride,
                                  cudaStream_t stream) {
  if (n <= 0 || rows <= 0) return cudaSuccess;
  soa_to_aos_kernel<int><<<transpose_grid(n, rows), dim3(32, 8), 0, stream>>>(soa, aos, n, rows, stride);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// finite-difference Jacobians of the inverse dynamics (mjd_inverseFD, engine_derivative_fd.c:611):
// the perturbed states are generated on the device, evaluated as one large batch by the phase
// kernels, and differenced on the device.

__global__ void fd_expand_kernel(const unsigned char* model, const double* __restrict__ qpos,
                                 const double* __restrict__ qvel, const double* __restrict__ qacc,
                                 long long stride_in, long long first, int nstate, double eps,
                                 double* xqpos, double* xqvel, double* xqacc, long long stride_out) {
  const mjbHdr* H = reinterpret_cast<const mjbHdr*>(model);
  const int* I = reinterpret_cast<const int*>(model + H->int_section);
  const int nq = H->nq, nv = H->nv, nvar = 1 + 3*nv;
  const long long total = (long long)nstate * nvar;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long s = first + e / nvar;
    const int v = (int)(e % nvar);
    for (int k = 0; k < nq; k++) xqpos[(size_t)k * stride_out + e] = qpos[(size_t)k * stride_in + s];
    for (int k = 0; k < nv; k++) {
      xqvel[(size_t)k * stride_out + e] = qvel[(size_t)k * stride_in + s];
      xqacc[(size_t)k * stride_out + e] = qacc[(size_t)k * stride_in + s];
    }
    if (v == 0) continue;
    if (v <= nv) { xqacc[(size_t)(v - 1) * stride_out + e] += eps; continue; }
    if (v <= 2*nv) { xqvel[(size_t)(v - nv - 1) * stride_out + e] += eps; continue; }
    // mj_integratePos(m, qpos, unit vector of dof i, eps)  (engine_support.c:1518-1548)
    const int i = v - 2*nv - 1;
    const int j = (I + H->ioff[MJB_I_dof_jntid])[i];
    const int jt = (I + H->ioff[MJB_I_jnt_type])[j];
    int padr = (I + H->ioff[MJB_I_jnt_qposadr])[j];
    int r = i - (I + H->ioff[MJB_I_jnt_dofadr])[j];
    if (jt == MJB_JNT_HINGE || jt == MJB_JNT_SLIDE || (jt == MJB_JNT_FREE && r < 3)) {
      xqpos[(size_t)(padr + r) * stride_out + e] += eps * 1.0;
      continue;
    }
    if (jt == MJB_JNT_FREE) { padr += 3; r -= 3; }
    // mju_quatIntegrate (engine_util_spatial.c:241): quat <- normalize(quat) * quat(axis e_r, eps)
    double q[4], axis[3] = {0, 0, 0}, qrot[4], res[4];
    for (int k = 0; k < 4; k++) q[k] = xqpos[(size_t)(padr + k) * stride_out + e];
    axis[r] = 1.0;
    const double angle = eps * normalize3(axis);
    double sn, cs;
    sincos(angle*0.5, &sn, &cs);
    qrot[0] = cs; qrot[1] = axis[0]*sn; qrot[2] = axis[1]*sn; qrot[3] = axis[2]*sn;
    normalize4(q);
    mulQuat(res, q, qrot);
    for (int k = 0; k < 4; k++) xqpos[(size_t)(padr + k) * stride_out + e] = res[k];
  }
}

cudaError_t launch_fd_expand(const unsigned char* model, const double* qpos, const double* qvel,
                             const double* qacc, long long stride_in, long long first, int nstate,
                             double eps, double* xqpos, double* xqvel, double* xqacc,
                             long long stride_out, cudaStream_t stream) {
  if (nstate <= 0) return cudaSuccess;
  fd_expand_kernel<<<kSMs * 8, 256, 0, stream>>>(model, qpos, qvel, qacc, stride_in, first, nstate, eps,
                                                 xqpos, xqvel, xqacc, stride_out);
  return cudaGetLastError();
}

__global__ void fd_diff_kernel(const double* __restrict__ f, long long stride, int nstate, int nvar,
                               int v0, int nrow, int ncol, double inv_eps, double* out) {
  const long long total = (long long)nstate * nrow * ncol;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % ncol);
    const long long sr = t / ncol;
    const int r = (int)(sr % nrow);
    const long long s = sr / nrow;
    const double base = f[(size_t)k * stride + s * nvar];
    const double plus = f[(size_t)k * stride + s * nvar + v0 + r];
    out[t] = inv_eps * (plus - base);                       // diff(), engine_derivative_fd.c:48
  }
}

cudaError_t launch_fd_diff(const double* f, long long stride, int nstate, int nvar, int v0, int nrow,
                           int ncol, double eps, double* out, cudaStream_t stream) {
  if (nstate <= 0 || nrow <= 0 || ncol <= 0) return cudaSuccess;
  fd_diff_kernel<<<kSMs * 8, 256, 0, stream>>>(f, stride, nstate, nvar, v0, nrow, ncol, 1/eps, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// number of states whose status word is non-zero (return value of mjb_inverse)

__global__ void count_nonzero_kernel(const int* __restrict__ status, int n, int* counter) {
  int local = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    local += status[i] != 0;
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(counter, local);
}

cudaError_t launch_count_nonzero(const int* status, int n, int* counter, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  int grid = (n + 255) / 256;
  if (grid > kSMs * 8) grid = kSMs * 8;
  count_nonzero_kernel<<<grid, 256, 0, stream>>>(status, n, counter);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// FP64 FMA peak probe: 8 independent DFMA chains per thread, enough warps to fill every SM

__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, b = 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
      a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

cudaError_t dfma_peak_probe(int iters, float* ms, double* flops, cudaStream_t stream) {
  const int blocks = kSMs * 8, threads = 256;
  double* out = nullptr;
  cudaError_t e = cudaMalloc(&out, sizeof(double) * blocks * threads);
  if (e != cudaSuccess) return e;
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  dfma_probe_kernel<<<blocks, threads, 0, stream>>>(out, 16, 1.0);   // warm-up
  cudaEventRecord(t0, stream);
  dfma_probe_kernel<<<blocks, threads, 0, stream>>>(out, iters, 1.0);
  cudaEventRecord(t1, stream);
  e = cudaEventSynchronize(t1);
  if (e == cudaSuccess) {
    cudaEventElapsedTime(ms, t0, t1);
    *flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(out);
  return e;
}

}  // namespace mjb
