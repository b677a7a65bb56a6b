// Launch interface between the C-ABI host code (mjb_api.cu) and the sm_100a kernels
// (mjb_kernels.cu). Plain pointers and sizes only.
#ifndef MJB_KERNELS_CUH_
#define MJB_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

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
struct ContactQueue { int nitems; int ncontacts; int overflow; int nslots; };
struct ContactItem { int state; int ci; };          // chunk-local state, candidate pair
struct ItemCon { int base; int count; };            // the item's contacts: contacts[base .. base+count)
struct ContactRec {
  int state, ci, k, efc_address;                    // k: contact index in its state; efc_address or -1
  double dist, pos[3], frame[6];                    // normal, tangent (third axis is their cross product)
};


struct LaunchArgs {
  const unsigned char* model;   // device blob (mjbHdr + sections)
  int model_bytes;
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
  int* slot_rec;                // slot (state's first slot + k) -> index into contacts
  int items_cap, contacts_cap;
  int has_contacts;             // run the contact kernel (ncand > 0 and contacts enabled)
  int has_spatial;              // mjbHdr::has_spatial (force-carrying spatial tendons: smooth kernel variant)
  int has_gravcomp;             // mjbHdr::has_gravcomp (selects the backward kernel instantiation)
  int max_pair_contacts;        // mjbHdr::max_pair_contacts (sizes the per-warp contact pool)
  int skip_sensors;             // mj_inverseSkip(skipsensor = 1): leave sensordata as it is
  Outputs out;
};

// phase ids for the optional per-kernel timing
enum { kPhaseSmooth = 0, kPhaseInertia, kPhaseScan, kPhaseContact, kPhaseBackward, kPhaseDiscrete, kPhaseCount };

// optional per-kernel timing: launch_inverse records a CUDA event before and after every phase
// kernel on the launching stream (events come from the host-side pool below)
struct PhaseTimer {
  cudaEvent_t (*next_event)(void* ctx);      // returns a fresh event from the owner's pool
  void (*mark)(void* ctx, int phase, cudaEvent_t begin, cudaEvent_t end);
  void* ctx;
};

// launches the phase kernels (smooth, [inertia], [contact], backward) for one chunk on `stream`;
// *launches receives the number of kernels launched
cudaError_t launch_inverse(const LaunchArgs& args, cudaStream_t stream, int* launches,
                           const PhaseTimer* timer = nullptr);

// AoS [n][rows] (host layout, as in looping mju_copy into d->qpos) <-> SoA [rows][stride]
cudaError_t launch_aos_to_soa(const double* aos, double* soa, int n, int rows, long long stride,
                              cudaStream_t stream);
cudaError_t launch_soa_to_aos(const double* soa, double* aos, int n, int rows, long long stride,
                              cudaStream_t stream);
cudaError_t launch_soa_to_aos_int(const int* soa, int* aos, int n, int rows, long long stride,
                                  cudaStream_t stream);

// mjd_inverseFD support (engine_derivative_fd.c:611): variant v of state s goes to column
// s*(1+3nv) + v of the expanded batch: v = 0 unperturbed, 1..nv qacc[i] += eps, nv+1..2nv
// qvel[i] += eps, 2nv+1..3nv qpos integrated by eps along dof i (mj_integratePos)
cudaError_t launch_fd_expand(const unsigned char* model, const double* qpos, const double* qvel,
                             const double* qacc, long long stride_in, long long first, int nstate,
                             double eps, double* xqpos, double* xqvel, double* xqacc,
                             long long stride_out, cudaStream_t stream);
// out[(s*nrow + r)*ncol + k] = (f[k][s*nvar + v0 + r] - f[k][s*nvar]) / eps for r < nrow
cudaError_t launch_fd_diff(const double* f, long long stride, int nstate, int nvar, int v0, int nrow,
                           int ncol, double eps, double* out, cudaStream_t stream);

// *counter += number of non-zero entries of status[0..n)
cudaError_t launch_count_nonzero(const int* status, int n, int* counter, cudaStream_t stream);

// FP64 FMA peak probe: runs `iters` dependent-chain-free DFMA rounds on every thread, returns
// elapsed milliseconds in *ms and the number of flops executed in *flops.
cudaError_t dfma_peak_probe(int iters, float* ms, double* flops, cudaStream_t stream);

size_t inverse_smem_bytes(int model_bytes, int model_in_smem);

}  // namespace mjb

#endif  // MJB_KERNELS_CUH_
