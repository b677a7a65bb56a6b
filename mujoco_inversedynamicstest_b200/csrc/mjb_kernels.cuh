// Launch interface between the C-ABI host code (mjb_api.cu) and the sm_100a kernels
// (mjb_kernels.cu). Plain pointers and sizes only.
#ifndef MJB_KERNELS_CUH_
#define MJB_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "mjb_launch.h"

namespace mjb {

// phase ids for the optional per-kernel timing
enum { kPhaseSmooth = 0, kPhaseInertia, kPhaseScan, kPhaseContact, kPhaseBackward, kPhaseDiscrete, kPhaseTree, kPhaseCount };

// optional per-kernel timing: launch_inverse records a CUDA event before and after every phase
// kernel on the launching stream (events come from the host-side pool below)
struct PhaseTimer {
  cudaEvent_t (*next_event)(void* ctx);      // returns a fresh event from the owner's pool
  void (*mark)(void* ctx, int phase, cudaEvent_t begin, cudaEvent_t end);
  void* ctx;
};

// launches the phase kernels (smooth, [inertia], [contact], backward) for one chunk on `stream`;
// *launches receives the number of kernels launched
// `spec` (or null): kernels specialised for this model (mjb_jit.h); a phase whose handle is set runs
// the specialised kernel, the others the generic ones -- both work on the same scratch layout
struct SpecKernels;
cudaError_t launch_inverse(const LaunchArgs& args, cudaStream_t stream, int* launches,
                           const PhaseTimer* timer = nullptr, const SpecKernels* spec = nullptr);

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

// states per CTA of the warp-per-state candidate scan for a model with `ncand` candidate pairs and
// `ngeom` geoms; 0: the thread-per-state scan is used (short lists, or positions do not fit shared memory)
int scan_wide_states(int ncand, int ngeom);
size_t scan_wide_smem_bytes(int ngeom, int states_per_cta);
int scan_wide_buf_cap(int ncand);                     // ints per warp of the stage-1 -> stage-2 buffer
long long scan_wide_buf_ints(int ncand, int ngeom);   // ... and in total for the launch's fixed grid

// the sub-warp mj_crb / mj_factorM kernel keeps cdof, crb and the sparse matrix of 32 states in shared memory
bool inertia_subwarp_fits(int nv, int nbody, int nC);

}  // namespace mjb

#endif  // MJB_KERNELS_CUH_
