// Launch interface between the C-ABI host code (mjb_api.cu) and the sm_100a kernels
// (mjb_kernels.cu). Plain pointers and sizes only.
#ifndef MJB_KERNELS_CUH_
#define MJB_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "mjb_model.h"

namespace mjb {

// Resident-thread geometry of the main kernel. The per-thread scratch is laid out
// [slot][kNT] (structure of arrays over thread slots), so consecutive lanes touch consecutive
// doubles; kNT is a compile-time constant so that `slot + k` offsets fold into immediates.
constexpr int kThreads = 128;                 // threads per CTA (4 warps)
constexpr int kSMs = 148;                     // B200
constexpr int kCtasPerSM = 4;                 // scratch slots provisioned per SM
constexpr int kGrid = kSMs * kCtasPerSM;      // 592 CTAs
constexpr int kNT = kGrid * kThreads;         // 75,776 thread slots

}  // namespace mjb

#define MJB_NT mjb::kNT
#include "mjb_pipeline.h"

namespace mjb {

struct LaunchArgs {
  const unsigned char* model;   // device blob (mjbHdr + sections)
  int model_bytes;
  int model_in_smem;            // 1: stage the blob into shared memory with a TMA bulk copy
  const double* qpos;           // [nq][stride]
  const double* qvel;           // [nv][stride]
  const double* qacc;           // [nv][stride]
  double* scratch;              // [nscratch][kNT]
  long long stride;             // row stride of every state-indexed array (>= nbatch)
  int nbatch;
  int nconmax, njmax;
  Outputs out;
};

// launches the fused mj_inverse kernel on `stream`; returns cudaGetLastError()
cudaError_t launch_inverse(const LaunchArgs& args, cudaStream_t stream);

// AoS [n][rows] (host layout, as in looping mju_copy into d->qpos) <-> SoA [rows][stride]
cudaError_t launch_aos_to_soa(const double* aos, double* soa, int n, int rows, long long stride,
                              cudaStream_t stream);
cudaError_t launch_soa_to_aos(const double* soa, double* aos, int n, int rows, long long stride,
                              cudaStream_t stream);
cudaError_t launch_soa_to_aos_int(const int* soa, int* aos, int n, int rows, long long stride,
                                  cudaStream_t stream);

// *counter += number of non-zero entries of status[0..n)
cudaError_t launch_count_nonzero(const int* status, int n, int* counter, cudaStream_t stream);

// FP64 FMA peak probe: runs `iters` dependent-chain-free DFMA rounds on every thread, returns
// elapsed milliseconds in *ms and the number of flops executed in *flops.
cudaError_t dfma_peak_probe(int iters, float* ms, double* flops, cudaStream_t stream);

size_t inverse_smem_bytes(int model_bytes, int model_in_smem);

}  // namespace mjb

#endif  // MJB_KERNELS_CUH_
