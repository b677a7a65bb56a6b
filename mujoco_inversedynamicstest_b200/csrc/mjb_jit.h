// Model specialisation: compiles csrc/mjb_spec_kernels.cuh for ONE model with NVRTC (sm_100a cubin),
// loads it through the CUDA driver API and hands the kernel handles to the launcher. Both libraries
// are resolved with dlopen at first use (libnvrtc.so.12, libcuda.so.1), so libmjb.so itself links
// only against the static CUDA runtime. Compiled cubins are cached on disk, keyed by a hash of
// (kernel sources, model blob, options, NVRTC version): <directory of libmjb.so>/jitcache or
// $MJB_JIT_CACHE.
#ifndef MJB_JIT_H_
#define MJB_JIT_H_

#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "mjb_launch.h"

namespace mjb {

struct SpecKernels {
  void* module = nullptr;        // CUmodule
  // CUfunction handles. smooth / inertia: one kernel per stage (body range), launched in order;
  // an empty list or a null handle means the generic kernel runs that phase.
  std::vector<void*> smooth;
  std::vector<void*> inertia;
  std::vector<void*> tree;       // trunk-forward / fused-subtree / trunk-backward stages (phase_tree): replace
                                 // smooth + inertia when the inertia outputs are requested
  void* contact_scan = nullptr;
  void* backward = nullptr;
  bool from_cache = false;
  double compile_seconds = 0;
  std::string key;               // cache key (hex)
};

// true when NVRTC and the driver API could be loaded; otherwise `why` says what is missing
bool jitAvailable(std::string& why);

// Compile (or fetch from the cache) and load the specialised kernels of the model blob on the
// current device. Returns false with a message when the model is not eligible or a step failed.
bool jitSpecialize(const std::vector<unsigned char>& blob, SpecKernels& out, std::string& err);

// compile only (no device needed): used by build() to fill the cache ahead of the GPU runs
bool jitCompileToCache(const std::vector<unsigned char>& blob, std::string& key, bool& cached,
                       double& seconds, std::string& err);

void jitUnload(SpecKernels& k);

// launch one specialised kernel (grid x 1 x 1, threads x 1 x 1) with the launch arguments by value
cudaError_t jitLaunch(void* fn, int grid, int threads, size_t smem, cudaStream_t stream,
                      const LaunchArgs& args);

// stage boundaries the prelude was generated with (for diagnostics): "smooth 1-5 5-9 ... | inertia ..."
std::string jitStagePlan(const std::vector<unsigned char>& blob);

}  // namespace mjb

#endif  // MJB_JIT_H_
