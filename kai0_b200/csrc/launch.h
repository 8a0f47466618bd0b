// Kernel launch helper with programmatic dependent launch (PDL): the kernel is enqueued with
// cudaLaunchAttributeProgrammaticStreamSerialization so that its CTAs may become resident and run their prologue
// while the previous kernel of the stream drains.  CONTRACT: a kernel launched through launch_pdl() must execute
// pdl_wait() (common.cuh) before its first access to global memory; it should call pdl_launch_dependents() right
// after so that the next kernel can be staged as early as possible.  PI05_PDL=0 in the environment disables the
// attribute (plain stream order) for A/B measurements.  Works inside CUDA-graph stream capture (programmatic edges).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace pi05 {

// -1: not decided yet (read PI05_PDL on first use); 0 / 1: forced (pi05_debug_set_pdl, used by profiling runs because a
// staged kernel's measured duration includes the time it waits for its predecessor)
int& pdl_state();
inline bool pdl_enabled() {
  int& s = pdl_state();
  if (s < 0) {
    const char* v = getenv("PI05_PDL");
    s = (v != nullptr && v[0] == '0') ? 0 : 1;
  }
  return s != 0;
}

template <class... P, class... A>
inline cudaError_t launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(std::forward<A>(args))...);
}

}  // namespace pi05
