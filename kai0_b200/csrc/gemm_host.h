// Host-side glue between the GEMM dispatcher (gemm_sm100.cu) and the 2-CTA kernel (gemm2_sm100.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "gemm_device.cuh"

namespace pi05 {
// Launches the cta_group::2 kernel (256 x 256 tile per CTA pair).  Tensor maps: A box {64, 128} (or 64x64 atoms),
// B box {64, 128} = the half of the B tile each CTA of the pair loads.
int launch_gemm2(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const gemm_detail::KParams& kp,
                 cudaStream_t stream, char* err, int err_len);
// Dynamic tile scheduler: every GEMM launch draws tile indices from its own pair of device words {next tile, clusters /
// CTAs finished}, taken round-robin from a per-device pool so that launches overlapping under programmatic dependent
// launch never share one.  Both words are zero whenever no kernel is using them: the last CTA (cluster) to run dry resets
// them, so the pool needs no host-side memset between launches (and is CUDA-graph safe: a captured launch keeps its pair).
unsigned int* next_sched_counter();
}  // namespace pi05
