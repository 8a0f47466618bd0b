// Data-parallel gradient exchange of the pi0.5 engine: chunked NCCL all-reduce overlapped with backward
// (replaces DistributedDataParallel's bucketed all-reduce, scripts/train_pytorch.py:440-447).  See include/pi05.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace pi05 {

struct GRange {  // one contiguous piece of a gradient arena
  char* lo = nullptr;
  char* hi = nullptr;
  int dtype = 0;  // PI05_F32 / PI05_BF16
  bool valid() const { return lo != nullptr && hi > lo; }
};

struct GradExchange {
  void* comm = nullptr;  // ncclComm_t (opaque)
  int nranks = 1;
  bool average_in_place = false;
  bool overlap = false;  // true: chunk by chunk under backward on `stream`; false: one exchange at the end of backward
  bool chunked = false;  // group ranges verified contiguous at bind time
  cudaStream_t stream = nullptr;       // engine-owned, highest priority
  std::vector<cudaEvent_t> events;     // ready[i]: producing kernels enqueued on the compute stream
  size_t ev_cursor = 0;
  cudaEvent_t done = nullptr;
  int64_t calls = 0, bytes = 0;        // of the current / last backward
  // ranges (filled by engine_resolve_params)
  std::vector<GRange> pg, ex, vit;
  GRange vtail, embed, f32_main, f32_vis, all_bf16, all_f32;
};

struct Engine;
int exchange_setup(Engine& e, void* comm, int nranks, int average_in_place, int overlap);
void exchange_destroy(Engine& e);
void exchange_begin(Engine& e);                                 // start of a backward
int exchange_range(Engine& e, const GRange& r);                 // "this range is final on e.stream"
int exchange_ranges(Engine& e, const GRange& a, const GRange& b);  // [a.lo, b.hi) when adjacent, else both
int exchange_finish(Engine& e);                                 // e.stream waits for the exchange stream
int exchange_all(Engine& e, void* comm, int nranks, int average, cudaStream_t st);  // one-shot, no overlap

}  // namespace pi05
