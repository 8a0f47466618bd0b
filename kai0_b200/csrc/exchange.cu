// Gradient exchange: NCCL all-reduce of contiguous pieces of the flat gradient arenas on an engine-owned high-priority
// stream, enqueued from inside engine_backward as soon as each piece is final (include/pi05.h, exchange.h).
// libnccl is resolved with dlopen at first use: the process that created the ncclComm_t has it loaded already.
#include "exchange.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>

#if __has_include(<nccl.h>)
#include <nccl.h>  // types and NCCL_CONFIG_INITIALIZER only: every symbol is resolved with dlsym at run time
#define PI05_HAVE_NCCL_H 1
#else
#define PI05_HAVE_NCCL_H 0
#endif

#include "common.cuh"
#include "engine.h"
#include "errors.h"
#include "launch.h"

namespace pi05 {

namespace {

// NCCL 2.x ABI constants (nccl.h): ncclDataType_t {ncclFloat32 = 7, ncclBfloat16 = 9}, ncclRedOp_t {ncclSum = 0}
constexpr int kNcclFloat32 = 7, kNcclBfloat16 = 9, kNcclSum = 0;
using AllReduceFn = int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t);
using ErrStrFn = const char* (*)(int);
using CountFn = int (*)(void*, int*);

struct NcclApi {
  AllReduceFn all_reduce = nullptr;
  ErrStrFn err_str = nullptr;
  CountFn comm_count = nullptr;
  void* get_unique_id = nullptr;
  void* init_rank = nullptr;
  void* init_rank_config = nullptr;
  void* comm_destroy = nullptr;
  void* get_version = nullptr;
  bool tried = false;
};

NcclApi& nccl() {
  static NcclApi api;
  if (!api.tried) {
    api.tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the caller's communicator came from
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
      api.err_str = reinterpret_cast<ErrStrFn>(dlsym(h, "ncclGetErrorString"));
      api.comm_count = reinterpret_cast<CountFn>(dlsym(h, "ncclCommCount"));
      api.get_unique_id = dlsym(h, "ncclGetUniqueId");
      api.init_rank = dlsym(h, "ncclCommInitRank");
      api.init_rank_config = dlsym(h, "ncclCommInitRankConfig");
      api.comm_destroy = dlsym(h, "ncclCommDestroy");
      api.get_version = dlsym(h, "ncclGetVersion");
    }
  }
  return api;
}

__global__ void __launch_bounds__(256) scale_bf16_k(bf16* __restrict__ p, int64_t n, float s) {
  pdl_enter();
  const int64_t n8 = n / 8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v[8];
    load8(p + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= s;
    store8(p + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n8 * 8; i < n; ++i) p[i] = __float2bfloat16_rn(__bfloat162float(p[i]) * s);
}
__global__ void __launch_bounds__(256) scale_f32_k(float* __restrict__ p, int64_t n, float s) {
  pdl_enter();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] *= s;
}

int fail(Engine& e, const char* what, int nccl_rc) {
  const NcclApi& api = nccl();
  snprintf(e.err, sizeof(e.err), "gradient exchange: %s (%s)", what,
           nccl_rc != 0 && api.err_str ? api.err_str(nccl_rc) : "no NCCL status");
  set_error(e.err);
  return 10;
}

// all-reduce (+ optional in-place average) of one range on `st`
int reduce_on(Engine& e, const GRange& r, void* comm, int nranks, bool average, cudaStream_t st) {
  NcclApi& api = nccl();
  if (!api.all_reduce) return fail(e, "libnccl.so.2 could not be resolved (dlopen)", 0);
  const int esz = r.dtype == PI05_BF16 ? 2 : 4;
  const size_t count = static_cast<size_t>(r.hi - r.lo) / esz;
  const int rc = api.all_reduce(r.lo, r.lo, count, r.dtype == PI05_BF16 ? kNcclBfloat16 : kNcclFloat32, kNcclSum, comm, st);
  if (rc != 0) return fail(e, "ncclAllReduce failed", rc);
  if (average && nranks > 1) {
    const float s = 1.0f / static_cast<float>(nranks);
    const int64_t n = static_cast<int64_t>(count);
    int64_t blocks = (n / 8 + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    if (r.dtype == PI05_BF16)
      launch_pdl(scale_bf16_k, dim3(static_cast<int>(blocks)), dim3(256), 0, st, reinterpret_cast<bf16*>(r.lo), n, s);
    else
      launch_pdl(scale_f32_k, dim3(static_cast<int>(blocks)), dim3(256), 0, st, reinterpret_cast<float*>(r.lo), n, s);
    count_launch();
  }
  e.xch.calls += 1;
  e.xch.bytes += static_cast<int64_t>(r.hi - r.lo);
  return 0;
}

}  // namespace

int exchange_setup(Engine& e, void* comm, int nranks, int average_in_place, int overlap) {
  GradExchange& x = e.xch;
  if (comm == nullptr) {
    x.comm = nullptr;
    x.nranks = 1;
    return 0;
  }
  if (nranks < 1) {
    snprintf(e.err, sizeof(e.err), "pi05_set_grad_exchange: nranks %d", nranks);
    set_error(e.err);
    return 1;
  }
  NcclApi& api = nccl();
  if (!api.all_reduce) return fail(e, "libnccl.so.2 could not be resolved (dlopen)", 0);
  if (api.comm_count) {
    int n = 0;
    const int rc = api.comm_count(comm, &n);
    if (rc != 0) return fail(e, "ncclCommCount failed (not a live ncclComm_t?)", rc);
    if (n != nranks) {
      snprintf(e.err, sizeof(e.err), "pi05_set_grad_exchange: communicator has %d ranks, caller said %d", n, nranks);
      set_error(e.err);
      return 1;
    }
  }
  if (x.stream == nullptr) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = greatest priority (numerically lowest)
    if (cudaStreamCreateWithPriority(&x.stream, cudaStreamNonBlocking, hi) != cudaSuccess) return fail(e, "stream create", 0);
    x.events.resize(96);
    for (auto& ev : x.events) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&x.done, cudaEventDisableTiming);
  }
  x.comm = comm;
  x.nranks = nranks;
  x.average_in_place = average_in_place != 0;
  x.overlap = overlap != 0;
  return 0;
}

void exchange_destroy(Engine& e) {
  GradExchange& x = e.xch;
  if (x.stream != nullptr) {
    cudaStreamSynchronize(x.stream);
    for (auto& ev : x.events) cudaEventDestroy(ev);
    cudaEventDestroy(x.done);
    cudaStreamDestroy(x.stream);
    x.stream = nullptr;
    x.events.clear();
  }
}

void exchange_begin(Engine& e) {
  e.xch.calls = 0;
  e.xch.bytes = 0;
}

int exchange_range(Engine& e, const GRange& r) {
  GradExchange& x = e.xch;
  if (x.comm == nullptr || !x.overlap || !x.chunked || !r.valid()) return 0;
  cudaEvent_t ev = x.events[x.ev_cursor++ % x.events.size()];
  cudaEventRecord(ev, e.stream);        // everything that writes this range is enqueued before this point
  cudaStreamWaitEvent(x.stream, ev, 0);
  return reduce_on(e, r, x.comm, x.nranks, x.average_in_place, x.stream);
}

int exchange_ranges(Engine& e, const GRange& a, const GRange& b) {
  if (a.valid() && b.valid() && a.dtype == b.dtype) {
    const GRange& first = a.lo <= b.lo ? a : b;
    const GRange& second = a.lo <= b.lo ? b : a;
    if (second.lo >= first.hi && second.lo - first.hi <= 64) {  // adjacent up to alignment padding: one collective
      GRange m = first;
      m.hi = second.hi;
      return exchange_range(e, m);
    }
  }
  int rc = exchange_range(e, a);
  if (rc != 0) return rc;
  return exchange_range(e, b);
}

int exchange_finish(Engine& e) {
  GradExchange& x = e.xch;
  if (x.comm == nullptr) return 0;
  if (!x.overlap) {
    // one exchange of both arenas on the compute stream right behind backward, at NCCL's full speed.  On the power-capped
    // B200s this was measured FASTER than the overlapped mode at N = 2 and N = 8 (profiles/r02_exchange_overlap.md).
    int rc = x.all_bf16.valid() ? reduce_on(e, x.all_bf16, x.comm, x.nranks, x.average_in_place, e.stream) : 0;
    if (rc == 0 && x.all_f32.valid()) rc = reduce_on(e, x.all_f32, x.comm, x.nranks, x.average_in_place, e.stream);
    return rc;
  }
  if (!x.chunked) {  // ranges could not be verified at bind time: one exchange of both arenas behind backward
    cudaEvent_t ev = x.events[x.ev_cursor++ % x.events.size()];
    cudaEventRecord(ev, e.stream);
    cudaStreamWaitEvent(x.stream, ev, 0);
    int rc = x.all_bf16.valid() ? reduce_on(e, x.all_bf16, x.comm, x.nranks, x.average_in_place, x.stream) : 0;
    if (rc == 0 && x.all_f32.valid()) rc = reduce_on(e, x.all_f32, x.comm, x.nranks, x.average_in_place, x.stream);
    if (rc != 0) return rc;
  }
  cudaEventRecord(x.done, x.stream);
  cudaStreamWaitEvent(e.stream, x.done, 0);
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(e.err, sizeof(e.err), "gradient exchange: %s", cudaGetErrorString(ce));
    set_error(e.err);
    return 9;
  }
  return 0;
}

int exchange_all(Engine& e, void* comm, int nranks, int average, cudaStream_t st) {
  if (comm == nullptr || nranks < 1) {
    snprintf(e.err, sizeof(e.err), "pi05_allreduce_grads: null communicator / bad nranks");
    set_error(e.err);
    return 1;
  }
  GradExchange& x = e.xch;
  if (!x.all_bf16.valid() && !x.all_f32.valid()) {
    snprintf(e.err, sizeof(e.err), "pi05_allreduce_grads: gradient buffers not bound");
    set_error(e.err);
    return 8;
  }
  exchange_begin(e);
  int rc = x.all_bf16.valid() ? reduce_on(e, x.all_bf16, comm, nranks, average != 0, st) : 0;
  if (rc == 0 && x.all_f32.valid()) rc = reduce_on(e, x.all_f32, comm, nranks, average != 0, st);
  return rc;
}

}  // namespace pi05

// ---- communicator helpers for hosts that have no NCCL binding of their own (include/pi05.h) ------------------------
extern "C" int pi05_nccl_unique_id(void* out128) {
#if PI05_HAVE_NCCL_H
  using namespace pi05;
  NcclApi& api = nccl();
  if (!api.get_unique_id || !out128) {
    set_error("pi05_nccl_unique_id: libnccl.so.2 could not be resolved");
    return 1;
  }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  const int rc = reinterpret_cast<ncclResult_t (*)(ncclUniqueId*)>(api.get_unique_id)(static_cast<ncclUniqueId*>(out128));
  if (rc != 0) {
    set_error(api.err_str ? api.err_str(rc) : "ncclGetUniqueId failed");
    return 2;
  }
  return 0;
#else
  pi05::set_error("built without nccl.h");
  return 3;
#endif
}

extern "C" int pi05_nccl_comm_create(const void* unique_id128, int32_t nranks, int32_t rank, int32_t max_ctas,
                                     void** comm_out) {
#if PI05_HAVE_NCCL_H
  using namespace pi05;
  NcclApi& api = nccl();
  if (!api.init_rank || !unique_id128 || !comm_out) {
    set_error("pi05_nccl_comm_create: libnccl.so.2 could not be resolved / null argument");
    return 1;
  }
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  ncclComm_t comm = nullptr;
  int rc = -1;
  int runtime_version = 0;
  if (api.get_version) reinterpret_cast<ncclResult_t (*)(int*)>(api.get_version)(&runtime_version);
  if (max_ctas != 0 && api.init_rank_config && runtime_version >= NCCL_VERSION_CODE) {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    cfg.blocking = 1;
    if (max_ctas > 0) {
      // few CTAs: the exchange overlaps backward, its average demand is ~30 GB/s of the 900 GB/s NVLink 5 port
      cfg.minCTAs = 1;
      cfg.maxCTAs = max_ctas;
    } else {
      // exchange at the end of backward on idle SMs: at least -max_ctas CTAs (measured on 8 B200s, 6.95 GB payload:
      // 14.7 ms = 830 GB/s bus bandwidth with 32, tools/allreduce_probe.py)
      cfg.minCTAs = -max_ctas;
    }
    rc = reinterpret_cast<ncclResult_t (*)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*)>(api.init_rank_config)(
        &comm, nranks, id, rank, &cfg);
    if (rc != 0) {  // a configuration this NCCL build refuses must not cost the communicator: plain init
      comm = nullptr;
      rc = reinterpret_cast<ncclResult_t (*)(ncclComm_t*, int, ncclUniqueId, int)>(api.init_rank)(&comm, nranks, id, rank);
    }
  } else {
    rc = reinterpret_cast<ncclResult_t (*)(ncclComm_t*, int, ncclUniqueId, int)>(api.init_rank)(&comm, nranks, id, rank);
  }
  if (rc != 0) {
    char buf[256];
    snprintf(buf, sizeof(buf), "pi05_nccl_comm_create: %s (runtime NCCL %d, header %d)",
             api.err_str ? api.err_str(rc) : "ncclCommInitRank failed", runtime_version, NCCL_VERSION_CODE);
    set_error(buf);
    return 2;
  }
  *comm_out = comm;
  return 0;
#else
  pi05::set_error("built without nccl.h");
  return 3;
#endif
}

extern "C" int pi05_nccl_comm_destroy(void* comm) {
#if PI05_HAVE_NCCL_H
  using namespace pi05;
  NcclApi& api = nccl();
  if (!comm || !api.comm_destroy) return 0;
  return reinterpret_cast<ncclResult_t (*)(ncclComm_t)>(api.comm_destroy)(static_cast<ncclComm_t>(comm)) == 0 ? 0 : 2;
#else
  return 0;
#endif
}
