// C-ABI entry points of the engine (include/pi05.h): lifecycle, parameter binding, forward/backward/decode, taps.
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/pi05.h"
#include "engine.h"
#include "errors.h"
#include "launch.h"

using pi05::Engine;

namespace pi05 {
int engine_init_tables(Engine& e, cudaStream_t st);
}  // namespace pi05

static Engine* E(pi05_engine* p) { return reinterpret_cast<Engine*>(p); }

static int validate(const pi05_config* c, char* err, int n) {
  if (!c) {
    snprintf(err, n, "null config");
    return 1;
  }
  const int w[] = {c->paligemma.width, c->expert.width, c->vit_width};
  for (int x : w)
    if (x <= 0 || x % 8 != 0) {
      snprintf(err, n, "widths must be positive multiples of 8 (got %d)", x);
      return 1;
    }
  if (c->paligemma.head_dim % 16 != 0 || c->paligemma.head_dim <= 0) {
    snprintf(err, n, "head_dim must be a multiple of 16");
    return 1;
  }
  if (c->vit_width % c->vit_heads != 0 || (c->vit_width / c->vit_heads) % 8 != 0) {
    snprintf(err, n, "vit head_dim must be a multiple of 8");
    return 1;
  }
  if (c->image_size % c->vit_patch != 0 || c->max_batch <= 0 || c->num_images <= 0 || c->action_horizon <= 0 ||
      c->paligemma.mlp_dim % 8 != 0 || c->expert.mlp_dim % 8 != 0 || c->vit_mlp_dim % 8 != 0) {
    snprintf(err, n, "bad geometry (image/patch, batch, images, horizon or mlp dims)");
    return 1;
  }
  return 0;
}

extern "C" {

size_t pi05_workspace_bytes(const pi05_config* cfg) {
  char err[256];
  if (validate(cfg, err, sizeof(err)) != 0) {
    pi05::set_error(err);
    return 0;
  }
  Engine tmp;
  tmp.cfg = *cfg;
  pi05::engine_plan(tmp, /*dry=*/true);
  return tmp.arena.off + 4096;
}

int pi05_create(const pi05_config* cfg, int device, void* workspace, size_t workspace_bytes, pi05_engine** out) {
  char err[256];
  if (validate(cfg, err, sizeof(err)) != 0 || !out || !workspace) {
    pi05::set_error(cfg && out && workspace ? err : "pi05_create: null argument");
    return 1;
  }
  cudaError_t ce = cudaSetDevice(device);
  if (ce != cudaSuccess) {
    pi05::set_error(cudaGetErrorString(ce));
    return 2;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10) {
    snprintf(err, sizeof(err), "pi05 engine requires an sm_100 (B200) device; device %d is sm_%d0", device, major);
    pi05::set_error(err);
    return 2;
  }
  Engine* e = new (std::nothrow) Engine();
  if (!e) {
    pi05::set_error("out of host memory");
    return 3;
  }
  e->cfg = *cfg;
  e->device = device;
  e->arena.base = static_cast<char*>(workspace);
  e->arena.cap = workspace_bytes;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
    pi05::set_error("workspace must be 256B aligned");
    delete e;
    return 1;
  }
  int rc = pi05::engine_plan(*e, /*dry=*/false);
  if (rc == 0) rc = pi05::engine_init_tables(*e, nullptr);
  if (rc != 0) {
    delete e;
    return rc;
  }
  *out = reinterpret_cast<pi05_engine*>(e);
  return 0;
}

void pi05_destroy(pi05_engine* e) {
  if (e) pi05::exchange_destroy(*E(e));
  delete E(e);
}

int pi05_set_grad_exchange(pi05_engine* pe, void* nccl_comm, int32_t nranks, int32_t average_in_place, int32_t overlap) {
  Engine* e = E(pe);
  if (!e) {
    pi05::set_error("pi05_set_grad_exchange: null engine");
    return 1;
  }
  return pi05::exchange_setup(*e, nccl_comm, nranks, average_in_place, overlap);
}

int pi05_allreduce_grads(pi05_engine* pe, void* nccl_comm, int32_t nranks, int32_t average, void* stream) {
  Engine* e = E(pe);
  if (!e) {
    pi05::set_error("pi05_allreduce_grads: null engine");
    return 1;
  }
  return pi05::exchange_all(*e, nccl_comm, nranks, average, static_cast<cudaStream_t>(stream));
}

int pi05_grad_exchange_stats(pi05_engine* pe, int64_t* calls, int64_t* bytes) {
  Engine* e = E(pe);
  if (!e) {
    pi05::set_error("pi05_grad_exchange_stats: null engine");
    return 1;
  }
  if (calls) *calls = e->xch.calls;
  if (bytes) *bytes = e->xch.bytes;
  return 0;
}

int pi05_bind_params(pi05_engine* pe, const pi05_param* params, int n) {
  Engine* e = E(pe);
  if (!e || (!params && n > 0)) {
    pi05::set_error("pi05_bind_params: null argument");
    return 1;
  }
  e->params.clear();
  for (int i = 0; i < n; ++i) {
    pi05::PRef r;
    r.data = params[i].data;
    r.grad = params[i].grad;
    r.dtype = params[i].dtype;
    r.numel = params[i].numel;
    e->params[params[i].name] = r;
  }
  return pi05::engine_resolve_params(*e);
}

int pi05_params_updated(pi05_engine* pe, void* stream) {
  (void)stream;
  return E(pe) ? 0 : 1;
}

int pi05_set_taps(pi05_engine* pe, int enabled) {
  if (!E(pe)) return 1;
  E(pe)->taps_enabled = enabled != 0;
  return 0;
}

int pi05_forward(pi05_engine* pe, const pi05_batch* b, const float* actions, const float* noise, const float* time,
                 float* loss_out, void* stream) {
  if (!E(pe) || !b || !actions || !noise || !time || !loss_out) {
    pi05::set_error("pi05_forward: null argument");
    return 1;
  }
  return pi05::engine_forward(*E(pe), b, actions, noise, time, loss_out, static_cast<cudaStream_t>(stream));
}

int pi05_backward(pi05_engine* pe, const float* dloss, void* stream) {
  if (!E(pe) || !dloss) {
    pi05::set_error("pi05_backward: null argument");
    return 1;
  }
  return pi05::engine_backward(*E(pe), dloss, static_cast<cudaStream_t>(stream));
}

int pi05_prefill(pi05_engine* pe, const pi05_batch* b, void* stream) {
  if (!E(pe) || !b) {
    pi05::set_error("pi05_prefill: null argument");
    return 1;
  }
  return pi05::engine_prefill(*E(pe), b, static_cast<cudaStream_t>(stream));
}

int pi05_denoise(pi05_engine* pe, const float* noise, int num_steps, float* actions_out, void* stream) {
  if (!E(pe) || !noise || !actions_out || num_steps <= 0) {
    pi05::set_error("pi05_denoise: bad argument");
    return 1;
  }
  return pi05::engine_denoise(*E(pe), noise, num_steps, actions_out, static_cast<cudaStream_t>(stream));
}

int pi05_forward_advantage(pi05_engine* pe, const pi05_batch* b, const float* actions, const float* noise,
                           const float* time, const float* progress, float w_action, float w_value, float* loss_out,
                           float* aux_out, void* stream) {
  if (!E(pe) || !b || !actions || !noise || !time || !progress || !loss_out) {
    pi05::set_error("pi05_forward_advantage: null argument");
    return 1;
  }
  return pi05::engine_forward_advantage(*E(pe), b, actions, noise, time, progress, w_action, w_value, loss_out, aux_out,
                                        static_cast<cudaStream_t>(stream));
}

int pi05_forward_value(pi05_engine* pe, const pi05_batch* b, const float* noise, const float* time, float* value_out,
                       void* stream) {
  if (!E(pe) || !b || !noise || !time || !value_out) {
    pi05::set_error("pi05_forward_value: null argument");
    return 1;
  }
  return pi05::engine_value(*E(pe), b, noise, time, value_out, static_cast<cudaStream_t>(stream));
}

int pi05_debug_set_pdl(int enabled) {
  pi05::pdl_state() = enabled ? 1 : 0;
  return 0;
}

int pi05_debug_profile_layer(pi05_engine* pe, int layer) {
  if (!E(pe)) return 1;
  E(pe)->profile_layer = layer;
  return 0;
}

size_t pi05_preprocess_scratch_floats(int32_t batch, int32_t out_size) {
  if (batch <= 0 || out_size <= 0) return 0;
  return pi05::preprocess_scratch_floats(batch, out_size);
}

int pi05_preprocess_image(const float* image, int32_t height, int32_t width, int32_t channels_last, int32_t batch,
                          int32_t out_size, int32_t train, int32_t geometric, const float* params, float* scratch,
                          float* out, void* stream) {
  if (!image || !out || !scratch || batch <= 0 || height <= 0 || width <= 0 || out_size <= 1) {
    pi05::set_error("pi05_preprocess_image: bad argument");
    return 1;
  }
  if (train && !params) {
    pi05::set_error("pi05_preprocess_image: train != 0 needs the 6 augmentation parameters");
    return 1;
  }
  pi05::preprocess_image(image, height, width, channels_last, batch, out_size, train, geometric, params, scratch, out,
                         static_cast<cudaStream_t>(stream));
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    pi05::set_error(cudaGetErrorString(ce));
    return 9;
  }
  return 0;
}

int pi05_denoise_rtc(pi05_engine* pe, const float* noise, int32_t num_steps, const float* prev_chunk,
                     const float* time_weights, const float* dim_mask, const float* guidance, int32_t mask_rows,
                     int32_t provided, float* actions_out, void* stream) {
  Engine* e = E(pe);
  if (!e || !noise || !actions_out) {
    pi05::set_error("pi05_denoise_rtc: null argument");
    return 1;
  }
  return pi05::engine_denoise_rtc(*e, noise, num_steps, prev_chunk, time_weights, dim_mask, guidance, mask_rows, provided,
                                  actions_out, static_cast<cudaStream_t>(stream));
}

int32_t pi05_patch_row_kp(int32_t patch) { return pi05::patch_row_kp(patch); }

int pi05_preprocess_patches(const void* image, int32_t image_dtype, int32_t height, int32_t width, int32_t channels_last,
                            int32_t batch, int32_t out_size, int32_t patch, int32_t train, int32_t geometric,
                            const float* params, float* scratch, void* rows, void* stream) {
  if (!image || !rows || !scratch || batch <= 0 || height <= 0 || width <= 0 || out_size <= 1 || patch <= 0 ||
      out_size % patch != 0 || (image_dtype != PI05_F32 && image_dtype != PI05_U8)) {
    pi05::set_error("pi05_preprocess_patches: bad argument (image dtype must be PI05_F32 or PI05_U8, out_size % patch == 0)");
    return 1;
  }
  if (train && !params) {
    pi05::set_error("pi05_preprocess_patches: train != 0 needs the 6 augmentation parameters");
    return 1;
  }
  pi05::preprocess_patches(image, image_dtype == PI05_U8 ? 1 : 0, height, width, channels_last, batch, out_size, patch, train,
                           geometric, params, scratch, static_cast<pi05::bf16*>(rows), static_cast<cudaStream_t>(stream));
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    pi05::set_error(cudaGetErrorString(ce));
    return 9;
  }
  return 0;
}

int pi05_get_tap(pi05_engine* pe, const char* name, void* dst, int64_t* numel, int32_t* dtype, void* stream) {
  Engine* e = E(pe);
  if (!e || !name) {
    pi05::set_error("pi05_get_tap: null argument");
    return 1;
  }
  auto it = e->taps.find(name);
  if (it == e->taps.end()) {
    char err[256];
    snprintf(err, sizeof(err), "pi05_get_tap: no tap named '%s' (taps enabled: %d)", name, e->taps_enabled ? 1 : 0);
    pi05::set_error(err);
    return 2;
  }
  if (numel) *numel = it->second.numel;
  if (dtype) *dtype = it->second.dtype;
  if (dst) {
    const int dtc = it->second.dtype;
    const size_t bytes = static_cast<size_t>(it->second.numel) * (dtc == PI05_BF16 ? 2 : (dtc == PI05_U8 ? 1 : 4));
    cudaError_t ce =
        cudaMemcpyAsync(dst, it->second.ptr, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
    if (ce != cudaSuccess) {
      pi05::set_error(cudaGetErrorString(ce));
      return 3;
    }
  }
  return 0;
}

}  // extern "C"
