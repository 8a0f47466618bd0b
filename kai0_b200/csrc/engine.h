// Engine state for the pi0.5 forward / backward / decode path (see engine.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/pi05.h"
#include "exchange.h"
#include "kernels.h"

namespace pi05 {

struct PRef {
  void* data = nullptr;
  void* grad = nullptr;
  int dtype = 0;
  int64_t numel = 0;
  template <class T>
  T* d() const { return static_cast<T*>(data); }
  template <class T>
  T* g() const { return static_cast<T*>(grad); }
};

struct Arena {
  char* base = nullptr;
  size_t off = 0;
  size_t cap = 0;
  bool dry = true;
  bool overflow = false;
  void* alloc(size_t bytes) {
    off = (off + 255) & ~static_cast<size_t>(255);
    void* p = dry ? nullptr : static_cast<void*>(base + off);
    off += bytes;
    if (!dry && off > cap) overflow = true;
    return p;
  }
  template <class T>
  T* get(size_t n) { return static_cast<T*>(alloc(n * sizeof(T))); }
};

struct VitLayerP {
  PRef ln1_w, ln1_b, q_w, k_w, v_w, q_b, k_b, v_b, out_w, out_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b;
};
struct VitLayerA {
  bf16 *x_in, *h1, *qkv, *P, *attn, *x_mid, *h2, *pre, *act, *x_out;
  float *mean1, *rstd1, *mean2, *rstd2;
};

struct GemmaLayerP {
  PRef q_w, k_w, v_w, o_w, gate_w, up_w, down_w;
  PRef in_w, post_w;                          // plain RMSNorm weights (fp32)
  PRef in_dw, in_db, post_dw, post_db;        // adaptive dense (fp32)
};
struct GemmaLayerA {  // per stream
  bf16 *x_in, *n1, *qkv, *Q, *P, *O, *o_lin, *x_mid, *n2, *GU, *Hh, *d_lin, *x_out;
  bf16 *gate1, *gate2;
  float *rstd1, *rstd2;
};

struct Tap {
  const void* ptr;
  int64_t numel;
  int dtype;
};

struct Engine {
  pi05_config cfg{};
  int device = 0;
  // derived sizes
  int T = 0, NI = 0, L = 0, P = 0, A = 0, S = 0, Spad = 0, Ppad = 0;
  int Lmax = 0;  // cfg.max_token_len: buffers are planned for it; L / P / S / Ppad / Spad are those of the CURRENT batch
  int D = 0, E = 0, W = 0, H = 0, hd = 0, VH = 0, vhd = 0, Bmax = 0;
  bool train = false;

  std::map<std::string, PRef> params;
  bool bound = false;
  // resolved parameters
  PRef patch_w, patch_b, pos_emb, post_ln_w, post_ln_b, proj_w, proj_b, embed;
  std::vector<VitLayerP> vit;
  std::vector<GemmaLayerP> pg, ex;  // paligemma LM, expert
  PRef pg_norm_w, ex_norm_dw, ex_norm_db;
  PRef ain_w, ain_b, aout_w, aout_b, tin_w, tin_b, tout_w, tout_b;
  PRef vh0_w, vh0_b, vh2_w, vh2_b, vh4_w, vh4_b;

  // workspace
  Arena arena;
  // constant tables
  bf16 *rope_cos = nullptr, *rope_sin = nullptr;  // [S+1, hd/2]
  double* time_scaling = nullptr;                  // [E/2]
  // per-call state
  int B = 0;                 // batch of the last forward / prefill
  const pi05_batch* last_batch = nullptr;
  pi05_batch batch_copy{};
  uint8_t* pad = nullptr;    // [B, P]
  int *pos = nullptr, *nvalid = nullptr;
  // vision
  bf16* vit_x0 = nullptr;
  float *rtc_tap_v = nullptr, *rtc_tap_j = nullptr;  // copies of step 0's velocity / VJP for the parity taps (cfg.rtc)
  bf16* patch_wb = nullptr;     // patch-row path: Conv2d weight split [W, 3*Kp] = [hi | hi | lo]
  float* g_patch_dw = nullptr;  // its weight-gradient GEMM output [W, 2*Kp] fp32 before the fold
  int Kp = 0;
  std::vector<VitLayerA> va;
  bf16* vit_post = nullptr;
  float *vit_post_mean = nullptr, *vit_post_rstd = nullptr;
  // streams
  std::vector<GemmaLayerA> a1, a2;
  bf16 *Kc = nullptr, *Vc = nullptr;  // per layer [depth][B, S, hd] (train) or KV cache (decode)
  std::vector<bf16*> Kl, Vl;
  bf16 *prefix_out = nullptr, *suffix_out = nullptr;
  float* rstd_f1 = nullptr;
  float* rstd_f2 = nullptr;
  // suffix front-end (fp32)
  float *x_t = nullptr, *u_t = nullptr, *temb = nullptr, *aemb32 = nullptr, *t1 = nullptr, *t1s = nullptr, *t2 = nullptr,
        *cond = nullptr, *mods = nullptr, *so32 = nullptr, *v_t = nullptr, *timevec = nullptr;
  // backward scratch
  bf16 *g_x1 = nullptr, *g_x1b = nullptr, *g_x2 = nullptr, *g_x2b = nullptr, *g_big = nullptr, *g_big2 = nullptr,
       *g_t1 = nullptr, *g_t2 = nullptr, *g_t3 = nullptr, *g_P = nullptr, *g2_do = nullptr, *g2_big = nullptr,
       *g2_big2 = nullptr, *g2_t1 = nullptr, *g2_t2 = nullptr, *g2_t3 = nullptr;
  float *g_dK = nullptr, *g_dV = nullptr, *g_dmods = nullptr, *g_f32a = nullptr, *g_f32b = nullptr, *g_f32c = nullptr,
        *g_acc = nullptr, *g_embed_scratch = nullptr;
  int* g_first = nullptr;
  size_t g_acc_elems = 0;

  // adaRMS dense layers laid out with one uniform stride (fp32 arena order) -> batched single-launch linears
  bool ada_uniform = false, ada_uniform_grad = false;
  int64_t ada_wstride = 0, ada_bstride = 0;
  float* g_dcond_part = nullptr;
  // decode: per-call tables of the time conditioning (depends only on the step index, not on x_t)
  static constexpr int kMaxDecodeSteps = 64;
  float *dec_times = nullptr, *dec_temb = nullptr, *dec_t1 = nullptr, *dec_t1s = nullptr, *dec_t2 = nullptr,
        *dec_cond = nullptr, *dec_mods = nullptr;
  // AdvantageEstimator head (cfg.value_head): fp32 activations of tanh(MLP3(suffix_out[:,0])) and the loss pieces
  float *vh_in = nullptr, *vh_h1 = nullptr, *vh_s1 = nullptr, *vh_h2 = nullptr, *vh_s2 = nullptr, *vh_h3 = nullptr,
        *vh_val = nullptr, *vh_prog = nullptr, *vh_la = nullptr, *vh_lv = nullptr;
  bool adv_mode = false;  // last forward was pi05_forward_advantage
  float w_action = 1.0f, w_value = 0.0f;
  float* splitk_ws = nullptr;  // fp32 scratch of the small-M split-K GEMM path
  size_t splitk_ws_bytes = 0;
  GradExchange xch;  // data-parallel gradient exchange overlapped with backward (exchange.h)
  bool taps_enabled = false;
  int profile_layer = -1;  // pi05_debug_profile_layer
  std::map<std::string, Tap> taps;
  cudaStream_t stream = nullptr;
  char err[1024] = "";
};

// engine.cu
int engine_plan(Engine& e, bool dry);
int engine_set_token_len(Engine& e, int token_len, const char* who);
int engine_resolve_params(Engine& e);
int engine_forward(Engine& e, const pi05_batch* b, const float* actions, const float* noise, const float* time,
                   float* loss_out, cudaStream_t st);
int engine_forward_advantage(Engine& e, const pi05_batch* b, const float* actions, const float* noise, const float* time,
                             const float* progress, float w_action, float w_value, float* loss_out, float* aux_out,
                             cudaStream_t st);
int engine_value(Engine& e, const pi05_batch* b, const float* noise, const float* time, float* value_out,
                 cudaStream_t st);
// engine_bwd.cu
int engine_backward(Engine& e, const float* dloss, cudaStream_t st);
// engine_decode.cu
int engine_prefill(Engine& e, const pi05_batch* b, cudaStream_t st);
int engine_denoise(Engine& e, const float* noise, int num_steps, float* actions_out, cudaStream_t st);
// engine_rtc.cu
int engine_denoise_rtc(Engine& e, const float* noise, int num_steps, const float* prev, const float* time_weights,
                       const float* dim_mask, const float* guidance, int mask_rows, int provided, float* actions_out,
                       cudaStream_t st);
void add_tap(Engine& e, const char* name, const void* p, int64_t n, int dtype);

}  // namespace pi05
