// Inference path: prefix pass that fills the per-layer KV cache, then the Euler denoise loop over the action
// expert only.  Follows pi0_pytorch.py:375-419 (sample_actions), :421-461 (denoise_step),
// gemma_pytorch.py:102-125 (single-stream branches) -> modeling_gemma.py:344-384,446-555.
#include <cmath>
#include <cstdio>

#include "engine.h"
#include "errors.h"
#include "gemm.h"

namespace pi05 {

#define CHECK_RC(x)           \
  do {                        \
    int _rc = (x);            \
    if (_rc != 0) return _rc; \
  } while (0)

int engine_gemm(Engine& e, const GemmArgs& a);
GemmArgs mk_gemm(int M, int N, int K, const void* A, int64_t lda, const void* Bm, int64_t ldb, void* D, int64_t ldd,
                 int epi);
int prefix_forward(Engine& e, const pi05_batch* b);
int suffix_frontend(Engine& e, const float* x_t, const float* time, int B);
void add_tap(Engine& e, const char* name, const void* p, int64_t n, int dtype);

// PaliGemma-only layer; K/V (post-RoPE K) go to the cache rows [0, P) (modeling_gemma.py:303-307).
static int prefix_layer(Engine& e, int l, int B, bool kv_only) {
  cudaStream_t st = e.stream;
  const pi05_config& c = e.cfg;
  const int P = e.P, S = e.S, D = e.D, H = e.H, hd = e.hd;
  const int M1 = B * P, QW = (H + 2) * hd;
  GemmaLayerA& p1 = e.a1[l];
  const GemmaLayerP& w1 = e.pg[l];
  bf16 *Kc = e.Kl[l], *Vc = e.Vl[l];
  rmsnorm_fwd(p1.x_in, w1.in_w.d<float>(), nullptr, 0, p1.n1, p1.rstd1, nullptr, M1, D, 1e-6f, st);
  CHECK_RC(engine_gemm(e, mk_gemm(M1, QW, D, p1.n1, D, w1.q_w.data, D, p1.qkv, QW, EPI_STORE)));
  rope_pack_fwd(p1.qkv, P, H, hd, e.pos, e.nvalid, 0, e.rope_cos, e.rope_sin, p1.Q, Kc, Vc, 0, S, B, st);
  if (kv_only) return 0;  // the last layer's output is never read by the decode loop
  {
    GemmArgs g = mk_gemm(P * H, P, hd, p1.Q, hd, Kc, hd, p1.P, e.Ppad, EPI_SCALE);
    g.batch = B;
    g.a_batch_stride = static_cast<int64_t>(P) * H * hd;
    g.b_batch_stride = static_cast<int64_t>(S) * hd;
    g.d_batch_stride = static_cast<int64_t>(P) * H * e.Ppad;
    g.scale = 1.0f / sqrtf(static_cast<float>(hd));
    CHECK_RC(engine_gemm(e, g));
  }
  softmax_fwd(p1.P, e.Ppad, P * H, B, P, P, e.pad, e.pad, H, st);
  {
    GemmArgs g = mk_gemm(P * H, hd, P, p1.P, e.Ppad, Vc, hd, p1.O, hd, EPI_STORE);
    g.b_major = 1;
    g.batch = B;
    g.a_batch_stride = static_cast<int64_t>(P) * H * e.Ppad;
    g.b_batch_stride = static_cast<int64_t>(S) * hd;
    g.d_batch_stride = static_cast<int64_t>(P) * H * hd;
    CHECK_RC(engine_gemm(e, g));
  }
  {
    GemmArgs g = mk_gemm(M1, D, H * hd, p1.O, H * hd, w1.o_w.data, H * hd, p1.x_mid, D, EPI_RES);
    g.res = p1.x_in;
    g.ldres = D;
    CHECK_RC(engine_gemm(e, g));
  }
  rmsnorm_fwd(p1.x_mid, w1.post_w.d<float>(), nullptr, 0, p1.n2, p1.rstd2, nullptr, M1, D, 1e-6f, st);
  {
    GemmArgs g = mk_gemm(M1, c.paligemma.mlp_dim, D, p1.n2, D, w1.gate_w.data, D, p1.GU, 2 * c.paligemma.mlp_dim,
                         EPI_GEGLU);
    g.D2 = p1.Hh;
    g.ldd2 = c.paligemma.mlp_dim;
    CHECK_RC(engine_gemm(e, g));
  }
  {
    GemmArgs g = mk_gemm(M1, D, c.paligemma.mlp_dim, p1.Hh, c.paligemma.mlp_dim, w1.down_w.data, c.paligemma.mlp_dim,
                         p1.x_out, D, EPI_RES);
    g.res = p1.x_mid;
    g.ldres = D;
    CHECK_RC(engine_gemm(e, g));
  }
  return 0;
}

// Expert-only layer reading the cache: K,V = cat(cache, new) (modeling_gemma.py:308-310).
// The adaptive norms are fused behind the split-K finish of the GEMM that produces their input (gemm.h norm_*):
// this layer's post-attention norm rides on o_proj, and the NEXT consumer's norm (layer l+1's input norm, or the final
// norm after the last layer) rides on down_proj.  Only layer 0 runs its input norm as a separate kernel.
static int suffix_layer(Engine& e, int l, int B, const float* mod_in, const float* mod_post, int rpb,
                        bool input_norm_done, const float* next_mod, bf16* next_out, bf16* next_gate) {
  cudaStream_t st = e.stream;
  const pi05_config& c = e.cfg;
  const int P = e.P, A = e.A, S = e.S, E = e.E, H = e.H, hd = e.hd;
  const int M2 = B * A, QW = (H + 2) * hd;
  GemmaLayerA& p2 = e.a2[l];
  const GemmaLayerP& w2 = e.ex[l];
  bf16 *Kc = e.Kl[l], *Vc = e.Vl[l];
  if (!input_norm_done) rmsnorm_fwd(p2.x_in, nullptr, mod_in, rpb, p2.n1, p2.rstd1, p2.gate1, M2, E, 1e-6f, st);
  {  // fused qkv projection; RoPE + the q / K-cache / V-cache split ride on its split-K finish (gemm.h GemmRope)
    GemmArgs g = mk_gemm(M2, QW, E, p2.n1, E, w2.q_w.data, E, p2.qkv, QW, EPI_STORE);
    const GemmRope rope{A, H, hd, e.pos, e.nvalid, 1, e.rope_cos, e.rope_sin, p2.Q, Kc, Vc, P, S, B};
    g.rope = &rope;
    CHECK_RC(engine_gemm(e, g));
  }
  {
    GemmArgs g = mk_gemm(A * H, S, hd, p2.Q, hd, Kc, hd, p2.P, e.Spad, EPI_SCALE);
    g.batch = B;
    g.a_batch_stride = static_cast<int64_t>(A) * H * hd;
    g.b_batch_stride = static_cast<int64_t>(S) * hd;
    g.d_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    g.scale = 1.0f / sqrtf(static_cast<float>(hd));
    CHECK_RC(engine_gemm(e, g));
  }
  softmax_fwd(p2.P, e.Spad, A * H, B, S, P, e.pad, nullptr, H, st);
  {
    GemmArgs g = mk_gemm(A * H, hd, S, p2.P, e.Spad, Vc, hd, p2.O, hd, EPI_STORE);
    g.b_major = 1;
    g.batch = B;
    g.a_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    g.b_batch_stride = static_cast<int64_t>(S) * hd;
    g.d_batch_stride = static_cast<int64_t>(A) * H * hd;
    CHECK_RC(engine_gemm(e, g));
  }
  {
    GemmArgs g = mk_gemm(M2, E, H * hd, p2.O, H * hd, w2.o_w.data, H * hd, p2.x_mid, E, EPI_RES);
    g.res = p2.x_in;
    g.ldres = E;
    g.gate = p2.gate1;
    g.gate_rows = rpb;
    g.ldgate = E;
    g.norm_mod = mod_post;  // post-attention adaRMS norm of x_mid -> n2, gate2
    g.norm_rows_per_batch = rpb;
    g.norm_out = p2.n2;
    g.norm_gate_out = p2.gate2;
    CHECK_RC(engine_gemm(e, g));
  }
  {
    GemmArgs g = mk_gemm(M2, c.expert.mlp_dim, E, p2.n2, E, w2.gate_w.data, E, p2.GU, 2 * c.expert.mlp_dim, EPI_GEGLU);
    g.D2 = p2.Hh;
    g.ldd2 = c.expert.mlp_dim;
    CHECK_RC(engine_gemm(e, g));
  }
  {
    GemmArgs g =
        mk_gemm(M2, E, c.expert.mlp_dim, p2.Hh, c.expert.mlp_dim, w2.down_w.data, c.expert.mlp_dim, p2.x_out, E, EPI_RES);
    g.res = p2.x_mid;
    g.ldres = E;
    g.gate = p2.gate2;
    g.gate_rows = rpb;
    g.ldgate = E;
    g.norm_mod = next_mod;  // the next consumer's adaRMS norm of x_out
    g.norm_rows_per_batch = rpb;
    g.norm_out = next_out;
    g.norm_gate_out = next_gate;
    CHECK_RC(engine_gemm(e, g));
  }
  return 0;
}

int engine_prefill(Engine& e, const pi05_batch* b, cudaStream_t st) {
  if (!e.bound || b->batch <= 0 || b->batch > e.Bmax) {
    snprintf(e.err, sizeof(e.err), "pi05_prefill: parameters not bound or batch %d outside [1, %d]", b->batch, e.Bmax);
    set_error(e.err);
    return 8;
  }
  e.stream = st;
  e.taps.clear();
  CHECK_RC(engine_set_token_len(e, b->token_len, "pi05_prefill"));
  e.B = b->batch;
  CHECK_RC(prefix_forward(e, b));
  const int depth = e.cfg.paligemma.depth;
  for (int l = 0; l < depth; ++l) CHECK_RC(prefix_layer(e, l, e.B, /*kv_only=*/l == depth - 1));
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(e.err, sizeof(e.err), "pi05_prefill: %s", cudaGetErrorString(ce));
    set_error(e.err);
    return 9;
  }
  return 0;
}

int engine_denoise(Engine& e, const float* noise, int num_steps, float* actions_out, cudaStream_t st) {
  if (!e.bound || e.B <= 0) {
    snprintf(e.err, sizeof(e.err), "pi05_denoise: call pi05_prefill first");
    set_error(e.err);
    return 8;
  }
  e.stream = st;
  const int B = e.B, A = e.A, E = e.E, ad = e.cfg.action_dim, depth = e.cfg.paligemma.depth;
  const int M2 = B * A;
  const int64_t n = static_cast<int64_t>(M2) * ad;
  cudaMemcpyAsync(actions_out, noise, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
  // pi0_pytorch.py:401-418: dt and time are fp32 tensors; time is a running fp32 sum; loop while time >= -dt/2
  const float dt = static_cast<float>(-1.0 / static_cast<double>(num_steps));
  int nsteps = 0;
  for (float time = 1.0f; time >= -dt / 2; time = time + dt) ++nsteps;
  if (nsteps > Engine::kMaxDecodeSteps) {
    snprintf(e.err, sizeof(e.err), "pi05_denoise: num_steps %d exceeds the engine's table (%d)", num_steps,
             Engine::kMaxDecodeSteps);
    set_error(e.err);
    return 8;
  }
  // The time conditioning depends only on the step index: sincos embedding -> time MLP -> all 2*depth+1 adaRMS
  // modulation layers are evaluated ONCE per call for every step (M = nsteps), instead of once per step with M = B.
  // (time.expand(bsize) in the reference: every batch row shares the step's modulation, so rows_per_batch = B*A.)
  const int nm = 2 * depth + 1;
  const int64_t srow = static_cast<int64_t>(3) * E;
  decode_times(e.dec_times, nsteps, dt, st);
  time_embedding(e.dec_times, e.time_scaling, e.dec_temb, nsteps, E / 2, st);
  linear_f32(e.dec_temb, e.tin_w.d<float>(), e.tin_b.d<float>(), e.dec_t1, nsteps, E, E, st);
  silu_fwd(e.dec_t1, e.dec_t1s, static_cast<int64_t>(nsteps) * E, st);
  linear_f32(e.dec_t1s, e.tout_w.d<float>(), e.tout_b.d<float>(), e.dec_t2, nsteps, E, E, st);
  silu_fwd(e.dec_t2, e.dec_cond, static_cast<int64_t>(nsteps) * E, st);
  if (e.ada_uniform && depth > 0) {
    linear_f32_batched(e.dec_cond, e.ex[0].in_dw.d<float>(), e.ex[0].in_db.d<float>(), e.dec_mods, nsteps, 3 * E, E, nm,
                       e.ada_wstride, e.ada_bstride, nsteps * srow, st);
  } else {
    for (int j = 0; j < nm; ++j) {
      const PRef& dw = (j == 2 * depth) ? e.ex_norm_dw : ((j & 1) ? e.ex[j / 2].post_dw : e.ex[j / 2].in_dw);
      const PRef& db = (j == 2 * depth) ? e.ex_norm_db : ((j & 1) ? e.ex[j / 2].post_db : e.ex[j / 2].in_db);
      linear_f32(e.dec_cond, dw.d<float>(), db.d<float>(), e.dec_mods + j * nsteps * srow, nsteps, 3 * E, E, st);
    }
  }
  auto mod_at = [&](int j, int s) { return e.dec_mods + (static_cast<int64_t>(j) * nsteps + s) * srow; };
  for (int step = 0; step < nsteps; ++step) {
    // embed_suffix (pi05): action_in_proj of the current x_t, cast to bf16 (pi0_pytorch.py:270-273,332-337)
    linear_f32(actions_out, e.ain_w.d<float>(), e.ain_b.d<float>(), e.aemb32, M2, E, ad, st);
    cast_f32_to_bf16(e.aemb32, e.a2[0].x_in, static_cast<int64_t>(M2) * E, st);
    for (int l = 0; l < depth; ++l) {
      const bool last = l == depth - 1;
      CHECK_RC(suffix_layer(e, l, B, mod_at(2 * l, step), mod_at(2 * l + 1, step), M2, /*input_norm_done=*/l > 0,
                            mod_at(last ? 2 * depth : 2 * (l + 1), step), last ? e.suffix_out : e.a2[l + 1].n1,
                            last ? nullptr : e.a2[l + 1].gate1));
    }
    if (depth == 0)
      rmsnorm_fwd(e.a2[0].x_in, nullptr, mod_at(0, step), M2, e.suffix_out, e.rstd_f2, nullptr, M2, E, 1e-6f, st);
    cast_bf16_to_f32(e.suffix_out, e.so32, static_cast<int64_t>(M2) * E, st);
    linear_f32(e.so32, e.aout_w.d<float>(), e.aout_b.d<float>(), e.v_t, M2, ad, E, st);
    if (e.taps_enabled && step == 0) {  // keep a copy: v_t is overwritten by the following steps (u_t is free in decode)
      cudaMemcpyAsync(e.u_t, e.v_t, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
      add_tap(e, "v_t_step0", e.u_t, n, PI05_F32);
    }
    euler_step(actions_out, e.v_t, dt, n, st);
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(e.err, sizeof(e.err), "pi05_denoise: %s", cudaGetErrorString(ce));
    set_error(e.err);
    return 9;
  }
  return 0;
}

}  // namespace pi05
