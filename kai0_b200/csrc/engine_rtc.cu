// Real-time-chunking (RTC) guided decoding on the engine.
//
// The reference implements RTC only in its JAX model: src/openpi/models/pi0_rtc.py:234-360 ("R:" below).  Per Euler step it
// needs, besides the velocity v(x_t, t) of the ordinary decode step (pi0_pytorch.py:421-461), the vector-Jacobian product of
// the denoiser x_1(x) = x - t * v(x) with the prefix-weighted error to the previous action chunk (R:331-339,
// `jax.vjp(denoiser, x)`): J^T e = e - t * (dv/dx)^T e.  (dv/dx)^T e is the INPUT gradient of the suffix-only network:
// this file runs the expert stack with a per-layer activation stash (the suffix half of engine.cu::joint_layer_forward
// reading the prefix K/V cache written by pi05_prefill) and then the suffix half of engine_bwd.cu::joint_layer_backward
// without any weight gradient.  Nothing is recomputed.  Needs an engine created with cfg.rtc = 1 (per-layer suffix
// buffers + the suffix backward scratch: ~80 MB at B = 1).
#include <cmath>
#include <cstdio>

#include "common.cuh"
#include "engine.h"
#include "errors.h"
#include "gemm.h"
#include "launch.h"

namespace pi05 {

#define CHECK_RC(x)           \
  do {                        \
    int _rc = (x);            \
    if (_rc != 0) return _rc; \
  } while (0)

int engine_gemm(Engine& e, const GemmArgs& a);
GemmArgs mk_gemm(int M, int N, int K, const void* A, int64_t lda, const void* Bm, int64_t ldb, void* D, int64_t ldd,
                 int epi);

namespace {

// x_in[b, a, j] = prev[b, a, j] where (a < mask_rows && j < provided), else x[b, a, j]      (R:328-333)
__global__ void __launch_bounds__(256) rtc_input_k(const float* __restrict__ x, const float* __restrict__ prev,
                                                   float* __restrict__ x_in, int64_t n, int A, int ad, int mask_rows,
                                                   int provided) {
  pdl_enter();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % ad), a = static_cast<int>((i / ad) % A);
    x_in[i] = (a < mask_rows && j < provided) ? prev[i] : x[i];
  }
}
// err = (prev - (x_in - t * v)) * w[a] * dm[j]                                               (R:336-338)
__global__ void __launch_bounds__(256) rtc_error_k(const float* __restrict__ x_in, const float* __restrict__ v,
                                                   const float* __restrict__ prev, const float* __restrict__ w,
                                                   const float* __restrict__ dm, float t, float* __restrict__ err,
                                                   int64_t n, int A, int ad) {
  pdl_enter();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % ad), a = static_cast<int>((i / ad) % A);
    const float x1 = x_in[i] - t * v[i];
    err[i] = (prev[i] - x1) * w[a] * dm[j];
  }
}
__device__ __forceinline__ float nan_to_zero(float v) { return (isnan(v) || isinf(v)) ? 0.f : v; }
// corr = err - t * g  (J^T err);  v_t = nan_to_num(v - gw * corr);  x += dt * v_t           (R:339,348-350)
__global__ void __launch_bounds__(256) rtc_update_k(float* __restrict__ x, const float* __restrict__ v,
                                                    const float* __restrict__ err, const float* __restrict__ g, float t,
                                                    float gw, float dt, int64_t n) {
  pdl_enter();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float corr = err[i] - t * g[i];
    const float vt = nan_to_zero(v[i] - gw * corr);
    x[i] = x[i] + dt * vt;
  }
}
__global__ void __launch_bounds__(256) nan_to_num_k(float* __restrict__ x, int64_t n) {
  pdl_enter();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    x[i] = nan_to_zero(x[i]);
}
inline int grid_for(int64_t n) { return static_cast<int>((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184); }

// The expert stream of one joint layer with every intermediate the backward needs kept in a2[l]
// (gemma_pytorch.py:158-238 restricted to the suffix; K/V rows [0, P) of the cache come from pi05_prefill).
int suffix_layer_fwd_stash(Engine& e, int l, int B, const float* mod_in, const float* mod_post, int rpb) {
  cudaStream_t st = e.stream;
  const pi05_config& c = e.cfg;
  const int P = e.P, A = e.A, S = e.S, E = e.E, H = e.H, hd = e.hd;
  const int M2 = B * A, QW = (H + 2) * hd;
  GemmaLayerA& p2 = e.a2[l];
  const GemmaLayerP& w2 = e.ex[l];
  bf16 *Kc = e.Kl[l], *Vc = e.Vl[l];
  rmsnorm_fwd(p2.x_in, nullptr, mod_in, rpb, p2.n1, p2.rstd1, p2.gate1, M2, E, 1e-6f, st);
  CHECK_RC(engine_gemm(e, mk_gemm(M2, QW, E, p2.n1, E, w2.q_w.data, E, p2.qkv, QW, EPI_STORE)));
  rope_pack_fwd(p2.qkv, A, H, hd, e.pos, e.nvalid, 1, e.rope_cos, e.rope_sin, p2.Q, Kc, Vc, P, S, B, st);
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
    g.D2 = p2.o_lin;
    g.ldd2 = E;
    CHECK_RC(engine_gemm(e, g));
  }
  rmsnorm_fwd(p2.x_mid, nullptr, mod_post, rpb, p2.n2, p2.rstd2, p2.gate2, M2, E, 1e-6f, st);
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
    g.D2 = p2.d_lin;
    g.ldd2 = E;
    CHECK_RC(engine_gemm(e, g));
  }
  return 0;
}

int dgrad(Engine& e, const bf16* dY, int M, int N, const void* W, int K, bf16* dX) {
  GemmArgs g = mk_gemm(M, K, N, dY, N, W, K, dX, K, EPI_STORE);
  g.b_major = 1;
  return engine_gemm(e, g);
}

// Input gradient of that layer: e.g_x2 = d(x_out) on entry, d(x_in) on exit.  The suffix half of
// engine_bwd.cu::joint_layer_backward, same kernels and rounding points, no weight gradients; the keys / values of the
// prefix are constants here (their rows of dK / dV are computed by the shared GEMM and ignored).
int suffix_layer_bwd_input(Engine& e, int l, int B, const float* mod_in, const float* mod_post, int rpb) {
  cudaStream_t st = e.stream;
  const pi05_config& c = e.cfg;
  const int P = e.P, A = e.A, S = e.S, E = e.E, H = e.H, hd = e.hd;
  const int M2 = B * A, QW = (H + 2) * hd, HD = H * hd, mlp2 = c.expert.mlp_dim;
  GemmaLayerA& p2 = e.a2[l];
  const GemmaLayerP& w2 = e.ex[l];
  bf16 *Kc = e.Kl[l], *Vc = e.Vl[l];
  const float scaling = 1.0f / sqrtf(static_cast<float>(hd));
  bf16 *g2 = e.g_x2, *g2m = e.g_x2b;
  float* dmod = e.g_dmods;  // gradients of the modulation are not needed: one scratch block, never read
  bf16* do2 = e.g2_do;
  gated_residual_bwd(g2, p2.d_lin, p2.gate2, rpb, do2, dmod, M2, E, st);
  {
    GemmArgs g = mk_gemm(M2, mlp2, E, do2, E, w2.down_w.data, mlp2, e.g2_big, 2 * mlp2, EPI_GEGLU_BWD);
    g.b_major = 1;
    g.res = p2.GU;
    g.ldres = 2 * mlp2;
    CHECK_RC(engine_gemm(e, g));
  }
  CHECK_RC(dgrad(e, e.g2_big, M2, 2 * mlp2, w2.gate_w.data, E, e.g2_t1));
  rmsnorm_bwd(e.g2_t1, p2.x_mid, nullptr, mod_post, rpb, p2.rstd2, g2, g2m, nullptr, dmod, M2, E, st);
  bf16* dol2 = e.g2_do;
  gated_residual_bwd(g2m, p2.o_lin, p2.gate1, rpb, dol2, dmod, M2, E, st);
  CHECK_RC(dgrad(e, dol2, M2, E, w2.o_w.data, HD, e.g2_t2));  // dO2 [B, A*H, hd]
  {
    GemmArgs g = mk_gemm(A * H, S, hd, e.g2_t2, hd, Vc, hd, e.g_P, e.Spad, EPI_STORE);  // dP = dO V^T
    g.batch = B;
    g.a_batch_stride = static_cast<int64_t>(A) * H * hd;
    g.b_batch_stride = static_cast<int64_t>(S) * hd;
    g.d_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    CHECK_RC(engine_gemm(e, g));
    softmax_bwd(p2.P, e.g_P, e.Spad, B * A * H, S, scaling, st);  // -> dS
    GemmArgs q = mk_gemm(A * H, hd, S, e.g_P, e.Spad, Kc, hd, e.g2_t3, hd, EPI_STORE);  // dQ = dS K
    q.b_major = 1;
    q.batch = B;
    q.a_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    q.b_batch_stride = static_cast<int64_t>(S) * hd;
    q.d_batch_stride = static_cast<int64_t>(A) * H * hd;
    CHECK_RC(engine_gemm(e, q));
    GemmArgs k = mk_gemm(S, hd, A * H, e.g_P, e.Spad, p2.Q, hd, e.g_dK, hd, EPI_F32);  // dK = dS^T Q
    k.a_major = 1;
    k.b_major = 1;
    k.batch = B;
    k.a_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    k.b_batch_stride = static_cast<int64_t>(A) * H * hd;
    k.d_batch_stride = static_cast<int64_t>(S) * hd;
    CHECK_RC(engine_gemm(e, k));
    GemmArgs v = mk_gemm(S, hd, A * H, p2.P, e.Spad, e.g2_t2, hd, e.g_dV, hd, EPI_F32);  // dV = P^T dO
    v.a_major = 1;
    v.b_major = 1;
    v.batch = B;
    v.a_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    v.b_batch_stride = static_cast<int64_t>(A) * H * hd;
    v.d_batch_stride = static_cast<int64_t>(S) * hd;
    CHECK_RC(engine_gemm(e, v));
  }
  bf16* dqkv2 = e.g2_t1;
  rope_pack_bwd(e.g2_t3, e.g_dK, e.g_dV, A, H, hd, e.pos, e.nvalid, 1, e.rope_cos, e.rope_sin, dqkv2, P, S, B, st);
  CHECK_RC(dgrad(e, dqkv2, M2, QW, w2.q_w.data, E, e.g2_t2));  // dn1
  rmsnorm_bwd(e.g2_t2, p2.x_in, nullptr, mod_in, rpb, p2.rstd1, g2m, g2, nullptr, dmod, M2, E, st);
  return 0;
}

}  // namespace

// guidance: HOST array [num_steps] of min(c * inv_r2, max_guidance_weight) per step (R:341-347), computed by the caller from
// the same fp32 running-sum times; time_weights: device [A] (R:47-61); dim_mask: device [action_dim] (R:326-327).
int engine_denoise_rtc(Engine& e, const float* noise, int num_steps, const float* prev, const float* time_weights,
                       const float* dim_mask, const float* guidance, int mask_rows, int provided, float* actions_out,
                       cudaStream_t st) {
  if (!e.bound || e.B <= 0) {
    snprintf(e.err, sizeof(e.err), "pi05_denoise_rtc: call pi05_prefill first");
    set_error(e.err);
    return 8;
  }
  if (!e.cfg.rtc || e.g_x2 == nullptr) {
    snprintf(e.err, sizeof(e.err), "pi05_denoise_rtc: engine was not created with cfg.rtc = 1");
    set_error(e.err);
    return 8;
  }
  if (num_steps < 1 || num_steps > Engine::kMaxDecodeSteps || !prev || !time_weights || !dim_mask || !guidance) {
    snprintf(e.err, sizeof(e.err), "pi05_denoise_rtc: bad argument (1 <= num_steps <= %d, non-null tables)",
             Engine::kMaxDecodeSteps);
    set_error(e.err);
    return 8;
  }
  e.stream = st;
  const int B = e.B, A = e.A, E = e.E, ad = e.cfg.action_dim, depth = e.cfg.paligemma.depth;
  const int M2 = B * A;
  const int64_t n = static_cast<int64_t>(M2) * ad;
  cudaMemcpyAsync(actions_out, noise, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
  const float dt = static_cast<float>(-1.0 / static_cast<double>(num_steps));  // R:256
  const int nsteps = num_steps;  // R:354-358: exactly num_steps steps (lax.scan), time an fp32 running sum from 1.0
  const int nm = 2 * depth + 1;
  const int64_t srow = static_cast<int64_t>(3) * E;
  decode_times(e.dec_times, nsteps, dt, st);
  time_embedding(e.dec_times, e.time_scaling, e.dec_temb, nsteps, E / 2, st);
  linear_f32(e.dec_temb, e.tin_w.d<float>(), e.tin_b.d<float>(), e.dec_t1, nsteps, E, E, st);
  silu_fwd(e.dec_t1, e.dec_t1s, static_cast<int64_t>(nsteps) * E, st);
  linear_f32(e.dec_t1s, e.tout_w.d<float>(), e.tout_b.d<float>(), e.dec_t2, nsteps, E, E, st);
  silu_fwd(e.dec_t2, e.dec_cond, static_cast<int64_t>(nsteps) * E, st);
  for (int j = 0; j < nm; ++j) {
    const PRef& dw = (j == 2 * depth) ? e.ex_norm_dw : ((j & 1) ? e.ex[j / 2].post_dw : e.ex[j / 2].in_dw);
    const PRef& db = (j == 2 * depth) ? e.ex_norm_db : ((j & 1) ? e.ex[j / 2].post_db : e.ex[j / 2].in_db);
    linear_f32(e.dec_cond, dw.d<float>(), db.d<float>(), e.dec_mods + j * nsteps * srow, nsteps, 3 * E, E, st);
  }
  auto mod_at = [&](int j, int s) { return e.dec_mods + (static_cast<int64_t>(j) * nsteps + s) * srow; };
  fill_zero(e.g_dmods, static_cast<size_t>(B) * 3 * E * sizeof(float), st);
  float* x_in = e.x_t;     // [M2, ad] fp32: the denoiser's input (x_t with the delayed prefix overwritten)
  float* err = e.u_t;      // [M2, ad]
  float* gin = e.g_f32c;   // [M2, ad]: (dv/dx)^T err
  const bf16* x2f = depth > 0 ? e.a2[depth - 1].x_out : e.a2[0].x_in;
  float time = 1.0f;
  for (int step = 0; step < nsteps; ++step, time = time + dt) {
    launch_pdl(rtc_input_k, dim3(grid_for(n)), dim3(256), 0, st, actions_out, prev, x_in, n, A, ad, mask_rows, provided);
    count_launch();
    // ---- forward with stash: v = denoise_step(x_in, time)
    linear_f32(x_in, e.ain_w.d<float>(), e.ain_b.d<float>(), e.aemb32, M2, E, ad, st);
    cast_f32_to_bf16(e.aemb32, e.a2[0].x_in, static_cast<int64_t>(M2) * E, st);
    for (int l = 0; l < depth; ++l)
      CHECK_RC(suffix_layer_fwd_stash(e, l, B, mod_at(2 * l, step), mod_at(2 * l + 1, step), M2));
    rmsnorm_fwd(x2f, nullptr, mod_at(2 * depth, step), M2, e.suffix_out, e.rstd_f2, nullptr, M2, E, 1e-6f, st);
    cast_bf16_to_f32(e.suffix_out, e.so32, static_cast<int64_t>(M2) * E, st);
    linear_f32(e.so32, e.aout_w.d<float>(), e.aout_b.d<float>(), e.v_t, M2, ad, E, st);
    // ---- error to the previous chunk at the denoised endpoint, then its pull-back through the network
    launch_pdl(rtc_error_k, dim3(grid_for(n)), dim3(256), 0, st, x_in, e.v_t, prev, time_weights, dim_mask, time, err, n, A,
               ad);
    count_launch();
    linear_f32_dgrad(err, e.aout_w.d<float>(), e.g_f32b, M2, ad, E, 0, st);          // d so32
    cast_f32_to_bf16(e.g_f32b, e.g_x2b, static_cast<int64_t>(M2) * E, st);            // grad of the .float() cast
    rmsnorm_bwd(e.g_x2b, x2f, nullptr, mod_at(2 * depth, step), M2, e.rstd_f2, nullptr, e.g_x2, nullptr, e.g_dmods, M2, E,
                st);
    for (int l = depth - 1; l >= 0; --l)
      CHECK_RC(suffix_layer_bwd_input(e, l, B, mod_at(2 * l, step), mod_at(2 * l + 1, step), M2));
    cast_bf16_to_f32(e.g_x2, e.g_f32a, static_cast<int64_t>(M2) * E, st);             // grad through the bf16 cast
    linear_f32_dgrad(e.g_f32a, e.ain_w.d<float>(), gin, M2, E, ad, 0, st);            // (dv/dx)^T err
    if (e.taps_enabled && step == 0) {
      cudaMemcpyAsync(e.rtc_tap_v, e.v_t, n * sizeof(float), cudaMemcpyDeviceToDevice, st);  // later steps overwrite both
      cudaMemcpyAsync(e.rtc_tap_j, gin, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
      add_tap(e, "rtc_v_step0", e.rtc_tap_v, n, PI05_F32);
      add_tap(e, "rtc_vjp_step0", e.rtc_tap_j, n, PI05_F32);
    }
    launch_pdl(rtc_update_k, dim3(grid_for(n)), dim3(256), 0, st, actions_out, e.v_t, err, gin, time, guidance[step], dt, n);
    count_launch();
  }
  launch_pdl(nan_to_num_k, dim3(grid_for(n)), dim3(256), 0, st, actions_out, n);  // R:359
  count_launch();
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(e.err, sizeof(e.err), "pi05_denoise_rtc: %s", cudaGetErrorString(ce));
    set_error(e.err);
    return 9;
  }
  return 0;
}

}  // namespace pi05
