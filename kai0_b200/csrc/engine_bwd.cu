// Backward pass of the pi0.5 training step: the autograd counterpart of engine.cu, written out by hand.
// It is what torch.autograd derives for pi0_pytorch.py:316-373 / gemma_pytorch.py:158-275 /
// modeling_siglip.py:435-481, with every dgrad / wgrad contraction on the tcgen05 GEMM (MN-major operands) and
// no recomputation (the reference re-runs every layer under torch.utils.checkpoint, gemma_pytorch.py:241-252).
//
// Gradient buffers are overwritten, never accumulated (the caller's zero_grad(set_to_none=True) semantics,
// train_pytorch.py:561).
#include <algorithm>
#include <cmath>
#include <cstdio>

#include <cuda_profiler_api.h>

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

// dX[M, K] = dY[M, N] @ W[N, K]      (W row-major as stored by nn.Linear: N-major B operand)
static int dgrad(Engine& e, const bf16* dY, int M, int N, const void* W, int K, bf16* dX) {
  GemmArgs g = mk_gemm(M, K, N, dY, N, W, K, dX, K, EPI_STORE);
  g.b_major = 1;
  return engine_gemm(e, g);
}
// dW[N, K] = dY[M, N]^T @ X[M, K]    (both operands MN-major; contraction over the M rows)
static int wgrad(Engine& e, const bf16* dY, int M, int N, const bf16* X, int K, void* dW) {
  GemmArgs g = mk_gemm(N, K, M, dY, N, X, K, dW, K, EPI_STORE);
  g.a_major = 1;
  g.b_major = 1;
  return engine_gemm(e, g);
}

// Down-projection dgrad with the GeGLU backward fused into the epilogue (modeling_gemma.py:125):
// dH[M, mlp] = dY[M, D] @ Wd[D, mlp] is consumed in registers; dGU[M, 2*mlp] = [dg | du] is what reaches HBM.
static int dgrad_geglu(Engine& e, const bf16* dY, int M, int N, const void* Wd, int mlp, const bf16* GU, bf16* dGU) {
  GemmArgs g = mk_gemm(M, mlp, N, dY, N, Wd, mlp, dGU, 2 * mlp, EPI_GEGLU_BWD);
  g.b_major = 1;
  g.res = GU;
  g.ldres = 2 * mlp;
  return engine_gemm(e, g);
}

// bias-style gradient: bf16 grad[cols] = colsum(dY) accumulated in fp32
static void bias_grad(Engine& e, const bf16* dY, int64_t ld, int64_t rows, int cols, bf16* grad) {
  fill_zero(e.g_acc, static_cast<size_t>(cols) * sizeof(float), e.stream);
  colsum_bf16(dY, ld, rows, cols, e.g_acc, e.stream);
  cast_f32_to_bf16(e.g_acc, grad, cols, e.stream);
}

// ------------------------------------------------------------------------------------------------------------
// one joint layer
// ------------------------------------------------------------------------------------------------------------
static int joint_layer_backward(Engine& e, int l, int B, bool g1_zero) {
  cudaStream_t st = e.stream;
  const pi05_config& c = e.cfg;
  const int P = e.P, A = e.A, S = e.S, D = e.D, E = e.E, H = e.H, hd = e.hd;
  const int M1 = B * P, M2 = B * A, QW = (H + 2) * hd, HD = H * hd;
  const int mlp1 = c.paligemma.mlp_dim, mlp2 = c.expert.mlp_dim;
  GemmaLayerA &p1 = e.a1[l], &p2 = e.a2[l];
  const GemmaLayerP &w1 = e.pg[l], &w2 = e.ex[l];
  const int64_t ms = static_cast<int64_t>(B) * 3 * E;
  const float* mod_in = e.mods + (2 * l) * ms;
  const float* mod_post = e.mods + (2 * l + 1) * ms;
  float* dmod_in = e.g_dmods + (2 * l) * ms;
  float* dmod_post = e.g_dmods + (2 * l + 1) * ms;
  bf16 *Kc = e.Kl[l], *Vc = e.Vl[l];
  const float scaling = 1.0f / sqrtf(static_cast<float>(hd));
  // gradient holders: g_x1/g_x2 = d(x_out) on entry, d(x_in) on exit; g_x1b/g_x2b = d(x_mid)
  bf16 *g1 = e.g_x1, *g1m = e.g_x1b, *g2 = e.g_x2, *g2m = e.g_x2b;

  // ================= expert stream: MLP half =================
  bf16* do2 = e.g2_do;
  gated_residual_bwd(g2, p2.d_lin, p2.gate2, A, do2, dmod_post, M2, E, st);                 // x_out = x_mid + d*gate
  CHECK_RC(dgrad_geglu(e, do2, M2, E, w2.down_w.data, mlp2, p2.GU, e.g2_big));  // small M: dH -> dGU fused in the epilogue
  CHECK_RC(wgrad(e, do2, M2, E, p2.Hh, mlp2, w2.down_w.grad));
  CHECK_RC(dgrad(e, e.g2_big, M2, 2 * mlp2, w2.gate_w.data, E, e.g2_t1));                    // dn2
  CHECK_RC(wgrad(e, e.g2_big, M2, 2 * mlp2, p2.n2, E, w2.gate_w.grad));
  rmsnorm_bwd(e.g2_t1, p2.x_mid, nullptr, mod_post, A, p2.rstd2, g2, g2m, nullptr, dmod_post, M2, E, st);
  // ================= prefix stream: MLP half =================
  if (!g1_zero) {
    // Measured (profiles/r01): fusing the GeGLU backward into this dgrad's epilogue makes the K = 2048 GEMM
    // epilogue-bound (1.88 ms vs 0.84 + 0.67 ms unfused), so the big stream keeps the separate streaming kernel.
    CHECK_RC(dgrad(e, g1, M1, D, w1.down_w.data, mlp1, e.g_big2));
    CHECK_RC(wgrad(e, g1, M1, D, p1.Hh, mlp1, w1.down_w.grad));
    geglu_bwd(e.g_big2, p1.GU, e.g_big, M1, mlp1, st);
    CHECK_RC(dgrad(e, e.g_big, M1, 2 * mlp1, w1.gate_w.data, D, e.g_t1));
    CHECK_RC(wgrad(e, e.g_big, M1, 2 * mlp1, p1.n2, D, w1.gate_w.grad));
    fill_zero(w1.post_w.grad, static_cast<size_t>(D) * sizeof(float), st);
    rmsnorm_bwd(e.g_t1, p1.x_mid, w1.post_w.d<float>(), nullptr, 0, p1.rstd2, g1, g1m, w1.post_w.g<float>(), nullptr,
                M1, D, st);
  } else {
    // x1_out of the last layer feeds nothing that reaches the loss (prefix_out is unused, pi0_pytorch.py:350-358)
    fill_zero(w1.down_w.grad, static_cast<size_t>(D) * mlp1 * 2, st);
    fill_zero(w1.gate_w.grad, static_cast<size_t>(2) * mlp1 * D * 2, st);
    fill_zero(w1.post_w.grad, static_cast<size_t>(D) * sizeof(float), st);
    fill_zero(w1.o_w.grad, static_cast<size_t>(D) * HD * 2, st);
  }

  // ================= attention half: o_proj =================
  bf16* dol2 = e.g2_do;  // reuse: d(o_lin) of the expert stream
  gated_residual_bwd(g2m, p2.o_lin, p2.gate1, A, dol2, dmod_in, M2, E, st);                  // x_mid = x_in + o*gate
  CHECK_RC(dgrad(e, dol2, M2, E, w2.o_w.data, HD, e.g2_t2));                                 // dO2 [B, A*H, hd]
  CHECK_RC(wgrad(e, dol2, M2, E, p2.O, HD, w2.o_w.grad));
  if (!g1_zero) {
    CHECK_RC(dgrad(e, g1m, M1, D, w1.o_w.data, HD, e.g_t2));                                 // dO1 [B, P*H, hd]
    CHECK_RC(wgrad(e, g1m, M1, D, p1.O, HD, w1.o_w.grad));
  }

  // ================= attention core (suffix queries first; g_P is shared) =================
  {
    // dP2 = dO2 V^T
    GemmArgs g = mk_gemm(A * H, S, hd, e.g2_t2, hd, Vc, hd, e.g_P, e.Spad, EPI_STORE);
    g.batch = B;
    g.a_batch_stride = static_cast<int64_t>(A) * H * hd;
    g.b_batch_stride = static_cast<int64_t>(S) * hd;
    g.d_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    CHECK_RC(engine_gemm(e, g));
    softmax_bwd(p2.P, e.g_P, e.Spad, B * A * H, S, scaling, st);  // -> dS2
    // dQ2 = dS2 K
    GemmArgs q = mk_gemm(A * H, hd, S, e.g_P, e.Spad, Kc, hd, e.g2_t3, hd, EPI_STORE);
    q.b_major = 1;
    q.batch = B;
    q.a_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    q.b_batch_stride = static_cast<int64_t>(S) * hd;
    q.d_batch_stride = static_cast<int64_t>(A) * H * hd;
    CHECK_RC(engine_gemm(e, q));
    // dK = dS2^T Q2 (all S key rows, fp32)
    GemmArgs k = mk_gemm(S, hd, A * H, e.g_P, e.Spad, p2.Q, hd, e.g_dK, hd, EPI_F32);
    k.a_major = 1;
    k.b_major = 1;
    k.batch = B;
    k.a_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    k.b_batch_stride = static_cast<int64_t>(A) * H * hd;
    k.d_batch_stride = static_cast<int64_t>(S) * hd;
    CHECK_RC(engine_gemm(e, k));
    // dV = P2^T dO2
    GemmArgs v = mk_gemm(S, hd, A * H, p2.P, e.Spad, e.g2_t2, hd, e.g_dV, hd, EPI_F32);
    v.a_major = 1;
    v.b_major = 1;
    v.batch = B;
    v.a_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    v.b_batch_stride = static_cast<int64_t>(A) * H * hd;
    v.d_batch_stride = static_cast<int64_t>(S) * hd;
    CHECK_RC(engine_gemm(e, v));
  }
  bf16* dQ1 = e.g_t3;
  if (!g1_zero) {
    GemmArgs g = mk_gemm(P * H, P, hd, e.g_t2, hd, Vc, hd, e.g_P, e.Ppad, EPI_STORE);
    g.batch = B;
    g.a_batch_stride = static_cast<int64_t>(P) * H * hd;
    g.b_batch_stride = static_cast<int64_t>(S) * hd;
    g.d_batch_stride = static_cast<int64_t>(P) * H * e.Ppad;
    CHECK_RC(engine_gemm(e, g));
    softmax_bwd(p1.P, e.g_P, e.Ppad, B * P * H, P, scaling, st);  // -> dS1
    GemmArgs q = mk_gemm(P * H, hd, P, e.g_P, e.Ppad, Kc, hd, dQ1, hd, EPI_STORE);
    q.b_major = 1;
    q.batch = B;
    q.a_batch_stride = static_cast<int64_t>(P) * H * e.Ppad;
    q.b_batch_stride = static_cast<int64_t>(S) * hd;
    q.d_batch_stride = static_cast<int64_t>(P) * H * hd;
    CHECK_RC(engine_gemm(e, q));
    GemmArgs k = mk_gemm(P, hd, P * H, e.g_P, e.Ppad, p1.Q, hd, e.g_dK, hd, EPI_F32);  // += rows [0, P)
    k.a_major = 1;
    k.b_major = 1;
    k.batch = B;
    k.accumulate = 1;
    k.a_batch_stride = static_cast<int64_t>(P) * H * e.Ppad;
    k.b_batch_stride = static_cast<int64_t>(P) * H * hd;
    k.d_batch_stride = static_cast<int64_t>(S) * hd;
    CHECK_RC(engine_gemm(e, k));
    GemmArgs v = mk_gemm(P, hd, P * H, p1.P, e.Ppad, e.g_t2, hd, e.g_dV, hd, EPI_F32);
    v.a_major = 1;
    v.b_major = 1;
    v.batch = B;
    v.accumulate = 1;
    v.a_batch_stride = static_cast<int64_t>(P) * H * e.Ppad;
    v.b_batch_stride = static_cast<int64_t>(P) * H * hd;
    v.d_batch_stride = static_cast<int64_t>(S) * hd;
    CHECK_RC(engine_gemm(e, v));
  } else {
    fill_zero(dQ1, static_cast<size_t>(M1) * HD * 2, st);
  }
  // RoPE^T and re-assembly of the fused qkv gradient rows
  bf16* dqkv1 = e.g_t1;
  bf16* dqkv2 = e.g2_t1;
  rope_pack_bwd(dQ1, e.g_dK, e.g_dV, P, H, hd, e.pos, e.nvalid, 0, e.rope_cos, e.rope_sin, dqkv1, 0, S, B, st);
  rope_pack_bwd(e.g2_t3, e.g_dK, e.g_dV, A, H, hd, e.pos, e.nvalid, 1, e.rope_cos, e.rope_sin, dqkv2, P, S, B, st);

  // ================= qkv projections + input norms =================
  CHECK_RC(dgrad(e, dqkv2, M2, QW, w2.q_w.data, E, e.g2_t2));  // dn1 (expert)
  CHECK_RC(wgrad(e, dqkv2, M2, QW, p2.n1, E, w2.q_w.grad));
  rmsnorm_bwd(e.g2_t2, p2.x_in, nullptr, mod_in, A, p2.rstd1, g2m, g2, nullptr, dmod_in, M2, E, st);
  CHECK_RC(dgrad(e, dqkv1, M1, QW, w1.q_w.data, D, e.g_t2));
  CHECK_RC(wgrad(e, dqkv1, M1, QW, p1.n1, D, w1.q_w.grad));
  fill_zero(w1.in_w.grad, static_cast<size_t>(D) * sizeof(float), st);
  rmsnorm_bwd(e.g_t2, p1.x_in, w1.in_w.d<float>(), nullptr, 0, p1.rstd1, g1_zero ? nullptr : g1m, g1,
              w1.in_w.g<float>(), nullptr, M1, D, st);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// vision tower
// ------------------------------------------------------------------------------------------------------------
static int vit_layer_backward(Engine& e, int l, int B) {
  cudaStream_t st = e.stream;
  const pi05_config& c = e.cfg;
  const int nimg = e.NI * B, T = e.T, W = e.W, VH = e.VH, vhd = e.vhd, mlp = c.vit_mlp_dim;
  const int Mv = nimg * T;
  const VitLayerP& p = e.vit[l];
  VitLayerA& a = e.va[l];
  bf16 *g = e.g_x1, *gm = e.g_x1b;
  float* acc = e.g_acc;
  // ---- MLP half: x_out = x_mid + fc2(act) + b
  CHECK_RC(dgrad(e, g, Mv, W, p.fc2_w.data, mlp, e.g_big2));  // d act
  CHECK_RC(wgrad(e, g, Mv, W, a.act, mlp, p.fc2_w.grad));
  bias_grad(e, g, W, Mv, W, p.fc2_b.g<bf16>());
  gelu_bwd(e.g_big2, a.pre, e.g_big, static_cast<int64_t>(Mv) * mlp, st);  // d pre (EPI_GELU_BWD fusion measured slower)
  CHECK_RC(dgrad(e, e.g_big, Mv, mlp, p.fc1_w.data, W, e.g_t1));         // d h2
  CHECK_RC(wgrad(e, e.g_big, Mv, mlp, a.h2, W, p.fc1_w.grad));
  bias_grad(e, e.g_big, mlp, Mv, mlp, p.fc1_b.g<bf16>());
  fill_zero(acc, static_cast<size_t>(2) * W * sizeof(float), st);
  layernorm_bwd(e.g_t1, a.x_mid, p.ln2_w.d<bf16>(), a.mean2, a.rstd2, g, gm, acc, acc + W, Mv, W, st);
  cast_f32_to_bf16(acc, p.ln2_w.g<bf16>(), W, st);
  cast_f32_to_bf16(acc + W, p.ln2_b.g<bf16>(), W, st);
  // ---- attention half: x_mid = x_in + out_proj(attn) + b
  CHECK_RC(dgrad(e, gm, Mv, W, p.out_w.data, W, e.g_t2));  // d attn
  CHECK_RC(wgrad(e, gm, Mv, W, a.attn, W, p.out_w.grad));
  bias_grad(e, gm, W, Mv, W, p.out_b.g<bf16>());
  const int Z = nimg * VH;
  const int64_t s_qkv1 = static_cast<int64_t>(T) * 3 * W, s_x1 = static_cast<int64_t>(T) * W;
  const int64_t s_p0 = static_cast<int64_t>(T) * T, s_p1 = static_cast<int64_t>(VH) * T * T;
  bf16* dqkv = e.g_t3;
  {  // dP = d_attn V^T
    GemmArgs q = mk_gemm(T, T, vhd, e.g_t2, W, a.qkv + 2 * W, 3 * W, e.g_P, T, EPI_STORE);
    q.batch = Z;
    q.batch_inner = VH;
    q.a_batch_stride = vhd;
    q.a_batch_stride1 = s_x1;
    q.b_batch_stride = vhd;
    q.b_batch_stride1 = s_qkv1;
    q.d_batch_stride = s_p0;
    q.d_batch_stride1 = s_p1;
    q.block_n = (T > 128) ? 256 : 128;
    CHECK_RC(engine_gemm(e, q));
  }
  softmax_bwd(a.P, e.g_P, T, Z * T, T, 1.0f / sqrtf(static_cast<float>(vhd)), st);  // -> dS
  {  // dQ = dS K
    GemmArgs q = mk_gemm(T, vhd, T, e.g_P, T, a.qkv + W, 3 * W, dqkv, 3 * W, EPI_STORE);
    q.b_major = 1;
    q.batch = Z;
    q.batch_inner = VH;
    q.a_batch_stride = s_p0;
    q.a_batch_stride1 = s_p1;
    q.b_batch_stride = vhd;
    q.b_batch_stride1 = s_qkv1;
    q.d_batch_stride = vhd;
    q.d_batch_stride1 = s_qkv1;
    q.block_n = 128;
    CHECK_RC(engine_gemm(e, q));
  }
  {  // dK = dS^T Q
    GemmArgs q = mk_gemm(T, vhd, T, e.g_P, T, a.qkv, 3 * W, dqkv + W, 3 * W, EPI_STORE);
    q.a_major = 1;
    q.b_major = 1;
    q.batch = Z;
    q.batch_inner = VH;
    q.a_batch_stride = s_p0;
    q.a_batch_stride1 = s_p1;
    q.b_batch_stride = vhd;
    q.b_batch_stride1 = s_qkv1;
    q.d_batch_stride = vhd;
    q.d_batch_stride1 = s_qkv1;
    q.block_n = 128;
    CHECK_RC(engine_gemm(e, q));
  }
  {  // dV = P^T d_attn
    GemmArgs q = mk_gemm(T, vhd, T, a.P, T, e.g_t2, W, dqkv + 2 * W, 3 * W, EPI_STORE);
    q.a_major = 1;
    q.b_major = 1;
    q.batch = Z;
    q.batch_inner = VH;
    q.a_batch_stride = s_p0;
    q.a_batch_stride1 = s_p1;
    q.b_batch_stride = vhd;
    q.b_batch_stride1 = s_x1;
    q.d_batch_stride = vhd;
    q.d_batch_stride1 = s_qkv1;
    q.block_n = 128;
    CHECK_RC(engine_gemm(e, q));
  }
  CHECK_RC(dgrad(e, dqkv, Mv, 3 * W, p.q_w.data, W, e.g_t1));  // d h1
  CHECK_RC(wgrad(e, dqkv, Mv, 3 * W, a.h1, W, p.q_w.grad));
  bias_grad(e, dqkv, 3 * W, Mv, 3 * W, p.q_b.g<bf16>());
  fill_zero(acc, static_cast<size_t>(2) * W * sizeof(float), st);
  layernorm_bwd(e.g_t1, a.x_in, p.ln1_w.d<bf16>(), a.mean1, a.rstd1, gm, g, acc, acc + W, Mv, W, st);
  cast_f32_to_bf16(acc, p.ln1_w.g<bf16>(), W, st);
  cast_f32_to_bf16(acc + W, p.ln1_b.g<bf16>(), W, st);
  return 0;
}

static int vision_backward(Engine& e, int B) {
  cudaStream_t st = e.stream;
  const pi05_config& c = e.cfg;
  const int nimg = e.NI * B, T = e.T, W = e.W, D = e.D;
  const int Mv = nimg * T;
  // gather d(prefix_embs)[b, n*T + t, :] into (n, b, t) row order = the tower's row order
  bf16* dY = e.g_t1;  // [Mv, D]
  for (int n = 0; n < e.NI; ++n)
    copy_rows_bf16(e.g_x1, static_cast<int64_t>(e.P) * D, n * T, T, D, dY + static_cast<int64_t>(n) * B * T * D, B, st);
  CHECK_RC(wgrad(e, dY, Mv, D, e.vit_post, W, e.proj_w.grad));
  bias_grad(e, dY, D, Mv, D, e.proj_b.g<bf16>());
  CHECK_RC(dgrad(e, dY, Mv, D, e.proj_w.data, W, e.g_t2));  // d vit_post
  float* acc = e.g_acc;
  fill_zero(acc, static_cast<size_t>(2) * W * sizeof(float), st);
  const bf16* xl = c.vit_depth > 0 ? e.va[c.vit_depth - 1].x_out : e.vit_x0;
  layernorm_bwd(e.g_t2, xl, e.post_ln_w.d<bf16>(), e.vit_post_mean, e.vit_post_rstd, nullptr, e.g_x1, acc, acc + W, Mv, W,
                st);
  cast_f32_to_bf16(acc, e.post_ln_w.g<bf16>(), W, st);
  cast_f32_to_bf16(acc + W, e.post_ln_b.g<bf16>(), W, st);
  CHECK_RC(exchange_range(e, e.xch.vtail));  // post-LN + projector gradients are final
  constexpr int kVitGroup = 3;  // SigLIP layers per collective (~90 MB)
  for (int l = c.vit_depth - 1; l >= 0; --l) {
    if (l == e.profile_layer) cudaProfilerStart();
    CHECK_RC(vit_layer_backward(e, l, B));
    if (l == e.profile_layer) cudaProfilerStop();
    if (l % kVitGroup == 0 && e.xch.comm != nullptr && e.xch.chunked) {
      const int top = std::min(l + kVitGroup - 1, c.vit_depth - 1);  // layers [l, top] are contiguous in the arena
      GRange r = e.xch.vit[l];
      for (int j = l + 1; j <= top; ++j) {
        if (e.xch.vit[j].lo < r.lo) r.lo = e.xch.vit[j].lo;
        if (e.xch.vit[j].hi > r.hi) r.hi = e.xch.vit[j].hi;
      }
      CHECK_RC(exchange_range(e, r));
    }
  }
  // patch embedding: fp32 weight / bias / position-embedding gradients (modeling_siglip.py:271-282)
  if (e.batch_copy.patch_rows != nullptr) {
    // dW = dY^T X with X = hi + lo (the first two column blocks of the patch rows): one MN-major tcgen05 GEMM into
    // [W, 2*Kp] fp32, folded into the fp32 [W, 3*p*p] gradient
    GemmArgs g = mk_gemm(W, 2 * e.Kp, Mv, e.g_x1, W, e.batch_copy.patch_rows, 3 * e.Kp, e.g_patch_dw, 2 * e.Kp, EPI_F32);
    g.a_major = 1;
    g.b_major = 1;
    CHECK_RC(engine_gemm(e, g));
    fold_patch_dw(e.g_patch_dw, e.patch_w.g<float>(), W, 3 * c.vit_patch * c.vit_patch, e.Kp, st);
    patch_embed_bwd_pos_bias(e.g_x1, e.patch_b.g<float>(), e.pos_emb.g<float>(), nimg, T, W, st);
  } else {
    patch_embed_bwd(e.batch_copy.images, e.g_x1, e.patch_w.g<float>(), e.patch_b.g<float>(), e.pos_emb.g<float>(), nullptr,
                    nimg, c.image_size, c.vit_patch, W, st);
  }
  CHECK_RC(exchange_range(e, e.xch.f32_vis));
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
int engine_backward(Engine& e, const float* dloss, cudaStream_t st) {
  if (!e.train || !e.bound || e.B <= 0) {
    snprintf(e.err, sizeof(e.err), "pi05_backward: engine not in training mode or no forward recorded");
    set_error(e.err);
    return 8;
  }
  if (e.embed.grad == nullptr) {
    snprintf(e.err, sizeof(e.err), "pi05_backward: gradient buffers not bound");
    set_error(e.err);
    return 8;
  }
  e.stream = st;
  exchange_begin(e);
  const pi05_config& c = e.cfg;
  const int B = e.B, A = e.A, E = e.E, D = e.D, ad = c.action_dim, depth = c.paligemma.depth;
  const int M2 = B * A;
  const int nmods = 2 * depth + 1;
  const int64_t ms = static_cast<int64_t>(B) * 3 * E;
  // ---- head: loss -> v_t -> action_out_proj -> suffix_out
  float* d_vh_in = nullptr;
  if (e.adv_mode) {
    // AdvantageEstimator (pi0_pytorch.py:560-587): dloss is [B, A]; the value head hangs off suffix_out[:, 0]
    const int64_t BE = static_cast<int64_t>(B) * E;
    float *dpre = e.g_f32c, *ds2 = e.g_f32c + BE, *dh2 = e.g_f32c + 2 * BE, *ds1 = e.g_f32c + 3 * BE,
          *dh1 = e.g_f32c + 4 * BE;
    d_vh_in = e.g_f32c + 5 * BE;
    advantage_loss_bwd(e.u_t, e.v_t, e.vh_val, e.vh_prog, dloss, e.w_action, e.w_value, e.g_f32a, dpre, B, A, ad, st);
    linear_f32_wgrad(dpre, e.vh_s2, e.vh4_w.g<float>(), e.vh4_b.g<float>(), B, 1, E, st);
    linear_f32_dgrad(dpre, e.vh4_w.d<float>(), ds2, B, 1, E, 0, st);
    silu_bwd(ds2, e.vh_h2, dh2, BE, st);
    linear_f32_wgrad(dh2, e.vh_s1, e.vh2_w.g<float>(), e.vh2_b.g<float>(), B, E, E, st);
    linear_f32_dgrad(dh2, e.vh2_w.d<float>(), ds1, B, E, E, 0, st);
    silu_bwd(ds1, e.vh_h1, dh1, BE, st);
    linear_f32_wgrad(dh1, e.vh_in, e.vh0_w.g<float>(), e.vh0_b.g<float>(), B, E, E, st);
    linear_f32_dgrad(dh1, e.vh0_w.d<float>(), d_vh_in, B, E, E, 0, st);
  } else {
    flow_loss_bwd(e.u_t, e.v_t, dloss, e.g_f32a, static_cast<int64_t>(M2) * ad, st);  // dv
    if (e.cfg.value_head) {  // plain PI0 loss on an engine that carries a value head: the head gets zero gradients
      fill_zero(e.vh0_w.grad, static_cast<size_t>(E) * E * sizeof(float), st);
      fill_zero(e.vh0_b.grad, static_cast<size_t>(E) * sizeof(float), st);
      fill_zero(e.vh2_w.grad, static_cast<size_t>(E) * E * sizeof(float), st);
      fill_zero(e.vh2_b.grad, static_cast<size_t>(E) * sizeof(float), st);
      fill_zero(e.vh4_w.grad, static_cast<size_t>(E) * sizeof(float), st);
      fill_zero(e.vh4_b.grad, sizeof(float), st);
    }
  }
  linear_f32_wgrad(e.g_f32a, e.so32, e.aout_w.g<float>(), e.aout_b.g<float>(), M2, ad, E, st);
  linear_f32_dgrad(e.g_f32a, e.aout_w.d<float>(), e.g_f32b, M2, ad, E, 0, st);  // d so32
  cast_f32_to_bf16(e.g_f32b, e.g_x2b, static_cast<int64_t>(M2) * E, st);         // grad of the .to(float32) cast
  if (d_vh_in != nullptr) add_row0_grad(e.g_x2b, d_vh_in, B, A, E, st);          // + grad of suffix_out[:, 0].float()
  fill_zero(e.g_dmods, static_cast<size_t>(nmods) * ms * sizeof(float), st);
  const bf16* x2f = depth > 0 ? e.a2[depth - 1].x_out : e.a2[0].x_in;
  rmsnorm_bwd(e.g_x2b, x2f, nullptr, e.mods + (2 * depth) * ms, A, e.rstd_f2, nullptr, e.g_x2, nullptr,
              e.g_dmods + (2 * depth) * ms, M2, E, st);
  fill_zero(e.pg_norm_w.grad, static_cast<size_t>(D) * sizeof(float), st);  // final prefix norm never reaches the loss
  // ---- transformer layers
  for (int l = depth - 1; l >= 0; --l) {
    if (l == e.profile_layer) cudaProfilerStart();
    CHECK_RC(joint_layer_backward(e, l, B, /*g1_zero=*/l == depth - 1));
    if (l == e.profile_layer) cudaProfilerStop();
    // this layer's bf16 weight gradients (both streams) are final: hand them to the exchange stream
    CHECK_RC(exchange_range(e, e.xch.ex[l]));
    CHECK_RC(exchange_range(e, e.xch.pg[l]));
  }
  if (depth == 0) fill_zero(e.g_x1, static_cast<size_t>(B) * e.P * D * 2, st);
  // ---- suffix front-end: action_in_proj, adaRMS dense layers, time MLP
  cast_bf16_to_f32(e.g_x2, e.g_f32a, static_cast<int64_t>(M2) * E, st);  // grad through the bf16 cast of suffix_embs
  linear_f32_wgrad(e.g_f32a, e.x_t, e.ain_w.g<float>(), e.ain_b.g<float>(), M2, E, ad, st);
  float* dcond = e.g_f32b;
  fill_zero(dcond, static_cast<size_t>(B) * E * sizeof(float), st);
  if (e.ada_uniform && e.ada_uniform_grad && depth > 0) {
    linear_f32_wgrad_batched(e.g_dmods, e.cond, e.ex[0].in_dw.g<float>(), e.ex[0].in_db.g<float>(), B, 3 * E, E, nmods, ms,
                             e.ada_wstride, e.ada_bstride, st);
    linear_f32_dgrad_batched_sum(e.g_dmods, e.ex[0].in_dw.d<float>(), dcond, e.g_dcond_part, B, 3 * E, E, nmods, ms,
                                 e.ada_wstride, st);
  } else
  for (int j = 0; j < nmods; ++j) {
    const PRef& dw = (j == 2 * depth) ? e.ex_norm_dw : ((j & 1) ? e.ex[j / 2].post_dw : e.ex[j / 2].in_dw);
    const PRef& db = (j == 2 * depth) ? e.ex_norm_db : ((j & 1) ? e.ex[j / 2].post_db : e.ex[j / 2].in_db);
    const float* dm = e.g_dmods + j * ms;
    linear_f32_wgrad(dm, e.cond, dw.g<float>(), db.g<float>(), B, 3 * E, E, st);
    linear_f32_dgrad(dm, dw.d<float>(), dcond, B, 3 * E, E, /*accumulate=*/1, st);
  }
  silu_bwd(dcond, e.t2, e.g_f32c, static_cast<int64_t>(B) * E, st);  // d t2
  linear_f32_wgrad(e.g_f32c, e.t1s, e.tout_w.g<float>(), e.tout_b.g<float>(), B, E, E, st);
  linear_f32_dgrad(e.g_f32c, e.tout_w.d<float>(), e.g_f32a, B, E, E, 0, st);  // d t1s
  silu_bwd(e.g_f32a, e.t1, e.g_f32c, static_cast<int64_t>(B) * E, st);        // d t1
  linear_f32_wgrad(e.g_f32c, e.temb, e.tin_w.g<float>(), e.tin_b.g<float>(), B, E, E, st);
  CHECK_RC(exchange_range(e, e.xch.f32_main));  // norms, adaRMS dense, heads: every fp32 gradient but the patch embedding
  // ---- prefix: token embedding table (dense bf16 gradient, padding_idx 0 excluded) and the vision tower
  fill_zero(e.embed.grad, static_cast<size_t>(c.vocab_size) * D * 2, st);
  embed_tokens_bwd(e.batch_copy.tokens, e.g_x1, static_cast<int64_t>(e.P) * D, e.NI * e.T, e.embed.g<bf16>(),
                   e.g_embed_scratch, e.g_first, B, e.L, D, static_cast<float>(sqrt(static_cast<double>(D))), st);
  CHECK_RC(exchange_range(e, e.xch.embed));
  CHECK_RC(vision_backward(e, B));
  CHECK_RC(exchange_finish(e));
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(e.err, sizeof(e.err), "pi05_backward: %s", cudaGetErrorString(ce));
    set_error(e.err);
    return 9;
  }
  return 0;
}

}  // namespace pi05
