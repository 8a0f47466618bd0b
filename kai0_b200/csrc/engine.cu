// pi0.5 engine: workspace plan, parameter binding, and the training-time forward pass.
//
// Follows, stage by stage (paths relative to /root/reference/src/openpi/models_pytorch/):
//   pi0_pytorch.py:316-373 (forward), :186-235 (embed_prefix), :237-314 (embed_suffix, pi05 branch),
//   gemma_pytorch.py:158-238 (joint layer), :262-275 (final norms),
//   transformers_replace/models/siglip/modeling_siglip.py:271-282,435-481,763-796 (vision tower),
//   transformers_replace/models/paligemma/modeling_paligemma.py:91-99,232-247 (projector).
// Every bf16 GEMM runs on the tcgen05 kernel of gemm_sm100.cu; everything else is in *_kernels.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include <cuda_profiler_api.h>

#include "engine.h"
#include "errors.h"
#include "gemm.h"

namespace pi05 {

#define CHECK_RC(x)        \
  do {                     \
    int _rc = (x);         \
    if (_rc != 0) return _rc; \
  } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

int engine_gemm(Engine& e, const GemmArgs& a) {
  char err[512] = "";
  GemmArgs b = a;
  b.splitk_ws = e.splitk_ws;
  b.splitk_ws_bytes = e.splitk_ws_bytes;
  int rc = gemm_bf16(b, e.stream, err, sizeof(err));
  if (rc != 0) {
    snprintf(e.err, sizeof(e.err), "%s", err);
    set_error(e.err);
  }
  return rc;
}

GemmArgs mk_gemm(int M, int N, int K, const void* A, int64_t lda, const void* Bm, int64_t ldb, void* D, int64_t ldd,
                 int epi) {
  GemmArgs g;
  g.M = M;
  g.N = N;
  g.K = K;
  g.A = A;
  g.lda = lda;
  g.B = Bm;
  g.ldb = ldb;
  g.D = D;
  g.ldd = ldd;
  g.epilogue = epi;
  return g;
}

void add_tap(Engine& e, const char* name, const void* p, int64_t n, int dtype) {
  if (e.taps_enabled) e.taps[name] = Tap{p, n, dtype};
}

// ------------------------------------------------------------------------------------------------------------
// workspace plan (dry = count bytes only)
// ------------------------------------------------------------------------------------------------------------
int engine_plan(Engine& e, bool dry) {
  const pi05_config& c = e.cfg;
  e.T = (c.image_size / c.vit_patch) * (c.image_size / c.vit_patch);
  e.NI = c.num_images;
  e.L = c.max_token_len;
  e.Lmax = c.max_token_len;
  e.P = e.NI * e.T + e.L;
  e.A = c.action_horizon;
  e.S = e.P + e.A;
  e.Ppad = round_up(e.P, 8);
  e.Spad = round_up(e.S, 8);
  e.D = c.paligemma.width;
  e.E = c.expert.width;
  e.W = c.vit_width;
  e.H = c.paligemma.num_heads;
  e.hd = c.paligemma.head_dim;
  e.VH = c.vit_heads;
  e.vhd = c.vit_width / c.vit_heads;
  e.Bmax = c.max_batch;
  e.train = c.train != 0;

  Arena& ar = e.arena;
  ar.dry = dry;
  ar.off = 0;
  ar.overflow = false;
  const int64_t B = e.Bmax, T = e.T, P = e.P, A = e.A, S = e.S, D = e.D, E = e.E, W = e.W, H = e.H, hd = e.hd;
  const int64_t nimg = e.NI * B, Mv = nimg * T, M1 = B * P, M2 = B * A;
  const int64_t QW = (H + 2) * hd;
  const int depth = c.paligemma.depth;
  const int vdepth = c.vit_depth;
  const int64_t vmlp = c.vit_mlp_dim, mlp1 = c.paligemma.mlp_dim, mlp2 = c.expert.mlp_dim;
  const bool tr = e.train;
  const bool rtc = !tr && c.rtc != 0;  // inference engine that also runs pi05_denoise_rtc: suffix stash per layer

  e.rope_cos = ar.get<bf16>((S + 1) * (hd / 2));
  e.rope_sin = ar.get<bf16>((S + 1) * (hd / 2));
  e.time_scaling = ar.get<double>(E / 2);
  e.pad = ar.get<uint8_t>(B * P);
  e.pos = ar.get<int>(B * P);
  e.nvalid = ar.get<int>(B);

  // ---- vision tower
  e.vit_x0 = ar.get<bf16>(Mv * W);
  e.Kp = patch_row_kp(c.vit_patch);
  e.patch_wb = ar.get<bf16>(W * 3 * e.Kp);
  if (tr) e.g_patch_dw = ar.get<float>(W * 2 * e.Kp);
  e.va.resize(vdepth);
  VitLayerA shared{};
  for (int l = 0; l < vdepth; ++l) {
    VitLayerA& a = e.va[l];
    if (tr || l == 0) {
      a.h1 = ar.get<bf16>(Mv * W);
      a.qkv = ar.get<bf16>(Mv * 3 * W);
      a.P = ar.get<bf16>(nimg * e.VH * T * T);
      a.attn = ar.get<bf16>(Mv * W);
      a.h2 = ar.get<bf16>(Mv * W);
      a.pre = ar.get<bf16>(Mv * vmlp);
      a.act = ar.get<bf16>(Mv * vmlp);
      a.mean1 = ar.get<float>(Mv);
      a.rstd1 = ar.get<float>(Mv);
      a.mean2 = ar.get<float>(Mv);
      a.rstd2 = ar.get<float>(Mv);
      if (l == 0) shared = a;
    } else {
      a = shared;
    }
    a.x_in = (l == 0) ? e.vit_x0 : e.va[l - 1].x_out;
    a.x_mid = ar.get<bf16>(Mv * W);
    a.x_out = ar.get<bf16>(Mv * W);
  }
  e.vit_post = ar.get<bf16>(Mv * W);
  e.vit_post_mean = ar.get<float>(Mv);
  e.vit_post_rstd = ar.get<float>(Mv);

  // ---- the two Gemma streams
  e.a1.resize(depth);
  e.a2.resize(depth);
  e.Kl.resize(depth);
  e.Vl.resize(depth);
  bf16* x1_0 = ar.get<bf16>(M1 * D);  // prefix_embs
  bf16* x2_0 = ar.get<bf16>(M2 * E);  // suffix_embs
  GemmaLayerA sh1{}, sh2{};
  for (int l = 0; l < depth; ++l) {
    GemmaLayerA &p1 = e.a1[l], &p2 = e.a2[l];
    if (rtc && l > 0) {  // own suffix buffers per layer (M2 rows: ~4 MB each), prefix buffers shared
      p1 = sh1;
      p2.n1 = ar.get<bf16>(M2 * E);
      p2.qkv = ar.get<bf16>(M2 * QW);
      p2.Q = ar.get<bf16>(M2 * H * hd);
      p2.P = ar.get<bf16>(B * A * H * e.Spad);
      p2.O = ar.get<bf16>(M2 * H * hd);
      p2.o_lin = ar.get<bf16>(M2 * E);
      p2.n2 = ar.get<bf16>(M2 * E);
      p2.GU = ar.get<bf16>(M2 * 2 * mlp2);
      p2.Hh = ar.get<bf16>(M2 * mlp2);
      p2.d_lin = ar.get<bf16>(M2 * E);
      p2.gate1 = ar.get<bf16>(B * E);
      p2.gate2 = ar.get<bf16>(B * E);
      p2.rstd1 = ar.get<float>(M2);
      p2.rstd2 = ar.get<float>(M2);
    } else if (tr || l == 0) {
      p1.n1 = ar.get<bf16>(M1 * D);
      p1.qkv = ar.get<bf16>(M1 * QW);
      p1.Q = ar.get<bf16>(M1 * H * hd);
      p1.P = ar.get<bf16>(B * P * H * e.Ppad);
      p1.O = ar.get<bf16>(M1 * H * hd);
      p1.n2 = ar.get<bf16>(M1 * D);
      p1.GU = ar.get<bf16>(M1 * 2 * mlp1);
      p1.Hh = ar.get<bf16>(M1 * mlp1);
      p1.rstd1 = ar.get<float>(M1);
      p1.rstd2 = ar.get<float>(M1);
      p1.o_lin = nullptr;
      p1.d_lin = nullptr;
      p1.gate1 = nullptr;
      p1.gate2 = nullptr;
      p2.n1 = ar.get<bf16>(M2 * E);
      p2.qkv = ar.get<bf16>(M2 * QW);
      p2.Q = ar.get<bf16>(M2 * H * hd);
      p2.P = ar.get<bf16>(B * A * H * e.Spad);
      p2.O = ar.get<bf16>(M2 * H * hd);
      p2.o_lin = ar.get<bf16>(M2 * E);
      p2.n2 = ar.get<bf16>(M2 * E);
      p2.GU = ar.get<bf16>(M2 * 2 * mlp2);
      p2.Hh = ar.get<bf16>(M2 * mlp2);
      p2.d_lin = ar.get<bf16>(M2 * E);
      p2.gate1 = ar.get<bf16>(B * E);
      p2.gate2 = ar.get<bf16>(B * E);
      p2.rstd1 = ar.get<float>(M2);
      p2.rstd2 = ar.get<float>(M2);
      if (l == 0) {
        sh1 = p1;
        sh2 = p2;
      }
    } else {
      p1 = sh1;
      p2 = sh2;
    }
    p1.x_in = (l == 0) ? x1_0 : e.a1[l - 1].x_out;
    p2.x_in = (l == 0) ? x2_0 : e.a2[l - 1].x_out;
    p1.x_mid = ar.get<bf16>(M1 * D);
    p1.x_out = ar.get<bf16>(M1 * D);
    p2.x_mid = ar.get<bf16>(M2 * E);
    p2.x_out = ar.get<bf16>(M2 * E);
    e.Kl[l] = ar.get<bf16>(B * S * hd);
    e.Vl[l] = ar.get<bf16>(B * S * hd);
  }
  e.prefix_out = ar.get<bf16>(M1 * D);
  e.suffix_out = ar.get<bf16>(M2 * E);
  e.rstd_f1 = ar.get<float>(M1);
  e.rstd_f2 = ar.get<float>(M2);

  // ---- suffix front-end / head (fp32)
  const int nmods = 2 * depth + 1;
  e.x_t = ar.get<float>(M2 * c.action_dim);
  e.u_t = ar.get<float>(M2 * c.action_dim);
  e.temb = ar.get<float>(B * E);
  e.aemb32 = ar.get<float>(M2 * E);
  e.t1 = ar.get<float>(B * E);
  e.t1s = ar.get<float>(B * E);
  e.t2 = ar.get<float>(B * E);
  e.cond = ar.get<float>(B * E);
  e.mods = ar.get<float>(nmods * B * 3 * E);
  e.so32 = ar.get<float>(M2 * E);
  e.v_t = ar.get<float>(M2 * c.action_dim);
  e.timevec = ar.get<float>(B);
  if (c.value_head) {
    e.vh_in = ar.get<float>(B * E);
    e.vh_h1 = ar.get<float>(B * E);
    e.vh_s1 = ar.get<float>(B * E);
    e.vh_h2 = ar.get<float>(B * E);
    e.vh_s2 = ar.get<float>(B * E);
    e.vh_h3 = ar.get<float>(B);
    e.vh_val = ar.get<float>(B);
    e.vh_prog = ar.get<float>(B);
    e.vh_la = ar.get<float>(M2);
    e.vh_lv = ar.get<float>(B);
  }
  // small-M split-K (decode) needs a few MB; the training wave-quantisation path up to s * M * N fp32 partials of a
  // weight-gradient GEMM (largest user: 4 x 2560 x 2048 x 4 B = 84 MB)
  e.splitk_ws_bytes = static_cast<size_t>(tr ? 128 : 64) << 20;
  e.splitk_ws = ar.get<float>(e.splitk_ws_bytes / sizeof(float));
  {
    const int64_t ns = Engine::kMaxDecodeSteps;
    e.dec_times = ar.get<float>(ns);
    e.dec_temb = ar.get<float>(ns * E);
    e.dec_t1 = ar.get<float>(ns * E);
    e.dec_t1s = ar.get<float>(ns * E);
    e.dec_t2 = ar.get<float>(ns * E);
    e.dec_cond = ar.get<float>(ns * E);
    e.dec_mods = ar.get<float>(static_cast<int64_t>(nmods) * ns * 3 * E);
  }

  // ---- backward scratch
  if (tr) {
    auto mx = [](int64_t a, int64_t b) { return a > b ? a : b; };
    const int64_t r1 = mx(M1 * D, Mv * W);
    e.g_x1 = ar.get<bf16>(r1);
    e.g_x1b = ar.get<bf16>(r1);
    e.g_x2 = ar.get<bf16>(M2 * E);
    e.g_x2b = ar.get<bf16>(M2 * E);
    e.g_big = ar.get<bf16>(mx(mx(M1 * 2 * mlp1, Mv * vmlp), M2 * 2 * mlp2));
    e.g_big2 = ar.get<bf16>(mx(mx(M1 * mlp1, Mv * vmlp), M2 * mlp2));
    const int64_t rt = mx(mx(M1 * QW, Mv * 3 * W), mx(M1 * D, M1 * H * hd));
    e.g_t1 = ar.get<bf16>(rt);
    e.g_t2 = ar.get<bf16>(rt);
    e.g_t3 = ar.get<bf16>(rt);
    e.g_P = ar.get<bf16>(mx(mx(B * P * H * e.Ppad, nimg * e.VH * T * T), B * A * H * e.Spad));
    e.g2_do = ar.get<bf16>(M2 * E);
    e.g2_big = ar.get<bf16>(M2 * 2 * mlp2);
    e.g2_big2 = ar.get<bf16>(M2 * mlp2);
    e.g2_t1 = ar.get<bf16>(M2 * mx(QW, E));
    e.g2_t2 = ar.get<bf16>(M2 * mx(H * hd, E));
    e.g2_t3 = ar.get<bf16>(M2 * H * hd);
    e.g_dK = ar.get<float>(B * S * hd);
    e.g_dV = ar.get<float>(B * S * hd);
    e.g_dmods = ar.get<float>(nmods * B * 3 * E);
    e.g_f32a = ar.get<float>(mx(M2 * E, B * 3 * E));
    e.g_f32b = ar.get<float>(mx(M2 * E, B * 3 * E));
    e.g_f32c = ar.get<float>(mx(M2 * E, B * 6 * E));  // also the value head's backward scratch (6 x [B, E])
    e.g_dcond_part = ar.get<float>(nmods * B * E);
    e.g_acc_elems = static_cast<size_t>(mx(mx(4 * W + 2 * vmlp, 4 * D), mx(T * W + W, 3 * W * 14 * 14 * 3)) + 1024);
    e.g_acc = ar.get<float>(e.g_acc_elems);
    e.g_embed_scratch = ar.get<float>(B * e.L * D);
    e.g_first = ar.get<int>(B * e.L);
  }
  if (rtc) {  // the suffix half of the backward scratch (engine_rtc.cu)
    auto mx = [](int64_t a, int64_t b) { return a > b ? a : b; };
    e.g_x2 = ar.get<bf16>(M2 * E);
    e.g_x2b = ar.get<bf16>(M2 * E);
    e.g_P = ar.get<bf16>(B * A * H * e.Spad);
    e.g2_do = ar.get<bf16>(M2 * E);
    e.g2_big = ar.get<bf16>(M2 * 2 * mlp2);
    e.g2_t1 = ar.get<bf16>(M2 * mx(QW, E));
    e.g2_t2 = ar.get<bf16>(M2 * mx(H * hd, E));
    e.g2_t3 = ar.get<bf16>(M2 * H * hd);
    e.g_dK = ar.get<float>(B * S * hd);
    e.g_dV = ar.get<float>(B * S * hd);
    e.g_dmods = ar.get<float>(B * 3 * E);
    e.g_f32a = ar.get<float>(M2 * E);
    e.g_f32b = ar.get<float>(M2 * E);
    e.g_f32c = ar.get<float>(M2 * E);
    e.rtc_tap_v = ar.get<float>(M2 * c.action_dim);
    e.rtc_tap_j = ar.get<float>(M2 * c.action_dim);
  }
  ar.alloc(256);
  if (!dry && ar.overflow) {
    snprintf(e.err, sizeof(e.err), "workspace too small: need %zu bytes, have %zu", ar.off, ar.cap);
    set_error(e.err);
    return 5;
  }
  return 0;
}

// pi05_batch.token_len: the prompt length of THIS batch (<= cfg.max_token_len).  Every extent and stride that depends on the
// prefix length (P, S and their 8-element pitches) is a run-time value read at launch time; the workspace was planned for
// the maximum, so a shorter prefix uses the leading part of each buffer.
int engine_set_token_len(Engine& e, int token_len, const char* who) {
  const int L = token_len > 0 ? token_len : e.Lmax;
  if (L > e.Lmax) {
    snprintf(e.err, sizeof(e.err), "%s: pi05_batch.token_len %d exceeds max_token_len %d", who, L, e.Lmax);
    set_error(e.err);
    return 8;
  }
  e.L = L;
  e.P = e.NI * e.T + L;
  e.S = e.P + e.A;
  e.Ppad = round_up(e.P, 8);
  e.Spad = round_up(e.S, 8);
  return 0;
}

// Host-computed constant tables, uploaded once at create time.
int engine_init_tables(Engine& e, cudaStream_t st) {
  const int half = e.hd / 2;
  const int rows = e.S + 1;
  std::vector<__nv_bfloat16> hc(static_cast<size_t>(rows) * half), hs(static_cast<size_t>(rows) * half);
  for (int i = 0; i < half; ++i) {
    // modeling_gemma.py:129-160 with rope_type "default": inv_freq = 1 / theta^(2i/hd), all in fp32
    const float ex = static_cast<float>(2 * i) / static_cast<float>(e.hd);
    const float inv = 1.0f / powf(10000.0f, ex);
    for (int r = 0; r < rows; ++r) {
      const float ang = inv * static_cast<float>(r - 1);
      hc[static_cast<size_t>(r) * half + i] = __float2bfloat16_rn(cosf(ang));
      hs[static_cast<size_t>(r) * half + i] = __float2bfloat16_rn(sinf(ang));
    }
  }
  cudaMemcpyAsync(e.rope_cos, hc.data(), hc.size() * 2, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(e.rope_sin, hs.data(), hs.size() * 2, cudaMemcpyHostToDevice, st);
  // pi0_pytorch.py:25-42: fraction = linspace(0,1,E/2) (fp64); period = 4e-3 * (4/4e-3)^fraction; 1/period*2*pi
  const int th = e.E / 2;
  std::vector<double> sc(th);
  const double step = th > 1 ? 1.0 / static_cast<double>(th - 1) : 0.0;
  for (int i = 0; i < th; ++i) {
    const double frac = (i < th / 2) ? i * step : 1.0 - (th - 1 - i) * step;  // torch.linspace's symmetric form
    const double period = 4e-3 * pow(4.0 / 4e-3, frac);
    sc[i] = 1.0 / period * 2 * M_PI;
  }
  cudaMemcpyAsync(e.time_scaling, sc.data(), sc.size() * sizeof(double), cudaMemcpyHostToDevice, st);
  cudaError_t ce = cudaStreamSynchronize(st);  // host vectors go out of scope
  if (ce != cudaSuccess) {
    snprintf(e.err, sizeof(e.err), "table upload: %s", cudaGetErrorString(ce));
    set_error(e.err);
    return 6;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// parameter resolution (reference state_dict names; fused-weight contiguity is required and checked)
// ------------------------------------------------------------------------------------------------------------
static bool get_param(Engine& e, const std::string& name, int dtype, int64_t numel, PRef* out, std::string* missing) {
  auto it = e.params.find(name);
  if (it == e.params.end()) {
    if (missing->size() < 600) *missing += name + " ";
    return false;
  }
  if (it->second.dtype != dtype || it->second.numel != numel) {
    if (missing->size() < 600)
      *missing += name + "(dtype/numel " + std::to_string(it->second.dtype) + "/" + std::to_string(it->second.numel) +
                  " want " + std::to_string(dtype) + "/" + std::to_string(numel) + ") ";
    return false;
  }
  *out = it->second;
  return true;
}

static bool contiguous_after(const PRef& a, const PRef& b, int elem) {
  const bool d = static_cast<char*>(a.data) + a.numel * elem == static_cast<char*>(b.data);
  const bool g = (a.grad == nullptr && b.grad == nullptr) ||
                 (a.grad && b.grad && static_cast<char*>(a.grad) + a.numel * elem == static_cast<char*>(b.grad));
  return d && g;
}

int engine_resolve_params(Engine& e) {
  const pi05_config& c = e.cfg;
  std::string miss, contig;
  const std::string PWE = "paligemma_with_expert.";
  const std::string VT = PWE + "paligemma.model.vision_tower.vision_model.";
  const std::string LM = PWE + "paligemma.model.language_model.";
  const std::string EX = PWE + "gemma_expert.model.";
  const int64_t W = e.W, D = e.D, E = e.E, pp = c.vit_patch * c.vit_patch * 3;
  const int F = PI05_F32, BF = PI05_BF16;
  get_param(e, VT + "embeddings.patch_embedding.weight", F, W * pp, &e.patch_w, &miss);
  get_param(e, VT + "embeddings.patch_embedding.bias", F, W, &e.patch_b, &miss);
  get_param(e, VT + "embeddings.position_embedding.weight", F, e.T * W, &e.pos_emb, &miss);
  e.vit.resize(c.vit_depth);
  for (int l = 0; l < c.vit_depth; ++l) {
    const std::string Lp = VT + "encoder.layers." + std::to_string(l) + ".";
    VitLayerP& p = e.vit[l];
    get_param(e, Lp + "layer_norm1.weight", BF, W, &p.ln1_w, &miss);
    get_param(e, Lp + "layer_norm1.bias", BF, W, &p.ln1_b, &miss);
    get_param(e, Lp + "layer_norm2.weight", BF, W, &p.ln2_w, &miss);
    get_param(e, Lp + "layer_norm2.bias", BF, W, &p.ln2_b, &miss);
    get_param(e, Lp + "self_attn.q_proj.weight", BF, W * W, &p.q_w, &miss);
    get_param(e, Lp + "self_attn.k_proj.weight", BF, W * W, &p.k_w, &miss);
    get_param(e, Lp + "self_attn.v_proj.weight", BF, W * W, &p.v_w, &miss);
    get_param(e, Lp + "self_attn.q_proj.bias", BF, W, &p.q_b, &miss);
    get_param(e, Lp + "self_attn.k_proj.bias", BF, W, &p.k_b, &miss);
    get_param(e, Lp + "self_attn.v_proj.bias", BF, W, &p.v_b, &miss);
    get_param(e, Lp + "self_attn.out_proj.weight", BF, W * W, &p.out_w, &miss);
    get_param(e, Lp + "self_attn.out_proj.bias", BF, W, &p.out_b, &miss);
    get_param(e, Lp + "mlp.fc1.weight", BF, static_cast<int64_t>(c.vit_mlp_dim) * W, &p.fc1_w, &miss);
    get_param(e, Lp + "mlp.fc1.bias", BF, c.vit_mlp_dim, &p.fc1_b, &miss);
    get_param(e, Lp + "mlp.fc2.weight", BF, static_cast<int64_t>(c.vit_mlp_dim) * W, &p.fc2_w, &miss);
    get_param(e, Lp + "mlp.fc2.bias", BF, W, &p.fc2_b, &miss);
    if (p.q_w.data && p.k_w.data && p.v_w.data &&
        !(contiguous_after(p.q_w, p.k_w, 2) && contiguous_after(p.k_w, p.v_w, 2) && contiguous_after(p.q_b, p.k_b, 2) &&
          contiguous_after(p.k_b, p.v_b, 2)))
      contig += Lp + "self_attn.{q,k,v}_proj ";
  }
  get_param(e, VT + "post_layernorm.weight", BF, W, &e.post_ln_w, &miss);
  get_param(e, VT + "post_layernorm.bias", BF, W, &e.post_ln_b, &miss);
  get_param(e, PWE + "paligemma.model.multi_modal_projector.linear.weight", BF, D * W, &e.proj_w, &miss);
  get_param(e, PWE + "paligemma.model.multi_modal_projector.linear.bias", BF, D, &e.proj_b, &miss);
  get_param(e, LM + "embed_tokens.weight", BF, static_cast<int64_t>(c.vocab_size) * D, &e.embed, &miss);

  for (int s = 0; s < 2; ++s) {
    const pi05_gemma_cfg& g = s == 0 ? c.paligemma : c.expert;
    std::vector<GemmaLayerP>& Lv = s == 0 ? e.pg : e.ex;
    const std::string pre = s == 0 ? LM : EX;
    const int64_t w = g.width, hq = static_cast<int64_t>(g.num_heads) * g.head_dim,
                  hk = static_cast<int64_t>(g.num_kv_heads) * g.head_dim;
    Lv.resize(g.depth);
    for (int l = 0; l < g.depth; ++l) {
      const std::string Lp = pre + "layers." + std::to_string(l) + ".";
      GemmaLayerP& p = Lv[l];
      get_param(e, Lp + "self_attn.q_proj.weight", BF, hq * w, &p.q_w, &miss);
      get_param(e, Lp + "self_attn.k_proj.weight", BF, hk * w, &p.k_w, &miss);
      get_param(e, Lp + "self_attn.v_proj.weight", BF, hk * w, &p.v_w, &miss);
      get_param(e, Lp + "self_attn.o_proj.weight", BF, w * hq, &p.o_w, &miss);
      get_param(e, Lp + "mlp.gate_proj.weight", BF, static_cast<int64_t>(g.mlp_dim) * w, &p.gate_w, &miss);
      get_param(e, Lp + "mlp.up_proj.weight", BF, static_cast<int64_t>(g.mlp_dim) * w, &p.up_w, &miss);
      get_param(e, Lp + "mlp.down_proj.weight", BF, static_cast<int64_t>(g.mlp_dim) * w, &p.down_w, &miss);
      if (s == 0) {
        get_param(e, Lp + "input_layernorm.weight", F, w, &p.in_w, &miss);
        get_param(e, Lp + "post_attention_layernorm.weight", F, w, &p.post_w, &miss);
      } else {
        get_param(e, Lp + "input_layernorm.dense.weight", F, 3 * w * w, &p.in_dw, &miss);
        get_param(e, Lp + "input_layernorm.dense.bias", F, 3 * w, &p.in_db, &miss);
        get_param(e, Lp + "post_attention_layernorm.dense.weight", F, 3 * w * w, &p.post_dw, &miss);
        get_param(e, Lp + "post_attention_layernorm.dense.bias", F, 3 * w, &p.post_db, &miss);
      }
      if (p.q_w.data && p.k_w.data && p.v_w.data &&
          !(contiguous_after(p.q_w, p.k_w, 2) && contiguous_after(p.k_w, p.v_w, 2)))
        contig += Lp + "self_attn.{q,k,v}_proj ";
      if (p.gate_w.data && p.up_w.data && !contiguous_after(p.gate_w, p.up_w, 2)) contig += Lp + "mlp.{gate,up}_proj ";
    }
  }
  get_param(e, LM + "norm.weight", F, D, &e.pg_norm_w, &miss);
  get_param(e, EX + "norm.dense.weight", F, 3 * E * E, &e.ex_norm_dw, &miss);
  get_param(e, EX + "norm.dense.bias", F, 3 * E, &e.ex_norm_db, &miss);
  get_param(e, "action_in_proj.weight", F, E * c.action_dim, &e.ain_w, &miss);
  get_param(e, "action_in_proj.bias", F, E, &e.ain_b, &miss);
  get_param(e, "action_out_proj.weight", F, E * c.action_dim, &e.aout_w, &miss);
  get_param(e, "action_out_proj.bias", F, c.action_dim, &e.aout_b, &miss);
  get_param(e, "time_mlp_in.weight", F, E * E, &e.tin_w, &miss);
  get_param(e, "time_mlp_in.bias", F, E, &e.tin_b, &miss);
  get_param(e, "time_mlp_out.weight", F, E * E, &e.tout_w, &miss);
  get_param(e, "time_mlp_out.bias", F, E, &e.tout_b, &miss);
  if (c.value_head) {
    get_param(e, "value_head.0.weight", F, E * E, &e.vh0_w, &miss);
    get_param(e, "value_head.0.bias", F, E, &e.vh0_b, &miss);
    get_param(e, "value_head.2.weight", F, E * E, &e.vh2_w, &miss);
    get_param(e, "value_head.2.bias", F, E, &e.vh2_b, &miss);
    get_param(e, "value_head.4.weight", F, E, &e.vh4_w, &miss);
    get_param(e, "value_head.4.bias", F, 1, &e.vh4_b, &miss);
  }
  if (!miss.empty() || !contig.empty()) {
    snprintf(e.err, sizeof(e.err), "bind_params: missing/mismatched: [%s] not contiguous (fused arenas required): [%s]",
             miss.c_str(), contig.c_str());
    set_error(e.err);
    return 7;
  }
  if (c.paligemma.num_kv_heads != 1 || c.expert.num_kv_heads != 1 || c.paligemma.head_dim != c.expert.head_dim ||
      c.paligemma.num_heads != c.expert.num_heads || c.paligemma.depth != c.expert.depth) {
    snprintf(e.err, sizeof(e.err), "unsupported attention geometry (reference hard-codes 8 q heads / 1 kv head)");
    set_error(e.err);
    return 7;
  }
  {  // are the 2*depth+1 adaRMS dense weights / biases (and their grads) uniformly strided?  (arena order makes them so)
    const int depth = c.paligemma.depth, nm = 2 * depth + 1;
    auto W = [&](int j) -> const PRef& { return j == 2 * depth ? e.ex_norm_dw : ((j & 1) ? e.ex[j / 2].post_dw : e.ex[j / 2].in_dw); };
    auto Bz = [&](int j) -> const PRef& { return j == 2 * depth ? e.ex_norm_db : ((j & 1) ? e.ex[j / 2].post_db : e.ex[j / 2].in_db); };
    e.ada_uniform = nm >= 2;
    e.ada_uniform_grad = nm >= 2 && W(0).grad != nullptr;
    if (nm >= 2) {
      e.ada_wstride = W(1).d<float>() - W(0).d<float>();
      e.ada_bstride = Bz(1).d<float>() - Bz(0).d<float>();
      for (int j = 0; j < nm; ++j) {
        if (W(j).d<float>() != W(0).d<float>() + j * e.ada_wstride || Bz(j).d<float>() != Bz(0).d<float>() + j * e.ada_bstride)
          e.ada_uniform = false;
        if (e.ada_uniform_grad && (W(j).g<float>() != W(0).g<float>() + j * e.ada_wstride ||
                                   Bz(j).g<float>() != Bz(0).g<float>() + j * e.ada_bstride))
          e.ada_uniform_grad = false;
      }
    }
  }
  // ---- contiguous gradient groups for the overlapped data-parallel exchange (exchange.h), in arena order; verified to
  // tile [first trainable bf16 gradient, end of the last one) / the fp32 arena without overlap
  {
    GradExchange& x = e.xch;
    auto span = [](std::initializer_list<const PRef*> ps, int dtype) {
      GRange r;
      r.dtype = dtype;
      const int esz = dtype == PI05_BF16 ? 2 : 4;
      for (const PRef* p : ps) {
        if (p->grad == nullptr) return GRange{};
        char* lo = static_cast<char*>(p->grad);
        char* hi = lo + p->numel * esz;
        if (r.lo == nullptr || lo < r.lo) r.lo = lo;
        if (hi > r.hi) r.hi = hi;
      }
      return r;
    };
    x.pg.clear();
    x.ex.clear();
    x.vit.clear();
    const int BFt = PI05_BF16, Ft = PI05_F32;
    for (auto& p : e.pg) x.pg.push_back(span({&p.q_w, &p.k_w, &p.v_w, &p.o_w, &p.gate_w, &p.up_w, &p.down_w}, BFt));
    for (auto& p : e.ex) x.ex.push_back(span({&p.q_w, &p.k_w, &p.v_w, &p.o_w, &p.gate_w, &p.up_w, &p.down_w}, BFt));
    for (auto& p : e.vit)
      x.vit.push_back(span({&p.ln1_w, &p.ln1_b, &p.ln2_w, &p.ln2_b, &p.q_w, &p.k_w, &p.v_w, &p.q_b, &p.k_b, &p.v_b, &p.out_w,
                            &p.out_b, &p.fc1_w, &p.fc1_b, &p.fc2_w, &p.fc2_b}, BFt));
    x.vtail = span({&e.post_ln_w, &e.post_ln_b, &e.proj_w, &e.proj_b}, BFt);
    x.embed = span({&e.embed}, BFt);
    x.f32_vis = span({&e.patch_w, &e.patch_b, &e.pos_emb}, Ft);
    std::vector<const PRef*> f32;
    for (auto& p : e.pg) { f32.push_back(&p.in_w); f32.push_back(&p.post_w); }
    for (auto& p : e.ex) { f32.push_back(&p.in_dw); f32.push_back(&p.in_db); f32.push_back(&p.post_dw); f32.push_back(&p.post_db); }
    for (const PRef* p : {&e.pg_norm_w, &e.ex_norm_dw, &e.ex_norm_db, &e.ain_w, &e.ain_b, &e.aout_w, &e.aout_b, &e.tin_w,
                          &e.tin_b, &e.tout_w, &e.tout_b})
      f32.push_back(p);
    if (c.value_head)
      for (const PRef* p : {&e.vh0_w, &e.vh0_b, &e.vh2_w, &e.vh2_b, &e.vh4_w, &e.vh4_b}) f32.push_back(p);
    x.f32_main = GRange{};
    x.f32_main.dtype = Ft;
    bool f32_ok = true;
    for (const PRef* p : f32) {
      if (p->grad == nullptr) { f32_ok = false; break; }
      char* lo = static_cast<char*>(p->grad);
      char* hi = lo + p->numel * 4;
      if (x.f32_main.lo == nullptr || lo < x.f32_main.lo) x.f32_main.lo = lo;
      if (hi > x.f32_main.hi) x.f32_main.hi = hi;
    }
    if (!f32_ok) x.f32_main = GRange{};
    // verification: bf16 groups sorted by address must not overlap and may only be separated by alignment padding
    std::vector<GRange> all = x.pg;
    all.insert(all.end(), x.ex.begin(), x.ex.end());
    all.insert(all.end(), x.vit.begin(), x.vit.end());
    all.push_back(x.vtail);
    all.push_back(x.embed);
    bool ok = f32_ok && x.f32_vis.valid();
    for (auto& r : all) ok = ok && r.valid();
    if (ok) {
      std::sort(all.begin(), all.end(), [](const GRange& a, const GRange& b) { return a.lo < b.lo; });
      for (size_t i = 1; i < all.size(); ++i)
        if (all[i].lo < all[i - 1].hi || all[i].lo - all[i - 1].hi > 64) ok = false;
      // the fp32 pieces: vision embeddings first, then everything else, disjoint
      if (!(x.f32_vis.hi <= x.f32_main.lo || x.f32_main.hi <= x.f32_vis.lo)) ok = false;
    }
    x.chunked = ok;
    x.all_bf16 = GRange{};
    x.all_f32 = GRange{};
    if (!all.empty() && all.front().valid() && e.embed.grad != nullptr) {
      x.all_bf16.dtype = BFt;
      x.all_bf16.lo = all.front().lo;
      x.all_bf16.hi = all.back().hi;
      for (auto& r : all) {
        if (r.lo < x.all_bf16.lo) x.all_bf16.lo = r.lo;
        if (r.hi > x.all_bf16.hi) x.all_bf16.hi = r.hi;
      }
    }
    if (x.f32_main.valid() && x.f32_vis.valid()) {
      x.all_f32.dtype = Ft;
      x.all_f32.lo = x.f32_vis.lo < x.f32_main.lo ? x.f32_vis.lo : x.f32_main.lo;
      x.all_f32.hi = x.f32_vis.hi > x.f32_main.hi ? x.f32_vis.hi : x.f32_main.hi;
    }
  }
  e.bound = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// forward pieces
// ------------------------------------------------------------------------------------------------------------
// SigLIP tower + projector for all NI*B images at once (row-independent, so identical to the reference's
// per-camera calls, pi0_pytorch.py:197-202); writes image tokens straight into prefix_embs.
int vision_forward(Engine& e, const float* images, const void* patch_rows, int B, bf16* prefix_embs) {
  const pi05_config& c = e.cfg;
  cudaStream_t st = e.stream;
  const int nimg = e.NI * B, T = e.T, W = e.W, VH = e.VH, vhd = e.vhd;
  const int Mv = nimg * T;
  const float ln_eps = 1e-6f;
  if (patch_rows != nullptr) {
    // Patch embedding (modeling_siglip.py:271-282) on the tensor cores at fp32-class accuracy: the preprocessing kernel
    // already wrote the im2col rows as the split [hi | lo | hi] (pi05_preprocess_patches); with the weight split
    // [Whi | Whi | Wlo] ONE bf16 GEMM accumulates hi*Whi + lo*Whi + hi*Wlo in fp32 (the dropped lo*Wlo term is 2^-16
    // relative), and the epilogue adds the fp32 bias and position embedding before the single rounding to bf16.
    const int k = 3 * c.vit_patch * c.vit_patch;
    split_patch_weight(e.patch_w.d<float>(), e.patch_wb, W, k, e.Kp, st);
    GemmArgs g = mk_gemm(Mv, W, 3 * e.Kp, patch_rows, 3 * e.Kp, e.patch_wb, 3 * e.Kp, e.vit_x0, W, EPI_PATCH);
    g.bias32 = e.patch_b.d<float>();
    g.rowadd32 = e.pos_emb.d<float>();
    g.rowadd_period = T;
    g.ld_rowadd = W;
    CHECK_RC(engine_gemm(e, g));
  } else {
    patch_embed_fwd(images, e.patch_w.d<float>(), e.patch_b.d<float>(), e.pos_emb.d<float>(), e.vit_x0, nimg,
                    c.image_size, c.vit_patch, W, st);
  }
  add_tap(e, "vit_embed", e.vit_x0, static_cast<int64_t>(Mv) * W, PI05_BF16);
  for (int l = 0; l < c.vit_depth; ++l) {
    const VitLayerP& p = e.vit[l];
    VitLayerA& a = e.va[l];
    if (l == e.profile_layer) cudaProfilerStart();
    layernorm_fwd(a.x_in, p.ln1_w.d<bf16>(), p.ln1_b.d<bf16>(), a.h1, a.mean1, a.rstd1, Mv, W, ln_eps, st);
    {  // fused q|k|v projection (+bias)
      GemmArgs g = mk_gemm(Mv, 3 * W, W, a.h1, W, p.q_w.data, W, a.qkv, 3 * W, EPI_BIAS);
      g.bias = p.q_b.data;
      CHECK_RC(engine_gemm(e, g));
    }
    {  // scores[img, head] = bf(bf(Q K^T) * hd^-1/2)   (modeling_siglip.py:333)
      GemmArgs g = mk_gemm(T, T, vhd, a.qkv, 3 * W, a.qkv + W, 3 * W, a.P, T, EPI_SCALE);
      g.batch = nimg * VH;
      g.batch_inner = VH;
      g.a_batch_stride = vhd;
      g.a_batch_stride1 = static_cast<int64_t>(T) * 3 * W;
      g.b_batch_stride = vhd;
      g.b_batch_stride1 = static_cast<int64_t>(T) * 3 * W;
      g.d_batch_stride = static_cast<int64_t>(T) * T;
      g.d_batch_stride1 = static_cast<int64_t>(VH) * T * T;
      g.scale = 1.0f / sqrtf(static_cast<float>(vhd));
      g.block_n = (T > 128) ? 256 : 128;
      CHECK_RC(engine_gemm(e, g));
    }
    softmax_fwd(a.P, T, T, nimg * VH, T, 0, nullptr, nullptr, 1, st);
    {  // attn[img, t, head, :] = P V
      GemmArgs g = mk_gemm(T, vhd, T, a.P, T, a.qkv + 2 * W, 3 * W, a.attn, W, EPI_STORE);
      g.b_major = 1;
      g.batch = nimg * VH;
      g.batch_inner = VH;
      g.a_batch_stride = static_cast<int64_t>(T) * T;
      g.a_batch_stride1 = static_cast<int64_t>(VH) * T * T;
      g.b_batch_stride = vhd;
      g.b_batch_stride1 = static_cast<int64_t>(T) * 3 * W;
      g.d_batch_stride = vhd;
      g.d_batch_stride1 = static_cast<int64_t>(T) * W;
      g.block_n = 128;
      CHECK_RC(engine_gemm(e, g));
    }
    {  // out_proj + bias + residual
      GemmArgs g = mk_gemm(Mv, W, W, a.attn, W, p.out_w.data, W, a.x_mid, W, EPI_RES);
      g.bias = p.out_b.data;
      g.res = a.x_in;
      g.ldres = W;
      CHECK_RC(engine_gemm(e, g));
    }
    layernorm_fwd(a.x_mid, p.ln2_w.d<bf16>(), p.ln2_b.d<bf16>(), a.h2, a.mean2, a.rstd2, Mv, W, ln_eps, st);
    {  // fc1 + bias + gelu_tanh
      GemmArgs g = mk_gemm(Mv, c.vit_mlp_dim, W, a.h2, W, p.fc1_w.data, W, a.pre, c.vit_mlp_dim, EPI_BIAS_GELU);
      g.bias = p.fc1_b.data;
      g.D2 = a.act;
      g.ldd2 = c.vit_mlp_dim;
      CHECK_RC(engine_gemm(e, g));
    }
    {  // fc2 + bias + residual
      GemmArgs g = mk_gemm(Mv, W, c.vit_mlp_dim, a.act, c.vit_mlp_dim, p.fc2_w.data, c.vit_mlp_dim, a.x_out, W, EPI_RES);
      g.bias = p.fc2_b.data;
      g.res = a.x_mid;
      g.ldres = W;
      CHECK_RC(engine_gemm(e, g));
    }
    if (e.taps_enabled) {
      char nm[64];
      snprintf(nm, sizeof(nm), "vit_layer%d", l);
      add_tap(e, nm, a.x_out, static_cast<int64_t>(Mv) * W, PI05_BF16);
    }
    if (l == e.profile_layer) cudaProfilerStop();
  }
  const bf16* xl = c.vit_depth > 0 ? e.va[c.vit_depth - 1].x_out : e.vit_x0;
  layernorm_fwd(xl, e.post_ln_w.d<bf16>(), e.post_ln_b.d<bf16>(), e.vit_post, e.vit_post_mean, e.vit_post_rstd, Mv, W,
                ln_eps, st);
  {  // projector (+bias), scattered into prefix_embs[b, n*T + t, :]  (z0 = b, z1 = image index n)
    GemmArgs g = mk_gemm(T, e.D, W, e.vit_post, W, e.proj_w.data, W, prefix_embs, e.D, EPI_BIAS);
    g.bias = e.proj_b.data;
    g.batch = nimg;
    g.batch_inner = B;
    g.a_batch_stride = static_cast<int64_t>(T) * W;
    g.a_batch_stride1 = static_cast<int64_t>(B) * T * W;
    g.d_batch_stride = static_cast<int64_t>(e.P) * e.D;
    g.d_batch_stride1 = static_cast<int64_t>(T) * e.D;
    CHECK_RC(engine_gemm(e, g));
  }
  return 0;
}

// embed_prefix (pi0_pytorch.py:186-235) + masks / positions (:342-343)
int prefix_forward(Engine& e, const pi05_batch* b) {
  const int B = b->batch;
  prefix_meta(b->image_masks, b->token_mask, B, e.NI, e.T, e.L, e.pad, e.pos, e.nvalid, e.stream);
  // index work of pi0_pytorch.py:219-235,342-343 (pad mask, cumsum(pad) - 1, valid-prefix count): bit-exact taps
  add_tap(e, "prefix_pad", e.pad, static_cast<int64_t>(B) * e.P, PI05_U8);
  add_tap(e, "prefix_pos", e.pos, static_cast<int64_t>(B) * e.P, PI05_I32);
  add_tap(e, "prefix_nvalid", e.nvalid, B, PI05_I32);
  bf16* prefix_embs = e.a1[0].x_in;
  if (b->images == nullptr && b->patch_rows == nullptr) {
    snprintf(e.err, sizeof(e.err), "pi05_batch: neither images nor patch_rows given");
    set_error(e.err);
    return 8;
  }
  CHECK_RC(vision_forward(e, b->images, b->patch_rows, B, prefix_embs));
  embed_tokens_fwd(b->tokens, e.embed.d<bf16>(), prefix_embs, B, e.L, e.D, static_cast<int64_t>(e.P) * e.D, e.NI * e.T,
                   static_cast<float>(sqrt(static_cast<double>(e.D))), e.stream);
  add_tap(e, "prefix_embs", prefix_embs, static_cast<int64_t>(B) * e.P * e.D, PI05_BF16);
  return 0;
}

// embed_suffix, pi05 branch (pi0_pytorch.py:237-314) + all adaRMS modulations (modeling_gemma.py:88)
int suffix_frontend(Engine& e, const float* x_t, const float* time, int B) {
  cudaStream_t st = e.stream;
  const int E = e.E, A = e.A, M2 = B * A, depth = e.cfg.paligemma.depth;
  time_embedding(time, e.time_scaling, e.temb, B, E / 2, st);
  linear_f32(x_t, e.ain_w.d<float>(), e.ain_b.d<float>(), e.aemb32, M2, E, e.cfg.action_dim, st);
  cast_f32_to_bf16(e.aemb32, e.a2[0].x_in, static_cast<int64_t>(M2) * E, st);
  linear_f32(e.temb, e.tin_w.d<float>(), e.tin_b.d<float>(), e.t1, B, E, E, st);
  silu_fwd(e.t1, e.t1s, static_cast<int64_t>(B) * E, st);
  linear_f32(e.t1s, e.tout_w.d<float>(), e.tout_b.d<float>(), e.t2, B, E, E, st);
  silu_fwd(e.t2, e.cond, static_cast<int64_t>(B) * E, st);
  const int64_t ms = static_cast<int64_t>(B) * 3 * E;
  if (e.ada_uniform && depth > 0) {
    // all 2*depth+1 modulation layers in ONE launch (weights uniformly strided in the fp32 arena)
    linear_f32_batched(e.cond, e.ex[0].in_dw.d<float>(), e.ex[0].in_db.d<float>(), e.mods, B, 3 * E, E, 2 * depth + 1,
                       e.ada_wstride, e.ada_bstride, ms, st);
    add_tap(e, "suffix_embs", e.a2[0].x_in, static_cast<int64_t>(M2) * E, PI05_BF16);
    add_tap(e, "adarms_cond", e.cond, static_cast<int64_t>(B) * E, PI05_F32);
    return 0;
  }
  for (int l = 0; l < depth; ++l) {
    linear_f32(e.cond, e.ex[l].in_dw.d<float>(), e.ex[l].in_db.d<float>(), e.mods + (2 * l) * ms, B, 3 * E, E, st);
    linear_f32(e.cond, e.ex[l].post_dw.d<float>(), e.ex[l].post_db.d<float>(), e.mods + (2 * l + 1) * ms, B, 3 * E, E,
               st);
  }
  linear_f32(e.cond, e.ex_norm_dw.d<float>(), e.ex_norm_db.d<float>(), e.mods + (2 * depth) * ms, B, 3 * E, E, st);
  add_tap(e, "suffix_embs", e.a2[0].x_in, static_cast<int64_t>(M2) * E, PI05_BF16);
  add_tap(e, "adarms_cond", e.cond, static_cast<int64_t>(B) * E, PI05_F32);
  return 0;
}

// One joint layer (gemma_pytorch.py:158-238).
int joint_layer_forward(Engine& e, int l, int B) {
  cudaStream_t st = e.stream;
  const pi05_config& c = e.cfg;
  const int P = e.P, A = e.A, S = e.S, D = e.D, E = e.E, H = e.H, hd = e.hd;
  const int M1 = B * P, M2 = B * A, QW = (H + 2) * hd;
  const float eps = 1e-6f;
  GemmaLayerA &p1 = e.a1[l], &p2 = e.a2[l];
  const GemmaLayerP &w1 = e.pg[l], &w2 = e.ex[l];
  const int64_t ms = static_cast<int64_t>(B) * 3 * E;
  const float* mod_in = e.mods + (2 * l) * ms;
  const float* mod_post = e.mods + (2 * l + 1) * ms;
  bf16 *Kc = e.Kl[l], *Vc = e.Vl[l];

  // input norms
  rmsnorm_fwd(p1.x_in, w1.in_w.d<float>(), nullptr, 0, p1.n1, p1.rstd1, nullptr, M1, D, eps, st);
  rmsnorm_fwd(p2.x_in, nullptr, mod_in, A, p2.n1, p2.rstd1, p2.gate1, M2, E, eps, st);
  // fused q|k|v projections
  CHECK_RC(engine_gemm(e, mk_gemm(M1, QW, D, p1.n1, D, w1.q_w.data, D, p1.qkv, QW, EPI_STORE)));
  CHECK_RC(engine_gemm(e, mk_gemm(M2, QW, E, p2.n1, E, w2.q_w.data, E, p2.qkv, QW, EPI_STORE)));
  // RoPE + concat along the sequence (gemma_pytorch.py:181-195)
  rope_pack_fwd(p1.qkv, P, H, hd, e.pos, e.nvalid, 0, e.rope_cos, e.rope_sin, p1.Q, Kc, Vc, 0, S, B, st);
  rope_pack_fwd(p2.qkv, A, H, hd, e.pos, e.nvalid, 1, e.rope_cos, e.rope_sin, p2.Q, Kc, Vc, P, S, B, st);
  const float scaling = 1.0f / sqrtf(static_cast<float>(hd));
  // The last layer's prefix-stream output only feeds prefix_out, which nothing on the loss path reads
  // (pi0_pytorch.py:350-358 keeps suffix_out only): its prefix-query attention, o_proj and MLP are dead code here.
  // They are computed only when taps are recorded (parity tests compare prefix_out too).
  const bool prefix_live = !(l == c.paligemma.depth - 1) || e.taps_enabled;
  // scores: prefix queries see prefix keys; suffix queries see everything (pi0_pytorch.py:52-81)
  if (prefix_live) {
    GemmArgs g = mk_gemm(P * H, P, hd, p1.Q, hd, Kc, hd, p1.P, e.Ppad, EPI_SCALE);
    g.batch = B;
    g.a_batch_stride = static_cast<int64_t>(P) * H * hd;
    g.b_batch_stride = static_cast<int64_t>(S) * hd;
    g.d_batch_stride = static_cast<int64_t>(P) * H * e.Ppad;
    g.scale = scaling;
    CHECK_RC(engine_gemm(e, g));
  }
  {
    GemmArgs g = mk_gemm(A * H, S, hd, p2.Q, hd, Kc, hd, p2.P, e.Spad, EPI_SCALE);
    g.batch = B;
    g.a_batch_stride = static_cast<int64_t>(A) * H * hd;
    g.b_batch_stride = static_cast<int64_t>(S) * hd;
    g.d_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    g.scale = scaling;
    CHECK_RC(engine_gemm(e, g));
  }
  if (prefix_live) softmax_fwd(p1.P, e.Ppad, P * H, B, P, P, e.pad, e.pad, H, st);
  softmax_fwd(p2.P, e.Spad, A * H, B, S, P, e.pad, nullptr, H, st);
  if (prefix_live) {  // O = P V  (V stored [keys, hd] -> N-major B operand)
    GemmArgs g = mk_gemm(P * H, hd, P, p1.P, e.Ppad, Vc, hd, p1.O, hd, EPI_STORE);
    g.b_major = 1;
    g.batch = B;
    g.a_batch_stride = static_cast<int64_t>(P) * H * e.Ppad;
    g.b_batch_stride = static_cast<int64_t>(S) * hd;
    g.d_batch_stride = static_cast<int64_t>(P) * H * hd;
    CHECK_RC(engine_gemm(e, g));
  }
  {
    GemmArgs g = mk_gemm(A * H, hd, S, p2.P, e.Spad, Vc, hd, p2.O, hd, EPI_STORE);
    g.b_major = 1;
    g.batch = B;
    g.a_batch_stride = static_cast<int64_t>(A) * H * e.Spad;
    g.b_batch_stride = static_cast<int64_t>(S) * hd;
    g.d_batch_stride = static_cast<int64_t>(A) * H * hd;
    CHECK_RC(engine_gemm(e, g));
  }
  {  // o_proj + (gated) residual
    if (prefix_live) {
      GemmArgs g = mk_gemm(M1, D, H * hd, p1.O, H * hd, w1.o_w.data, H * hd, p1.x_mid, D, EPI_RES);
      g.res = p1.x_in;
      g.ldres = D;
      CHECK_RC(engine_gemm(e, g));
    }
    GemmArgs g2 = mk_gemm(M2, E, H * hd, p2.O, H * hd, w2.o_w.data, H * hd, p2.x_mid, E, EPI_RES);
    g2.res = p2.x_in;
    g2.ldres = E;
    g2.gate = p2.gate1;
    g2.gate_rows = A;
    g2.ldgate = E;
    g2.D2 = p2.o_lin;
    g2.ldd2 = E;
    CHECK_RC(engine_gemm(e, g2));
  }
  // post-attention norms
  if (prefix_live) rmsnorm_fwd(p1.x_mid, w1.post_w.d<float>(), nullptr, 0, p1.n2, p1.rstd2, nullptr, M1, D, eps, st);
  rmsnorm_fwd(p2.x_mid, nullptr, mod_post, A, p2.n2, p2.rstd2, p2.gate2, M2, E, eps, st);
  {  // GeGLU up-projection (fused gate|up weight) then down-projection + (gated) residual
    if (prefix_live) {
      GemmArgs g = mk_gemm(M1, c.paligemma.mlp_dim, D, p1.n2, D, w1.gate_w.data, D, p1.GU, 2 * c.paligemma.mlp_dim,
                           EPI_GEGLU);
      g.D2 = p1.Hh;
      g.ldd2 = c.paligemma.mlp_dim;
      CHECK_RC(engine_gemm(e, g));
    }
    GemmArgs g2 = mk_gemm(M2, c.expert.mlp_dim, E, p2.n2, E, w2.gate_w.data, E, p2.GU, 2 * c.expert.mlp_dim, EPI_GEGLU);
    g2.D2 = p2.Hh;
    g2.ldd2 = c.expert.mlp_dim;
    CHECK_RC(engine_gemm(e, g2));
  }
  {
    if (prefix_live) {
      GemmArgs g = mk_gemm(M1, D, c.paligemma.mlp_dim, p1.Hh, c.paligemma.mlp_dim, w1.down_w.data, c.paligemma.mlp_dim,
                           p1.x_out, D, EPI_RES);
      g.res = p1.x_mid;
      g.ldres = D;
      CHECK_RC(engine_gemm(e, g));
    }
    GemmArgs g2 =
        mk_gemm(M2, E, c.expert.mlp_dim, p2.Hh, c.expert.mlp_dim, w2.down_w.data, c.expert.mlp_dim, p2.x_out, E, EPI_RES);
    g2.res = p2.x_mid;
    g2.ldres = E;
    g2.gate = p2.gate2;
    g2.gate_rows = A;
    g2.ldgate = E;
    g2.D2 = p2.d_lin;
    g2.ldd2 = E;
    CHECK_RC(engine_gemm(e, g2));
  }
  if (e.taps_enabled) {
    char nm[64];
    snprintf(nm, sizeof(nm), "layer%d_prefix", l);
    add_tap(e, nm, p1.x_out, static_cast<int64_t>(M1) * D, PI05_BF16);
    snprintf(nm, sizeof(nm), "layer%d_suffix", l);
    add_tap(e, nm, p2.x_out, static_cast<int64_t>(M2) * E, PI05_BF16);
  }
  return 0;
}

// Everything of the training forward up to suffix_out / v_t.  x_t_direct != null: x_t is given (sample_values).
static int forward_network(Engine& e, const pi05_batch* b, const float* actions, const float* noise, const float* time,
                           const float* x_t_direct, const char* who, cudaStream_t st) {
  if (!e.bound) {
    snprintf(e.err, sizeof(e.err), "%s: parameters not bound", who);
    set_error(e.err);
    return 8;
  }
  if (b->batch <= 0 || b->batch > e.Bmax) {
    snprintf(e.err, sizeof(e.err), "%s: batch %d outside [1, %d]", who, b->batch, e.Bmax);
    set_error(e.err);
    return 8;
  }
  e.stream = st;
  e.taps.clear();
  CHECK_RC(engine_set_token_len(e, b->token_len, who));
  const int B = b->batch;
  e.B = B;
  e.batch_copy = *b;
  const pi05_config& c = e.cfg;
  const int depth = c.paligemma.depth;
  const int M2 = B * e.A;
  if (x_t_direct != nullptr)
    cudaMemcpyAsync(e.x_t, x_t_direct, static_cast<size_t>(M2) * c.action_dim * sizeof(float), cudaMemcpyDeviceToDevice, st);
  else
    flow_inputs(actions, noise, time, e.x_t, e.u_t, B, e.A * c.action_dim, st);
  CHECK_RC(prefix_forward(e, b));
  CHECK_RC(suffix_frontend(e, e.x_t, time, B));
  for (int l = 0; l < depth; ++l) {
    if (l == e.profile_layer) cudaProfilerStart();
    CHECK_RC(joint_layer_forward(e, l, B));
    if (l == e.profile_layer) cudaProfilerStop();
  }
  const int64_t ms = static_cast<int64_t>(B) * 3 * e.E;
  const bf16* x1f = depth > 0 ? e.a1[depth - 1].x_out : e.a1[0].x_in;
  const bf16* x2f = depth > 0 ? e.a2[depth - 1].x_out : e.a2[0].x_in;
  if (e.taps_enabled) {
    rmsnorm_fwd(x1f, e.pg_norm_w.d<float>(), nullptr, 0, e.prefix_out, e.rstd_f1, nullptr, B * e.P, e.D, 1e-6f, st);
    add_tap(e, "prefix_out", e.prefix_out, static_cast<int64_t>(B) * e.P * e.D, PI05_BF16);
  }
  rmsnorm_fwd(x2f, nullptr, e.mods + (2 * depth) * ms, e.A, e.suffix_out, e.rstd_f2, nullptr, M2, e.E, 1e-6f, st);
  add_tap(e, "suffix_out", e.suffix_out, static_cast<int64_t>(M2) * e.E, PI05_BF16);
  return 0;
}

static void action_head(Engine& e, cudaStream_t st) {
  const pi05_config& c = e.cfg;
  const int M2 = e.B * e.A;
  cast_bf16_to_f32(e.suffix_out, e.so32, static_cast<int64_t>(M2) * e.E, st);
  linear_f32(e.so32, e.aout_w.d<float>(), e.aout_b.d<float>(), e.v_t, M2, c.action_dim, e.E, st);
  add_tap(e, "v_t", e.v_t, static_cast<int64_t>(M2) * c.action_dim, PI05_F32);
}

// value = tanh(MLP3(float(suffix_out[:, 0])))   (pi0_pytorch.py:473-481,571-572)
static void value_head_forward(Engine& e, cudaStream_t st) {
  const int B = e.B, E = e.E;
  gather_rows_f32(e.suffix_out, static_cast<int64_t>(e.A) * E, 0, 1, E, e.vh_in, B, st);
  linear_f32(e.vh_in, e.vh0_w.d<float>(), e.vh0_b.d<float>(), e.vh_h1, B, E, E, st);
  silu_fwd(e.vh_h1, e.vh_s1, static_cast<int64_t>(B) * E, st);
  linear_f32(e.vh_s1, e.vh2_w.d<float>(), e.vh2_b.d<float>(), e.vh_h2, B, E, E, st);
  silu_fwd(e.vh_h2, e.vh_s2, static_cast<int64_t>(B) * E, st);
  linear_f32(e.vh_s2, e.vh4_w.d<float>(), e.vh4_b.d<float>(), e.vh_h3, B, 1, E, st);
  tanh_fwd(e.vh_h3, e.vh_val, B, st);
  add_tap(e, "value", e.vh_val, B, PI05_F32);
}

static int finish(Engine& e, const char* who) {
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(e.err, sizeof(e.err), "%s: %s", who, cudaGetErrorString(ce));
    set_error(e.err);
    return 9;
  }
  return 0;
}

int engine_forward(Engine& e, const pi05_batch* b, const float* actions, const float* noise, const float* time,
                   float* loss_out, cudaStream_t st) {
  CHECK_RC(forward_network(e, b, actions, noise, time, nullptr, "pi05_forward", st));
  e.adv_mode = false;
  action_head(e, st);
  flow_loss(e.u_t, e.v_t, loss_out, static_cast<int64_t>(e.B) * e.A * e.cfg.action_dim, st);
  return finish(e, "pi05_forward");
}

int engine_forward_advantage(Engine& e, const pi05_batch* b, const float* actions, const float* noise, const float* time,
                             const float* progress, float w_action, float w_value, float* loss_out, float* aux_out,
                             cudaStream_t st) {
  if (!e.cfg.value_head) {
    snprintf(e.err, sizeof(e.err), "pi05_forward_advantage: engine was created without cfg.value_head");
    set_error(e.err);
    return 8;
  }
  CHECK_RC(forward_network(e, b, actions, noise, time, nullptr, "pi05_forward_advantage", st));
  e.adv_mode = true;
  e.w_action = w_action;
  e.w_value = w_value;
  action_head(e, st);
  value_head_forward(e, st);
  cudaMemcpyAsync(e.vh_prog, progress, static_cast<size_t>(e.B) * sizeof(float), cudaMemcpyDeviceToDevice, st);
  advantage_loss(e.u_t, e.v_t, e.vh_val, e.vh_prog, w_action, w_value, loss_out, e.vh_la, e.vh_lv, aux_out, e.B, e.A,
                 e.cfg.action_dim, st);
  return finish(e, "pi05_forward_advantage");
}

int engine_value(Engine& e, const pi05_batch* b, const float* noise, const float* time, float* value_out,
                 cudaStream_t st) {
  if (!e.cfg.value_head) {
    snprintf(e.err, sizeof(e.err), "pi05_forward_value: engine was created without cfg.value_head");
    set_error(e.err);
    return 8;
  }
  CHECK_RC(forward_network(e, b, nullptr, nullptr, time, noise, "pi05_forward_value", st));
  e.adv_mode = false;
  value_head_forward(e, st);
  cudaMemcpyAsync(value_out, e.vh_val, static_cast<size_t>(e.B) * sizeof(float), cudaMemcpyDeviceToDevice, st);
  return finish(e, "pi05_forward_value");
}

}  // namespace pi05
