// Prefix bookkeeping, RoPE + QKV layout split, masked softmax (forward / backward).
// Reference arithmetic: pi0_pytorch.py:52-81,156-159,207,221,343 (masks, positions),
// modeling_gemma.py:147-194 (rotary), modeling_gemma.py:243-248 (scale, mask add, fp32 softmax -> bf16).
#include "common.cuh"
#include "errors.h"
#include "kernels.h"
#include "launch.h"

namespace pi05 {

// softmax_kernels.cu: vectorised single-pass variants (return false when the shape is not eligible)
bool softmax_fwd_vec(bf16* s, int64_t ld, int rows_per_batch, int batch, int n_keys, int n_prefix, const uint8_t* pad,
                     const uint8_t* qpad, int q_per_token, cudaStream_t st);
bool softmax_bwd_vec(const bf16* p, bf16* dp, int64_t ld, int64_t rows, int n_keys, float scale, cudaStream_t st);

namespace {

constexpr float kMaskValue = -2.3819763e38f;  // pi0_pytorch.py:159

__global__ void prefix_meta_k(const uint8_t* __restrict__ image_masks, const uint8_t* __restrict__ token_mask, int batch,
                              int num_images, int tpi, int L, uint8_t* __restrict__ pad, int* __restrict__ pos,
                              int* __restrict__ nvalid) {
  pdl_enter();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int P = num_images * tpi + L;
  int run = 0;
  for (int j = 0; j < P; ++j) {
    uint8_t v;
    if (j < num_images * tpi)
      v = image_masks[(j / tpi) * batch + b] ? 1 : 0;
    else
      v = token_mask[b * L + (j - num_images * tpi)] ? 1 : 0;
    run += v;
    pad[b * P + j] = v;
    pos[b * P + j] = run - 1;  // cumsum(pad) - 1 (can be -1 for leading padding)
  }
  nvalid[b] = run;
}

// One work item = 8 rotation pairs (or 16 V elements) of one (row, slot); slot < H: query head, H: key, H+1: value.
__global__ void __launch_bounds__(256) rope_pack_fwd_k(const bf16* __restrict__ qkv, int T, int H, int hd,
                                                       const int* __restrict__ pos, const int* __restrict__ nvalid,
                                                       int pos_mode, const bf16* __restrict__ cos_t,
                                                       const bf16* __restrict__ sin_t, bf16* __restrict__ Q,
                                                       bf16* __restrict__ K, bf16* __restrict__ V, int key_off,
                                                       int kv_len, int batch) {
  pdl_enter();
  const int half = hd / 2;
  const int chunks = half / 8;
  const int64_t total = static_cast<int64_t>(batch) * T * (H + 2) * chunks;
  const int ldq = (H + 2) * hd;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(idx % chunks);
    const int slot = static_cast<int>((idx / chunks) % (H + 2));
    const int64_t row = idx / (static_cast<int64_t>(chunks) * (H + 2));
    const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
    const bf16* src = qkv + row * ldq + slot * hd + ch * 8;
    float x1[8], x2[8];
    load8(src, x1);
    load8(src + half, x2);
    if (slot == H + 1) {
      bf16* dst = V + (static_cast<int64_t>(b) * kv_len + key_off + t) * hd + ch * 8;
      store8(dst, x1);
      store8(dst + half, x2);
      continue;
    }
    const int p = (pos_mode == 0 ? pos[row] : nvalid[b] + t) + 1;
    float c[8], s[8], o1[8], o2[8];
    load8(cos_t + static_cast<int64_t>(p) * half + ch * 8, c);
    load8(sin_t + static_cast<int64_t>(p) * half + ch * 8, s);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o1[i] = bfr(x1[i] * c[i]) + bfr(-x2[i] * s[i]);  // q*cos + rotate_half(q)*sin, bf16 ops
      o2[i] = bfr(x2[i] * c[i]) + bfr(x1[i] * s[i]);
    }
    bf16* dst = (slot < H) ? Q + (row * H + slot) * hd + ch * 8
                           : K + (static_cast<int64_t>(b) * kv_len + key_off + t) * hd + ch * 8;
    store8(dst, o1);
    store8(dst + half, o2);
  }
}

__global__ void __launch_bounds__(256) rope_pack_bwd_k(const bf16* __restrict__ dQ, const float* __restrict__ dK,
                                                       const float* __restrict__ dV, int T, int H, int hd,
                                                       const int* __restrict__ pos, const int* __restrict__ nvalid,
                                                       int pos_mode, const bf16* __restrict__ cos_t,
                                                       const bf16* __restrict__ sin_t, bf16* __restrict__ dqkv,
                                                       int key_off, int kv_len, int batch) {
  pdl_enter();
  const int half = hd / 2;
  const int chunks = half / 8;
  const int64_t total = static_cast<int64_t>(batch) * T * (H + 2) * chunks;
  const int ldq = (H + 2) * hd;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(idx % chunks);
    const int slot = static_cast<int>((idx / chunks) % (H + 2));
    const int64_t row = idx / (static_cast<int64_t>(chunks) * (H + 2));
    const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
    bf16* dst = dqkv + row * ldq + slot * hd + ch * 8;
    float d1[8], d2[8];
    if (slot < H) {
      const bf16* src = dQ + (row * H + slot) * hd + ch * 8;
      load8(src, d1);
      load8(src + half, d2);
    } else {
      const float* src = (slot == H ? dK : dV) + (static_cast<int64_t>(b) * kv_len + key_off + t) * hd + ch * 8;
      load8f(src, d1);
      load8f(src + half, d2);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        d1[i] = bfr(d1[i]);
        d2[i] = bfr(d2[i]);
      }
    }
    if (slot == H + 1) {
      store8(dst, d1);
      store8(dst + half, d2);
      continue;
    }
    const int p = (pos_mode == 0 ? pos[row] : nvalid[b] + t) + 1;
    float c[8], s[8], o1[8], o2[8];
    load8(cos_t + static_cast<int64_t>(p) * half + ch * 8, c);
    load8(sin_t + static_cast<int64_t>(p) * half + ch * 8, s);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o1[i] = bfr(d1[i] * c[i]) + bfr(d2[i] * s[i]);
      o2[i] = bfr(d2[i] * c[i]) + bfr(-d1[i] * s[i]);
    }
    store8(dst, o1);
    store8(dst + half, o2);
  }
}

// One warp per score row.
__global__ void __launch_bounds__(256) softmax_fwd_k(bf16* __restrict__ s, int64_t ld, int rows_per_batch, int batch,
                                                     int n_keys, int n_prefix, const uint8_t* __restrict__ pad,
                                                     const uint8_t* __restrict__ qpad, int q_per_token) {
  pdl_enter();
  const int64_t row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<int64_t>(batch) * rows_per_batch) return;
  const int b = static_cast<int>(row / rows_per_batch);
  bf16* sr = s + row * ld;
  const uint8_t* padb = pad ? pad + static_cast<int64_t>(b) * n_prefix : nullptr;
  bool qmasked = false;
  if (qpad) {
    const int tokens = rows_per_batch / q_per_token;
    const int tok = static_cast<int>(row % rows_per_batch) / q_per_token;
    qmasked = qpad[static_cast<int64_t>(b) * tokens + tok] == 0;
  }
  auto masked_val = [&](int j) -> float {
    const float v = __bfloat162float(sr[j]);
    const bool valid = !qmasked && (j >= n_prefix || padb == nullptr || padb[j] != 0);
    return valid ? v : v + kMaskValue;
  };
  float mx = -INFINITY;
  for (int j = lane; j < n_keys; j += 32) mx = fmaxf(mx, masked_val(j));
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n_keys; j += 32) sum += expf(masked_val(j) - mx);
  sum = warp_sum(sum);
  for (int j = lane; j < n_keys; j += 32) {
    const float pj = expf(masked_val(j) - mx) / sum;
    sr[j] = __float2bfloat16_rn(pj);
  }
  // zero the pitch padding so later K-loops over ld never see garbage NaNs
  for (int64_t j = n_keys + lane; j < ld; j += 32) sr[j] = __float2bfloat16_rn(0.f);
}

__global__ void __launch_bounds__(256) softmax_bwd_k(const bf16* __restrict__ p, bf16* __restrict__ dp, int64_t ld,
                                                     int64_t rows, int n_keys, float scale) {
  pdl_enter();
  const int64_t row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16* pr = p + row * ld;
  bf16* dr = dp + row * ld;
  float dot = 0.f;
  for (int j = lane; j < n_keys; j += 32) dot += __bfloat162float(pr[j]) * __bfloat162float(dr[j]);
  dot = warp_sum(dot);
  for (int j = lane; j < n_keys; j += 32) {
    const float pj = __bfloat162float(pr[j]);
    const float ds = pj * (__bfloat162float(dr[j]) - dot);
    dr[j] = __float2bfloat16_rn(bfr(ds) * scale);
  }
  for (int64_t j = n_keys + lane; j < ld; j += 32) dr[j] = __float2bfloat16_rn(0.f);
}

}  // namespace

void prefix_meta(const uint8_t* image_masks, const uint8_t* token_mask, int batch, int num_images, int tokens_per_image,
                 int max_token_len, uint8_t* pad, int* pos, int* nvalid, cudaStream_t st) {
  launch_pdl(prefix_meta_k, dim3(ceil_div(batch, 32)), dim3(32), 0, st, image_masks, token_mask, batch, num_images, tokens_per_image,
                                                    max_token_len, pad, pos, nvalid); count_launch();
}

void rope_pack_fwd(const bf16* qkv, int T, int H, int hd, const int* pos, const int* nvalid, int pos_mode,
                   const bf16* cos_t, const bf16* sin_t, bf16* Q, bf16* K, bf16* V, int key_off, int kv_len, int batch,
                   cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(batch) * T * (H + 2) * (hd / 16);
  const int blocks = static_cast<int>(total / 256 + 1 < 148 * 8 ? total / 256 + 1 : 148 * 8);
  launch_pdl(rope_pack_fwd_k, dim3(blocks), dim3(256), 0, st, qkv, T, H, hd, pos, nvalid, pos_mode, cos_t, sin_t, Q, K, V, key_off, kv_len,
                                          batch); count_launch();
}

void rope_pack_bwd(const bf16* dQ, const float* dK, const float* dV, int T, int H, int hd, const int* pos,
                   const int* nvalid, int pos_mode, const bf16* cos_t, const bf16* sin_t, bf16* dqkv, int key_off,
                   int kv_len, int batch, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(batch) * T * (H + 2) * (hd / 16);
  const int blocks = static_cast<int>(total / 256 + 1 < 148 * 8 ? total / 256 + 1 : 148 * 8);
  launch_pdl(rope_pack_bwd_k, dim3(blocks), dim3(256), 0, st, dQ, dK, dV, T, H, hd, pos, nvalid, pos_mode, cos_t, sin_t, dqkv, key_off,
                                          kv_len, batch); count_launch();
}

void softmax_fwd(bf16* s, int64_t ld, int rows_per_batch, int batch, int n_keys, int n_prefix, const uint8_t* pad,
                 const uint8_t* qpad, int q_per_token, cudaStream_t st) {
  if (softmax_fwd_vec(s, ld, rows_per_batch, batch, n_keys, n_prefix, pad, qpad, q_per_token > 0 ? q_per_token : 1, st))
    return;
  const int64_t rows = static_cast<int64_t>(batch) * rows_per_batch;
  launch_pdl(softmax_fwd_k, dim3(ceil_div(rows, 8)), dim3(256), 0, st, s, ld, rows_per_batch, batch, n_keys, n_prefix, pad, qpad,
                                                   q_per_token > 0 ? q_per_token : 1); count_launch();
}

void softmax_bwd(const bf16* p, bf16* dp, int64_t ld, int rows, int n_keys, float scale, cudaStream_t st) {
  if (softmax_bwd_vec(p, dp, ld, rows, n_keys, scale, st)) return;
  launch_pdl(softmax_bwd_k, dim3(ceil_div(rows, 8)), dim3(256), 0, st, p, dp, ld, rows, n_keys, scale); count_launch();
}

}  // namespace pi05
