// Vectorised single-pass masked softmax (forward / backward): one warp per score row, the whole row held in
// registers (CH x 8 bf16 per lane, 16-byte loads), one HBM read and one write per element.
// Same arithmetic as attn_kernels.cu::softmax_{fwd,bwd}_k (modeling_gemma.py:243-248); those remain the fallback for
// rows longer than 2048 keys or pitches that are not a multiple of 8.
#include <math_constants.h>

#include "common.cuh"
#include "errors.h"
#include "kernels.h"
#include "launch.h"

namespace pi05 {

namespace {

constexpr float kMaskValue = -2.3819763e38f;  // pi0_pytorch.py:159

template <int CH>
__global__ void __launch_bounds__(256) softmax_fwd_vec_k(bf16* __restrict__ s, int64_t ld, int rows_per_batch, int batch,
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
  // The kernel is instruction-bound, not bandwidth-bound, when every element pays for the mask logic (measured 2.4 TB/s):
  // masks are resolved per 8-key chunk -- all valid (the common case: no per-element work), all masked, or mixed.
  float v[CH][8];
  float mx = -CUDART_INF_F;
  constexpr uint64_t kAllValid = 0x0101010101010101ull;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int j0 = (i * 32 + lane) * 8;
    if (j0 + 8 <= n_keys) {
      load8(sr + j0, v[i]);
      // pad flags of this chunk: one 8-byte load inside the (8-aligned) prefix; suffix keys are always valid
      uint64_t flags = kAllValid;
      if (padb != nullptr && j0 < n_prefix) {
        if (j0 + 8 <= n_prefix && ((reinterpret_cast<uintptr_t>(padb) + j0) & 7) == 0) {
          flags = *reinterpret_cast<const uint64_t*>(padb + j0);
        } else {
          flags = 0;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            flags |= static_cast<uint64_t>((j0 + e >= n_prefix || padb[j0 + e] != 0) ? 1 : 0) << (8 * e);
        }
      }
      if (qmasked) flags = 0;
      if (flags != kAllValid) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (((flags >> (8 * e)) & 0xFFull) == 0) v[i][e] += kMaskValue;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) mx = fmaxf(mx, v[i][e]);
    } else if (j0 < n_keys) {  // ragged last chunk of the row: never read the (unwritten) pitch padding
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int j = j0 + e;
        if (j < n_keys) {
          float x = __bfloat162float(sr[j]);
          const bool ok = !qmasked && (j >= n_prefix || padb == nullptr || padb[j] != 0);
          if (!ok) x += kMaskValue;
          v[i][e] = x;
          mx = fmaxf(mx, x);
        } else {
          v[i][e] = -CUDART_INF_F;  // excluded from max / sum
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = -CUDART_INF_F;
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
  // exp(x - mx) = 2^((x - mx) * log2e) with one ex2.approx per element (rel. error ~1e-6, far below the bf16 rounding of
  // P); exp(-inf) = 0.  The subtraction comes FIRST: a fully masked row (padded query) has mx = -2.38e38, and
  // mx * log2e would overflow to -inf and turn the row into NaNs that the next layer's keys spread to valid rows.
  constexpr float kLog2e = 1.4426950408889634f;
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float ex;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"((v[i][e] - mx) * kLog2e));
      v[i][e] = ex;
      sum += ex;
    }
  sum = warp_sum(sum);
  // one IEEE reciprocal per row instead of 1024 divisions (<= 1 ulp from x/sum in fp32, i.e. far below the bf16
  // rounding applied next); the per-element fp32 division was the dominant cost of this kernel
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int j0 = (i * 32 + lane) * 8;
    if (j0 < ld) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = v[i][e] * inv;
      store8(sr + j0, v[i]);
    }
  }
}

template <int CH>
__global__ void __launch_bounds__(256) softmax_bwd_vec_k(const bf16* __restrict__ p, bf16* __restrict__ dp, int64_t ld,
                                                         int64_t rows, int n_keys, float scale) {
  pdl_enter();
  const int64_t row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16* pr = p + row * ld;
  bf16* dr = dp + row * ld;
  float pv[CH][8], dv[CH][8];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int j0 = (i * 32 + lane) * 8;
    if (j0 < ld) {
      if (j0 + 8 <= n_keys) {
        load8(pr + j0, pv[i]);
        load8(dr + j0, dv[i]);
      } else {  // last chunk: the dP pitch padding was never written
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool ok = j0 + e < n_keys;
          pv[i][e] = ok ? __bfloat162float(pr[j0 + e]) : 0.f;
          dv[i][e] = ok ? __bfloat162float(dr[j0 + e]) : 0.f;
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) dot += pv[i][e] * dv[i][e];
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        pv[i][e] = 0.f;
        dv[i][e] = 0.f;
      }
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int j0 = (i * 32 + lane) * 8;
    if (j0 < ld) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = bfr(pv[i][e] * (dv[i][e] - dot)) * scale;
      store8(dr + j0, o);
    }
  }
}

}  // namespace

bool softmax_fwd_vec(bf16* s, int64_t ld, int rows_per_batch, int batch, int n_keys, int n_prefix, const uint8_t* pad,
                     const uint8_t* qpad, int q_per_token, cudaStream_t st) {
  if (ld % 8 != 0 || ld > 2048 || (reinterpret_cast<uintptr_t>(s) & 15)) return false;
  const int64_t rows = static_cast<int64_t>(batch) * rows_per_batch;
  const int grid = ceil_div(rows, 8);
  const int ch = static_cast<int>((ld + 255) / 256);
#define LAUNCH_F(C)                                                                                                   \
  launch_pdl(softmax_fwd_vec_k<C>, dim3(grid), dim3(256), 0, st, s, ld, rows_per_batch, batch, n_keys, n_prefix, pad, qpad, q_per_token); \
  count_launch();                                                                                                     \
  break;
  switch (ch) {
    case 1: LAUNCH_F(1)
    case 2: LAUNCH_F(2)
    case 3: LAUNCH_F(3)
    case 4: LAUNCH_F(4)
    case 5: LAUNCH_F(5)
    case 6: LAUNCH_F(6)
    case 7: LAUNCH_F(7)
    default: LAUNCH_F(8)
  }
#undef LAUNCH_F
  return true;
}

bool softmax_bwd_vec(const bf16* p, bf16* dp, int64_t ld, int64_t rows, int n_keys, float scale, cudaStream_t st) {
  if (ld % 8 != 0 || ld > 2048 || (reinterpret_cast<uintptr_t>(p) & 15) || (reinterpret_cast<uintptr_t>(dp) & 15))
    return false;
  const int grid = ceil_div(rows, 8);
  const int ch = static_cast<int>((ld + 255) / 256);
#define LAUNCH_B(C)                                                               \
  launch_pdl(softmax_bwd_vec_k<C>, dim3(grid), dim3(256), 0, st, p, dp, ld, rows, n_keys, scale);    \
  count_launch();                                                                 \
  break;
  switch (ch) {
    case 1: LAUNCH_B(1)
    case 2: LAUNCH_B(2)
    case 3: LAUNCH_B(3)
    case 4: LAUNCH_B(4)
    case 5: LAUNCH_B(5)
    case 6: LAUNCH_B(6)
    case 7: LAUNCH_B(7)
    default: LAUNCH_B(8)
  }
#undef LAUNCH_B
  return true;
}

}  // namespace pi05
