// LayerNorm / (adaptive) RMSNorm forward + backward, gated-residual backward.
// One warp per row, 128-bit bf16 loads, warp-shuffle reductions, fp32 statistics.
// Reference arithmetic: modeling_siglip.py:466,474,787 (nn.LayerNorm), modeling_gemma.py:49-104 (GemmaRMSNorm),
// modeling_gemma.py:209-227 (_gated_residual).
#include "common.cuh"
#include "errors.h"
#include "kernels.h"
#include "launch.h"

namespace pi05 {

namespace {
constexpr int WARPS = 8;

// ---------------------------------------------------------------- register-resident forward (width <= 2048)
// The row is loaded ONCE, with all of its 16-byte loads in flight, and stays in registers for the statistics and the
// normalisation (the loop kernels below re-read it 2-3 times with one dependent load per iteration: 12-15 us for the
// 968-row prefix norms of the decode path, where there is no occupancy to hide it).  Same summation order per lane
// (chunk 0..CH-1, element 0..7) as the loop kernels: results are bit-identical.
__device__ __forceinline__ void unpack8v(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}

template <int CH>
__global__ void __launch_bounds__(WARPS * 32) layernorm_fwd_vec_k(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                  const bf16* __restrict__ b, bf16* __restrict__ y,
                                                                  float* __restrict__ mean_o, float* __restrict__ rstd_o,
                                                                  int rows, int width, float eps) {
  pdl_enter();
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16* xr = x + static_cast<int64_t>(row) * width;
  uint4 px[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = lane * 8 + k * 256;
    if (c < width) px[k] = *reinterpret_cast<const uint4*>(xr + c);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    if (lane * 8 + k * 256 < width) {
      float v[8];
      unpack8v(px[k], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[i];
    }
  }
  const float mean = warp_sum(s) / width;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    if (lane * 8 + k * 256 < width) {
      float v[8];
      unpack8v(px[k], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[i] - mean;
        ss += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / width + eps);
  bf16* yr = y + static_cast<int64_t>(row) * width;
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = lane * 8 + k * 256;
    if (c < width) {
      float v[8], ww[8], bb[8], o[8];
      unpack8v(px[k], v);
      load8(w + c, ww);
      load8(b + c, bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (v[i] - mean) * rstd * ww[i] + bb[i];
      store8(yr + c, o);
    }
  }
  if (lane == 0 && mean_o) {
    mean_o[row] = mean;
    rstd_o[row] = rstd;
  }
}

template <int CH>
__global__ void __launch_bounds__(WARPS * 32) rmsnorm_fwd_vec_k(const bf16* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ mod, int rows_per_batch,
                                                                bf16* __restrict__ y, float* __restrict__ rstd_o,
                                                                bf16* __restrict__ gate_out, int rows, int width,
                                                                float eps) {
  pdl_enter();
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16* xr = x + static_cast<int64_t>(row) * width;
  uint4 px[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = lane * 8 + k * 256;
    if (c < width) px[k] = *reinterpret_cast<const uint4*>(xr + c);
  }
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    if (lane * 8 + k * 256 < width) {
      float v[8];
      unpack8v(px[k], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += v[i] * v[i];
    }
  }
  const float var = warp_sum(ss) / width;
  const float rstd = rsqrtf(var + eps);  // modeling_gemma.py:68-70
  bf16* yr = y + static_cast<int64_t>(row) * width;
  const int b = row / rows_per_batch;
  const float* m = mod ? mod + static_cast<int64_t>(b) * 3 * width : nullptr;
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = lane * 8 + k * 256;
    if (c < width) {
      float v[8], o[8];
      unpack8v(px[k], v);
      if (m == nullptr) {
        float ww[8];
        load8f(w + c, ww);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __fmul_rn(__fmul_rn(v[i], rstd), __fadd_rn(1.0f, ww[i]));  // :80
      } else {
        float sc[8], sh[8];
        load8f(m + c, sc);
        load8f(m + width + c, sh);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o[i] = __fadd_rn(__fmul_rn(__fmul_rn(v[i], rstd), __fadd_rn(1.0f, sc[i])), sh[i]);  // :102
      }
      store8(yr + c, o);
    }
  }
  if (lane == 0 && rstd_o) rstd_o[row] = rstd;
  if (m != nullptr && gate_out != nullptr && (row % rows_per_batch) == 0) {
    bf16* g = gate_out + static_cast<int64_t>(b) * width;
    for (int c = lane * 8; c < width; c += 256) {
      float gv[8];
      load8f(m + 2 * width + c, gv);
      store8(g + c, gv);
    }
  }
}

// ---------------------------------------------------------------- LayerNorm fwd
__global__ void __launch_bounds__(WARPS * 32) layernorm_fwd_k(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                              const bf16* __restrict__ b, bf16* __restrict__ y,
                                                              float* __restrict__ mean_o, float* __restrict__ rstd_o,
                                                              int rows, int width, float eps) {
  pdl_enter();
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16* xr = x + static_cast<int64_t>(row) * width;
  float s = 0.f;
  for (int c = lane * 8; c < width; c += 256) {
    float v[8];
    load8(xr + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
  }
  const float mean = warp_sum(s) / width;
  float ss = 0.f;
  for (int c = lane * 8; c < width; c += 256) {
    float v[8];
    load8(xr + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = v[i] - mean;
      ss += d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / width + eps);
  bf16* yr = y + static_cast<int64_t>(row) * width;
  for (int c = lane * 8; c < width; c += 256) {
    float v[8], ww[8], bb[8], o[8];
    load8(xr + c, v);
    load8(w + c, ww);
    load8(b + c, bb);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (v[i] - mean) * rstd * ww[i] + bb[i];
    store8(yr + c, o);
  }
  if (lane == 0 && mean_o) {
    mean_o[row] = mean;
    rstd_o[row] = rstd;
  }
}

// ---------------------------------------------------------------- LayerNorm bwd (dx)
__global__ void __launch_bounds__(WARPS * 32) layernorm_bwd_dx_k(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                                 const bf16* __restrict__ w,
                                                                 const float* __restrict__ mean_i,
                                                                 const float* __restrict__ rstd_i,
                                                                 const bf16* __restrict__ dres, bf16* __restrict__ dx,
                                                                 int rows, int width) {
  pdl_enter();
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int64_t off = static_cast<int64_t>(row) * width;
  const float mean = mean_i[row], rstd = rstd_i[row];
  float s1 = 0.f, s2 = 0.f;  // sum(g), sum(g*xhat) with g = dy*w
  for (int c = lane * 8; c < width; c += 256) {
    float d[8], v[8], ww[8];
    load8(dy + off + c, d);
    load8(x + off + c, v);
    load8(w + c, ww);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float g = d[i] * ww[i];
      s1 += g;
      s2 += g * (v[i] - mean) * rstd;
    }
  }
  s1 = warp_sum(s1) / width;
  s2 = warp_sum(s2) / width;
  for (int c = lane * 8; c < width; c += 256) {
    float d[8], v[8], ww[8], o[8];
    load8(dy + off + c, d);
    load8(x + off + c, v);
    load8(w + c, ww);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (v[i] - mean) * rstd;
      o[i] = bfr(rstd * (d[i] * ww[i] - s1 - xh * s2));
    }
    if (dres) {
      float r[8];
      load8(dres + off + c, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += r[i];
    }
    store8(dx + off + c, o);
  }
}

// dw += sum_r dy*xhat ; db += sum_r dy.  Block = 64-row slab x 256 threads, 8 columns per thread per pass.
constexpr int SLAB = 64;
template <bool LN>
__global__ void __launch_bounds__(256) norm_bwd_dwdb_k(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                       const float* __restrict__ mean_i,
                                                       const float* __restrict__ rstd_i, float* __restrict__ dw32,
                                                       float* __restrict__ db32, int rows, int width) {
  pdl_enter();
  const int r0 = blockIdx.x * SLAB;
  const int r1 = min(rows, r0 + SLAB);
  for (int c = threadIdx.x * 8; c < width; c += 256 * 8) {
    float aw[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = r0; r < r1; ++r) {
      const int64_t off = static_cast<int64_t>(r) * width + c;
      float d[8], v[8];
      load8(dy + off, d);
      load8(x + off, v);
      const float m = LN ? mean_i[r] : 0.f;
      const float rs = rstd_i[r];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        aw[i] += d[i] * (v[i] - m) * rs;
        ab[i] += d[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(dw32 + c + i, aw[i]);
      if (LN) atomicAdd(db32 + c + i, ab[i]);
    }
  }
}

// ---------------------------------------------------------------- RMSNorm fwd
__global__ void __launch_bounds__(WARPS * 32) rmsnorm_fwd_k(const bf16* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ mod, int rows_per_batch,
                                                            bf16* __restrict__ y, float* __restrict__ rstd_o,
                                                            bf16* __restrict__ gate_out, int rows, int width,
                                                            float eps) {
  pdl_enter();
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16* xr = x + static_cast<int64_t>(row) * width;
  float ss = 0.f;
  for (int c = lane * 8; c < width; c += 256) {
    float v[8];
    load8(xr + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += v[i] * v[i];
  }
  const float var = warp_sum(ss) / width;
  const float rstd = rsqrtf(var + eps);  // modeling_gemma.py:68-70
  bf16* yr = y + static_cast<int64_t>(row) * width;
  const int b = row / rows_per_batch;
  const float* m = mod ? mod + static_cast<int64_t>(b) * 3 * width : nullptr;
  for (int c = lane * 8; c < width; c += 256) {
    float v[8], o[8];
    load8(xr + c, v);
    if (m == nullptr) {
      float ww[8];
      load8f(w + c, ww);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = __fmul_rn(__fmul_rn(v[i], rstd), __fadd_rn(1.0f, ww[i]));  // :80
    } else {
      float sc[8], sh[8];
      load8f(m + c, sc);
      load8f(m + width + c, sh);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o[i] = __fadd_rn(__fmul_rn(__fmul_rn(v[i], rstd), __fadd_rn(1.0f, sc[i])), sh[i]);  // :102
    }
    store8(yr + c, o);
  }
  if (lane == 0 && rstd_o) rstd_o[row] = rstd;
  // gate (bf16) once per batch: the first row of each batch writes it (modeling_gemma.py:104)
  if (m != nullptr && gate_out != nullptr && (row % rows_per_batch) == 0) {
    bf16* g = gate_out + static_cast<int64_t>(b) * width;
    for (int c = lane * 8; c < width; c += 256) {
      float gv[8];
      load8f(m + 2 * width + c, gv);
      store8(g + c, gv);
    }
  }
}

// ---------------------------------------------------------------- RMSNorm bwd (dx)
__global__ void __launch_bounds__(WARPS * 32) rmsnorm_bwd_dx_k(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ mod, int rows_per_batch,
                                                               const float* __restrict__ rstd_i,
                                                               const bf16* __restrict__ dres, bf16* __restrict__ dx,
                                                               int rows, int width) {
  pdl_enter();
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int64_t off = static_cast<int64_t>(row) * width;
  const float rstd = rstd_i[row];
  const int b = row / rows_per_batch;
  const float* m = mod ? mod + static_cast<int64_t>(b) * 3 * width : nullptr;
  float s = 0.f;  // sum(g * xhat), g = dy*(1+w or 1+scale)
  for (int c = lane * 8; c < width; c += 256) {
    float d[8], v[8], ww[8];
    load8(dy + off + c, d);
    load8(x + off + c, v);
    load8f(m ? m + c : w + c, ww);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d[i] * (1.0f + ww[i]) * v[i] * rstd;
  }
  s = warp_sum(s) / width;
  for (int c = lane * 8; c < width; c += 256) {
    float d[8], v[8], ww[8], o[8];
    load8(dy + off + c, d);
    load8(x + off + c, v);
    load8f(m ? m + c : w + c, ww);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = bfr(rstd * (d[i] * (1.0f + ww[i]) - v[i] * rstd * s));
    if (dres) {
      float r[8];
      load8(dres + off + c, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += r[i];
    }
    store8(dx + off + c, o);
  }
}

// adaptive: dmod[b].scale += sum_rows dy*xhat ; dmod[b].shift += sum_rows dy.
// grid (width / 256, batch): lane = 8-column group, warp = row lane (rows t = warp, warp + 8, ...); the 8 row-lane
// partials are summed in a fixed order through shared memory (deterministic, no atomics).
__global__ void __launch_bounds__(256) adarms_bwd_dmod_k(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                         const float* __restrict__ rstd_i, int rows_per_batch,
                                                         float* __restrict__ dmod, int width) {
  pdl_enter();
  __shared__ float red[2][8][256];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  float a1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < width) {
    for (int t = warp; t < rows_per_batch; t += 8) {
      const int r = b * rows_per_batch + t;
      const int64_t off = static_cast<int64_t>(r) * width + c;
      float d[8], v[8];
      load8(dy + off, d);
      load8(x + off, v);
      const float rs = rstd_i[r];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a1[i] += d[i] * v[i] * rs;
        a2[i] += d[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[0][warp][lane * 8 + i] = a1[i];
    red[1][warp][lane * 8 + i] = a2[i];
  }
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < width) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      t1 += red[0][w][threadIdx.x];
      t2 += red[1][w][threadIdx.x];
    }
    float* dm = dmod + static_cast<int64_t>(b) * 3 * width;
    dm[cc] += t1;
    dm[width + cc] += t2;
  }
}

// y = x + o*gate:  d_o = bf(dy*gate) ; dmod[b].gate += sum_rows dy*o
__global__ void __launch_bounds__(256) gated_residual_bwd_k(const bf16* __restrict__ dy, const bf16* __restrict__ o,
                                                            const bf16* __restrict__ gate, int rows_per_batch,
                                                            bf16* __restrict__ d_o, float* __restrict__ dmod,
                                                            int width) {
  pdl_enter();
  __shared__ float red[8][256];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < width) {
    float g[8];
    load8(gate + static_cast<int64_t>(b) * width + c, g);
    for (int t = warp; t < rows_per_batch; t += 8) {
      const int64_t off = (static_cast<int64_t>(b) * rows_per_batch + t) * width + c;
      float d[8], ov[8], r[8];
      load8(dy + off, d);
      load8(o + off, ov);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        r[i] = d[i] * g[i];
        acc[i] += d[i] * ov[i];
      }
      store8(d_o + off, r);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < width) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    dmod[static_cast<int64_t>(b) * 3 * width + 2 * width + cc] += t;
  }
}

}  // namespace

void layernorm_fwd(const bf16* x, const bf16* w, const bf16* b, bf16* y, float* mean, float* rstd, int rows, int width,
                   float eps, cudaStream_t st) {
  const dim3 grid(ceil_div(rows, WARPS)), block(WARPS * 32);
  const int ch = (width + 255) / 256;
  if (width % 8 == 0 && ch <= 8) {
#define PI05_LN_CASE(C)                                                                                         \
  case C:                                                                                                       \
    launch_pdl(layernorm_fwd_vec_k<C>, grid, block, 0, st, x, w, b, y, mean, rstd, rows, width, eps);           \
    break;
    switch (ch) {
      PI05_LN_CASE(1) PI05_LN_CASE(2) PI05_LN_CASE(3) PI05_LN_CASE(4) PI05_LN_CASE(5) PI05_LN_CASE(6) PI05_LN_CASE(7)
      default: launch_pdl(layernorm_fwd_vec_k<8>, grid, block, 0, st, x, w, b, y, mean, rstd, rows, width, eps);
    }
#undef PI05_LN_CASE
  } else {
    launch_pdl(layernorm_fwd_k, grid, block, 0, st, x, w, b, y, mean, rstd, rows, width, eps);
  }
  count_launch();
}

void layernorm_bwd(const bf16* dy, const bf16* x, const bf16* w, const float* mean, const float* rstd,
                   const bf16* dres, bf16* dx, float* dw32, float* db32, int rows, int width, cudaStream_t st) {
  if (dw32 != nullptr && db32 != nullptr &&
      layernorm_bwd_fused(dy, x, w, mean, rstd, dres, dx, dw32, db32, rows, width, st))
    return;
  launch_pdl(layernorm_bwd_dx_k, dim3(ceil_div(rows, WARPS)), dim3(WARPS * 32), 0, st, dy, x, w, mean, rstd, dres, dx, rows, width); count_launch();
  launch_pdl(norm_bwd_dwdb_k<true>, dim3(ceil_div(rows, SLAB)), dim3(256), 0, st, dy, x, mean, rstd, dw32, db32, rows, width); count_launch();
}

void rmsnorm_fwd(const bf16* x, const float* w, const float* mod, int rows_per_batch, bf16* y, float* rstd,
                 bf16* gate_out, int rows, int width, float eps, cudaStream_t st) {
  const dim3 grid(ceil_div(rows, WARPS)), block(WARPS * 32);
  const int rpb = rows_per_batch > 0 ? rows_per_batch : rows;
  const int ch = (width + 255) / 256;
  if (width % 8 == 0 && ch <= 8) {
#define PI05_RMS_CASE(C)                                                                                              \
  case C:                                                                                                             \
    launch_pdl(rmsnorm_fwd_vec_k<C>, grid, block, 0, st, x, w, mod, rpb, y, rstd, gate_out, rows, width, eps);        \
    break;
    switch (ch) {
      PI05_RMS_CASE(1) PI05_RMS_CASE(2) PI05_RMS_CASE(3) PI05_RMS_CASE(4) PI05_RMS_CASE(5) PI05_RMS_CASE(6) PI05_RMS_CASE(7)
      default: launch_pdl(rmsnorm_fwd_vec_k<8>, grid, block, 0, st, x, w, mod, rpb, y, rstd, gate_out, rows, width, eps);
    }
#undef PI05_RMS_CASE
  } else {
    launch_pdl(rmsnorm_fwd_k, grid, block, 0, st, x, w, mod, rpb, y, rstd, gate_out, rows, width, eps);
  }
  count_launch();
}

void rmsnorm_bwd(const bf16* dy, const bf16* x, const float* w, const float* mod, int rows_per_batch,
                 const float* rstd, const bf16* dres, bf16* dx, float* dw32, float* dmod, int rows, int width,
                 cudaStream_t st) {
  const int rpb = rows_per_batch > 0 ? rows_per_batch : rows;
  if (mod == nullptr && dw32 != nullptr && rmsnorm_bwd_fused(dy, x, w, rstd, dres, dx, dw32, rows, width, st)) return;
  launch_pdl(rmsnorm_bwd_dx_k, dim3(ceil_div(rows, WARPS)), dim3(WARPS * 32), 0, st, dy, x, w, mod, rpb, rstd, dres, dx, rows, width); count_launch();
  if (mod == nullptr) {
    launch_pdl(norm_bwd_dwdb_k<false>, dim3(ceil_div(rows, SLAB)), dim3(256), 0, st, dy, x, nullptr, rstd, dw32, nullptr, rows, width); count_launch();
  } else {
    launch_pdl(adarms_bwd_dmod_k, dim3(ceil_div(width, 256), rows / rpb), dim3(256), 0, st, dy, x, rstd, rpb, dmod, width); count_launch();
  }
}

void gated_residual_bwd(const bf16* dy, const bf16* o, const bf16* gate, int rows_per_batch, bf16* d_o, float* dmod,
                        int rows, int width, cudaStream_t st) {
  launch_pdl(gated_residual_bwd_k, dim3(ceil_div(width, 256), rows / rows_per_batch), dim3(256), 0, st, dy, o, gate, rows_per_batch, d_o, dmod, width); count_launch();
}

}  // namespace pi05
