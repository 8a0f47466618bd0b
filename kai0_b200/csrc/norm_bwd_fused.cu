// Fused LayerNorm / plain-RMSNorm backward: dx AND the weight (bias) gradient in ONE pass over dy and x.
// The two-kernel version (norm_kernels.cu: *_bwd_dx_k + norm_bwd_dwdb_k) read dy and x twice from HBM and its
// column-sum kernel ran at 9-19 % of the HBM peak (profiles/r01_ncu_layer.md).  Here each warp walks rows with a grid
// stride, keeps the row's dy / x packed in registers between the statistics pass and the dx pass, accumulates the
// per-column dw (db) partials of ITS rows in registers, and the block reduces them through shared memory into one
// atomicAdd per column per block.  Arithmetic per element is identical to the two-kernel version.
#include "common.cuh"
#include "errors.h"
#include "kernels.h"
#include "launch.h"

namespace pi05 {
namespace {

constexpr int NW = 8;  // warps per block

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}

// LN = true : y = (x - mean) * rstd * w + b   (w bf16; dw, db)
// LN = false: y = x * rstd * (1 + w)          (w fp32; dw)
template <bool LN, int CH>
__global__ void __launch_bounds__(NW * 32) norm_bwd_fused_k(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                            const void* __restrict__ w_, const float* __restrict__ mean_i,
                                                            const float* __restrict__ rstd_i,
                                                            const bf16* __restrict__ dres, bf16* __restrict__ dx,
                                                            float* __restrict__ dw32, float* __restrict__ db32, int rows,
                                                            int width) {
  pdl_enter();
  __shared__ float red[NW][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per-lane weights of its columns (row-invariant)
  float wv[CH][8];
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = lane * 8 + k * 256;
    if (c < width) {
      if (LN) {
        load8(static_cast<const bf16*>(w_) + c, wv[k]);
      } else {
        load8f(static_cast<const float*>(w_) + c, wv[k]);
#pragma unroll
        for (int i = 0; i < 8; ++i) wv[k][i] = 1.0f + wv[k][i];
      }
    }
  }
  float aw[CH][8], ab[LN ? CH : 1][8];
#pragma unroll
  for (int k = 0; k < CH; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      aw[k][i] = 0.f;
      if (LN) ab[k][i] = 0.f;
    }
  const float inv_w = 1.0f / static_cast<float>(width);
  for (int row = blockIdx.x * NW + warp; row < rows; row += gridDim.x * NW) {
    const int64_t off = static_cast<int64_t>(row) * width;
    uint4 pd[CH], pv[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int c = lane * 8 + k * 256;
      if (c < width) {
        pd[k] = *reinterpret_cast<const uint4*>(dy + off + c);
        pv[k] = *reinterpret_cast<const uint4*>(x + off + c);
      }
    }
    const float mean = LN ? mean_i[row] : 0.f;
    const float rstd = rstd_i[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int c = lane * 8 + k * 256;
      if (c < width) {
        float d[8], v[8];
        unpack8(pd[k], d);
        unpack8(pv[k], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float g = d[i] * wv[k][i];
          if (LN) {
            s1 += g;
            s2 += g * (v[i] - mean) * rstd;
          } else {
            s2 += g * v[i] * rstd;
          }
        }
      }
    }
    if (LN) s1 = warp_sum(s1) * inv_w;
    s2 = warp_sum(s2) * inv_w;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int c = lane * 8 + k * 256;
      if (c < width) {
        float d[8], v[8], o[8];
        unpack8(pd[k], d);
        unpack8(pv[k], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = LN ? (v[i] - mean) * rstd : v[i] * rstd;
          o[i] = LN ? bfr(rstd * (d[i] * wv[k][i] - s1 - xh * s2)) : bfr(rstd * (d[i] * wv[k][i] - xh * s2));
          aw[k][i] += d[i] * xh;
          if (LN) ab[k][i] += d[i];
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
  }
  // block reduction of the column partials, 256 columns at a time
#pragma unroll
  for (int pass = 0; pass < (LN ? 2 : 1); ++pass) {
    float* dst = pass == 0 ? dw32 : db32;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = (pass == 0) ? aw[k][i] : ab[LN ? k : 0][i];
      __syncthreads();
      const int c = k * 256 + threadIdx.x;
      if (c < width) {
        float t = 0.f;
#pragma unroll
        for (int wi = 0; wi < NW; ++wi) t += red[wi][threadIdx.x];
        atomicAdd(dst + c, t);
      }
      __syncthreads();
    }
  }
}

template <bool LN>
bool launch_fused(const bf16* dy, const bf16* x, const void* w, const float* mean, const float* rstd, const bf16* dres,
                  bf16* dx, float* dw32, float* db32, int rows, int width, cudaStream_t st) {
  if (width % 8 != 0 || width > 2048) return false;
  const int ch = (width + 255) / 256;
  const int want = (rows + NW - 1) / NW;
  int grid = want;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // exactly one resident wave (grid-stride rows): blocks per SM from the occupancy calculator of this instantiation
#define PI05_NORM_CASE(C)                                                                                            \
  case C: {                                                                                                          \
    static int per_sm = 0;                                                                                           \
    if (per_sm == 0) {                                                                                               \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, norm_bwd_fused_k<LN, C>, NW * 32, 0);                   \
      if (per_sm < 1) per_sm = 1;                                                                                    \
    }                                                                                                                \
    if (grid > per_sm * sms) grid = per_sm * sms;                                                                    \
  }                                                                                                                  \
    launch_pdl(norm_bwd_fused_k<LN, C>, dim3(grid), dim3(NW * 32), 0, st, dy, x, w, mean, rstd, dres, dx, dw32, db32, \
               rows, width);                                                                                         \
    break;
  switch (ch) {
    PI05_NORM_CASE(1)
    PI05_NORM_CASE(2)
    PI05_NORM_CASE(3)
    PI05_NORM_CASE(4)
    PI05_NORM_CASE(5)
    PI05_NORM_CASE(6)
    PI05_NORM_CASE(7)
    PI05_NORM_CASE(8)
    default:
      return false;
  }
#undef PI05_NORM_CASE
  count_launch();
  return true;
}

}  // namespace

bool layernorm_bwd_fused(const bf16* dy, const bf16* x, const bf16* w, const float* mean, const float* rstd,
                         const bf16* dres, bf16* dx, float* dw32, float* db32, int rows, int width, cudaStream_t st) {
  return launch_fused<true>(dy, x, w, mean, rstd, dres, dx, dw32, db32, rows, width, st);
}
bool rmsnorm_bwd_fused(const bf16* dy, const bf16* x, const float* w, const float* rstd, const bf16* dres, bf16* dx,
                       float* dw32, int rows, int width, cudaStream_t st) {
  return launch_fused<false>(dy, x, w, nullptr, rstd, dres, dx, dw32, nullptr, rows, width, st);
}

}  // namespace pi05
