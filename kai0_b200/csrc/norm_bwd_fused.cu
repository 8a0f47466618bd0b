// Fused LayerNorm / plain-RMSNorm backward: dx AND the weight (bias) gradient from ONE read of dy and x.
// The two-kernel version (norm_kernels.cu: *_bwd_dx_k + norm_bwd_dwdb_k) reads dy and x twice and its column-sum kernel
// ran at 9-19 % of the HBM peak (profiles/r01_ncu_layer.md).  Here persistent blocks of 8 warps walk 8-row slabs:
//   phase A: warp w owns row w of the slab — loads dy / x (all 16-byte loads of the row in flight), keeps the packed
//            values in registers for the statistics and the dx pass, and parks a bf16 copy of both in shared memory;
//   phase B: thread t owns columns t, t+256, ... and adds the slab's 8 rows from shared memory into its dw (db)
//            accumulators, which live in registers for the whole kernel;
// one atomicAdd per column per block at the end.  Arithmetic per element is identical to the two-kernel version.
#include <cstdlib>

#include "common.cuh"
#include "errors.h"
#include "kernels.h"
#include "launch.h"

namespace pi05 {
namespace {

// NW = warps per block = rows per slab.  Measured (B=32 bench shapes): width 2048 is fastest with one 16-warp block per
// SM (128 KB of shared memory), width 1152 with three 8-warp blocks per SM running out of phase (A overlaps B).

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
template <bool LN, int CH, int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 3 : 1) norm_bwd_slab_k(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                           const void* __restrict__ w_, const float* __restrict__ mean_i,
                                                           const float* __restrict__ rstd_i,
                                                           const bf16* __restrict__ dres, bf16* __restrict__ dx,
                                                           float* __restrict__ dw32, float* __restrict__ db32, int rows,
                                                           int width) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* dyS = reinterpret_cast<bf16*>(smem_raw);               // [NW][width]
  bf16* xS = dyS + static_cast<size_t>(NW) * width;            // [NW][width]
  __shared__ float meanS[NW], rstdS[NW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int MAXC = 2048 / (NW * 32);  // columns per thread in phase B (width <= 2048)
  float aw[MAXC], ab[MAXC];
#pragma unroll
  for (int j = 0; j < MAXC; ++j) aw[j] = ab[j] = 0.f;
  const float inv_w = 1.0f / static_cast<float>(width);
  const int nslab = (rows + NW - 1) / NW;
  for (int slab = blockIdx.x; slab < nslab; slab += gridDim.x) {
    const int row = slab * NW + warp;
    const bool live = row < rows;
    // ---------------- phase A: one row per warp ----------------
    if (live) {
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
      if (lane == 0) {
        meanS[warp] = mean;
        rstdS[warp] = rstd;
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = lane * 8 + k * 256;
        if (c < width) {
          *reinterpret_cast<uint4*>(dyS + warp * width + c) = pd[k];
          *reinterpret_cast<uint4*>(xS + warp * width + c) = pv[k];
          float d[8], v[8], ww[8];
          unpack8(pd[k], d);
          unpack8(pv[k], v);
          if (LN) load8(static_cast<const bf16*>(w_) + c, ww);
          else load8f(static_cast<const float*>(w_) + c, ww);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float g = LN ? d[i] * ww[i] : d[i] * (1.0f + ww[i]);
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
          float d[8], v[8], ww[8], o[8];
          unpack8(pd[k], d);
          unpack8(pv[k], v);
          if (LN) load8(static_cast<const bf16*>(w_) + c, ww);
          else load8f(static_cast<const float*>(w_) + c, ww);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xh = LN ? (v[i] - mean) * rstd : v[i] * rstd;
            const float g = LN ? d[i] * ww[i] : d[i] * (1.0f + ww[i]);
            o[i] = LN ? bfr(rstd * (g - s1 - xh * s2)) : bfr(rstd * (g - xh * s2));
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
    __syncthreads();
    // ---------------- phase B: column sums of the slab from shared memory ----------------
    const int nrow = min(NW, rows - slab * NW);
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      const int c = threadIdx.x + j * NW * 32;
      if (c < width) {
        float a = 0.f, b = 0.f;
        for (int r = 0; r < nrow; ++r) {
          const float d = __bfloat162float(dyS[r * width + c]);
          const float v = __bfloat162float(xS[r * width + c]);
          const float xh = LN ? (v - meanS[r]) * rstdS[r] : v * rstdS[r];
          a += d * xh;
          b += d;
        }
        aw[j] += a;
        ab[j] += b;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < MAXC; ++j) {
    const int c = threadIdx.x + j * NW * 32;
    if (c < width) {
      atomicAdd(dw32 + c, aw[j]);
      if (LN) atomicAdd(db32 + c, ab[j]);
    }
  }
}

// Register-accumulating variant: no shared-memory phase at all.  A lane owns the SAME columns (lane * 8 + k * 256) of every
// row its warp visits, so the weight (bias) gradient is accumulated in registers across rows while the row is still in
// registers for dx; warps never synchronise inside the row loop (one warp's loads overlap another's arithmetic), the
// residual gradient is loaded together with dy / x (one memory latency per row, not two), and the 8 warps of a block are
// reduced through shared memory once at the end (one atomicAdd per column per block).  Same arithmetic per element.
template <bool LN, int CH>
__global__ void __launch_bounds__(256, 1) norm_bwd_reg_k(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                         const void* __restrict__ w_, const float* __restrict__ mean_i,
                                                         const float* __restrict__ rstd_i, const bf16* __restrict__ dres,
                                                         bf16* __restrict__ dx, float* __restrict__ dw32,
                                                         float* __restrict__ db32, int rows, int width) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* red = reinterpret_cast<float*>(smem_raw);  // [8][width] (weight grads), then reused for the bias grads
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float aw[CH][8], ab[LN ? CH : 1][8];
#pragma unroll
  for (int k = 0; k < CH; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      aw[k][i] = 0.f;
      if (LN) ab[k][i] = 0.f;
    }
  const float inv_w = 1.0f / static_cast<float>(width);
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const int64_t off = static_cast<int64_t>(row) * width;
    uint4 pd[CH], pv[CH], pr[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int c = lane * 8 + k * 256;
      if (c < width) {
        pd[k] = *reinterpret_cast<const uint4*>(dy + off + c);
        pv[k] = *reinterpret_cast<const uint4*>(x + off + c);
        if (dres) pr[k] = *reinterpret_cast<const uint4*>(dres + off + c);
      }
    }
    const float mean = LN ? mean_i[row] : 0.f;
    const float rstd = rstd_i[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int c = lane * 8 + k * 256;
      if (c < width) {
        float d[8], v[8], ww[8];
        unpack8(pd[k], d);
        unpack8(pv[k], v);
        if (LN) load8(static_cast<const bf16*>(w_) + c, ww);
        else load8f(static_cast<const float*>(w_) + c, ww);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = LN ? (v[i] - mean) * rstd : v[i] * rstd;
          const float g = LN ? d[i] * ww[i] : d[i] * (1.0f + ww[i]);
          if (LN) s1 += g;
          s2 += g * xh;
          aw[k][i] += d[i] * xh;
          if (LN) ab[k][i] += d[i];
        }
      }
    }
    if (LN) s1 = warp_sum(s1) * inv_w;
    s2 = warp_sum(s2) * inv_w;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int c = lane * 8 + k * 256;
      if (c < width) {
        float d[8], v[8], ww[8], o[8];
        unpack8(pd[k], d);
        unpack8(pv[k], v);
        if (LN) load8(static_cast<const bf16*>(w_) + c, ww);
        else load8f(static_cast<const float*>(w_) + c, ww);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = LN ? (v[i] - mean) * rstd : v[i] * rstd;
          const float g = LN ? d[i] * ww[i] : d[i] * (1.0f + ww[i]);
          o[i] = LN ? bfr(rstd * (g - s1 - xh * s2)) : bfr(rstd * (g - xh * s2));
        }
        if (dres) {
          float r[8];
          unpack8(pr[k], r);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += r[i];
        }
        store8(dx + off + c, o);
      }
    }
  }
  // ---- block reduction of the per-warp accumulators, then one atomicAdd per column
#pragma unroll
  for (int pass = 0; pass < (LN ? 2 : 1); ++pass) {
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int c = lane * 8 + k * 256;
      if (c < width) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[warp * width + c + i] = pass == 0 ? aw[k][i] : ab[LN ? k : 0][i];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < width; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) t += red[wv * width + c];
      atomicAdd((pass == 0 ? dw32 : db32) + c, t);
    }
    __syncthreads();
  }
}

template <bool LN, int CH>
void launch_reg(const bf16* dy, const bf16* x, const void* w, const float* mean, const float* rstd, const bf16* dres,
                bf16* dx, float* dw32, float* db32, int rows, int width, int sms, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(8) * width * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(norm_bwd_reg_k<LN, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 256 * CH * 4);
    attr_set = true;
  }
  const int per_sm = (LN ? CH <= 3 : CH <= 4) ? 2 : 1;  // register budget (ptxas: RMS 128 regs at CH 4, LN 126 at CH 3)
  const int want = (rows + 7) / 8;
  const int grid = want < per_sm * sms ? want : per_sm * sms;
  launch_pdl(norm_bwd_reg_k<LN, CH>, dim3(grid), dim3(256), smem, st, dy, x, w, mean, rstd, dres, dx, dw32, db32, rows, width);
}

template <bool LN, int CH, int NW>
void launch_slab(const bf16* dy, const bf16* x, const void* w, const float* mean, const float* rstd, const bf16* dres,
                 bf16* dx, float* dw32, float* db32, int rows, int width, int sms, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(2) * NW * width * sizeof(bf16);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(norm_bwd_slab_k<LN, CH, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * NW * 256 * CH * 2);
    attr_set = true;
  }
  const int nslab = (rows + NW - 1) / NW;
  const int per_sm = NW == 8 ? (smem <= 72 * 1024 ? 3 : (smem <= 110 * 1024 ? 2 : 1)) : 1;
  const int grid = nslab < per_sm * sms ? nslab : per_sm * sms;
  launch_pdl(norm_bwd_slab_k<LN, CH, NW>, dim3(grid), dim3(NW * 32), smem, st, dy, x, w, mean, rstd, dres, dx, dw32, db32,
             rows, width);
}

template <bool LN>
bool launch_fused(const bf16* dy, const bf16* x, const void* w, const float* mean, const float* rstd, const bf16* dres,
                  bf16* dx, float* dw32, float* db32, int rows, int width, cudaStream_t st) {
  if (width % 8 != 0 || width > 2048) return false;
  const int ch = (width + 255) / 256;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  static const bool use_slab = getenv("PI05_NORM_BWD_SLAB") != nullptr;  // A/B: the round-1 shared-memory slab kernel
#define PI05_NORM_CASE(C, W)                                                                               \
  case C:                                                                                                  \
    if (use_slab) launch_slab<LN, C, W>(dy, x, w, mean, rstd, dres, dx, dw32, db32, rows, width, sms, st); \
    else launch_reg<LN, C>(dy, x, w, mean, rstd, dres, dx, dw32, db32, rows, width, sms, st);             \
    break;
  switch (ch) {
    PI05_NORM_CASE(1, 8)
    PI05_NORM_CASE(2, 8)
    PI05_NORM_CASE(3, 8)
    PI05_NORM_CASE(4, 8)
    PI05_NORM_CASE(5, 8)
    PI05_NORM_CASE(6, 16)
    PI05_NORM_CASE(7, 16)
    PI05_NORM_CASE(8, 16)
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
