// Device-side pieces shared by the 1-CTA (gemm_sm100.cu) and 2-CTA (gemm2_sm100.cu) tcgen05 GEMM kernels:
// tile geometry, kernel parameters, tile decoding and the fused epilogue that drains one accumulator tile out of TMEM.
#pragma once
#include "gemm.h"
#include "ptx.cuh"

namespace pi05 {
namespace gemm_detail {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int NUM_EPI_WARPS = 8;                    // two warps per TMEM lane quarter, each takes half the columns
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int GROUP_M = 8;

template <int BN>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGES = (BN == 256) ? 4 : ((BN == 128) ? 6 : 8);
  static constexpr int TILE_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
  static constexpr int SMEM_BYTES = TILE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // two accumulator stages
};

struct KParams {
  int M, N, K, batch, nz0;
  int a_mn, b_mn, a_b0, a_b1, b_b0, b_b1;
  int num_m, num_n, num_kb;
  uint32_t idesc;
  void* D;
  long long ldd, dbs, dbs1;
  void* D2;
  long long ldd2, d2bs, d2bs1;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* res;
  long long ldres, resbs, resbs1;
  const __nv_bfloat16* gate;
  int gate_rows;
  long long ldgate;
  float scale;
  int accumulate;
  uint32_t mn_lbo, mn_sbo;  // MN-major descriptor geometry (overridable for bring-up: PI05_DBG_MN_LBO/SBO)
};

struct TileCoord {
  int z, z0, z1, m_blk, n_blk;
};

__device__ __forceinline__ TileCoord decode_tile(int tile, const KParams& p) {
  const int per_batch = p.num_m * p.num_n;
  TileCoord c;
  c.z = tile / per_batch;
  c.z1 = c.z / p.nz0;
  c.z0 = c.z - c.z1 * p.nz0;
  const int t = tile - c.z * per_batch;
  const int group_span = GROUP_M * p.num_n;
  const int group = t / group_span;
  const int first_m = group * GROUP_M;
  const int gsz = min(GROUP_M, p.num_m - first_m);
  const int r = t - group * group_span;
  c.m_blk = first_m + r % gsz;
  c.n_blk = r / gsz;
  return c;
}

// ---- epilogue helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* p, int nvalid, float (&out)[32]) {
  if (nvalid == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 v = q[i];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        out[i * 8 + j * 2 + 0] = __uint_as_float(w[j] << 16);
        out[i * 8 + j * 2 + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) out[i] = (i < nvalid) ? __bfloat162float(p[i]) : 0.0f;
  }
}

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* p, int nvalid, const float (&v)[32]) {
  if (nvalid == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 o;
      o.x = pack_bf16x2(v[i * 8 + 0], v[i * 8 + 1]);
      o.y = pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
      o.z = pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]);
      o.w = pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
      q[i] = o;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) p[i] = __float2bfloat16_rn(v[i]);
  }
}

__device__ __forceinline__ void regs_to_float(const uint32_t (&r)[32], float (&f)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
}


// Drains this warp's share (column half `chalf`) of one 128 x BN accumulator tile: thread <-> row, 32 columns per
// tcgen05.ld, fused epilogue math, 16-byte global stores.
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const KParams& p, int z0, int z1, int row, bool row_ok, int n0,
                                              uint32_t t_base, int chalf) {
  constexpr int BN_OUT = (EPI == EPI_GEGLU) ? BN / 2 : BN;
  constexpr int NCH = BN_OUT / 32;
  constexpr int CH_PER_WARP = NCH / (NUM_EPI_WARPS / 4);
#pragma unroll 1
  for (int c = chalf * CH_PER_WARP; c < (chalf + 1) * CH_PER_WARP; ++c) {
    const int col = n0 + c * 32;
    if (col >= p.N) break;  // warp-uniform
    const int nvalid = min(32, p.N - col);
    uint32_t r[32];
    float v[32];
    tmem_ld32(t_base + c * 32, r);
    tmem_ld_wait();
    regs_to_float(r, v);

    if constexpr (EPI == EPI_GEGLU) {
      uint32_t r2[32];
      float u[32];
      tmem_ld32(t_base + BN / 2 + c * 32, r2);
      tmem_ld_wait();
      regs_to_float(r2, u);
      if (row_ok) {
        float h[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float g = bf16_round(v[i]);
          const float uu = bf16_round(u[i]);
          const float a = bf16_round(gelu_tanh_f(g));
          v[i] = g;
          u[i] = uu;
          h[i] = a * uu;
        }
        __nv_bfloat16* d = static_cast<__nv_bfloat16*>(p.D) + z0 * p.dbs + z1 * p.dbs1 + static_cast<long long>(row) * p.ldd;
        store_bf16x32(d + col, nvalid, v);
        store_bf16x32(d + p.N + col, nvalid, u);
        __nv_bfloat16* d2 =
            static_cast<__nv_bfloat16*>(p.D2) + z0 * p.d2bs + z1 * p.d2bs1 + static_cast<long long>(row) * p.ldd2;
        store_bf16x32(d2 + col, nvalid, h);
      }
    } else if constexpr (EPI == EPI_F32) {
      if (row_ok) {
        float* d = static_cast<float*>(p.D) + z0 * p.dbs + z1 * p.dbs1 + static_cast<long long>(row) * p.ldd + col;
        if (p.accumulate) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid) d[i] += v[i];
        } else if (nvalid == 32 && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(d)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid) d[i] = v[i];
        }
      }
    } else {
      if (row_ok) {
        if constexpr (EPI == EPI_SCALE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i]) * p.scale;
        }
        if constexpr (EPI == EPI_BIAS || EPI == EPI_BIAS_GELU) {
          float b[32];
          load_bf16x32(p.bias + col, nvalid, b);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += b[i];
        }
        if constexpr (EPI == EPI_RES) {
          if (p.bias != nullptr) {
            float b[32];
            load_bf16x32(p.bias + col, nvalid, b);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += b[i];
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i]);
          if (p.D2 != nullptr) {  // keep the pre-gate linear output for the backward of the gate
            __nv_bfloat16* d2 = static_cast<__nv_bfloat16*>(p.D2) + z0 * p.d2bs + z1 * p.d2bs1 +
                                static_cast<long long>(row) * p.ldd2 + col;
            store_bf16x32(d2, nvalid, v);
          }
          if (p.gate != nullptr) {
            float gt[32];
            load_bf16x32(p.gate + static_cast<long long>(row / p.gate_rows) * p.ldgate + col, nvalid, gt);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i] * gt[i]);
          }
          float rs[32];
          load_bf16x32(p.res + z0 * p.resbs + z1 * p.resbs1 + static_cast<long long>(row) * p.ldres + col, nvalid, rs);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += rs[i];
        }
        __nv_bfloat16* d =
            static_cast<__nv_bfloat16*>(p.D) + z0 * p.dbs + z1 * p.dbs1 + static_cast<long long>(row) * p.ldd + col;
        store_bf16x32(d, nvalid, v);
        if constexpr (EPI == EPI_BIAS_GELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_tanh_f(bf16_round(v[i]));
          __nv_bfloat16* d2 =
              static_cast<__nv_bfloat16*>(p.D2) + z0 * p.d2bs + z1 * p.d2bs1 + static_cast<long long>(row) * p.ldd2 + col;
          store_bf16x32(d2, nvalid, v);
        }
      }
    }
  }
}

}  // namespace gemm_detail
}  // namespace pi05
