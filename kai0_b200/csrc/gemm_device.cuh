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
  static constexpr int SMEM_BYTES = TILE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 8 * 32 * 128 /*epilogue staging*/;
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
  const float* bias32;      // EPI_PATCH
  const float* rowadd32;
  int rowadd_period;
  long long ld_rowadd;
  int group_m;              // tile rasterisation: m-blocks per group (tiles of a group share their A panels in L2)
  int static_sched;         // 1: static `tile += grid` walk instead of the dynamic tile ring (PI05_GEMM_STATIC=1: A/B runs)
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
  const int gm = p.group_m;
  const int group_span = gm * p.num_n;
  const int group = t / group_span;
  const int first_m = group * gm;
  const int gsz = min(gm, p.num_m - first_m);
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


// ---- coalesced output path: registers -> per-warp shared-memory staging -> row-contiguous 16-byte global stores ----
// A thread owns one accumulator ROW, so storing straight from registers makes every warp-wide store instruction touch
// 32 different 128-byte lines (32 wavefronts for 512 bytes).  Staging a 32-row x 32-column chunk in shared memory
// (128-byte row pitch, 16-byte chunks XOR-swizzled by row to stay bank-conflict free) and reading it back with lanes
// walking along the rows turns that into 4-8 wavefronts per instruction.
constexpr int STAGE_BYTES_PER_WARP = 32 * 128;

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// This lane's 32 values (its row of the chunk) as bf16: 4 x 16 B at chunk slots 0..3 of staging row `lane`.
__device__ __forceinline__ void stage_write_bf16(uint32_t stage, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 o;
    o.x = pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]);
    o.y = pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]);
    o.z = pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]);
    o.w = pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]);
    st_shared_v4(stage + lane * 128 + ((j ^ (lane & 7)) << 4), o);
  }
}
__device__ __forceinline__ void stage_write_f32(uint32_t stage, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 o;
    o.x = __float_as_uint(v[j * 4 + 0]);
    o.y = __float_as_uint(v[j * 4 + 1]);
    o.z = __float_as_uint(v[j * 4 + 2]);
    o.w = __float_as_uint(v[j * 4 + 3]);
    st_shared_v4(stage + lane * 128 + ((j ^ (lane & 7)) << 4), o);
  }
}
// Copy the staged chunk out: dst points at (first row of the warp, first column of the chunk); `rows_valid` rows and
// `cols_valid` (<= 32) columns are inside the matrix.  All 32 lanes must call this.
__device__ __forceinline__ void stage_flush_bf16(uint32_t stage, int lane, __nv_bfloat16* dst, long long ld,
                                                 int rows_valid, int cols_valid) {
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2), c = lane & 3;  // 4 lanes x 16 B = one 64-byte row segment
    const uint4 val = ld_shared_v4(stage + r * 128 + ((c ^ (r & 7)) << 4));
    if (r < rows_valid && c * 8 < cols_valid) {
      __nv_bfloat16* g = dst + r * ld + c * 8;
      if (c * 8 + 8 <= cols_valid && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        *reinterpret_cast<uint4*>(g) = val;
      } else {
        const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&val);
        for (int k = 0; k < 8; ++k)
          if (c * 8 + k < cols_valid) g[k] = e[k];
      }
    }
  }
  __syncwarp();  // the staging rows are rewritten by the next chunk
}
__device__ __forceinline__ void stage_flush_f32(uint32_t stage, int lane, float* dst, long long ld, int rows_valid,
                                                int cols_valid, int accumulate) {
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + (lane >> 3), c = lane & 7;  // 8 lanes x 16 B = one 128-byte row segment
    const uint4 val = ld_shared_v4(stage + r * 128 + ((c ^ (r & 7)) << 4));
    if (r < rows_valid && c * 4 < cols_valid) {
      float* g = dst + r * ld + c * 4;
      const float f[4] = {__uint_as_float(val.x), __uint_as_float(val.y), __uint_as_float(val.z), __uint_as_float(val.w)};
      if (c * 4 + 4 <= cols_valid && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        float4 o = make_float4(f[0], f[1], f[2], f[3]);
        if (accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(g);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(g) = o;
      } else {
        for (int k = 0; k < 4; ++k)
          if (c * 4 + k < cols_valid) g[k] = accumulate ? g[k] + f[k] : f[k];
      }
    }
  }
  __syncwarp();
}

// Coalesced load of a 32-row x 32-column bf16 chunk (the reverse of stage_flush_bf16): lanes walk along the rows into
// the staging area, then every thread picks up its own row.  Rows / columns outside the matrix read as zero.
__device__ __forceinline__ void stage_load_bf16(uint32_t stage, int lane, const __nv_bfloat16* src, long long ld,
                                                int rows_valid, int cols_valid, float (&out)[32]) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2), c = lane & 3;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < rows_valid && c * 8 < cols_valid) {
      const __nv_bfloat16* g = src + r * ld + c * 8;
      if (c * 8 + 8 <= cols_valid && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        val = *reinterpret_cast<const uint4*>(g);
      } else {
        __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(&val);
        for (int k = 0; k < 8; ++k)
          if (c * 8 + k < cols_valid) e[k] = g[k];
      }
    }
    st_shared_v4(stage + r * 128 + ((c ^ (r & 7)) << 4), val);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 w4 = ld_shared_v4(stage + lane * 128 + ((j ^ (lane & 7)) << 4));
    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      out[j * 8 + k * 2 + 0] = __uint_as_float(w[k] << 16);
      out[j * 8 + k * 2 + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
    }
  }
  __syncwarp();
}

// Drains this warp's share (column half `chalf`) of one 128 x BN accumulator tile: thread <-> row, 32 columns per
// tcgen05.ld, fused epilogue math in registers, then the staged coalesced store above.  `row0` is the warp's first
// row (row = row0 + lane); `stage` is this warp's staging area in shared memory.
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const KParams& p, int z0, int z1, int row0, int lane, int n0,
                                              uint32_t t_base, int chalf, uint32_t stage) {
  constexpr int BN_OUT = (EPI == EPI_GEGLU) ? BN / 2 : BN;
  constexpr int NCH = BN_OUT / 32;
  constexpr int CH_PER_WARP = NCH / (NUM_EPI_WARPS / 4);
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  const int rows_valid = p.M - row0;  // may be <= 0 or > 32; the flush clamps by comparison
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
      __nv_bfloat16* d = static_cast<__nv_bfloat16*>(p.D) + z0 * p.dbs + z1 * p.dbs1 + static_cast<long long>(row0) * p.ldd;
      stage_write_bf16(stage, lane, v);
      stage_flush_bf16(stage, lane, d + col, p.ldd, rows_valid, nvalid);
      stage_write_bf16(stage, lane, u);
      stage_flush_bf16(stage, lane, d + p.N + col, p.ldd, rows_valid, nvalid);
      __nv_bfloat16* d2 =
          static_cast<__nv_bfloat16*>(p.D2) + z0 * p.d2bs + z1 * p.d2bs1 + static_cast<long long>(row0) * p.ldd2;
      stage_write_bf16(stage, lane, h);
      stage_flush_bf16(stage, lane, d2 + col, p.ldd2, rows_valid, nvalid);
    } else if constexpr (EPI == EPI_GEGLU_BWD) {
      const __nv_bfloat16* gu = p.res + z0 * p.resbs + z1 * p.resbs1 + static_cast<long long>(row0) * p.ldres;
      float g[32], u[32];
      stage_load_bf16(stage, lane, gu + col, p.ldres, rows_valid, nvalid, g);
      stage_load_bf16(stage, lane, gu + p.N + col, p.ldres, rows_valid, nvalid, u);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float d = bf16_round(v[i]);                   // dH as the bf16 tensor autograd would hold
        const float a = bf16_round(gelu_tanh_f(g[i]));
        const float da = bf16_round(d * u[i]);
        u[i] = d * a;                                       // du
        v[i] = da * gelu_tanh_grad_f(g[i]);                 // dg
      }
      __nv_bfloat16* d = static_cast<__nv_bfloat16*>(p.D) + z0 * p.dbs + z1 * p.dbs1 + static_cast<long long>(row0) * p.ldd;
      stage_write_bf16(stage, lane, v);
      stage_flush_bf16(stage, lane, d + col, p.ldd, rows_valid, nvalid);
      stage_write_bf16(stage, lane, u);
      stage_flush_bf16(stage, lane, d + p.N + col, p.ldd, rows_valid, nvalid);
    } else if constexpr (EPI == EPI_GELU_BWD) {
      const __nv_bfloat16* pre = p.res + z0 * p.resbs + z1 * p.resbs1 + static_cast<long long>(row0) * p.ldres;
      float x[32];
      stage_load_bf16(stage, lane, pre + col, p.ldres, rows_valid, nvalid, x);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i]) * gelu_tanh_grad_f(x[i]);
      __nv_bfloat16* d =
          static_cast<__nv_bfloat16*>(p.D) + z0 * p.dbs + z1 * p.dbs1 + static_cast<long long>(row0) * p.ldd + col;
      stage_write_bf16(stage, lane, v);
      stage_flush_bf16(stage, lane, d, p.ldd, rows_valid, nvalid);
    } else if constexpr (EPI == EPI_PATCH) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (i < nvalid) ? __fadd_rn(v[i], p.bias32[col + i]) : 0.0f;
      if (row_ok) {
        const float* ra = p.rowadd32 + static_cast<long long>(row % p.rowadd_period) * p.ld_rowadd + col;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nvalid) v[i] = __fadd_rn(v[i], ra[i]);
      }
      __nv_bfloat16* d =
          static_cast<__nv_bfloat16*>(p.D) + z0 * p.dbs + z1 * p.dbs1 + static_cast<long long>(row0) * p.ldd + col;
      stage_write_bf16(stage, lane, v);
      stage_flush_bf16(stage, lane, d, p.ldd, rows_valid, nvalid);
    } else if constexpr (EPI == EPI_F32) {
      float* d = static_cast<float*>(p.D) + z0 * p.dbs + z1 * p.dbs1 + static_cast<long long>(row0) * p.ldd + col;
      stage_write_f32(stage, lane, v);
      stage_flush_f32(stage, lane, d, p.ldd, rows_valid, nvalid, p.accumulate);
    } else {
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
                              static_cast<long long>(row0) * p.ldd2 + col;
          stage_write_bf16(stage, lane, v);
          stage_flush_bf16(stage, lane, d2, p.ldd2, rows_valid, nvalid);
        }
        if (p.gate != nullptr && row_ok) {
          float gt[32];
          load_bf16x32(p.gate + static_cast<long long>(row / p.gate_rows) * p.ldgate + col, nvalid, gt);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i] * gt[i]);
        }
        if (row_ok) {
          float rs[32];
          load_bf16x32(p.res + z0 * p.resbs + z1 * p.resbs1 + static_cast<long long>(row) * p.ldres + col, nvalid, rs);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += rs[i];
        }
      }
      __nv_bfloat16* d =
          static_cast<__nv_bfloat16*>(p.D) + z0 * p.dbs + z1 * p.dbs1 + static_cast<long long>(row0) * p.ldd + col;
      stage_write_bf16(stage, lane, v);
      stage_flush_bf16(stage, lane, d, p.ldd, rows_valid, nvalid);
      if constexpr (EPI == EPI_BIAS_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_tanh_f(bf16_round(v[i]));
        __nv_bfloat16* d2 =
            static_cast<__nv_bfloat16*>(p.D2) + z0 * p.d2bs + z1 * p.d2bs1 + static_cast<long long>(row0) * p.ldd2 + col;
        stage_write_bf16(stage, lane, v);
        stage_flush_bf16(stage, lane, d2, p.ldd2, rows_valid, nvalid);
      }
    }
  }
}

}  // namespace gemm_detail
}  // namespace pi05
