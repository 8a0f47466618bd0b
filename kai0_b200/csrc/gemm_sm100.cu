// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma (fp32 accumulators in TMEM,
//   double-buffered) -> tcgen05.ld epilogue with the fused element-wise tails listed in gemm.h.
//
// It replaces every bf16 nn.Linear / torch.matmul call site of the reference hot path
// (modeling_gemma.py:125,243,250,295-297,328; modeling_siglip.py:333-341,378-380,412,429-431;
//  modeling_paligemma.py:96-99) and all of their autograd counterparts (dgrad / wgrad), which is why both
// operands may be K-major or MN-major.
//
// Warp roles (320 threads):  warp 0 = TMA producer (1 thread)   warp 1 = MMA issuer (1 thread) + TMEM owner
//                            warps 2..9 = epilogue (TMEM lane quarter = warp_idx & 3, column half = (warp-2)>>2)
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "errors.h"
#include "gemm.h"
#include "gemm_device.cuh"
#include "gemm_host.h"
#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace pi05 {

using namespace gemm_detail;

namespace {

struct ProfEntry {
  cudaEvent_t a, b;
  int M, N, K, batch, epi, majors;
};
bool g_prof_on = false;
std::vector<ProfEntry> g_prof;


// ---- the kernel --------------------------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const KParams p,
            unsigned int* __restrict__ sched) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024B-aligned tiles
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + C::STAGES * A_STAGE_BYTES;
  const uint32_t bar_base = smem_base + C::TILE_BYTES;
  // barrier slots (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr
  const uint32_t full_bar0 = bar_base;
  const uint32_t empty_bar0 = bar_base + 8 * C::STAGES;
  const uint32_t tfull_bar0 = bar_base + 16 * C::STAGES;
  const uint32_t tempty_bar0 = tfull_bar0 + 16;
  const uint32_t tmem_slot = tempty_bar0 + 16;
  // dynamic tile schedule (same scheme as gemm2_sm100.cu, CTA-local): the producer thread draws tile indices from a
  // global counter and publishes them through a 4-slot ring; consumers = MMA warp + epilogue warps
  constexpr int TQ = 4;
  const uint32_t tq_full0 = tmem_slot + 8;
  const uint32_t tq_empty0 = tq_full0 + 8 * TQ;
  const uint32_t tq_slot0 = tq_empty0 + 8 * TQ;
  static_assert(16 * C::STAGES + 32 + 8 + 16 * TQ + 4 * TQ <= 256, "barrier area overflow");

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // first tile index drawn before the prologue so the atomic's latency is hidden (see gemm2_sm100.cu)
  unsigned int first_raw = 0;
  const bool dyn = p.static_sched == 0;
  if (threadIdx.x == 0 && dyn) first_raw = atomicAdd(sched, 1u);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar0 + 8 * s, 1);
      mbar_init(empty_bar0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar0 + 8 * a, 1);
      mbar_init(tempty_bar0 + 8 * a, 32 * NUM_EPI_WARPS);
    }
    for (int q = 0; q < TQ; ++q) {
      mbar_init(tq_full0 + 8 * q, 1);
      mbar_init(tq_empty0 + 8 * q, 1 + NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above touched only shared memory / TMEM / the kernel parameters: it overlaps the previous kernel's tail
  pdl_enter();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int total_tiles = p.num_m * p.num_n * p.batch;
  constexpr int BN_OUT = (EPI == EPI_GEGLU) ? BN / 2 : BN;  // output columns covered by one tile
  int tq_slot = 0;
  uint32_t tq_phase = 0;
  // consumer side of the tile ring (whole converged warp): next index, then release the slot
  int stile = blockIdx.x;
  auto ring_next = [&]() -> int {
    if (!dyn) {
      const int t = stile < total_tiles ? stile : -1;
      stile += gridDim.x;
      return t;
    }
    mbar_wait(tq_full0 + 8 * tq_slot, tq_phase);
    int t;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(t) : "r"(tq_slot0 + 4 * tq_slot) : "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(tq_empty0 + 8 * tq_slot);
    if (++tq_slot == TQ) {
      tq_slot = 0;
      tq_phase ^= 1;
    }
    return t;
  };

  if (warp_idx == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto to_tile = [&](unsigned int raw) -> int {
        return raw < static_cast<unsigned int>(total_tiles) ? static_cast<int>(raw) : -1;
      };
      auto publish = [&](int t) {
        mbar_wait(tq_empty0 + 8 * tq_slot, tq_phase ^ 1);
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(tq_slot0 + 4 * tq_slot), "r"(t) : "memory");
        mbar_arrive(tq_full0 + 8 * tq_slot);
        if (++tq_slot == TQ) {
          tq_slot = 0;
          tq_phase ^= 1;
        }
      };
      const int pub_kb = p.num_kb > 8 ? 8 : p.num_kb - 1;
      int tile;
      if (dyn) {
        tile = to_tile(first_raw);
        publish(tile);
      } else {
        tile = static_cast<int>(blockIdx.x) < total_tiles ? static_cast<int>(blockIdx.x) : -1;
      }
      while (tile >= 0) {
        unsigned int next_raw = 0;
        if (dyn) next_raw = atomicAdd(sched, 1u);  // in flight while this tile's first loads are issued
        int next_tile = -1;
        const TileCoord tc = decode_tile(tile, p);
        const int m0 = tc.m_blk * BM;
        const int n0 = tc.n_blk * BN_OUT;
        const int za0 = p.a_b0 ? tc.z0 : 0, za1 = p.a_b1 ? tc.z1 : 0;
        const int zb0 = p.b_b0 ? tc.z0 : 0, zb1 = p.b_b1 ? tc.z1 : 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
          const uint32_t full = full_bar0 + 8 * stage;
          mbar_arrive_expect_tx(full, A_STAGE_BYTES + C::B_STAGE_BYTES);
          const uint32_t sa = smem_a0 + stage * A_STAGE_BYTES;
          const uint32_t sb = smem_b0 + stage * C::B_STAGE_BYTES;
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_4d(sa, &tma_a, full, k0, m0, za0, za1);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_4d(sa + i * (BK * 128), &tma_a, full, m0 + 64 * i, k0, za0, za1);
          }
          if (EPI == EPI_GEGLU) {
            tma_load_4d(sb, &tma_b, full, k0, n0, zb0, zb1);
            tma_load_4d(sb + (BN / 2) * 128, &tma_b, full, k0, p.N + n0, zb0, zb1);
          } else if (!p.b_mn) {
            tma_load_4d(sb, &tma_b, full, k0, n0, zb0, zb1);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_4d(sb + i * (BK * 128), &tma_b, full, n0 + 64 * i, k0, zb0, zb1);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
          if (dyn && kb == pub_kb) {
            next_tile = to_tile(next_raw);
            publish(next_tile);
          }
        }
        tile = dyn ? next_tile : (tile + static_cast<int>(gridDim.x) < total_tiles ? tile + static_cast<int>(gridDim.x) : -1);
      }
      // the last CTA to run dry returns the counter pair to zero for the next launch that uses it
      if (dyn && atomicAdd(sched + 1, 1u) == gridDim.x - 1u) {
        sched[0] = 0u;
        sched[1] = 0u;
        __threadfence();
      }
    }
  } else if (warp_idx == 1) {
    // ============================== MMA issuer ==============================
    // Whole warp walks the loop with warp-uniform values; one elected lane issues (see gemm2_sm100.cu for why).
    {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // Per-operand descriptor geometry.  K-major: rows of 128 B, 8-row groups 1024 B apart, k-step = 32 B.
      // MN-major: 64-element (128 B) MN atoms, k rows 128 B apart, 8-k groups 1024 B apart (SBO),
      //           MN atoms BK*128 B apart (LBO), k-step (16 k rows) = 2048 B.
      const uint32_t a_lbo = p.a_mn ? p.mn_lbo : 16, a_sbo = p.a_mn ? p.mn_sbo : 1024, a_kstep = p.a_mn ? 2048 : 32;
      const uint32_t b_lbo = p.b_mn ? p.mn_lbo : 16, b_sbo = p.b_mn ? p.mn_sbo : 1024, b_kstep = p.b_mn ? 2048 : 32;
      const uint32_t a_hi = smem_desc_hi_sw128(a_sbo), b_hi = smem_desc_hi_sw128(b_sbo);
      const uint32_t a_step = a_kstep >> 4, b_step = b_kstep >> 4;
      const uint32_t idesc = p.idesc;
      const int num_kb = p.num_kb;
      for (;;) {
        if (ring_next() < 0) break;
        mbar_wait(tempty_bar0 + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_lo = smem_desc_lo(smem_a0 + stage * A_STAGE_BYTES, a_lbo);
          const uint32_t b_lo = smem_desc_lo(smem_b0 + stage * C::B_STAGE_BYTES, b_lbo);
          if (elect_one()) {
            umma_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, kb > 0 ? 1u : 0u);
#pragma unroll
            for (int j = 1; j < BK / 16; ++j)
              umma_lohi(d_tmem, a_lo + j * a_step, a_hi, b_lo + j * b_step, b_hi, idesc, 1u);
            umma_commit(empty_bar0 + 8 * stage);  // frees the smem slot when these MMAs retire
            if (kb == num_kb - 1) umma_commit(tfull_bar0 + 8 * acc);
          }
          __syncwarp();
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ============================== epilogue ==============================
    const int q = warp_idx & 3;  // TMEM lane quarter this warp may read
    const int chalf = (warp_idx - 2) >> 2;  // which half of the tile's column chunks this warp drains
    int acc = 0;
    uint32_t acc_phase = 0;
    for (;;) {
      const int tile = ring_next();
      if (tile < 0) break;
      const TileCoord tc = decode_tile(tile, p);
      const int row0 = tc.m_blk * BM + q * 32;
      const int n0 = tc.n_blk * BN_OUT;
      mbar_wait(tfull_bar0 + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      epilogue_tile<BN, EPI>(p, tc.z0, tc.z1, row0, lane, n0, t_base, chalf,
                             bar_base + 256 + (warp_idx - 2) * STAGE_BYTES_PER_WARP);
      tc_fence_before();
      mbar_arrive(tempty_bar0 + 8 * acc);  // 128 arrivals release this accumulator stage to the MMA warp
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
using EncodeFn = PFN_cuTensorMapEncodeTiled_v12000;

EncodeFn get_encode_fn() {
  static EncodeFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) ptr = nullptr;
    return reinterpret_cast<EncodeFn>(ptr);
  }();
  return fn;
}

using TmapKey = std::tuple<const void*, int, long long, long long, long long, long long, long long, long long,
                           long long, int>;
std::map<TmapKey, CUtensorMap>& tmap_cache() {
  static std::map<TmapKey, CUtensorMap> c;
  return c;
}
std::mutex& tmap_mutex() {
  static std::mutex m;
  return m;
}

// major 0: stored [rows, K] (K contiguous)  -> dims {K, rows, nz0, nz1}, box {64, box_rows, 1, 1}
// major 1: stored [K, rows] (rows contiguous) -> dims {rows, K, nz0, nz1}, box {64, 64, 1, 1}
// A batch level whose stride is 0 is collapsed to extent 1 (the kernel then passes coordinate 0).
bool make_tmap(CUtensorMap* out, const void* ptr, int major, long long rows, long long K, long long ld, long long nz0,
               long long bs0, long long nz1, long long bs1, int box_rows, char* err, int err_len) {
  if (bs0 == 0) nz0 = 1;
  if (bs1 == 0) nz1 = 1;
  TmapKey key{ptr, major, rows, K, ld, nz0, bs0, nz1, bs1, box_rows};
  {
    std::lock_guard<std::mutex> g(tmap_mutex());
    auto it = tmap_cache().find(key);
    if (it != tmap_cache().end()) {
      *out = it->second;
      return true;
    }
  }
  EncodeFn enc = get_encode_fn();
  if (!enc) {
    if (err) snprintf(err, err_len, "cuTensorMapEncodeTiled entry point unavailable");
    return false;
  }
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  cuuint32_t box[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (major == 0) {
    dims[0] = K;
    dims[1] = rows;
    box[0] = BK;
    box[1] = box_rows;
  } else {
    dims[0] = rows;
    dims[1] = K;
    box[0] = 64;
    box[1] = BK;
  }
  dims[2] = nz0;
  dims[3] = nz1;
  box[2] = 1;
  box[3] = 1;
  strides[0] = static_cast<cuuint64_t>(ld) * 2;
  strides[1] = nz0 > 1 ? static_cast<cuuint64_t>(bs0) * 2 : strides[0] * dims[1];
  strides[2] = nz1 > 1 ? static_cast<cuuint64_t>(bs1) * 2 : strides[1] * dims[2];
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15)) {
    if (err)
      snprintf(err, err_len, "gemm operand not 16B aligned (ptr %p ld %lld batch strides %lld %lld)", ptr, ld, bs0, bs1);
    return false;
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err)
      snprintf(err, err_len,
               "cuTensorMapEncodeTiled failed (%d): major %d rows %lld K %lld ld %lld nz0 %lld bs0 %lld nz1 %lld bs1 %lld",
               static_cast<int>(r), major, rows, K, ld, nz0, bs0, nz1, bs1);
    return false;
  }
  std::lock_guard<std::mutex> g(tmap_mutex());
  // A descriptor is a pure function of its key (pointer + geometry), so entries never go stale; the bound only keeps a
  // process that streams ever-new buffers through pi05_gemm from growing the map without limit (callers hold copies,
  // never references, so dropping everything is safe).
  if (tmap_cache().size() >= (1u << 16)) tmap_cache().clear();
  tmap_cache()[key] = *out;
  return true;
}

}  // namespace

unsigned int* next_sched_counter() {
  constexpr int kSlots = 64, kStrideWords = 32;  // one 128-byte line per launch: no false sharing between overlapping kernels
  static std::mutex mu;
  static std::map<int, unsigned int*> pools;
  static std::map<int, unsigned int> cursor;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> g(mu);
  auto it = pools.find(dev);
  if (it == pools.end()) {
    unsigned int* p = nullptr;
    if (cudaMalloc(&p, kSlots * kStrideWords * sizeof(unsigned int)) != cudaSuccess) return nullptr;
    if (cudaMemset(p, 0, kSlots * kStrideWords * sizeof(unsigned int)) != cudaSuccess) return nullptr;
    it = pools.emplace(dev, p).first;
    cursor[dev] = 0;
  }
  unsigned int& c = cursor[dev];
  unsigned int* out = it->second + static_cast<size_t>(c % kSlots) * kStrideWords;
  ++c;
  return out;
}

namespace {

int num_sms() {
  static int n = [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

template <int BN, int EPI>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const KParams& kp, cudaStream_t stream, char* err,
           int err_len) {
  using C = Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      if (err) snprintf(err, err_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 2;
    }
    configured = true;
  }
  const int total = kp.num_m * kp.num_n * kp.batch;
  const int grid = total < num_sms() ? total : num_sms();
  unsigned int* sched = next_sched_counter();
  if (sched == nullptr) {
    if (err) snprintf(err, err_len, "gemm: tile-scheduler counters unavailable");
    return 2;
  }
  launch_pdl(gemm_kernel<BN, EPI>, dim3(grid), dim3(NUM_THREADS), C::SMEM_BYTES, stream, ta, tb, kp, sched); count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) snprintf(err, err_len, "gemm launch: %s", cudaGetErrorString(e));
    return 3;
  }
  return 0;
}

template <int BN>
int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const KParams& kp, cudaStream_t s, char* err,
                 int err_len) {
  switch (epi) {
    case EPI_STORE: return launch<BN, EPI_STORE>(ta, tb, kp, s, err, err_len);
    case EPI_SCALE: return launch<BN, EPI_SCALE>(ta, tb, kp, s, err, err_len);
    case EPI_BIAS: return launch<BN, EPI_BIAS>(ta, tb, kp, s, err, err_len);
    case EPI_BIAS_GELU: return launch<BN, EPI_BIAS_GELU>(ta, tb, kp, s, err, err_len);
    case EPI_RES: return launch<BN, EPI_RES>(ta, tb, kp, s, err, err_len);
    case EPI_GEGLU: return launch<BN, EPI_GEGLU>(ta, tb, kp, s, err, err_len);
    case EPI_F32: return launch<BN, EPI_F32>(ta, tb, kp, s, err, err_len);
    case EPI_GEGLU_BWD: return launch<BN, EPI_GEGLU_BWD>(ta, tb, kp, s, err, err_len);
    case EPI_GELU_BWD: return launch<BN, EPI_GELU_BWD>(ta, tb, kp, s, err, err_len);
    case EPI_PATCH: return launch<BN, EPI_PATCH>(ta, tb, kp, s, err, err_len);
    default:
      if (err) snprintf(err, err_len, "unknown epilogue %d", epi);
      return 1;
  }
}

}  // namespace

namespace {
// Finish of the small-M split-K path: sum the fp32 partials in a fixed order, then the STORE / RES epilogue.
__global__ void __launch_bounds__(256) splitk_finish_k(const float* __restrict__ ws, int splits, int M, int N, int epi,
                                                       __nv_bfloat16* __restrict__ D, long long ldd,
                                                       __nv_bfloat16* __restrict__ D2, long long ldd2,
                                                       const __nv_bfloat16* __restrict__ bias,
                                                       const __nv_bfloat16* __restrict__ res, long long ldres,
                                                       const __nv_bfloat16* __restrict__ gate, int gate_rows,
                                                       long long ldgate) {
  pdl_enter();
  const long long total = static_cast<long long>(M) * N;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(idx / N), n = static_cast<int>(idx % N);
    float acc = 0.f;
    for (int z = 0; z < splits; ++z) acc += ws[z * total + idx];
    if (epi == EPI_RES) {
      if (bias) acc += __bfloat162float(bias[n]);
      float t = bf16_round(acc);
      if (D2) D2[m * ldd2 + n] = __float2bfloat16_rn(t);
      if (gate) t = bf16_round(t * __bfloat162float(gate[static_cast<long long>(m / gate_rows) * ldgate + n]));
      acc = t + __bfloat162float(res[m * ldres + n]);
    }
    D[m * ldd + n] = __float2bfloat16_rn(acc);
  }
}
// Decode: split-K finish + EPI_RES epilogue + the adaptive RMSNorm of the result in ONE kernel (one warp per row, the
// row stays in registers between the residual add and the normalisation).  Replaces splitk_finish_k + rmsnorm_fwd_k.
template <int CH>
__global__ void __launch_bounds__(CH * 32) splitk_finish_norm_k(const float* __restrict__ ws, int splits, int M, int N,
                                                                __nv_bfloat16* __restrict__ D, long long ldd,
                                                                const __nv_bfloat16* __restrict__ res, long long ldres,
                                                                const __nv_bfloat16* __restrict__ gate, int gate_rows,
                                                                long long ldgate, const float* __restrict__ mod,
                                                                int norm_rpb, __nv_bfloat16* __restrict__ y,
                                                                __nv_bfloat16* __restrict__ gate_out) {
  // one block per output row, warp k owns columns [256k, 256k + 256): every load of the row is in flight at once
  pdl_enter();
  __shared__ float ssum[CH];
  const int row = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(M) * N;
  const int c = lane * 8 + k * 256;
  const bool live = c < N;
  float v[8];
  float ss = 0.f;
  const int b = row / norm_rpb;
  const float* m = mod + static_cast<long long>(b) * 3 * N;
  float sc[8], sh[8];
  if (live) {
    // issue the (independent) modulation / residual / gate loads before the partial-sum chain
    {
      const float4 s0 = *reinterpret_cast<const float4*>(m + c), s1 = *reinterpret_cast<const float4*>(m + c + 4);
      const float4 h0 = *reinterpret_cast<const float4*>(m + N + c), h1 = *reinterpret_cast<const float4*>(m + N + c + 4);
      sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
      sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
    }
    const uint4 rv = *reinterpret_cast<const uint4*>(res + row * ldres + c);
    uint4 gv = make_uint4(0, 0, 0, 0);
    if (gate) gv = *reinterpret_cast<const uint4*>(gate + static_cast<long long>(row / gate_rows) * ldgate + c);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const float* p0 = ws + static_cast<long long>(row) * N + c;
    int z = 0;
    for (; z + 4 <= splits; z += 4) {  // four splits (8 x 16-byte loads) in flight per step, summed in z order
      float4 a[4], bq[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4* p = reinterpret_cast<const float4*>(p0 + (z + j) * total);
        a[j] = p[0];
        bq[j] = p[1];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[0] += a[j].x; acc[1] += a[j].y; acc[2] += a[j].z; acc[3] += a[j].w;
        acc[4] += bq[j].x; acc[5] += bq[j].y; acc[6] += bq[j].z; acc[7] += bq[j].w;
      }
    }
    for (; z < splits; ++z) {
      const float4* p = reinterpret_cast<const float4*>(p0 + z * total);
      const float4 a = p[0], bq = p[1];
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      acc[4] += bq.x; acc[5] += bq.y; acc[6] += bq.z; acc[7] += bq.w;
    }
    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
    const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float o2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 2 * j + h;
        float t = bf16_round(acc[i]);
        if (gate) t = bf16_round(t * (h == 0 ? __uint_as_float(gw[j] << 16) : __uint_as_float(gw[j] & 0xFFFF0000u)));
        const float r = h == 0 ? __uint_as_float(rw[j] << 16) : __uint_as_float(rw[j] & 0xFFFF0000u);
        o2[h] = bf16_round(t + r);  // the bf16 value the next layer sees
        v[i] = o2[h];
        ss += o2[h] * o2[h];
      }
      __nv_bfloat162 pk = __floats2bfloat162_rn(o2[0], o2[1]);
      ow[j] = *reinterpret_cast<uint32_t*>(&pk);
    }
    *reinterpret_cast<uint4*>(D + row * ldd + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) ssum[k] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int j = 0; j < CH; ++j) tot += ssum[j];
  const float rstd = rsqrtf(tot / N + 1e-6f);  // modeling_gemma.py:68-70
  if (live) {
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a0 = __fadd_rn(__fmul_rn(__fmul_rn(v[2 * j], rstd), __fadd_rn(1.0f, sc[2 * j])), sh[2 * j]);  // :102
      const float a1 = __fadd_rn(__fmul_rn(__fmul_rn(v[2 * j + 1], rstd), __fadd_rn(1.0f, sc[2 * j + 1])), sh[2 * j + 1]);
      __nv_bfloat162 pk = __floats2bfloat162_rn(a0, a1);
      ow[j] = *reinterpret_cast<uint32_t*>(&pk);
    }
    *reinterpret_cast<uint4*>(y + static_cast<long long>(row) * N + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    if (gate_out != nullptr && (row % norm_rpb) == 0) {
      uint32_t gw2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 gg = *reinterpret_cast<const float2*>(m + 2 * N + c + 2 * j);
        __nv_bfloat162 pk = __floats2bfloat162_rn(gg.x, gg.y);
        gw2[j] = *reinterpret_cast<uint32_t*>(&pk);
      }
      *reinterpret_cast<uint4*>(gate_out + static_cast<long long>(b) * N + c) = make_uint4(gw2[0], gw2[1], gw2[2], gw2[3]);
    }
  }
}

// Decode: split-K finish of the fused qkv projection + RoPE + q/k/v split (rope_pack_fwd_k's arithmetic) in one kernel.
// One block per token row, one warp per 256-wide head slot (H query heads, K, V); lanes 0-15 hold the first half of
// the head, lanes 16-31 the second half (rotate_half pairs are exchanged with one shuffle).
__global__ void __launch_bounds__(1024) splitk_finish_rope_k(const float* __restrict__ ws, int splits, int M, int N,
                                                             GemmRope r) {
  pdl_enter();
  const int row = blockIdx.x, slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(M) * N;
  const int c = lane * 8;
  const float* p0 = ws + static_cast<long long>(row) * N + slot * r.hd + c;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int z = 0;
  for (; z + 4 <= splits; z += 4) {
    float4 a[4], bq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4* p = reinterpret_cast<const float4*>(p0 + (z + j) * total);
      a[j] = p[0];
      bq[j] = p[1];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[0] += a[j].x; acc[1] += a[j].y; acc[2] += a[j].z; acc[3] += a[j].w;
      acc[4] += bq[j].x; acc[5] += bq[j].y; acc[6] += bq[j].z; acc[7] += bq[j].w;
    }
  }
  for (; z < splits; ++z) {
    const float4* p = reinterpret_cast<const float4*>(p0 + z * total);
    const float4 a = p[0], bq = p[1];
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += bq.x; acc[5] += bq.y; acc[6] += bq.z; acc[7] += bq.w;
  }
  float x[8], other[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = bf16_round(acc[i]);  // the projection's bf16 output
    other[i] = __shfl_xor_sync(0xffffffffu, x[i], 16);
  }
  const int b = row / r.T, t = row % r.T;
  float o[8];
  if (slot == r.H + 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = x[i];
  } else {
    const int half = r.hd / 2;
    const int pidx = (r.pos_mode == 0 ? r.pos[row] : r.nvalid[b] + t) + 1;
    const int cc = c & (half - 1);
    const uint4 cv = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(r.cos_t) +
                                                     static_cast<long long>(pidx) * half + cc);
    const uint4 sv = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(r.sin_t) +
                                                     static_cast<long long>(pidx) * half + cc);
    const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w}, sw[4] = {sv.x, sv.y, sv.z, sv.w};
    const bool first = lane < 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float cs = (i & 1) ? __uint_as_float(cw[i >> 1] & 0xFFFF0000u) : __uint_as_float(cw[i >> 1] << 16);
      const float sn = (i & 1) ? __uint_as_float(sw[i >> 1] & 0xFFFF0000u) : __uint_as_float(sw[i >> 1] << 16);
      // q*cos + rotate_half(q)*sin in bf16 ops (modeling_gemma.py:170-194): first half pairs with -second, second with +first
      o[i] = first ? bf16_round(x[i] * cs) + bf16_round(-other[i] * sn) : bf16_round(x[i] * cs) + bf16_round(other[i] * sn);
    }
  }
  __nv_bfloat16* dst;
  if (slot < r.H) dst = static_cast<__nv_bfloat16*>(r.Q) + (static_cast<long long>(row) * r.H + slot) * r.hd + c;
  else
    dst = static_cast<__nv_bfloat16*>(slot == r.H ? r.K : r.V) +
          (static_cast<long long>(b) * r.kv_len + r.key_off + t) * r.hd + c;
  uint32_t ow[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 pk = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
    ow[j] = *reinterpret_cast<uint32_t*>(&pk);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
}

// RES finish of the few-tile long-K path (no bias / gate / second output): D = bf(res + bf(sum of partials)), 4 outputs per
// thread, 16-byte partial loads, 8-byte residual loads and stores, fixed summation order.
__global__ void __launch_bounds__(256) splitk_sum_res4_k(const float* __restrict__ ws, int splits, long long total4, int n4,
                                                         __nv_bfloat16* __restrict__ D, long long ldd,
                                                         const __nv_bfloat16* __restrict__ res, long long ldres) {
  pdl_enter();
  const long long total = total4 * 4;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total4;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 acc = reinterpret_cast<const float4*>(ws)[q];
    for (int z = 1; z < splits; ++z) {
      const float4 v = reinterpret_cast<const float4*>(ws + z * total)[q];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    const long long m = q / n4;
    const int n = static_cast<int>(q % n4) * 4;
    const uint2 r = *reinterpret_cast<const uint2*>(res + m * ldres + n);
    const float r0 = __uint_as_float(r.x << 16), r1 = __uint_as_float(r.x & 0xFFFF0000u);
    const float r2 = __uint_as_float(r.y << 16), r3 = __uint_as_float(r.y & 0xFFFF0000u);
    __nv_bfloat162 lo = __floats2bfloat162_rn(bf16_round(acc.x) + r0, bf16_round(acc.y) + r1);
    __nv_bfloat162 hi = __floats2bfloat162_rn(bf16_round(acc.z) + r2, bf16_round(acc.w) + r3);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(D + m * ldd + n) = o;
  }
}

// STORE-only finish of the wave-quantisation path: 4 outputs per thread, 16-byte partial loads, fixed summation order.
__global__ void __launch_bounds__(256) splitk_sum_store4_k(const float* __restrict__ ws, int splits, long long total4,
                                                           int n4, __nv_bfloat16* __restrict__ D, long long ldd) {
  pdl_enter();
  const long long total = total4 * 4;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total4;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 acc = reinterpret_cast<const float4*>(ws)[q];
    for (int z = 1; z < splits; ++z) {
      const float4 v = reinterpret_cast<const float4*>(ws + z * total)[q];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    const long long m = q / n4;
    const int n = static_cast<int>(q % n4) * 4;
    __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(D + m * ldd + n) = o;
  }
}
}  // namespace

static int gemm_bf16_impl(const GemmArgs& a, cudaStream_t stream, char* err, int err_len, bool* norm_done);

int gemm_bf16(const GemmArgs& a, cudaStream_t stream, char* err, int err_len) {
  bool norm_done = false;
  const int rc = gemm_bf16_impl(a, stream, err, err_len, &norm_done);
  if (rc == 0 && a.rope != nullptr && !norm_done) {
    const GemmRope& r = *a.rope;  // the fused finish+RoPE kernel was not applicable: plain kernel on the stored qkv
    rope_pack_fwd(static_cast<const bf16*>(a.D), r.T, r.H, r.hd, r.pos, r.nvalid, r.pos_mode,
                  static_cast<const bf16*>(r.cos_t), static_cast<const bf16*>(r.sin_t), static_cast<bf16*>(r.Q),
                  static_cast<bf16*>(r.K), static_cast<bf16*>(r.V), r.key_off, r.kv_len, r.batch, stream);
  }
  if (rc == 0 && a.norm_mod != nullptr && !norm_done) {
    // the fused finish+norm kernel was not applicable: plain kernel, same arithmetic
    rmsnorm_fwd(static_cast<const bf16*>(a.D), nullptr, a.norm_mod, a.norm_rows_per_batch, static_cast<bf16*>(a.norm_out),
                nullptr, static_cast<bf16*>(a.norm_gate_out), a.M, a.N, 1e-6f, stream);
  }
  return rc;
}

static int gemm_bf16_impl(const GemmArgs& a, cudaStream_t stream, char* err, int err_len, bool* norm_done) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.batch <= 0) {
    if (err) snprintf(err, err_len, "gemm: empty problem M=%d N=%d K=%d batch=%d", a.M, a.N, a.K, a.batch);
    return 1;
  }
  // ---- small-M latency path (decode): one 128-row tile, few N tiles, long serial K loop -> split K across CTAs
  if (a.M <= BM && a.batch == 1 && a.splitk_ws != nullptr && a.a_major == 0 && a.b_major == 0 &&
      (a.epilogue == EPI_STORE || a.epilogue == EPI_RES) && a.K % BK == 0 && a.block_n == 0) {
    const int tiles = (a.N + 127) / 128;
    const int kb = a.K / BK;
    int s = 1;
    while (s * 2 <= 32 && kb % (s * 2) == 0 && kb / (s * 2) >= 4 && tiles * s * 2 <= 160) s *= 2;
    if (s > 1 && static_cast<size_t>(s) * a.M * a.N * sizeof(float) <= a.splitk_ws_bytes) {
      const int Kc = a.K / s;
      GemmArgs part = a;
      part.splitk_ws = nullptr;
      part.K = Kc;
      part.batch = s;
      part.a_batch_stride = Kc;
      part.b_batch_stride = Kc;
      part.epilogue = EPI_F32;
      part.accumulate = 0;
      part.D = a.splitk_ws;
      part.ldd = a.N;
      part.d_batch_stride = static_cast<int64_t>(a.M) * a.N;
      part.block_n = 128;
      part.bias = nullptr;
      part.res = nullptr;
      part.gate = nullptr;
      part.D2 = nullptr;
      part.norm_mod = nullptr;
      part.rope = nullptr;
      int rc = gemm_bf16(part, stream, err, err_len);
      if (rc != 0) return rc;
      const long long total = static_cast<long long>(a.M) * a.N;
      const int grid = static_cast<int>((total + 255) / 256);
      const bool fuse_norm = a.norm_mod != nullptr && a.epilogue == EPI_RES && a.bias == nullptr && a.D2 == nullptr &&
                             a.N % 8 == 0 && a.N <= 1024 && a.ldd % 8 == 0 && a.ldres % 8 == 0 &&
                             (a.gate == nullptr || a.ldgate % 8 == 0);
      if (a.rope != nullptr && a.epilogue == EPI_STORE && a.rope->hd == 256 && a.N == (a.rope->H + 2) * a.rope->hd &&
          a.rope->H + 2 <= 32 && a.bias == nullptr) {
        launch_pdl(splitk_finish_rope_k, dim3(a.M), dim3(32 * (a.rope->H + 2)), 0, stream, a.splitk_ws, s, a.M, a.N,
                   *a.rope);
        count_launch();
        *norm_done = true;
        return 0;
      }
      if (fuse_norm) {
        const int ch = (a.N + 255) / 256;
        const dim3 g2(a.M);
#define PI05_FN(C)                                                                                                     \
  launch_pdl(splitk_finish_norm_k<C>, g2, dim3(C * 32), 0, stream, a.splitk_ws, s, a.M, a.N,                             \
             static_cast<__nv_bfloat16*>(a.D), a.ldd, static_cast<const __nv_bfloat16*>(a.res), a.ldres,             \
             static_cast<const __nv_bfloat16*>(a.gate), a.gate_rows > 0 ? a.gate_rows : 1, a.ldgate, a.norm_mod,    \
             a.norm_rows_per_batch > 0 ? a.norm_rows_per_batch : a.M, static_cast<__nv_bfloat16*>(a.norm_out),       \
             static_cast<__nv_bfloat16*>(a.norm_gate_out))
        if (ch == 1) PI05_FN(1);
        else if (ch == 2) PI05_FN(2);
        else if (ch == 3) PI05_FN(3);
        else PI05_FN(4);
#undef PI05_FN
        count_launch();
        *norm_done = true;
        return 0;
      }
      launch_pdl(splitk_finish_k, dim3(grid), dim3(256), 0, stream, a.splitk_ws, s, a.M, a.N, a.epilogue, static_cast<__nv_bfloat16*>(a.D), a.ldd,
                                                static_cast<__nv_bfloat16*>(a.D2), a.ldd2,
                                                static_cast<const __nv_bfloat16*>(a.bias),
                                                static_cast<const __nv_bfloat16*>(a.res), a.ldres,
                                                static_cast<const __nv_bfloat16*>(a.gate), a.gate_rows > 0 ? a.gate_rows : 1,
                                                a.ldgate);
      count_launch();
      return 0;
    }
  }
  if (a.epilogue == EPI_GEGLU && a.b_major != 0) {
    if (err) snprintf(err, err_len, "gemm: GEGLU epilogue needs a K-major [2N,K] weight");
    return 1;
  }
  // ---- few-tile, long-K path (the prefix pass of the decode: M = 968 rows against a 16384-deep down-projection): with
  // 128 x 128 tiles each CTA streams 8 MB of operands for 0.5 GFLOP and the GEMM is L2-bound; 256 x 256 pair tiles quarter
  // the operand traffic per flop but leave only 32 of them for 74 clusters.  Split K so that pairs * s fills one wave, fp32
  // partials + the general finish kernel (sum in a fixed order, then the STORE / RES epilogue).
  if (a.M > BM && a.batch == 1 && a.splitk_ws != nullptr && a.a_major == 0 && a.b_major == 0 && a.block_n == 0 &&
      (a.epilogue == EPI_STORE || a.epilogue == EPI_RES) && a.K % BK == 0 && a.K >= 4096 && a.norm_mod == nullptr &&
      a.rope == nullptr) {
    const long long pairs = ((a.M + 2 * BM - 1) / (2 * BM)) * ((a.N + 255) / 256);
    const long long slots = num_sms() / 2;
    const int kb = a.K / BK;
    int sp = 1;
    for (int c = 2; c <= 8; ++c)
      if (pairs * c <= slots && kb % c == 0 && a.K / c >= 2048 &&
          static_cast<size_t>(c) * a.M * a.N * sizeof(float) <= a.splitk_ws_bytes)
        sp = c;
    if (sp > 1 && pairs * 2 <= slots) {
      const int Kc = a.K / sp;
      GemmArgs part = a;
      part.splitk_ws = nullptr;
      part.K = Kc;
      part.batch = sp;
      part.batch_inner = 0;
      part.a_batch_stride = Kc;
      part.b_batch_stride = Kc;
      part.epilogue = EPI_F32;
      part.accumulate = 0;
      part.D = a.splitk_ws;
      part.ldd = a.N;
      part.d_batch_stride = static_cast<int64_t>(a.M) * a.N;
      part.block_n = 256;
      part.bias = nullptr;
      part.res = nullptr;
      part.gate = nullptr;
      part.D2 = nullptr;
      int rc = gemm_bf16(part, stream, err, err_len);
      if (rc != 0) return rc;
      const long long total = static_cast<long long>(a.M) * a.N;
      long long grid = (total + 255) / 256;
      if (grid > num_sms() * 16) grid = num_sms() * 16;
      const bool vec4 = a.N % 4 == 0 && a.ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(a.D) & 7) == 0;
      if (vec4 && a.epilogue == EPI_RES && a.bias == nullptr && a.gate == nullptr && a.D2 == nullptr && a.ldres % 4 == 0 &&
          (reinterpret_cast<uintptr_t>(a.res) & 7) == 0) {
        long long g4 = (total / 4 + 255) / 256;
        if (g4 > num_sms() * 16) g4 = num_sms() * 16;
        launch_pdl(splitk_sum_res4_k, dim3(static_cast<int>(g4)), dim3(256), 0, stream, a.splitk_ws, sp, total / 4, a.N / 4,
                   static_cast<__nv_bfloat16*>(a.D), a.ldd, static_cast<const __nv_bfloat16*>(a.res), a.ldres);
        count_launch();
        return 0;
      }
      if (vec4 && a.epilogue == EPI_STORE) {
        long long g4 = (total / 4 + 255) / 256;
        if (g4 > num_sms() * 16) g4 = num_sms() * 16;
        launch_pdl(splitk_sum_store4_k, dim3(static_cast<int>(g4)), dim3(256), 0, stream, a.splitk_ws, sp, total / 4, a.N / 4,
                   static_cast<__nv_bfloat16*>(a.D), a.ldd);
        count_launch();
        return 0;
      }
      launch_pdl(splitk_finish_k, dim3(static_cast<int>(grid)), dim3(256), 0, stream, a.splitk_ws, sp, a.M, a.N, a.epilogue,
                 static_cast<__nv_bfloat16*>(a.D), a.ldd, static_cast<__nv_bfloat16*>(a.D2), a.ldd2,
                 static_cast<const __nv_bfloat16*>(a.bias), static_cast<const __nv_bfloat16*>(a.res), a.ldres,
                 static_cast<const __nv_bfloat16*>(a.gate), a.gate_rows > 0 ? a.gate_rows : 1, a.ldgate);
      count_launch();
      return 0;
    }
  }
  int bn = a.block_n;
  if (bn == 0) {
    bn = (a.N > 128 || a.epilogue == EPI_GEGLU) ? 256 : 128;
    // Under-filled grids (decode / prefill at batch 1: M = 50 or 968): when 256-wide tiles cannot occupy the 148 SMs,
    // narrower tiles double the number of CTAs streaming the weights.
    const long long mt = (a.M + BM - 1) / BM;
    const int out256 = (a.epilogue == EPI_GEGLU) ? 128 : 256;
    const long long tiles256 = mt * ((a.N + out256 - 1) / out256) * a.batch;
    if (tiles256 < 120) bn = 128;
  }
  if (bn != 128 && bn != 256) {
    if (err) snprintf(err, err_len, "gemm: unsupported block_n %d", bn);
    return 1;
  }
  const int bn_out = (a.epilogue == EPI_GEGLU) ? bn / 2 : bn;
  const long long b_rows = (a.epilogue == EPI_GEGLU) ? 2LL * a.N : a.N;
  // 2-CTA (cta_group::2) kernel for every problem with at least two 128-row blocks and 256-wide tiles: each CTA of
  // the pair stages half of the B tile, so B is always fetched in 128-row boxes.  PI05_GEMM_1CTA=1 forces 1-CTA.
  static const bool force_1cta = getenv("PI05_GEMM_1CTA") != nullptr;
  const bool use2 = (bn == 256) && (a.M > BM) && !force_1cta;
  const int b_box_rows = (a.epilogue == EPI_GEGLU || use2) ? bn / 2 : bn;

  // ---- wave-quantisation path: weight-gradient shapes (few output tiles, very long K) leave most of the last wave
  // idle (e.g. 85 pair-tiles on 74 clusters = 57 %).  Split K into s equal chunks run as a batch with fp32 partials
  // and sum them in a fixed order.  Cost model in units of the un-split, perfectly balanced GEMM time:
  // t(s) = 1 / wave_efficiency(s) + s * 933 / K   (fp32 partial write+read at ~6 TB/s against ~1.4 PFLOP/s).
  if (a.epilogue == EPI_STORE && a.batch == 1 && a.splitk_ws != nullptr && a.K % BK == 0 && a.K >= 8192 &&
      a.block_n == 0 && a.D2 == nullptr) {
    const long long mt = (a.M + BM - 1) / BM, nt = (a.N + bn - 1) / bn;
    const long long tiles = use2 ? ((mt + 1) / 2) * nt : mt * nt;
    const long long slots = use2 ? num_sms() / 2 : num_sms();
    const int kb = a.K / BK;
    auto t_of = [&](int sp) {
      const long long work = tiles * sp;
      const long long waves = (work + slots - 1) / slots;
      return static_cast<double>(waves * slots) / static_cast<double>(work) + sp * 933.0 / a.K;
    };
    int best = 1;
    double tbest = t_of(1);
    for (int sp = 2; sp <= 8; ++sp) {
      if (kb % sp != 0 || a.K / sp < 2048) continue;
      if (static_cast<size_t>(sp) * a.M * a.N * sizeof(float) > a.splitk_ws_bytes) continue;
      const double t = t_of(sp);
      if (t < tbest) {
        tbest = t;
        best = sp;
      }
    }
    if (best > 1 && tbest < 0.9 * t_of(1)) {
      const int Kc = a.K / best;
      GemmArgs part = a;
      part.splitk_ws = nullptr;
      part.K = Kc;
      part.batch = best;
      part.batch_inner = 0;
      part.a_batch_stride = a.a_major == 0 ? Kc : static_cast<int64_t>(Kc) * a.lda;
      part.b_batch_stride = a.b_major == 0 ? Kc : static_cast<int64_t>(Kc) * a.ldb;
      part.epilogue = EPI_F32;
      part.accumulate = 0;
      part.D = a.splitk_ws;
      part.ldd = a.N;
      part.d_batch_stride = static_cast<int64_t>(a.M) * a.N;
      part.block_n = bn;
      part.norm_mod = nullptr;
      part.rope = nullptr;
      int rc = gemm_bf16(part, stream, err, err_len);
      if (rc != 0) return rc;
      const long long total = static_cast<long long>(a.M) * a.N;
      long long grid = (total / 4 + 255) / 256;
      if (grid > num_sms() * 16) grid = num_sms() * 16;
      if (a.N % 4 == 0 && a.ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(a.D) & 7) == 0) {
        launch_pdl(splitk_sum_store4_k, dim3(static_cast<int>(grid)), dim3(256), 0, stream, a.splitk_ws, best, total / 4, a.N / 4,
                                                                        static_cast<__nv_bfloat16*>(a.D), a.ldd);
      } else {
        launch_pdl(splitk_finish_k, dim3(static_cast<int>(grid)), dim3(256), 0, stream,
            a.splitk_ws, best, a.M, a.N, EPI_STORE, static_cast<__nv_bfloat16*>(a.D), a.ldd, nullptr, 0, nullptr, nullptr,
            0, nullptr, 1, 0);
      }
      count_launch();
      return 0;
    }
  }

  const int nz0 = (a.batch_inner > 0) ? a.batch_inner : a.batch;
  if (a.batch % nz0 != 0) {
    if (err) snprintf(err, err_len, "gemm: batch %d not a multiple of batch_inner %d", a.batch, nz0);
    return 1;
  }
  const int nz1 = a.batch / nz0;
  CUtensorMap ta, tb;
  if (!make_tmap(&ta, a.A, a.a_major, a.M, a.K, a.lda, nz0, a.a_batch_stride, nz1, a.a_batch_stride1, BM, err,
                 err_len))
    return 4;
  if (!make_tmap(&tb, a.B, a.b_major, b_rows, a.K, a.ldb, nz0, a.b_batch_stride, nz1, a.b_batch_stride1, b_box_rows,
                 err, err_len))
    return 4;

  KParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.M = a.M;
  kp.N = a.N;
  kp.K = a.K;
  kp.batch = a.batch;
  kp.a_mn = a.a_major;
  kp.b_mn = a.b_major;
  kp.nz0 = nz0;
  kp.a_b0 = (nz0 > 1 && a.a_batch_stride != 0) ? 1 : 0;
  kp.a_b1 = (nz1 > 1 && a.a_batch_stride1 != 0) ? 1 : 0;
  kp.b_b0 = (nz0 > 1 && a.b_batch_stride != 0) ? 1 : 0;
  kp.b_b1 = (nz1 > 1 && a.b_batch_stride1 != 0) ? 1 : 0;
  kp.num_m = (a.M + BM - 1) / BM;
  kp.num_n = (a.N + bn_out - 1) / bn_out;
  kp.num_kb = (a.K + BK - 1) / BK;
  kp.idesc = make_idesc_bf16(use2 ? 2 * BM : BM, bn, a.a_major, a.b_major);
  kp.D = a.D;
  kp.ldd = a.ldd;
  kp.dbs = a.d_batch_stride;
  kp.dbs1 = a.d_batch_stride1;
  kp.D2 = a.D2;
  kp.ldd2 = a.ldd2;
  kp.d2bs = a.d2_batch_stride;
  kp.d2bs1 = a.d2_batch_stride1;
  kp.bias = static_cast<const __nv_bfloat16*>(a.bias);
  kp.res = static_cast<const __nv_bfloat16*>(a.res);
  kp.ldres = a.ldres;
  kp.resbs = a.res_batch_stride;
  kp.resbs1 = a.res_batch_stride1;
  kp.gate = static_cast<const __nv_bfloat16*>(a.gate);
  kp.gate_rows = a.gate_rows > 0 ? a.gate_rows : 1;
  kp.ldgate = a.ldgate;
  kp.scale = a.scale;
  kp.accumulate = a.accumulate;
  kp.bias32 = a.bias32;
  kp.rowadd32 = a.rowadd32;
  kp.rowadd_period = a.rowadd_period > 0 ? a.rowadd_period : 1;
  kp.ld_rowadd = a.ld_rowadd;
  if (a.epilogue == EPI_PATCH && (a.bias32 == nullptr || a.rowadd32 == nullptr)) {
    if (err) snprintf(err, err_len, "gemm: EPI_PATCH needs bias32 and rowadd32");
    return 1;
  }
  kp.mn_lbo = BK * 128;
  kp.mn_sbo = 1024;
  // Dynamic tile scheduling pays one atomic round trip per kernel (~0.3 us on a 3-10 us decode GEMM: measured +0.5 ms on
  // the 10-step decode), and buys nothing when no CTA / cluster gets more than one tile: those launches walk statically.
  static const int group_m_env = getenv("PI05_GROUP_M") ? atoi(getenv("PI05_GROUP_M")) : 0;
  kp.group_m = group_m_env > 0 ? group_m_env : GROUP_M;
  static const bool static_sched = getenv("PI05_GEMM_STATIC") != nullptr;
  {
    const long long tiles = use2 ? static_cast<long long>((kp.num_m + 1) / 2) * kp.num_n * kp.batch
                                 : static_cast<long long>(kp.num_m) * kp.num_n * kp.batch;
    const long long slots = use2 ? num_sms() / 2 : num_sms();
    kp.static_sched = (static_sched || tiles <= slots) ? 1 : 0;
  }
  if (const char* e = getenv("PI05_DBG_MN_LBO")) kp.mn_lbo = static_cast<uint32_t>(atoi(e));
  if (const char* e = getenv("PI05_DBG_MN_SBO")) kp.mn_sbo = static_cast<uint32_t>(atoi(e));
  ProfEntry pe{};
  const bool prof = g_prof_on;
  if (prof) {
    cudaEventCreate(&pe.a);
    cudaEventCreate(&pe.b);
    pe.M = a.M;
    pe.N = (a.epilogue == EPI_GEGLU) ? 2 * a.N : a.N;
    pe.K = a.K;
    pe.batch = a.batch;
    pe.epi = a.epilogue;
    pe.majors = a.a_major * 2 + a.b_major;
    cudaEventRecord(pe.a, stream);
  }
  const int rc = use2 ? launch_gemm2(a.epilogue, ta, tb, kp, stream, err, err_len)
                 : (bn == 256) ? dispatch_epi<256>(a.epilogue, ta, tb, kp, stream, err, err_len)
                             : dispatch_epi<128>(a.epilogue, ta, tb, kp, stream, err, err_len);
  if (prof) {
    cudaEventRecord(pe.b, stream);
    g_prof.push_back(pe);
  }
  return rc;
}

// ---- per-launch timing of the tcgen05 GEMM (bench.py's roofline line): CUDA events around every launch ----------
void gemm_profile_enable(int on) {
  g_prof_on = on != 0;
  if (on) {
    for (auto& p : g_prof) {
      cudaEventDestroy(p.a);
      cudaEventDestroy(p.b);
    }
    g_prof.clear();
  }
}

// Synchronises the events and writes one line per GEMM class: "M N K batch epi majors launches total_ms\n".
int gemm_profile_report(char* buf, int len) {
  struct Acc {
    int launches = 0;
    double ms = 0;
  };
  std::map<std::tuple<int, int, int, int, int, int>, Acc> acc;
  for (auto& p : g_prof) {
    cudaEventSynchronize(p.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    Acc& x = acc[std::make_tuple(p.M, p.N, p.K, p.batch, p.epi, p.majors)];
    x.launches += 1;
    x.ms += ms;
  }
  int off = 0;
  for (auto& kv : acc) {
    const int n = snprintf(buf + off, off < len ? len - off : 0, "%d %d %d %d %d %d %d %.6f\n", std::get<0>(kv.first),
                           std::get<1>(kv.first), std::get<2>(kv.first), std::get<3>(kv.first), std::get<4>(kv.first),
                           std::get<5>(kv.first), kv.second.launches, kv.second.ms);
    if (n < 0 || off + n >= len) break;
    off += n;
  }
  return off;
}

}  // namespace pi05
