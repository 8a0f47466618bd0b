// 2-CTA (cta_group::2) variant of the tcgen05 bf16 GEMM: a cluster of two CTAs on one TPC computes a 256 x 256
// output tile.  Each CTA stages its own 128 x 64 A slice and HALF of the 256 x 64 B slice; the leader CTA's single
// MMA thread issues UMMA 256x256x16 that reads A and B from both CTAs' shared memory and writes 128 accumulator rows
// into each CTA's TMEM.  Per-CTA shared-memory fill traffic per flop is 2/3 of the 1-CTA kernel's (32 KB instead of
// 48 KB per 128x256x64 MMA block) and the ring is 6 stages deep instead of 4.
//
// Synchronisation (all mbarriers live at identical offsets in both CTAs):
//   full[s]   (leader's copy is the one waited on, count 2): leader producer arrive.expect_tx(both CTAs' bytes) +
//             peer producer remote arrive; both CTAs' TMA loads complete_tx on the LEADER's barrier (.cta_group::2).
//   empty[s]  (count 1, per CTA): tcgen05.commit ... multicast::cluster to both CTAs frees the slot for both producers.
//   tfull[a]  (count 1, per CTA): multicast commit after the last k-block -> each CTA's epilogue warps.
//   tempty[a] (leader's copy, count 2 x 256): every epilogue thread of both CTAs arrives (peer: remote arrive).
//
// Tile schedule: DYNAMIC.  The leader's producer thread draws pair-tile indices from a global counter (atomicAdd) and
// publishes them through a 4-slot ring in BOTH CTAs' shared memory (tq_full[slot]: count 1 per CTA, remote store +
// an asynchronous remote store that completes on the peer's barrier; tq_empty[slot]: leader's copy, one arrive per consumer warp of both CTAs).
// A cluster that becomes resident late (SMs held by a concurrent kernel, e.g. the NCCL all-reduce of the gradient
// exchange that overlaps backward) simply draws fewer tiles, where a static `tile += gridDim` walk would serialise its
// whole share behind the others.  Which cluster computes a tile never changes the tile's arithmetic: results stay
// bit-identical run to run.  The counter is returned to zero by the last cluster to finish (see gemm_host.h).
#include <cstdio>

#include "errors.h"
#include "gemm_host.h"
#include "launch.h"

namespace pi05 {

using namespace gemm_detail;

namespace {

constexpr int BN2 = 256;
constexpr int STAGES2 = 6;
constexpr int B_HALF_BYTES = (BN2 / 2) * BK * 2;              // 16 KiB: this CTA's half of the B tile
constexpr int STAGE_BYTES2 = A_STAGE_BYTES + B_HALF_BYTES;    // 32 KiB
constexpr int TILE_BYTES2 = STAGES2 * STAGE_BYTES2;           // 192 KiB
constexpr int SMEM_BYTES2 = TILE_BYTES2 + 1024 + 256 + NUM_EPI_WARPS * STAGE_BYTES_PER_WARP;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;                // clears the CTA-rank bit of a shared::cluster address
constexpr int TQ = 4;                                         // tile-index ring slots
constexpr int TQ_CONSUMER_WARPS = 2 * (1 + NUM_EPI_WARPS);    // leader: MMA warp + epilogue warps; peer: producer + epilogue warps

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier.
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// Arrive on the barrier at the same offset in CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}

// ---- tile ring, remote half: the index reaches the peer CTA as an ASYNCHRONOUS remote store that completes on the peer's
// mbarrier (st.async ... mbarrier::complete_tx, the same completion mechanism TMA uses), so no cluster-scope release /
// acquire is needed: those compile to MEMBAR.ALL.GPU + CCTL.IVALL (L1 invalidate) and cost ~1.5 us per tile boundary
// on the MMA warp's critical path (measured: K = 2048 GEMM classes 10 % slower).
__device__ __forceinline__ void remote_expect_tx(uint32_t bar, uint32_t cta, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [ra], %2;\n\t}"
      ::"r"(bar), "r"(cta), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void st_async_remote_u32(uint32_t addr, uint32_t bar, uint32_t cta, uint32_t v) {
  asm volatile(
      "{\n\t.reg .b32 ra, rb;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %2;\n\t"
      "mapa.shared::cluster.u32 rb, %1, %2;\n\t"
      "st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [ra], %3, [rb];\n\t}"
      ::"r"(addr), "r"(bar), "r"(cta), "r"(v)
      : "memory");
}
__device__ __forceinline__ int ld_shared_s32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// One consumer warp's view of the tile ring: wait for the next published index, read it, release the slot.
struct TileRing {
  uint32_t full0, empty0, slot0;
  int slot;
  uint32_t phase;
  bool leader;
  __device__ __forceinline__ int next() {  // all 32 lanes of a converged warp, or the single producer thread
    mbar_wait(full0 + 8 * slot, phase);
    const int t = ld_shared_s32(slot0 + 4 * slot);
    if (t < -1) __trap();  // never true: makes the slot release below wait for the load (write-after-read on the slot)
    return t;
  }
  __device__ __forceinline__ void release_warp() {  // after every lane has consumed the value
    __syncwarp();
    if ((threadIdx.x & 31) == 0) release_one();
    advance();
  }
  __device__ __forceinline__ void release_one() {
    if (leader)
      mbar_arrive(empty0 + 8 * slot);
    else
      mbar_arrive_cluster(empty0 + 8 * slot, 0);
  }
  __device__ __forceinline__ void advance() {
    if (++slot == TQ) {
      slot = 0;
      phase ^= 1;
    }
  }
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const KParams p,
             unsigned int* __restrict__ sched) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + STAGES2 * A_STAGE_BYTES;
  const uint32_t bar_base = smem_base + TILE_BYTES2;
  const uint32_t full_bar0 = bar_base;
  const uint32_t empty_bar0 = bar_base + 8 * STAGES2;
  const uint32_t tfull_bar0 = bar_base + 16 * STAGES2;
  const uint32_t tempty_bar0 = tfull_bar0 + 16;
  const uint32_t tmem_slot = tempty_bar0 + 16;
  // tile-index ring (after the 12 + 4 pipeline barriers and the TMEM slot; the barrier area is 256 bytes)
  const uint32_t tq_full0 = tmem_slot + 8;
  const uint32_t tq_empty0 = tq_full0 + 8 * TQ;
  const uint32_t tq_slot0 = tq_empty0 + 8 * TQ;
  static_assert(16 * STAGES2 + 32 + 8 + 16 * TQ + 4 * TQ <= 256, "barrier area overflow");

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // First tile index: drawn as the very first instruction, consumed after the prologue, so the atomic's round trip is off
  // the critical path even in a 3 us kernel.  (Safe before griddepcontrol.wait: the only other writer of this counter pair
  // is the reset of a launch 64 GEMMs ago; a PDL predecessor uses another pair.)
  unsigned int first_raw = 0;
  const bool dyn = p.static_sched == 0;
  if (threadIdx.x == 0 && leader && dyn) first_raw = atomicAdd(sched, 1u);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(full_bar0 + 8 * s, 2);
      mbar_init(empty_bar0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar0 + 8 * a, 1);
      mbar_init(tempty_bar0 + 8 * a, 2 * 32 * NUM_EPI_WARPS);
    }
    for (int q = 0; q < TQ; ++q) {
      mbar_init(tq_full0 + 8 * q, 1);
      mbar_init(tq_empty0 + 8 * q, TQ_CONSUMER_WARPS);
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc2(tmem_slot, 2 * BN2);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits + TMEM allocation visible in both CTAs before any remote arrive / multicast
  tc_fence_after();
  pdl_enter();  // the prologue above overlaps the previous kernel's tail (launch.h)
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // the pair walks "pair tiles": two vertically adjacent 128-row blocks x one 256-column block
  KParams pp = p;
  pp.num_m = (p.num_m + 1) / 2;
  const int total_tiles = pp.num_m * pp.num_n * pp.batch;
  const int num_clusters = gridDim.x >> 1;
  constexpr int BN_OUT = (EPI == EPI_GEGLU) ? BN2 / 2 : BN2;
  TileRing ring{tq_full0, tq_empty0, tq_slot0, 0, 0u, leader};

  if (warp_idx == 0) {
    // ============================== TMA producer (both CTAs) ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // leader: draw + publish tile indices (the draw for tile n+1 is in flight while tile n's loads are issued);
      // peer: follow the ring
      int pub_slot = 0;
      uint32_t pub_phase = 0;
      auto to_tile = [&](unsigned int raw) -> int {
        return raw < static_cast<unsigned int>(total_tiles) ? static_cast<int>(raw) : -1;
      };
      auto publish = [&](int t) {
        mbar_wait(tq_empty0 + 8 * pub_slot, pub_phase ^ 1);  // every consumer warp of both CTAs is done with the old value
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(tq_slot0 + 4 * pub_slot), "r"(t) : "memory");
        mbar_arrive(tq_full0 + 8 * pub_slot);
        remote_expect_tx(tq_full0 + 8 * pub_slot, 1, 4);  // the peer's copy: 1 arrival + 4 bytes complete the phase
        st_async_remote_u32(tq_slot0 + 4 * pub_slot, tq_full0 + 8 * pub_slot, 1, static_cast<uint32_t>(t));
        if (++pub_slot == TQ) {
          pub_slot = 0;
          pub_phase ^= 1;
        }
      };
      int tile, next_tile = -1;
      // the next index is published a few k-blocks into the current tile: late enough for the draw to have returned,
      // early enough that the peer's producer and the consumers never wait for it at the tile boundary
      const int pub_kb = p.num_kb > 8 ? 8 : p.num_kb - 1;
      const int cluster_id = blockIdx.x >> 1;
      if (!dyn) {
        tile = cluster_id < total_tiles ? cluster_id : -1;
      } else if (leader) {
        tile = to_tile(first_raw);
        publish(tile);
      } else {
        tile = ring.next();
        ring.release_one();
        ring.advance();
      }
      while (tile >= 0) {
        unsigned int next_raw = 0;
        if (leader && dyn) next_raw = atomicAdd(sched, 1u);  // result first read at k-block pub_kb
        const TileCoord tc = decode_tile(tile, pp);
        const int m0 = (2 * tc.m_blk + static_cast<int>(rank)) * BM;
        const int n0 = tc.n_blk * BN_OUT;
        const int za0 = p.a_b0 ? tc.z0 : 0, za1 = p.a_b1 ? tc.z1 : 0;
        const int zb0 = p.b_b0 ? tc.z0 : 0, zb1 = p.b_b1 ? tc.z1 : 0;
        // this CTA's half of the B tile (rows of the [N, K] operand)
        const int nb = (EPI == EPI_GEGLU) ? (leader ? n0 : p.N + n0) : n0 + static_cast<int>(rank) * (BN2 / 2);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar0 + 8 * stage, phase ^ 1);
          const uint32_t full = full_bar0 + 8 * stage;
          if (leader) mbar_arrive_expect_tx(full, 2 * STAGE_BYTES2);
          const uint32_t sa = smem_a0 + stage * A_STAGE_BYTES;
          const uint32_t sb = smem_b0 + stage * B_HALF_BYTES;
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_4d_2sm(sa, &tma_a, full, k0, m0, za0, za1);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_4d_2sm(sa + i * (BK * 128), &tma_a, full, m0 + 64 * i, k0, za0, za1);
          }
          if (!p.b_mn) {
            tma_load_4d_2sm(sb, &tma_b, full, k0, nb, zb0, zb1);
          } else {
#pragma unroll
            for (int i = 0; i < (BN2 / 2) / 64; ++i)
              tma_load_4d_2sm(sb + i * (BK * 128), &tma_b, full, nb + 64 * i, k0, zb0, zb1);
          }
          if (!leader) mbar_arrive_cluster(full, 0);  // "my loads for this stage are in flight"
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
          if (leader && dyn && kb == pub_kb) {
            next_tile = to_tile(next_raw);
            publish(next_tile);
          }
        }
        if (!dyn) {
          tile = tile + num_clusters < total_tiles ? tile + num_clusters : -1;
        } else if (leader) {
          tile = next_tile;
        } else {
          tile = ring.next();
          ring.release_one();
          ring.advance();
        }
      }
      if (leader && dyn) {
        // the last cluster to run dry returns the counter (and the exit count) to zero for the next launch that uses it
        if (atomicAdd(sched + 1, 1u) == static_cast<unsigned int>(num_clusters) - 1u) {
          sched[0] = 0u;
          sched[1] = 0u;
          __threadfence();
        }
      }
    }
  } else if (warp_idx == 1) {
    // ============================== MMA issuer (leader CTA only) ==============================
    // The WHOLE warp walks the loop with warp-uniform values (so descriptors live in uniform registers) and one
    // elected lane issues; a lane-0-only loop costs ~170 SASS instructions per k-block (ELECT/R2UR per operand) and
    // makes the single issuing thread, not the tensor pipe, the bottleneck (ncu: 76% tensor-active before this).
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t a_lbo = p.a_mn ? p.mn_lbo : 16, a_sbo = p.a_mn ? p.mn_sbo : 1024, a_kstep = p.a_mn ? 2048 : 32;
      const uint32_t b_lbo = p.b_mn ? p.mn_lbo : 16, b_sbo = p.b_mn ? p.mn_sbo : 1024, b_kstep = p.b_mn ? 2048 : 32;
      const uint32_t a_hi = smem_desc_hi_sw128(a_sbo), b_hi = smem_desc_hi_sw128(b_sbo);
      const uint32_t a_step = a_kstep >> 4, b_step = b_kstep >> 4;
      const uint32_t idesc = p.idesc;
      const int num_kb = p.num_kb;
      int stile = blockIdx.x >> 1;
      for (;;) {
        int tile;
        if (dyn) {
          tile = ring.next();
          ring.release_warp();
        } else {
          tile = stile < total_tiles ? stile : -1;
          stile += num_clusters;
        }
        if (tile < 0) break;
        mbar_wait(tempty_bar0 + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN2;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_lo = smem_desc_lo(smem_a0 + stage * A_STAGE_BYTES, a_lbo);
          const uint32_t b_lo = smem_desc_lo(smem_b0 + stage * B_HALF_BYTES, b_lbo);
          if (elect_one()) {
            umma2_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, kb > 0 ? 1u : 0u);
#pragma unroll
            for (int j = 1; j < BK / 16; ++j)
              umma2_lohi(d_tmem, a_lo + j * a_step, a_hi, b_lo + j * b_step, b_hi, idesc, 1u);
            umma_commit2(empty_bar0 + 8 * stage);
            if (kb == num_kb - 1) umma_commit2(tfull_bar0 + 8 * acc);
          }
          __syncwarp();
          if (++stage == STAGES2) {
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
    // ============================== epilogue (both CTAs, own 128 rows) ==============================
    const int q = warp_idx & 3;
    const int chalf = (warp_idx - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    int stile = blockIdx.x >> 1;
    for (;;) {
      int tile;
      if (dyn) {
        tile = ring.next();
        ring.release_warp();
      } else {
        tile = stile < total_tiles ? stile : -1;
        stile += num_clusters;
      }
      if (tile < 0) break;
      const TileCoord tc = decode_tile(tile, pp);
      const int row0 = (2 * tc.m_blk + static_cast<int>(rank)) * BM + q * 32;
      const int n0 = tc.n_blk * BN_OUT;
      mbar_wait(tfull_bar0 + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN2;
      epilogue_tile<BN2, EPI>(p, tc.z0, tc.z1, row0, lane, n0, t_base, chalf,
                              bar_base + 256 + (warp_idx - 2) * STAGE_BYTES_PER_WARP);
      tc_fence_before();
      if (leader)
        mbar_arrive(tempty_bar0 + 8 * acc);
      else
        mbar_arrive_cluster(tempty_bar0 + 8 * acc, 0);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // nobody frees TMEM / exits while the peer may still multicast into this CTA
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 2 * BN2);
  }
}

template <int EPI>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const KParams& kp, cudaStream_t stream, char* err, int err_len) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES2);
    if (e != cudaSuccess) {
      if (err) snprintf(err, err_len, "cudaFuncSetAttribute(gemm2): %s", cudaGetErrorString(e));
      return 2;
    }
    configured = true;
  }
  static int sms = [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  const int pairs = ((kp.num_m + 1) / 2) * kp.num_n * kp.batch;
  int clusters = sms / 2;
  if (pairs < clusters) clusters = pairs;
  unsigned int* sched = next_sched_counter();
  if (sched == nullptr) {
    if (err) snprintf(err, err_len, "gemm2: tile-scheduler counters unavailable");
    return 2;
  }
  launch_pdl(gemm2_kernel<EPI>, dim3(2 * clusters), dim3(NUM_THREADS), SMEM_BYTES2, stream, ta, tb, kp, sched);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) snprintf(err, err_len, "gemm2 launch: %s", cudaGetErrorString(e));
    return 3;
  }
  return 0;
}

}  // namespace

int launch_gemm2(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const KParams& kp, cudaStream_t s, char* err,
                 int err_len) {
  switch (epi) {
    case EPI_STORE: return launch2<EPI_STORE>(ta, tb, kp, s, err, err_len);
    case EPI_SCALE: return launch2<EPI_SCALE>(ta, tb, kp, s, err, err_len);
    case EPI_BIAS: return launch2<EPI_BIAS>(ta, tb, kp, s, err, err_len);
    case EPI_BIAS_GELU: return launch2<EPI_BIAS_GELU>(ta, tb, kp, s, err, err_len);
    case EPI_RES: return launch2<EPI_RES>(ta, tb, kp, s, err, err_len);
    case EPI_GEGLU: return launch2<EPI_GEGLU>(ta, tb, kp, s, err, err_len);
    case EPI_F32: return launch2<EPI_F32>(ta, tb, kp, s, err, err_len);
    case EPI_GEGLU_BWD: return launch2<EPI_GEGLU_BWD>(ta, tb, kp, s, err, err_len);
    case EPI_GELU_BWD: return launch2<EPI_GELU_BWD>(ta, tb, kp, s, err, err_len);
    case EPI_PATCH: return launch2<EPI_PATCH>(ta, tb, kp, s, err, err_len);
    default:
      if (err) snprintf(err, err_len, "unknown epilogue %d", epi);
      return 1;
  }
}

}  // namespace pi05
