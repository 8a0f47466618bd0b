// Internal interface of the tcgen05 bf16 GEMM engine (see gemm_sm100.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pi05 {

// Epilogue applied while draining the fp32 accumulator out of TMEM.  Every "bf(...)" below is a
// round-to-nearest-even bf16 rounding, placed where the reference materialises a bf16 tensor.
enum GemmEpilogue : int {
  EPI_STORE = 0,      // D = bf(acc)
  EPI_SCALE = 1,      // D = bf(bf(acc) * scale)                      (attention scores, modeling_gemma.py:243)
  EPI_BIAS = 2,       // D = bf(acc + bias[n])                        (nn.Linear with bias)
  EPI_BIAS_GELU = 3,  // D = bf(acc + bias[n]); D2 = bf(gelu_tanh(D)) (SiglipMLP fc1 + activation)
  EPI_RES = 4,        // t = bf(acc [+ bias[n]]); if gate: t = bf(t * gate[row / gate_rows, n]); D = bf(res + t)
  EPI_GEGLU = 5,      // B tile = [gate rows | up rows]; g = bf(acc_g), u = bf(acc_u), D[:, n] = g, D[:, N + n] = u,
                      // D2[:, n] = bf(bf(gelu_tanh(g)) * u)          (GemmaMLP, modeling_gemma.py:125)
  EPI_F32 = 6,        // Df32 = acc (+ Df32 if accumulate)
  // Backward of the GeGLU fused into the down-projection dgrad: acc = dH tile; res = GU = [g | u] (ld 2N) from the
  // forward; D = dGU = [dg | du] (ld 2N):  a = bf(gelu(g)); du = bf(bf(acc)*a); dg = bf(bf(bf(acc)*u) * gelu'(g))
  EPI_GEGLU_BWD = 7,
  // Backward of gelu_tanh fused into the fc2 dgrad (SigLIP): acc = d(act); res = pre-activation; D = bf(bf(acc)*gelu'(pre))
  EPI_GELU_BWD = 8,
  // SigLIP patch embedding (modeling_siglip.py:271-282): D = bf( (acc + bias32[n]) + rowadd32[row % rowadd_period, n] ),
  // fp32 bias and fp32 position embedding added in the reference's order before the single rounding to bf16
  EPI_PATCH = 9,
  EPI_COUNT = 10,
};

// Optional (EPI_STORE, decode): the RoPE + q/k/v split that consumes a fused qkv projection (rope_pack_fwd in kernels.h,
// same arguments) fused behind the split-K finish: Q / K-cache / V-cache rows are written directly, D is not.
struct GemmRope {
  int T, H, hd;
  const int* pos;
  const int* nvalid;
  int pos_mode;
  const void* cos_t;
  const void* sin_t;
  void *Q, *K, *V;
  int key_off, kv_len, batch;
};

struct GemmArgs {
  // Problem: for every batch index z: D[z][M,N] = A[z][M,K] * B[z][N,K]^T
  int M = 0, N = 0, K = 0, batch = 1;
  // Two-level batch: z = z1 * batch_inner + z0 (batch_inner = 0 means "= batch", i.e. one level).  Every operand
  // has a stride per level (elements); a 0 stride means the operand does not move along that level.
  int batch_inner = 0;
  // Operands (bf16).  major 0: row-major [rows, K] (K contiguous).  major 1: row-major [K, rows] (rows contiguous).
  const void* A = nullptr;
  const void* B = nullptr;
  int a_major = 0, b_major = 0;
  int64_t lda = 0, ldb = 0;                      // elements between consecutive rows of the stored matrix
  int64_t a_batch_stride = 0, b_batch_stride = 0;  // level-0 stride (elements); 0 = shared by all batches
  int64_t a_batch_stride1 = 0, b_batch_stride1 = 0;  // level-1 stride
  // Output
  int epilogue = EPI_STORE;
  void* D = nullptr;
  int64_t ldd = 0, d_batch_stride = 0, d_batch_stride1 = 0;
  void* D2 = nullptr;  // EPI_BIAS_GELU / EPI_GEGLU second output; EPI_RES: optional copy of the pre-gate value t
  int64_t ldd2 = 0, d2_batch_stride = 0, d2_batch_stride1 = 0;
  const void* bias = nullptr;  // bf16 [N]
  const void* res = nullptr;   // bf16 [M, ldres]
  int64_t ldres = 0, res_batch_stride = 0, res_batch_stride1 = 0;
  const void* gate = nullptr;  // bf16 [ceil(M / gate_rows), ldgate]
  int gate_rows = 1;
  int64_t ldgate = 0;
  float scale = 1.0f;
  int accumulate = 0;  // EPI_F32 only
  // Optional fp32 scratch for the small-M split-K path (decode: M = action horizon, one 128-row tile, long K):
  // the K loop is split across CTAs as a batch dimension and a finish kernel applies the epilogue.
  float* splitk_ws = nullptr;
  size_t splitk_ws_bytes = 0;
  int block_n = 0;     // 0 = auto (256 when N >= 256 else 128)
  // Optional (EPI_RES, decode): the adaptive RMSNorm that consumes D (modeling_gemma.py:84-104) fused behind the
  // epilogue.  norm_mod: fp32 [*, 3*N] = [scale | shift | gate] rows, one per norm_rows_per_batch output rows;
  // norm_out = bf16 [M, N]; norm_gate_out (optional) = bf16 gate per batch row.  Same arithmetic as rmsnorm_fwd.
  const float* norm_mod = nullptr;
  int norm_rows_per_batch = 0;
  void* norm_out = nullptr;
  void* norm_gate_out = nullptr;
  const GemmRope* rope = nullptr;
  // EPI_PATCH: fp32 bias [N] and fp32 per-row addend table [rowadd_period, ld_rowadd]
  const float* bias32 = nullptr;
  const float* rowadd32 = nullptr;
  int rowadd_period = 1;
  int64_t ld_rowadd = 0;
};

// Enqueue on `stream`.  Returns 0 on success; on failure returns non-zero and fills `err` (if given).
int gemm_bf16(const GemmArgs& args, cudaStream_t stream, char* err = nullptr, int err_len = 0);

}  // namespace pi05

namespace pi05 {
// Per-launch CUDA-event timing of the tcgen05 GEMM (used by bench.py for the roofline line).
void gemm_profile_enable(int on);
int gemm_profile_report(char* buf, int len);
}  // namespace pi05
