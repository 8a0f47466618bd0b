// Launchers of the non-GEMM kernels of the pi0.5 path (norms, RoPE, softmax, embeddings, element-wise tails,
// fp32 SIMT linears).  All enqueue on `st`; none synchronises.  bf16 = __nv_bfloat16.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pi05 {

typedef __nv_bfloat16 bf16;

// ---------------- norms (norm_kernels.cu) ----------------
// nn.LayerNorm on bf16 rows, fp32 statistics (modeling_siglip.py:466,474,787).  mean/rstd may be null (inference).
void layernorm_fwd(const bf16* x, const bf16* w, const bf16* b, bf16* y, float* mean, float* rstd, int rows, int width,
                   float eps, cudaStream_t st);
// dx = LN backward; if dres != null, dx += dres (residual-stream gradient).  dw32/db32: fp32 accumulators (+=).
void layernorm_bwd(const bf16* dy, const bf16* x, const bf16* w, const float* mean, const float* rstd,
                   const bf16* dres, bf16* dx, float* dw32, float* db32, int rows, int width, cudaStream_t st);
// GemmaRMSNorm (modeling_gemma.py:49-104).  Plain: y = bf(x*rstd*(1+w)).  Adaptive (mod != null): per batch b,
// mod[b] = [scale | shift | gate] fp32 (3*width): y = bf(x*rstd*(1+scale)+shift); gate_out[b] = bf(gate).
void rmsnorm_fwd(const bf16* x, const float* w, const float* mod, int rows_per_batch, bf16* y, float* rstd,
                 bf16* gate_out, int rows, int width, float eps, cudaStream_t st);
// Backward.  Plain: dw32 += sum_r dy*xhat.  Adaptive: dmod[b] scale/shift sections += (gate section untouched).
void rmsnorm_bwd(const bf16* dy, const bf16* x, const float* w, const float* mod, int rows_per_batch,
                 const float* rstd, const bf16* dres, bf16* dx, float* dw32, float* dmod, int rows, int width,
                 cudaStream_t st);
// One-pass variants (norm_bwd_fused.cu): dx and dw (db) from a single read of dy and x; false = shape not supported.
bool layernorm_bwd_fused(const bf16* dy, const bf16* x, const bf16* w, const float* mean, const float* rstd,
                         const bf16* dres, bf16* dx, float* dw32, float* db32, int rows, int width, cudaStream_t st);
bool rmsnorm_bwd_fused(const bf16* dy, const bf16* x, const float* w, const float* rstd, const bf16* dres, bf16* dx,
                       float* dw32, int rows, int width, cudaStream_t st);
// y = x + o*gate backward pieces for the adaptive stream: d_o = bf(dy*gate[b]); dmod[b].gate += sum_rows dy*o.
void gated_residual_bwd(const bf16* dy, const bf16* o, const bf16* gate, int rows_per_batch, bf16* d_o, float* dmod,
                        int rows, int width, cudaStream_t st);

// ---------------- attention helpers (attn_kernels.cu) ----------------
// Prefix bookkeeping (pi0_pytorch.py:207,221,343): pad[b,j], pos[b,j] = cumsum(pad)-1, nvalid[b].
void prefix_meta(const uint8_t* image_masks, const uint8_t* token_mask, int batch, int num_images, int tokens_per_image,
                 int max_token_len, uint8_t* pad, int* pos, int* nvalid, cudaStream_t st);
// RoPE (modeling_gemma.py:170-194, bf16 arithmetic) + layout split of a fused QKV row [H*hd | hd | hd].
// Writes Q[b, t*H + h, :] (rotated), K[b, key_off + t, :] (rotated), V[b, key_off + t, :].  pos_mode 0: pos[b*T+t];
// pos_mode 1: nvalid[b] + t.  rope table: [max_pos+2, hd/2] (cos, sin) indexed by pos+1.
void rope_pack_fwd(const bf16* qkv, int T, int H, int hd, const int* pos, const int* nvalid, int pos_mode,
                   const bf16* cos_t, const bf16* sin_t, bf16* Q, bf16* K, bf16* V, int key_off, int kv_len, int batch,
                   cudaStream_t st);
// Inverse: dQ (bf16 [b, t*H+h, hd]), dK/dV (fp32 [b, kv_len, hd], rows key_off..key_off+T) -> dqkv rows (bf16).
void rope_pack_bwd(const bf16* dQ, const float* dK, const float* dV, int T, int H, int hd, const int* pos,
                   const int* nvalid, int pos_mode, const bf16* cos_t, const bf16* sin_t, bf16* dqkv, int key_off,
                   int kv_len, int batch, cudaStream_t st);
// In-place softmax over bf16 score rows (modeling_gemma.py:246-248): P = bf(softmax_fp32(s + mask)).
// rows = batch*rows_per_batch, row length n_keys (ld = row pitch).  key j < n_prefix is valid iff pad[b*n_prefix+j]
// (pad == null: all valid); keys >= n_prefix always valid.  qpad (optional): padded query rows get a uniform row.
void softmax_fwd(bf16* s, int64_t ld, int rows_per_batch, int batch, int n_keys, int n_prefix, const uint8_t* pad,
                 const uint8_t* qpad, int q_per_token, cudaStream_t st);
// dS = bf(bf(P*(dP - sum(dP*P))) * scale), written over dP.
void softmax_bwd(const bf16* p, bf16* dp, int64_t ld, int rows, int n_keys, float scale, cudaStream_t st);

// ---------------- embeddings & element-wise tails (misc_kernels.cu) ----------------
// out[b, row_off + l, :] = bf(E[tok[b,l]] * sqrt(width))   (gemma_pytorch.py:88-89, pi0_pytorch.py:213-216)
void embed_tokens_fwd(const int64_t* tok, const bf16* table, bf16* out, int batch, int L, int width, int64_t out_bstride,
                      int row_off, float scale, cudaStream_t st);
// dense bf16 gradient of the table: rows of used tokens only (others must be pre-zeroed); padding_idx 0 skipped.
void embed_tokens_bwd(const int64_t* tok, const bf16* dout, int64_t dout_bstride, int row_off, bf16* dtable,
                      float* scratch /*[batch*L, width]*/, int* first /*[batch*L]*/, int batch, int L, int width,
                      float scale, cudaStream_t st);
// GeGLU backward (modeling_gemma.py:125): GU = [g | u] (ld 2*n); dGU = [dg | du].
void geglu_bwd(const bf16* dh, const bf16* gu, bf16* dgu, int64_t rows, int n, cudaStream_t st);
// h = bf(bf(gelu(g))*u) recompute (when H is not stashed)
void geglu_fwd(const bf16* gu, bf16* h, int64_t rows, int n, cudaStream_t st);
// dpre = bf(dact * gelu'(pre))
void gelu_bwd(const bf16* dact, const bf16* pre, bf16* dpre, int64_t n, cudaStream_t st);
// acc32[c] += sum_r x[r, c]   (bias gradients)
void colsum_bf16(const bf16* x, int64_t ld, int64_t rows, int cols, float* acc32, cudaStream_t st);
void cast_f32_to_bf16(const float* in, bf16* out, int64_t n, cudaStream_t st);
void cast_bf16_to_f32(const bf16* in, float* out, int64_t n, cudaStream_t st);
void add_bf16(const bf16* a, const bf16* b, bf16* out, int64_t n, cudaStream_t st);  // out = bf(a + b)
void add_f32(float* a, const float* b, int64_t n, cudaStream_t st);                  // a += b
void fill_zero(void* p, size_t bytes, cudaStream_t st);
void fill_f32(float* p, float v, int64_t n, cudaStream_t st);
// out[s] = the fp32 running sum 1, 1+dt, (1+dt)+dt, ... of sample_actions' time variable (pi0_pytorch.py:401-418)
void decode_times(float* out, int n, float dt, cudaStream_t st);
// sincos time embedding in fp64 (pi0_pytorch.py:25-42,264-267): out[b, :] fp32 [2*half]
void time_embedding(const float* time, const double* scaling /*[half]*/, float* out, int batch, int half, cudaStream_t st);
void silu_fwd(const float* x, float* y, int64_t n, cudaStream_t st);
void silu_bwd(const float* dy, const float* x, float* dx, int64_t n, cudaStream_t st);  // dx = dy * silu'(x)
// flow matching (pi0_pytorch.py:326-328,373): x_t, u_t; loss = (u_t - v_t)^2; dv = -2 (u_t - v_t) dloss
void flow_inputs(const float* actions, const float* noise, const float* time, float* x_t, float* u_t, int batch, int per,
                 cudaStream_t st);
void flow_loss(const float* u_t, const float* v_t, float* loss, int64_t n, cudaStream_t st);
void flow_loss_bwd(const float* u_t, const float* v_t, const float* dloss, float* dv, int64_t n, cudaStream_t st);
void euler_step(float* x, const float* v, float dt, int64_t n, cudaStream_t st);  // x += dt*v (fp32, P:417)
// copy rows [b, row_off:row_off+T, :] of a [batch, S, width] bf16 tensor to a contiguous fp32 [batch*T, width]
void gather_rows_f32(const bf16* in, int64_t in_bstride, int row_off, int T, int width, float* out, int batch,
                     cudaStream_t st);
// out[(b*T + t), :] = in[b, row_off + t, :]  (bf16 row gather out of a [batch, S, width] tensor)
void copy_rows_bf16(const bf16* in, int64_t in_bstride, int row_off, int T, int width, bf16* out, int batch,
                    cudaStream_t st);
void scatter_rows_bf16(const float* in, bf16* out, int64_t out_bstride, int row_off, int T, int width, int batch,
                       cudaStream_t st);

// ---------------- AdvantageEstimator head (value_kernels.cu; pi0_pytorch.py:473-481,560-587) ----------------
void tanh_fwd(const float* x, float* y, int64_t n, cudaStream_t st);
// loss[b,t] = w_a*mean_d (u-v)^2 + w_v*(value[b]-clamp(progress[b],-1,1))^2; la[b,t] = unweighted action loss,
// lv[b] = weighted value loss; aux (optional) = {mean(la), mean(lv)}
void advantage_loss(const float* u, const float* v, const float* value, const float* progress, float w_a, float w_v,
                    float* loss, float* la, float* lv, float* aux, int B, int A, int ad, cudaStream_t st);
// dv [B,A,ad] and dpre[b] = d loss / d(pre-tanh value) for dloss [B,A]
void advantage_loss_bwd(const float* u, const float* v, const float* value, const float* progress, const float* dloss,
                        float w_a, float w_v, float* dv, float* dpre, int B, int A, int ad, cudaStream_t st);
// g[b, 0, :] = bf(g[b, 0, :] + bf(dx[b, :]))  for g bf16 [B, A, E]
void add_row0_grad(bf16* g, const float* dx, int B, int A, int E, cudaStream_t st);

// ---------------- observation preprocessing (preprocess_kernels.cu; preprocessing_pytorch.py:35-148) ----------------
// One image key: fp32 [-1,1] input [B,3,h,w] (channels_last = 0) or [B,h,w,3] (1) -> out fp32 [B,3,S,S].
// Resize-with-pad when (h,w) != (S,S); with train != 0 the augmentation with the 6 device-resident parameters
// {start_h, start_w, angle_deg, brightness, contrast, saturation}; geometric = 0 for wrist cameras.
size_t preprocess_scratch_floats(int batch, int out_size);
void preprocess_image(const float* data, int height, int width, int channels_last, int batch, int out_size, int train,
                      int geometric, const float* params, float* scratch, float* out, cudaStream_t st);
// Same preprocessing, written straight into the patch-embedding GEMM operand (row f2): `data` fp32 [-1,1] or uint8 (is_u8),
// rows = bf16 [batch * (out_size/patch)^2, 3 * patch_row_kp(patch)] for THIS image key, column blocks [hi | lo | hi].
// The padding columns [3*patch*patch, Kp) of every block are never written: the caller zero-fills the buffer once.
inline int patch_row_kp(int patch) { return (3 * patch * patch + 7) / 8 * 8; }
void preprocess_patches(const void* data, int is_u8, int height, int width, int channels_last, int batch, int out_size,
                        int patch, int train, int geometric, const float* params, float* scratch, bf16* rows,
                        cudaStream_t st);

// ---------------- fp32 SIMT linears (sgemm_f32.cu) ----------------
// Y[M,N] = X[M,K] W[N,K]^T + bias   (nn.Linear in fp32: action_in/out_proj, time MLP, adaRMS dense)
void linear_f32(const float* X, const float* W, const float* bias, float* Y, int M, int N, int K, cudaStream_t st);
// dX[M,K] (+)= dY[M,N] W[N,K]
void linear_f32_dgrad(const float* dY, const float* W, float* dX, int M, int N, int K, int accumulate, cudaStream_t st);
// dW[N,K] = dY[M,N]^T X[M,K];  db[N] = colsum(dY)  (db may be null)
void linear_f32_wgrad(const float* dY, const float* X, float* dW, float* db, int M, int N, int K, cudaStream_t st);
// Batched variants for the 2*depth+1 adaRMS dense layers (uniformly strided weights): one launch each.
void linear_f32_batched(const float* X, const float* W, const float* bias, float* Y, int M, int N, int K, int batch,
                        int64_t w_stride, int64_t b_stride, int64_t y_stride, cudaStream_t st);
void linear_f32_wgrad_batched(const float* dY, const float* X, float* dW, float* db, int M, int N, int K, int batch,
                              int64_t dy_stride, int64_t w_stride, int64_t b_stride, cudaStream_t st);
void linear_f32_dgrad_batched_sum(const float* dY, const float* W, float* dX, float* scratch, int M, int N, int K,
                                  int batch, int64_t dy_stride, int64_t w_stride, cudaStream_t st);
// Patch embedding = conv 14x14/14 as an fp32 GEMM over an implicit im2col (modeling_siglip.py:220-226,271-282):
// out[img, patch, c] = bf( sum W[c, ch,py,px] * img[ch, ...] + bias[c] + pos[patch, c] )
void patch_embed_fwd(const float* images, const float* W, const float* bias, const float* pos, bf16* out, int n_img,
                     int image_size, int patch, int width, cudaStream_t st);
// dW[c, 3*p*p] , dbias[c], dpos[patch, c] from dout (bf16 [n_img*patches, width])
void patch_embed_bwd(const float* images, const bf16* dout, float* dW, float* dbias, float* dpos, float* scratch,
                     int n_img, int image_size, int patch, int width, cudaStream_t st);
void patch_embed_bwd_pos_bias(const bf16* dout, float* dbias, float* dpos, int n_img, int tokens, int width, cudaStream_t st);
// patch-row path (row f2): Conv2d weight fp32 [rows, k] -> bf16 [rows, 3*kp] = [hi | hi | lo]; and the fold of the
// weight-gradient GEMM's two column blocks back into the fp32 [rows, k] gradient
void split_patch_weight(const float* w, bf16* out, int rows, int k, int kp, cudaStream_t st);
void fold_patch_dw(const float* d, float* dw, int rows, int k, int kp, cudaStream_t st);

}  // namespace pi05
