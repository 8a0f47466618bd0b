/* pi05.h — C ABI of libpi05.so, the B200-native (sm_100a) engine for the pi0.5 hot path of OpenDriveLab/kai0.
 *
 * The reference has no FFI for this path: its boundary is the Python class
 *   openpi.models_pytorch.pi0_pytorch.PI0Pytorch   (src/openpi/models_pytorch/pi0_pytorch.py:84-461)
 * reached from scripts/train_pytorch.py:417,540 and src/openpi/policies/policy.py:110.  This header is the
 * C boundary a maintainer would bind underneath that class (see INTEGRATION.md for the ctypes stub); each entry
 * point names the reference method it replaces.
 *
 * Conventions: every pointer in a call is a DEVICE pointer owned by the caller (torch allocations) unless it
 * says "host"; nothing here synchronises the device; all work is enqueued on the cudaStream_t passed in (as a
 * void*).  Return value 0 = ok; otherwise pi05_last_error() holds a message (thread-local, host string).
 */
#ifndef PI05_H_
#define PI05_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PI05_ABI_VERSION 2

typedef struct pi05_engine pi05_engine; /* opaque */

/* ---- model description: mirrors what PI0Pytorch.__init__ reads from its config (pi0_pytorch.py:85-109) and
 *      openpi.models.gemma.get_config (src/openpi/models/gemma.py:58-110) ------------------------------------ */
typedef struct pi05_gemma_cfg {
  int32_t width, depth, mlp_dim, num_heads, num_kv_heads, head_dim;
} pi05_gemma_cfg;

typedef struct pi05_config {
  pi05_gemma_cfg paligemma; /* gemma_2b  */
  pi05_gemma_cfg expert;    /* gemma_300m, adaRMS conditioned (pi05 = True) */
  /* SigLIP-So400m/14 vision tower (gemma_pytorch.py:38-41 + HF SiglipVisionConfig defaults) */
  int32_t vit_width, vit_depth, vit_mlp_dim, vit_heads, vit_patch, image_size;
  int32_t vocab_size;     /* 257152 */
  int32_t action_dim;     /* 32  */
  int32_t action_horizon; /* 50  */
  int32_t max_token_len;  /* 200 */
  int32_t num_images;     /* 3 for PI0Pytorch; up to 6 for AdvantageEstimator */
  int32_t max_batch;      /* largest per-GPU batch the workspace is sized for */
  int32_t train;          /* 1: size the activation stash for backward */
  int32_t value_head;     /* 1: AdvantageEstimator value head (pi0_pytorch.py:473-481) */
  int32_t rtc;            /* 1 (inference engines): per-layer suffix stash + suffix backward scratch for pi05_denoise_rtc */
} pi05_config;

/* dtype codes used in tensor descriptors */
enum { PI05_F32 = 0, PI05_BF16 = 1, PI05_I32 = 2, PI05_U8 = 3 }; /* I32 / U8: index taps only (pi05_get_tap) */

/* One named parameter (reference state_dict key), its data and (optionally) its gradient buffer.
 * The Python module owns both as torch tensors; the engine only keeps the pointers. */
typedef struct pi05_param {
  const char* name; /* host string, e.g. "paligemma_with_expert.paligemma.model.language_model.layers.0.self_attn.q_proj.weight" */
  int32_t dtype;
  int64_t numel;
  void* data;
  void* grad; /* may be NULL (inference) */
} pi05_param;

/* One preprocessed batch = what PI0Pytorch._preprocess_observation returns (pi0_pytorch.py:161-170). */
typedef struct pi05_batch {
  int32_t batch;
  const float* images;         /* [num_images][batch,3,H,W] fp32 in [-1,1], image-major (one pointer, contiguous) */
  const uint8_t* image_masks;  /* [num_images][batch] bool */
  const int64_t* tokens;       /* [batch, max_token_len] */
  const uint8_t* token_mask;   /* [batch, max_token_len] bool */
  /* Optional (ABI 2): the images already laid out as the operand of the patch-embedding GEMM by pi05_preprocess_patches,
   * bf16 [num_images * batch * T, 3 * Kp] with T = (image_size / patch)^2, Kp = pi05_patch_row_kp(patch), rows ordered
   * (image, sample, patch).  When set, `images` is not read (it may be NULL) and the fp32 im2col convolution is replaced
   * by one tcgen05 GEMM on the split operands (see pi05_preprocess_patches). */
  const void* patch_rows;
  /* Optional (ABI 2): number of prompt slots actually present in `tokens` / `token_mask`, which are then [batch, token_len]
   * (0 = max_token_len).  A caller whose prompts are left-aligned (valid slots first, as the reference's tokenizer pads:
   * models/tokenizer.py:35-38) may drop the trailing slots that are padding in EVERY sample of the batch: padded slots are
   * masked keys (probability exactly 0) and their own rows feed nothing, so every output and every gradient is unchanged
   * while the prefix shrinks from num_images * T + max_token_len to num_images * T + token_len rows. */
  int32_t token_len;
} pi05_batch;

/* ---- lifecycle ------------------------------------------------------------------------------------------ */
int pi05_abi_version(void);
const char* pi05_last_error(void);

/* Bytes of device workspace pi05_create will carve up for this config (activations, stash, KV cache). */
size_t pi05_workspace_bytes(const pi05_config* cfg);
/* `workspace` must stay alive and 256B-aligned for the life of the engine. */
int pi05_create(const pi05_config* cfg, int device, void* workspace, size_t workspace_bytes, pi05_engine** out);
void pi05_destroy(pi05_engine* e);
/* Bind every parameter by reference name; unknown or missing names are an error (lists them). */
int pi05_bind_params(pi05_engine* e, const pi05_param* params, int n);
/* Called after the optimiser changed weights in place (refreshes derived tables). */
int pi05_params_updated(pi05_engine* e, void* stream);

/* ---- training: replaces PI0Pytorch.forward (pi0_pytorch.py:316-373) and its autograd backward ----------- */
/* actions / noise: [batch, action_horizon, action_dim] fp32, time: [batch] fp32.  noise and time are drawn by the
 * caller exactly as the reference does (torch.normal / Beta(1.5,1)*0.999+0.001, pi0_pytorch.py:172-184,320-324).
 * loss_out: [batch, action_horizon, action_dim] fp32 = (u_t - v_t)^2, i.e. F.mse_loss(reduction="none").        */
int pi05_forward(pi05_engine* e, const pi05_batch* b, const float* actions, const float* noise, const float* time,
                 float* loss_out, void* stream);
/* dloss: gradient w.r.t. loss_out, same shape.  Writes every bound .grad buffer (overwrites, does not accumulate). */
int pi05_backward(pi05_engine* e, const float* dloss, void* stream);
/* Enable (1) / disable (0) recording of named intermediates for pi05_get_tap. */
int pi05_set_taps(pi05_engine* e, int enabled);

/* ---- inference: replaces PI0Pytorch.sample_actions (pi0_pytorch.py:375-419) ----------------------------- */
/* prefix pass + KV cache (gemma_pytorch.py:102-113) */
int pi05_prefill(pi05_engine* e, const pi05_batch* b, void* stream);
/* `num_steps` Euler steps from `noise` [batch,H,A] fp32 -> actions_out (pi0_pytorch.py:401-419,421-461) */
int pi05_denoise(pi05_engine* e, const float* noise, int num_steps, float* actions_out, void* stream);

/* Real-time-chunking guided decoding.  The reference has it in its JAX model only
 * (src/openpi/models/pi0_rtc.py:234-360); this is that sampler on the PyTorch-path network: per Euler step the velocity of
 * pi05_denoise's step AND the vector-Jacobian product of the denoiser x - t v(x) with the prefix-weighted error to the
 * previous chunk (pi0_rtc.py:331-339), i.e. an input-gradient backward through the action expert, then
 * v <- nan_to_num(v - guidance * J^T err), x <- x + dt v (:348-350).  Engine created with cfg.rtc = 1; pi05_prefill first.
 *   prev_chunk   device fp32 [batch, action_horizon, action_dim], already NaN-cleaned and padded / cut to action_dim (:317-324)
 *   time_weights device fp32 [action_horizon]  = get_prefix_weights(delay, execute_horizon, horizon, schedule) (:47-61,337)
 *   dim_mask     device fp32 [action_dim]      = 1 for the first min(14, provided, action_dim) dims (:326-327)
 *   guidance     HOST   fp32 [num_steps]       = min(c * inv_r2, max_guidance_weight) of every step (:341-347)
 *   mask_rows / provided: with mask_prefix_delay the first `mask_rows` horizon rows of the first `provided` dims of the
 *     denoiser input are overwritten by prev_chunk (:328-333); 0 / 0 otherwise.
 * Exactly num_steps steps (lax.scan, :354-358), time an fp32 running sum from 1.0 by -1/num_steps. */
int pi05_denoise_rtc(pi05_engine* e, const float* noise, int32_t num_steps, const float* prev_chunk,
                     const float* time_weights, const float* dim_mask, const float* guidance, int32_t mask_rows,
                     int32_t provided, float* actions_out, void* stream);

/* ---- AdvantageEstimator (pi0_pytorch.py:464-644); engine created with cfg.value_head = 1 ------------------ */
/* Replaces AdvantageEstimator.forward (pi0_pytorch.py:499-592).  progress: [batch] fp32 (clamped to [-1,1] inside,
 * :574).  loss_out: [batch, action_horizon] fp32 = w_action * mean_d (u_t - v_t)^2 + w_value * (value - progress)^2
 * (:564-587, the [batch,1] value loss broadcast over the horizon).  aux_out (device, may be NULL): 2 floats,
 * {mean(loss_action), mean(w_value * value_loss)} = the reference's loss_aux_dict (:582-583).  Followed by
 * pi05_backward with dloss of shape [batch, action_horizon]. */
int pi05_forward_advantage(pi05_engine* e, const pi05_batch* b, const float* actions, const float* noise,
                           const float* time, const float* progress, float w_action, float w_value, float* loss_out,
                           float* aux_out, void* stream);
/* Replaces AdvantageEstimator.sample_values (pi0_pytorch.py:596-644): one joint forward without KV cache on
 * x_t = noise, t = time (both drawn by the caller as the reference does, :604-605); value_out: [batch] fp32. */
int pi05_forward_value(pi05_engine* e, const pi05_batch* b, const float* noise, const float* time, float* value_out,
                       void* stream);

/* Debug taps: copy a named intermediate of the last forward into `dst` (fp32 or bf16 as stored). Returns its
 * element count through *numel and dtype through *dtype; dst may be NULL to query.  Used by parity tests only. */
int pi05_get_tap(pi05_engine* e, const char* name, void* dst, int64_t* numel, int32_t* dtype, void* stream);
/* Profiling hook (ncu --profile-from-start off): bracket joint layer `layer` (and ViT layer `layer` when it exists)
 * of every following forward / backward with cudaProfilerStart/Stop; layer < 0 disables.  Instrumentation only. */
int pi05_debug_profile_layer(pi05_engine* e, int layer);
/* Programmatic dependent launch on (1, default; PI05_PDL=0 in the environment also disables) / off (0) for all
 * following launches of the library.  Per-kernel timings from profilers are only meaningful with it off. */
int pi05_debug_set_pdl(int enabled);

/* ---- stand-alone operator: observation preprocessing of ONE image key on the device --------------------------
 * Replaces the per-image body of preprocess_observation_pytorch (src/openpi/models_pytorch/preprocessing_pytorch.py:
 * 35-148) and resize_with_pad_torch (src/openpi/shared/image_tools.py:55-126).
 *   image: fp32 in [-1,1], [batch,3,height,width] (channels_last = 0) or [batch,height,width,3] (1)
 *   out:   fp32 [batch,3,out_size,out_size]  (what pi05_batch.images holds for this key)
 *   (height,width) != out_size: aspect-preserving bilinear resize, clamp to [-1,1], pad with -1.
 *   train != 0: the reference's augmentation with the parameters the caller drew exactly as the reference draws
 *     them (:70-72,85,124,129,137; ONE draw per batch): params = device pointer to 6 floats
 *     {crop start_h, crop start_w, angle in degrees, brightness, contrast, saturation}; geometric = 1 for the
 *     non-wrist cameras (95 % crop + resize, rotation when |angle| > 0.1), 0 for wrist cameras (colour only).
 *   scratch: device, at least pi05_preprocess_scratch_floats(batch, out_size) floats. */
size_t pi05_preprocess_scratch_floats(int32_t batch, int32_t out_size);
int pi05_preprocess_image(const float* image, int32_t height, int32_t width, int32_t channels_last, int32_t batch,
                          int32_t out_size, int32_t train, int32_t geometric, const float* params, float* scratch,
                          float* out, void* stream);

/* Same preprocessing for ONE image key, fused with what sits either side of it on the path:
 *   in:  image dtype PI05_F32 (in [-1,1]) or PI05_U8 -- the uint8 -> fp32 `x / 255 * 2 - 1` of Observation.from_dict
 *        (src/openpi/models/model.py:129-133) is then taken on the fly, same three fp32 roundings;
 *   out: `rows` = this key's slice of pi05_batch.patch_rows: bf16 [batch * T, 3 * Kp], row = sample * T + patch,
 *        column = c * patch^2 + (y % patch) * patch + (x % patch) (the flattening of the Conv2d weight,
 *        modeling_siglip.py:220-226), each fp32 pixel v stored as the split hi = bf16(v), lo = bf16(v - hi) in three column
 *        blocks [hi | lo | hi]; the engine multiplies them with the weight blocks [Whi | Whi | Wlo] in ONE bf16 GEMM with fp32
 *        accumulation: v * w to 2^-16 relative, i.e. the fp32 convolution of the reference without its CUDA-core cost.
 *        Columns [3 * patch^2, Kp) of every block are padding the caller zero-fills once.
 * Everything else (layout sniffing, resize-with-pad, train-time augmentation, params, scratch) as pi05_preprocess_image. */
int32_t pi05_patch_row_kp(int32_t patch);
int pi05_preprocess_patches(const void* image, int32_t image_dtype, int32_t height, int32_t width, int32_t channels_last,
                            int32_t batch, int32_t out_size, int32_t patch, int32_t train, int32_t geometric,
                            const float* params, float* scratch, void* rows, void* stream);

/* ---- stand-alone operator: the tcgen05 GEMM that every nn.Linear / matmul of the path maps to ----------- */
typedef struct pi05_gemm_desc {
  int32_t M, N, K, batch;
  int32_t batch_inner; /* two-level batch z = z1*batch_inner + z0; 0 = one level */
  const void* A; /* bf16 */
  const void* B; /* bf16 */
  int32_t a_major, b_major; /* 0: [rows,K] K contiguous; 1: [K,rows] rows contiguous */
  int64_t lda, ldb, a_batch_stride, b_batch_stride; /* elements */
  int64_t a_batch_stride1, b_batch_stride1;
  int32_t epilogue; /* see GemmEpilogue in csrc/gemm.h */
  void* D;
  int64_t ldd, d_batch_stride, d_batch_stride1;
  void* D2;
  int64_t ldd2, d2_batch_stride, d2_batch_stride1;
  const void* bias;
  const void* res;
  int64_t ldres, res_batch_stride, res_batch_stride1;
  const void* gate;
  int32_t gate_rows;
  int64_t ldgate;
  float scale;
  int32_t accumulate;
  int32_t block_n;
  /* optional fp32 scratch (device): enables the split-K paths (small-M decode GEMMs; weight-gradient shapes whose
   * tile count quantises badly onto the SMs).  NULL = never split.  Results are deterministic either way. */
  void* workspace;
  size_t workspace_bytes;
} pi05_gemm_desc;
int pi05_gemm_bf16(const pi05_gemm_desc* d, void* stream);

/* ---- caller-side step right after the path: global-norm clip + AdamW in one pass --------
 * Replaces torch.nn.utils.clip_grad_norm_(params, max_norm) + torch.optim.AdamW.step() of scripts/train_pytorch.py:
 * 557-560 over the two flat arenas (state m, v in the parameter dtype).  `step` is the 1-based update count,
 * max_norm <= 0 disables clipping.  scratch: >= 4096 floats (device); after the call scratch[0] = total gradient
 * norm, scratch[1] = applied clip coefficient.                                                                   */
int pi05_fused_clip_adamw(void* p_bf16, const void* g_bf16, void* m_bf16, void* v_bf16, int64_t n_bf16, float* p_f32,
                          const float* g_f32, float* m_f32, float* v_f32, int64_t n_f32, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int64_t step, float max_norm, float* scratch,
                          void* stream);

/* Same, with a gradient pre-scale: the arenas hold SUMS over `1 / grad_scale` data-parallel ranks (pi05_set_grad_exchange
 * with average_in_place = 0) and the averaged gradient DDP would have produced, bf(g * grad_scale), is taken on the fly
 * (norm, clip and update all see the averaged value; scratch[0] = norm of the AVERAGED gradient). */
int pi05_fused_clip_adamw_scaled(void* p_bf16, const void* g_bf16, void* m_bf16, void* v_bf16, int64_t n_bf16, float* p_f32,
                                 const float* g_f32, float* m_f32, float* v_f32, int64_t n_f32, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, int64_t step, float max_norm,
                                 float grad_scale, float* scratch, void* stream);

/* ---- data-parallel gradient exchange: replaces DistributedDataParallel's bucketed all-reduce
 * (scripts/train_pytorch.py:440-447) ----------------------------------------------------------------------------
 * `nccl_comm` is an ncclComm_t created by the caller over the data-parallel ranks (libnccl is resolved at run time with
 * dlopen: the library has no link-time NCCL dependency).
 *
 * pi05_set_grad_exchange: from now on pi05_backward itself exchanges the gradients (sum all-reduce of both arenas, the
 * never-used expert lm_head excluded) before it returns control of `stream`:
 *   overlap = 0: ONE exchange at the end of backward on the caller's stream, at NCCL's full speed;
 *   overlap = 1: OVERLAPPED with backward: as soon as the kernels producing one contiguous group of the gradient arena
 *     (an expert layer, a PaliGemma layer, the fp32 norm / adaRMS / head block, the embedding table, groups of SigLIP
 *     layers) are enqueued, a ncclAllReduce of that range is enqueued on an engine-owned high-priority stream behind an
 *     event; at the end the caller's stream waits for that stream.  Create the communicator with few CTAs for this mode
 *     (pi05_nccl_comm_create max_ctas): the tcgen05 GEMMs schedule tiles dynamically, so the SMs the NCCL kernels hold cost
 *     their share and no more.  Measured on power-capped B200s the overlapped mode is SLOWER (N = 2: 479 vs 471 ms,
 *     N = 8: 483-486 vs 478 ms per step): the NCCL kernels' power comes out of the GEMMs' clock.  It is kept for parts
 *     that are not power-limited; overlap = 0 is what bench.py uses.
 * average_in_place != 0: every range is also multiplied by 1 / nranks (what DDP's averaging leaves in .grad); 0: the
 * arenas hold sums and the caller folds 1 / nranks into its optimiser (pi05_fused_clip_adamw_scaled).
 * nccl_comm = NULL switches the exchange off again.
 *
 * pi05_allreduce_grads: the overlap = 0 exchange as a separate call on `stream` after pi05_backward. */
int pi05_set_grad_exchange(pi05_engine* e, void* nccl_comm, int32_t nranks, int32_t average_in_place, int32_t overlap);
int pi05_allreduce_grads(pi05_engine* e, void* nccl_comm, int32_t nranks, int32_t average, void* stream);
/* Communicator helpers for hosts without an NCCL binding of their own (the Python host uses them through ctypes):
 * rank 0 calls pi05_nccl_unique_id (128 bytes out) and ships the id to every rank by its own means (torch.distributed's
 * store here); every rank then calls pi05_nccl_comm_create (collective).  max_ctas > 0 bounds the SMs the communicator's
 * kernels may occupy (ncclConfig_t.maxCTAs, for the overlapped exchange); max_ctas < 0 asks for at least -max_ctas CTAs
 * (ncclConfig_t.minCTAs, for the exchange at the end of backward); 0 = NCCL's defaults. */
int pi05_nccl_unique_id(void* out128);
int pi05_nccl_comm_create(const void* unique_id128, int32_t nranks, int32_t rank, int32_t max_ctas, void** comm_out);
int pi05_nccl_comm_destroy(void* comm);
/* Statistics of the last pi05_backward's exchange: *calls = ncclAllReduce launches, *bytes = payload bytes. */
int pi05_grad_exchange_stats(pi05_engine* e, int64_t* calls, int64_t* bytes);

/* ---- instrumentation (bench.py): kernel-launch counter and per-launch CUDA-event timing of the GEMM -------- */
unsigned long long pi05_launch_count(void);
void pi05_gemm_profile_enable(int on);
/* one text line per GEMM class: "M N K batch epilogue majors launches total_ms"; returns bytes written (host buf) */
int pi05_gemm_profile_report(char* buf, int len);

#ifdef __cplusplus
}
#endif
#endif /* PI05_H_ */
