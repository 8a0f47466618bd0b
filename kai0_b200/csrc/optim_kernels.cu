// Fused global-norm gradient clipping + AdamW over the two flat parameter arenas: the
// caller-side step right after the hot path (scripts/train_pytorch.py:557-560: clip_grad_norm_(1.0) then
// torch.optim.AdamW(betas=(0.9,0.95), eps=1e-8, weight_decay=1e-10), state kept in the parameter dtype).
//
// Pass 1: deterministic two-stage sum of squares of all gradients (fp32 accumulation).
// Pass 2: one streaming pass per arena: g' = g * min(1, max_norm/(norm+1e-6)) (rounded to the gradient dtype, as the
//         in-place torch clip does), then the AdamW update in fp32 and a single rounding of p, m, v back to storage.
// Arithmetic follows torch's fused AdamW (lerp for exp_avg, sqrt(v)/sqrt(bc2) + eps denominator).
#include "../../include/pi05.h"
#include "common.cuh"
#include "errors.h"
#include "launch.h"

namespace pi05 {

namespace {

constexpr int NORM_BLOCKS = 1184;  // 8 per SM

// gs: gradient pre-scale (1 / world size when the data-parallel exchange left SUMS in the arenas: the average DDP would
// have produced is bf(g * gs), taken on the fly here and in the update kernels instead of in a separate 14 GB pass)
__global__ void __launch_bounds__(256) sumsq_k(const bf16* __restrict__ gb, int64_t nb, const float* __restrict__ gf,
                                               int64_t nf, float* __restrict__ partial, float gs) {
  pdl_enter();
  float acc = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nb8 = nb / 8;
  for (int64_t i = tid; i < nb8; i += stride) {
    float v[8];
    load8(gb + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float x = bfr(v[k] * gs);
      acc += x * x;
    }
  }
  for (int64_t i = nb8 * 8 + tid; i < nb; i += stride) {
    const float v = bfr(__bfloat162float(gb[i]) * gs);
    acc += v * v;
  }
  for (int64_t i = tid; i < nf; i += stride) {
    const float v = gf[i] * gs;
    acc += v * v;
  }
  acc = warp_sum(acc);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += sm[k];
    partial[blockIdx.x] = t;
  }
}

// partial[0..n) -> out[0] = total norm, out[1] = clip coefficient
__global__ void __launch_bounds__(256) finish_norm_k(const float* __restrict__ partial, int n, float max_norm,
                                                     float* __restrict__ out) {
  pdl_enter();
  __shared__ double sm[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) t += static_cast<double>(partial[i]);
  sm[threadIdx.x] = t;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = static_cast<float>(sqrt(sm[0]));
    out[0] = norm;
    float coef = 1.0f;
    if (max_norm > 0.f) {
      coef = max_norm / (norm + 1e-6f);  // torch.nn.utils.clip_grad_norm_
      coef = coef > 1.0f ? 1.0f : coef;
    }
    out[1] = coef;
  }
}

struct AdamArgs {
  float lr, beta1, beta2, eps, wd, step_size, bc2_sqrt;
};

__device__ __forceinline__ void adam_math(float& p, float g, float& m, float& v, const AdamArgs& a) {
  p -= a.lr * a.wd * p;
  m = m + (1.0f - a.beta1) * (g - m);  // lerp(m, g, 1 - beta1), weight < 0.5 form
  v = a.beta2 * v + (1.0f - a.beta2) * g * g;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= a.step_size * m / denom;
}

__global__ void __launch_bounds__(256) adamw_bf16_k(bf16* __restrict__ p, const bf16* __restrict__ g, bf16* __restrict__ m,
                                                    bf16* __restrict__ v, int64_t n, AdamArgs a,
                                                    const float* __restrict__ coef_ptr, float gs) {
  pdl_enter();
  const float coef = coef_ptr[1];
  const int64_t n8 = n / 8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float pp[8], gg[8], mm[8], vv[8];
    load8(p + i * 8, pp);
    load8(g + i * 8, gg);
    load8(m + i * 8, mm);
    load8(v + i * 8, vv);
#pragma unroll
    for (int k = 0; k < 8; ++k) adam_math(pp[k], bfr(bfr(gg[k] * gs) * coef), mm[k], vv[k], a);
    store8(p + i * 8, pp);
    store8(m + i * 8, mm);
    store8(v + i * 8, vv);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = n8 * 8; i < n; ++i) {
      float pp = __bfloat162float(p[i]), mm = __bfloat162float(m[i]), vv = __bfloat162float(v[i]);
      adam_math(pp, bfr(bfr(__bfloat162float(g[i]) * gs) * coef), mm, vv, a);
      p[i] = __float2bfloat16_rn(pp);
      m[i] = __float2bfloat16_rn(mm);
      v[i] = __float2bfloat16_rn(vv);
    }
  }
}

__global__ void __launch_bounds__(256) adamw_f32_k(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, AdamArgs a,
                                                   const float* __restrict__ coef_ptr, float gs) {
  pdl_enter();
  const float coef = coef_ptr[1];
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam_math(pp, (g[i] * gs) * coef, mm, vv, a);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
  }
}

}  // namespace
}  // namespace pi05

extern "C" int pi05_fused_clip_adamw(void* p_bf16, const void* g_bf16, void* m_bf16, void* v_bf16, int64_t n_bf16,
                                     float* p_f32, const float* g_f32, float* m_f32, float* v_f32, int64_t n_f32, float lr,
                                     float beta1, float beta2, float eps, float weight_decay, int64_t step, float max_norm,
                                     float* scratch, void* stream) {
  return pi05_fused_clip_adamw_scaled(p_bf16, g_bf16, m_bf16, v_bf16, n_bf16, p_f32, g_f32, m_f32, v_f32, n_f32, lr, beta1,
                                      beta2, eps, weight_decay, step, max_norm, 1.0f, scratch, stream);
}

extern "C" int pi05_fused_clip_adamw_scaled(void* p_bf16, const void* g_bf16, void* m_bf16, void* v_bf16, int64_t n_bf16,
                                            float* p_f32, const float* g_f32, float* m_f32, float* v_f32, int64_t n_f32,
                                            float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                                            float max_norm, float grad_scale, float* scratch, void* stream) {
  using namespace pi05;
  const float gs = grad_scale;
  if (step < 1 || n_bf16 < 0 || n_f32 < 0 || !scratch) {
    set_error("pi05_fused_clip_adamw: bad argument (step >= 1, scratch of 4096 floats required)");
    return 1;
  }
  if ((reinterpret_cast<uintptr_t>(p_bf16) | reinterpret_cast<uintptr_t>(g_bf16) | reinterpret_cast<uintptr_t>(m_bf16) |
       reinterpret_cast<uintptr_t>(v_bf16)) & 15) {
    set_error("pi05_fused_clip_adamw: bf16 arenas must be 16-byte aligned");
    return 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = scratch + 8;  // scratch[0] = norm, scratch[1] = clip coefficient, [8, 8+NORM_BLOCKS) partial sums
  launch_pdl(sumsq_k, dim3(NORM_BLOCKS), dim3(256), 0, st, static_cast<const bf16*>(g_bf16), n_bf16, g_f32, n_f32, partial, gs);
  count_launch();
  launch_pdl(finish_norm_k, dim3(1), dim3(256), 0, st, partial, NORM_BLOCKS, max_norm, scratch);
  count_launch();
  AdamArgs a;
  a.lr = lr;
  a.beta1 = beta1;
  a.beta2 = beta2;
  a.eps = eps;
  a.wd = weight_decay;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  a.step_size = static_cast<float>(static_cast<double>(lr) / bc1);
  a.bc2_sqrt = static_cast<float>(sqrt(bc2));
  if (n_bf16 > 0) {
    launch_pdl(adamw_bf16_k, dim3(148 * 16), dim3(256), 0, st, static_cast<bf16*>(p_bf16), static_cast<const bf16*>(g_bf16),
                                          static_cast<bf16*>(m_bf16), static_cast<bf16*>(v_bf16), n_bf16, a, scratch, gs);
    count_launch();
  }
  if (n_f32 > 0) {
    launch_pdl(adamw_f32_k, dim3(148 * 8), dim3(256), 0, st, p_f32, g_f32, m_f32, v_f32, n_f32, a, scratch, gs);
    count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return 2;
  }
  return 0;
}
