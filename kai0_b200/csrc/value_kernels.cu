// AdvantageEstimator head (pi0_pytorch.py:473-481, 560-587): loss assembly, its backward and the merge of the value
// head's gradient into row 0 of d(suffix_out).  All fp32, tiny; one thread per (b, t) row or per element.
#include "common.cuh"
#include "errors.h"
#include "kernels.h"
#include "launch.h"

namespace pi05 {
namespace {

__global__ void tanh_fwd_k(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  pdl_enter();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = tanhf(x[i]);
}

// loss[b,t] = w_a * mean_d (u-v)^2 + w_v * (value[b] - clamp(progress[b]))^2
__global__ void advantage_loss_k(const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ value,
                                 const float* __restrict__ progress, float w_a, float w_v, float* __restrict__ loss,
                                 float* __restrict__ la, float* __restrict__ lv, int B, int A, int ad) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * A) return;
  const int b = i / A;
  const float* up = u + static_cast<int64_t>(i) * ad;
  const float* vp = v + static_cast<int64_t>(i) * ad;
  float s = 0.0f;
  for (int d = 0; d < ad; ++d) {
    const float diff = up[d] - vp[d];
    s += diff * diff;
  }
  const float mean = s / static_cast<float>(ad);
  const float tgt = fminf(fmaxf(progress[b], -1.0f), 1.0f);
  const float dv = value[b] - tgt;
  const float vl = dv * dv * w_v;
  la[i] = mean;
  if (i % A == 0) lv[b] = vl;
  loss[i] = mean * w_a + vl;
}

// deterministic single-block means: out[0] = mean(la[0..n1)), out[1] = mean(lv[0..n2))
__global__ void advantage_aux_k(const float* __restrict__ la, int n1, const float* __restrict__ lv, int n2,
                                float* __restrict__ out) {
  pdl_enter();
  __shared__ float sh[2][32];
  float a = 0.0f, b = 0.0f;
  for (int i = threadIdx.x; i < n1; i += blockDim.x) a += la[i];
  for (int i = threadIdx.x; i < n2; i += blockDim.x) b += lv[i];
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sh[0][w] = a;
    sh[1][w] = b;
  }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    a = l < nw ? sh[0][l] : 0.0f;
    b = l < nw ? sh[1][l] : 0.0f;
    a = warp_sum(a);
    b = warp_sum(b);
    if (l == 0) {
      out[0] = a / static_cast<float>(n1);
      out[1] = b / static_cast<float>(n2);
    }
  }
}

// dv[b,t,d] = dloss[b,t] * w_a / ad * 2 (v - u);  dpre[b] = (sum_t dloss[b,t]) * w_v * 2 (value - tgt) * (1 - value^2)
__global__ void advantage_loss_bwd_k(const float* __restrict__ u, const float* __restrict__ v,
                                     const float* __restrict__ value, const float* __restrict__ progress,
                                     const float* __restrict__ dloss, float w_a, float w_v, float* __restrict__ dv,
                                     float* __restrict__ dpre, int B, int A, int ad) {
  pdl_enter();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * A * ad;
  if (i < total) {
    const int64_t row = i / ad;
    const float g = dloss[row] * w_a / static_cast<float>(ad);
    dv[i] = g * 2.0f * (v[i] - u[i]);
  }
  if (i < B) {
    const int b = static_cast<int>(i);
    float s = 0.0f;
    for (int t = 0; t < A; ++t) s += dloss[b * A + t];
    const float tgt = fminf(fmaxf(progress[b], -1.0f), 1.0f);
    const float val = value[b];
    dpre[b] = s * w_v * 2.0f * (val - tgt) * (1.0f - val * val);
  }
}

// g[b*A + 0, :] = bf( g + bf(dx[b, :]) ): autograd sums the two bf16 gradients of suffix_out (slice + select)
__global__ void add_row0_grad_k(bf16* __restrict__ g, const float* __restrict__ dx, int B, int A, int E) {
  pdl_enter();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(B) * E) return;
  const int b = static_cast<int>(i / E), c = static_cast<int>(i % E);
  bf16* p = g + static_cast<int64_t>(b) * A * E + c;
  *p = __float2bfloat16_rn(__bfloat162float(*p) + bfr(dx[i]));
}

}  // namespace

void tanh_fwd(const float* x, float* y, int64_t n, cudaStream_t st) {
  launch_pdl(tanh_fwd_k, dim3(ceil_div(n, 256)), dim3(256), 0, st, x, y, n);
  count_launch();
}
void advantage_loss(const float* u, const float* v, const float* value, const float* progress, float w_a, float w_v,
                    float* loss, float* la, float* lv, float* aux, int B, int A, int ad, cudaStream_t st) {
  launch_pdl(advantage_loss_k, dim3(ceil_div(static_cast<int64_t>(B) * A, 128)), dim3(128), 0, st, u, v, value, progress, w_a, w_v, loss, la,
                                                                             lv, B, A, ad);
  count_launch();
  if (aux != nullptr) {
    launch_pdl(advantage_aux_k, dim3(1), dim3(256), 0, st, la, B * A, lv, B, aux);
    count_launch();
  }
}
void advantage_loss_bwd(const float* u, const float* v, const float* value, const float* progress, const float* dloss,
                        float w_a, float w_v, float* dv, float* dpre, int B, int A, int ad, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(B) * A * ad;
  launch_pdl(advantage_loss_bwd_k, dim3(ceil_div(total, 256)), dim3(256), 0, st, u, v, value, progress, dloss, w_a, w_v, dv, dpre, B, A, ad);
  count_launch();
}
void add_row0_grad(bf16* g, const float* dx, int B, int A, int E, cudaStream_t st) {
  launch_pdl(add_row0_grad_k, dim3(ceil_div(static_cast<int64_t>(B) * E, 256)), dim3(256), 0, st, g, dx, B, A, E);
  count_launch();
}

}  // namespace pi05
