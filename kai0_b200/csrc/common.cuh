// Shared device helpers for the element-wise / reduction kernels: 128-bit bf16 vector I/O, warp reductions.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pdl.cuh"

namespace pi05 {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float bfr(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 o;
  o.x = pack2(f[0], f[1]);
  o.y = pack2(f[2], f[3]);
  o.z = pack2(f[4], f[5]);
  o.w = pack2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// same 2-MUFU tanh as the GEMM epilogue (ptx.cuh::fast_tanh) so recomputed activations match the stored ones
__device__ __forceinline__ float fast_tanh_c(float x) {
  const float e = exp2f(x * 2.8853900817779268f);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float gelu_tanh_fw(float x) {
  const float kBeta = 0.7978845608028654f;
  const float kKappa = 0.044715f;
  const float inner = kBeta * (x + kKappa * x * x * x);
  return 0.5f * x * (1.0f + fast_tanh_c(inner));
}
__device__ __forceinline__ float gelu_tanh_bw(float x) {
  const float kBeta = 0.7978845608028654f;
  const float kKappa = 0.044715f;
  const float x2 = x * x;
  const float inner = kBeta * (x + kKappa * x2 * x);
  const float t = fast_tanh_c(inner);
  const float left = 0.5f * x * ((1.0f - t * t) * (kBeta * (1.0f + 3.0f * kKappa * x2)));
  const float right = 0.5f * (1.0f + t);
  return left + right;
}

// gelu(x) and gelu'(x) from ONE tanh evaluation (same expressions as the two functions above)
__device__ __forceinline__ void gelu_tanh_fw_bw(float x, float& fw, float& bw) {
  const float kBeta = 0.7978845608028654f;
  const float kKappa = 0.044715f;
  const float x2 = x * x;
  const float inner = kBeta * (x + kKappa * x2 * x);
  const float t = fast_tanh_c(inner);
  fw = 0.5f * x * (1.0f + t);
  const float left = 0.5f * x * ((1.0f - t * t) * (kBeta * (1.0f + 3.0f * kKappa * x2)));
  const float right = 0.5f * (1.0f + t);
  bw = left + right;
}

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace pi05
