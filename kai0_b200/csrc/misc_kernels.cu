// Embedding gather/scatter, GeGLU / GELU backward, bias (column) sums, casts, time embedding, flow-matching tails.
// Reference arithmetic: gemma_pytorch.py:88-89 + pi0_pytorch.py:213-216 (embedding * sqrt(d)),
// modeling_gemma.py:125 (GeGLU), pi0_pytorch.py:25-42 (sincos, fp64), :326-328,373 (flow matching), :417 (Euler).
#include "common.cuh"
#include "errors.h"
#include "kernels.h"
#include "launch.h"

namespace pi05 {

namespace {

inline int grid_for(int64_t n, int per_block = 256, int cap = 148 * 16) {
  int64_t g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  return static_cast<int>(g < cap ? g : cap);
}
#define GRID_STRIDE(i, n)                                                                 \
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < (n); \
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)

__global__ void embed_tokens_fwd_k(const int64_t* __restrict__ tok, const bf16* __restrict__ table,
                                   bf16* __restrict__ out, int batch, int L, int width, int64_t out_bstride, int row_off,
                                   float scale) {
  pdl_enter();
  const int chunks = width / 8;
  const int64_t total = static_cast<int64_t>(batch) * L * chunks;
  GRID_STRIDE(idx, total) {
    const int ch = static_cast<int>(idx % chunks);
    const int64_t bl = idx / chunks;
    const int b = static_cast<int>(bl / L), l = static_cast<int>(bl % L);
    float v[8];
    load8(table + tok[bl] * width + ch * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= scale;
    store8(out + b * out_bstride + static_cast<int64_t>(row_off + l) * width + ch * 8, v);
  }
}

// first[i] = smallest j with tok[j] == tok[i]
__global__ void first_occurrence_k(const int64_t* __restrict__ tok, int* __restrict__ first, int n) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t t = tok[i];
  int f = i;
  for (int j = 0; j < i; ++j)
    if (tok[j] == t) {
      f = j;
      break;
    }
  first[i] = f;
}
__global__ void embed_accum_k(const int64_t* __restrict__ tok, const int* __restrict__ first,
                              const bf16* __restrict__ dout, int64_t dout_bstride, int row_off,
                              float* __restrict__ scratch, int batch, int L, int width, float scale) {
  pdl_enter();
  const int64_t total = static_cast<int64_t>(batch) * L * width;
  GRID_STRIDE(idx, total) {
    const int c = static_cast<int>(idx % width);
    const int64_t bl = idx / width;
    if (tok[bl] == 0) continue;  // padding_idx (modeling_gemma.py:422-425)
    const int b = static_cast<int>(bl / L), l = static_cast<int>(bl % L);
    // d(emb*scale) -> bf16 rounded product, as autograd does on the bf16 tensor
    const float g = bfr(__bfloat162float(dout[b * dout_bstride + static_cast<int64_t>(row_off + l) * width + c]) * scale);
    atomicAdd(scratch + static_cast<int64_t>(first[bl]) * width + c, g);
  }
}
__global__ void embed_write_k(const int64_t* __restrict__ tok, const int* __restrict__ first,
                              const float* __restrict__ scratch, bf16* __restrict__ dtable, int n, int width) {
  pdl_enter();
  const int64_t total = static_cast<int64_t>(n) * width;
  GRID_STRIDE(idx, total) {
    const int c = static_cast<int>(idx % width);
    const int i = static_cast<int>(idx / width);
    if (first[i] != i || tok[i] == 0) continue;
    dtable[tok[i] * width + c] = __float2bfloat16_rn(scratch[static_cast<int64_t>(i) * width + c]);
  }
}

// One block row-strides over the tensor; each thread owns two 8-wide chunks per row (all six 16-byte loads issued
// before any math) so that enough bytes are in flight per SM to cover HBM latency.
__global__ void __launch_bounds__(256) geglu_bwd_k(const bf16* __restrict__ dh, const bf16* __restrict__ gu,
                                                   bf16* __restrict__ dgu, int64_t rows, int n) {
  pdl_enter();
  const int chunks = n >> 3;
  const int half = (chunks + 1) >> 1;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const bf16* dhr = dh + r * n;
    const bf16* gr = gu + r * 2 * n;
    bf16* dr = dgu + r * 2 * n;
    for (int c0 = blockIdx.x * blockDim.x + threadIdx.x; c0 < half; c0 += gridDim.x * blockDim.x) {
      const int c1 = c0 + half;
      const bool two = c1 < chunks;
      float d[2][8], g[2][8], u[2][8];
      load8(dhr + c0 * 8, d[0]);
      load8(gr + c0 * 8, g[0]);
      load8(gr + n + c0 * 8, u[0]);
      if (two) {
        load8(dhr + c1 * 8, d[1]);
        load8(gr + c1 * 8, g[1]);
        load8(gr + n + c1 * 8, u[1]);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (k == 1 && !two) break;
        const int c = k == 0 ? c0 : c1;
        float dg[8], du[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float fw, bw;
          gelu_tanh_fw_bw(g[k][i], fw, bw);
          du[i] = d[k][i] * bfr(fw);                // d(a*u)/du
          const float da = bfr(d[k][i] * u[k][i]);  // d(a*u)/da, a bf16 tensor in the reference graph
          dg[i] = da * bw;
        }
        store8(dr + c * 8, dg);
        store8(dr + n + c * 8, du);
      }
    }
  }
}

__global__ void geglu_fwd_k(const bf16* __restrict__ gu, bf16* __restrict__ h, int64_t rows, int n) {
  pdl_enter();
  const int chunks = n / 8;
  const int64_t total = rows * chunks;
  GRID_STRIDE(idx, total) {
    const int ch = static_cast<int>(idx % chunks);
    const int64_t r = idx / chunks;
    float g[8], u[8], o[8];
    load8(gu + r * 2 * n + ch * 8, g);
    load8(gu + r * 2 * n + n + ch * 8, u);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = bfr(gelu_tanh_fw(g[i])) * u[i];
    store8(h + r * n + ch * 8, o);
  }
}

__global__ void gelu_bwd_k(const bf16* __restrict__ dact, const bf16* __restrict__ pre, bf16* __restrict__ dpre,
                           int64_t n8) {
  pdl_enter();
  GRID_STRIDE(idx, n8) {
    float d[8], x[8], o[8];
    load8(dact + idx * 8, d);
    load8(pre + idx * 8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = d[i] * gelu_tanh_bw(x[i]);
    store8(dpre + idx * 8, o);
  }
}

// block: 256 threads = 32 column-groups(8 cols) x 8 row lanes; rows slab of 256 per block.
__global__ void __launch_bounds__(256) colsum_bf16_k(const bf16* __restrict__ x, int64_t ld, int64_t rows, int cols,
                                                     float* __restrict__ acc32) {
  pdl_enter();
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + cg) * 8;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 256;
  const int64_t r1 = r0 + 256 < rows ? r0 + 256 : rows;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < cols) {
    int64_t r = r0 + rl;
    if (c0 + 8 <= cols) {
      // four independent 16-byte loads in flight per thread (the single-load loop was latency-bound: 9-32 % of HBM peak)
      for (; r + 24 < r1; r += 32) {
        float v0[8], v1[8], v2[8], v3[8];
        load8(x + r * ld + c0, v0);
        load8(x + (r + 8) * ld + c0, v1);
        load8(x + (r + 16) * ld + c0, v2);
        load8(x + (r + 24) * ld + c0, v3);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += (v0[i] + v1[i]) + (v2[i] + v3[i]);
      }
    }
    for (; r < r1; r += 8) {
      if (c0 + 8 <= cols) {
        float v[8];
        load8(x + r * ld + c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += v[i];
      } else {
        for (int i = 0; i < cols - c0; ++i) a[i] += __bfloat162float(x[r * ld + c0 + i]);
      }
    }
  }
  __shared__ float sm[8][32][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[rl][cg][i] = a[i];
  __syncthreads();
  if (rl == 0 && c0 < cols) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += sm[k][cg][i];
      if (c0 + i < cols) atomicAdd(acc32 + c0 + i, t);
    }
  }
}

__global__ void cast_f32_to_bf16_k(const float* __restrict__ in, bf16* __restrict__ out, int64_t n) {
  pdl_enter();
  GRID_STRIDE(i, n) out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void cast_bf16_to_f32_k(const bf16* __restrict__ in, float* __restrict__ out, int64_t n) {
  pdl_enter();
  GRID_STRIDE(i, n) out[i] = __bfloat162float(in[i]);
}
__global__ void add_bf16_k(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, int64_t n8) {
  pdl_enter();
  GRID_STRIDE(idx, n8) {
    float x[8], y[8];
    load8(a + idx * 8, x);
    load8(b + idx * 8, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] += y[i];
    store8(out + idx * 8, x);
  }
}
__global__ void add_f32_k(float* __restrict__ a, const float* __restrict__ b, int64_t n) {
  pdl_enter();
  GRID_STRIDE(i, n) a[i] += b[i];
}

__global__ void fill_f32_k(float* __restrict__ p, float v, int64_t n) {
  pdl_enter();
  GRID_STRIDE(i, n) p[i] = v;
}
__global__ void decode_times_k(float* __restrict__ out, int n, float dt) {
  pdl_enter();
  float t = 1.0f;
  for (int s = 0; s < n; ++s) {
    out[s] = t;
    t = __fadd_rn(t, dt);
  }
}
__global__ void time_embedding_k(const float* __restrict__ time, const double* __restrict__ scaling,
                                 float* __restrict__ out, int batch, int half) {
  pdl_enter();
  const int64_t total = static_cast<int64_t>(batch) * half;
  GRID_STRIDE(idx, total) {
    const int i = static_cast<int>(idx % half);
    const int b = static_cast<int>(idx / half);
    const double a = scaling[i] * static_cast<double>(time[b]);
    out[static_cast<int64_t>(b) * 2 * half + i] = static_cast<float>(sin(a));
    out[static_cast<int64_t>(b) * 2 * half + half + i] = static_cast<float>(cos(a));
  }
}
__global__ void silu_fwd_k(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  pdl_enter();
  GRID_STRIDE(i, n) {
    const float v = x[i];
    y[i] = v / (1.0f + expf(-v));
  }
}
__global__ void silu_bwd_k(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, int64_t n) {
  pdl_enter();
  GRID_STRIDE(i, n) {
    const float v = x[i];
    const float s = 1.0f / (1.0f + expf(-v));
    dx[i] = dy[i] * s * (1.0f + v * (1.0f - s));
  }
}
__global__ void flow_inputs_k(const float* __restrict__ actions, const float* __restrict__ noise,
                              const float* __restrict__ time, float* __restrict__ x_t, float* __restrict__ u_t,
                              int batch, int per) {
  pdl_enter();
  const int64_t total = static_cast<int64_t>(batch) * per;
  GRID_STRIDE(i, total) {
    const float t = time[i / per];
    const float a = actions[i], e = noise[i];
    x_t[i] = __fadd_rn(__fmul_rn(t, e), __fmul_rn(__fsub_rn(1.0f, t), a));  // P:327
    u_t[i] = __fsub_rn(e, a);                                               // P:328
  }
}
__global__ void flow_loss_k(const float* __restrict__ u_t, const float* __restrict__ v_t, float* __restrict__ loss,
                            int64_t n) {
  pdl_enter();
  GRID_STRIDE(i, n) {
    const float d = __fsub_rn(u_t[i], v_t[i]);
    loss[i] = __fmul_rn(d, d);
  }
}
__global__ void flow_loss_bwd_k(const float* __restrict__ u_t, const float* __restrict__ v_t,
                                const float* __restrict__ dloss, float* __restrict__ dv, int64_t n) {
  pdl_enter();
  GRID_STRIDE(i, n) dv[i] = -2.0f * (u_t[i] - v_t[i]) * dloss[i];
}
__global__ void euler_step_k(float* __restrict__ x, const float* __restrict__ v, float dt, int64_t n) {
  pdl_enter();
  GRID_STRIDE(i, n) x[i] = __fadd_rn(x[i], __fmul_rn(dt, v[i]));
}
__global__ void gather_rows_f32_k(const bf16* __restrict__ in, int64_t in_bstride, int row_off, int T, int width,
                                  float* __restrict__ out, int batch) {
  pdl_enter();
  const int64_t total = static_cast<int64_t>(batch) * T * width;
  GRID_STRIDE(idx, total) {
    const int c = static_cast<int>(idx % width);
    const int64_t bt = idx / width;
    const int b = static_cast<int>(bt / T), t = static_cast<int>(bt % T);
    out[idx] = __bfloat162float(in[b * in_bstride + static_cast<int64_t>(row_off + t) * width + c]);
  }
}
__global__ void copy_rows_bf16_k(const bf16* __restrict__ in, int64_t in_bstride, int row_off, int T, int width,
                                 bf16* __restrict__ out, int batch) {
  pdl_enter();
  const int chunks = width / 8;
  const int64_t total = static_cast<int64_t>(batch) * T * chunks;
  GRID_STRIDE(idx, total) {
    const int ch = static_cast<int>(idx % chunks);
    const int64_t bt = idx / chunks;
    const int b = static_cast<int>(bt / T), t = static_cast<int>(bt % T);
    const uint4 v = *reinterpret_cast<const uint4*>(in + b * in_bstride + static_cast<int64_t>(row_off + t) * width + ch * 8);
    *reinterpret_cast<uint4*>(out + bt * width + ch * 8) = v;
  }
}
__global__ void scatter_rows_bf16_k(const float* __restrict__ in, bf16* __restrict__ out, int64_t out_bstride,
                                    int row_off, int T, int width, int batch) {
  pdl_enter();
  const int64_t total = static_cast<int64_t>(batch) * T * width;
  GRID_STRIDE(idx, total) {
    const int c = static_cast<int>(idx % width);
    const int64_t bt = idx / width;
    const int b = static_cast<int>(bt / T), t = static_cast<int>(bt % T);
    out[b * out_bstride + static_cast<int64_t>(row_off + t) * width + c] = __float2bfloat16_rn(in[idx]);
  }
}

}  // namespace

void embed_tokens_fwd(const int64_t* tok, const bf16* table, bf16* out, int batch, int L, int width, int64_t out_bstride,
                      int row_off, float scale, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(batch) * L * (width / 8);
  launch_pdl(embed_tokens_fwd_k, dim3(grid_for(total)), dim3(256), 0, st, tok, table, out, batch, L, width, out_bstride, row_off, scale); count_launch();
}

void embed_tokens_bwd(const int64_t* tok, const bf16* dout, int64_t dout_bstride, int row_off, bf16* dtable,
                      float* scratch, int* first, int batch, int L, int width, float scale, cudaStream_t st) {
  const int n = batch * L;
  cudaMemsetAsync(scratch, 0, static_cast<size_t>(n) * width * sizeof(float), st);
  launch_pdl(first_occurrence_k, dim3(ceil_div(n, 128)), dim3(128), 0, st, tok, first, n); count_launch();
  const int64_t total = static_cast<int64_t>(n) * width;
  launch_pdl(embed_accum_k, dim3(grid_for(total)), dim3(256), 0, st, tok, first, dout, dout_bstride, row_off, scratch, batch, L, width,
                                                 scale); count_launch();
  launch_pdl(embed_write_k, dim3(grid_for(total)), dim3(256), 0, st, tok, first, scratch, dtable, n, width); count_launch();
}

void geglu_bwd(const bf16* dh, const bf16* gu, bf16* dgu, int64_t rows, int n, cudaStream_t st) {
  const int half = (n / 8 + 1) / 2;
  const int gx = ceil_div(half, 256);
  int64_t gy = (148 * 8 + gx - 1) / gx;
  if (gy > rows) gy = rows;
  launch_pdl(geglu_bwd_k, dim3(dim3(gx, static_cast<unsigned>(gy))), dim3(256), 0, st, dh, gu, dgu, rows, n); count_launch();
}
void geglu_fwd(const bf16* gu, bf16* h, int64_t rows, int n, cudaStream_t st) {
  launch_pdl(geglu_fwd_k, dim3(grid_for(rows * (n / 8))), dim3(256), 0, st, gu, h, rows, n); count_launch();
}
void gelu_bwd(const bf16* dact, const bf16* pre, bf16* dpre, int64_t n, cudaStream_t st) {
  launch_pdl(gelu_bwd_k, dim3(grid_for(n / 8)), dim3(256), 0, st, dact, pre, dpre, n / 8); count_launch();
}
void colsum_bf16(const bf16* x, int64_t ld, int64_t rows, int cols, float* acc32, cudaStream_t st) {
  dim3 grid(ceil_div(cols, 256), ceil_div(rows, 256));
  launch_pdl(colsum_bf16_k, dim3(grid), dim3(256), 0, st, x, ld, rows, cols, acc32); count_launch();
}
void cast_f32_to_bf16(const float* in, bf16* out, int64_t n, cudaStream_t st) {
  launch_pdl(cast_f32_to_bf16_k, dim3(grid_for(n)), dim3(256), 0, st, in, out, n); count_launch();
}
void cast_bf16_to_f32(const bf16* in, float* out, int64_t n, cudaStream_t st) {
  launch_pdl(cast_bf16_to_f32_k, dim3(grid_for(n)), dim3(256), 0, st, in, out, n); count_launch();
}
void add_bf16(const bf16* a, const bf16* b, bf16* out, int64_t n, cudaStream_t st) {
  launch_pdl(add_bf16_k, dim3(grid_for(n / 8)), dim3(256), 0, st, a, b, out, n / 8); count_launch();
}
void add_f32(float* a, const float* b, int64_t n, cudaStream_t st) { launch_pdl(add_f32_k, dim3(grid_for(n)), dim3(256), 0, st, a, b, n); count_launch(); }
void fill_zero(void* p, size_t bytes, cudaStream_t st) { cudaMemsetAsync(p, 0, bytes, st); }
void fill_f32(float* p, float v, int64_t n, cudaStream_t st) { launch_pdl(fill_f32_k, dim3(grid_for(n)), dim3(256), 0, st, p, v, n); count_launch(); }
void decode_times(float* out, int n, float dt, cudaStream_t st) {
  launch_pdl(decode_times_k, dim3(1), dim3(1), 0, st, out, n, dt); count_launch();
}
void time_embedding(const float* time, const double* scaling, float* out, int batch, int half, cudaStream_t st) {
  launch_pdl(time_embedding_k, dim3(grid_for(static_cast<int64_t>(batch) * half)), dim3(256), 0, st, time, scaling, out, batch, half); count_launch();
}
void silu_fwd(const float* x, float* y, int64_t n, cudaStream_t st) { launch_pdl(silu_fwd_k, dim3(grid_for(n)), dim3(256), 0, st, x, y, n); count_launch(); }
void silu_bwd(const float* dy, const float* x, float* dx, int64_t n, cudaStream_t st) {
  launch_pdl(silu_bwd_k, dim3(grid_for(n)), dim3(256), 0, st, dy, x, dx, n); count_launch();
}
void flow_inputs(const float* actions, const float* noise, const float* time, float* x_t, float* u_t, int batch, int per,
                 cudaStream_t st) {
  launch_pdl(flow_inputs_k, dim3(grid_for(static_cast<int64_t>(batch) * per)), dim3(256), 0, st, actions, noise, time, x_t, u_t, batch, per); count_launch();
}
void flow_loss(const float* u_t, const float* v_t, float* loss, int64_t n, cudaStream_t st) {
  launch_pdl(flow_loss_k, dim3(grid_for(n)), dim3(256), 0, st, u_t, v_t, loss, n); count_launch();
}
void flow_loss_bwd(const float* u_t, const float* v_t, const float* dloss, float* dv, int64_t n, cudaStream_t st) {
  launch_pdl(flow_loss_bwd_k, dim3(grid_for(n)), dim3(256), 0, st, u_t, v_t, dloss, dv, n); count_launch();
}
void euler_step(float* x, const float* v, float dt, int64_t n, cudaStream_t st) {
  launch_pdl(euler_step_k, dim3(grid_for(n)), dim3(256), 0, st, x, v, dt, n); count_launch();
}
void gather_rows_f32(const bf16* in, int64_t in_bstride, int row_off, int T, int width, float* out, int batch,
                     cudaStream_t st) {
  launch_pdl(gather_rows_f32_k, dim3(grid_for(static_cast<int64_t>(batch) * T * width)), dim3(256), 0, st, in, in_bstride, row_off, T, width,
                                                                                       out, batch); count_launch();
}
void copy_rows_bf16(const bf16* in, int64_t in_bstride, int row_off, int T, int width, bf16* out, int batch,
                    cudaStream_t st) {
  launch_pdl(copy_rows_bf16_k, dim3(grid_for(static_cast<int64_t>(batch) * T * (width / 8))), dim3(256), 0, st, in, in_bstride, row_off, T,
                                                                                            width, out, batch); count_launch();
}
void scatter_rows_bf16(const float* in, bf16* out, int64_t out_bstride, int row_off, int T, int width, int batch,
                       cudaStream_t st) {
  launch_pdl(scatter_rows_bf16_k, dim3(grid_for(static_cast<int64_t>(batch) * T * width)), dim3(256), 0, st, in, out, out_bstride, row_off,
                                                                                         T, width, batch); count_launch();
}

}  // namespace pi05
