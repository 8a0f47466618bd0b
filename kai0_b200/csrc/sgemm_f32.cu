// fp32 SIMT GEMMs for the small fp32 linears of the path (action_in/out_proj, time MLP, adaRMS dense:
// pi0_pytorch.py:270-273,289-293,368-371; modeling_gemma.py:88) and the fp32 patch-embedding convolution
// expressed as a GEMM over an implicit im2col (modeling_siglip.py:220-226,271-282).
//
// C[i,j] (+)= sum_k A(i,k) * B(j,k) with operand access through small loader functors; 64x64x16 tiles,
// 256 threads, 4x4 register micro-tiles.  These are <1% of the step's FLOPs and must stay fp32-exact.
#include "common.cuh"
#include "errors.h"
#include "kernels.h"
#include "launch.h"

namespace pi05 {

namespace {

struct LinearF32 {  // element (r,k) at p[r*sr + k*sk]; batch z at p + z*sz
  const float* p;
  int64_t sr, sk;
  int64_t sz = 0;
  __device__ __forceinline__ LinearF32 at(int z) const { return LinearF32{p + z * sz, sr, sk, sz}; }
  __device__ __forceinline__ float operator()(int r, int k) const { return p[r * sr + k * sk]; }
  __device__ __forceinline__ bool k_fast() const { return sk == 1; }  // which index is contiguous in memory
};
struct LinearBF16 {
  const bf16* p;
  int64_t sr, sk;
  __device__ __forceinline__ LinearBF16 at(int) const { return *this; }
  __device__ __forceinline__ float operator()(int r, int k) const { return __bfloat162float(p[r * sr + k * sk]); }
  __device__ __forceinline__ bool k_fast() const { return sk == 1; }
};
struct Im2col {  // row = img*P*P + prow*P + pcol ; k = ch*p*p + py*p + px
  const float* img;
  int S, p, P;  // image size, patch, patches per side
  __device__ __forceinline__ Im2col at(int) const { return *this; }
  __device__ __forceinline__ float operator()(int r, int k) const {
    const int pp = p * p;
    const int im = r / (P * P), pr = (r / P) % P, pc = r % P;
    const int ch = k / pp, py = (k % pp) / p, px = k % p;
    return img[(static_cast<int64_t>(im) * 3 + ch) * S * S + static_cast<int64_t>(pr * p + py) * S + pc * p + px];
  }
  __device__ __forceinline__ bool k_fast() const { return true; }
};
struct Im2colT {  // transposed roles: "row" = k-feature index, "k" = patch-row index  (for wgrad: B(j=feature, kk=row))
  Im2col base;
  __device__ __forceinline__ Im2colT at(int) const { return *this; }
  __device__ __forceinline__ float operator()(int feat, int row) const { return base(row, feat); }
  __device__ __forceinline__ bool k_fast() const { return false; }
};

struct EpiF32 {  // C fp32 [M,N] row-major (+bias[j]) ; accumulate / atomic variants
  float* C;
  int64_t ldc;
  const float* bias;
  int mode;  // 0 store, 1 accumulate (+=), 2 atomicAdd
  int64_t cz = 0, bz = 0;  // batch strides of C and bias
  __device__ __forceinline__ EpiF32 at(int z) const {
    return EpiF32{C + z * cz, ldc, bias ? bias + z * bz : nullptr, mode, cz, bz};
  }
  __device__ __forceinline__ void operator()(int i, int j, float v, bool first) const {
    if (bias && first) v = __fadd_rn(v, bias[j]);
    float* c = C + i * ldc + j;
    if (mode == 0) *c = v;
    else if (mode == 1) *c += v;
    else atomicAdd(c, v);
  }
};
struct EpiPatch {  // out bf16 [rows, width] = bf( (acc + bias[c]) + pos[patch, c] )
  bf16* out;
  int width, patches;
  const float* bias;
  const float* pos;
  __device__ __forceinline__ EpiPatch at(int) const { return *this; }
  __device__ __forceinline__ void operator()(int i, int j, float v, bool) const {
    v = __fadd_rn(v, bias[j]);
    v = __fadd_rn(v, pos[static_cast<int64_t>(i % patches) * width + j]);
    out[static_cast<int64_t>(i) * width + j] = __float2bfloat16_rn(v);
  }
};

constexpr int TM = 64, TN = 64, TK = 16;

template <class LA, class LB, class EPI>
__global__ void __launch_bounds__(256) sgemm_k(LA la_, LB lb_, EPI epi_, int M, int N, int K, int k_per_split,
                                               int batched) {
  pdl_enter();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  // blockIdx.z is either a split-K slice (atomic epilogue) or, in batched mode, an independent problem
  const LA la = batched ? la_.at(blockIdx.z) : la_;
  const LB lb = batched ? lb_.at(blockIdx.z) : lb_;
  const EPI epi = batched ? epi_.at(blockIdx.z) : epi_;
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  const int kbeg = batched ? 0 : blockIdx.z * k_per_split;
  const int kend = batched ? K : min(K, kbeg + k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
    // consecutive threads walk whichever index is contiguous in memory for that operand (coalesced tile loads)
    const bool ak = la.k_fast(), bk = lb.k_fast();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = threadIdx.x + r * 256;  // 0..1023
      {
        const int kk = ak ? (e & 15) : (e >> 6), ii = ak ? (e >> 4) : (e & 63);
        const int gi = i0 + ii, gk = k0 + kk;
        As[kk][ii] = (gi < M && gk < kend) ? la(gi, gk) : 0.f;
      }
      {
        const int kk = bk ? (e & 15) : (e >> 6), ii = bk ? (e >> 4) : (e & 63);
        const int gj = j0 + ii, gk = k0 + kk;
        Bs[kk][ii] = (gj < N && gk < kend) ? lb(gj, gk) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        a[x] = As[kk][ty * 4 + x];
        b[x] = Bs[kk][tx * 4 + x];
      }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int i = i0 + ty * 4 + x, j = j0 + tx * 4 + y;
      if (i < M && j < N) epi(i, j, acc[x][y], batched || blockIdx.z == 0);
    }
}

template <class LA, class LB, class EPI>
void run(LA la, LB lb, EPI epi, int M, int N, int K, int splits, cudaStream_t st) {
  if (splits < 1) splits = 1;
  int kps = (K + splits - 1) / splits;
  kps = ((kps + TK - 1) / TK) * TK;
  splits = (K + kps - 1) / kps;
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, splits);
  launch_pdl(sgemm_k<LA, LB, EPI>, dim3(grid), dim3(256), 0, st, la, lb, epi, M, N, K, kps, 0); count_launch();
}

template <class LA, class LB, class EPI>
void run_batched(LA la, LB lb, EPI epi, int M, int N, int K, int batch, cudaStream_t st) {
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, batch);
  launch_pdl(sgemm_k<LA, LB, EPI>, dim3(grid), dim3(256), 0, st, la, lb, epi, M, N, K, K, 1); count_launch();
}

__global__ void reduce_batches_k(const float* __restrict__ part, float* __restrict__ out, int64_t n, int batch) {
  pdl_enter();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < batch; ++z) s += part[z * n + i];  // fixed order: deterministic
    out[i] = s;
  }
}

__global__ void colsum_f32_k(const float* __restrict__ x, int M, int N, float* __restrict__ out) {
  pdl_enter();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  float s = 0.f;
  for (int i = 0; i < M; ++i) s += x[static_cast<int64_t>(i) * N + j];
  out[j] = s;
}

// dpos[patch, c] = sum_img dout[img, patch, c]; dbias[c] = sum_{img,patch} dout
__global__ void patch_dpos_k(const bf16* __restrict__ dout, float* __restrict__ dpos, int n_img, int patches, int width) {
  pdl_enter();
  const int64_t total = static_cast<int64_t>(patches) * width;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int im = 0; im < n_img; ++im) s += __bfloat162float(dout[im * total + idx]);
    dpos[idx] = s;
  }
}

}  // namespace

// Few outputs, long K (decode: action_out_proj is [50,1024] x [32,1024]^T): one warp per output element, 16-byte
// loads along K, shuffle reduction in a fixed order.  The tiled kernel would run this as ONE block with 64 k-steps.
__global__ void __launch_bounds__(256) linear_f32_small_k(const float* __restrict__ X, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ Y, int M,
                                                          int N, int K) {
  pdl_enter();
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= M * N) return;
  const int m = o / N, n = o % N;
  const float4* x = reinterpret_cast<const float4*>(X + static_cast<int64_t>(m) * K);
  const float4* w = reinterpret_cast<const float4*>(W + static_cast<int64_t>(n) * K);
  float acc = 0.f;
  for (int k = lane; k < K / 4; k += 32) {
    const float4 a = x[k], b = w[k];
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) Y[o] = acc + (bias ? bias[n] : 0.f);
}

void linear_f32(const float* X, const float* W, const float* bias, float* Y, int M, int N, int K, cudaStream_t st) {
  if (static_cast<int64_t>(M) * N <= 4096 && K >= 256 && K % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    launch_pdl(linear_f32_small_k, dim3((M * N + 7) / 8), dim3(256), 0, st, X, W, bias, Y, M, N, K);
    count_launch();
    return;
  }
  run(LinearF32{X, K, 1}, LinearF32{W, K, 1}, EpiF32{Y, N, bias, 0}, M, N, K, 1, st);
}

void linear_f32_dgrad(const float* dY, const float* W, float* dX, int M, int N, int K, int accumulate, cudaStream_t st) {
  // dX[i, k] = sum_n dY[i, n] W[n, k]  -> A(i, n) = dY, B(k, n) = W[n*K + k]
  run(LinearF32{dY, N, 1}, LinearF32{W, 1, K}, EpiF32{dX, K, nullptr, accumulate ? 1 : 0}, M, K, N, 1, st);
}

void linear_f32_wgrad(const float* dY, const float* X, float* dW, float* db, int M, int N, int K, cudaStream_t st) {
  // dW[n, k] = sum_i dY[i, n] X[i, k] -> A(n, i) = dY[i*N + n], B(k, i) = X[i*K + k]
  run(LinearF32{dY, 1, N}, LinearF32{X, 1, K}, EpiF32{dW, K, nullptr, 0}, N, K, M, 1, st);
  if (db) launch_pdl(colsum_f32_k, dim3((N + 127) / 128), dim3(128), 0, st, dY, M, N, db); count_launch();
}

// `batch` independent linears that share X: Y[z] = X W[z]^T + bias[z]  (all 37 adaRMS modulation layers in one launch)
void linear_f32_batched(const float* X, const float* W, const float* bias, float* Y, int M, int N, int K, int batch,
                        int64_t w_stride, int64_t b_stride, int64_t y_stride, cudaStream_t st) {
  run_batched(LinearF32{X, K, 1, 0}, LinearF32{W, K, 1, w_stride}, EpiF32{Y, N, bias, 0, y_stride, b_stride}, M, N, K,
              batch, st);
}
// dW[z] = dY[z]^T X ; db[z] = colsum(dY[z])
void linear_f32_wgrad_batched(const float* dY, const float* X, float* dW, float* db, int M, int N, int K, int batch,
                              int64_t dy_stride, int64_t w_stride, int64_t b_stride, cudaStream_t st) {
  run_batched(LinearF32{dY, 1, N, dy_stride}, LinearF32{X, 1, K, 0}, EpiF32{dW, K, nullptr, 0, w_stride, 0}, N, K, M,
              batch, st);
  for (int z = 0; z < batch; ++z) {
    launch_pdl(colsum_f32_k, dim3((N + 127) / 128), dim3(128), 0, st, dY + z * dy_stride, M, N, db + z * b_stride);
    count_launch();
  }
}
// dX = sum_z dY[z] W[z]: per-z partials into `scratch` [batch, M, K], then a fixed-order reduction (deterministic)
void linear_f32_dgrad_batched_sum(const float* dY, const float* W, float* dX, float* scratch, int M, int N, int K,
                                  int batch, int64_t dy_stride, int64_t w_stride, cudaStream_t st) {
  run_batched(LinearF32{dY, N, 1, dy_stride}, LinearF32{W, 1, K, w_stride},
              EpiF32{scratch, K, nullptr, 0, static_cast<int64_t>(M) * K, 0}, M, K, N, batch, st);
  const int64_t n = static_cast<int64_t>(M) * K;
  launch_pdl(reduce_batches_k, dim3(static_cast<int>((n + 255) / 256)), dim3(256), 0, st, scratch, dX, n, batch);
  count_launch();
}

void patch_embed_fwd(const float* images, const float* W, const float* bias, const float* pos, bf16* out, int n_img,
                     int image_size, int patch, int width, cudaStream_t st) {
  const int P = image_size / patch;
  const int K = 3 * patch * patch;
  run(Im2col{images, image_size, patch, P}, LinearF32{W, K, 1}, EpiPatch{out, width, P * P, bias, pos}, n_img * P * P,
      width, K, 1, st);
}

void patch_embed_bwd(const float* images, const bf16* dout, float* dW, float* dbias, float* dpos, float* scratch,
                     int n_img, int image_size, int patch, int width, cudaStream_t st) {
  (void)scratch;
  const int P = image_size / patch;
  const int K = 3 * patch * patch;
  const int rows = n_img * P * P;
  // dW[c, f] = sum_row dout[row, c] * im2col(row, f)
  cudaMemsetAsync(dW, 0, static_cast<size_t>(width) * K * sizeof(float), st);
  const int splits = rows >= 4096 ? 16 : 1;
  run(LinearBF16{dout, 1, width}, Im2colT{Im2col{images, image_size, patch, P}}, EpiF32{dW, K, nullptr, 2}, width, K,
      rows, splits, st);
  patch_embed_bwd_pos_bias(dout, dbias, dpos, n_img, P * P, width, st);
}

// d(position embedding)[t, c] = sum over images of dout[img, t, c]; d(bias)[c] = column sum of that
void patch_embed_bwd_pos_bias(const bf16* dout, float* dbias, float* dpos, int n_img, int tokens, int width,
                              cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(tokens) * width;
  launch_pdl(patch_dpos_k, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, st, dout, dpos, n_img, tokens, width); count_launch();
  launch_pdl(colsum_f32_k, dim3((width + 127) / 128), dim3(128), 0, st, dpos, tokens, width, dbias); count_launch();
}

namespace {
// fp32 weight [rows, k] -> bf16 [rows, 3 * kp] = [hi | hi | lo] with hi = bf16(w), lo = bf16(w - hi); padding columns zero
__global__ void __launch_bounds__(256) split_weight_k(const float* __restrict__ w, bf16* __restrict__ out, int rows, int k,
                                                      int kp) {
  pdl_enter();
  const int64_t total = static_cast<int64_t>(rows) * kp;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / kp), c = static_cast<int>(i % kp);
    bf16 hi = __float2bfloat16_rn(0.f), lo = hi;
    if (c < k) {
      const float v = w[static_cast<int64_t>(r) * k + c];
      hi = __float2bfloat16_rn(v);
      lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
    bf16* o = out + static_cast<int64_t>(r) * 3 * kp + c;
    o[0] = hi;
    o[kp] = hi;
    o[2 * kp] = lo;
  }
}
// dW[r, c] = d[r, c] + d[r, kp + c]   (the two column blocks [hi | lo] of the split input)
__global__ void __launch_bounds__(256) fold_patch_dw_k(const float* __restrict__ d, float* __restrict__ dw, int rows, int k,
                                                       int kp) {
  pdl_enter();
  const int64_t total = static_cast<int64_t>(rows) * k;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / k), c = static_cast<int>(i % k);
    const float* s = d + static_cast<int64_t>(r) * 2 * kp + c;
    dw[i] = s[0] + s[kp];
  }
}
}  // namespace

void split_patch_weight(const float* w, bf16* out, int rows, int k, int kp, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(rows) * kp;
  launch_pdl(split_weight_k, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, st, w, out, rows, k, kp); count_launch();
}
void fold_patch_dw(const float* d, float* dw, int rows, int k, int kp, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(rows) * k;
  launch_pdl(fold_patch_dw_k, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, st, d, dw, rows, k, kp); count_launch();
}

}  // namespace pi05
