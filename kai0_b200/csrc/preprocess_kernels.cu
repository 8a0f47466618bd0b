// Observation preprocessing on the device (preprocessing_pytorch.py:35-148, image_tools.py:55-126): aspect-preserving
// bilinear resize + pad, and the train-time augmentation (95 % crop -> bilinear resize, small rotation by grid_sample,
// brightness / contrast / saturation) fused into two passes over each image instead of ~15 ATen kernels:
//   pass 1: geometry -> [0,1] image in scratch + per-block partial sums of (image * brightness)
//   pass 2: per-sample mean from the partials (fixed order), colour transform, back to [-1,1], NCHW out.
// The arithmetic follows ATen's upsample_bilinear2d (align_corners=False) and grid_sampler_2d (bilinear, zeros).
#include <algorithm>

#include "common.cuh"
#include "errors.h"
#include "kernels.h"
#include "launch.h"

namespace pi05 {
namespace {

struct Img {
  const void* p;
  int h, w;
  int cl;  // 1: [B,h,w,3]; 0: [B,3,h,w]
  int u8;  // 1: uint8 pixels, converted on the fly exactly as Observation.from_dict does (models/model.py:129-133:
           //    img.astype(float32) / 255.0 * 2.0 - 1.0, three separately rounded fp32 operations); 0: fp32 in [-1,1]
  __device__ __forceinline__ float at(int b, int c, int y, int x) const {
    const int64_t base = static_cast<int64_t>(b) * 3 * h * w;
    const int64_t idx = cl ? base + (static_cast<int64_t>(y) * w + x) * 3 + c : base + (static_cast<int64_t>(c) * h + y) * w + x;
    if (u8) {
      const float v = static_cast<float>(static_cast<const uint8_t*>(p)[idx]);
      return __fsub_rn(__fmul_rn(__fdiv_rn(v, 255.0f), 2.0f), 1.0f);
    }
    return static_cast<const float*>(p)[idx];
  }
};

// Where a preprocessed pixel goes: the fp32 NCHW image (rows == nullptr), or straight into the operand of the patch-embedding
// GEMM: bf16 rows [batch * T, 3 * Kp], row = b * T + patch index, column = c * p * p + (y % p) * p + (x % p)
// (the Conv2d weight's [3, p, p] flattening, modeling_siglip.py:220-226), stored as the split v = hi + lo with
// hi = bf16(v), lo = bf16(v - hi) in three column blocks [hi | lo | hi] (see engine.cu::vision_forward for why).
struct Sink {
  float* nchw;
  bf16* rows;
  int S, p, P, Kp;
  __device__ __forceinline__ void put(int b, int c, int y, int x, float v) const {
    if (rows == nullptr) {
      nchw[((static_cast<int64_t>(b) * 3 + c) * S + y) * S + x] = v;
      return;
    }
    const int64_t row = static_cast<int64_t>(b) * P * P + (y / p) * P + (x / p);
    const int col = c * p * p + (y % p) * p + (x % p);
    const bf16 hi = __float2bfloat16_rn(v);
    const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    bf16* r = rows + row * (3 * Kp) + col;
    r[0] = hi;
    r[Kp] = lo;
    r[2 * Kp] = hi;
  }
};

// ATen area_pixel_compute_source_index (align_corners = false, cubic = false)
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = scale * (dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - i0;
  l0 = 1.f - l1;
}

// image_tools.py:55-126 for fp32 input: out NCHW [B,3,S,S]
__global__ void resize_pad_k(Img src, int S, int rh, int rw, int ph0, int pw0, Sink out, int B) {
  pdl_enter();
  const int64_t total = static_cast<int64_t>(B) * 3 * S * S;
  const float sh = static_cast<float>(src.h) / rh, sw = static_cast<float>(src.w) / rw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % S), y = static_cast<int>((i / S) % S);
    const int c = static_cast<int>((i / (static_cast<int64_t>(S) * S)) % 3), b = static_cast<int>(i / (3LL * S * S));
    const int ry = y - ph0, rx = x - pw0;
    float v = -1.0f;  // padding value for float images
    if (ry >= 0 && ry < rh && rx >= 0 && rx < rw) {
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      src_index(sh, ry, src.h, y0, y1, ly0, ly1);
      src_index(sw, rx, src.w, x0, x1, lx0, lx1);
      v = ly0 * (lx0 * src.at(b, c, y0, x0) + lx1 * src.at(b, c, y0, x1)) +
          ly1 * (lx0 * src.at(b, c, y1, x0) + lx1 * src.at(b, c, y1, x1));
      v = fminf(fmaxf(v, -1.0f), 1.0f);
    }
    out.put(b, c, y, x, v);
  }
}

// plain layout change to NCHW (no resize, no augmentation)
__global__ void to_nchw_k(Img src, Sink out, int B) {
  pdl_enter();
  const int S = src.h;
  const int64_t total = static_cast<int64_t>(B) * 3 * S * S;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % S), y = static_cast<int>((i / S) % S);
    const int c = static_cast<int>((i / (static_cast<int64_t>(S) * S)) % 3), b = static_cast<int>(i / (3LL * S * S));
    out.put(b, c, y, x, src.at(b, c, y, x));
  }
}

struct Geo {
  int S, crop_h, crop_w, start_h, start_w;
  float sch, scw;  // crop_h / S, crop_w / S
  // value of the cropped-and-resized [0,1] image at integer pixel (ty, tx)  (preprocessing_pytorch.py:62-82)
  __device__ __forceinline__ float crop_resize(const Img& im, int b, int c, int ty, int tx) const {
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    src_index(sch, ty, crop_h, y0, y1, ly0, ly1);
    src_index(scw, tx, crop_w, x0, x1, lx0, lx1);
    const float v00 = im.at(b, c, start_h + y0, start_w + x0) * 0.5f + 0.5f;
    const float v01 = im.at(b, c, start_h + y0, start_w + x1) * 0.5f + 0.5f;
    const float v10 = im.at(b, c, start_h + y1, start_w + x0) * 0.5f + 0.5f;
    const float v11 = im.at(b, c, start_h + y1, start_w + x1) * 0.5f + 0.5f;
    return ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
  }
};

// torch.linspace(-1, 1, n)[i] in fp32 (two-sided evaluation as ATen does)
__device__ __forceinline__ float linspace_pm1(int i, int n) {
  const float step = 2.0f / static_cast<float>(n - 1);
  return (i < n / 2) ? (-1.0f + step * i) : (1.0f - step * (n - 1 - i));
}

constexpr int kGeoBlocks = 32;  // blocks per sample in pass 1 (= partial sums per sample)

// pass 1.  geometric = 1: crop/resize + rotation (non-wrist cameras); 0: just x/2 + 0.5.
__global__ void __launch_bounds__(256) augment_geo_k(Img im, const float* __restrict__ params, int geometric,
                                                     float* __restrict__ tmp, float* __restrict__ partial) {
  pdl_enter();
  const int S = im.h, b = blockIdx.y;
  const float bright = params[3];
  Geo g;
  g.S = S;
  g.crop_h = static_cast<int>(S * 0.95);
  g.crop_w = g.crop_h;
  const bool crop = (S - g.crop_h) > 0;
  if (!crop) g.crop_h = g.crop_w = S;
  g.start_h = crop ? static_cast<int>(params[0]) : 0;
  g.start_w = crop ? static_cast<int>(params[1]) : 0;
  g.sch = static_cast<float>(g.crop_h) / S;
  g.scw = static_cast<float>(g.crop_w) / S;
  const float angle = params[2];
  const bool rotate = geometric && fabsf(angle) > 0.1f;
  float cs = 1.f, sn = 0.f;
  if (rotate) {
    const float rad = angle * 3.14159265358979323846f / 180.0f;
    cs = cosf(rad);
    sn = sinf(rad);
  }
  float acc = 0.f;
  const int npix = S * S;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    const int y = pix / S, x = pix % S;
    float out[3];
    if (!geometric) {
#pragma unroll
      for (int c = 0; c < 3; ++c) out[c] = im.at(b, c, y, x) * 0.5f + 0.5f;
    } else if (!rotate) {
#pragma unroll
      for (int c = 0; c < 3; ++c) out[c] = g.crop_resize(im, b, c, y, x);
    } else {
      // grid_sample(bilinear, zeros, align_corners=False) at the rotated normalised coordinates
      const float gx = linspace_pm1(x, S), gy = linspace_pm1(y, S);
      // separate roundings as in the reference's tensor expressions (no FMA contraction on the coordinates)
      const float rx = __fsub_rn(__fmul_rn(gx, cs), __fmul_rn(gy, sn));
      const float ry = __fadd_rn(__fmul_rn(gx, sn), __fmul_rn(gy, cs));
      const float ix = __fsub_rn(__fmul_rn(__fadd_rn(rx, 1.f), static_cast<float>(S)), 1.f) / 2.f;
      const float iy = __fsub_rn(__fmul_rn(__fadd_rn(ry, 1.f), static_cast<float>(S)), 1.f) / 2.f;
      const float fx = floorf(ix), fy = floorf(iy);
      const int x_w = static_cast<int>(fx), y_n = static_cast<int>(fy);
      const int x_e = x_w + 1, y_s = y_n + 1;
      const float nw = (x_e - ix) * (y_s - iy), ne = (ix - x_w) * (y_s - iy);
      const float sw = (x_e - ix) * (iy - y_n), se = (ix - x_w) * (iy - y_n);
      const bool in_n = y_n >= 0 && y_n < S, in_s = y_s >= 0 && y_s < S;
      const bool in_w = x_w >= 0 && x_w < S, in_e = x_e >= 0 && x_e < S;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = 0.f;
        if (in_n && in_w) v += g.crop_resize(im, b, c, y_n, x_w) * nw;
        if (in_n && in_e) v += g.crop_resize(im, b, c, y_n, x_e) * ne;
        if (in_s && in_w) v += g.crop_resize(im, b, c, y_s, x_w) * sw;
        if (in_s && in_e) v += g.crop_resize(im, b, c, y_s, x_e) * se;
        out[c] = v;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      tmp[(static_cast<int64_t>(b) * 3 + c) * npix + pix] = out[c];
      acc += out[c] * bright;
    }
  }
  __shared__ float sh[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    partial[b * kGeoBlocks + blockIdx.x] = t;
  }
}

// pass 2: colour (preprocessing_pytorch.py:122-146)
__global__ void __launch_bounds__(256) augment_colour_k(const float* __restrict__ tmp, const float* __restrict__ partial,
                                                        const float* __restrict__ params, int S, Sink out) {
  pdl_enter();
  const int b = blockIdx.y, npix = S * S;
  const float bright = params[3], contrast = params[4], satur = params[5];
  float tot = 0.f;
  for (int i = 0; i < kGeoBlocks; ++i) tot += partial[b * kGeoBlocks + i];
  const float mean = tot / static_cast<float>(3 * npix);
  const float* in = tmp + static_cast<int64_t>(b) * 3 * npix;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = in[c * npix + pix] * bright;
      v[c] = (x - mean) * contrast + mean;
    }
    const float gray = (v[0] + v[1] + v[2]) / 3.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x = gray + (v[c] - gray) * satur;
      x = fminf(fmaxf(x, 0.f), 1.f);
      out.put(b, c, pix / S, pix % S, x * 2.0f - 1.0f);
    }
  }
}

inline int blocks_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  return static_cast<int>(g < 148 * 8 ? g : 148 * 8);
}

}  // namespace

size_t preprocess_scratch_floats(int batch, int out_size) {
  return static_cast<size_t>(2) * batch * 3 * out_size * out_size + static_cast<size_t>(batch) * kGeoBlocks;
}

static void preprocess_any(Img src, int batch, int out_size, int train, int geometric, const float* params,
                           float* scratch, Sink out, cudaStream_t st) {
  const int S = out_size, height = src.h, width = src.w;
  const int64_t n = static_cast<int64_t>(batch) * 3 * S * S;
  float* resized = scratch;      // [B,3,S,S]
  float* tmp = scratch + n;      // [B,3,S,S]
  float* partial = scratch + 2 * n;
  Img cur = src;
  if (height != S || width != S) {
    const double ratio = std::max(static_cast<double>(width) / S, static_cast<double>(height) / S);  // image_tools.py:88
    const int rh = static_cast<int>(height / ratio), rw = static_cast<int>(width / ratio);
    const int ph0 = (S - rh) / 2, pw0 = (S - rw) / 2;
    const Sink dst = train ? Sink{resized, nullptr, S, 1, 1, 0} : out;
    launch_pdl(resize_pad_k, dim3(blocks_for(n)), dim3(256), 0, st, src, S, rh, rw, ph0, pw0, dst, batch);
    count_launch();
    if (!train) return;
    cur = Img{resized, S, S, 0, 0};
  } else if (!train) {
    launch_pdl(to_nchw_k, dim3(blocks_for(n)), dim3(256), 0, st, src, out, batch);
    count_launch();
    return;
  }
  dim3 grid(kGeoBlocks, batch);
  launch_pdl(augment_geo_k, dim3(grid), dim3(256), 0, st, cur, params, geometric, tmp, partial);
  count_launch();
  launch_pdl(augment_colour_k, dim3(grid), dim3(256), 0, st, tmp, partial, params, S, out);
  count_launch();
}

void preprocess_image(const float* data, int height, int width, int channels_last, int batch, int out_size, int train,
                      int geometric, const float* params, float* scratch, float* out, cudaStream_t st) {
  preprocess_any(Img{data, height, width, channels_last, 0}, batch, out_size, train, geometric, params, scratch,
                 Sink{out, nullptr, out_size, 1, 1, 0}, st);
}

void preprocess_patches(const void* data, int is_u8, int height, int width, int channels_last, int batch, int out_size,
                        int patch, int train, int geometric, const float* params, float* scratch, bf16* rows,
                        cudaStream_t st) {
  const int Kp = patch_row_kp(patch);
  preprocess_any(Img{data, height, width, channels_last, is_u8}, batch, out_size, train, geometric, params, scratch,
                 Sink{nullptr, rows, out_size, patch, out_size / patch, Kp}, st);
}

}  // namespace pi05
