// Layout / precision conversion kernels: NCHW fp32 <-> NHWC bf16, conv weight shadows,
// and the space-to-depth staging of the ResNet stem input.
#include "common.h"

#include <string.h>

namespace ecgmm {

// [N][C][P] fp32 -> [N][P][C] bf16 through a 32x32 shared-memory tile (P = H*W).
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, int P) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xn = x + (size_t)n * C * P;
  __nv_bfloat16* yn = y + (size_t)n * C * P;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, pp = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && pp < P) ? xn[(size_t)c * P + pp] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, c = c0 + threadIdx.x;
    if (c < C && pp < P) yn[(size_t)pp * C + c] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int C, int P) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const __nv_bfloat16* xn = x + (size_t)n * C * P;
  float* yn = y + (size_t)n * C * P;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && pp < P) ? __bfloat162float(xn[(size_t)pp * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, pp = p0 + threadIdx.x;
    if (c < C && pp < P) yn[(size_t)c * P + pp] = tile[threadIdx.x][i];
  }
}

__global__ void conv_weight_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w_fwd,
                                        __nv_bfloat16* __restrict__ w_dgrad, int O, int I, int RS) {
  const size_t total = (size_t)O * I * RS;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int t = idx % RS;
    const int i = (idx / RS) % I;
    const int o = idx / ((size_t)RS * I);
    const __nv_bfloat16 v = __float2bfloat16(w[idx]);
    if (w_fwd) w_fwd[((size_t)o * RS + t) * I + i] = v;
    if (w_dgrad) w_dgrad[((size_t)i * RS + t) * O + o] = v;
  }
}

// Stem weights [64][3][7][7] -> [64][ra(4)][sa(4)][ch16] with ch16 = (dr*2+ds)*3 + c,
// r = 2*ra+dr, s = 2*sa+ds (taps with r == 7 or s == 7 and channels 12..15 are zero).
__global__ void stem_weight_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ ws) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 256) return;
  const int ch = idx & 15, sa = (idx >> 4) & 3, ra = (idx >> 6) & 3, o = idx >> 8;
  float v = 0.f;
  if (ch < 12) {
    const int dr = ch / 6, ds = (ch % 6) / 3, c = ch % 3;
    const int r = 2 * ra + dr, s = 2 * sa + ds;
    if (r < 7 && s < 7) v = w[((o * 3 + c) * 7 + r) * 7 + s];
  }
  ws[idx] = __float2bfloat16(v);
}

__device__ __forceinline__ float load_pixel(const float* p) { return *p; }
__device__ __forceinline__ float load_pixel(const __nv_bfloat16* p) { return __bfloat162float(*p); }
// uint8 pixels: torchvision ToTensor (u / 255) followed by Normalize(mean 0.5, std 0.5), the transform of
// dataset.py:119-123, in the same fp32 operation order (so the value equals the reference's tensor element)
__device__ __forceinline__ float load_pixel(const uint8_t* p) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)*p, 255.f), 0.5f), 0.5f);
}

// xs[n][a][b][(dr*2+ds)*3+c] = x[n][c][2a+dr-3][2b+ds-3]  (zero outside the image, channels 12..15 zero)
template <typename T>
__global__ void stem_s2d_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ xs, int H, int W, int Hs,
                                int Ws) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int a = blockIdx.y;
  const int n = blockIdx.z;
  if (b >= Ws) return;
  const T* xn = x + (size_t)n * 3 * H * W;
  __align__(16) __nv_bfloat16 v[16];
#pragma unroll
  for (int dr = 0; dr < 2; ++dr)
#pragma unroll
    for (int ds = 0; ds < 2; ++ds) {
      const int ih = 2 * a + dr - 3, iw = 2 * b + ds - 3;
      const bool in = (ih >= 0 && ih < H && iw >= 0 && iw < W);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float f = 0.f;
        if (in) f = load_pixel(xn + ((size_t)c * H + ih) * W + iw);
        v[(dr * 2 + ds) * 3 + c] = __float2bfloat16(f);
      }
    }
#pragma unroll
  for (int c = 12; c < 16; ++c) v[c] = __float2bfloat16(0.f);
  uint4* dst = reinterpret_cast<uint4*>(xs + (((size_t)n * Hs + a) * Ws + b) * 16);
  dst[0] = reinterpret_cast<const uint4*>(v)[0];
  dst[1] = reinterpret_cast<const uint4*>(v)[1];
}

// uint8 specialisation: the transform is a function of the byte alone, so every CTA builds the 256-entry table once
// (two IEEE divisions per ENTRY instead of per pixel value: the generic kernel above spent its time in 24 divisions per
// thread -- 0.34 ms for 64 images against 0.07 ms of memory traffic, profiles/r02_ncu_launches_gb64.txt) and a pixel
// costs one byte load and one shared-memory lookup.  One thread = one output pixel (32 bytes) of a row segment.
__global__ void __launch_bounds__(256) stem_s2d_u8_kernel(const uint8_t* __restrict__ x,
                                                           __nv_bfloat16* __restrict__ xs, int H, int W, int Hs,
                                                           int Ws) {
  __shared__ __nv_bfloat16 lut[256];
  {
    const uint8_t u = (uint8_t)threadIdx.x;
    lut[threadIdx.x] = __float2bfloat16(load_pixel(&u));
  }
  __syncthreads();
  const int n = blockIdx.z;
  const uint8_t* xn = x + (size_t)n * 3 * H * W;
  for (int a = blockIdx.y; a < Hs; a += gridDim.y) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < Ws; b += gridDim.x * blockDim.x) {
      __align__(16) __nv_bfloat16 v[16];
#pragma unroll
      for (int dr = 0; dr < 2; ++dr)
#pragma unroll
        for (int ds = 0; ds < 2; ++ds) {
          const int ih = 2 * a + dr - 3, iw = 2 * b + ds - 3;
          const bool in = (ih >= 0 && ih < H && iw >= 0 && iw < W);
#pragma unroll
          for (int c = 0; c < 3; ++c)
            v[(dr * 2 + ds) * 3 + c] = in ? lut[__ldg(xn + ((size_t)c * H + ih) * W + iw)] : __float2bfloat16(0.f);
        }
#pragma unroll
      for (int c = 12; c < 16; ++c) v[c] = __float2bfloat16(0.f);
      uint4* dst = reinterpret_cast<uint4*>(xs + (((size_t)n * Hs + a) * Ws + b) * 16);
      dst[0] = reinterpret_cast<const uint4*>(v)[0];
      dst[1] = reinterpret_cast<const uint4*>(v)[1];
    }
  }
}

}  // namespace ecgmm

using namespace ecgmm;

extern "C" int ecgmm_nchw_f32_to_nhwc_bf16(const float* x, ecgmm_bf16* y, int N, int C, int H, int W,
                                           void* stream) {
  ECGMM_CHECK(x && y, ECGMM_ERR_ARG, "nchw_f32_to_nhwc_bf16: null pointer");
  if (N == 0) return ECGMM_OK;
  const int P = H * W;
  dim3 grid(ceil_div(P, 32), ceil_div(C, 32), N), block(32, 8);
  ECGMM_CHECK(grid.y <= 65535 && grid.z <= 65535, ECGMM_ERR_SHAPE, "layout grid too large");
  nchw_to_nhwc_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(y), C, P);
  return check_launch("nchw_to_nhwc_kernel");
}

extern "C" int ecgmm_nhwc_bf16_to_nchw_f32(const ecgmm_bf16* x, float* y, int N, int C, int H, int W,
                                           void* stream) {
  ECGMM_CHECK(x && y, ECGMM_ERR_ARG, "nhwc_bf16_to_nchw_f32: null pointer");
  if (N == 0) return ECGMM_OK;
  const int P = H * W;
  dim3 grid(ceil_div(P, 32), ceil_div(C, 32), N), block(32, 8);
  ECGMM_CHECK(grid.y <= 65535 && grid.z <= 65535, ECGMM_ERR_SHAPE, "layout grid too large");
  nhwc_to_nchw_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), y, C, P);
  return check_launch("nhwc_to_nchw_kernel");
}

extern "C" int ecgmm_conv_weight_prep(const float* w, ecgmm_bf16* w_fwd, ecgmm_bf16* w_dgrad, int O, int I, int R,
                                      int S, void* stream) {
  ECGMM_CHECK(w && (w_fwd || w_dgrad), ECGMM_ERR_ARG, "conv_weight_prep: null pointer");
  const size_t total = (size_t)O * I * R * S;
  if (total == 0) return ECGMM_OK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  conv_weight_prep_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, reinterpret_cast<__nv_bfloat16*>(w_fwd), reinterpret_cast<__nv_bfloat16*>(w_dgrad), O, I, R * S);
  return check_launch("conv_weight_prep_kernel");
}

// All stale convolution weights of a stage in ONE launch (27 launches of conv_weight_prep_kernel per training step
// were 0.2 ms at per-GPU batch 64, mostly launch gaps and 2-byte scattered stores).  A CTA converts a 32 (o) x 32 (i) x RS
// block: coalesced fp32 reads (32 i x RS contiguous floats per output channel), a bf16 tile in shared memory, then
// 64-byte runs along i for w_fwd [O][RS][I] and along o for w_dgrad [I][RS][O].
constexpr int kPrepBatch = 32;
struct PrepBatchParams {
  const float* w[kPrepBatch];
  __nv_bfloat16* fwd[kPrepBatch];
  __nv_bfloat16* dg[kPrepBatch];
  int O[kPrepBatch], I[kPrepBatch], RS[kPrepBatch];
  int block_end[kPrepBatch];  // exclusive prefix sum of the CTAs per tensor
  int n;
};

__global__ void __launch_bounds__(256) conv_weight_prep_batch_kernel(const __grid_constant__ PrepBatchParams p) {
  __shared__ __nv_bfloat16 tile[32][32 * 9 + 2];  // [o][i * RS + t], row padded against bank conflicts
  int k = 0;
  while (k < p.n - 1 && (int)blockIdx.x >= p.block_end[k]) ++k;
  const int b = blockIdx.x - (k ? p.block_end[k - 1] : 0);
  const int O = p.O[k], I = p.I[k], RS = p.RS[k];
  const int ib = I / 32;
  const int o0 = (b / ib) * 32, i0 = (b % ib) * 32;
  const float* w = p.w[k];
  const int run = 32 * RS;  // contiguous floats per output channel of this block
  for (int e = threadIdx.x; e < 32 * run; e += 256) {
    const int o = e / run, r = e - o * run;
    tile[o][r] = __float2bfloat16(w[((size_t)(o0 + o) * I + i0) * RS + r]);
  }
  __syncthreads();
  __nv_bfloat16* fwd = p.fwd[k];
  __nv_bfloat16* dg = p.dg[k];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;  // 8 warps
  for (int ot = grp; ot < 32 * RS; ot += 8) {  // (o, t) rows of w_fwd: 32 consecutive i
    const int o = ot / RS, t = ot - o * RS;
    if (fwd) fwd[((size_t)(o0 + o) * RS + t) * I + i0 + lane] = tile[o][lane * RS + t];
  }
  for (int it = grp; it < 32 * RS; it += 8) {  // (i, t) rows of w_dgrad: 32 consecutive o
    const int i = it / RS, t = it - i * RS;
    if (dg) dg[((size_t)(i0 + i) * RS + t) * O + o0 + lane] = tile[lane][i * RS + t];
  }
}

extern "C" int ecgmm_conv_weight_prep_batch(const ecgmm_weight_prep_desc* descs, int n, void* stream) {
  ECGMM_CHECK(n >= 0 && (descs || n == 0), ECGMM_ERR_ARG, "conv_weight_prep_batch: null descriptor array");
  for (int base = 0; base < n; base += kPrepBatch) {
    PrepBatchParams p;
    memset(&p, 0, sizeof(p));
    int blocks = 0;
    const int m = n - base < kPrepBatch ? n - base : kPrepBatch;
    for (int k = 0; k < m; ++k) {
      const ecgmm_weight_prep_desc& d = descs[base + k];
      const int RS = d.R * d.S;
      ECGMM_CHECK(d.w && (d.w_fwd || d.w_dgrad), ECGMM_ERR_ARG, "conv_weight_prep_batch: null pointer in descriptor %d", base + k);
      ECGMM_CHECK(d.O > 0 && d.I > 0 && d.O % 32 == 0 && d.I % 32 == 0 && RS >= 1 && RS <= 9, ECGMM_ERR_SHAPE,
                  "conv_weight_prep_batch: descriptor %d: O=%d I=%d must be multiples of 32, R*S=%d <= 9 "
                  "(use ecgmm_conv_weight_prep otherwise)", base + k, d.O, d.I, RS);
      p.w[k] = d.w;
      p.fwd[k] = reinterpret_cast<__nv_bfloat16*>(d.w_fwd);
      p.dg[k] = reinterpret_cast<__nv_bfloat16*>(d.w_dgrad);
      p.O[k] = d.O;
      p.I[k] = d.I;
      p.RS[k] = RS;
      blocks += (d.O / 32) * (d.I / 32);
      p.block_end[k] = blocks;
    }
    p.n = m;
    if (blocks == 0) continue;
    conv_weight_prep_batch_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    int rc = check_launch("conv_weight_prep_batch_kernel");
    if (rc) return rc;
  }
  return ECGMM_OK;
}

extern "C" int ecgmm_stem_weight_prep(const float* w, ecgmm_bf16* w_s2d, void* stream) {
  ECGMM_CHECK(w && w_s2d, ECGMM_ERR_ARG, "stem_weight_prep: null pointer");
  stem_weight_prep_kernel<<<64, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, reinterpret_cast<__nv_bfloat16*>(w_s2d));
  return check_launch("stem_weight_prep_kernel");
}

extern "C" int ecgmm_stem_s2d(const void* x, int x_dtype, ecgmm_bf16* xs, int N, int H, int W, void* stream) {
  ECGMM_CHECK(x && xs, ECGMM_ERR_ARG, "stem_s2d: null pointer");
  ECGMM_CHECK(x_dtype >= 0 && x_dtype <= 2, ECGMM_ERR_ARG, "stem_s2d: x_dtype %d (0 fp32, 1 bf16, 2 uint8)", x_dtype);
  if (N == 0) return ECGMM_OK;
  int Hs, Ws;
  ecgmm_stem_s2d_dims(H, W, &Hs, &Ws);
  ECGMM_CHECK(Hs <= 65535 && N <= 65535, ECGMM_ERR_SHAPE, "stem_s2d: image too tall / batch too large");
  dim3 grid(ceil_div(Ws, 128), Hs, N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_dtype == 2) {
    // row segments of 256 pixels; 8 output rows per CTA so that the lookup table is built once per 8 rows
    dim3 g8(ceil_div(Ws, 256), ceil_div(Hs, 8), N);
    stem_s2d_u8_kernel<<<g8, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(x), reinterpret_cast<__nv_bfloat16*>(xs), H,
                                           W, Hs, Ws);
  }
  else if (x_dtype == 1)
    stem_s2d_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                        reinterpret_cast<__nv_bfloat16*>(xs), H, W, Hs, Ws);
  else
    stem_s2d_kernel<float><<<grid, 128, 0, st>>>(reinterpret_cast<const float*>(x),
                                                 reinterpret_cast<__nv_bfloat16*>(xs), H, W, Hs, Ws);
  return check_launch("stem_s2d_kernel");
}
