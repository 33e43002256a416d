// HBM-bound kernels around the convolutions: BatchNorm statistics / apply / backward (2-D and
// 1-D), ReLU, residual add, squeeze-excite scaling, 3x3/s2 max-pool fused behind the stem
// BatchNorm, global average pool.  Activations are channels-last bf16 [N][P][C] (P = H*W or L),
// 8 channels (16 bytes) per thread; statistics and reductions are fp32 partials finalised in fp64.
//
// Reduction scheme (deterministic, no atomics): a grid of (SPLIT, N) CTAs each reduces a slab of
// pixels of one sample for all C channels and writes one partial row [C]; a one-CTA-per-128-
// channels finalize kernel folds the N*SPLIT rows.  Per-sample rows double as the squeeze
// (mean over L) of the SE blocks of the 1-D ResNet.
#include "common.h"

#include <stdlib.h>
#include "vec.cuh"

#include <unordered_map>

namespace ecgmm {

constexpr int kRedThreads = 256;

// ------------------------------------------------------------------------------------------
// forward statistics:  psum/psq [N][SPLIT][C] = sum / sum of squares over the slab
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRedThreads) chan_stats_kernel(const __nv_bfloat16* __restrict__ x,
                                                                  float* __restrict__ psum,
                                                                  float* __restrict__ psq, int P, int C,
                                                                  int rows_per_split) {
  extern __shared__ float sred[];  // [rows][C] x 2
  const int CG = C >> 3;
  const int rows = kRedThreads / CG;
  const int cg = threadIdx.x % CG, r = threadIdx.x / CG;
  const int n = blockIdx.y, split = blockIdx.x;
  const int p0 = split * rows_per_split;
  const int p1 = min(P, p0 + rows_per_split);
  const uint4* base = reinterpret_cast<const uint4*>(x + (size_t)n * P * C) + cg;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  int p = p0 + r;
  for (; p + 3 * rows < p1; p += 4 * rows) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld_stream(base + (size_t)(p + u * rows) * CG);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] = fmaf(f[j], f[j], q[j]);
      }
    }
  }
  for (; p < p1; p += rows) {
    float f[8];
    unpack8(ld_stream(base + (size_t)p * CG), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += f[j];
      q[j] = fmaf(f[j], f[j], q[j]);
    }
  }
  float* ss = sred;
  float* sq = sred + rows * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ss[r * C + cg * 8 + j] = s[j];
    sq[r * C + cg * 8 + j] = q[j];
  }
  __syncthreads();
  const size_t orow = ((size_t)n * gridDim.x + split) * C;
  for (int c = threadIdx.x; c < C; c += kRedThreads) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < rows; ++i) {
      a += ss[i * C + c];
      b += sq[i * C + c];
    }
    psum[orow + c] = a;
    psq[orow + c] = b;
  }
}

// Fold the partial rows; emit mean / invstd (saved for backward), the affine (scale, shift) the
// apply kernel uses, optional per-sample sums (SE squeeze) and the running-statistics update.
//   conv_bias: the convolution in front has a bias that the conv kernel does NOT add; in
//   training mode it only shifts the batch mean (and so running_mean), never the output.
// One CTA = 8 channels x 128 row lanes (a 32-byte sector per row read); lanes stride over the N*SPLIT
// partial rows (or over the samples when per-sample sums are wanted) with 4 independent fp64 chains,
// then a shared-memory fold.  C/8 CTAs keep this latency-bound fold at a few microseconds.
// (Measured and not kept, round 2: a cluster of 8 CTAs per channel group sharing the rows, rank 0 adding their sums
// through distributed shared memory - equal at C <= 128, 5-15 us slower per launch at C >= 256: the fold is not the
// latency chain it looks like, the cluster launch costs more than the trips it saves.  profiles/r02ee_ab*.txt.)
constexpr int kFinLanes = 128;
constexpr int kFinCh = 8;

// Sum over the kFinLanes row lanes of a finalize CTA (thread = lane * kFinCh + channel): 7 halving steps in shared
// memory instead of one thread walking 128 doubles.  Result valid in lane 0.
__device__ __forceinline__ void fin_fold(double (&sh_a)[kFinLanes][kFinCh + 1], double (&sh_b)[kFinLanes][kFinCh + 1],
                                         int lane, int cl, double& a, double& b) {
  sh_a[lane][cl] = a;
  sh_b[lane][cl] = b;
  __syncthreads();
#pragma unroll
  for (int h = kFinLanes / 2; h >= 1; h >>= 1) {
    if (lane < h) {
      sh_a[lane][cl] += sh_a[lane + h][cl];
      sh_b[lane][cl] += sh_b[lane + h][cl];
    }
    __syncthreads();
  }
  a = sh_a[0][cl];
  b = sh_b[0][cl];
}

__global__ void __launch_bounds__(kFinCh * kFinLanes) bn_finalize_kernel(
    const float* __restrict__ psum, const float* __restrict__ psq, int rows_total, int split, int C, double count,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ conv_bias, float eps,
    float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
    long long* __restrict__ num_batches, float* __restrict__ mean_out, float* __restrict__ invstd_out,
    float* __restrict__ scale_out, float* __restrict__ shift_out, float* __restrict__ nsum_out) {
  __shared__ double sh_s[kFinLanes][kFinCh + 1], sh_q[kFinLanes][kFinCh + 1];
  const int cl = threadIdx.x & (kFinCh - 1), lane = threadIdx.x / kFinCh;
  const int c = blockIdx.x * kFinCh + cl;
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches) *num_batches += 1;
  double s = 0.0, q = 0.0;
  if (c < C) {
    if (nsum_out) {
      const int N = rows_total / split;
      for (int n = lane; n < N; n += kFinLanes) {
        float ns = 0.f;
        for (int k = 0; k < split; ++k) {
          const size_t i = ((size_t)n * split + k) * C + c;
          ns += psum[i];
          q += psq[i];
        }
        nsum_out[(size_t)n * C + c] = ns;
        s += ns;
      }
    } else {
      double s1 = 0.0, q1 = 0.0, s2 = 0.0, q2 = 0.0, s3 = 0.0, q3 = 0.0;
      int i = lane;
      for (; i + 3 * kFinLanes < rows_total; i += 4 * kFinLanes) {
        s += psum[(size_t)i * C + c];
        q += psq[(size_t)i * C + c];
        s1 += psum[(size_t)(i + kFinLanes) * C + c];
        q1 += psq[(size_t)(i + kFinLanes) * C + c];
        s2 += psum[(size_t)(i + 2 * kFinLanes) * C + c];
        q2 += psq[(size_t)(i + 2 * kFinLanes) * C + c];
        s3 += psum[(size_t)(i + 3 * kFinLanes) * C + c];
        q3 += psq[(size_t)(i + 3 * kFinLanes) * C + c];
      }
      for (; i < rows_total; i += kFinLanes) {
        s += psum[(size_t)i * C + c];
        q += psq[(size_t)i * C + c];
      }
      s += s1 + s2 + s3;
      q += q1 + q2 + q3;
    }
  }
  fin_fold(sh_s, sh_q, lane, cl, s, q);
  if (lane != 0 || c >= C) return;
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = g * invstd;
  mean_out[c] = (float)mean;
  invstd_out[c] = invstd;
  scale_out[c] = sc;
  shift_out[c] = b - (float)mean * sc;
  if (running_mean) {
    const float bm = (float)mean + (conv_bias ? conv_bias[c] : 0.f);
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * bm;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval mode: (scale, shift) from the running statistics; the conv bias is folded into shift.
__global__ void bn_eval_coeffs_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ conv_bias, const float* __restrict__ running_mean,
                                      const float* __restrict__ running_var, float eps,
                                      float* __restrict__ scale_out, float* __restrict__ shift_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = (gamma ? gamma[c] : 1.f) / sqrtf(running_var[c] + eps);
  scale_out[c] = sc;
  shift_out[c] = (beta ? beta[c] : 0.f) + ((conv_bias ? conv_bias[c] : 0.f) - running_mean[c]) * sc;
}

// ------------------------------------------------------------------------------------------
// apply:  y = act((x*scale[c] + shift[c]) * se[n][c] + res)
// ------------------------------------------------------------------------------------------
template <bool SE, bool RES, bool RELU>
__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ x,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ shift,
                                                        const float* __restrict__ se,
                                                        const __nv_bfloat16* __restrict__ res,
                                                        __nv_bfloat16* __restrict__ y,
                                                        uint8_t* __restrict__ mask_out, int CG,
                                                        size_t vec_per_sample, size_t total_vec) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total_vec; i += stride) {
    const int cg = (int)(i % CG);
    float f[8], sc[8], sh[8];
    unpack8(ld_stream(reinterpret_cast<const uint4*>(x) + i), f);
    load8f(scale + cg * 8, sc);
    load8f(shift + cg * 8, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
    if (SE) {
      const size_t n = i / vec_per_sample;
      float g[8];
      load8f(se + (n * CG + cg) * 8, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= g[j];
    }
    if (RES) {
      float r[8];
      unpack8(ld_stream(reinterpret_cast<const uint4*>(res) + i), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    if (RELU) {
      if (mask_out) {  // 1 bit per element: what the backward pass needs instead of re-reading y
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) m |= (f[j] > 0.f ? 1u : 0u) << j;
        mask_out[i] = (uint8_t)m;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    reinterpret_cast<uint4*>(y)[i] = pack8(f);
  }
}

// bn_apply_kernel without squeeze-excite for C/8 a power of two <= 256: the grid stride is a multiple of C/8, so a
// thread keeps its channel group and scale / shift live in registers (the generic kernel pays a 64-bit modulo and four
// 16-byte coefficient loads per vector: ~140 instructions per 16 bytes, which at the power-capped clock is as much of
// a limit as the HBM), two vectors per trip.  Same arithmetic per element.
// (Measured and not kept, profiles/r02gg_ab128.txt: the coefficients in shared memory under a 48-register cap for 5 CTAs
// per SM - the compiler spills, the residual variant is 1.8x slower; four vectors per trip at 3 CTAs per SM - slower too.)
template <bool RES, bool RELU>
__global__ void __launch_bounds__(256) bn_apply_fast_kernel(const uint4* __restrict__ x,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ shift,
                                                             const uint4* __restrict__ res, uint4* __restrict__ y,
                                                             uint8_t* __restrict__ mask_out, int CG,
                                                             size_t total_vec) {
  constexpr int U = 2;
  const size_t stride = (size_t)gridDim.x * 256;
  const size_t i0 = blockIdx.x * (size_t)256 + threadIdx.x;
  const int cg = (int)(i0 & (size_t)(CG - 1));
  float sc[8], sh[8];
  load8f(scale + cg * 8, sc);
  load8f(shift + cg * 8, sh);
  for (size_t i = i0; i < total_vec; i += U * stride) {
    uint4 vx[U], vr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t k = i + u * stride;
      if (k < total_vec) {
        vx[u] = ld_stream(x + k);
        if (RES) vr[u] = ld_stream(res + k);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t k = i + u * stride;
      if (k >= total_vec) break;
      float f[8];
      unpack8(vx[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
      if (RES) {
        float r[8];
        unpack8(vr[u], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += r[j];
      }
      if (RELU) {
        if (mask_out) {
          uint32_t m = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) m |= (f[j] > 0.f ? 1u : 0u) << j;
          mask_out[k] = (uint8_t)m;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      y[k] = pack8(f);
    }
  }
}

// ------------------------------------------------------------------------------------------
// stem: y = maxpool3x3/s2/p1(relu(x*scale + shift)), arg = position of the maximum inside the
// window (first maximum in row-major window order, torch's tie rule), 0..8
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) bn_relu_maxpool_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ shift,
                                                               __nv_bfloat16* __restrict__ y,
                                                               uint8_t* __restrict__ arg, int H, int W, int Ho,
                                                               int Wo, int CG, size_t total_vec) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total_vec; i += stride) {
    const int cg = (int)(i % CG);
    size_t t = i / CG;
    const int ow = (int)(t % Wo);
    t /= Wo;
    const int oh = (int)(t % Ho);
    const size_t n = t / Ho;
    float sc[8], sh[8];
    load8f(scale + cg * 8, sc);
    load8f(shift + cg * 8, sh);
    // max over the window of relu(sc*x+sh) = relu(sc*max(x)+sh) for sc > 0 and relu(sc*min(x)+sh) for sc < 0:
    // the window scan runs on the RAW bf16 pairs (2 channels per instruction), with the sign bit flipped
    // where sc < 0 so that one packed max serves both cases; the affine map is applied once at the end.
    // Strict > keeps the first maximum (torch's tie rule).  (sc == 0 exactly: any element is a maximum.)
    uint32_t flip[4], best2[4], idx2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      flip[k] = (sc[2 * k] < 0.f ? 0x8000u : 0u) | (sc[2 * k + 1] < 0.f ? 0x80000000u : 0u);
      best2[k] = 0xFF80FF80u;  // (-inf, -inf)
      idx2[k] = 0u;
    }
    // all 9 window loads are issued up front with clamped coordinates (independent loads in
    // flight); out-of-image taps are masked afterwards
    uint4 v[9];
    bool ok[9];
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int h = 2 * oh - 1 + dh;
      const int hc = min(max(h, 0), H - 1);
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
        const int w = 2 * ow - 1 + dw;
        const int wc = min(max(w, 0), W - 1);
        ok[dh * 3 + dw] = (h == hc) && (w == wc);
        v[dh * 3 + dw] = reinterpret_cast<const uint4*>(x)[((n * H + hc) * W + wc) * CG + cg];
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      if (!ok[t]) continue;
      const uint32_t code2 = (uint32_t)t | ((uint32_t)t << 16);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t xw = (k == 0 ? v[t].x : (k == 1 ? v[t].y : (k == 2 ? v[t].z : v[t].w))) ^ flip[k];
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&xw);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&best2[k]);
        const uint32_t m = __hgt2_mask(a, b);  // 0xFFFF per half where a > b
        const __nv_bfloat162 mx = __hmax2(a, b);
        best2[k] = *reinterpret_cast<const uint32_t*>(&mx);
        idx2[k] = (idx2[k] & ~m) | (code2 & m);
      }
    }
    float best[8];
    int bi[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t xw = best2[k] ^ flip[k];
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw));
      best[2 * k] = fmaxf(fmaf(f.x, sc[2 * k], sh[2 * k]), 0.f);
      best[2 * k + 1] = fmaxf(fmaf(f.y, sc[2 * k + 1], sh[2 * k + 1]), 0.f);
      bi[2 * k] = idx2[k] & 0xffff;
      bi[2 * k + 1] = idx2[k] >> 16;
    }
    reinterpret_cast<uint4*>(y)[i] = pack8(best);
    if (arg) {
      uint2 a;
      a.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      a.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      reinterpret_cast<uint2*>(arg)[i] = a;
    }
  }
}

// The same operation for C = 64 with the input window staged in shared memory.  In the kernel above every input element
// is fetched 2.25 times (3x3 windows, stride 2) and the vertical overlap is between CTAs: at 250x2500 images the L2 -> SM
// traffic was 23 GB per batch of 512 for a 10 GB tensor and the kernel sat at 0.58 of the HBM peak.  Here a CTA owns
// 4 x 16 pooled pixels, loads their 9 x 33 input pixels once (1.16x overlap), and scans the windows from shared memory.
// Same arithmetic and tie rule -> bit-identical outputs and argmax codes.
constexpr int kPoolTOH = 4, kPoolTOW = 16, kPoolIH = 2 * kPoolTOH + 1, kPoolIW = 2 * kPoolTOW + 1;

// BRANCHFREE (default): the tile is stored in the "flipped" domain (sign bit toggled where scale < 0, which is what
// the scan compares anyway) and positions outside the image hold -inf there, so the 9-tap scan needs no bounds
// tests (the v1 scan carried 18 divergent-branch regions per thread) and the 4 XORs per tap move to the fill.  A
// thread's channel group is tid & 7 in the fill AND in the scan (256 % 8 == 0), so the flip words are per-thread
// constants.  -inf never wins a strict > against the initial -inf: same winners, same codes as v1.
template <bool BRANCHFREE>
__global__ void __launch_bounds__(256) bn_relu_maxpool_tiled_kernel(const __nv_bfloat16* __restrict__ x,
                                                                    const float* __restrict__ scale,
                                                                    const float* __restrict__ shift,
                                                                    __nv_bfloat16* __restrict__ y,
                                                                    uint8_t* __restrict__ arg, int H, int W, int Ho,
                                                                    int Wo, int tiles_w, int tiles_h) {
  __shared__ uint4 tile[kPoolIH * kPoolIW * 8];
  int b = blockIdx.x;
  const int tw = b % tiles_w;
  b /= tiles_w;
  const int th = b % tiles_h;
  const size_t n = b / tiles_h;
  const int oh0 = th * kPoolTOH, ow0 = tw * kPoolTOW;
  const int ih0 = 2 * oh0 - 1, iw0 = 2 * ow0 - 1;
  const uint4* xin = reinterpret_cast<const uint4*>(x) + n * (size_t)H * W * 8;
  const int cg = threadIdx.x & 7;
  float sc[8], sh[8];
  load8f(scale + cg * 8, sc);
  load8f(shift + cg * 8, sh);
  uint32_t flip[4];
#pragma unroll
  for (int q = 0; q < 4; ++q)
    flip[q] = (sc[2 * q] < 0.f ? 0x8000u : 0u) | (sc[2 * q + 1] < 0.f ? 0x80000000u : 0u);
  constexpr int kTileVec = kPoolIH * kPoolIW * 8;
  if (BRANCHFREE) {
    constexpr int kIters = (kTileVec + 255) / 256;
    uint4 v[kIters];
    int r = 0, c = threadIdx.x >> 3;  // tile pixel of this thread: advances by 32 pixels per iteration (32 < kPoolIW)
#pragma unroll
    for (int it = 0; it < kIters; ++it) {  // all loads of the thread in flight before the first store
      const int h = ih0 + r, w = iw0 + c;
      const bool in = threadIdx.x + 256 * it < kTileVec && (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W;
      v[it] = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);  // (-inf, -inf) in the flipped domain
      if (in) {
        v[it] = xin[(h * W + w) * 8 + cg];
        v[it].x ^= flip[0];
        v[it].y ^= flip[1];
        v[it].z ^= flip[2];
        v[it].w ^= flip[3];
      }
      c += 32;
      if (c >= kPoolIW) {
        c -= kPoolIW;
        ++r;
      }
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int e = threadIdx.x + 256 * it;
      if (e < kTileVec) tile[e] = v[it];
    }
  } else {
    for (int e = threadIdx.x; e < kTileVec; e += 256) {
      const int pix = e >> 3;
      const int r = pix / kPoolIW, c = pix - r * kPoolIW;
      const int h = ih0 + r, w = iw0 + c;
      if (h >= 0 && h < H && w >= 0 && w < W) tile[e] = xin[((size_t)h * W + w) * 8 + cg];
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int item = threadIdx.x + 256 * k;
    const int px = item >> 3;
    const int tow = px & (kPoolTOW - 1), toh = px >> 4;
    const int oh = oh0 + toh, ow = ow0 + tow;
    if (oh >= Ho || ow >= Wo) continue;
    uint32_t best2[4], idx2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      best2[q] = 0xFF80FF80u;  // (-inf, -inf)
      idx2[q] = 0u;
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int dh = t / 3, dw = t - 3 * dh;
      if (!BRANCHFREE) {
        const int h = 2 * oh - 1 + dh, w = 2 * ow - 1 + dw;
        if (h < 0 || h >= H || w < 0 || w >= W) continue;
      }
      const uint4 v = tile[((2 * toh + dh) * kPoolIW + 2 * tow + dw) * 8 + cg];
      const uint32_t code2 = (uint32_t)t | ((uint32_t)t << 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t raw = (q == 0 ? v.x : (q == 1 ? v.y : (q == 2 ? v.z : v.w)));
        const uint32_t xw = BRANCHFREE ? raw : (raw ^ flip[q]);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&xw);
        const __nv_bfloat162 bb = *reinterpret_cast<const __nv_bfloat162*>(&best2[q]);
        const uint32_t m = __hgt2_mask(a, bb);
        const __nv_bfloat162 mx = __hmax2(a, bb);
        best2[q] = *reinterpret_cast<const uint32_t*>(&mx);
        idx2[q] = (idx2[q] & ~m) | (code2 & m);
      }
    }
    float best[8];
    int bi[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t xw = best2[q] ^ flip[q];
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw));
      best[2 * q] = fmaxf(fmaf(f.x, sc[2 * q], sh[2 * q]), 0.f);
      best[2 * q + 1] = fmaxf(fmaf(f.y, sc[2 * q + 1], sh[2 * q + 1]), 0.f);
      bi[2 * q] = idx2[q] & 0xffff;
      bi[2 * q + 1] = idx2[q] >> 16;
    }
    const size_t o = ((n * Ho + oh) * Wo + ow) * 8 + cg;
    reinterpret_cast<uint4*>(y)[o] = pack8(best);
    if (arg) {
      uint2 a2;
      a2.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      a2.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      reinterpret_cast<uint2*>(arg)[o] = a2;
    }
  }
}

// Gradient reaching the pre-pool activation at (h, w) for 8 channels: sum of the pooled
// gradients of the (at most 4) windows whose recorded argmax is this pixel, gated by ReLU.
__device__ __forceinline__ void pool_gather_dz(const __nv_bfloat16* __restrict__ dyp,
                                               const uint8_t* __restrict__ arg, size_t n, int h, int w, int Ho,
                                               int Wo, int CG, int cg, const float (&xv)[8], const float (&sc)[8],
                                               const float (&sh)[8], float (&dz)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) dz[j] = 0.f;
  const int oh0 = h >> 1, oh1 = (h + 1) >> 1;  // equal when h is even
  const int ow0 = w >> 1, ow1 = (w + 1) >> 1;
  for (int oh = oh0; oh <= oh1; ++oh) {
    if (oh >= Ho) continue;
    const int dh = h - 2 * oh + 1;
    for (int ow = ow0; ow <= ow1; ++ow) {
      if (ow >= Wo) continue;
      const int code = dh * 3 + (w - 2 * ow + 1);
      const size_t o = ((n * Ho + oh) * Wo + ow) * CG + cg;
      const uint2 a = reinterpret_cast<const uint2*>(arg)[o];
      float g[8];
      unpack8(reinterpret_cast<const uint4*>(dyp)[o], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int aj = ((j < 4 ? a.x : a.y) >> (8 * (j & 3))) & 0xff;
        if (aj == code) dz[j] += g[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (fmaf(xv[j], sc[j], sh[j]) <= 0.f) dz[j] = 0.f;
}

// ------------------------------------------------------------------------------------------
// backward reductions:  p1/p2 [N][SPLIT][C] = sum dz, sum dz * xhat
//   MODE 0: dz = dy     MODE 1: dz = dy * (y > 0)     MODE 2: stem (pool gather + relu gate)
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kRedThreads) bn_bwd_reduce_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
    const uint8_t* __restrict__ arg, const float* __restrict__ mean, const float* __restrict__ invstd,
    const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ p1,
    float* __restrict__ p2, int P, int C, int rows_per_split, int H, int W, int Ho, int Wo) {
  extern __shared__ float sred[];
  const int CG = C >> 3;
  const int rows = kRedThreads / CG;
  const int cg = threadIdx.x % CG, r = threadIdx.x / CG;
  const int n = blockIdx.y, split = blockIdx.x;
  const int pa = split * rows_per_split;
  const int pb = min(P, pa + rows_per_split);
  float mu[8], is[8], sc[8], sh[8], s[8], q[8];
  load8f(mean + cg * 8, mu);
  load8f(invstd + cg * 8, is);
  if (MODE == 4) {  // mean := beta, invstd := gamma: xhat of the selected pre-pool element = (y - beta) / gamma
#pragma unroll
    for (int j = 0; j < 8; ++j) is[j] = is[j] != 0.f ? 1.f / is[j] : 0.f;
  }
  if (MODE == 2) {
    load8f(scale + cg * 8, sc);
    load8f(shift + cg * 8, sh);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const size_t sbase = (size_t)n * P * CG + cg;
  int p = pa + r;
  if (MODE != 2) {
    constexpr int U = 4;  // independent 16-byte loads in flight per operand
    for (; p + (U - 1) * rows < pb; p += U * rows) {
      uint4 vx[U], vd[U], vy[U];
      uint32_t vm[U];
      if (MODE == 3) {  // the mask bytes first: the first thing the arithmetic below waits for (ncu: 35 % of the stall
                        // samples sat on their consumer when they were issued behind the 16-byte loads)
#pragma unroll
        for (int u = 0; u < U; ++u) vm[u] = arg[sbase + (size_t)(p + u * rows) * CG];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t i = sbase + (size_t)(p + u * rows) * CG;
        vx[u] = ld_stream(reinterpret_cast<const uint4*>(x) + i);
        vd[u] = ld_stream(reinterpret_cast<const uint4*>(dy) + i);
        if (MODE == 1) vy[u] = ld_stream(reinterpret_cast<const uint4*>(y) + i);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float xv[8], dz[8];
        unpack8(vx[u], xv);
        if (MODE == 3) relu_mask_words(vm[u], vd[u]);  // masked lanes become +0.0 (what `dz = 0.f` was)
        unpack8(vd[u], dz);
        if (MODE == 1) {
          float yv[8];
          unpack8(vy[u], yv);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (yv[j] <= 0.f) dz[j] = 0.f;
        }
        if (MODE == 4) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (xv[j] <= 0.f) dz[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += dz[j];
          q[j] = fmaf(dz[j], (xv[j] - mu[j]) * is[j], q[j]);
        }
      }
    }
  }
  for (; p < pb; p += rows) {
    const size_t i = sbase + (size_t)p * CG;
    float xv[8], dz[8];
    unpack8(ld_stream(reinterpret_cast<const uint4*>(x) + i), xv);
    if (MODE == 2) {
      pool_gather_dz(dy, arg, n, p / W, p % W, Ho, Wo, CG, cg, xv, sc, sh, dz);
    } else {
      unpack8(ld_stream(reinterpret_cast<const uint4*>(dy) + i), dz);
      if (MODE == 1) {
        float yv[8];
        unpack8(ld_stream(reinterpret_cast<const uint4*>(y) + i), yv);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (yv[j] <= 0.f) dz[j] = 0.f;
      }
      if (MODE == 3) {
        const uint32_t m = arg[i];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (!((m >> j) & 1u)) dz[j] = 0.f;
      }
      if (MODE == 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (xv[j] <= 0.f) dz[j] = 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += dz[j];
      q[j] = fmaf(dz[j], (xv[j] - mu[j]) * is[j], q[j]);
    }
  }
  float* ss = sred;
  float* sq = sred + rows * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ss[r * C + cg * 8 + j] = s[j];
    sq[r * C + cg * 8 + j] = q[j];
  }
  __syncthreads();
  const size_t orow = ((size_t)n * gridDim.x + split) * C;
  for (int c = threadIdx.x; c < C; c += kRedThreads) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < rows; ++i) {
      a += ss[i * C + c];
      b += sq[i * C + c];
    }
    p1[orow + c] = a;
    p2[orow + c] = b;
  }
}

// bn_bwd_reduce_kernel<0 | 3> with x and dy staged through the thread-private cp.async ring (bn_bwd_apply_async_kernel): a
// thread walks its pixels p = pa + r, pa + r + rows, ... one vector at a time with kRedStages - 1 copies in flight, and
// accumulates in exactly the order of the register kernel (bit-identical partial rows).
constexpr int kRedStages = 6;

template <int MODE>
__global__ void __launch_bounds__(kRedThreads, 3) bn_bwd_reduce_async_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ arg,
    const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ p1, float* __restrict__ p2,
    int P, int C, int rows_per_split) {
  static_assert(MODE == 0 || MODE == 3, "register kernel for the other modes");
  constexpr int S = kRedStages;
  extern __shared__ uint4 ring[];  // [S][2][256] uint4, then the [2][rows][C] float fold area
  float* sred = reinterpret_cast<float*>(ring + S * 2 * kRedThreads);
  const int CG = C >> 3;
  const int rows = kRedThreads / CG;
  const int cg = threadIdx.x % CG, r = threadIdx.x / CG;
  const int n = blockIdx.y, split = blockIdx.x;
  const int pa = split * rows_per_split;
  const int pb = min(P, pa + rows_per_split);
  float mu[8], is[8], s[8], q[8];
  load8f(mean + cg * 8, mu);
  load8f(invstd + cg * 8, is);
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const uint4* xv4 = reinterpret_cast<const uint4*>(x) + (size_t)n * P * CG + cg;
  const uint4* dv4 = reinterpret_cast<const uint4*>(dy) + (size_t)n * P * CG + cg;
  const uint8_t* mk = arg + (size_t)n * P * CG + cg;
  uint4* slot0 = ring + threadIdx.x;
  uint32_t mreg[S];
#pragma unroll
  for (int t = 0; t < S; ++t) mreg[t] = 0;
  const int p0 = pa + r;
#pragma unroll
  for (int t = 0; t < S - 1; ++t) {
    const int p = p0 + t * rows;
    if (p < pb) {
      if (MODE == 3) mreg[t] = mk[(size_t)p * CG];
      cp_async16(slot0 + (2 * t) * kRedThreads, xv4 + (size_t)p * CG);
      cp_async16(slot0 + (2 * t + 1) * kRedThreads, dv4 + (size_t)p * CG);
    }
    cp_async_commit();
  }
  for (int base = p0; base < pb; base += S * rows) {
#pragma unroll
    for (int t = 0; t < S; ++t) {
      const int p = base + t * rows;
      if (p >= pb) break;
      const int pn = p + (S - 1) * rows;
      const int tn = (t + S - 1) % S;
      if (pn < pb) {
        if (MODE == 3) mreg[tn] = mk[(size_t)pn * CG];
        cp_async16(slot0 + (2 * tn) * kRedThreads, xv4 + (size_t)pn * CG);
        cp_async16(slot0 + (2 * tn + 1) * kRedThreads, dv4 + (size_t)pn * CG);
      }
      cp_async_commit();
      cp_async_wait<S - 1>();
      const uint4 vx = slot0[(2 * t) * kRedThreads];
      uint4 vd = slot0[(2 * t + 1) * kRedThreads];
      if (MODE == 3) relu_mask_words(mreg[t], vd);
      float xv[8], dz[8];
      unpack8(vx, xv);
      unpack8(vd, dz);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += dz[j];
        q[j] = fmaf(dz[j], (xv[j] - mu[j]) * is[j], q[j]);
      }
    }
  }
  float* ss = sred;
  float* sq = sred + rows * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ss[r * C + cg * 8 + j] = s[j];
    sq[r * C + cg * 8 + j] = q[j];
  }
  __syncthreads();
  const size_t orow = ((size_t)n * gridDim.x + split) * C;
  for (int c = threadIdx.x; c < C; c += kRedThreads) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < rows; ++i) {
      a += ss[i * C + c];
      b += sq[i * C + c];
    }
    p1[orow + c] = a;
    p2[orow + c] = b;
  }
}

// Fold the backward partials.  With du = dz*se[n][c] + q[n][c] (SE blocks; se = 1, q = 0
// otherwise) the BatchNorm backward is   dx = A*se*dz + B*x + D + A*q   with per-channel
//   A = gamma*invstd,  B = -gamma*invstd^2*m2,  D = -A*m1 + gamma*invstd^2*mean*m2,
//   m1 = mean(du), m2 = mean(du*xhat);  dgamma = sum(du*xhat), dbeta = sum(du).
__global__ void __launch_bounds__(kFinCh * kFinLanes) bn_bwd_finalize_kernel(
    const float* __restrict__ p1, const float* __restrict__ p2, int N, int split, int C, double per_sample,
    double count, const float* __restrict__ gamma, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ se, const float* __restrict__ q,
    const float* __restrict__ nsum, float* __restrict__ dgamma, float* __restrict__ dbeta,
    float* __restrict__ coefA, float* __restrict__ coefB, float* __restrict__ coefD) {
  __shared__ double sh_a[kFinLanes][kFinCh + 1], sh_b[kFinLanes][kFinCh + 1];
  const int cl = threadIdx.x & (kFinCh - 1), lane = threadIdx.x / kFinCh;
  const int c = blockIdx.x * kFinCh + cl;
  double s1 = 0.0, s2 = 0.0;
  double mu = 0.0, is = 0.0;
  if (c < C) {
    mu = mean[c];
    is = invstd[c];
    if (se) {
      for (int n = lane; n < N; n += kFinLanes) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < split; ++k) {
          const size_t i = ((size_t)n * split + k) * C + c;
          a += p1[i];
          b += p2[i];
        }
        const double g = se[(size_t)n * C + c], qq = q[(size_t)n * C + c];
        const double sum_xhat = ((double)nsum[(size_t)n * C + c] - per_sample * mu) * is;
        s1 += g * a + per_sample * qq;
        s2 += g * b + qq * sum_xhat;
      }
    } else {
      const int rows_total = N * split;
      double t1 = 0.0, t2 = 0.0, u1 = 0.0, u2 = 0.0, v1 = 0.0, v2 = 0.0;
      int i = lane;
      for (; i + 3 * kFinLanes < rows_total; i += 4 * kFinLanes) {
        s1 += p1[(size_t)i * C + c];
        s2 += p2[(size_t)i * C + c];
        t1 += p1[(size_t)(i + kFinLanes) * C + c];
        t2 += p2[(size_t)(i + kFinLanes) * C + c];
        u1 += p1[(size_t)(i + 2 * kFinLanes) * C + c];
        u2 += p2[(size_t)(i + 2 * kFinLanes) * C + c];
        v1 += p1[(size_t)(i + 3 * kFinLanes) * C + c];
        v2 += p2[(size_t)(i + 3 * kFinLanes) * C + c];
      }
      for (; i < rows_total; i += kFinLanes) {
        s1 += p1[(size_t)i * C + c];
        s2 += p2[(size_t)i * C + c];
      }
      s1 += t1 + u1 + v1;
      s2 += t2 + u2 + v2;
    }
  }
  fin_fold(sh_a, sh_b, lane, cl, s1, s2);
  if (lane != 0 || c >= C) return;
  const double M = count > 0.0 ? count : per_sample * N;
  const double m1 = s1 / M, m2 = s2 / M;
  const double g = gamma ? gamma[c] : 1.0;
  if (dgamma) dgamma[c] = (float)s2;
  if (dbeta) dbeta[c] = (float)s1;
  const double A = g * is;
  coefA[c] = (float)A;
  coefB[c] = (float)(-g * is * is * m2);
  coefD[c] = (float)(-A * m1 + g * is * is * mu * m2);
}

// dx = A*se*dz + B*x + D + A*q ;  optionally also stores dz (gradient of the residual branch).
template <int MODE, bool SE>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
    const uint8_t* __restrict__ arg, const float* __restrict__ coefA, const float* __restrict__ coefB,
    const float* __restrict__ coefD, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ se, const float* __restrict__ q, __nv_bfloat16* __restrict__ dx,
    __nv_bfloat16* __restrict__ dz_out, int CG, size_t vec_per_sample, size_t total_vec, int H, int W, int Ho,
    int Wo) {
  constexpr int U = (MODE == 2) ? 1 : 2;  // vectors per thread per trip: U x (2..3) loads in flight
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < total_vec; i0 += U * stride) {
    uint4 vx[U], vd[U], vy[U];
    uint32_t vm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * stride;
      if (i < total_vec) {
        vx[u] = ld_stream(reinterpret_cast<const uint4*>(x) + i);
        if (MODE != 2) vd[u] = ld_stream(reinterpret_cast<const uint4*>(dy) + i);
        if (MODE == 1) vy[u] = ld_stream(reinterpret_cast<const uint4*>(y) + i);
        if (MODE == 3) vm[u] = arg[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * stride;
      if (i >= total_vec) break;
      const int cg = (int)(i % CG);
      const size_t n = i / vec_per_sample;
      float xv[8], dz[8], A[8], B[8], D[8];
      unpack8(vx[u], xv);
      load8f(coefA + cg * 8, A);
      load8f(coefB + cg * 8, B);
      load8f(coefD + cg * 8, D);
      if (MODE == 2) {
        float sc[8], sh[8];
        load8f(scale + cg * 8, sc);
        load8f(shift + cg * 8, sh);
        const size_t p = (i / CG) % ((size_t)H * W);
        pool_gather_dz(dy, arg, n, (int)(p / W), (int)(p % W), Ho, Wo, CG, cg, xv, sc, sh, dz);
      } else {
        unpack8(vd[u], dz);
        if (MODE == 1) {
          float yv[8];
          unpack8(vy[u], yv);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (yv[j] <= 0.f) dz[j] = 0.f;
        }
        if (MODE == 3) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (!((vm[u] >> j) & 1u)) dz[j] = 0.f;
        }
      }
      if (dz_out) reinterpret_cast<uint4*>(dz_out)[i] = pack8(dz);
      float o[8];
      if (SE) {
        float g[8], qq[8];
        load8f(se + (n * CG + cg) * 8, g);
        load8f(q + (n * CG + cg) * 8, qq);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], fmaf(g[j], dz[j], qq[j]), fmaf(B[j], xv[j], D[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], dz[j], fmaf(B[j], xv[j], D[j]));
      }
      reinterpret_cast<uint4*>(dx)[i] = pack8(o);
    }
  }
}

// bn_bwd_apply_kernel<0 | 3, false> for C/8 a power of two <= 256 (see bn_apply_fast_kernel): the three coefficient
// vectors in registers, no divisions; the ReLU bit mask is applied to the PACKED gradient words (relu_mask_words), so
// the dz output is the masked input words as they are (no unpack / select / repack).  Same arithmetic per element.
template <bool MASK, bool DZ>
__global__ void __launch_bounds__(256) bn_bwd_apply_fast_kernel(
    const uint4* __restrict__ x, const uint4* __restrict__ dy, const uint8_t* __restrict__ mask,
    const float* __restrict__ coefA, const float* __restrict__ coefB, const float* __restrict__ coefD,
    uint4* __restrict__ dx, uint4* __restrict__ dz_out, int CG, size_t total_vec) {
  constexpr int U = 2;
  const size_t stride = (size_t)gridDim.x * 256;
  const size_t i0 = blockIdx.x * (size_t)256 + threadIdx.x;
  const int cg = (int)(i0 & (size_t)(CG - 1));
  float A[8], B[8], D[8];
  load8f(coefA + cg * 8, A);
  load8f(coefB + cg * 8, B);
  load8f(coefD + cg * 8, D);
  for (size_t i = i0; i < total_vec; i += U * stride) {
    uint4 vx[U], vd[U];
    uint32_t vm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t k = i + u * stride;
      if (k < total_vec) {
        if (MASK) vm[u] = mask[k];
        vx[u] = ld_stream(x + k);
        vd[u] = ld_stream(dy + k);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t k = i + u * stride;
      if (k >= total_vec) break;
      if (MASK) relu_mask_words(vm[u], vd[u]);
      if (DZ) dz_out[k] = vd[u];
      float xv[8], dz[8], o[8];
      unpack8(vx[u], xv);
      unpack8(vd[u], dz);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], dz[j], fmaf(B[j], xv[j], D[j]));
      dx[k] = pack8(o);
    }
  }
}

// bn_bwd_apply_fast_kernel with x and dy staged through a thread-private cp.async ring (vec.cuh).  ncu showed the
// register version waiting on the long scoreboard at 48 % occupancy with 25 % of the issue slots used: the bytes a
// thread can have in flight are bounded by the registers that receive them, and every warp alternates "wait for the
// trip's loads" and "compute".  Here a thread keeps kAsyncStages - 1 future vectors in flight in shared memory (3 CTAs x
// 256 threads x 7 x 32 bytes = 172 KB per SM) while it computes on the oldest one; the mask bytes (too small for
// cp.async) ride in a register ring, loaded S - 1 vectors ahead as well.  Same arithmetic per element.  Measured
// (profiles/r02ii_*): backward-apply -1 ... -12 %, the reduction below -4 ... -16 % (1.00 of the measured HBM peak at
// batch 512); the forward apply kernel built the same way was 3-8 % SLOWER than its register version and was dropped.
constexpr int kAsyncStages = 8;

template <bool MASK, bool DZ>
__global__ void __launch_bounds__(256, 3) bn_bwd_apply_async_kernel(
    const uint4* __restrict__ x, const uint4* __restrict__ dy, const uint8_t* __restrict__ mask,
    const float* __restrict__ coefA, const float* __restrict__ coefB, const float* __restrict__ coefD,
    uint4* __restrict__ dx, uint4* __restrict__ dz_out, int CG, size_t total_vec) {
  constexpr int S = kAsyncStages;
  extern __shared__ uint4 ring[];  // [S][2][256]
  const size_t stride = (size_t)gridDim.x * 256;
  const size_t i0 = blockIdx.x * (size_t)256 + threadIdx.x;
  const int cg = (int)(i0 & (size_t)(CG - 1));
  float A[8], B[8], D[8];
  load8f(coefA + cg * 8, A);
  load8f(coefB + cg * 8, B);
  load8f(coefD + cg * 8, D);
  uint4* slot0 = ring + threadIdx.x;
  uint32_t mreg[S];
#pragma unroll
  for (int s = 0; s < S; ++s) mreg[s] = 0;
#pragma unroll
  for (int s = 0; s < S - 1; ++s) {
    const size_t k = i0 + s * stride;
    if (k < total_vec) {
      if (MASK) mreg[s] = mask[k];
      cp_async16(slot0 + (2 * s) * 256, x + k);
      cp_async16(slot0 + (2 * s + 1) * 256, dy + k);
    }
    cp_async_commit();
  }
  for (size_t base = i0; base < total_vec; base += S * stride) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const size_t k = base + s * stride;
      if (k >= total_vec) break;
      const size_t kn = k + (S - 1) * stride;
      const int sn = (s + S - 1) % S;
      if (kn < total_vec) {
        if (MASK) mreg[sn] = mask[kn];
        cp_async16(slot0 + (2 * sn) * 256, x + kn);
        cp_async16(slot0 + (2 * sn + 1) * 256, dy + kn);
      }
      cp_async_commit();
      cp_async_wait<S - 1>();
      const uint4 vx = slot0[(2 * s) * 256];
      uint4 vd = slot0[(2 * s + 1) * 256];
      if (MASK) relu_mask_words(mreg[s], vd);
      if (DZ) dz_out[k] = vd;
      float xv[8], dz[8], o[8];
      unpack8(vx, xv);
      unpack8(vd, dz);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], dz[j], fmaf(B[j], xv[j], D[j]));
      dx[k] = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Stem backward (max-pool 3x3/s2/p1 behind BN+ReLU), organised by 2x2 INPUT blocks: block (a, b)
// = pixels (2a+i, 2b+j) is touched only by the pooling windows (a..a+1, b..b+1), so one thread
// loads 4 x vectors + 4 argmax words + 4 pooled-gradient vectors (12 independent loads) and
// produces the gradient of all 4 pixels.  Window position codes (dh*3+dw) are compile-time.
// ------------------------------------------------------------------------------------------
// raw (still packed) operands of one 2x2 block for 8 channels: 12 independent loads
struct StemRaw {
  uint4 vx[4];  // x at pixel t = 2*i + j
  uint4 vg[4];  // pooled gradient of window t = 2*wi + wj (zeroed when the window does not exist)
  uint2 va[4];  // argmax codes of window t
  bool ok[4];   // pixel t inside the image
};

__device__ __forceinline__ void stem_raw_load(const __nv_bfloat16* __restrict__ x,
                                              const __nv_bfloat16* __restrict__ dyp,
                                              const uint8_t* __restrict__ arg, size_t n, int a, int b, int H, int W,
                                              int Ho, int Wo, int CG, int cg, StemRaw& raw) {
  // one 64-bit base per tensor, the 4 pixels / windows are small constant offsets from it
  const uint4* xb = reinterpret_cast<const uint4*>(x) + ((n * H + 2 * a) * W + 2 * b) * CG + cg;
  const size_t ob = ((n * Ho + a) * Wo + b) * CG + cg;
  const uint4* gb = reinterpret_cast<const uint4*>(dyp) + ob;
  const uint2* ab = reinterpret_cast<const uint2*>(arg) + ob;
  const bool h1 = 2 * a + 1 < H, w1 = 2 * b + 1 < W;    // second pixel row / column inside the image
  const bool oh1 = a + 1 < Ho, ow1 = b + 1 < Wo;        // second window row / column exists
  const int xrow = W * CG, orow = Wo * CG;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int i = t >> 1, j = t & 1;
    raw.ok[t] = (i == 0 || h1) && (j == 0 || w1);
    raw.vx[t] = raw.ok[t] ? ld_stream(xb + i * xrow + j * CG) : make_uint4(0, 0, 0, 0);
    const bool wok = (i == 0 || oh1) && (j == 0 || ow1);
    const int off = wok ? i * orow + j * CG : 0;
    raw.vg[t] = gb[off];
    raw.va[t] = ab[off];
    if (!wok) raw.vg[t] = make_uint4(0, 0, 0, 0);
  }
}

__device__ __forceinline__ uint32_t word_of(const uint4& v, int k) {
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

// Channels 2k, 2k+1 of the block: x[t][2] and the routed + ReLU-gated gradient dz[t][2] of its 4 pixels.
// code of pixel (i,j) inside window (a+wi, b+wj): dh = i - 2*wi + 1, dw = j - 2*wj + 1 (compile-time table).
__device__ __forceinline__ void stem_pair(const StemRaw& raw, int k, const float* sc, const float* sh,
                                          float (&xv)[4][2], float (&dz)[4][2]) {
  float g[4][2];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const uint32_t wx = word_of(raw.vx[t], k), wg = word_of(raw.vg[t], k);
    const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wx));
    const float2 fg = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wg));
    xv[t][0] = fx.x;
    xv[t][1] = fx.y;
    g[t][0] = fg.x;
    g[t][1] = fg.y;
  }
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int c = 2 * k + e;
    int code[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) code[t] = ((c < 4 ? raw.va[t].x : raw.va[t].y) >> (8 * (c & 3))) & 0xff;
    // window 0 = (a,b), 1 = (a,b+1), 2 = (a+1,b), 3 = (a+1,b+1)
    const float d00 = (code[0] == 4) ? g[0][e] : 0.f;
    const float d01 = ((code[0] == 5) ? g[0][e] : 0.f) + ((code[1] == 3) ? g[1][e] : 0.f);
    const float d10 = ((code[0] == 7) ? g[0][e] : 0.f) + ((code[2] == 1) ? g[2][e] : 0.f);
    const float d11 = ((code[0] == 8) ? g[0][e] : 0.f) + ((code[1] == 6) ? g[1][e] : 0.f) +
                      ((code[2] == 2) ? g[2][e] : 0.f) + ((code[3] == 0) ? g[3][e] : 0.f);
    dz[0][e] = (raw.ok[0] && fmaf(xv[0][e], sc[c], sh[c]) > 0.f) ? d00 : 0.f;
    dz[1][e] = (raw.ok[1] && fmaf(xv[1][e], sc[c], sh[c]) > 0.f) ? d01 : 0.f;
    dz[2][e] = (raw.ok[2] && fmaf(xv[2][e], sc[c], sh[c]) > 0.f) ? d10 : 0.f;
    dz[3][e] = (raw.ok[3] && fmaf(xv[3][e], sc[c], sh[c]) > 0.f) ? d11 : 0.f;
  }
}

// p1/p2 [N][SPLIT][C]; the slab unit is a 2x2 block (= one pooled position).  Accumulates sum dz and
// sum dz*x; xhat is applied once per CTA: sum dz*xhat = (sum dz*x - mean * sum dz) * invstd.
__global__ void __launch_bounds__(kRedThreads, 2) stem_bwd_reduce_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dyp, const uint8_t* __restrict__ arg,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ scale,
    const float* __restrict__ shift, float* __restrict__ p1, float* __restrict__ p2, int C, int blocks_per_split,
    int H, int W, int Ho, int Wo) {
  extern __shared__ float sred[];
  const int CG = C >> 3;
  const int rows = kRedThreads / CG;
  const int cg = threadIdx.x % CG, r = threadIdx.x / CG;
  const int n = blockIdx.y, split = blockIdx.x;
  const int nb = Ho * Wo;
  const int pa = split * blocks_per_split;
  const int pb = min(nb, pa + blocks_per_split);
  float sc[8], sh[8], s[8], q[8];
  load8f(scale + cg * 8, sc);
  load8f(shift + cg * 8, sh);
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  for (int p = pa + r; p < pb; p += rows) {
    StemRaw raw;
    stem_raw_load(x, dyp, arg, n, p / Wo, p % Wo, H, W, Ho, Wo, CG, cg, raw);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float xv[4][2], dz[4][2];
      stem_pair(raw, k, sc, sh, xv, dz);
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          s[2 * k + e] += dz[t][e];
          q[2 * k + e] = fmaf(dz[t][e], xv[t][e], q[2 * k + e]);
        }
    }
  }
  float* ss = sred;
  float* sq = sred + rows * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ss[r * C + cg * 8 + j] = s[j];
    sq[r * C + cg * 8 + j] = q[j];
  }
  __syncthreads();
  const size_t orow = ((size_t)n * gridDim.x + split) * C;
  for (int c = threadIdx.x; c < C; c += kRedThreads) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < rows; ++i) {
      a += ss[i * C + c];
      b += sq[i * C + c];
    }
    p1[orow + c] = a;
    p2[orow + c] = (b - mean[c] * a) * invstd[c];
  }
}

// Routed gradient of the 4 pixels of a block for all 8 channels, as packed bf16 pairs dzw[pixel][pair]:
// the argmax codes of 8 channels are 8 bytes, so "window t selected position K" is ONE byte-wise SIMD
// compare per 4 channels; the byte masks are widened to 16-bit lanes and ANDed onto the bf16 gradients.
__device__ __forceinline__ void stem_route(const StemRaw& raw, uint32_t (&dzw)[4][4]) {
  uint32_t g[4][4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    g[t][0] = raw.vg[t].x;
    g[t][1] = raw.vg[t].y;
    g[t][2] = raw.vg[t].z;
    g[t][3] = raw.vg[t].w;
  }
  auto sel = [&](int t, uint32_t code, uint32_t (&out)[4]) {
    const uint32_t k4 = code * 0x01010101u;
    const uint32_t m0 = __vcmpeq4(raw.va[t].x, k4), m1 = __vcmpeq4(raw.va[t].y, k4);
    out[0] = g[t][0] & __byte_perm(m0, 0, 0x1100);
    out[1] = g[t][1] & __byte_perm(m0, 0, 0x3322);
    out[2] = g[t][2] & __byte_perm(m1, 0, 0x1100);
    out[3] = g[t][3] & __byte_perm(m1, 0, 0x3322);
  };
  auto add = [](uint32_t (&acc)[4], const uint32_t (&v)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&acc[k]),
                                       *reinterpret_cast<const __nv_bfloat162*>(&v[k]));
      acc[k] = *reinterpret_cast<const uint32_t*>(&r);
    }
  };
  uint32_t tmp[4];
  // pixel (0,0): window 0 code 4
  sel(0, 4, dzw[0]);
  // pixel (0,1): window 0 code 5, window 1 code 3
  sel(0, 5, dzw[1]);
  sel(1, 3, tmp);
  add(dzw[1], tmp);
  // pixel (1,0): window 0 code 7, window 2 code 1
  sel(0, 7, dzw[2]);
  sel(2, 1, tmp);
  add(dzw[2], tmp);
  // pixel (1,1): window 0 code 8, window 1 code 6, window 2 code 2, window 3 code 0
  sel(0, 8, dzw[3]);
  sel(1, 6, tmp);
  add(dzw[3], tmp);
  sel(2, 2, tmp);
  add(dzw[3], tmp);
  sel(3, 0, tmp);
  add(dzw[3], tmp);
}

__global__ void __launch_bounds__(256, 2) stem_bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dyp, const uint8_t* __restrict__ arg,
    const float* __restrict__ coefA, const float* __restrict__ coefB, const float* __restrict__ coefD,
    const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ dx, int CG,
    size_t total_blocks_vec, int H, int W, int Ho, int Wo) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  // CG divides the grid stride (a power of two <= 256), so a thread keeps its channel group for the whole loop and
  // the five per-channel coefficient vectors live in registers: the kernel was issue-bound (~1040 instructions per
  // 2x2 block, 80 of them coefficient re-loads inside the pixel loop), not bandwidth-bound.
  const int cg = (int)(i0 % CG);
  float A[8], B[8], D[8], sc[8], sh[8];
  load8f(coefA + cg * 8, A);
  load8f(coefB + cg * 8, B);
  load8f(coefD + cg * 8, D);
  load8f(scale + cg * 8, sc);
  load8f(shift + cg * 8, sh);
  for (size_t i = i0; i < total_blocks_vec; i += stride) {
    size_t t = i / CG;
    const int b = (int)(t % Wo);
    t /= Wo;
    const int a = (int)(t % Ho);
    const size_t n = t / Ho;
    StemRaw raw;
    stem_raw_load(x, dyp, arg, n, a, b, H, W, Ho, Wo, CG, cg, raw);
    uint32_t dzw[4][4];
    stem_route(raw, dzw);
    uint4* dxb = reinterpret_cast<uint4*>(dx) + ((n * H + 2 * a) * W + 2 * b) * CG + cg;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      if (!raw.ok[px]) continue;
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t xw = word_of(raw.vx[px], k);
        const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw));
        float2 fz = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dzw[px][k]));
        if (fmaf(fx.x, sc[2 * k], sh[2 * k]) <= 0.f) fz.x = 0.f;  // ReLU gate of the pre-pool activation
        if (fmaf(fx.y, sc[2 * k + 1], sh[2 * k + 1]) <= 0.f) fz.y = 0.f;
        const __nv_bfloat162 v =
            __floats2bfloat162_rn(fmaf(A[2 * k], fz.x, fmaf(B[2 * k], fx.x, D[2 * k])),
                                  fmaf(A[2 * k + 1], fz.y, fmaf(B[2 * k + 1], fx.y, D[2 * k + 1])));
        o[k] = *reinterpret_cast<const uint32_t*>(&v);
      }
      dxb[(px >> 1) * (W * CG) + (px & 1) * CG] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// The same operation with one CTA per (sample, pooled row): the thread index inside the row is (b, cg) by a shift, so
// the grid-stride kernel's 64-bit divisions (i / CG, % Wo, / Ho per 2x2 block) disappear, and the row bases of the
// three tensors are loop invariants.  VARIANT 1 keeps the five coefficient vectors in registers (2 CTAs per SM);
// VARIANT 2 keeps them in shared memory and fits 3 CTAs per SM (the kernel is latency-, not issue-bound: 16 warps per
// SM issued 44 % of the slots at 0.64 of the HBM peak).  Same arithmetic per element as stem_bwd_apply_kernel.
template <int VARIANT>
__global__ void __launch_bounds__(256, VARIANT == 2 ? 3 : 2) stem_bwd_apply_rows_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dyp, const uint8_t* __restrict__ arg,
    const float* __restrict__ coefA, const float* __restrict__ coefB, const float* __restrict__ coefD,
    const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ dx, int C,
    int cg_shift, int H, int W, int Ho, int Wo) {
  extern __shared__ float scoef[];  // VARIANT 2: [5][C] = A, B, D, scale, shift
  const int CG = C >> 3;
  const int a = (int)(blockIdx.x % (unsigned)Ho);
  const size_t n = blockIdx.x / (unsigned)Ho;
  const int cg = threadIdx.x & (CG - 1);
  float A[8], B[8], D[8], sc[8], sh[8];
  if (VARIANT == 2) {
    for (int c = threadIdx.x; c < C; c += 256) {
      scoef[c] = coefA[c];
      scoef[C + c] = coefB[c];
      scoef[2 * C + c] = coefD[c];
      scoef[3 * C + c] = scale[c];
      scoef[4 * C + c] = shift[c];
    }
    __syncthreads();
  } else {
    load8f(coefA + cg * 8, A);
    load8f(coefB + cg * 8, B);
    load8f(coefD + cg * 8, D);
    load8f(scale + cg * 8, sc);
    load8f(shift + cg * 8, sh);
  }
  const int items = Wo << cg_shift;
  for (int j = threadIdx.x; j < items; j += 256) {
    const int b = j >> cg_shift;
    StemRaw raw;
    stem_raw_load(x, dyp, arg, n, a, b, H, W, Ho, Wo, CG, cg, raw);
    uint32_t dzw[4][4];
    stem_route(raw, dzw);
    uint32_t o[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 cA, cB, cD, cS, cH;
      if (VARIANT == 2) {
        const float* base = scoef + cg * 8 + 2 * k;
        cA = *reinterpret_cast<const float2*>(base);
        cB = *reinterpret_cast<const float2*>(base + C);
        cD = *reinterpret_cast<const float2*>(base + 2 * C);
        cS = *reinterpret_cast<const float2*>(base + 3 * C);
        cH = *reinterpret_cast<const float2*>(base + 4 * C);
      } else {
        cA = make_float2(A[2 * k], A[2 * k + 1]);
        cB = make_float2(B[2 * k], B[2 * k + 1]);
        cD = make_float2(D[2 * k], D[2 * k + 1]);
        cS = make_float2(sc[2 * k], sc[2 * k + 1]);
        cH = make_float2(sh[2 * k], sh[2 * k + 1]);
      }
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        const uint32_t xw = word_of(raw.vx[px], k);
        const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw));
        float2 fz = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dzw[px][k]));
        if (fmaf(fx.x, cS.x, cH.x) <= 0.f) fz.x = 0.f;  // ReLU gate of the pre-pool activation
        if (fmaf(fx.y, cS.y, cH.y) <= 0.f) fz.y = 0.f;
        const __nv_bfloat162 v = __floats2bfloat162_rn(fmaf(cA.x, fz.x, fmaf(cB.x, fx.x, cD.x)),
                                                       fmaf(cA.y, fz.y, fmaf(cB.y, fx.y, cD.y)));
        o[px][k] = *reinterpret_cast<const uint32_t*>(&v);
      }
    }
    uint4* dxb = reinterpret_cast<uint4*>(dx) + ((n * H + 2 * a) * W + 2 * b) * CG + cg;
#pragma unroll
    for (int px = 0; px < 4; ++px)
      if (raw.ok[px]) dxb[(px >> 1) * (W * CG) + (px & 1) * CG] = make_uint4(o[px][0], o[px][1], o[px][2], o[px][3]);
  }
}

// ------------------------------------------------------------------------------------------
// global average pool  [N][P][C] bf16 -> [N][C] fp32   and its backward (broadcast / P)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRedThreads) avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                                   float* __restrict__ out, int P, int C) {
  extern __shared__ float sred[];
  const int CG = C >> 3;
  const int rows = kRedThreads / CG;
  const int cg = threadIdx.x % CG, r = threadIdx.x / CG;
  const int n = blockIdx.x;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  const uint4* base = reinterpret_cast<const uint4*>(x + (size_t)n * P * C) + cg;
  for (int p = r; p < P; p += rows) {
    float f[8];
    unpack8(ld_stream(base + (size_t)p * CG), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += f[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sred[r * C + cg * 8 + j] = s[j];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kRedThreads) {
    float a = 0.f;
    for (int i = 0; i < rows; ++i) a += sred[i * C + c];
    out[(size_t)n * C + c] = a / (float)P;
  }
}

__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dout,
                                                           __nv_bfloat16* __restrict__ dx, int CG, float inv_p,
                                                           size_t vec_per_sample, size_t total_vec) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total_vec; i += stride) {
    const int cg = (int)(i % CG);
    const size_t n = i / vec_per_sample;
    float g[8];
    load8f(dout + (n * CG + cg) * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= inv_p;
    reinterpret_cast<uint4*>(dx)[i] = pack8(g);
  }
}

// Grid for a grid-stride streaming kernel: exactly one wave of resident CTAs (num_sms x occupancy of
// THAT kernel), so no partially filled last wave; fewer CTAs when the tensor is small.
template <typename K>
static int stream_grid(size_t total_vec, K kernel, int vec_per_thread = 1) {
  static std::unordered_map<const void*, int> cache;  // one host thread per GPU process (header contract)
  int& occ = cache[reinterpret_cast<const void*>(kernel)];
  if (occ == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0) != cudaSuccess || occ < 1))
    occ = 2;
  const size_t cap = (size_t)num_sms() * occ;
  size_t b = (total_vec + (size_t)256 * vec_per_thread - 1) / ((size_t)256 * vec_per_thread);
  return (int)(b < cap ? (b ? b : 1) : cap);
}

// ECGMM_BN_ASYNC=0: the register versions of the BatchNorm backward-apply / backward-reduce kernels (no cp.async ring)
static bool bn_async_enabled() {
  const char* e = getenv("ECGMM_BN_ASYNC");
  return !(e && e[0] == '0');
}

// Persistent grid of a kernel that needs `bytes` of dynamic shared memory per CTA: opt in above 48 KB once, then
// SMs x resident CTAs (no partially filled last wave, as stream_grid does for the register kernels).
template <typename K>
static int async_grid_cap(K kernel, size_t bytes) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, bytes) != cudaSuccess || occ < 1) occ = 1;
  return num_sms() * occ;
}

// ECGMM_BN_FAST=0: the generic BatchNorm apply / backward-apply kernels everywhere
static bool bn_fast_enabled() {
  const char* e = getenv("ECGMM_BN_FAST");
  return !(e && e[0] == '0');
}

static int check_c(int C, const char* who) {
  ECGMM_CHECK(C >= 8 && C % 8 == 0 && C <= 2048 && (kRedThreads % (C >> 3) == 0 || (C >> 3) % kRedThreads == 0),
              ECGMM_ERR_SHAPE, "%s: channel count %d must be 8 * a power of two <= 2048", who, C);
  ECGMM_CHECK((C >> 3) <= kRedThreads, ECGMM_ERR_SHAPE, "%s: channel count %d too large", who, C);
  return ECGMM_OK;
}

}  // namespace ecgmm

using namespace ecgmm;
typedef __nv_bfloat16 bf16;

extern "C" int ecgmm_reduce_split(int N, int P, int C) {
  if (N <= 0 || P <= 0 || C < 8) return 1;
  const int rows = kRedThreads / (C >> 3) > 0 ? kRedThreads / (C >> 3) : 1;
  // many small CTAs (32 per SM): with fewer, longer CTAs the partially filled last wave costs up to 30 % (measured:
  // 8 per SM doubled the statistics pass at batch 512)
  const int want = ceil_div(num_sms() * 32, N);
  int max_split = ceil_div(P, rows * 8);
  if (max_split < 1) max_split = 1;
  int s = want < 1 ? 1 : want;
  if (s > max_split) s = max_split;
  return s;
}

static inline int rows_per_split(int P, int split) { return ceil_div(P, split); }

extern "C" int ecgmm_chan_stats(const ecgmm_bf16* x, float* psum, float* psq, int N, int P, int C, int split,
                                void* stream) {
  ECGMM_CHECK(x && psum && psq, ECGMM_ERR_ARG, "chan_stats: null pointer");
  int rc = check_c(C, "chan_stats");
  if (rc) return rc;
  ECGMM_CHECK(split >= 1 && N <= 65535, ECGMM_ERR_SHAPE, "chan_stats: bad split %d / batch %d", split, N);
  if (N == 0) return ECGMM_OK;
  const int rows = kRedThreads / (C >> 3);
  const size_t smem = (size_t)2 * rows * C * sizeof(float);
  chan_stats_kernel<<<dim3(split, N), kRedThreads, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(x), psum, psq, P, C, rows_per_split(P, split));
  return check_launch("chan_stats_kernel");
}

extern "C" int ecgmm_bn_finalize(const float* psum, const float* psq, int N, int split, int C, long long count,
                                 const float* gamma, const float* beta, const float* conv_bias, float eps,
                                 float momentum, float* running_mean, float* running_var, long long* num_batches,
                                 float* mean, float* invstd, float* scale, float* shift, float* nsum,
                                 void* stream) {
  ECGMM_CHECK(psum && psq && mean && invstd && scale && shift, ECGMM_ERR_ARG, "bn_finalize: null pointer");
  ECGMM_CHECK(count > 0, ECGMM_ERR_SHAPE, "bn_finalize: empty batch");
  bn_finalize_kernel<<<ceil_div(C, kFinCh), kFinCh * kFinLanes, 0, (cudaStream_t)stream>>>(
      psum, psq, N * split, split, C, (double)count, gamma, beta, conv_bias, eps, momentum, running_mean,
      running_var, num_batches, mean, invstd, scale, shift, nsum);
  return check_launch("bn_finalize_kernel");
}

extern "C" int ecgmm_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* conv_bias,
                                    const float* running_mean, const float* running_var, float eps, float* scale,
                                    float* shift, void* stream) {
  ECGMM_CHECK(running_mean && running_var && scale && shift, ECGMM_ERR_ARG, "bn_eval_coeffs: null pointer");
  bn_eval_coeffs_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(C, gamma, beta, conv_bias, running_mean,
                                                                           running_var, eps, scale, shift);
  return check_launch("bn_eval_coeffs_kernel");
}

extern "C" int ecgmm_bn_apply(const ecgmm_bf16* x_, const float* scale, const float* shift, const float* se,
                              const ecgmm_bf16* res_, ecgmm_bf16* y_, uint8_t* mask_out, int N, int P, int C,
                              int relu, void* stream) {
  ECGMM_CHECK(x_ && scale && shift && y_, ECGMM_ERR_ARG, "bn_apply: null pointer");
  ECGMM_CHECK(C % 8 == 0, ECGMM_ERR_SHAPE, "bn_apply: C=%d not a multiple of 8", C);
  const size_t vps = (size_t)P * (C >> 3), total = vps * N;
  if (total == 0) return ECGMM_OK;
  const bf16* x = reinterpret_cast<const bf16*>(x_);
  const bf16* res = reinterpret_cast<const bf16*>(res_);
  bf16* y = reinterpret_cast<bf16*>(y_);
  cudaStream_t st = (cudaStream_t)stream;
  const int CG = C >> 3;
  // ECGMM_BN_FAST=0: the generic kernels everywhere
  if (!se && CG <= 256 && (CG & (CG - 1)) == 0 && bn_fast_enabled()) {
    const uint4* x4 = reinterpret_cast<const uint4*>(x);
    const uint4* r4 = reinterpret_cast<const uint4*>(res);
    uint4* y4 = reinterpret_cast<uint4*>(y);
#define ECGMM_APPLY_FAST(RES_, RELU_)                                                                          \
  bn_apply_fast_kernel<RES_, RELU_><<<stream_grid(total, bn_apply_fast_kernel<RES_, RELU_>, 2), 256, 0, st>>>( \
      x4, scale, shift, r4, y4, mask_out, CG, total)
    if (res && relu)
      ECGMM_APPLY_FAST(true, true);
    else if (res)
      ECGMM_APPLY_FAST(true, false);
    else if (relu)
      ECGMM_APPLY_FAST(false, true);
    else
      ECGMM_APPLY_FAST(false, false);
#undef ECGMM_APPLY_FAST
    return check_launch("bn_apply_fast_kernel");
  }
#define ECGMM_APPLY(SE_, RES_, RELU_)                                                                 \
  bn_apply_kernel<SE_, RES_, RELU_><<<stream_grid(total, bn_apply_kernel<SE_, RES_, RELU_>), 256, 0, st>>>( \
      x, scale, shift, se, res, y, mask_out, CG, vps, total)
  const int key = (se ? 4 : 0) | (res ? 2 : 0) | (relu ? 1 : 0);
  switch (key) {
    case 0: ECGMM_APPLY(false, false, false); break;
    case 1: ECGMM_APPLY(false, false, true); break;
    case 2: ECGMM_APPLY(false, true, false); break;
    case 3: ECGMM_APPLY(false, true, true); break;
    case 4: ECGMM_APPLY(true, false, false); break;
    case 5: ECGMM_APPLY(true, false, true); break;
    case 6: ECGMM_APPLY(true, true, false); break;
    default: ECGMM_APPLY(true, true, true); break;
  }
#undef ECGMM_APPLY
  return check_launch("bn_apply_kernel");
}

extern "C" int ecgmm_bn_relu_maxpool(const ecgmm_bf16* x, const float* scale, const float* shift, ecgmm_bf16* y,
                                     uint8_t* argmax, int N, int H, int W, int C, void* stream) {
  ECGMM_CHECK(x && scale && shift && y, ECGMM_ERR_ARG, "bn_relu_maxpool: null pointer");
  ECGMM_CHECK(C % 8 == 0, ECGMM_ERR_SHAPE, "bn_relu_maxpool: C=%d not a multiple of 8", C);
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const size_t total = (size_t)N * Ho * Wo * (C >> 3);
  if (total == 0) return ECGMM_OK;
  if (C == 64 && H >= 2 * kPoolTOH && W >= 2 * kPoolTOW && !getenv("ECGMM_POOL_LEGACY")) {
    const int tiles_w = (Wo + kPoolTOW - 1) / kPoolTOW, tiles_h = (Ho + kPoolTOH - 1) / kPoolTOH;
    const long long blocks = (long long)N * tiles_h * tiles_w;
    ECGMM_CHECK(blocks <= 0x7fffffffLL, ECGMM_ERR_SHAPE, "bn_relu_maxpool: extent");
    if (getenv("ECGMM_POOL_TILED_V1") || (long long)H * W * 8 > 0x7fffffffLL)  // v2 indexes a sample with 32 bits
      bn_relu_maxpool_tiled_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const bf16*>(x), scale, shift, reinterpret_cast<bf16*>(y), argmax, H, W, Ho, Wo, tiles_w,
          tiles_h);
    else
      bn_relu_maxpool_tiled_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const bf16*>(x), scale, shift, reinterpret_cast<bf16*>(y), argmax, H, W, Ho, Wo, tiles_w,
          tiles_h);
    return check_launch("bn_relu_maxpool_tiled_kernel");
  }
  bn_relu_maxpool_kernel<<<stream_grid(total, bn_relu_maxpool_kernel), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(x), scale, shift, reinterpret_cast<bf16*>(y), argmax, H, W, Ho, Wo, C >> 3,
      total);
  return check_launch("bn_relu_maxpool_kernel");
}

extern "C" int ecgmm_bn_bwd_reduce(const ecgmm_bf16* x, const ecgmm_bf16* dy, const ecgmm_bf16* y,
                                   const uint8_t* argmax, const float* mean, const float* invstd,
                                   const float* scale, const float* shift, float* p1, float* p2, int N, int H,
                                   int W, int C, int split, int mode, void* stream) {
  ECGMM_CHECK(x && dy && mean && invstd && p1 && p2, ECGMM_ERR_ARG, "bn_bwd_reduce: null pointer");
  ECGMM_CHECK(mode >= 0 && mode <= 4, ECGMM_ERR_ARG, "bn_bwd_reduce: bad mode %d", mode);
  ECGMM_CHECK(mode != 3 || argmax, ECGMM_ERR_ARG, "bn_bwd_reduce: mode 3 needs the bit mask");
  ECGMM_CHECK(mode != 1 || y, ECGMM_ERR_ARG, "bn_bwd_reduce: mode 1 needs y");
  ECGMM_CHECK(mode != 2 || (argmax && scale && shift), ECGMM_ERR_ARG, "bn_bwd_reduce: mode 2 needs argmax/scale/shift");
  int rc = check_c(C, "bn_bwd_reduce");
  if (rc) return rc;
  ECGMM_CHECK(split >= 1 && N <= 65535, ECGMM_ERR_SHAPE, "bn_bwd_reduce: bad split %d / batch %d", split, N);
  if (N == 0) return ECGMM_OK;
  const int P = H * W;
  const int rows = kRedThreads / (C >> 3);
  const size_t smem = (size_t)2 * rows * C * sizeof(float);
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  dim3 grid(split, N);
  cudaStream_t st = (cudaStream_t)stream;
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  const bf16* dyb = reinterpret_cast<const bf16*>(dy);
  const bf16* yb = reinterpret_cast<const bf16*>(y);
  const int rps = rows_per_split(P, split);
  if ((mode == 0 || mode == 3) && bn_async_enabled()) {
    const size_t bytes = (size_t)kRedStages * 2 * kRedThreads * sizeof(uint4) + smem;
    if (mode == 0) {
      static const int once = (cudaFuncSetAttribute(bn_bwd_reduce_async_kernel<0>,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024), 0);
      (void)once;
      bn_bwd_reduce_async_kernel<0><<<grid, kRedThreads, bytes, st>>>(xb, dyb, argmax, mean, invstd, p1, p2, P, C, rps);
    } else {
      static const int once = (cudaFuncSetAttribute(bn_bwd_reduce_async_kernel<3>,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024), 0);
      (void)once;
      bn_bwd_reduce_async_kernel<3><<<grid, kRedThreads, bytes, st>>>(xb, dyb, argmax, mean, invstd, p1, p2, P, C, rps);
    }
    return check_launch("bn_bwd_reduce_async_kernel");
  }
  if (mode == 0)
    bn_bwd_reduce_kernel<0><<<grid, kRedThreads, smem, st>>>(xb, dyb, yb, argmax, mean, invstd, scale, shift, p1, p2,
                                                             P, C, rps, H, W, Ho, Wo);
  else if (mode == 1)
    bn_bwd_reduce_kernel<1><<<grid, kRedThreads, smem, st>>>(xb, dyb, yb, argmax, mean, invstd, scale, shift, p1, p2,
                                                             P, C, rps, H, W, Ho, Wo);
  else if (mode == 3)
    bn_bwd_reduce_kernel<3><<<grid, kRedThreads, smem, st>>>(xb, dyb, yb, argmax, mean, invstd, scale, shift, p1, p2,
                                                             P, C, rps, H, W, Ho, Wo);
  else if (mode == 4)
    bn_bwd_reduce_kernel<4><<<grid, kRedThreads, smem, st>>>(xb, dyb, yb, argmax, mean, invstd, scale, shift, p1, p2,
                                                             P, C, rps, H, W, Ho, Wo);
  else
    stem_bwd_reduce_kernel<<<grid, kRedThreads, smem, st>>>(xb, dyb, argmax, mean, invstd, scale, shift, p1, p2, C,
                                                            rows_per_split(Ho * Wo, split), H, W, Ho, Wo);
  return check_launch("bn_bwd_reduce_kernel");
}

extern "C" int ecgmm_bn_bwd_finalize(const float* p1, const float* p2, int N, int split, int C,
                                     long long per_sample, const float* gamma, const float* mean,
                                     const float* invstd, const float* se, const float* q, const float* nsum,
                                     float* dgamma, float* dbeta, float* coefA, float* coefB, float* coefD,
                                     long long count, void* stream) {
  ECGMM_CHECK(count >= 0 && (count == 0 || !se), ECGMM_ERR_ARG, "bn_bwd_finalize: bad count %lld", count);
  ECGMM_CHECK(p1 && p2 && mean && invstd && coefA && coefB && coefD, ECGMM_ERR_ARG, "bn_bwd_finalize: null pointer");
  ECGMM_CHECK(!se || (q && nsum), ECGMM_ERR_ARG, "bn_bwd_finalize: SE mode needs q and nsum");
  ECGMM_CHECK(N > 0 && per_sample > 0, ECGMM_ERR_SHAPE, "bn_bwd_finalize: empty batch");
  bn_bwd_finalize_kernel<<<ceil_div(C, kFinCh), kFinCh * kFinLanes, 0, (cudaStream_t)stream>>>(
      p1, p2, N, split, C, (double)per_sample, (double)count, gamma, mean, invstd, se, q, nsum, dgamma, dbeta, coefA,
      coefB, coefD);
  return check_launch("bn_bwd_finalize_kernel");
}

extern "C" int ecgmm_bn_bwd_apply(const ecgmm_bf16* x, const ecgmm_bf16* dy, const ecgmm_bf16* y,
                                  const uint8_t* argmax, const float* coefA, const float* coefB, const float* coefD,
                                  const float* scale, const float* shift, const float* se, const float* q,
                                  ecgmm_bf16* dx, ecgmm_bf16* dz_out, int N, int H, int W, int C, int mode,
                                  void* stream) {
  ECGMM_CHECK(x && dy && coefA && coefB && coefD && dx, ECGMM_ERR_ARG, "bn_bwd_apply: null pointer");
  ECGMM_CHECK(mode >= 0 && mode <= 3, ECGMM_ERR_ARG, "bn_bwd_apply: bad mode %d", mode);
  ECGMM_CHECK(mode != 3 || argmax, ECGMM_ERR_ARG, "bn_bwd_apply: mode 3 needs the bit mask");
  ECGMM_CHECK(mode != 1 || y, ECGMM_ERR_ARG, "bn_bwd_apply: mode 1 needs y");
  ECGMM_CHECK(mode != 2 || (argmax && scale && shift && !se), ECGMM_ERR_ARG, "bn_bwd_apply: bad mode-2 arguments");
  ECGMM_CHECK(!se || q, ECGMM_ERR_ARG, "bn_bwd_apply: SE mode needs q");
  ECGMM_CHECK(C % 8 == 0, ECGMM_ERR_SHAPE, "bn_bwd_apply: C=%d not a multiple of 8", C);
  const size_t vps = (size_t)H * W * (C >> 3), total = vps * N;
  if (total == 0) return ECGMM_OK;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  cudaStream_t st = (cudaStream_t)stream;
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  const bf16* dyb = reinterpret_cast<const bf16*>(dy);
  const bf16* yb = reinterpret_cast<const bf16*>(y);
  bf16* dxb = reinterpret_cast<bf16*>(dx);
  bf16* dzb = reinterpret_cast<bf16*>(dz_out);
#define ECGMM_BWD(MODE_, SE_)                                                                                  \
  bn_bwd_apply_kernel<MODE_, SE_><<<stream_grid(total, bn_bwd_apply_kernel<MODE_, SE_>, 2), 256, 0, st>>>(     \
      xb, dyb, yb, argmax, coefA, coefB, coefD, scale, shift, se, q, dxb, dzb, C >> 3, vps, total, H, W, Ho, Wo)
  if (mode == 2) {
    ECGMM_CHECK(!dz_out, ECGMM_ERR_ARG, "bn_bwd_apply: mode 2 does not produce dz_out");
    ECGMM_CHECK(256 % (C >> 3) == 0, ECGMM_ERR_SHAPE, "bn_bwd_apply: mode 2 needs C = 8 * a power of two <= 2048 (got %d)", C);
    const size_t tb = (size_t)N * Ho * Wo * (C >> 3);
    // ECGMM_STEM_BWD_APPLY = 0: the grid-stride kernel; 1: one CTA per pooled row, coefficients in registers;
    // 2 (default): one CTA per pooled row, coefficients in shared memory
    const char* ev = getenv("ECGMM_STEM_BWD_APPLY");
    const int variant = ev ? atoi(ev) : 2;
    const long long rows = (long long)N * Ho;
    if (variant == 0 || rows > 0x7fffffffLL) {
      stem_bwd_apply_kernel<<<stream_grid(tb, stem_bwd_apply_kernel), 256, 0, st>>>(
          xb, dyb, argmax, coefA, coefB, coefD, scale, shift, dxb, C >> 3, tb, H, W, Ho, Wo);
    } else {
      int cg_shift = 0;
      while ((1 << cg_shift) < (C >> 3)) ++cg_shift;
      if (variant == 1)
        stem_bwd_apply_rows_kernel<1><<<(unsigned)rows, 256, 0, st>>>(xb, dyb, argmax, coefA, coefB, coefD, scale,
                                                                      shift, dxb, C, cg_shift, H, W, Ho, Wo);
      else
        stem_bwd_apply_rows_kernel<2><<<(unsigned)rows, 256, (size_t)5 * C * sizeof(float), st>>>(
            xb, dyb, argmax, coefA, coefB, coefD, scale, shift, dxb, C, cg_shift, H, W, Ho, Wo);
    }
  } else if (!se && (mode == 0 || mode == 3) && (C >> 3) <= 256 && ((C >> 3) & ((C >> 3) - 1)) == 0 &&
             bn_fast_enabled()) {
    const uint4* x4 = reinterpret_cast<const uint4*>(x);
    const uint4* d4 = reinterpret_cast<const uint4*>(dy);
    uint4* dx4 = reinterpret_cast<uint4*>(dx);
    uint4* dz4 = reinterpret_cast<uint4*>(dz_out);
#define ECGMM_BWD_FAST(MASK_, DZ_)                                                                               \
  do {                                                                                                           \
    if (bn_async_enabled()) {                                                                                    \
      constexpr size_t kBytes = (size_t)kAsyncStages * 2 * 256 * sizeof(uint4);                                  \
      static const int cap = async_grid_cap(bn_bwd_apply_async_kernel<MASK_, DZ_>, kBytes);                      \
      const size_t want = (total + 255) / 256;                                                                   \
      bn_bwd_apply_async_kernel<MASK_, DZ_>                                                                      \
          <<<(unsigned)(want < (size_t)cap ? want : (size_t)cap), 256, kBytes, st>>>(                            \
              x4, d4, argmax, coefA, coefB, coefD, dx4, dz4, C >> 3, total);                                     \
    } else {                                                                                                     \
      bn_bwd_apply_fast_kernel<MASK_, DZ_>                                                                       \
          <<<stream_grid(total, bn_bwd_apply_fast_kernel<MASK_, DZ_>, 2), 256, 0, st>>>(                         \
              x4, d4, argmax, coefA, coefB, coefD, dx4, dz4, C >> 3, total);                                     \
    }                                                                                                            \
  } while (0)
    if (mode == 3 && dz_out)
      ECGMM_BWD_FAST(true, true);
    else if (mode == 3)
      ECGMM_BWD_FAST(true, false);
    else if (dz_out)
      ECGMM_BWD_FAST(false, true);
    else
      ECGMM_BWD_FAST(false, false);
#undef ECGMM_BWD_FAST
    return check_launch("bn_bwd_apply_fast_kernel");
  } else if (mode == 3 && se)
    ECGMM_BWD(3, true);
  else if (mode == 3)
    ECGMM_BWD(3, false);
  else if (mode == 1 && se)
    ECGMM_BWD(1, true);
  else if (mode == 1)
    ECGMM_BWD(1, false);
  else if (se)
    ECGMM_BWD(0, true);
  else
    ECGMM_BWD(0, false);
#undef ECGMM_BWD
  return check_launch("bn_bwd_apply_kernel");
}

extern "C" int ecgmm_avgpool_fwd(const ecgmm_bf16* x, float* out, int N, int P, int C, void* stream) {
  ECGMM_CHECK(x && out, ECGMM_ERR_ARG, "avgpool_fwd: null pointer");
  int rc = check_c(C, "avgpool_fwd");
  if (rc) return rc;
  if (N == 0) return ECGMM_OK;
  ECGMM_CHECK(P > 0, ECGMM_ERR_SHAPE, "avgpool_fwd: empty plane");
  const int rows = kRedThreads / (C >> 3);
  avgpool_fwd_kernel<<<N, kRedThreads, (size_t)rows * C * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(x), out, P, C);
  return check_launch("avgpool_fwd_kernel");
}

extern "C" int ecgmm_avgpool_bwd(const float* dout, ecgmm_bf16* dx, int N, int P, int C, void* stream) {
  ECGMM_CHECK(dout && dx, ECGMM_ERR_ARG, "avgpool_bwd: null pointer");
  ECGMM_CHECK(C % 8 == 0, ECGMM_ERR_SHAPE, "avgpool_bwd: C=%d not a multiple of 8", C);
  const size_t vps = (size_t)P * (C >> 3), total = vps * N;
  if (total == 0) return ECGMM_OK;
  avgpool_bwd_kernel<<<stream_grid(total, avgpool_bwd_kernel), 256, 0, (cudaStream_t)stream>>>(dout, reinterpret_cast<bf16*>(dx), C >> 3,
                                                                         1.f / (float)P, vps, total);
  return check_launch("avgpool_bwd_kernel");
}
