// Small fp32 kernels of the encoder tails and the fusion head: tiled SGEMM (Linear forward /
// data-grad / weight-grad), column sums (bias grads), LayerNorm, the AttentionFusion gate
// (softmax over 3 scalars, scale, concat), the variance regulariser, softmax-CE and focal loss
// with their gradients, dropout, element masks, the squeeze-excite MLP and per-row z-score.
// Everything here is latency-bound at training batch sizes (B <= 512 rows of <= 768 floats).
#include "common.h"
#include "vec.cuh"

namespace ecgmm {

// ------------------------------------------------------------------------------------------
// C[M][N] (+)= op(A) * op(B) (+ bias[N]) (relu)
//   ta == 0: A is [M][K] row-major, ta == 1: A is [K][M];  tb == 0: B is [K][N], tb == 1: B is [N][K].
// T x T tile (T = 64: 4x4 outputs per thread, K step 16;  T = 32: 2x2 outputs, K step 32), 256 threads.
// These GEMMs are latency-bound (a handful of CTAs, K <= 768): the next K slab is fetched into registers
// while the current one is multiplied, and the host picks T = 32 whenever 64x64 tiles would leave most
// SMs without a CTA (batch 64: Linear(768,128) is 2 CTAs of 64x64 but 8 of 32x32).
// ------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                     float* __restrict__ C, const float* __restrict__ bias, int M,
                                                     int N, int K, int ta, int tb, int accumulate, int relu) {
  constexpr int KS = 1024 / T;  // K extent of one slab: 1024 elements per operand, 4 per thread
  constexpr int R = T / 16;     // outputs per thread per dimension
  __shared__ float As[KS][T + 1];
  __shared__ float Bs[KS][T + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * T, n0 = blockIdx.x * T;
  float acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;
  // element e of a slab -> (k, row): the fastest-varying index follows the operand's contiguous dimension
  auto coord = [](int e, int transposed_k_major, int& kk, int& rr) {
    if (transposed_k_major) {
      kk = e / T;
      rr = e % T;
    } else {
      rr = e / KS;
      kk = e % KS;
    }
  };
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = threadIdx.x + u * 256;
      int kk, rr;
      coord(e, ta, kk, rr);
      const int m = m0 + rr, k = k0 + kk;
      ra[u] = (m < M && k < K) ? (ta ? A[(size_t)k * M + m] : A[(size_t)m * K + k]) : 0.f;
      coord(e, !tb, kk, rr);
      const int n = n0 + rr, k2 = k0 + kk;
      rb[u] = (n < N && k2 < K) ? (tb ? B[(size_t)n * K + k2] : B[(size_t)k2 * N + n]) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += KS) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = threadIdx.x + u * 256;
      int kk, rr;
      coord(e, ta, kk, rr);
      As[kk][rr] = ra[u];
      coord(e, !tb, kk, rr);
      Bs[kk][rr] = rb[u];
    }
    __syncthreads();
    if (k0 + KS < K) fetch(k0 + KS);  // in flight during the multiply below
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      float a[R], b[R];
#pragma unroll
      for (int i = 0; i < R; ++i) a[i] = As[kk][ty * R + i];
#pragma unroll
      for (int j = 0; j < R; ++j) b[j] = Bs[kk][tx * R + j];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int m = m0 + ty * R + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int n = n0 + tx * R + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (accumulate) v += C[(size_t)m * N + n];
      if (relu) v = fmaxf(v, 0.f);
      C[(size_t)m * N + n] = v;
    }
  }
}

// out[n] (+)= sum_m X[m][n]
__global__ void colsum_kernel(const float* __restrict__ X, float* __restrict__ out, int M, int N, int accumulate) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (n < N)
    for (int m = threadIdx.y; m < M; m += 8) s += X[(size_t)m * N + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[n] = accumulate ? out[n] + t : t;
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dimension (one CTA per row)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ y,
                                                             float* __restrict__ mean_out,
                                                             float* __restrict__ rstd_out, int D, float eps) {
  __shared__ float red[33];
  const size_t row = blockIdx.x;
  const float* xr = x + row * D;
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) s += xr[d];
  const float mean = block_sum(s, red) / D;
  float q = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float t = xr[d] - mean;
    q = fmaf(t, t, q);
  }
  const float rstd = rsqrtf(block_sum(q, red) / D + eps);
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    y[row * D + d] = (xr[d] - mean) * rstd * (gamma ? gamma[d] : 1.f) + (beta ? beta[d] : 0.f);
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
__global__ void __launch_bounds__(256) layernorm_bwd_dx_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ dy,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ rstd,
                                                                float* __restrict__ dx, int D, int accumulate) {
  __shared__ float red[33];
  const size_t row = blockIdx.x;
  const float mu = mean[row], rs = rstd[row];
  float a = 0.f, b = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float g = dy[row * D + d] * (gamma ? gamma[d] : 1.f);
    a += g;
    b = fmaf(g, (x[row * D + d] - mu) * rs, b);
  }
  const float m1 = block_sum(a, red) / D;
  const float m2 = block_sum(b, red) / D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float g = dy[row * D + d] * (gamma ? gamma[d] : 1.f);
    const float xh = (x[row * D + d] - mu) * rs;
    const float v = rs * (g - m1 - xh * m2);
    dx[row * D + d] = accumulate ? dx[row * D + d] + v : v;
  }
}

// dgamma[d] = sum_rows dy * xhat, dbeta[d] = sum_rows dy
__global__ void layernorm_bwd_params_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                            float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int D) {
  __shared__ float r1[8][33], r2[8][33];
  const int d = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, b = 0.f;
  if (d < D)
    for (int r = threadIdx.y; r < rows; r += 8) {
      const float g = dy[(size_t)r * D + d];
      a = fmaf(g, (x[(size_t)r * D + d] - mean[r]) * rstd[r], a);
      b += g;
    }
  r1[threadIdx.y][threadIdx.x] = a;
  r2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && d < D) {
    float ta = 0.f, tb = 0.f;
    for (int i = 0; i < 8; ++i) {
      ta += r1[i][threadIdx.x];
      tb += r2[i][threadIdx.x];
    }
    if (dgamma) dgamma[d] = ta;
    if (dbeta) dbeta[d] = tb;
  }
}

// ------------------------------------------------------------------------------------------
// AttentionFusion gate: w = softmax(weights[3]); fused = cat(w0*f0, w1*f1, w2*f2)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void softmax3(const float* w, float (&o)[3]) {
  const float m = fmaxf(w[0], fmaxf(w[1], w[2]));
  const float e0 = expf(w[0] - m), e1 = expf(w[1] - m), e2 = expf(w[2] - m);
  const float inv = 1.f / (e0 + e1 + e2);
  o[0] = e0 * inv;
  o[1] = e1 * inv;
  o[2] = e2 * inv;
}

__global__ void fusion_gate_fwd_kernel(const float* __restrict__ f0, const float* __restrict__ f1,
                                       const float* __restrict__ f2, const float* __restrict__ weights,
                                       float* __restrict__ fused, float* __restrict__ soft_w, int B, int D0, int D1,
                                       int D2) {
  float w[3];
  softmax3(weights, w);
  const int D = D0 + D1 + D2;
  const size_t total = (size_t)B * D;
  const size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i0 == 0 && soft_w) {
    soft_w[0] = w[0];
    soft_w[1] = w[1];
    soft_w[2] = w[2];
  }
  for (size_t i = i0; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / D;
    const int d = (int)(i % D);
    float v;
    if (d < D0)
      v = w[0] * f0[b * D0 + d];
    else if (d < D0 + D1)
      v = w[1] * f1[b * D1 + (d - D0)];
    else
      v = w[2] * f2[b * D2 + (d - D0 - D1)];
    fused[i] = v;
  }
}

// df_i (+)= w_i * dfused_i   (elementwise part of the gate backward)
__global__ void fusion_gate_bwd_feat_kernel(const float* __restrict__ dfused, const float* __restrict__ weights,
                                            float* __restrict__ df0, float* __restrict__ df1,
                                            float* __restrict__ df2, int B, int D0, int D1, int D2, int accumulate) {
  float w[3];
  softmax3(weights, w);
  const int D = D0 + D1 + D2;
  const size_t total = (size_t)B * D;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / D;
    const int d = (int)(i % D);
    const float g = dfused[i];
    float* dst;
    float v;
    if (d < D0) {
      dst = df0 + b * D0 + d;
      v = w[0] * g;
    } else if (d < D0 + D1) {
      dst = df1 + b * D1 + (d - D0);
      v = w[1] * g;
    } else {
      dst = df2 + b * D2 + (d - D0 - D1);
      v = w[2] * g;
    }
    *dst = accumulate ? *dst + v : v;
  }
}

// dweights[j] = w_j * (s_j - sum_i w_i s_i),  s_i = sum dfused_i * f_i   (single CTA)
__global__ void __launch_bounds__(1024) fusion_gate_bwd_w_kernel(const float* __restrict__ dfused,
                                                                  const float* __restrict__ f0,
                                                                  const float* __restrict__ f1,
                                                                  const float* __restrict__ f2,
                                                                  const float* __restrict__ weights,
                                                                  float* __restrict__ dweights, int B, int D0,
                                                                  int D1, int D2) {
  __shared__ float red[33];
  const int D = D0 + D1 + D2;
  const size_t total = (size_t)B * D;
  float s[3] = {0.f, 0.f, 0.f};
  for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
    const size_t b = i / D;
    const int d = (int)(i % D);
    const float g = dfused[i];
    if (d < D0)
      s[0] = fmaf(g, f0[b * D0 + d], s[0]);
    else if (d < D0 + D1)
      s[1] = fmaf(g, f1[b * D1 + (d - D0)], s[1]);
    else
      s[2] = fmaf(g, f2[b * D2 + (d - D0 - D1)], s[2]);
  }
  float t[3];
  for (int k = 0; k < 3; ++k) t[k] = block_sum(s[k], red);
  if (threadIdx.x == 0) {
    float w[3];
    softmax3(weights, w);
    const float dot = w[0] * t[0] + w[1] * t[1] + w[2] * t[2];
    for (int k = 0; k < 3; ++k) dweights[k] = w[k] * (t[k] - dot);
  }
}

// ------------------------------------------------------------------------------------------
// variance regulariser: v_i = mean_b var_unbiased_d(f_i);  loss = |v0-v1| + |v0-v2| + |v1-v2|
// One CTA.  Saves row means [3][B] and the coefficients c_i = dloss/dv_i for the backward.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) var_loss_fwd_kernel(const float* __restrict__ f0,
                                                             const float* __restrict__ f1,
                                                             const float* __restrict__ f2, float* __restrict__ loss,
                                                             float* __restrict__ row_mean, float* __restrict__ coef,
                                                             int B, int D0, int D1, int D2) {
  __shared__ float red[33];
  const float* f[3] = {f0, f1, f2};
  const int Dm[3] = {D0, D1, D2};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float v[3];
  for (int k = 0; k < 3; ++k) {
    const int D = Dm[k];
    float acc = 0.f;
    for (int b = warp; b < B; b += nw) {  // one warp per row
      const float* r = f[k] + (size_t)b * D;
      float s = 0.f;
      for (int d = lane; d < D; d += 32) s += r[d];
      const float mu = warp_sum(s) / D;
      float q = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float t = r[d] - mu;
        q = fmaf(t, t, q);
      }
      q = warp_sum(q);
      if (lane == 0) {
        acc += q / (D - 1);
        row_mean[(size_t)k * B + b] = mu;
      }
    }
    v[k] = block_sum(acc, red) / B;
  }
  if (threadIdx.x == 0) {
    *loss = fabsf(v[0] - v[1]) + fabsf(v[0] - v[2]) + fabsf(v[1] - v[2]);
    auto sgn = [](float a) { return a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f); };
    coef[0] = sgn(v[0] - v[1]) + sgn(v[0] - v[2]);
    coef[1] = -sgn(v[0] - v[1]) + sgn(v[1] - v[2]);
    coef[2] = -sgn(v[0] - v[2]) - sgn(v[1] - v[2]);
  }
}

// df[b][d] (+)= g * c * 2 (f - rowmean[b]) / ((D-1) B)
__global__ void var_loss_bwd_kernel(const float* __restrict__ f, const float* __restrict__ row_mean,
                                    const float* __restrict__ coef, const float* __restrict__ gout,
                                    float* __restrict__ df, int B, int D, int accumulate) {
  const float k = gout[0] * coef[0] * 2.f / ((float)(D - 1) * (float)B);
  const size_t total = (size_t)B * D;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float v = k * (f[i] - row_mean[i / D]);
    df[i] = accumulate ? df[i] + v : v;
  }
}

// ------------------------------------------------------------------------------------------
// losses (mean reduction).  One CTA; dlogits already divided by B (times gscale).
//   focal == 0: softmax cross entropy;  focal == 1: alpha (1-pt)^gamma ce
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ logits,
                                                       const long long* __restrict__ labels,
                                                       float* __restrict__ loss, float* __restrict__ dlogits, int B,
                                                       int C, int focal, float alpha, float gamma, float gscale,
                                                       long long ignore_index, int* __restrict__ bad_label) {
  // Rows labelled ignore_index do not count (torch: mean over the other rows, zero gradient); any other label outside
  // [0, C) is an error (torch: device assert): flagged in *bad_label, zero gradient, and the loss becomes NaN so that
  // the failure is loud even when nobody reads the flag.
  __shared__ float red[33];
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  float cnt = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const long long y = labels[b];
    if (y == ignore_index) continue;
    if (y < 0 || y >= C) s_bad = 1;  // benign race: every writer stores 1 (after the barrier inside block_sum)
    else cnt += 1.f;
  }
  // CrossEntropyLoss(mean) averages over the rows that are not ignored; the focal loss is a plain mean over ALL rows
  // of per-row values (signal_model.py:99-106: reduction='none', then .mean()) in which an ignored row is 0
  const float n_valid = focal ? (float)B : block_sum(cnt, red);
  __syncthreads();
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* z = logits + (size_t)b * C;
    const long long y = labels[b];
    if (y < 0 || y >= C) {  // ignored or invalid: no contribution, zero gradient row
      if (dlogits)
        for (int c = 0; c < C; ++c) dlogits[(size_t)b * C + c] = 0.f;
      continue;
    }
    float m = z[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(z[c] - m);
    const float lse = m + logf(se);
    const float ce = lse - z[y];
    float li = ce, dce = 1.f;  // dloss_i / dce
    if (focal) {
      const float pt = expf(-ce);
      const float om = 1.f - pt;
      const float pw = powf(om, gamma);
      li = alpha * pw * ce;
      const float pw1 = (gamma == 0.f) ? 0.f : gamma * powf(om, gamma - 1.f);
      dce = alpha * (pw + pw1 * pt * ce);
    }
    acc += li;
    if (dlogits) {
      const float k = dce * gscale / n_valid;
      for (int c = 0; c < C; ++c) {
        const float p = expf(z[c] - lse);
        dlogits[(size_t)b * C + c] = k * (p - (c == y ? 1.f : 0.f));
      }
    }
  }
  const float t = block_sum(acc, red);
  if (threadIdx.x == 0) {
    *loss = s_bad ? __int_as_float(0x7fc00000) : t / n_valid;  // n_valid == 0 -> NaN, as torch
    if (bad_label && s_bad) *bad_label = 1;
  }
}

// ------------------------------------------------------------------------------------------
// dropout and element masks
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_u32(uint64_t k) {  // splitmix64 finaliser
  k += 0x9E3779B97F4A7C15ull;
  k = (k ^ (k >> 30)) * 0xBF58476D1CE4E5B9ull;
  k = (k ^ (k >> 27)) * 0x94D049BB133111EBull;
  k ^= k >> 31;
  return (uint32_t)(k >> 32);
}

// mask_in != NULL: y = x * mask_in (mask already holds 0 or 1/(1-p)).
// otherwise keep element i iff u(seed, i) >= p; mask_out receives 0 or 1/(1-p).
__global__ void dropout_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mask_in,
                                   float* __restrict__ y, float* __restrict__ mask_out, size_t n, float p,
                                   uint64_t seed, const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev) seed += *seed_dev;  // CUDA-graph replay: the per-step offset lives in device memory
  const float keep_scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float m;
    if (mask_in) {
      m = mask_in[i];
    } else {
      const float u = (hash_u32(seed * 0x100000001B3ull + i) >> 8) * (1.f / 16777216.f);
      m = (u >= p) ? keep_scale : 0.f;
    }
    if (mask_out) mask_out[i] = m;
    y[i] = x[i] * m;
  }
}

// dx = dy * (y > 0 if y) * (mask if mask)
__global__ void mask_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                const float* __restrict__ mask, float* __restrict__ dx, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float g = dy[i];
    if (mask) g *= mask[i];
    if (y && y[i] <= 0.f) g = 0.f;
    dx[i] = g;
  }
}

// ------------------------------------------------------------------------------------------
// squeeze-excite MLP (one CTA per sample, C threads):
//   pooled = scale*nsum/L + shift ; h = relu(W1 pooled + b1) ; s = sigmoid(W2 h + b2)
// ------------------------------------------------------------------------------------------
__global__ void se_fwd_kernel(const float* __restrict__ nsum, const float* __restrict__ scale,
                              const float* __restrict__ shift, const float* __restrict__ w1,
                              const float* __restrict__ b1, const float* __restrict__ w2,
                              const float* __restrict__ b2, float* __restrict__ pooled, float* __restrict__ hid,
                              float* __restrict__ gate, int C, int R, float inv_len) {
  extern __shared__ float sm[];  // pooled[C], h[R]
  float* sp = sm;
  float* shd = sm + C;
  const int n = blockIdx.x, c = threadIdx.x;
  const float pv = fmaf(scale[c], nsum[(size_t)n * C + c] * inv_len, shift[c]);
  sp[c] = pv;
  pooled[(size_t)n * C + c] = pv;
  __syncthreads();
  const int lane = c & 31, warp = c >> 5, nw = C >> 5;
  for (int r = warp; r < R; r += nw) {
    float a = 0.f;
    for (int k = lane; k < C; k += 32) a = fmaf(w1[(size_t)r * C + k], sp[k], a);
    a = warp_sum(a);
    if (lane == 0) {
      const float hv = fmaxf(a + b1[r], 0.f);
      shd[r] = hv;
      hid[(size_t)n * R + r] = hv;
    }
  }
  __syncthreads();
  float a = b2[c];
  for (int r = 0; r < R; ++r) a = fmaf(w2[(size_t)c * R + r], shd[r], a);
  gate[(size_t)n * C + c] = 1.f / (1.f + expf(-a));
}

// Backward through the gate: ds = gamma*sum(P2) + beta*sum(P1) -> dpre2 -> dh -> dpre1 -> dpool;
// q = dpool / L is the per-(n,c) constant added to the gradient of the BatchNorm output.
__global__ void se_bwd_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int split,
                              const float* __restrict__ gamma, const float* __restrict__ beta,
                              const float* __restrict__ w1, const float* __restrict__ w2,
                              const float* __restrict__ hid, const float* __restrict__ gate,
                              float* __restrict__ q, float* __restrict__ dpre2, float* __restrict__ dpre1, int C,
                              int R, float inv_len) {
  extern __shared__ float sm[];  // d2[C], d1[R]
  float* s2 = sm;
  float* s1 = sm + C;
  const int n = blockIdx.x, c = threadIdx.x;
  float a = 0.f, b = 0.f;
  for (int k = 0; k < split; ++k) {
    const size_t i = ((size_t)n * split + k) * C + c;
    a += p1[i];
    b += p2[i];
  }
  const float g = gate[(size_t)n * C + c];
  const float ds = gamma[c] * b + beta[c] * a;
  const float d2 = ds * g * (1.f - g);
  s2[c] = d2;
  dpre2[(size_t)n * C + c] = d2;
  __syncthreads();
  const int lane = c & 31, warp = c >> 5, nw = C >> 5;
  for (int r = warp; r < R; r += nw) {
    float t = 0.f;
    for (int k = lane; k < C; k += 32) t = fmaf(w2[(size_t)k * R + r], s2[k], t);
    t = warp_sum(t);
    if (lane == 0) {
      const float d1 = hid[(size_t)n * R + r] > 0.f ? t : 0.f;
      s1[r] = d1;
      dpre1[(size_t)n * R + r] = d1;
    }
  }
  __syncthreads();
  float dp = 0.f;
  for (int r = 0; r < R; ++r) dp = fmaf(w1[(size_t)r * C + c], s1[r], dp);
  q[(size_t)n * C + c] = dp * inv_len;
}

// ------------------------------------------------------------------------------------------
// z-score per row: (x - mean) / (std_population + 1e-8)      signal_model.py:203-206
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) zscore_kernel(const float* __restrict__ x, float* __restrict__ y, int L,
                                                      float eps) {
  __shared__ float red[33];
  const float* xr = x + (size_t)blockIdx.x * L;
  float* yr = y + (size_t)blockIdx.x * L;
  float s = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s += xr[i];
  const float mean = block_sum(s, red) / L;
  float qv = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float t = xr[i] - mean;
    qv = fmaf(t, t, qv);
  }
  const float inv = 1.f / (sqrtf(block_sum(qv, red) / L) + eps);
  for (int i = threadIdx.x; i < L; i += blockDim.x) yr[i] = (xr[i] - mean) * inv;
}

// ------------------------------------------------------------------------------------------
// BatchNorm1d over a small fp32 feature matrix [B][C] (clinical MLP,
// multimodal_paper_modal_balance.py:258): one thread per feature column.
// ------------------------------------------------------------------------------------------
__global__ void bn_rows_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ num_batches,
                                   float* __restrict__ y, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, int B, int C, float eps, float momentum,
                                   int train, int relu) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && train && num_batches) *num_batches += 1;
  if (c >= C) return;
  float mean, invstd;
  if (train) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += x[(size_t)b * C + c];
    mean = s / B;
    float q = 0.f;
    for (int b = 0; b < B; ++b) {
      const float t = x[(size_t)b * C + c] - mean;
      q = fmaf(t, t, q);
    }
    const float var = q / B;
    invstd = rsqrtf(var + eps);
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (B > 1 ? q / (B - 1) : var);
    }
  } else {
    mean = running_mean[c];
    invstd = rsqrtf(running_var[c] + eps);
  }
  if (mean_out) mean_out[c] = mean;
  if (invstd_out) invstd_out[c] = invstd;
  const float sc = gamma[c] * invstd, sh = beta[c] - mean * sc;
  for (int b = 0; b < B; ++b) {
    float v = fmaf(x[(size_t)b * C + c], sc, sh);
    if (relu) v = fmaxf(v, 0.f);
    y[(size_t)b * C + c] = v;
  }
}

// dy is the gradient of the (optionally ReLU-ed) output y
__global__ void bn_rows_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                   const float* __restrict__ y, const float* __restrict__ gamma,
                                   const float* __restrict__ mean, const float* __restrict__ invstd,
                                   float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                   int B, int C, int relu) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mu = mean[c], is = invstd[c];
  float s1 = 0.f, s2 = 0.f;
  for (int b = 0; b < B; ++b) {
    const size_t i = (size_t)b * C + c;
    float g = dy[i];
    if (relu && y[i] <= 0.f) g = 0.f;
    s1 += g;
    s2 = fmaf(g, (x[i] - mu) * is, s2);
  }
  if (dgamma) dgamma[c] = s2;
  if (dbeta) dbeta[c] = s1;
  if (dx) {
    const float m1 = s1 / B, m2 = s2 / B, k = gamma[c] * is;
    for (int b = 0; b < B; ++b) {
      const size_t i = (size_t)b * C + c;
      float g = dy[i];
      if (relu && y[i] <= 0.f) g = 0.f;
      dx[i] = k * (g - m1 - (x[i] - mu) * is * m2);
    }
  }
}

static int ew_grid(size_t n) {
  size_t b = (n + 255) / 256;
  const size_t cap = (size_t)num_sms() * 8;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace ecgmm

using namespace ecgmm;

extern "C" int ecgmm_sgemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K,
                           int transA, int transB, int accumulate, int relu, void* stream) {
  ECGMM_CHECK(A && B && C, ECGMM_ERR_ARG, "sgemm: null pointer");
  ECGMM_CHECK(M >= 0 && N >= 0 && K >= 0, ECGMM_ERR_SHAPE, "sgemm: negative extent");
  if (M == 0 || N == 0) return ECGMM_OK;
  const bool small = (long long)ceil_div(N, 64) * ceil_div(M, 64) < 2LL * num_sms();
  const int T = small ? 32 : 64;
  dim3 grid(ceil_div(N, T), ceil_div(M, T));
  ECGMM_CHECK(grid.y <= 65535, ECGMM_ERR_SHAPE, "sgemm: M=%d too large", M);
  if (small)
    sgemm_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, C, bias, M, N, K, transA, transB, accumulate, relu);
  else
    sgemm_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, C, bias, M, N, K, transA, transB, accumulate, relu);
  return check_launch("sgemm_kernel");
}

extern "C" int ecgmm_colsum(const float* X, float* out, int M, int N, int accumulate, void* stream) {
  ECGMM_CHECK(X && out, ECGMM_ERR_ARG, "colsum: null pointer");
  if (N == 0) return ECGMM_OK;
  colsum_kernel<<<ceil_div(N, 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(X, out, M, N, accumulate);
  return check_launch("colsum_kernel");
}

extern "C" int ecgmm_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                   float* rstd, int rows, int D, float eps, void* stream) {
  ECGMM_CHECK(x && y, ECGMM_ERR_ARG, "layernorm_fwd: null pointer");
  ECGMM_CHECK(D > 0, ECGMM_ERR_SHAPE, "layernorm_fwd: D=%d", D);
  if (rows == 0) return ECGMM_OK;
  layernorm_fwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, mean, rstd, D, eps);
  return check_launch("layernorm_fwd_kernel");
}

extern "C" int ecgmm_layernorm_bwd(const float* x, const float* dy, const float* gamma, const float* mean,
                                   const float* rstd, float* dx, float* dgamma, float* dbeta, int rows, int D,
                                   int accumulate_dx, void* stream) {
  ECGMM_CHECK(x && dy && mean && rstd, ECGMM_ERR_ARG, "layernorm_bwd: null pointer");
  if (rows == 0) return ECGMM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) layernorm_bwd_dx_kernel<<<rows, 256, 0, st>>>(x, dy, gamma, mean, rstd, dx, D, accumulate_dx);
  if (dgamma || dbeta)
    layernorm_bwd_params_kernel<<<ceil_div(D, 32), dim3(32, 8), 0, st>>>(x, dy, mean, rstd, dgamma, dbeta, rows, D);
  return check_launch("layernorm_bwd", (dx ? 1 : 0) + ((dgamma || dbeta) ? 1 : 0));
}

extern "C" int ecgmm_fusion_gate_fwd(const float* f0, const float* f1, const float* f2, const float* weights,
                                     float* fused, float* soft_w, int B, int D0, int D1, int D2, void* stream) {
  ECGMM_CHECK(f0 && f1 && f2 && weights && fused, ECGMM_ERR_ARG, "fusion_gate_fwd: null pointer");
  const size_t total = (size_t)B * (D0 + D1 + D2);
  fusion_gate_fwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(f0, f1, f2, weights, fused, soft_w, B, D0,
                                                                         D1, D2);
  return check_launch("fusion_gate_fwd_kernel");
}

extern "C" int ecgmm_fusion_gate_bwd(const float* dfused, const float* f0, const float* f1, const float* f2,
                                     const float* weights, float* df0, float* df1, float* df2, float* dweights,
                                     int B, int D0, int D1, int D2, int accumulate, void* stream) {
  ECGMM_CHECK(dfused && f0 && f1 && f2 && weights, ECGMM_ERR_ARG, "fusion_gate_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)B * (D0 + D1 + D2);
  if (total == 0) return ECGMM_OK;
  if (df0 && df1 && df2)
    fusion_gate_bwd_feat_kernel<<<ew_grid(total), 256, 0, st>>>(dfused, weights, df0, df1, df2, B, D0, D1, D2,
                                                                accumulate);
  if (dweights) fusion_gate_bwd_w_kernel<<<1, 1024, 0, st>>>(dfused, f0, f1, f2, weights, dweights, B, D0, D1, D2);
  return check_launch("fusion_gate_bwd", ((df0 && df1 && df2) ? 1 : 0) + (dweights ? 1 : 0));
}

extern "C" int ecgmm_var_loss_fwd(const float* f0, const float* f1, const float* f2, float* loss, float* row_mean,
                                  float* coef, int B, int D0, int D1, int D2, void* stream) {
  ECGMM_CHECK(f0 && f1 && f2 && loss && row_mean && coef, ECGMM_ERR_ARG, "var_loss_fwd: null pointer");
  ECGMM_CHECK(B > 0 && D0 > 1 && D1 > 1 && D2 > 1, ECGMM_ERR_SHAPE, "var_loss_fwd: needs B>0 and D>1");
  var_loss_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(f0, f1, f2, loss, row_mean, coef, B, D0, D1, D2);
  return check_launch("var_loss_fwd_kernel");
}

extern "C" int ecgmm_var_loss_bwd(const float* f, const float* row_mean, const float* coef, const float* gout,
                                  float* df, int B, int D, int accumulate, void* stream) {
  ECGMM_CHECK(f && row_mean && coef && gout && df, ECGMM_ERR_ARG, "var_loss_bwd: null pointer");
  const size_t total = (size_t)B * D;
  if (total == 0) return ECGMM_OK;
  var_loss_bwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(f, row_mean, coef, gout, df, B, D, accumulate);
  return check_launch("var_loss_bwd_kernel");
}

extern "C" int ecgmm_ce_loss(const float* logits, const long long* labels, float* loss, float* dlogits, int B, int C,
                             int focal, float alpha, float gamma, float gscale, long long ignore_index,
                             int* bad_label, void* stream) {
  ECGMM_CHECK(logits && labels && loss, ECGMM_ERR_ARG, "ce_loss: null pointer");
  ECGMM_CHECK(B > 0 && C > 0, ECGMM_ERR_SHAPE, "ce_loss: empty batch");
  ce_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, labels, loss, dlogits, B, C, focal, alpha, gamma,
                                                     gscale, ignore_index, bad_label);
  return check_launch("ce_loss_kernel");
}

extern "C" int ecgmm_dropout_fwd(const float* x, const float* mask_in, float* y, float* mask_out, long long n,
                                 float p, unsigned long long seed, const unsigned long long* seed_dev,
                                 void* stream) {
  ECGMM_CHECK(x && y, ECGMM_ERR_ARG, "dropout_fwd: null pointer");
  ECGMM_CHECK(p >= 0.f && p <= 1.f, ECGMM_ERR_ARG, "dropout_fwd: p=%f", p);
  if (n == 0) return ECGMM_OK;
  dropout_fwd_kernel<<<ew_grid((size_t)n), 256, 0, (cudaStream_t)stream>>>(x, mask_in, y, mask_out, (size_t)n, p,
                                                                         seed, seed_dev);
  return check_launch("dropout_fwd_kernel");
}

extern "C" int ecgmm_mask_bwd(const float* dy, const float* y, const float* mask, float* dx, long long n,
                              void* stream) {
  ECGMM_CHECK(dy && dx, ECGMM_ERR_ARG, "mask_bwd: null pointer");
  if (n == 0) return ECGMM_OK;
  mask_bwd_kernel<<<ew_grid((size_t)n), 256, 0, (cudaStream_t)stream>>>(dy, y, mask, dx, (size_t)n);
  return check_launch("mask_bwd_kernel");
}

extern "C" int ecgmm_se_fwd(const float* nsum, const float* scale, const float* shift, const float* w1,
                            const float* b1, const float* w2, const float* b2, float* pooled, float* hid,
                            float* gate, int N, int C, int R, int L, void* stream) {
  ECGMM_CHECK(nsum && scale && shift && w1 && b1 && w2 && b2 && pooled && hid && gate, ECGMM_ERR_ARG,
              "se_fwd: null pointer");
  ECGMM_CHECK(C % 32 == 0 && C <= 1024 && R > 0 && L > 0, ECGMM_ERR_SHAPE, "se_fwd: C=%d R=%d L=%d", C, R, L);
  if (N == 0) return ECGMM_OK;
  se_fwd_kernel<<<N, C, (C + R) * sizeof(float), (cudaStream_t)stream>>>(nsum, scale, shift, w1, b1, w2, b2, pooled,
                                                                       hid, gate, C, R, 1.f / (float)L);
  return check_launch("se_fwd_kernel");
}

extern "C" int ecgmm_se_bwd(const float* p1, const float* p2, int split, const float* gamma, const float* beta,
                            const float* w1, const float* w2, const float* hid, const float* gate, float* q,
                            float* dpre2, float* dpre1, int N, int C, int R, int L, void* stream) {
  ECGMM_CHECK(p1 && p2 && gamma && beta && w1 && w2 && hid && gate && q && dpre2 && dpre1, ECGMM_ERR_ARG,
              "se_bwd: null pointer");
  ECGMM_CHECK(C % 32 == 0 && C <= 1024 && R > 0 && L > 0, ECGMM_ERR_SHAPE, "se_bwd: C=%d R=%d L=%d", C, R, L);
  if (N == 0) return ECGMM_OK;
  se_bwd_kernel<<<N, C, (C + R) * sizeof(float), (cudaStream_t)stream>>>(p1, p2, split, gamma, beta, w1, w2, hid,
                                                                       gate, q, dpre2, dpre1, C, R, 1.f / (float)L);
  return check_launch("se_bwd_kernel");
}

extern "C" int ecgmm_zscore(const float* x, float* y, long long rows, int L, float eps, void* stream) {
  ECGMM_CHECK(x && y, ECGMM_ERR_ARG, "zscore: null pointer");
  ECGMM_CHECK(L > 0, ECGMM_ERR_SHAPE, "zscore: L=%d", L);
  if (rows == 0) return ECGMM_OK;
  ECGMM_CHECK(rows <= 2147483647ll, ECGMM_ERR_SHAPE, "zscore: too many rows");
  zscore_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, y, L, eps);
  return check_launch("zscore_kernel");
}

extern "C" int ecgmm_bn_rows_fwd(const float* x, const float* gamma, const float* beta, float* running_mean,
                                 float* running_var, long long* num_batches, float* y, float* mean, float* invstd,
                                 int B, int C, float eps, float momentum, int train, int relu, void* stream) {
  ECGMM_CHECK(x && gamma && beta && y, ECGMM_ERR_ARG, "bn_rows_fwd: null pointer");
  ECGMM_CHECK(train || (running_mean && running_var), ECGMM_ERR_ARG, "bn_rows_fwd: eval mode needs running stats");
  ECGMM_CHECK(!train || B > 1, ECGMM_ERR_SHAPE, "bn_rows_fwd: training needs more than 1 row (got %d)", B);
  if (B == 0 || C == 0) return ECGMM_OK;
  bn_rows_fwd_kernel<<<ceil_div(C, 64), 64, 0, (cudaStream_t)stream>>>(x, gamma, beta, running_mean, running_var,
                                                                     num_batches, y, mean, invstd, B, C, eps,
                                                                     momentum, train, relu);
  return check_launch("bn_rows_fwd_kernel");
}

extern "C" int ecgmm_bn_rows_bwd(const float* x, const float* dy, const float* y, const float* gamma,
                                 const float* mean, const float* invstd, float* dx, float* dgamma, float* dbeta,
                                 int B, int C, int relu, void* stream) {
  ECGMM_CHECK(x && dy && gamma && mean && invstd, ECGMM_ERR_ARG, "bn_rows_bwd: null pointer");
  ECGMM_CHECK(!relu || y, ECGMM_ERR_ARG, "bn_rows_bwd: relu needs y");
  if (B == 0 || C == 0) return ECGMM_OK;
  bn_rows_bwd_kernel<<<ceil_div(C, 64), 64, 0, (cudaStream_t)stream>>>(x, dy, y, gamma, mean, invstd, dx, dgamma,
                                                                     dbeta, B, C, relu);
  return check_launch("bn_rows_bwd_kernel");
}
