// Explainer maths over the fusion head and the serving-side helpers.
//
// Expected gradients (the estimator behind shap.GradientExplainer, which shap_fusion_modal_balance.py:135,159 drives
// over FusionClassifierWrapper; `shap` itself is unpinned and absent, so the spec is restated in
// ecgmm/explain.py and oracle/model.py):
//     phi[s][d][c] = 1/K * sum_k (e[s][d] - bg[j_sk][d]) * d logit_c / d x_d ( bg[j_sk] + a_sk (e[s] - bg[j_sk]) )
// For fusion_classifier = Linear(D,HID) -> ReLU -> Dropout(eval) -> Linear(HID,C) the gradient is
//     W1^T ( [W1 x + b1 > 0] * w2[c] ),
// so the whole estimate is two fp32 SGEMMs (ecgmm_sgemm) around three bandwidth-bound kernels:
//   eg_points_kernel   the S*K interpolation points                               [S*K][D]
//   eg_gate_kernel     gate[c][r][h] = hidden[r][h] > 0 ? w2[c][h] : 0            [C][S*K][HID]
//   eg_reduce_kernel   phi = mean_k (e - bg_j) * (gate_c W1)                      [S][D][C]
// plus the per-modality |phi| shares of shap_fusion_modal_balance.py:177-200 (modality_share_kernel).
//
// Serving (image-only endpoint + Grad-CAM heat map): softmax / argmax of the logits rows, a row gather (the
// classifier row of each sample's class = d logit / d feature) and
//   gradcam_kernel     cam[n][p] = relu( scale * sum_c g[n][c] * A[n][p][c] )     A = layer4 output, NHWC bf16
// where g = d logit / d pooled comes out of the existing LayerNorm / Linear backward kernels and scale = 1 / (h w)
// is the average pool's uniform gradient.
#include "common.h"
#include "vec.cuh"

namespace ecgmm {

__global__ void __launch_bounds__(256) eg_points_kernel(const float4* __restrict__ e, const float4* __restrict__ bg,
                                                        const int* __restrict__ idx, const float* __restrict__ alpha,
                                                        float4* __restrict__ out, size_t rows, int K, int D4, int NB) {
  const size_t total = rows * (size_t)D4, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / D4;
    const int d = (int)(i - r * D4);
    const size_t s = r / K;
    int j = idx[r];
    j = j < 0 ? 0 : (j >= NB ? NB - 1 : j);  // memory safety only: the host wrapper validates the plan
    const float a = alpha[r];
    const float4 x = e[s * D4 + d], b = bg[(size_t)j * D4 + d];
    out[i] = make_float4(b.x + a * (x.x - b.x), b.y + a * (x.y - b.y), b.z + a * (x.z - b.z), b.w + a * (x.w - b.w));
  }
}

__global__ void __launch_bounds__(256) eg_gate_kernel(const float4* __restrict__ hidden, const float4* __restrict__ w2,
                                                      float4* __restrict__ gate, size_t rows, int H4, int C) {
  const size_t per_class = rows * (size_t)H4, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per_class; i += stride) {
    const int h = (int)(i % H4);
    const float4 v = hidden[i];
    for (int c = 0; c < C; ++c) {
      const float4 w = w2[(size_t)c * H4 + h];
      gate[(size_t)c * per_class + i] =
          make_float4(v.x > 0.f ? w.x : 0.f, v.y > 0.f ? w.y : 0.f, v.z > 0.f ? w.z : 0.f, v.w > 0.f ? w.w : 0.f);
    }
  }
}

// block (64 embedding columns, 4 draw lanes); grid (ceil(D/64), S).  Thread (tx, ty) sums the draws k = ty, ty+4, ...
// of column d for every class, the 4 partial sums meet in shared memory.
template <int MAXC>
__global__ void __launch_bounds__(256) eg_reduce_kernel(const float* __restrict__ e, const float* __restrict__ bg,
                                                        const int* __restrict__ idx, const float* __restrict__ grad,
                                                        float* __restrict__ phi, size_t rows, int K, int D, int C,
                                                        int NB) {
  __shared__ float part[4][MAXC][64];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int d = blockIdx.x * 64 + tx;
  const size_t s = blockIdx.y;
  float acc[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) acc[c] = 0.f;
  if (d < D) {
    const float x = e[s * D + d];
    for (int k = ty; k < K; k += 4) {
      const size_t r = s * K + k;
      int j = idx[r];
      j = j < 0 ? 0 : (j >= NB ? NB - 1 : j);
      const float diff = x - bg[(size_t)j * D + d];
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < C) acc[c] = fmaf(diff, grad[((size_t)c * rows + r) * D + d], acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) part[ty][c][tx] = acc[c];
  __syncthreads();
  if (ty == 0 && d < D) {
    const float inv = 1.f / (float)K;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) phi[(s * D + d) * C + c] = (part[0][c][tx] + part[1][c][tx] + part[2][c][tx] + part[3][c][tx]) * inv;
  }
}

// one warp per (sample, class): the three slice means of |phi| and their shares in percent
__global__ void __launch_bounds__(256) modality_share_kernel(const float* __restrict__ phi, float* __restrict__ share,
                                                             size_t pairs, int C, int D0, int D1, int D2, int use_sum) {
  const int lane = threadIdx.x & 31;
  const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  if (w >= pairs) return;
  const size_t s = w / C;
  const int c = (int)(w - s * C);
  const int D = D0 + D1 + D2;
  const float* row = phi + s * (size_t)D * C + c;
  float m[3] = {0.f, 0.f, 0.f};
  for (int d = lane; d < D; d += 32) {
    const float v = fabsf(row[(size_t)d * C]);
    if (d < D0)
      m[0] += v;
    else if (d < D0 + D1)
      m[1] += v;
    else
      m[2] += v;
  }
  m[0] = warp_sum(m[0]) / (use_sum ? 1.f : (float)D0);
  m[1] = warp_sum(m[1]) / (use_sum ? 1.f : (float)D1);
  m[2] = warp_sum(m[2]) / (use_sum ? 1.f : (float)D2);
  const float total = m[0] + m[1] + m[2];
  if (lane < 3) {
    const float mine = lane == 0 ? m[0] : (lane == 1 ? m[1] : m[2]);
    share[w * 3 + lane] = total > 0.f ? mine / total * 100.f : 0.f;
  }
}

__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ logits, float* __restrict__ probs,
                                                           int* __restrict__ argmax, size_t rows, int C) {
  const size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* x = logits + r * C;
  float mx = x[0];
  int am = 0;
  for (int c = 1; c < C; ++c)
    if (x[c] > mx) {  // first maximum wins, like torch.argmax on ties
      mx = x[c];
      am = c;
    }
  float den = 0.f;
  for (int c = 0; c < C; ++c) den += expf(x[c] - mx);
  if (probs)
    for (int c = 0; c < C; ++c) probs[r * C + c] = expf(x[c] - mx) / den;
  if (argmax) argmax[r] = am;
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table, const int* __restrict__ idx,
                                                          float* __restrict__ out, size_t rows, int D, int NT) {
  const size_t total = rows * (size_t)D, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / D;
    int j = idx[r];
    j = j < 0 ? 0 : (j >= NT ? NT - 1 : j);
    out[i] = table[(size_t)j * D + (i - r * D)];
  }
}

// one warp per pixel, lanes over 8-channel (16-byte) groups; grid (pixel slabs, N)
__global__ void __launch_bounds__(256) gradcam_kernel(const __nv_bfloat16* __restrict__ act, const float* __restrict__ g,
                                                      float* __restrict__ cam, int P, int C, float scale) {
  const int lane = threadIdx.x & 31;
  const size_t n = blockIdx.y;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const float* gn = g + n * C;
  for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < P; p += warps_per_grid) {
    const __nv_bfloat16* row = act + (n * P + p) * (size_t)C;
    float acc = 0.f;
    for (int q = lane * 8; q < C; q += 256) {
      float a[8], w[8];
      unpack8(*reinterpret_cast<const uint4*>(row + q), a);
      load8f(gn + q, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(a[j], w[j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) cam[n * P + p] = fmaxf(acc * scale, 0.f);
  }
}

static unsigned ew_blocks(size_t items) {
  size_t b = (items + 255) / 256;
  const size_t cap = (size_t)num_sms() * 8;
  if (b > cap) b = cap;
  return (unsigned)(b ? b : 1);
}

}  // namespace ecgmm

using namespace ecgmm;

extern "C" int ecgmm_eg_points(const float* e, const float* bg, const int* idx, const float* alpha, float* points,
                               long long S, int K, int D, int NB, void* stream) {
  ECGMM_CHECK(e && bg && idx && alpha && points, ECGMM_ERR_ARG, "eg_points: null pointer");
  ECGMM_CHECK(D > 0 && D % 4 == 0, ECGMM_ERR_SHAPE, "eg_points: D=%d must be a multiple of 4", D);
  ECGMM_CHECK(S >= 0 && K >= 0 && NB >= 1, ECGMM_ERR_SHAPE, "eg_points: bad extents S=%lld K=%d NB=%d", S, K, NB);
  if (S == 0 || K == 0) return ECGMM_OK;
  const size_t rows = (size_t)S * K;
  eg_points_kernel<<<ew_blocks(rows * (D / 4)), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(e), reinterpret_cast<const float4*>(bg), idx, alpha,
      reinterpret_cast<float4*>(points), rows, K, D / 4, NB);
  return check_launch("eg_points_kernel");
}

extern "C" int ecgmm_eg_gate(const float* hidden, const float* w2, float* gate, long long rows, int HID, int C,
                             void* stream) {
  ECGMM_CHECK(hidden && w2 && gate, ECGMM_ERR_ARG, "eg_gate: null pointer");
  ECGMM_CHECK(HID > 0 && HID % 4 == 0, ECGMM_ERR_SHAPE, "eg_gate: hidden width %d must be a multiple of 4", HID);
  ECGMM_CHECK(C >= 1 && C <= 8, ECGMM_ERR_SHAPE, "eg_gate: %d classes (1..8 supported)", C);
  if (rows <= 0) return ECGMM_OK;
  eg_gate_kernel<<<ew_blocks((size_t)rows * (HID / 4)), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(hidden), reinterpret_cast<const float4*>(w2), reinterpret_cast<float4*>(gate),
      (size_t)rows, HID / 4, C);
  return check_launch("eg_gate_kernel");
}

extern "C" int ecgmm_eg_reduce(const float* e, const float* bg, const int* idx, const float* grad, float* phi,
                               long long S, int K, int D, int C, int NB, void* stream) {
  ECGMM_CHECK(e && bg && idx && grad && phi, ECGMM_ERR_ARG, "eg_reduce: null pointer");
  ECGMM_CHECK(D > 0 && K >= 1 && NB >= 1, ECGMM_ERR_SHAPE, "eg_reduce: bad extents K=%d D=%d NB=%d", K, D, NB);
  ECGMM_CHECK(C >= 1 && C <= 8, ECGMM_ERR_SHAPE, "eg_reduce: %d classes (1..8 supported)", C);
  ECGMM_CHECK(S >= 0 && S <= 65535, ECGMM_ERR_SHAPE, "eg_reduce: at most 65535 samples per call (got %lld)", S);
  if (S == 0) return ECGMM_OK;
  const dim3 grid(ceil_div(D, 64), (unsigned)S), block(64, 4);
  const size_t rows = (size_t)S * K;
  if (C <= 2)
    eg_reduce_kernel<2><<<grid, block, 0, (cudaStream_t)stream>>>(e, bg, idx, grad, phi, rows, K, D, C, NB);
  else
    eg_reduce_kernel<8><<<grid, block, 0, (cudaStream_t)stream>>>(e, bg, idx, grad, phi, rows, K, D, C, NB);
  return check_launch("eg_reduce_kernel");
}

extern "C" int ecgmm_modality_share(const float* phi, float* share, long long S, int C, int D0, int D1, int D2,
                                    int use_sum, void* stream) {
  ECGMM_CHECK(phi && share, ECGMM_ERR_ARG, "modality_share: null pointer");
  ECGMM_CHECK(C >= 1 && D0 > 0 && D1 > 0 && D2 > 0, ECGMM_ERR_SHAPE, "modality_share: bad extents C=%d dims=%d,%d,%d",
              C, D0, D1, D2);
  if (S <= 0) return ECGMM_OK;
  const size_t pairs = (size_t)S * C;
  modality_share_kernel<<<(unsigned)((pairs + 7) / 8), 256, 0, (cudaStream_t)stream>>>(phi, share, pairs, C, D0, D1,
                                                                                        D2, use_sum);
  return check_launch("modality_share_kernel");
}

extern "C" int ecgmm_softmax_rows(const float* logits, float* probs, int* argmax, long long rows, int C, void* stream) {
  ECGMM_CHECK(logits && (probs || argmax), ECGMM_ERR_ARG, "softmax_rows: null pointer");
  ECGMM_CHECK(C >= 1, ECGMM_ERR_SHAPE, "softmax_rows: C=%d", C);
  if (rows <= 0) return ECGMM_OK;
  softmax_rows_kernel<<<(unsigned)(((size_t)rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(logits, probs, argmax,
                                                                                               (size_t)rows, C);
  return check_launch("softmax_rows_kernel");
}

extern "C" int ecgmm_gather_rows(const float* table, const int* idx, float* out, long long rows, int D, int NT,
                                 void* stream) {
  ECGMM_CHECK(table && idx && out, ECGMM_ERR_ARG, "gather_rows: null pointer");
  ECGMM_CHECK(D > 0 && NT >= 1, ECGMM_ERR_SHAPE, "gather_rows: bad extents D=%d NT=%d", D, NT);
  if (rows <= 0) return ECGMM_OK;
  gather_rows_kernel<<<ew_blocks((size_t)rows * D), 256, 0, (cudaStream_t)stream>>>(table, idx, out, (size_t)rows, D,
                                                                                    NT);
  return check_launch("gather_rows_kernel");
}

extern "C" int ecgmm_gradcam(const ecgmm_bf16* act, const float* g, float* cam, int N, int P, int C, float scale,
                             void* stream) {
  ECGMM_CHECK(act && g && cam, ECGMM_ERR_ARG, "gradcam: null pointer");
  ECGMM_CHECK(C > 0 && C % 8 == 0, ECGMM_ERR_SHAPE, "gradcam: C=%d must be a multiple of 8", C);
  ECGMM_CHECK(N >= 0 && N <= 65535 && P >= 0, ECGMM_ERR_SHAPE, "gradcam: bad extents N=%d P=%d", N, P);
  if (N == 0 || P == 0) return ECGMM_OK;
  int slabs = ceil_div(P, 8);  // 8 warps (pixels) per CTA
  const int cap = ceil_div(num_sms() * 8, N);
  if (slabs > cap) slabs = cap < 1 ? 1 : cap;
  gradcam_kernel<<<dim3(slabs, N), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(act), g, cam,
                                                                   P, C, scale);
  return check_launch("gradcam_kernel");
}
