// Batched perturbation inference of the fusion head (BASELINE.json configs[3], SURVEY.md section 8d "cfg4"):
// V masked variants of every sample's fused embedding e (D = 768) against a background embedding b,
//     variant[s][v] = z[v] * e[s] + (1 - z[v]) * b,
// pushed through fusion_classifier = Linear(D,128) -> ReLU -> Dropout(eval: identity) -> Linear(128,C)
// (multimodal_paper_modal_balance.py:283-289, driven by shap_fusion_modal_balance.py:135,159) and reduced to
// softmax(logits)[:, class].
//
// Three steps, the middle one on the tensor cores:
//   perturb_build_kernel   masks (1 byte per element, shared by all samples) + e + b -> variants bf16
//                          [S*V][D]  (HBM-bound: 2 B written per element)
//   ecgmm_conv2d_fwd       the [S*V, D] x [D, 128] GEMM as a 1x1 "convolution" over a 1 x (S*V) image:
//                          the same TMA + tcgen05 + TMEM implicit-GEMM kernel as the ResNet convolutions
//   head_tail_kernel       bias + ReLU + Linear(128, C) + softmax, one warp per variant row, fp32
#include "common.h"
#include "vec.cuh"

namespace ecgmm {

// 8 elements (16 B of bf16) per thread; masks are read 8 bytes at a time.
__global__ void __launch_bounds__(256) perturb_build_kernel(const float* __restrict__ e, const float* __restrict__ bg,
                                                            const uint8_t* __restrict__ masks,
                                                            __nv_bfloat16* __restrict__ out, int V, int DG,
                                                            size_t total_vec) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total_vec; i += stride) {
    const int dg = (int)(i % DG);
    const size_t sv = i / DG;
    const int v = (int)(sv % V);
    const size_t s = sv / V;
    const uint2 m = reinterpret_cast<const uint2*>(masks)[(size_t)v * DG + dg];
    float fe[8], fb[8], f[8];
    load8f(e + (s * DG + dg) * 8, fe);
    load8f(bg + dg * 8, fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t byte = ((j < 4 ? m.x : m.y) >> (8 * (j & 3))) & 0xffu;
      f[j] = byte ? fe[j] : fb[j];
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(f);
  }
}

// One warp per row: h = relu(hidden[row] + b1) (HID values, 4 per lane for HID = 128), logits = W2 h + b2,
// out = softmax(logits)[cls] (cls >= 0) or the C logits (cls < 0).
template <int MAXC>
__global__ void __launch_bounds__(256) head_tail_kernel(const __nv_bfloat16* __restrict__ hidden,
                                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, float* __restrict__ out,
                                                        size_t rows, int HID, int C, int cls) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t row = warp; row < rows; row += nwarps) {
    float acc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) acc[c] = 0.f;
    for (int k = lane * 4; k < HID; k += 128) {
      const uint2 hv = *reinterpret_cast<const uint2*>(hidden + row * HID + k);
      const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&hv);
      const float2 h01 = __bfloat1622float2(hb[0]), h23 = __bfloat1622float2(hb[1]);
      const float4 bb = *reinterpret_cast<const float4*>(b1 + k);
      const float h0 = fmaxf(h01.x + bb.x, 0.f), h1 = fmaxf(h01.y + bb.y, 0.f);
      const float h2 = fmaxf(h23.x + bb.z, 0.f), h3 = fmaxf(h23.y + bb.w, 0.f);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          const float4 w = *reinterpret_cast<const float4*>(w2 + (size_t)c * HID + k);
          acc[c] = fmaf(h0, w.x, fmaf(h1, w.y, fmaf(h2, w.z, fmaf(h3, w.w, acc[c]))));
        }
      }
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) acc[c] = warp_sum(acc[c]) + (c < C ? b2[c] : 0.f);
    if (lane == 0) {
      if (cls < 0) {
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < C) out[row * C + c] = acc[c];
      } else {
        float mx = -INFINITY, den = 0.f, num = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < C) mx = fmaxf(mx, acc[c]);
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < C) {
            const float ex = __expf(acc[c] - mx);
            den += ex;
            if (c == cls) num = ex;
          }
        out[row] = num / den;
      }
    }
  }
}

}  // namespace ecgmm

using namespace ecgmm;

extern "C" int ecgmm_perturb_build(const float* e, const float* bg, const uint8_t* masks, ecgmm_bf16* variants,
                                   long long S, int V, int D, void* stream) {
  ECGMM_CHECK(e && bg && masks && variants, ECGMM_ERR_ARG, "perturb_build: null pointer");
  ECGMM_CHECK(D > 0 && D % 8 == 0, ECGMM_ERR_SHAPE, "perturb_build: D=%d must be a multiple of 8", D);
  ECGMM_CHECK(S >= 0 && V >= 0, ECGMM_ERR_SHAPE, "perturb_build: negative extent");
  const size_t total = (size_t)S * V * (D >> 3);
  if (total == 0) return ECGMM_OK;
  const size_t cap = (size_t)num_sms() * 8;
  size_t blocks = (total + 255) / 256;
  if (blocks > cap) blocks = cap;
  perturb_build_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      e, bg, masks, reinterpret_cast<__nv_bfloat16*>(variants), V, D >> 3, total);
  return check_launch("perturb_build_kernel");
}

extern "C" int ecgmm_head_tail(const ecgmm_bf16* hidden, const float* b1, const float* w2, const float* b2,
                               float* out, long long rows, int HID, int C, int cls, void* stream) {
  ECGMM_CHECK(hidden && b1 && w2 && b2 && out, ECGMM_ERR_ARG, "head_tail: null pointer");
  ECGMM_CHECK(HID > 0 && HID % 4 == 0, ECGMM_ERR_SHAPE, "head_tail: hidden width %d must be a multiple of 4", HID);
  ECGMM_CHECK(C >= 1 && C <= 8, ECGMM_ERR_SHAPE, "head_tail: %d classes (1..8 supported)", C);
  ECGMM_CHECK(cls < C, ECGMM_ERR_ARG, "head_tail: class index %d out of range", cls);
  if (rows <= 0) return ECGMM_OK;
  const size_t cap = (size_t)num_sms() * 8;
  size_t blocks = ((size_t)rows + 7) / 8;
  if (blocks > cap) blocks = cap;
  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(hidden);
  if (C <= 2)
    head_tail_kernel<2><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(h, b1, w2, b2, out, (size_t)rows, HID, C, cls);
  else
    head_tail_kernel<8><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(h, b1, w2, b2, out, (size_t)rows, HID, C, cls);
  return check_launch("head_tail_kernel");
}
