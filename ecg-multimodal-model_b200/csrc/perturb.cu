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

// A CTA owns one sample and a slab of variants; thread t owns 8 embedding elements (e and bg stay in registers)
// and walks down the slab: per 16-byte store it issues ONE 8-byte mask load (L2-resident: the V x D mask matrix
// is shared by all samples).  A row of D elements is written by D/8 consecutive threads (contiguous bytes).
__global__ void __launch_bounds__(256) perturb_build_kernel(const float* __restrict__ e, const float* __restrict__ bg,
                                                            const uint8_t* __restrict__ masks,
                                                            __nv_bfloat16* __restrict__ out, int V, int DG,
                                                            int v_per_cta) {
  const size_t s = blockIdx.y;
  const int rows_per_pass = blockDim.x / DG;  // variants covered by the CTA at once
  const int dg = threadIdx.x % DG, r = threadIdx.x / DG;
  if (r >= rows_per_pass) return;
  float fe[8], fb[8];
  load8f(e + (s * DG + dg) * 8, fe);
  load8f(bg + dg * 8, fb);
  uint32_t pe[4], pb[4];  // packed bf16 pairs of both candidates: the select is 4 byte-permutes per store
  {
    const uint4 ve = pack8(fe), vb = pack8(fb);
    pe[0] = ve.x; pe[1] = ve.y; pe[2] = ve.z; pe[3] = ve.w;
    pb[0] = vb.x; pb[1] = vb.y; pb[2] = vb.z; pb[3] = vb.w;
  }
  const int v0 = blockIdx.x * v_per_cta, v1 = min(V, v0 + v_per_cta);
  const uint2* mrow = reinterpret_cast<const uint2*>(masks);
  uint4* orow = reinterpret_cast<uint4*>(out) + s * (size_t)V * DG;
#pragma unroll 4
  for (int v = v0 + r; v < v1; v += rows_per_pass) {
    const uint2 m = mrow[(size_t)v * DG + dg];
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t mw = (k < 2 ? m.x : m.y) >> (16 * (k & 1));  // mask bytes of elements 2k, 2k+1
      const uint32_t lo = (mw & 0xffu) ? (pe[k] & 0xffffu) : (pb[k] & 0xffffu);
      const uint32_t hi = (mw & 0xff00u) ? (pe[k] & 0xffff0000u) : (pb[k] & 0xffff0000u);
      o[k] = lo | hi;
    }
    orow[(size_t)v * DG + dg] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// relu(hidden + b1) -> Linear(HID, C) -> softmax.  A warp takes RPW rows per trip (RPW independent 8-byte loads
// per lane in flight); lane l owns hidden units [4l, 4l+4) (+128, ...), whose b1 / w2 slices stay in registers
// when HID == 128.  out = softmax(logits)[cls] (cls >= 0) or the C logits (cls < 0).
template <int MAXC>
__global__ void __launch_bounds__(256) head_tail_kernel(const __nv_bfloat16* __restrict__ hidden,
                                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, float* __restrict__ out,
                                                        size_t rows, int HID, int C, int cls) {
  constexpr int RPW = 4;
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t row0 = warp * RPW; row0 < rows; row0 += nwarps * RPW) {
    float acc[RPW][MAXC];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int c = 0; c < MAXC; ++c) acc[r][c] = 0.f;
    for (int k = lane * 4; k < HID; k += 128) {
      uint2 hv[RPW];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const size_t row = row0 + r < rows ? row0 + r : rows - 1;
        hv[r] = *reinterpret_cast<const uint2*>(hidden + row * HID + k);
      }
      const float4 bb = *reinterpret_cast<const float4*>(b1 + k);
      float4 w[MAXC];
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        w[c] = c < C ? *reinterpret_cast<const float4*>(w2 + (size_t)c * HID + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&hv[r]);
        const float2 h01 = __bfloat1622float2(hb[0]), h23 = __bfloat1622float2(hb[1]);
        const float h0 = fmaxf(h01.x + bb.x, 0.f), h1 = fmaxf(h01.y + bb.y, 0.f);
        const float h2 = fmaxf(h23.x + bb.z, 0.f), h3 = fmaxf(h23.y + bb.w, 0.f);
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          acc[r][c] = fmaf(h0, w[c].x, fmaf(h1, w[c].y, fmaf(h2, w[c].z, fmaf(h3, w[c].w, acc[r][c]))));
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int c = 0; c < MAXC; ++c) acc[r][c] = warp_sum(acc[r][c]) + (c < C ? b2[c] : 0.f);
    if (lane < RPW && row0 + lane < rows) {
      float a[MAXC];
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {  // lane r keeps row r (static indexing: no local-memory array)
        a[c] = acc[0][c];
#pragma unroll
        for (int r = 1; r < RPW; ++r)
          if (lane == r) a[c] = acc[r][c];
      }
      const size_t row = row0 + lane;
      if (cls < 0) {
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < C) out[row * C + c] = a[c];
      } else {
        float mx = -INFINITY, den = 0.f, num = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < C) mx = fmaxf(mx, a[c]);
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < C) {
            const float ex = __expf(a[c] - mx);
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
  ECGMM_CHECK((D >> 3) <= 256, ECGMM_ERR_SHAPE, "perturb_build: D=%d too wide (max 2048)", D);
  ECGMM_CHECK(S <= 65535, ECGMM_ERR_SHAPE, "perturb_build: at most 65535 samples per call (got %lld)", S);
  if (S == 0 || V == 0) return ECGMM_OK;
  const int DG = D >> 3;
  // enough CTAs per sample to fill the GPU ~8 deep, but at least 32 variants per CTA
  int slabs = (int)(((long long)num_sms() * 8 + S - 1) / S);
  if (slabs < 1) slabs = 1;
  int v_per_cta = ceil_div(V, slabs);
  if (v_per_cta < 32) v_per_cta = 32;
  dim3 grid(ceil_div(V, v_per_cta), (unsigned)S);
  perturb_build_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      e, bg, masks, reinterpret_cast<__nv_bfloat16*>(variants), V, DG, v_per_cta);
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
  size_t blocks = ((size_t)rows + 31) / 32;  // 8 warps x 4 rows per trip
  if (blocks > cap) blocks = cap;
  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(hidden);
  if (C <= 2)
    head_tail_kernel<2><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(h, b1, w2, b2, out, (size_t)rows, HID, C, cls);
  else
    head_tail_kernel<8><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(h, b1, w2, b2, out, (size_t)rows, HID, C, cls);
  return check_launch("head_tail_kernel");
}
