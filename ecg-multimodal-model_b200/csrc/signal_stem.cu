// Stem of the 1-D ResNet-SE: Conv1d(Cin in {1..16}, 64, k=7, s=2, p=3) over raw fp32 signals
// [B][Cin][L] (reference: multimodal_paper_modal_balance.py:99-104, train_signal_12_af.py:184).
// K = Cin*7 <= 112 is too small / misaligned for the tensor-core path, and the whole layer is
// ~1 % of the signal branch traffic, so it is a shared-memory CUDA-core kernel.
// Output is channels-last bf16 [B][Lo][64] WITHOUT the bias: in training the following BatchNorm
// cancels it (it only shifts running_mean, handled by bn_finalize), in eval it is folded into
// the BatchNorm shift.
#include "common.h"
#include "vec.cuh"

namespace ecgmm {

constexpr int kStemCo = 64;
constexpr int kStemTile = 128;                 // output positions per CTA tile
constexpr int kStemSpan = 2 * kStemTile + 5;   // input samples feeding one tile
constexpr int kStemMaxCin = 16;

__global__ void __launch_bounds__(256) signal_stem_fwd_kernel(const float* __restrict__ x,
                                                               const float* __restrict__ w,
                                                               __nv_bfloat16* __restrict__ y, int Cin, int L,
                                                               int Lo) {
  extern __shared__ float sm[];
  float* xs = sm;                      // [Cin][kStemSpan]
  float* ws = sm + ((Cin * kStemSpan + 3) & ~3);  // [Cin*7][64], 16-byte aligned
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kStemTile;
  const int K = Cin * 7;
  for (int e = threadIdx.x; e < K * kStemCo; e += 256) {
    const int o = e / K, k = e % K;
    ws[k * kStemCo + o] = w[e];
  }
  const int i0 = 2 * p0 - 3;
  for (int e = threadIdx.x; e < Cin * kStemSpan; e += 256) {
    const int ci = e / kStemSpan, j = e % kStemSpan;
    const int i = i0 + j;
    xs[e] = (i >= 0 && i < L) ? x[((size_t)b * Cin + ci) * L + i] : 0.f;
  }
  __syncthreads();
  const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      float wv[8];
      load8f(ws + (ci * 7 + k) * kStemCo + cg * 8, wv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float xv = xs[ci * kStemSpan + 2 * (pl + 32 * i) + k];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv, wv[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + pl + 32 * i;
    if (p < Lo) reinterpret_cast<uint4*>(y + ((size_t)b * Lo + p) * kStemCo)[cg] = pack8(acc[i]);
  }
}

// dW[o][ci][k] += sum_{b,p} dy[b][p][o] * x[b][ci][2p+k-3].  Each CTA walks (b, tile) pairs,
// keeps its share of the 64*Cin*7 outputs in registers and flushes once with atomics.
template <int NOUT>  // outputs per thread = ceil(64*Cin*7 / 256)
__global__ void __launch_bounds__(256) signal_stem_wgrad_kernel(const float* __restrict__ x,
                                                                 const __nv_bfloat16* __restrict__ dy,
                                                                 float* __restrict__ dw, float* __restrict__ ws,
                                                                 int B, int Cin, int L, int Lo, int tiles_per_row) {
  extern __shared__ float sm[];
  float* xs = sm;                     // [Cin][kStemSpan]
  float* ds = sm + ((Cin * kStemSpan + 3) & ~3);  // [kStemTile][64]
  const int K = Cin * 7, total_out = K * kStemCo;
  float acc[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; ++j) acc[j] = 0.f;
  const int o = threadIdx.x & 63;
  const int total_tiles = B * tiles_per_row;
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int b = t / tiles_per_row, p0 = (t % tiles_per_row) * kStemTile;
    const int i0 = 2 * p0 - 3;
    __syncthreads();
    for (int e = threadIdx.x; e < Cin * kStemSpan; e += 256) {
      const int ci = e / kStemSpan, j = e % kStemSpan;
      const int i = i0 + j;
      xs[e] = (i >= 0 && i < L) ? x[((size_t)b * Cin + ci) * L + i] : 0.f;
    }
    for (int e = threadIdx.x; e < kStemTile * 8; e += 256) {
      const int pl = e >> 3, cg = e & 7;
      float f[8];
      if (p0 + pl < Lo) {
        unpack8(reinterpret_cast<const uint4*>(dy + ((size_t)b * Lo + p0 + pl) * kStemCo)[cg], f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ds[pl * kStemCo + cg * 8 + j] = f[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      const int kk = (threadIdx.x >> 6) + 4 * j;  // (ci,k) index handled by this thread
      if (kk < K) {
        const float* xr = xs + (kk / 7) * kStemSpan + (kk % 7);
        float a = acc[j];
#pragma unroll 8
        for (int p = 0; p < kStemTile; ++p) a = fmaf(ds[p * kStemCo + o], xr[2 * p], a);
        acc[j] = a;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    const int kk = (threadIdx.x >> 6) + 4 * j;
    if (kk < K && o * K + kk < total_out) {
      if (ws)  // deterministic path: per-CTA partials, folded in CTA order by signal_stem_wgrad_reduce_kernel
        ws[(size_t)blockIdx.x * total_out + (size_t)o * K + kk] = acc[j];
      else
        atomicAdd(dw + (size_t)o * K + kk, acc[j]);
    }
  }
}

__global__ void __launch_bounds__(256) signal_stem_wgrad_reduce_kernel(const float* __restrict__ ws,
                                                                        float* __restrict__ dw, int total_out,
                                                                        int n_parts) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total_out) return;
  float a0 = 0.f, a1 = 0.f;
  int k = 0;
  for (; k + 1 < n_parts; k += 2) {
    a0 += ws[(size_t)k * total_out + idx];
    a1 += ws[(size_t)(k + 1) * total_out + idx];
  }
  if (k < n_parts) a0 += ws[(size_t)k * total_out + idx];
  dw[idx] += a0 + a1;
}

// ---------------------------------------------------------------------------------------------------------------------
// The same layer on the TENSOR cores (measured: the CUDA-core kernels above were 1.26 of the 3.2 ms training step of
// the 12-lead model at batch 256).  Grouping FOUR consecutive samples of all leads into one 64-channel "pixel"
//        xs4[b][q][j * Cin + ci] = x[b][ci][4q + j]          (bf16, channels >= 4 * Cin are zero)
// turns Conv1d(Cin, 64, k 7, stride 2, pad 3) into a 1x3 stride-1 pad-1 convolution with 64 input and 128 output
// channels over the q axis: output pixel q holds the stem outputs 2q (channels 0..63) and 2q+1 (channels 64..127), i.e.
// [B][Lq][128] IS the [B][2 Lq][64] channels-last stem output.  Output 2q+e reads x[4q + 2e + k - 3], k = 0..6: block tap
// t (block q+t-1), sample j of the block  ->  k = 4 (t-1) + j - 2e + 3.  Forward, weight gradient (and nothing else: the
// stem has no data gradient) then run through the generic tcgen05 kernels; three small kernels do the regrouping.
// Requires 2 * ceil(L/4) == ceil(L/2) (L mod 4 in {0, 3}); other lengths keep the CUDA-core kernels.
__global__ void __launch_bounds__(256) signal_s4d_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xs4,
                                                          int Cin, int L, int Lq, size_t total_vec) {
  // one thread = 8 channels (16 bytes) of one pixel; vector v of a pixel covers channels 8v..8v+7
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total_vec; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i & 7);
    const size_t pix = i >> 3;
    const int q = (int)(pix % Lq);
    const size_t b = pix / Lq;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ch = v * 8 + e;
      const int j = ch / Cin, ci = ch - j * Cin;
      const int t = 4 * q + j;
      f[e] = (j < 4 && t < L) ? __ldg(x + (b * Cin + ci) * (size_t)L + t) : 0.f;
    }
    reinterpret_cast<uint4*>(xs4)[i] = pack8(f);
  }
}

// w [64][Cin][7] fp32 -> w4 [128][1][3][64] bf16 (the [O][R][S][I] operand layout of the forward kernels)
__global__ void signal_stem_w4_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w4, int Cin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // ((e*64 + co) * 3 + t) * 64 + ch
  if (i >= 128 * 3 * 64) return;
  const int ch = i & 63, t = (i >> 6) % 3, o = i / 192;
  const int e = o >> 6, co = o & 63;
  const int j = ch / Cin, ci = ch - j * Cin;
  const int k = 4 * (t - 1) + j - 2 * e + 3;
  const bool ok = j < 4 && k >= 0 && k < 7;
  w4[i] = __float2bfloat16_rn(ok ? w[(co * Cin + ci) * 7 + k] : 0.f);
}

// dw [64][Cin][7] += the entries of dw4 [128][64][1][3] (OIHW of the regrouped convolution) that map onto it
__global__ void signal_stem_dw4_fold_kernel(const float* __restrict__ dw4, float* __restrict__ dw, int Cin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (co * Cin + ci) * 7 + k
  if (i >= 64 * Cin * 7) return;
  const int k = i % 7, ci = (i / 7) % Cin, co = i / (7 * Cin);
  float acc = 0.f;
  for (int e = 0; e < 2; ++e)
    for (int t = 0; t < 3; ++t) {
      const int j = k - 4 * (t - 1) + 2 * e - 3;
      if (j >= 0 && j < 4) acc += dw4[((size_t)(e * 64 + co) * 64 + j * Cin + ci) * 3 + t];
    }
  dw[i] += acc;
}

}  // namespace ecgmm

using namespace ecgmm;

extern "C" int ecgmm_signal_s4d_len(int L) { return L > 0 && (L % 4 == 0 || L % 4 == 3) ? (L + 3) / 4 : 0; }

extern "C" int ecgmm_signal_s4d(const float* x, ecgmm_bf16* xs4, int B, int Cin, int L, void* stream) {
  ECGMM_CHECK(x && xs4, ECGMM_ERR_ARG, "signal_s4d: null pointer");
  ECGMM_CHECK(Cin >= 1 && Cin <= kStemMaxCin, ECGMM_ERR_SHAPE, "signal_s4d: Cin=%d", Cin);
  const int Lq = ecgmm_signal_s4d_len(L);
  ECGMM_CHECK(Lq > 0, ECGMM_ERR_SHAPE, "signal_s4d: L=%d is not 0 or 3 modulo 4 (use the direct stem kernels)", L);
  if (B == 0) return ECGMM_OK;
  const size_t total = (size_t)B * Lq * 8;
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)num_sms() * 16) blocks = (size_t)num_sms() * 16;
  signal_s4d_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(xs4), Cin, L,
                                                                        Lq, total);
  return check_launch("signal_s4d_kernel");
}

extern "C" int ecgmm_signal_stem_w4(const float* w, ecgmm_bf16* w4, int Cin, void* stream) {
  ECGMM_CHECK(w && w4, ECGMM_ERR_ARG, "signal_stem_w4: null pointer");
  ECGMM_CHECK(Cin >= 1 && Cin <= kStemMaxCin, ECGMM_ERR_SHAPE, "signal_stem_w4: Cin=%d", Cin);
  signal_stem_w4_kernel<<<ceil_div(128 * 3 * 64, 256), 256, 0, (cudaStream_t)stream>>>(
      w, reinterpret_cast<__nv_bfloat16*>(w4), Cin);
  return check_launch("signal_stem_w4_kernel");
}

extern "C" int ecgmm_signal_stem_dw4_fold(const float* dw4, float* dw, int Cin, void* stream) {
  ECGMM_CHECK(dw4 && dw, ECGMM_ERR_ARG, "signal_stem_dw4_fold: null pointer");
  ECGMM_CHECK(Cin >= 1 && Cin <= kStemMaxCin, ECGMM_ERR_SHAPE, "signal_stem_dw4_fold: Cin=%d", Cin);
  signal_stem_dw4_fold_kernel<<<ceil_div(64 * Cin * 7, 256), 256, 0, (cudaStream_t)stream>>>(dw4, dw, Cin);
  return check_launch("signal_stem_dw4_fold_kernel");
}

extern "C" int ecgmm_signal_stem_fwd(const float* x, const float* w, ecgmm_bf16* y, int B, int Cin, int L,
                                     void* stream) {
  ECGMM_CHECK(x && w && y, ECGMM_ERR_ARG, "signal_stem_fwd: null pointer");
  ECGMM_CHECK(Cin >= 1 && Cin <= kStemMaxCin && L >= 1, ECGMM_ERR_SHAPE, "signal_stem_fwd: Cin=%d L=%d", Cin, L);
  ECGMM_CHECK(B <= 65535, ECGMM_ERR_SHAPE, "signal_stem_fwd: batch %d too large", B);
  if (B == 0) return ECGMM_OK;
  const int Lo = (L - 1) / 2 + 1;
  const size_t smem = ((size_t)((Cin * kStemSpan + 3) & ~3) + (size_t)Cin * 7 * kStemCo) * sizeof(float);
  static bool configured[kMaxDevices] = {};  // the shared-memory limit of a kernel is a per-device attribute
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(signal_stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    configured[ds] = true;
  }
  signal_stem_fwd_kernel<<<dim3(ceil_div(Lo, kStemTile), B), 256, smem, (cudaStream_t)stream>>>(
      x, w, reinterpret_cast<__nv_bfloat16*>(y), Cin, L, Lo);
  return check_launch("signal_stem_fwd_kernel");
}

static int signal_stem_wgrad_grid(int B, int L) {
  const int Lo = (L - 1) / 2 + 1;
  int grid = B * ceil_div(Lo, kStemTile);
  if (grid > num_sms() * 2) grid = num_sms() * 2;
  return grid;
}

extern "C" long long ecgmm_signal_stem_wgrad_workspace(int B, int Cin, int L) {
  if (B <= 0 || Cin < 1 || Cin > kStemMaxCin || L < 1) return 0;
  return (long long)signal_stem_wgrad_grid(B, L) * Cin * 7 * kStemCo * (long long)sizeof(float);
}

extern "C" int ecgmm_signal_stem_wgrad(const float* x, const ecgmm_bf16* dy, float* dw, int B, int Cin, int L,
                                       void* workspace, long long workspace_bytes, void* stream) {
  ECGMM_CHECK(x && dy && dw, ECGMM_ERR_ARG, "signal_stem_wgrad: null pointer");
  ECGMM_CHECK(Cin >= 1 && Cin <= kStemMaxCin && L >= 1, ECGMM_ERR_SHAPE, "signal_stem_wgrad: Cin=%d L=%d", Cin, L);
  if (B == 0) return ECGMM_OK;
  const int Lo = (L - 1) / 2 + 1;
  const int tiles = ceil_div(Lo, kStemTile);
  float* ws = (workspace && workspace_bytes >= ecgmm_signal_stem_wgrad_workspace(B, Cin, L))
                  ? static_cast<float*>(workspace) : nullptr;
  const size_t smem = ((size_t)((Cin * kStemSpan + 3) & ~3) + (size_t)kStemTile * kStemCo) * sizeof(float);
  static bool configured[kMaxDevices] = {};  // the shared-memory limit of a kernel is a per-device attribute
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(signal_stem_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    64 * 1024));
    ECGMM_CUDA(cudaFuncSetAttribute(signal_stem_wgrad_kernel<28>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    64 * 1024));
    configured[ds] = true;
  }
  const int grid = signal_stem_wgrad_grid(B, L);
  const __nv_bfloat16* dyb = reinterpret_cast<const __nv_bfloat16*>(dy);
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 1)
    signal_stem_wgrad_kernel<2><<<grid, 256, smem, st>>>(x, dyb, dw, ws, B, Cin, L, Lo, tiles);
  else
    signal_stem_wgrad_kernel<28><<<grid, 256, smem, st>>>(x, dyb, dw, ws, B, Cin, L, Lo, tiles);
  int rc = check_launch("signal_stem_wgrad_kernel");
  if (rc || !ws) return rc;
  const int total_out = Cin * 7 * kStemCo;
  signal_stem_wgrad_reduce_kernel<<<ceil_div(total_out, 256), 256, 0, st>>>(ws, dw, total_out, grid);
  return check_launch("signal_stem_wgrad_reduce_kernel");
}
