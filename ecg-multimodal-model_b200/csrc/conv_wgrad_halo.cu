// Weight gradient of stride-1 RxS convolutions (3x3 of the ResNet18 BasicBlocks, 1x3 of the 1-D
// BasicBlocks) with input-halo reuse.
//
//   dW[(tap,cin)][cout] = sum_pixels X[pixel + tap][cin] * dY[pixel][cout]        (GEMM K = pixels)
//
// The older igemm_tn_kernel stages one TMA box per filter tap, i.e. re-reads every input pixel
// R*S times from L2 (measured: 1.8 GB of L2->SM traffic for 0.32 GB of tensors, 14 B/clk/SM, tensor
// pipe 10 % busy).  Here a pipeline stage is KP consecutive output pixels of ONE image row; the
// R input rows it touches are staged once each as a box of KP+S-1 pixels, and filter tap (r,s) is
// the SAME shared-memory box read from row s onwards: the SWIZZLE_128B operand descriptor simply
// starts s*128 bytes later (the swizzle is a function of the absolute shared-memory address, probed
// by tools/desc_probe.py, so any 128-byte row is a legal start).  Both operands are MN-major.
//
// A stage covers RPS consecutive output rows of a KP-pixel column strip: RPS dy boxes + RPS+R-1 input rows, so
// an input row is fetched (RPS+R-1)/RPS times instead of R times.  The kernel is bound by L2->SM traffic, not by
// the MMA floor (ncu: tensor pipe 51 % busy at 66 KB per 1280-clk stage = 52 B/clk/SM with RPS = 1, KP = 128);
// (KP, RPS) is chosen per layer to minimise the bytes staged per output pixel with >= 3 stages in flight.
//
// Work decomposition: group = (64-channel slice of Cin) x (64-channel slice of Cout); every group is a
// [R*S*64] x [64] output held in TMEM as ceil(R*S/2) accumulators of 128 lanes x 64 columns; the
// pixel range is split across CTAs (split-K) and partial results are combined with fp32 atomics.
#include "common.h"
#include "ptx.cuh"

#include <stdlib.h>

namespace ecgmm {

constexpr int kHaloMaxStages = 6;

struct alignas(64) WgHaloParams {
  CUtensorMap x_map;   // x  [N][H][W][Cin],   box (64, KP+S-1, 1, 1)
  CUtensorMap dy_map;  // dy [N][Ho][Wo][Cout], box (64, KP, 1, 1)
  float* dw;           // [Cout][Cin][R][S]
  float* ws;           // optional split-K workspace [group][ksplit][slot][64 cols][128 rows]; NULL -> atomics
  int R, S, padH, padW;
  int KP, kmma;              // pixels per stage row, KP/16
  int rps;                   // output rows per stage
  int x_box_bytes, x_box_stride, dy_box_bytes, dy_box_stride, stage_bytes, stages;
  int tiles_w, Ho, row_groups, total_kblocks;  // row_groups = ceil(Ho / rps)
  int cin_chunks, cout_chunks, ksplit;
  int Cin, Cout;
};

__global__ void __launch_bounds__(192, 1) wgrad_halo_kernel(const __grid_constant__ WgHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* empty = full + kHaloMaxStages;
  uint64_t* tfull = empty + kHaloMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int RS = p.R * p.S;
  const int n_slots = (RS + 1) >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.x_map);
    tma_prefetch_desc(&p.dy_map);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // blockIdx -> (k-split id, cin chunk, cout chunk)
  const int ks = blockIdx.x % p.ksplit;
  const int g = blockIdx.x / p.ksplit;
  const int cc = g % p.cin_chunks;
  const int nt = g / p.cin_chunks;
  const int per = (p.total_kblocks + p.ksplit - 1) / p.ksplit;
  const int kb0 = ks * per;
  const int kb1 = min(p.total_kblocks, kb0 + per);
  const bool has_work = kb0 < kb1;

  if (has_work) {
    if (warp == 0) {
      if (elect_one()) {  // single elected thread: no ELECT serialisation loops around UTMALDG
        int stage = 0;
        uint32_t phase = 0;
        const int n_xrows = p.rps + p.R - 1;
        const uint32_t tx = p.rps * p.dy_box_bytes + n_xrows * p.x_box_bytes;
        const int x_base = p.rps * p.dy_box_stride;
        int twi = kb0 % p.tiles_w;
        int row = kb0 / p.tiles_w;  // img * row_groups + row group
        int og = row % p.row_groups, img = row / p.row_groups;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int w0 = twi * p.KP;
          const int oh0 = og * p.rps;
          uint8_t* st = smem + stage * p.stage_bytes;
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], tx);
          // rows past Ho / outside the image are zero-filled by TMA and contribute nothing
          for (int j = 0; j < p.rps; ++j)
            tma_load_4d(st + j * p.dy_box_stride, &p.dy_map, &full[stage], nt * 64, w0, oh0 + j, img);
          for (int r = 0; r < n_xrows; ++r)
            tma_load_4d(st + x_base + r * p.x_box_stride, &p.x_map, &full[stage], cc * 64, w0 - p.padW,
                        oh0 + r - p.padH, img);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          if (++twi == p.tiles_w) {
            twi = 0;
            if (++og == p.row_groups) {
              og = 0;
              ++img;
            }
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
        const uint32_t s_addr = smem_u32(smem);
        // per-slot A descriptors relative to the stage base, built once
        constexpr int kMaxSlots = 8;
        uint64_t a_rel[kMaxSlots];
#pragma unroll
        for (int i = 0; i < kMaxSlots; ++i) {
          const int t0 = min(2 * i, RS - 1), t1 = min(2 * i + 1, RS - 1);
          const uint32_t a0 = p.rps * p.dy_box_stride + (t0 / p.S) * p.x_box_stride + (t0 % p.S) * 128;
          const uint32_t a1 = p.rps * p.dy_box_stride + (t1 / p.S) * p.x_box_stride + (t1 % p.S) * 128;
          a_rel[i] = make_sw128_desc(s_addr + a0, a1 - a0, 1024);
        }
        const uint64_t b_rel = make_sw128_desc(s_addr, 0, 1024);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t so = (uint64_t)((stage * p.stage_bytes) >> 4);
          for (int j = 0; j < p.rps; ++j) {  // output row j of the stage: dy box j against input rows j .. j+R-1
            const uint64_t ao = so + (uint64_t)((j * p.x_box_stride) >> 4), bo = so + (uint64_t)((j * p.dy_box_stride) >> 4);
#pragma unroll
            for (int i = 0; i < kMaxSlots; ++i) {
              if (i < n_slots) {
                const uint64_t a_desc = a_rel[i] + ao, b_desc = b_rel + bo;
                for (int k = 0; k < p.kmma; ++k)  // 16 pixel rows = 2048 B further into both boxes
                  umma_bf16(tmem_base + i * 64, a_desc + k * 128, b_desc + k * 128, idesc,
                            (kb > kb0) || (j > 0) || (k > 0));
              }
            }
          }
          umma_commit(&empty[stage]);
          if (kb == kb1 - 1) umma_commit(tfull);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else {
      const int quad = warp & 3;
      const int m_row = quad * 32 + lane;
      mbar_wait(tfull, 0);
      tc_fence_after();
      for (int i = 0; i < n_slots; ++i) {
        const int tap = 2 * i + (m_row >> 6);
        const bool ok = tap < RS;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + i * 64;
        if (p.ws) {
          // split-K partial, column-major so that the 32 lanes (= rows) of a warp store 128 contiguous bytes
          float* dst0 = p.ws + ((size_t)(g * p.ksplit + ks) * n_slots + i) * (64 * 128) + m_row;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(t_addr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) dst0[(c * 32 + j) * 128] = __uint_as_float(r[j]);
          }
        } else {
          const int cin = cc * 64 + (m_row & 63);
          float* dst0 = p.dw + ((size_t)(nt * 64) * p.Cin + cin) * RS + (ok ? tap : 0);
          const size_t col_stride = (size_t)p.Cin * RS;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(t_addr + c * 32, r);
            tmem_ld_wait();
            if (ok) {
              float* dst = dst0 + (size_t)(c * 32) * col_stride;
#pragma unroll
              for (int j = 0; j < 32; ++j) atomicAdd(dst + j * col_stride, __uint_as_float(r[j]));
            }
          }
        }
      }
    }
  } else if (p.ws && warp >= 2) {
    // a CTA without pixels still owns a workspace slice: the reduction kernel reads every slice
    const int m_row = (warp & 3) * 32 + lane;
    for (int i = 0; i < n_slots; ++i) {
      float* dst0 = p.ws + ((size_t)(g * p.ksplit + ks) * n_slots + i) * (64 * 128) + m_row;
      for (int c = 0; c < 64; ++c) dst0[c * 128] = 0.f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Second stage of the split-K reduction: dw[cout][cin][tap] += sum over the ksplit partials.
__global__ void __launch_bounds__(256) wgrad_halo_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                                 int ksplit, int n_slots, int RS, int cin_chunks,
                                                                 int Cin, size_t total) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // ((g*n_slots + slot)*64 + c)*128 + m
  if (idx >= total) return;
  const int m = (int)(idx & 127);
  const int c = (int)((idx >> 7) & 63);
  const size_t gs = idx >> 13;
  const int slot = (int)(gs % n_slots);
  const int g = (int)(gs / n_slots);
  const int tap = 2 * slot + (m >> 6);
  if (tap >= RS) return;
  const size_t slice = (size_t)n_slots * 64 * 128;
  const float* src = ws + (size_t)g * ksplit * slice + ((size_t)slot * 64 + c) * 128 + m;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // independent chains: 4+ loads in flight per thread
  int k = 0;
  for (; k + 3 < ksplit; k += 4) {
    a0 += src[(size_t)k * slice];
    a1 += src[(size_t)(k + 1) * slice];
    a2 += src[(size_t)(k + 2) * slice];
    a3 += src[(size_t)(k + 3) * slice];
  }
  for (; k < ksplit; ++k) a0 += src[(size_t)k * slice];
  const float acc = (a0 + a1) + (a2 + a3);
  const int cc = g % cin_chunks, nt = g / cin_chunks;
  const int cin = cc * 64 + (m & 63), cout = nt * 64 + c;
  dw[((size_t)cout * Cin + cin) * RS + tap] += acc;
}

// Picks (KP, RPS): KP pixels (multiple of 16, <= 128) x RPS output rows per stage.
// Measured on B200 (tools/conv_bench.py sweep, profiles/r01_wgrad_shape_sweep.txt): the kernel is NOT bound by
// L2->SM traffic -- halving the staged bytes per pixel (RPS 1 -> 2..7) changes the time by < 4 %, while small KP
// (more, shorter TMA boxes and MMA runs) costs up to 2x.  What matters is the pixel slots wasted at the right
// edge of a row, so KP minimises ceil(Wo/KP)*KP (ties: larger KP) and RPS = 2 is taken only where three stages
// still fit (a small, free reduction of L2 traffic).  The cap is the shared-memory operand bandwidth of the
// N = 64 MMA shape: 6 KB of operands per 32-clk M128xN64xK16 MMA = 192 B/clk against ~128 B/clk.
static int round1k(int v) { return (v + 1023) & ~1023; }
static int halo_stage_bytes(int kp, int rps, int R, int S) {
  return rps * round1k(kp * 128) + (rps + R - 1) * round1k((kp + S - 1) * 128);
}
static void pick_shape(int Ho, int Wo, int R, int S, int* kp_out, int* rps_out) {
  const int budget = 220 * 1024;
  int best_kp = 64;
  long best_cost = -1;
  for (int kp = 128; kp >= 32; kp -= 16) {
    const long cost = (long)ceil_div(Wo, kp) * kp;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_kp = kp;
    }
  }
  int best_rps = 1;
  if (Ho >= 2 && 3 * halo_stage_bytes(best_kp, 2, R, S) <= budget) best_rps = 2;
  *kp_out = best_kp;
  *rps_out = best_rps;
  // development knobs (tools/conv_bench.py sweeps): force a shape
  const char* ekp = getenv("ECGMM_WG_KP");
  const char* erps = getenv("ECGMM_WG_RPS");
  if (ekp && erps) {
    const int kp = atoi(ekp), rps = atoi(erps);
    if (kp >= 16 && kp <= 128 && kp % 16 == 0 && rps >= 1 && rps <= (Ho > 1 ? Ho : 1) &&
        2 * halo_stage_bytes(kp, rps, R, S) <= budget) {
      *kp_out = kp;
      *rps_out = rps;
    }
  }
}

bool wgrad_halo_supported(int Cin, int Cout, int R, int S, int stride) {
  return stride == 1 && Cin % 64 == 0 && Cout % 64 == 0 && R * S > 1 && R * S <= 15 && (R == 1 || R == 3) &&
         (S == 3);
}

static void halo_split(int N, int H, int W, int Cin, int Cout, int R, int S, int padH, int padW, int* kp, int* total_kb,
                       int* groups, int* ksplit) {
  const int Ho = H + 2 * padH - R + 1, Wo = W + 2 * padW - S + 1;
  int rps;
  pick_shape(Ho, Wo, R, S, kp, &rps);
  *total_kb = N * ceil_div(Ho, rps) * ceil_div(Wo, *kp);
  *groups = (Cin / 64) * (Cout / 64);
  int ks = num_sms() / *groups;
  if (ks < 1) ks = 1;
  if (ks > *total_kb) ks = *total_kb > 0 ? *total_kb : 1;
  *ksplit = ks;
}

size_t wgrad_halo_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S, int padH, int padW) {
  int kp, total_kb, groups, ksplit;
  halo_split(N, H, W, Cin, Cout, R, S, padH, padW, &kp, &total_kb, &groups, &ksplit);
  return (size_t)groups * ksplit * ((R * S + 1) / 2) * 64 * 128 * sizeof(float);
}

int launch_wgrad_halo(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, int N, int H, int W, int Cin,
                      int Cout, int R, int S, int padH, int padW, void* workspace, size_t ws_bytes,
                      cudaStream_t st) {
  const int Ho = H + 2 * padH - R + 1, Wo = W + 2 * padW - S + 1;
  WgHaloParams p;
  memset(&p, 0, sizeof(p));
  p.R = R;
  p.S = S;
  p.padH = padH;
  p.padW = padW;
  pick_shape(Ho, Wo, R, S, &p.KP, &p.rps);
  p.kmma = p.KP / 16;
  const int xw = p.KP + S - 1;
  p.x_box_bytes = xw * 128;
  p.x_box_stride = (p.x_box_bytes + 1023) & ~1023;
  p.dy_box_bytes = p.KP * 128;
  p.dy_box_stride = (p.dy_box_bytes + 1023) & ~1023;
  p.stage_bytes = p.rps * p.dy_box_stride + (p.rps + R - 1) * p.x_box_stride;
  int stages = (220 * 1024) / p.stage_bytes;
  if (stages > kHaloMaxStages) stages = kHaloMaxStages;
  ECGMM_CHECK(stages >= 2, ECGMM_ERR_SHAPE, "wgrad_halo: stage of %d bytes does not fit twice", p.stage_bytes);
  p.stages = stages;
  p.tiles_w = ceil_div(Wo, p.KP);
  p.Ho = Ho;
  p.row_groups = ceil_div(Ho, p.rps);
  p.total_kblocks = N * p.row_groups * p.tiles_w;
  p.cin_chunks = Cin / 64;
  p.cout_chunks = Cout / 64;
  const int groups = p.cin_chunks * p.cout_chunks;
  int ksplit = num_sms() / groups;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > p.total_kblocks) ksplit = p.total_kblocks;
  p.ksplit = ksplit;
  p.Cin = Cin;
  p.Cout = Cout;
  p.dw = dw;
  const uint64_t e = 2;
  int rc = make_tmap_4d(&p.x_map, x, Cin, W, H, N, (uint64_t)Cin * e, (uint64_t)W * Cin * e, (uint64_t)H * W * Cin * e,
                        64, xw, 1);
  if (rc) return rc;
  rc = make_tmap_4d(&p.dy_map, dy, Cout, Wo, Ho, N, (uint64_t)Cout * e, (uint64_t)Wo * Cout * e,
                    (uint64_t)Ho * Wo * Cout * e, 64, p.KP, 1);
  if (rc) return rc;
  const int smem = p.stages * p.stage_bytes + 256 + 1024;
  static int configured = 0;
  if (configured < smem) {
    ECGMM_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  const size_t need = wgrad_halo_workspace_bytes(N, H, W, Cin, Cout, R, S, padH, padW);
  p.ws = (workspace && ws_bytes >= need) ? reinterpret_cast<float*>(workspace) : nullptr;
  wgrad_halo_kernel<<<groups * ksplit, 192, smem, st>>>(p);
  rc = check_launch("wgrad_halo_kernel");
  if (rc || !p.ws) return rc;
  const int n_slots = (R * S + 1) / 2;
  const size_t total = (size_t)groups * n_slots * 64 * 128;
  wgrad_halo_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.ws, dw, ksplit, n_slots, R * S,
                                                                           p.cin_chunks, Cin, total);
  return check_launch("wgrad_halo_reduce_kernel");
}

}  // namespace ecgmm
