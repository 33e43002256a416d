// ResNet18 stem (7x7/s2 conv over 3 channels, evaluated as 4 row-taps of a K = 64 GEMM over the
// space-to-depth image, see layout.cu / ecgmm.h) with ROLLING input rows.
//
// Output row oh needs s2d rows oh..oh+3; consecutive output rows share 3 of them.  The generic
// kernels stage 4 boxes per 128-pixel tile (and re-fetch the weights per tile): 96 KB moved for
// 2.4 algorithmic MFLOP, which pins the stem on the L2->SM cap.  Here a persistent CTA walks DOWN
// a 128-pixel column strip; every s2d row segment is loaded exactly once into a ring of shared-
// memory slots and the four taps of an output row are the four most recent slots.  Weights (32 KB)
// are CTA-resident.  Forward: D[pixel][cout]; weight gradient: D[(ra, e)][cout] accumulated over
// the whole strip set in TMEM, fp32 atomics into dW[64][3][7][7] at the end.
#include "common.h"
#include "ptx.cuh"

namespace ecgmm {

constexpr int kSrTile = 128;             // pixels per row segment
constexpr int kSrSlot = kSrTile * 128;   // 16 KiB: 128 pixels x 64 (overlapping) bf16
constexpr int kSrRing = 8;

struct alignas(64) StemRingParams {
  CUtensorMap x_map;   // overlapping-column view of xs: (64, Wo, Hs, N), box (64, 128, 1, 1)
  CUtensorMap w_map;   // fwd: w_s2d [64][256], box (64, 64)
  CUtensorMap dy_map;  // wgrad: dy [N][Ho][Wo][64], box (64, 128, 1, 1); fwd: the OUTPUT y, same shape (TMA store)
  int Ho, Wo, tiles_w, n_strips;
  __nv_bfloat16* out;  // fwd: y [N][Ho][Wo][64]
  float* dw;           // wgrad: [64][3][7][7]
  float* ws;           // wgrad: per-CTA partial tiles [grid][2][128][64] (deterministic fold); NULL: fp32 atomics
  float* psum;         // fwd<STATS>: per-CTA BatchNorm partial sums of the stored output, [gridDim.x][64]
  float* psq;          //             ... and sums of squares
};

// ------------------------------------------------------------------------------------- forward
struct StemFwdSmem {
  static constexpr int kW = 4 * 8192;
  static constexpr int kOut = kW + kSrRing * kSrSlot;  // 2 output staging tiles of 16 KiB
  static constexpr int kBarOff = kOut + 2 * kSrSlot;
  static constexpr int kBytes = kBarOff + 512 + 1024;
};

// STATS: per-channel sum / sum of squares of the stored output, accumulated in 128 registers per epilogue thread over
// the whole kernel and folded across threads once at the end (same scheme as igemm_nt_stack_kernel<true>).
// Epilogue: warps 2..9, two per TMEM lane quadrant, 32 of the 64 channels each (the per-tile epilogue chain of a single
// warp per quadrant was what bounded this kernel once the statistics were added: 2.7 -> 3.4 ms at batch 512).
constexpr int kSrFwdThreads = 320;
template <bool STATS>
__global__ void __launch_bounds__(kSrFwdThreads, 1) stem_fwd_ring_kernel(const __grid_constant__ StemRingParams p) {
  using L = StemFwdSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sX = smem + L::kW;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + kSrRing;
  uint64_t* tfull = empty + kSrRing;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hs = p.Ho + 3;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.x_map);
    tma_prefetch_desc(&p.w_map);
    tma_prefetch_desc(&p.dy_map);
    for (int i = 0; i < kSrRing; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    mbar_init(wfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(wfull, L::kW);
      for (int t = 0; t < 4; ++t) tma_load_2d(sW + t * 8192, &p.w_map, wfull, t * 64, 0);
      uint32_t q = 0;  // running row-load counter: slot = q % ring, phase = (q / ring) & 1
      for (int s = blockIdx.x; s < p.n_strips; s += gridDim.x) {
        const int img = s / p.tiles_w, w0 = (s % p.tiles_w) * kSrTile;
        for (int r = 0; r < Hs; ++r, ++q) {
          const int slot = q % kSrRing;
          mbar_wait(&empty[slot], ((q / kSrRing) & 1) ^ 1);
          mbar_expect_tx(&full[slot], kSrSlot);
          tma_load_4d(sX + slot * kSrSlot, &p.x_map, &full[slot], 0, w0, r, img);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      const uint64_t w_desc0 = make_sw128_desc(smem_u32(sW), 0, 1024);
      const uint64_t x_desc0 = make_sw128_desc(smem_u32(sX), 0, 1024);
      mbar_wait(wfull, 0);
      tc_fence_after();
      uint32_t q0 = 0;  // counter of the strip's row 0
      int it = 0;
      for (int s = blockIdx.x; s < p.n_strips; s += gridDim.x, q0 += Hs) {
        int waited = 0;
        for (int oh = 0; oh < p.Ho; ++oh, ++it) {
          while (waited <= oh + 3) {
            const uint32_t q = q0 + waited;
            mbar_wait(&full[q % kSrRing], (q / kSrRing) & 1);
            ++waited;
          }
          const int acc = it & 1;
          mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 64;
#pragma unroll
          for (int ra = 0; ra < 4; ++ra) {
            const uint64_t a_desc = x_desc0 + (uint64_t)(((q0 + oh + ra) % kSrRing) * (kSrSlot >> 4));
            const uint64_t w_desc = w_desc0 + (uint64_t)(ra * (8192 >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_desc + 2 * k, w_desc + 2 * k, idesc, (ra | k) != 0);
          }
          umma_commit(&tfull[acc]);
          umma_commit(&empty[(q0 + oh) % kSrRing]);  // row oh is not needed by later output rows
          if (oh == p.Ho - 1)
            for (int r = p.Ho; r < Hs; ++r) umma_commit(&empty[(q0 + r) % kSrRing]);
        }
      }
    }
  } else {
    // epilogue: TMEM -> bf16 -> swizzled staging tile -> one TMA tensor store per tile (see conv_nt_halo.cu)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;  // which 32 of the 64 channels
    const int m_row = quad * 32 + lane;
    const int et = (int)threadIdx.x - 64;
    const bool leader = (threadIdx.x == 64);
    uint8_t* sOut = smem + L::kOut;
    float ssum[STATS ? 32 : 1], ssq[STATS ? 32 : 1];
    if constexpr (STATS) {
#pragma unroll
      for (int i = 0; i < 32; ++i) ssum[i] = ssq[i] = 0.f;
    }
    int it = 0;
    for (int s = blockIdx.x; s < p.n_strips; s += gridDim.x) {
      const int img = s / p.tiles_w, w0 = (s % p.tiles_w) * kSrTile;
      const bool in_image = w0 + m_row < p.Wo;
      for (int oh = 0; oh < p.Ho; ++oh, ++it) {
        const int acc = it & 1;
        uint8_t* buf = sOut + acc * kSrSlot;
        if (leader) tma_store_wait_read<1>();
        named_bar_sync(1, 256);
        mbar_wait(&tfull[acc], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 64 + half * 32;
        uint8_t* row = buf + m_row * 128;
        {
          const int c = half;
          uint32_t r[32];
          tmem_ld_32x32(t_addr, r);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 v;
            __nv_bfloat162* vb = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              vb[j] = __floats2bfloat162_rn(__uint_as_float(r[g * 8 + 2 * j]), __uint_as_float(r[g * 8 + 2 * j + 1]));
            *reinterpret_cast<uint4*>(row + (((c * 4 + g) ^ (m_row & 7)) << 4)) = v;
            if constexpr (STATS) {
              if (in_image) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 s2 = __bfloat1622float2(vb[j]);
                  const int ch = g * 8 + 2 * j;
                  ssum[ch] += s2.x;
                  ssq[ch] = fmaf(s2.x, s2.x, ssq[ch]);
                  ssum[ch + 1] += s2.y;
                  ssq[ch + 1] = fmaf(s2.y, s2.y, ssq[ch + 1]);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        fence_proxy_async_smem();
        named_bar_sync(1, 256);
        if (leader) {
          tma_store_4d(&p.dy_map, buf, 0, w0, oh, img);
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_all<0>();
    if constexpr (STATS) {
      float* scr = reinterpret_cast<float*>(sX);  // the input ring is idle: all MMAs of this CTA have completed
      named_bar_sync(1, 256);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        scr[m_row * 129 + half * 32 + i] = ssum[i];
        scr[m_row * 129 + 64 + half * 32 + i] = ssq[i];
      }
      named_bar_sync(1, 256);
      if (et < 128) {
        double acc = 0.0;
        for (int r = 0; r < 128; ++r) acc += (double)scr[r * 129 + et];
        float* dst = (et < 64) ? p.psum : p.psq;
        dst[(size_t)blockIdx.x * 64 + (et & 63)] = (float)acc;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------- weight gradient
// Ring slot kSrRing mirrors slot 0, so that the atom pair (slot ring-1, slot 0) is also contiguous
// (the M = 128 A operand is two 64-row atoms one slot apart; LBO cannot be negative).
struct StemWgSmem {
  static constexpr int kX = (kSrRing + 1) * kSrSlot;  // + mirror slot
  static constexpr int kDyRing = 3;
  static constexpr int kBarOff = kX + kDyRing * kSrSlot;
  static constexpr int kBytes = kBarOff + 512 + 1024;
};

__global__ void __launch_bounds__(192, 1) stem_wgrad_ring_kernel(const __grid_constant__ StemRingParams p) {
  using L = StemWgSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sX = smem;
  uint8_t* sD = smem + L::kX;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + kSrRing;
  uint64_t* dfull = empty + kSrRing;
  uint64_t* dempty = dfull + L::kDyRing;
  uint64_t* tfull = dempty + L::kDyRing;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hs = p.Ho + 3;
  const bool has_work = (int)blockIdx.x < p.n_strips;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.x_map);
    tma_prefetch_desc(&p.dy_map);
    for (int i = 0; i < kSrRing; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < L::kDyRing; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (has_work) {
    if (warp == 0) {
      if (elect_one()) {
        uint32_t q = 0, qd = 0;
        for (int s = blockIdx.x; s < p.n_strips; s += gridDim.x) {
          const int img = s / p.tiles_w, w0 = (s % p.tiles_w) * kSrTile;
          for (int r = 0; r < Hs; ++r, ++q) {
            const int slot = q % kSrRing;
            mbar_wait(&empty[slot], ((q / kSrRing) & 1) ^ 1);
            mbar_expect_tx(&full[slot], slot == 0 ? 2 * kSrSlot : kSrSlot);
            tma_load_4d(sX + slot * kSrSlot, &p.x_map, &full[slot], 0, w0, r, img);
            if (slot == 0) tma_load_4d(sX + kSrRing * kSrSlot, &p.x_map, &full[slot], 0, w0, r, img);
            // the dY row of output row r-3 travels with the s2d row that completes its window
            if (r >= 3) {
              const int ds = qd % L::kDyRing;
              mbar_wait(&dempty[ds], ((qd / L::kDyRing) & 1) ^ 1);
              mbar_expect_tx(&dfull[ds], kSrSlot);
              tma_load_4d(sD + ds * kSrSlot, &p.dy_map, &dfull[ds], 0, w0, r - 3, img);
              ++qd;
            }
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
        const uint32_t x_addr = smem_u32(sX), d_addr = smem_u32(sD);
        uint32_t q0 = 0, qd = 0;
        bool first = true;
        for (int s = blockIdx.x; s < p.n_strips; s += gridDim.x, q0 += Hs) {
          int waited = 0;
          for (int oh = 0; oh < p.Ho; ++oh, ++qd) {
            while (waited <= oh + 3) {
              const uint32_t q = q0 + waited;
              mbar_wait(&full[q % kSrRing], (q / kSrRing) & 1);
              ++waited;
            }
            const int ds = qd % L::kDyRing;
            mbar_wait(&dfull[ds], (qd / L::kDyRing) & 1);
            tc_fence_after();
            const uint64_t b_desc = make_sw128_desc(d_addr + ds * kSrSlot, 0, 1024);
#pragma unroll
            for (int pair = 0; pair < 2; ++pair) {
              // atoms (ra = 2*pair, 2*pair+1): two consecutive ring slots (slot ring-1 pairs with the mirror)
              const uint32_t s0 = (q0 + oh + 2 * pair) % kSrRing;
              const uint64_t a_desc = make_sw128_desc(x_addr + s0 * kSrSlot, kSrSlot, 1024);
#pragma unroll
              for (int k = 0; k < kSrTile / 16; ++k)
                umma_bf16(tmem_base + pair * 64, a_desc + k * 128, b_desc + k * 128, idesc, !first || k > 0);
            }
            first = false;
            umma_commit(&dempty[ds]);
            umma_commit(&empty[(q0 + oh) % kSrRing]);
            if (oh == p.Ho - 1)
              for (int r = p.Ho; r < Hs; ++r) umma_commit(&empty[(q0 + r) % kSrRing]);
          }
        }
        umma_commit(tfull);
      }
    } else {
      const int quad = warp & 3;
      const int m_row = quad * 32 + lane;
      mbar_wait(tfull, 0);
      tc_fence_after();
      for (int pair = 0; pair < 2; ++pair) {
        // row m of accumulator `pair`: tap ra = 2*pair + (m >> 6), e = m & 63 = sa*16 + ch,
        // ch = (dr*2+ds)*3 + c  ->  w[cout][c][2*ra+dr][2*sa+ds]
        const int ra = 2 * pair + (m_row >> 6), e = m_row & 63;
        const int sa = e >> 4, ch = e & 15;
        const int dr = ch / 6, dsx = (ch % 6) / 3, c = ch % 3;
        const int r7 = 2 * ra + dr, s7 = 2 * sa + dsx;
        const bool ok = ch < 12 && r7 < 7 && s7 < 7;
        float* dst0 = p.dw + ((size_t)c * 7 + r7) * 7 + s7;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + pair * 64;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t r[32];
          tmem_ld_32x32(t_addr + cc * 32, r);
          tmem_ld_wait();
          if (p.ws) {  // deterministic path: partial [pair][128 rows][64 cout] of this CTA, folded by stem_wgrad_reduce
            float4* dst = reinterpret_cast<float4*>(p.ws + (((size_t)blockIdx.x * 2 + pair) * 128 + m_row) * 64 + cc * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          } else if (ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst0 + (size_t)(cc * 32 + j) * 147, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

static int stem_view(CUtensorMap* m, const void* xs, int N, int Ho, int Wo) {
  const int Hs = Ho + 3, Ws = Wo + 3;
  // column b of the view = the 64 contiguous elements (4 pixels x 16 ch) starting at pixel b
  return make_tmap_4d(m, xs, 64, Wo, Hs, N, 32, (uint64_t)Ws * 32, (uint64_t)Hs * Ws * 32, 64, kSrTile, 1);
}

int stem_fwd_ring_stats_rows(int N, int H, int W) {
  (void)H;
  const int strips = N * ceil_div((W - 1) / 2 + 1, kSrTile);
  return strips < num_sms() ? strips : num_sms();
}

int launch_stem_fwd_ring(const void* xs, const void* w_s2d, __nv_bfloat16* y, int N, int H, int W, cudaStream_t st,
                         float* psum, float* psq) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  StemRingParams p;
  memset(&p, 0, sizeof(p));
  p.psum = psum;
  p.psq = psq;
  p.Ho = Ho;
  p.Wo = Wo;
  p.tiles_w = ceil_div(Wo, kSrTile);
  p.n_strips = N * p.tiles_w;
  p.out = y;
  int rc = stem_view(&p.x_map, xs, N, Ho, Wo);
  if (rc) return rc;
  rc = make_tmap_2d(&p.w_map, w_s2d, 256, 64, 512, 64, 64);
  if (rc) return rc;
  rc = make_tmap_4d(&p.dy_map, y, 64, Wo, Ho, N, 128, (uint64_t)Wo * 128, (uint64_t)Ho * Wo * 128, 64, kSrTile, 1);
  if (rc) return rc;
  static bool configured[kMaxDevices] = {};  // the shared-memory limit of a kernel is a per-device attribute
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(stem_fwd_ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    StemFwdSmem::kBytes));
    ECGMM_CUDA(cudaFuncSetAttribute(stem_fwd_ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    StemFwdSmem::kBytes));
    configured[ds] = true;
  }
  const int grid = p.n_strips < num_sms() ? p.n_strips : num_sms();
  if (psum && psq)
    stem_fwd_ring_kernel<true><<<grid, kSrFwdThreads, StemFwdSmem::kBytes, st>>>(p);
  else
    stem_fwd_ring_kernel<false><<<grid, kSrFwdThreads, StemFwdSmem::kBytes, st>>>(p);
  return check_launch("stem_fwd_ring_kernel");
}

// dw[cout][c][r7][s7] += sum over CTAs (in CTA order) of the partial tiles written by stem_wgrad_ring_kernel.
__global__ void __launch_bounds__(256) stem_wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                                int n_parts) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // = cout * 147 + (c * 7 + r7) * 7 + s7
  if (idx >= 64 * 147) return;
  const int cout = idx / 147, rest = idx - cout * 147;
  const int c = rest / 49, r7 = (rest % 49) / 7, s7 = rest % 7;
  const int ra = r7 >> 1, dr = r7 & 1, sa = s7 >> 1, dsx = s7 & 1;
  const int ch = (dr * 2 + dsx) * 3 + c;
  const int pair = ra >> 1, m_row = (ra & 1) * 64 + sa * 16 + ch;
  const float* src = ws + ((size_t)pair * 128 + m_row) * 64 + cout;
  const size_t stride = (size_t)2 * 128 * 64;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int k = 0;
  for (; k + 3 < n_parts; k += 4) {
    a0 += src[(size_t)k * stride];
    a1 += src[(size_t)(k + 1) * stride];
    a2 += src[(size_t)(k + 2) * stride];
    a3 += src[(size_t)(k + 3) * stride];
  }
  for (; k < n_parts; ++k) a0 += src[(size_t)k * stride];
  dw[idx] += (a0 + a1) + (a2 + a3);
}

static int stem_wgrad_grid(int N, int W) {
  const int Wo = (W - 1) / 2 + 1;
  const int strips = N * ceil_div(Wo, kSrTile);
  return strips < num_sms() ? strips : num_sms();
}
size_t stem_wgrad_ring_workspace_bytes(int N, int H, int W) {
  (void)H;
  return (size_t)stem_wgrad_grid(N, W) * 2 * 128 * 64 * sizeof(float);
}

int launch_stem_wgrad_ring(const void* xs, const void* dy, float* dw, int N, int H, int W, void* workspace,
                           size_t ws_bytes, cudaStream_t st) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  StemRingParams p;
  memset(&p, 0, sizeof(p));
  p.ws = (workspace && ws_bytes >= stem_wgrad_ring_workspace_bytes(N, H, W)) ? static_cast<float*>(workspace) : nullptr;
  p.Ho = Ho;
  p.Wo = Wo;
  p.tiles_w = ceil_div(Wo, kSrTile);
  p.n_strips = N * p.tiles_w;
  p.dw = dw;
  int rc = stem_view(&p.x_map, xs, N, Ho, Wo);
  if (rc) return rc;
  rc = make_tmap_4d(&p.dy_map, dy, 64, Wo, Ho, N, 128, (uint64_t)Wo * 128, (uint64_t)Ho * Wo * 128, 64, kSrTile, 1);
  if (rc) return rc;
  p.w_map = p.x_map;
  static bool configured[kMaxDevices] = {};  // the shared-memory limit of a kernel is a per-device attribute
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(stem_wgrad_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    StemWgSmem::kBytes));
    configured[ds] = true;
  }
  const int grid = stem_wgrad_grid(N, W);
  stem_wgrad_ring_kernel<<<grid, 192, StemWgSmem::kBytes, st>>>(p);
  rc = check_launch("stem_wgrad_ring_kernel");
  if (rc || !p.ws) return rc;
  stem_wgrad_reduce_kernel<<<ceil_div(64 * 147, 256), 256, 0, st>>>(p.ws, dw, grid);
  return check_launch("stem_wgrad_reduce_kernel");
}

}  // namespace ecgmm
