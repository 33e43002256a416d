// Forward / data-gradient of the 64->64 channel 3x3 (and 1x3) stride-1 convolutions with
// shared-memory halo reuse and CTA-resident weights (ResNet18 layer1: 4 convs x {fwd, dgrad};
// the 1-D layer1 block).
//
// igemm_nt_kernel stages one 16 KB activation box per filter tap: with N = 64 output channels a
// 128x64 tile moves 9 x (16 + 8) KB for 9.4 MFLOP, 44 flop/B, and the kernel sits on the
// L2->SM bandwidth cap (measured 38 B/clk/SM of ~42).  Here
//   * the 9 weight tiles [64 cout][64 cin] (72 KB) are loaded ONCE per persistent CTA;
//   * an M tile is 128 consecutive output pixels of one image row; the 3 input rows it needs are
//     staged once each as a box of 130 pixels and filter tap (r,s) is row r's box read from
//     pixel s onwards (descriptor start += s*128 B; legal because the 128-byte swizzle is a
//     function of the absolute shared-memory address, see tools/desc_probe.py);
// so a tile moves 50 KB for the same 9.4 MFLOP.
// The N = 64 MMA shape itself is the cap: issuing N = 128 MMAs over the same tiles (2x the tensor work, half of it
// discarded) made this kernel only 24 % slower, i.e. M128xN64xK16 sustains ~62 % of the rate of M128xN128xK16
// (6 KB of shared-memory operands per 32 MMA-clocks); profiles/r01_mma_n64_vs_n128.txt.
// (A rolling-row variant -- a CTA walking down a column strip with a ring of input rows, one new row per tile --
// was measured too: 17 % faster with a cold L2 (tools/conv_bench.py) but 10 % SLOWER inside the training step,
// where the linear tile order below finds part of the producer's output still in L2; profiles/README.md.)
#include "common.h"
#include "ptx.cuh"
#include "vec.cuh"

#include <stdlib.h>

namespace ecgmm {

constexpr int kNhTile = 128;                               // output pixels per tile
constexpr int kNhBoxW = kNhTile + 2;                       // staged pixels per input row
constexpr int kNhBoxBytes = kNhBoxW * 128;                 // 16640
constexpr int kNhBoxStride = (kNhBoxBytes + 1023) & ~1023;  // 17408
constexpr int kNhStages = 2;
constexpr int kNhWTile = 64 * 128;                         // one tap of weights: [64 n][64 k] bf16
constexpr int kNhMaxTaps = 9;

struct alignas(64) NtHaloParams {
  CUtensorMap x_map;  // [N][H][W][64], box (64, 130, 1, 1)
  CUtensorMap w_map;  // [64][ntaps*64] (k contiguous), box (64, 64)
  CUtensorMap y_map;  // output [N][H][W][64], box (64, 128, 1, 1): epilogue TMA store (and load when accumulating)
  int ntaps, rows;    // rows = halo rows staged per tile (3 for 3x3, 1 for 1x3)
  int8_t tap_row[kNhMaxTaps], tap_shift[kNhMaxTaps];
  int padW, row0;     // input row of halo row 0 relative to the output row (-(R/2))
  int tiles_w, H, W, n_img, total_tiles;
  __nv_bfloat16* out;  // [N][H][W][64]
  int accumulate;
  // igemm_nt_halo_kernel<true> only (inference): the epilogue applies the folded
  // BatchNorm y = acc * scale[c] + shift[c], adds the residual tile fetched through res_map when accumulate != 0,
  // then ReLU when relu != 0.  Appended so that the offsets the training kernel reads do not move.
  const float* scale;
  const float* shift;
  int relu;
  CUtensorMap res_map;  // residual [N][H][W][64], box (64, 128, 1, 1)
};

struct NtHaloSmem {
  static constexpr int kW = kNhMaxTaps * kNhWTile;                 // 73728
  static constexpr int kStage = 3 * kNhBoxStride;                  // 52224
  static constexpr int kOut = kW + kNhStages * kStage;             // 178176: 3 output staging tiles of 16 KiB
  static constexpr int kBarOff = kOut + 3 * kNhTile * 128;         // 227328
  static constexpr int kBytes = kBarOff + 256 + 1024;
  static constexpr int kCoefOff = kBarOff + 256;          // fused epilogue: scale[64], shift[64] fp32
  static constexpr int kBytesFused = kBytes + 512;
};

// FUSED = false: forward / data-gradient as used by training.  FUSED = true: forward with the folded-BatchNorm
// (+ residual, + ReLU) epilogue of the serving path; not selected unless the caller asks for it (ecgmm_conv2d_fwd_bn).
template <bool FUSED>
__global__ void __launch_bounds__(192, 1) igemm_nt_halo_kernel(const __grid_constant__ NtHaloParams p) {
  using L = NtHaloSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sA = smem + L::kW;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + kNhStages;
  uint64_t* tfull = empty + kNhStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint64_t* ofull = wfull + 1;  // [3] old output tile landed in the staging buffer (accumulate mode)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ofull + 3);
  uint8_t* sOut = smem + L::kOut;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.x_map);
    tma_prefetch_desc(&p.w_map);
    for (int i = 0; i < kNhStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    mbar_init(wfull, 1);
    for (int i = 0; i < 3; ++i) mbar_init(&ofull[i], 1);
    tma_prefetch_desc(&p.y_map);
    mbar_fence_init();
  }
  if constexpr (FUSED) {
    float* coef = reinterpret_cast<float*>(smem + L::kCoefOff);
    if (threadIdx.x >= 64) {
      const int i = threadIdx.x - 64;  // 128 epilogue threads: scale[0..64) then shift[0..64)
      coef[i] = i < 64 ? p.scale[i] : p.shift[i - 64];
    }
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
      // weights: once per CTA
      mbar_expect_tx(wfull, p.ntaps * kNhWTile);
      for (int t = 0; t < p.ntaps; ++t) tma_load_2d(sW + t * kNhWTile, &p.w_map, wfull, t * 64, 0);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = p.rows * kNhBoxBytes;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int twi = t % p.tiles_w;
        const int m = t / p.tiles_w;
        const int oh = m % p.H;
        const int img = m / p.H;
        const int w0 = twi * kNhTile;
        uint8_t* st = sA + stage * L::kStage;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], tx);
        for (int r = 0; r < p.rows; ++r)
          tma_load_4d(st + r * kNhBoxStride, &p.x_map, &full[stage], 0, w0 - p.padW, oh + p.row0 + r, img);
        if (++stage == kNhStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      const uint64_t w_desc0 = make_sw128_desc(smem_u32(sW), 0, 1024);
      const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA), 0, 1024);
      // tap -> descriptor offset (in 16-byte units) inside a stage
      uint32_t a_off[kNhMaxTaps];
#pragma unroll
      for (int t = 0; t < kNhMaxTaps; ++t)
        a_off[t] = (t < p.ntaps) ? ((p.tap_row[t] * kNhBoxStride + p.tap_shift[t] * 128) >> 4) : 0;
      mbar_wait(wfull, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 64;
        const uint64_t a_st = a_desc0 + (uint64_t)(stage * (L::kStage >> 4));
#pragma unroll
        for (int tap = 0; tap < kNhMaxTaps; ++tap) {
          if (tap < p.ntaps) {
            const uint64_t a_desc = a_st + a_off[tap];
            const uint64_t w_desc = w_desc0 + (uint64_t)(tap * (kNhWTile >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_desc + 2 * k, w_desc + 2 * k, idesc, (tap | k) != 0);
          }
        }
        umma_commit(&empty[stage]);
        umma_commit(&tfull[acc]);
        if (++stage == kNhStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5, 128 threads)
    // TMEM -> registers -> bf16 -> SWIZZLE_128B staging tile in shared memory -> ONE TMA tensor store per
    // tile (full 128-byte lines; pixels past the image edge are clipped by the tensor map).  Writing
    // 16 bytes per lane straight to global memory touches 32 different lines per instruction and made
    // the epilogue, not the MMA, the pacing stage (ncu: long-scoreboard stalls behind STG).
    const int quad = warp & 3;
    const int m_row = quad * 32 + lane;
    const bool leader = (threadIdx.x == 64);
    // Staging tiles rotate over 3 buffers (tile `it` uses buffer it % 3).  In accumulate mode the OLD contents of
    // tile it+1 are fetched (TMA load) while tile `it` is being converted, into the buffer whose last store
    // (tile it-2) has been read out: the load latency used to sit in front of every tile (280 us against 181 us
    // without accumulation at batch 64).
    auto tile_coords = [&](int t, int& w0, int& oh, int& img) {
      const int twi = t % p.tiles_w;
      const int m = t / p.tiles_w;  // img * H + oh
      oh = m % p.H;
      img = m / p.H;
      w0 = twi * kNhTile;
    };
    if (leader && p.accumulate && (int)blockIdx.x < p.total_tiles) {
      int w0, oh, img;
      tile_coords(blockIdx.x, w0, oh, img);
      mbar_expect_tx(&ofull[0], kNhTile * 128);
      tma_load_4d(sOut, FUSED ? &p.res_map : &p.y_map, &ofull[0], 0, w0, oh, img);
    }
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      int w0, oh, img;
      tile_coords(t, w0, oh, img);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int ob = it % 3;
      const uint32_t ob_phase = (it / 3) & 1;
      uint8_t* buf = sOut + ob * (kNhTile * 128);
      if (leader) {
        tma_store_wait_read<1>();  // every store but the newest (tile it-1) has been read out of its buffer
        const int tn = t + gridDim.x;
        if (p.accumulate && tn < p.total_tiles) {
          int nw0, noh, nimg;
          tile_coords(tn, nw0, noh, nimg);
          const int nb = (it + 1) % 3;
          mbar_expect_tx(&ofull[nb], kNhTile * 128);
          tma_load_4d(sOut + nb * (kNhTile * 128), FUSED ? &p.res_map : &p.y_map, &ofull[nb], 0, nw0, noh, nimg);
        }
      }
      named_bar_sync(1, 128);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (p.accumulate) mbar_wait(&ofull[ob], ob_phase);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 64;
      uint8_t* row = buf + m_row * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          // 16-byte chunk j of row m lives at chunk (j ^ (m & 7)) of the swizzled tile
          uint4* d4 = reinterpret_cast<uint4*>(row + (((c * 4 + q) ^ (m_row & 7)) << 4));
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[q * 8 + j]);
          if constexpr (FUSED) {  // broadcast reads: every thread of the warp asks for the same 16 bytes
            const float4* cf = reinterpret_cast<const float4*>(smem + L::kCoefOff) + (c * 8 + q * 2);
            const float4 s0 = cf[0], s1 = cf[1], h0 = cf[16], h1 = cf[17];
            f[0] = fmaf(f[0], s0.x, h0.x); f[1] = fmaf(f[1], s0.y, h0.y);
            f[2] = fmaf(f[2], s0.z, h0.z); f[3] = fmaf(f[3], s0.w, h0.w);
            f[4] = fmaf(f[4], s1.x, h1.x); f[5] = fmaf(f[5], s1.y, h1.y);
            f[6] = fmaf(f[6], s1.z, h1.z); f[7] = fmaf(f[7], s1.w, h1.w);
          }
          if (p.accumulate) {
            const uint4 old = *d4;
            const __nv_bfloat162* oldb = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 o = __bfloat1622float2(oldb[j]);
              f[2 * j] += o.x;
              f[2 * j + 1] += o.y;
            }
          }
          if constexpr (FUSED) {
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
          }
          uint4 v;
          __nv_bfloat162* vb = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
          for (int j = 0; j < 4; ++j) vb[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          *d4 = v;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);  // accumulator may be overwritten by the MMA warp
      fence_proxy_async_smem();                   // make the staging tile visible to the TMA engine
      named_bar_sync(1, 128);
      if (leader) {
        tma_store_4d(&p.y_map, buf, 0, w0, oh, img);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_all<0>();

  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

bool nt_halo_supported(int Cin, int Cout, int R, int S, int stride, int W) {
  return stride == 1 && Cin == 64 && Cout == 64 && S == 3 && (R == 1 || R == 3) && W >= 96;
}

// dgrad != 0: w is the [Cin][R][S][Cout] shadow and taps are mirrored (dx[h,w] += dy[h+pad-r, w+pad-s] W[r,s]).
// scale != NULL selects the fused inference epilogue (forward only): y = act(conv * scale + shift [+ res]).
int launch_nt_halo(const __nv_bfloat16* x, const __nv_bfloat16* w, __nv_bfloat16* y, int N, int H, int W, int R, int S,
                   int dgrad, int accumulate, cudaStream_t st, const float* scale, const float* shift,
                   const __nv_bfloat16* res, int relu) {
  NtHaloParams p;
  memset(&p, 0, sizeof(p));
  p.ntaps = R * S;
  p.rows = R;
  p.padW = S / 2;
  p.row0 = -(R / 2);
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      const int t = r * S + s;
      p.tap_row[t] = (int8_t)(dgrad ? (R - 1 - r) : r);
      p.tap_shift[t] = (int8_t)(dgrad ? (S - 1 - s) : s);
    }
  p.tiles_w = ceil_div(W, kNhTile);
  p.H = H;
  p.W = W;
  p.n_img = N;
  p.total_tiles = N * H * p.tiles_w;
  p.out = y;
  p.accumulate = accumulate;
  const bool fused = scale != nullptr;
  if (fused) {
    ECGMM_CHECK(shift && !dgrad && !accumulate, ECGMM_ERR_ARG, "nt_halo: the fused epilogue is forward-only");
    p.scale = scale;
    p.shift = shift;
    p.relu = relu;
    p.accumulate = res != nullptr;  // "old tile" = the residual
  }
  const uint64_t e = 2;
  int rc = make_tmap_4d(&p.x_map, x, 64, W, H, N, 64 * e, (uint64_t)W * 64 * e, (uint64_t)H * W * 64 * e, 64, kNhBoxW, 1);
  if (rc) return rc;
  rc = make_tmap_2d(&p.w_map, w, (uint64_t)R * S * 64, 64, (uint64_t)R * S * 64 * e, 64, 64);
  if (rc) return rc;
  rc = make_tmap_4d(&p.y_map, y, 64, W, H, N, 64 * e, (uint64_t)W * 64 * e, (uint64_t)H * W * 64 * e, 64, kNhTile, 1);
  if (rc) return rc;
  if (fused && res) {
    rc = make_tmap_4d(&p.res_map, res, 64, W, H, N, 64 * e, (uint64_t)W * 64 * e, (uint64_t)H * W * 64 * e, 64, kNhTile, 1);
    if (rc) return rc;
  } else {
    p.res_map = p.y_map;
  }
  static bool configured[kMaxDevices] = {};  // the shared-memory limit of a kernel is a per-device attribute
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(igemm_nt_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    NtHaloSmem::kBytes));
    ECGMM_CUDA(cudaFuncSetAttribute(igemm_nt_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    NtHaloSmem::kBytesFused));
    configured[ds] = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (fused)
    igemm_nt_halo_kernel<true><<<grid, 192, NtHaloSmem::kBytesFused, st>>>(p);
  else
    igemm_nt_halo_kernel<false><<<grid, 192, NtHaloSmem::kBytes, st>>>(p);
  return check_launch("igemm_nt_halo_kernel");
}

}  // namespace ecgmm
