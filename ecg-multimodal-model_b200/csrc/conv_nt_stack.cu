// Forward / data-gradient of the 64 -> 64 channel 3x3 stride-1 convolutions (ResNet18 layer1) with ROLLING
// ACCUMULATORS and N = 192 MMAs.
//
// Why: an M128 x N64 x K16 tcgen05.mma reads 6 KB of shared-memory operands per 32 MMA-clocks and sustains ~62 % of the
// tensor rate (profiles/r01_mma_n64_vs_n128.txt); N >= 128 runs at the full rate.  The GEMM N of these layers is only
// 64 output channels -- but one INPUT row feeds three OUTPUT rows.  For input row i and horizontal shift s the stacked
// B operand
//        [ W(r=2, s) ; W(r=1, s) ; W(r=0, s) ]          (three consecutive 64-row weight tiles, K-major, N = 192)
// yields, in ONE MMA group, the contributions of that input row to the output rows i-1, i, i+1, which live as adjacent
// 64-column blocks of a ring of output-row accumulators in TMEM (7 blocks = 448 columns).  A CTA walks DOWN a
// 128-pixel column strip: 12 MMAs of N = 192 (= 1152 MMA-clocks, the same as 36 MMAs of N = 64) per input row, every
// output row complete after its third input row, read out (64 columns, as in igemm_nt_halo_kernel), zeroed with
// tcgen05.st and handed back, so that every MMA accumulates.  At the first / last rows of a unit, and where the three
// blocks wrap around the ring, narrower (N = 64 / 128) groups are issued.
//
// Roles: warp 0 = TMA producer (input rows, one box each, used once), warp 1 = MMA issuer, warps 2..9 = epilogue: TWO
// warps per TMEM lane quadrant, 32 of the 64 columns each -- with one warp per quadrant the epilogue (TMEM read, bf16
// pack, swizzled staging store, accumulator zeroing, + optional statistics) was the longest per-tile chain of the
// kernel as soon as anything was added to it.
// Shared memory: 9 weight tiles (72 KB, resident) + ring of 6 input rows (102 KB) + 3 output staging tiles (48 KB).
#include "common.h"
#include "ptx.cuh"

#include <stdlib.h>
#include <string.h>

namespace ecgmm {

constexpr int kStTile = 128;                               // output pixels per tile
constexpr int kStBoxW = kStTile + 2;                       // staged pixels per input row
constexpr int kStBoxBytes = kStBoxW * 128;                 // 16640
constexpr int kStBoxStride = (kStBoxBytes + 1023) & ~1023;  // 17408
constexpr int kStRing = 6;                                 // input-row slots
constexpr int kStWTile = 64 * 128;                         // one tap of weights: [64 n][64 k] bf16
constexpr int kStBlocks = 7;                               // output-row accumulators in TMEM (64 columns each)

struct alignas(64) NtStackParams {
  CUtensorMap x_map;  // [N][H][W][64], box (64, 130, 1, 1)
  CUtensorMap w_map;  // [64][9*64] (k contiguous), box (64, 64)
  CUtensorMap y_map;  // output [N][H][W][64], box (64, 128, 1, 1)
  int8_t tap_of_q[9];  // smem weight tile q = shift*3 + k  (k: output row i-1+k)  <-  filter tap index r*3+s
  int tiles_w, H, W, n_img;
  int seg, segs_h, total_units;  // unit = (image, column strip, segment of `seg` output rows)
  int accumulate;
  float* psum;  // STATS 1: per-CTA BatchNorm partial sums of the stored (bf16-rounded) output, [gridDim.x][64]
  float* psq;   //          ... and sums of squares.   STATS 2: sum dz and sum dz * xhat (see below)
  // STATS 2 (data gradient): the output dx is the upstream gradient of a BatchNorm (+ReLU) whose input was red_x
  const __nv_bfloat16* red_x;  // [N][H][W][64] raw convolution output that BatchNorm normalised
  const uint8_t* red_mask;     // ReLU decisions, one byte per 8 channels (NULL: no ReLU)
  const float* red_mean;       // [64] batch mean / inverse standard deviation of red_x
  const float* red_invstd;
};

struct NtStackSmem {
  static constexpr int kW = 9 * kStWTile;                          // 73728
  static constexpr int kRingBytes = kStRing * kStBoxStride;        // 104448
  static constexpr int kOut = kW + kRingBytes;                     // 178176
  static constexpr int kBarOff = kOut + 3 * kStTile * 128;         // 227328
  static constexpr int kBytes = kBarOff + 512 + 1024;
};

// 32 lanes x 32 consecutive fp32 columns <- 32 registers per thread (mirror of tmem_ld_32x32)
__device__ __forceinline__ void tmem_st_zero_32x32(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// STATS = 1: the epilogue also accumulates the per-channel sum and sum of squares of everything this CTA stores.  An
// epilogue thread owns one pixel slot of every tile and all 64 channels of it, so the sums live in 128 registers per
// thread for the whole kernel (no shuffles, no shared-memory traffic per tile -- the two variants that were measured
// slower than a separate statistics pass in round 1) and are folded across the 128 threads ONCE, at the end.
// STATS = 2 (data gradient): the tile being stored is the gradient dy of a BatchNorm(+ReLU) output; the thread also
// loads its pixel's row of that BatchNorm's INPUT x (128 contiguous bytes) and ReLU bits (8 bytes) and accumulates the
// two sums the BatchNorm backward needs, sum dz and sum dz * x with dz = dy * relu'(.), so that the separate reduction
// pass over (x, dy) -- ecgmm_bn_bwd_reduce -- disappears.  Written out as sum dz and sum dz * xhat.
constexpr int kStThreads = 320;   // 10 warps
constexpr int kStEpiThreads = 256;

template <int STATS>
__global__ void __launch_bounds__(kStThreads, 1) igemm_nt_stack_kernel(const __grid_constant__ NtStackParams p) {
  using L = NtStackSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sA = smem + L::kW;
  uint8_t* sOut = smem + L::kOut;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);  // [ring] input row landed
  uint64_t* empty = full + kStRing;                                 // [ring] input row consumed by the MMAs
  uint64_t* ofull = empty + kStRing;                                // [blocks] output row complete in TMEM
  uint64_t* bfree = ofull + kStBlocks;                              // [blocks] accumulator read out and zeroed
  uint64_t* wfull = bfree + kStBlocks;
  uint64_t* oldfull = wfull + 1;                                    // [3] old output tile landed (accumulate mode)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(oldfull + 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.x_map);
    tma_prefetch_desc(&p.w_map);
    tma_prefetch_desc(&p.y_map);
    for (int i = 0; i < kStRing; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kStBlocks; ++i) {
      mbar_init(&ofull[i], 1);
      mbar_init(&bfree[i], 8);  // one arrival per epilogue warp
    }
    mbar_init(wfull, 1);
    for (int i = 0; i < 3; ++i) mbar_init(&oldfull[i], 1);
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

  auto unit_rows = [&](int u) {  // number of output rows of unit u
    const int sg = (u / p.tiles_w) % p.segs_h;
    return min(p.seg, p.H - sg * p.seg);
  };

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: weights once, then input rows
    if (elect_one()) {
      mbar_expect_tx(wfull, 9 * kStWTile);
      for (int q = 0; q < 9; ++q) tma_load_2d(sW + q * kStWTile, &p.w_map, wfull, p.tap_of_q[q] * 64, 0);
      uint32_t n = 0;  // running input-row counter: slot = n % ring, fill parity = (n / ring) & 1
      for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
        const int twi = u % p.tiles_w;
        const int sg = (u / p.tiles_w) % p.segs_h;
        const int img = u / (p.tiles_w * p.segs_h);
        const int oh0 = sg * p.seg;
        const int rows = unit_rows(u);
        const int w0 = twi * kStTile;
        for (int j = -1; j <= rows; ++j, ++n) {  // input rows oh0-1 .. oh0+rows (outside the image: zero-filled)
          const uint32_t slot = n % kStRing, par = (n / kStRing) & 1u;
          mbar_wait(&empty[slot], par ^ 1u);
          mbar_expect_tx(&full[slot], kStBoxBytes);
          tma_load_4d(sA + slot * kStBoxStride, &p.x_map, &full[slot], 0, w0 - 1, oh0 + j, img);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      const uint64_t w_desc0 = make_sw128_desc(smem_u32(sW), 0, 1024);
      const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA), 0, 1024);
      mbar_wait(wfull, 0);
      tc_fence_after();
      uint32_t n = 0;   // running input-row counter (as in the producer)
      uint32_t g0 = 0;  // virtual index of the unit's first output row; accumulator block = g % kStBlocks
      for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
        const int rows = unit_rows(u);
        for (int j = -1; j <= rows; ++j, ++n) {
          // this input row feeds output rows j-1+k, k = 0..2, restricted to [0, rows)
          const int k0 = max(0, 1 - j), k1 = min(2, rows - j);  // inclusive range of k
          const uint32_t slot = n % kStRing;
          mbar_wait(&full[slot], (n / kStRing) & 1u);
          if (j + 1 < rows) {  // first touch of output row j+1: its accumulator must have been handed back (zeroed)
            const uint32_t g = g0 + (uint32_t)(j + 1);
            mbar_wait(&bfree[g % kStBlocks], (g / kStBlocks) & 1u);
          }
          tc_fence_after();
          if (k0 <= k1) {
            const uint32_t gfirst = g0 + (uint32_t)(j - 1 + k0);
            const int cnt = k1 - k0 + 1;
            const int b0 = (int)(gfirst % kStBlocks);
            const int cnt1 = min(cnt, kStBlocks - b0);  // blocks before the ring wraps
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const uint64_t a_desc = a_desc0 + (uint64_t)((slot * kStBoxStride + s * 128) >> 4);
              for (int part = 0; part < 2; ++part) {
                const int c = part == 0 ? cnt1 : cnt - cnt1;
                if (c <= 0) continue;
                const int kq = part == 0 ? k0 : k0 + cnt1;
                const int blk = part == 0 ? b0 : 0;
                const uint64_t w_desc = w_desc0 + (uint64_t)((s * 3 + kq) * (kStWTile >> 4));
                const uint32_t idesc = make_idesc_bf16(128, 64 * c, 0, 0);
                const uint32_t d_tmem = tmem_base + blk * 64;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_desc + 2 * k, w_desc + 2 * k, idesc, 1u);
              }
            }
          }
          umma_commit(&empty[slot]);  // an input row is read by exactly this group of MMAs
          if (j >= 1) {               // output row j-1 has now received all three input rows
            const uint32_t g = g0 + (uint32_t)(j - 1);
            umma_commit(&ofull[g % kStBlocks]);
          }
        }
        g0 += (uint32_t)rows;
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 2..9, 256 threads)
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;   // which 32 of the 64 columns (channels) this warp converts
    const int m_row = quad * 32 + lane;
    const int et = (int)threadIdx.x - 64;  // 0..255
    const bool leader = (threadIdx.x == 64);
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + half * 32;
    // hand over all accumulators zeroed (completion #0 of every bfree barrier)
    for (int b = 0; b < kStBlocks; ++b) tmem_st_zero_32x32(lane_base + b * 64);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int b = 0; b < kStBlocks; ++b) mbar_arrive(&bfree[b]);

    auto tile_coords = [&](int u, int i, int& w0, int& oh, int& img) {
      const int twi = u % p.tiles_w;
      const int sg = (u / p.tiles_w) % p.segs_h;
      img = u / (p.tiles_w * p.segs_h);
      w0 = twi * kStTile;
      oh = sg * p.seg + i;
    };
    if (leader && p.accumulate && (int)blockIdx.x < p.total_units) {
      int w0, oh, img;
      tile_coords(blockIdx.x, 0, w0, oh, img);
      mbar_expect_tx(&oldfull[0], kStTile * 128);
      tma_load_4d(sOut, &p.y_map, &oldfull[0], 0, w0, oh, img);
    }
    float ssum[STATS ? 32 : 1], ssq[STATS ? 32 : 1];  // this thread's pixel slot x its 32 channels, whole kernel
    if constexpr (STATS != 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) ssum[i] = ssq[i] = 0.f;
    }
    uint32_t g = 0;  // virtual output-row index (as in the MMA issuer); also the staging-buffer counter
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
      const int rows = unit_rows(u);
      for (int i = 0; i < rows; ++i, ++g) {
        int w0, oh, img;
        tile_coords(u, i, w0, oh, img);
        const bool in_image = w0 + m_row < p.W;  // pixel slots right of the image edge are computed but never stored
        const int blk = (int)(g % kStBlocks);
        const int ob = (int)(g % 3);
        uint8_t* buf = sOut + ob * (kStTile * 128);
        uint4 xrow[STATS == 2 ? 4 : 1];
        uint32_t mbits = 0xffffffffu;
        if constexpr (STATS == 2) {  // issued before the accumulator wait: the latency hides behind the MMAs
          if (in_image) {
            const size_t pix = ((size_t)img * p.H + oh) * p.W + (w0 + m_row);
            const uint4* xs = reinterpret_cast<const uint4*>(p.red_x + pix * 64 + half * 32);
#pragma unroll
            for (int k = 0; k < 4; ++k) xrow[k] = __ldg(xs + k);
            if (p.red_mask) mbits = __ldg(reinterpret_cast<const uint32_t*>(p.red_mask + pix * 8 + half * 4));
          }
        }
        if (leader) {
          tma_store_wait_read<1>();  // every store but the newest has been read out of its buffer
          if (p.accumulate) {
            int nu = u, ni = i + 1;
            if (ni == rows) {
              nu = u + gridDim.x;
              ni = 0;
            }
            if (nu < p.total_units) {
              int nw0, noh, nimg;
              tile_coords(nu, ni, nw0, noh, nimg);
              const int nb = (int)((g + 1) % 3);
              mbar_expect_tx(&oldfull[nb], kStTile * 128);
              tma_load_4d(sOut + nb * (kStTile * 128), &p.y_map, &oldfull[nb], 0, nw0, noh, nimg);
            }
          }
        }
        named_bar_sync(1, kStEpiThreads);
        mbar_wait(&ofull[blk], (g / kStBlocks) & 1u);
        tc_fence_after();
        if (p.accumulate) mbar_wait(&oldfull[ob], (g / 3) & 1u);
        const uint32_t t_addr = lane_base + blk * 64;
        uint8_t* row = buf + m_row * 128;
        {
          const int c = half;
          uint32_t r[32];
          tmem_ld_32x32(t_addr, r);
          tmem_ld_wait();
          tmem_st_zero_32x32(t_addr);  // hand the accumulator back empty
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            // 16-byte chunk j of row m lives at chunk (j ^ (m & 7)) of the swizzled tile
            uint4* d4 = reinterpret_cast<uint4*>(row + (((c * 4 + q) ^ (m_row & 7)) << 4));
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[q * 8 + j]);
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
            uint4 v;
            __nv_bfloat162* vb = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) vb[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            *d4 = v;
            if constexpr (STATS == 1) {
              if (in_image) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 s2 = __bfloat1622float2(vb[j]);  // statistics of what is stored, not of the fp32 value
                  const int ch = q * 8 + 2 * j;
                  ssum[ch] += s2.x;
                  ssq[ch] = fmaf(s2.x, s2.x, ssq[ch]);
                  ssum[ch + 1] += s2.y;
                  ssq[ch + 1] = fmaf(s2.y, s2.y, ssq[ch + 1]);
                }
              }
            }
            if constexpr (STATS == 2) {
              if (in_image) {
                const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(&xrow[q]);
                const uint32_t bits = mbits >> (8 * q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float2 d = __bfloat1622float2(vb[j]);  // the gradient as the BatchNorm backward will read it
                  const float2 xx = __bfloat1622float2(xb[j]);
                  if (!((bits >> (2 * j)) & 1u)) d.x = 0.f;
                  if (!((bits >> (2 * j + 1)) & 1u)) d.y = 0.f;
                  const int ch = q * 8 + 2 * j;
                  ssum[ch] += d.x;
                  ssq[ch] = fmaf(d.x, xx.x, ssq[ch]);
                  ssum[ch + 1] += d.y;
                  ssq[ch + 1] = fmaf(d.y, xx.y, ssq[ch + 1]);
                }
              }
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bfree[blk]);  // the MMA issuer may start the next output row in this block
        fence_proxy_async_smem();                  // make the staging tile visible to the TMA engine
        named_bar_sync(1, kStEpiThreads);
        if (leader) {
          tma_store_4d(&p.y_map, buf, 0, w0, oh, img);
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_all<0>();
    if constexpr (STATS != 0) {
      // fold the 128 threads' sums: the input-row ring is idle now (every MMA that read it has completed, or the last
      // ofull wait above would not have returned); rows of 129 floats keep the column reads conflict-free
      float* scr = reinterpret_cast<float*>(sA);
      named_bar_sync(1, kStEpiThreads);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        scr[m_row * 129 + half * 32 + i] = ssum[i];
        scr[m_row * 129 + 64 + half * 32 + i] = ssq[i];
      }
      named_bar_sync(1, kStEpiThreads);
      if constexpr (STATS == 1) {
        if (et < 128) {  // column et of the [128 pixel slots][sum 0..63 | sq 64..127] table
          double acc = 0.0;
          for (int r = 0; r < 128; ++r) acc += (double)scr[r * 129 + et];
          float* dst = (et < 64) ? p.psum : p.psq;
          dst[(size_t)blockIdx.x * 64 + (et & 63)] = (float)acc;
        }
      } else if (et < 64) {
        double s1 = 0.0, s2 = 0.0;
        for (int r = 0; r < 128; ++r) {
          s1 += (double)scr[r * 129 + et];
          s2 += (double)scr[r * 129 + 64 + et];
        }
        // sum dz * xhat = invstd * (sum dz * x - mean * sum dz), in double
        p.psum[(size_t)blockIdx.x * 64 + et] = (float)s1;
        p.psq[(size_t)blockIdx.x * 64 + et] = (float)((double)p.red_invstd[et] * (s2 - (double)p.red_mean[et] * s1));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

bool nt_stack_supported(int Cin, int Cout, int R, int S, int stride, int W) {
  return stride == 1 && Cin == 64 && Cout == 64 && S == 3 && R == 3 && W >= 96;
}

// Work decomposition (shared with the statistics-row query): unit = (image, 128-pixel column strip, segment of rows).
static void nt_stack_units(NtStackParams& p, int N, int H, int W) {
  p.tiles_w = ceil_div(W, kStTile);
  p.H = H;
  p.W = W;
  p.n_img = N;
  int seg = H < 16 ? H : 16;
  while (seg > 2 && (long long)N * p.tiles_w * ceil_div(H, seg) < 4LL * num_sms()) seg = (seg + 1) / 2;
  seg = ceil_div(H, ceil_div(H, seg));
  p.seg = seg;
  p.segs_h = ceil_div(H, seg);
  p.total_units = N * p.tiles_w * p.segs_h;
}

// rows of the [rows][64] statistics partials a forward launch with psum/psq writes (= its grid size)
int nt_stack_stats_rows(int N, int H, int W) {
  NtStackParams p;
  memset(&p, 0, sizeof(p));
  nt_stack_units(p, N, H, W);
  return p.total_units < num_sms() ? p.total_units : num_sms();
}

// dgrad != 0: w is the [Cin][R][S][Cout] shadow and taps are mirrored (dx[h,w] += dy[h+1-r, w+1-s] W[r,s]).
// psum / psq (may be NULL): nt_stack_stats_rows() rows of 64 floats each -- forward: BatchNorm partial sums of y;
// dgrad with red != NULL: the BatchNorm-backward sums of (red->x, y) (struct NtStackReduce in the caller's terms).
int launch_nt_stack(const __nv_bfloat16* x, const __nv_bfloat16* w, __nv_bfloat16* y, int N, int H, int W, int dgrad,
                    int accumulate, cudaStream_t st, float* psum, float* psq, const __nv_bfloat16* red_x,
                    const uint8_t* red_mask, const float* red_mean, const float* red_invstd) {
  NtStackParams p;
  memset(&p, 0, sizeof(p));
  p.psum = psum;
  p.psq = psq;
  p.red_x = red_x;
  p.red_mask = red_mask;
  p.red_mean = red_mean;
  p.red_invstd = red_invstd;
  // weight tile q = shift*3 + k feeds output row i-1+k from input row i read `shift` pixels to the right of w0-1.
  //   forward : out[oh][w] = sum x[oh+r-1][w+s-1] W[r][s]   ->  input row i = oh+r-1: k = 2-r ... r = 2-k, s = shift
  //   dgrad   : dx[h][w]   = sum dy[h+1-r][w+1-s] Wt[r][s]  ->  input row i = h+1-r : k = r,          s = 2-shift
  for (int sh = 0; sh < 3; ++sh)
    for (int k = 0; k < 3; ++k) p.tap_of_q[sh * 3 + k] = (int8_t)(dgrad ? (k * 3 + (2 - sh)) : ((2 - k) * 3 + sh));
  nt_stack_units(p, N, H, W);
  p.accumulate = accumulate;
  const uint64_t e = 2;
  int rc = make_tmap_4d(&p.x_map, x, 64, W, H, N, 64 * e, (uint64_t)W * 64 * e, (uint64_t)H * W * 64 * e, 64, kStBoxW, 1);
  if (rc) return rc;
  rc = make_tmap_2d(&p.w_map, w, (uint64_t)9 * 64, 64, (uint64_t)9 * 64 * e, 64, 64);
  if (rc) return rc;
  rc = make_tmap_4d(&p.y_map, y, 64, W, H, N, 64 * e, (uint64_t)W * 64 * e, (uint64_t)H * W * 64 * e, 64, kStTile, 1);
  if (rc) return rc;
  static bool configured[kMaxDevices] = {};  // the shared-memory limit of a kernel is a per-device attribute
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(igemm_nt_stack_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    NtStackSmem::kBytes));
    ECGMM_CUDA(cudaFuncSetAttribute(igemm_nt_stack_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    NtStackSmem::kBytes));
    ECGMM_CUDA(cudaFuncSetAttribute(igemm_nt_stack_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    NtStackSmem::kBytes));
    configured[ds] = true;
  }
  const int grid = p.total_units < num_sms() ? p.total_units : num_sms();
  if (psum && psq && !dgrad)
    igemm_nt_stack_kernel<1><<<grid, kStThreads, NtStackSmem::kBytes, st>>>(p);
  else if (psum && psq && dgrad && red_x && red_mean && red_invstd)
    igemm_nt_stack_kernel<2><<<grid, kStThreads, NtStackSmem::kBytes, st>>>(p);
  else
    igemm_nt_stack_kernel<0><<<grid, kStThreads, NtStackSmem::kBytes, st>>>(p);
  return check_launch("igemm_nt_stack_kernel");
}

}  // namespace ecgmm
