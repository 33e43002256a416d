// Implicit-GEMM convolutions on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM, operands staged by TMA into SWIZZLE_128B shared memory).
//
//  igemm_nt_kernel : forward and data-gradient.   D[pixel][cout] = sum_k A[pixel][k] * B[cout][k]
//      A tile = 128 output pixels (TH x TW rectangle of one image) x 64 input channels, loaded by
//      ONE 4-D TMA box per filter tap from the NHWC activation (the tap is a coordinate offset;
//      out-of-bounds rows/columns are zero-filled by TMA, which implements the padding).
//      Stride-2 convolutions read through per-parity views of the input (base + parity offset,
//      doubled W/H strides) so that every tap is still a dense box.
//  igemm_tn_kernel : weight-gradient.   D[(tap,cin)][cout] = sum_pixels X[pixel+tap][cin] * dY[pixel][cout]
//      both operands are "MN-major" (the reduction index = pixel is the slow smem dimension).
//
// Both kernels are persistent (grid = #SMs), warp-specialised: warp 0 = TMA producer,
// warp 1 = MMA issuer (+ TMEM allocation), warps 2..5 = epilogue (one TMEM lane quadrant each).
#include "common.h"
#include "ptx.cuh"
#include "vec.cuh"

#include <stdlib.h>

namespace ecgmm {

struct Tap {
  int16_t map;  // which activation tensor map
  int16_t dh;   // row offset added to the tile origin
  int16_t dw;   // column offset
  int16_t id;   // filter tap index r*S+s (wgrad output addressing)
  int32_t wk;   // K offset of this tap inside the weight matrix
};

constexpr int kMaxTaps = 16;
constexpr int kATile = 128 * 128;  // 128 pixels x 64 bf16 = 16 KiB

struct alignas(64) NtParams {
  CUtensorMap a_maps[4];
  CUtensorMap b_map;
  Tap taps[kMaxTaps];
  int ntaps;
  int k_chunks;  // 64-wide K chunks per tap (= Cin/64 forward, Cout/64 dgrad)
  int TH, TW, tw_shift;
  int tiles_h, tiles_w, n_img, n_tiles_n, total_tiles;
  int OH, OW;  // valid extent of the tile grid
  __nv_bfloat16* out;
  long long out_sN, out_sH, out_sW;  // element strides of the output pixel grid
  int accumulate;
  // BatchNorm batch statistics of the output, produced by the epilogue (forward only; NULL = off):
  // psum/psq [gridDim.x * 4][stats_C]: per (CTA, epilogue warp) partial sum / sum of squares of every channel
  // over the VALID pixels of the tiles that warp converted (bf16-rounded values, i.e. of the stored tensor).
  float* psum;
  float* psq;
  int stats_C;
  // igemm_nt_kernel<BN, STAGES, true> only (inference): the epilogue stores
  // act(acc * scale[c] + shift[c] + res) -- folded BatchNorm, optional residual (same layout as out), optional ReLU.
  // Appended so that the offsets the training kernels read do not move.
  const float* scale;
  const float* shift;
  const __nv_bfloat16* res;
  int relu;
  // igemm_nt_pair_kernel (cta_group::2): the weight matrix with a box of HALF an N tile (each CTA of a pair loads and
  // feeds half of B), the number of 128-pixel tiles and of (tile pair, N tile) work items
  CUtensorMap b_map_half;
  int m_tiles, total_pairs;
  int pair_ok;  // b_map_half has been built
};

template <int BN, int STAGES>
struct NtSmem {
  static constexpr int kBTile = BN * 128;
  static constexpr int kStage = kATile + kBTile;
  static constexpr int kBarOff = STAGES * kStage;
  static constexpr int kStatsOff = kBarOff + 256;   // double [4 warps][2][BN]
  static constexpr int kBytes = kStatsOff + 4 * 2 * BN * 8 + 1024;  // barriers + statistics + alignment slack
};

template <int BN, int STAGES, bool FUSED = false>
__global__ void __launch_bounds__(192, 1) igemm_nt_kernel(const __grid_constant__ NtParams p) {
  using L = NtSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kATile;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_maps[i]);
    tma_prefetch_desc(&p.b_map);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nkb = p.ntaps * p.k_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // One elected thread runs the whole loop: under a plain `lane == 0` guard ptxas wraps every
    // uniform-datapath instruction (UTMALDG / UTCHMMA) in an ELECT..BRA.U.ANY serialisation loop.
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int nt = t % p.n_tiles_n;
        int m = t / p.n_tiles_n;
        const int twi = m % p.tiles_w;
        m /= p.tiles_w;
        const int thi = m % p.tiles_h;
        const int img = m / p.tiles_h;
        const int h0 = thi * p.TH, w0 = twi * p.TW;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const Tap tp = p.taps[tap];
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], L::kStage);
            tma_load_4d(sA + stage * kATile, &p.a_maps[tp.map], &full[stage], kc * 64, w0 + tp.dw, h0 + tp.dh, img);
            tma_load_2d(sB + stage * L::kBTile, &p.b_map, &full[stage], tp.wk + kc * 64, nt * BN);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single elected thread)
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA), 0, 1024);
      const uint64_t b_desc0 = make_sw128_desc(smem_u32(sB), 0, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + (uint64_t)(stage * (kATile >> 4));
          const uint64_t b_desc = b_desc0 + (uint64_t)(stage * (L::kBTile >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 4 x (K = 16) per 64-wide chunk; +32 B inside the swizzle atom
            umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);
          if (kb == nkb - 1) umma_commit(&tfull[acc]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int quad = warp & 3;
    const int m_row = quad * 32 + lane;
    const int hl = m_row >> p.tw_shift;
    const int wl = m_row & (p.TW - 1);
    // statistics: lane j of this warp owns columns c*32 + j of the CTA's (fixed) N tile
    double* st_s = reinterpret_cast<double*>(smem + L::kStatsOff) + quad * 2 * BN;
    double* st_q = st_s + BN;
    if (p.psum)
      for (int c = lane; c < BN; c += 32) st_s[c] = st_q[c] = 0.0;
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int nt = t % p.n_tiles_n;
      int m = t / p.n_tiles_n;
      const int twi = m % p.tiles_w;
      m /= p.tiles_w;
      const int thi = m % p.tiles_h;
      const int img = m / p.tiles_h;
      const int oh = thi * p.TH + hl, ow = twi * p.TW + wl;
      const bool valid = (oh < p.OH) && (ow < p.OW);
      __nv_bfloat16* dst = p.out + img * p.out_sN + oh * p.out_sH + ow * p.out_sW + nt * BN;

      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c * 32, r);
        tmem_ld_wait();
        if (p.psum) {  // warp-uniform; statistics of the values as they are stored (bf16), invalid rows count as 0
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = valid ? __bfloat162float(__float2bfloat16_rn(__uint_as_float(r[j]))) : 0.f;
          double s = 0.0, q = 0.0;
          warp_colstats32(v, lane, s, q);
          st_s[c * 32 + lane] += s;
          st_q[c * 32 + lane] += q;
        }
        if (valid) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[q * 8 + j]);
            if constexpr (FUSED) {  // warp-uniform addresses: one L1 line serves the whole warp
              const int ch = nt * BN + c * 32 + q * 8;
              const float4* sc = reinterpret_cast<const float4*>(p.scale + ch);
              const float4* sh = reinterpret_cast<const float4*>(p.shift + ch);
              const float4 s0 = __ldg(sc), s1 = __ldg(sc + 1), h0 = __ldg(sh), h1 = __ldg(sh + 1);
              f[0] = fmaf(f[0], s0.x, h0.x); f[1] = fmaf(f[1], s0.y, h0.y);
              f[2] = fmaf(f[2], s0.z, h0.z); f[3] = fmaf(f[3], s0.w, h0.w);
              f[4] = fmaf(f[4], s1.x, h1.x); f[5] = fmaf(f[5], s1.y, h1.y);
              f[6] = fmaf(f[6], s1.z, h1.z); f[7] = fmaf(f[7], s1.w, h1.w);
              if (p.res) {
                const uint4 old = *reinterpret_cast<const uint4*>(p.res + (dst - p.out) + c * 32 + q * 8);
                const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 o = __bfloat1622float2(ob[j]);
                  f[2 * j] += o.x;
                  f[2 * j + 1] += o.y;
                }
              }
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
              }
            }
            if (p.accumulate) {
              const uint4 old = d4[q];
              const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 o = __bfloat1622float2(ob[j]);
                f[2 * j] += o.x;
                f[2 * j + 1] += o.y;
              }
            }
            uint4 v;
            __nv_bfloat162* vb = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) vb[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            d4[q] = v;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    if (p.psum) {
      // the host launches a grid that is a multiple of n_tiles_n, so every tile of this CTA has the same nt
      const int nt = blockIdx.x % p.n_tiles_n;
      const size_t row = ((size_t)blockIdx.x * 4 + quad) * p.stats_C;
      for (int c = lane; c < p.stats_C; c += 32) {
        const int k = c - nt * BN;
        const bool mine = (k >= 0 && k < BN);
        p.psum[row + c] = mine ? (float)st_s[k] : 0.f;
        p.psq[row + c] = mine ? (float)st_q[k] : 0.f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// ---------------------------------------------------------------------------------------------------------------------
// The same GEMM on CTA PAIRS (tcgen05 cta_group::2): the two CTAs of a cluster work on two adjacent 128-pixel tiles and
// the same N tile; ONE MMA of M = 256 per K step, issued by rank 0, reads rank r's 128 A rows and rank r's HALF of the
// B tile from rank r's shared memory and writes rank r's accumulator rows to rank r's TMEM.
// Why (findings): the shared memory of an SM moves 128 B/clk, and TMA writes and MMA operand reads share
// it.  The single-CTA kernel writes and reads (16 KB A + BN x 128 B of B) per 64-wide K chunk: 256 B/clk at BN = 128 (tensor pipe <=
// 50 % busy: measured 965 of 1940 TFLOP/s at the power-capped clock), 188 B/clk at BN = 256 (<= 68 %: measured 1320).
// With half of B per CTA: 192 B/clk (<= 67 %) at BN = 128 and 128 B/clk (not bound) at BN = 256.
// Barriers: every CTA owns empty[] / tfull[] (arrived on by rank 0's multicast tcgen05.commit); full[] and tempty[]
// are rank 0's (both producers' TMA bytes and all 8 epilogue warps of the pair arrive there).
template <int BN, int STAGES>
struct NtPairSmem {
  static constexpr int kBHalf = (BN / 2) * 128;
  static constexpr int kStage = kATile + kBHalf;
  static constexpr int kBarOff = STAGES * kStage;
  static constexpr int kStatsOff = kBarOff + 256;  // double [4 warps][2][BN]
  static constexpr int kBytes = kStatsOff + 4 * 2 * BN * 8 + 1024;
};

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
    igemm_nt_pair_kernel(const __grid_constant__ NtParams p) {
  using L = NtPairSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // identical in both CTAs (same kernel, same layout)

  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kATile;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_maps[i]);
    tma_prefetch_desc(&p.b_map_half);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);  // 4 epilogue warps of each CTA
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 2 * BN);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier initialisation and the TMEM address are visible to the whole pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nkb = p.ntaps * p.k_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int q = cid; q < p.total_pairs; q += n_clusters) {
        const int nt = q % p.n_tiles_n;
        int m = 2 * (q / p.n_tiles_n) + (int)rank;  // beyond the last tile: coordinates outside the tensor, zero fill
        const int twi = m % p.tiles_w;
        m /= p.tiles_w;
        const int thi = m % p.tiles_h;
        const int img = m / p.tiles_h;
        const int h0 = thi * p.TH, w0 = twi * p.TW;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const Tap tp = p.taps[tap];
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * L::kStage);  // the bytes of BOTH CTAs land on this barrier
            tma_load_4d_pair(sA + stage * kATile, &p.a_maps[tp.map], &full[stage], kc * 64, w0 + tp.dw, h0 + tp.dh,
                             img);
            tma_load_2d_pair(sB + stage * L::kBHalf, &p.b_map_half, &full[stage], tp.wk + kc * 64,
                             nt * BN + (int)rank * (BN / 2));
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (rank 0 only, one elected thread)
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, 0, 0);
      const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA), 0, 1024);
      const uint64_t b_desc0 = make_sw128_desc(smem_u32(sB), 0, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int q = cid; q < p.total_pairs; q += n_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + (uint64_t)(stage * (kATile >> 4));
          const uint64_t b_desc = b_desc0 + (uint64_t)(stage * (L::kBHalf >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair(&empty[stage]);
          if (kb == nkb - 1) umma_commit_pair(&tfull[acc]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (both CTAs: own 128 rows, all BN columns)
    const int quad = warp & 3;
    const int m_row = quad * 32 + lane;
    const int hl = m_row >> p.tw_shift;
    const int wl = m_row & (p.TW - 1);
    double* st_s = reinterpret_cast<double*>(smem + L::kStatsOff) + quad * 2 * BN;
    double* st_q = st_s + BN;
    if (p.psum)
      for (int c = lane; c < BN; c += 32) st_s[c] = st_q[c] = 0.0;
    int it = 0;
    for (int q = cid; q < p.total_pairs; q += n_clusters, ++it) {
      const int nt = q % p.n_tiles_n;
      const int mt = 2 * (q / p.n_tiles_n) + (int)rank;
      int m = mt;
      const int twi = m % p.tiles_w;
      m /= p.tiles_w;
      const int thi = m % p.tiles_h;
      const int img = m / p.tiles_h;
      const int oh = thi * p.TH + hl, ow = twi * p.TW + wl;
      const bool valid = (mt < p.m_tiles) && (oh < p.OH) && (ow < p.OW);
      __nv_bfloat16* dst = p.out + img * p.out_sN + oh * p.out_sH + ow * p.out_sW + nt * BN;

      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c * 32, r);
        tmem_ld_wait();
        if (p.psum) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = valid ? __bfloat162float(__float2bfloat16_rn(__uint_as_float(r[j]))) : 0.f;
          double s = 0.0, qq = 0.0;
          warp_colstats32(v, lane, s, qq);
          st_s[c * 32 + lane] += s;
          st_q[c * 32 + lane] += qq;
        }
        if (valid) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[g * 8 + j]);
            if (p.accumulate) {
              const uint4 old = d4[g];
              const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 o = __bfloat1622float2(ob[j]);
                f[2 * j] += o.x;
                f[2 * j + 1] += o.y;
              }
            }
            uint4 v;
            __nv_bfloat162* vb = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) vb[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            d4[g] = v;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_rank0(&tempty[acc]);  // rank 0's MMA thread waits for all 8 warps of the pair
    }
    if (p.psum) {
      // the host launches a cluster count that is a multiple of n_tiles_n: every pair of this cluster has the same nt
      const int nt = cid % p.n_tiles_n;
      const size_t row = ((size_t)blockIdx.x * 4 + quad) * p.stats_C;
      for (int c = lane; c < p.stats_C; c += 32) {
        const int k = c - nt * BN;
        const bool mine = (k >= 0 && k < BN);
        p.psum[row + c] = mine ? (float)st_s[k] : 0.f;
        p.psq[row + c] = mine ? (float)st_q[k] : 0.f;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may leave (or free TMEM) while the other can still signal or be written to
  if (warp == 1) tmem_dealloc_pair(tmem_base, 2 * BN);
}

// CTA pairs are used for N tiles of 128 and 256 columns (64-column layers have their own kernels) when the problem has
// at least one pair of pixel tiles per N tile.  ECGMM_NT_PAIR=0 selects the single-CTA kernel (kept under test).
static bool nt_pair_enabled() {
  const char* e = getenv("ECGMM_NT_PAIR");
  return !(e && atoi(e) == 0);
}
static int nt_pair_clusters(int total_pairs, int n_tiles_n) {
  int c = total_pairs < num_sms() / 2 ? total_pairs : num_sms() / 2;
  if (n_tiles_n > 1 && c > n_tiles_n) c -= c % n_tiles_n;
  return c;
}

template <int BN, int STAGES>
static int launch_nt_pair_t(NtParams& p, cudaStream_t s) {
  using L = NtPairSmem<BN, STAGES>;
  static bool configured[kMaxDevices] = {};
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(igemm_nt_pair_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    L::kBytes));
    configured[ds] = true;
  }
  const int grid = 2 * nt_pair_clusters(p.total_pairs, p.n_tiles_n);
  igemm_nt_pair_kernel<BN, STAGES><<<grid, 192, L::kBytes, s>>>(p);
  return check_launch("igemm_nt_pair_kernel");
}

// Persistent grid: one CTA per SM, rounded down to a multiple of the number of N tiles so that a CTA always works
// on the same N tile (t % n_tiles_n with t = blockIdx.x + k * grid): the statistics epilogue relies on it.
static int nt_grid(int total_tiles, int n_tiles_n) {
  int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  if (n_tiles_n > 1 && grid > n_tiles_n) grid -= grid % n_tiles_n;
  return grid;
}

template <int BN, int STAGES>
static int launch_nt_t(const NtParams& p, cudaStream_t s) {
  using L = NtSmem<BN, STAGES>;
  static bool configured[kMaxDevices] = {};  // the shared-memory limit of a kernel is a per-device attribute
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(igemm_nt_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    L::kBytes));
    ECGMM_CUDA(cudaFuncSetAttribute(igemm_nt_kernel<BN, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    L::kBytes));
    configured[ds] = true;
  }
  int grid = nt_grid(p.total_tiles, p.n_tiles_n);
  if (p.scale)
    igemm_nt_kernel<BN, STAGES, true><<<grid, 192, L::kBytes, s>>>(p);
  else
    igemm_nt_kernel<BN, STAGES><<<grid, 192, L::kBytes, s>>>(p);
  return check_launch("igemm_nt_kernel");
}

// n_gemm = GEMM N extent (Cout forward, Cin dgrad)
static int launch_nt(NtParams& p, int n_gemm, cudaStream_t s) {
  ECGMM_CHECK(n_gemm % 64 == 0, ECGMM_ERR_SHAPE, "GEMM N=%d must be a multiple of 64", n_gemm);
  int bn = (n_gemm % 256 == 0) ? 256 : (n_gemm % 128 == 0 ? 128 : 64);
  p.n_tiles_n = n_gemm / bn;
  p.m_tiles = p.n_img * p.tiles_h * p.tiles_w;
  p.total_tiles = p.m_tiles * p.n_tiles_n;
  if (p.total_tiles == 0) return ECGMM_OK;
  if (nt_pair_enabled() && bn >= 128 && !p.scale && p.pair_ok && p.m_tiles >= 2) {
    p.total_pairs = ((p.m_tiles + 1) / 2) * p.n_tiles_n;
    return bn == 256 ? launch_nt_pair_t<256, 6>(p, s) : launch_nt_pair_t<128, 8>(p, s);
  }
  if (bn == 256) return launch_nt_t<256, 4>(p, s);
  if (bn == 128) return launch_nt_t<128, 6>(p, s);
  return launch_nt_t<64, 6>(p, s);
}

static void set_tile_grid(NtParams& p, int n_img, int OH, int OW) {
  pick_tile(OH, OW, 128, &p.TH, &p.TW);
  p.tw_shift = 0;
  while ((1 << p.tw_shift) < p.TW) ++p.tw_shift;
  p.tiles_h = ceil_div(OH, p.TH);
  p.tiles_w = ceil_div(OW, p.TW);
  p.n_img = n_img;
  p.OH = OH;
  p.OW = OW;
}

static int check_conv_cfg(int Cin, int Cout, int R, int S, int stride, int padH, int padW) {
  ECGMM_CHECK(Cin > 0 && Cout > 0 && Cin % 64 == 0 && Cout % 64 == 0, ECGMM_ERR_SHAPE,
              "conv channels must be positive multiples of 64 (Cin=%d Cout=%d)", Cin, Cout);
  ECGMM_CHECK((R == 1 || R == 3) && (S == 1 || S == 3) && R * S <= kMaxTaps, ECGMM_ERR_SHAPE,
              "unsupported filter %dx%d", R, S);
  ECGMM_CHECK(stride == 1 || stride == 2, ECGMM_ERR_SHAPE, "unsupported stride %d", stride);
  ECGMM_CHECK(padH == R / 2 && padW == S / 2, ECGMM_ERR_SHAPE, "padding must be (R/2,S/2), got (%d,%d)", padH,
              padW);
  return ECGMM_OK;
}

// Activation views for a convolution input x[N][H][W][C] read with `stride`:
// stride 1 -> map 0 is the plain tensor; stride 2 -> map (ph*2+pw) is the parity sub-lattice.
static int make_input_maps(CUtensorMap* maps, bool* used, const __nv_bfloat16* x, int N, int H, int W, int C,
                           int stride, int box_w, int box_h) {
  const uint64_t e = sizeof(__nv_bfloat16);
  if (stride == 1) {
    used[0] = true;
    return make_tmap_4d(&maps[0], x, C, W, H, N, (uint64_t)C * e, (uint64_t)W * C * e, (uint64_t)H * W * C * e, 64,
                        box_w, box_h);
  }
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const int id = ph * 2 + pw;
      if (!used[id]) continue;
      const int Hp = (H - ph + 1) / 2, Wp = (W - pw + 1) / 2;
      ECGMM_CHECK(Hp > 0 && Wp > 0, ECGMM_ERR_SHAPE, "empty parity view (H=%d W=%d)", H, W);
      int rc = make_tmap_4d(&maps[id], x + ((size_t)ph * W + pw) * C, C, Wp, Hp, N, 2ull * C * e, 2ull * W * C * e,
                            (uint64_t)H * W * C * e, 64, box_w, box_h);
      if (rc) return rc;
    }
  return ECGMM_OK;
}

// floor division by 2 for possibly negative values
static inline int fdiv2(int v) { return (v >= 0) ? v / 2 : -((-v + 1) / 2); }

// Fill the forward tap table (also used by wgrad: same input addressing).
static int build_fwd_taps(Tap* taps, bool* used, int R, int S, int stride, int padH, int padW, int k_per_tap) {
  int n = 0;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      Tap t;
      if (stride == 1) {
        t.map = 0;
        t.dh = r - padH;
        t.dw = s - padW;
      } else {
        const int ph = (r - padH) & 1, pw = (s - padW) & 1;
        t.map = ph * 2 + pw;
        t.dh = fdiv2(r - padH - ph);
        t.dw = fdiv2(s - padW - pw);
      }
      used[t.map] = true;
      t.id = r * S + s;
      t.wk = (r * S + s) * k_per_tap;
      taps[n++] = t;
    }
  return n;
}

}  // namespace ecgmm

using namespace ecgmm;

namespace ecgmm {
bool nt_halo_supported(int Cin, int Cout, int R, int S, int stride, int W);
int launch_nt_halo(const __nv_bfloat16* x, const __nv_bfloat16* w, __nv_bfloat16* y, int N, int H, int W, int R, int S,
                   int dgrad, int accumulate, cudaStream_t st, const float* scale = nullptr,
                   const float* shift = nullptr, const __nv_bfloat16* res = nullptr, int relu = 0);
// rolling-accumulator kernel (conv_nt_stack.cu)
bool nt_stack_supported(int Cin, int Cout, int R, int S, int stride, int W);
// The rolling-accumulator kernel (conv_nt_stack.cu) is the default for the 64 -> 64 3x3 layers (measured inside the
// training step at batch 512: forward 1068 -> 1296, accumulating data gradient 961 -> 1109 TFLOP/s);
// ECGMM_NT_STACK=0 selects igemm_nt_halo_kernel again (kept under test; it also serves the 1x3 layers and the
// folded-BatchNorm epilogue of the serving path).
static bool nt_stack_enabled() {
  const char* e = getenv("ECGMM_NT_STACK");
  return !(e && atoi(e) == 0);
}
int launch_nt_stack(const __nv_bfloat16* x, const __nv_bfloat16* w, __nv_bfloat16* y, int N, int H, int W, int dgrad,
                    int accumulate, cudaStream_t st, float* psum = nullptr, float* psq = nullptr,
                    const __nv_bfloat16* red_x = nullptr, const uint8_t* red_mask = nullptr,
                    const float* red_mean = nullptr, const float* red_invstd = nullptr);
int nt_stack_stats_rows(int N, int H, int W);
}  // namespace ecgmm

// Rows of the statistics partials the forward kernel writes for this shape (4 epilogue warps per CTA of the generic
// kernel).  0 = not offered: the reduction sits in the epilogue, which has slack only when a tile holds many MMAs.
// Measured at batch 64 (profiles/README.md): layers 2-4 (K >= 576) +0.04 ms on three convolutions against a 0.13 ms
// statistics pass, the 64->128 stride-2 convolution (K = 576) breaks even; the 64->64 halo kernel (36 MMAs per tile) +0.33 ms against 0.28 ms and the stem (16 MMAs) +0.3 ms
// against 0.2 ms, whether the reduction is a register shuffle transpose or a read-back of the staging tile.
extern "C" int ecgmm_conv2d_fwd_stats_rows(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int padH,
                                           int padW) {
  if (N <= 0 || check_conv_cfg(Cin, Cout, R, S, stride, padH, padW)) return 0;
  // the rolling-accumulator kernel keeps the sums in registers (one row of partials per CTA)
  if (nt_stack_enabled() && nt_stack_supported(Cin, Cout, R, S, stride, W)) return nt_stack_stats_rows(N, H, W);
  if (nt_halo_supported(Cin, Cout, R, S, stride, W) && !getenv("ECGMM_NT_LEGACY")) return 0;
  if (R * S * Cin < 1152) return 0;
  const int Ho = (H + 2 * padH - R) / stride + 1, Wo = (W + 2 * padW - S) / stride + 1;
  NtParams p;
  memset(&p, 0, sizeof(p));
  set_tile_grid(p, N, Ho, Wo);
  const int bn = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
  const int n_tiles_n = Cout / bn;
  const int m_tiles = p.n_img * p.tiles_h * p.tiles_w;
  if (nt_pair_enabled() && bn >= 128 && m_tiles >= 2)  // CTA pairs: 4 epilogue warps in each of 2 * clusters CTAs
    return 4 * 2 * nt_pair_clusters(((m_tiles + 1) / 2) * n_tiles_n, n_tiles_n);
  return 4 * nt_grid(m_tiles * n_tiles_n, n_tiles_n);
}

struct FusedEpilogue {  // folded BatchNorm (+ residual, + ReLU) applied by the forward epilogue; scale == NULL: off
  const float* scale = nullptr;
  const float* shift = nullptr;
  const __nv_bfloat16* res = nullptr;
  int relu = 0;
};

static int conv2d_fwd_impl(const ecgmm_bf16* x_, const ecgmm_bf16* w_, ecgmm_bf16* y_, float* psum, float* psq, int N,
                           int H, int W, int Cin, int Cout, int R, int S, int stride, int padH, int padW,
                           void* stream, const FusedEpilogue& fe = FusedEpilogue()) {
  ECGMM_CHECK(x_ && w_ && y_, ECGMM_ERR_ARG, "conv2d_fwd: null pointer");
  int rc = check_conv_cfg(Cin, Cout, R, S, stride, padH, padW);
  if (rc) return rc;
  if (N == 0) return ECGMM_OK;
  if (!fe.scale && nt_stack_enabled() && nt_stack_supported(Cin, Cout, R, S, stride, W))
    return launch_nt_stack(reinterpret_cast<const __nv_bfloat16*>(x_), reinterpret_cast<const __nv_bfloat16*>(w_),
                           reinterpret_cast<__nv_bfloat16*>(y_), N, H, W, 0, 0, static_cast<cudaStream_t>(stream), psum,
                           psq);
  if (nt_halo_supported(Cin, Cout, R, S, stride, W) && !getenv("ECGMM_NT_LEGACY")) {
    ECGMM_CHECK(!psum, ECGMM_ERR_SHAPE, "conv2d_fwd_stats: not offered for this shape (see ecgmm_conv2d_fwd_stats_rows)");
    return launch_nt_halo(reinterpret_cast<const __nv_bfloat16*>(x_), reinterpret_cast<const __nv_bfloat16*>(w_),
                          reinterpret_cast<__nv_bfloat16*>(y_), N, H, W, R, S, 0, 0, static_cast<cudaStream_t>(stream),
                          fe.scale, fe.shift, fe.res, fe.relu);
  }
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(x_);
  const int Ho = (H + 2 * padH - R) / stride + 1, Wo = (W + 2 * padW - S) / stride + 1;
  NtParams p;
  memset(&p, 0, sizeof(p));
  set_tile_grid(p, N, Ho, Wo);
  bool used[4] = {false, false, false, false};
  p.ntaps = build_fwd_taps(p.taps, used, R, S, stride, padH, padW, Cin);
  p.k_chunks = Cin / 64;
  rc = make_input_maps(p.a_maps, used, x, N, H, W, Cin, stride, p.TW, p.TH);
  if (rc) return rc;
  for (int i = 1; i < 4; ++i)
    if (!used[i]) p.a_maps[i] = p.a_maps[p.taps[0].map];
  if (!used[0]) p.a_maps[0] = p.a_maps[p.taps[0].map];
  const int bn = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
  rc = make_tmap_2d(&p.b_map, w_, (uint64_t)R * S * Cin, Cout, (uint64_t)R * S * Cin * 2, 64, bn);
  if (rc) return rc;
  if (bn >= 128) {
    rc = make_tmap_2d(&p.b_map_half, w_, (uint64_t)R * S * Cin, Cout, (uint64_t)R * S * Cin * 2, 64, bn / 2);
    if (rc) return rc;
    p.pair_ok = 1;
  }
  p.out = reinterpret_cast<__nv_bfloat16*>(y_);
  p.out_sW = Cout;
  p.out_sH = (long long)Wo * Cout;
  p.out_sN = (long long)Ho * Wo * Cout;
  p.accumulate = 0;
  p.psum = psum;
  p.psq = psq;
  p.stats_C = Cout;
  p.scale = fe.scale;
  p.shift = fe.shift;
  p.res = fe.res;
  p.relu = fe.relu;
  return launch_nt(p, Cout, static_cast<cudaStream_t>(stream));
}

extern "C" int ecgmm_conv2d_fwd_bn(const ecgmm_bf16* x, const ecgmm_bf16* w, ecgmm_bf16* y, const float* scale,
                                   const float* shift, const ecgmm_bf16* res, int relu, int N, int H, int W, int Cin,
                                   int Cout, int R, int S, int stride, int padH, int padW, void* stream) {
  ECGMM_CHECK(scale && shift, ECGMM_ERR_ARG, "conv2d_fwd_bn: null scale / shift");
  ECGMM_CHECK((reinterpret_cast<uintptr_t>(scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(shift) & 15) == 0,
              ECGMM_ERR_ALIGN, "conv2d_fwd_bn: scale / shift must be 16-byte aligned");
  ECGMM_CHECK(res == nullptr || (reinterpret_cast<uintptr_t>(res) & 15) == 0, ECGMM_ERR_ALIGN,
              "conv2d_fwd_bn: residual must be 16-byte aligned");
  FusedEpilogue fe;
  fe.scale = scale;
  fe.shift = shift;
  fe.res = reinterpret_cast<const __nv_bfloat16*>(res);
  fe.relu = relu;
  return conv2d_fwd_impl(x, w, y, nullptr, nullptr, N, H, W, Cin, Cout, R, S, stride, padH, padW, stream, fe);
}

extern "C" int ecgmm_conv2d_fwd(const ecgmm_bf16* x, const ecgmm_bf16* w, ecgmm_bf16* y, int N, int H, int W, int Cin,
                                int Cout, int R, int S, int stride, int padH, int padW, void* stream) {
  return conv2d_fwd_impl(x, w, y, nullptr, nullptr, N, H, W, Cin, Cout, R, S, stride, padH, padW, stream);
}

extern "C" int ecgmm_conv2d_fwd_stats(const ecgmm_bf16* x, const ecgmm_bf16* w, ecgmm_bf16* y, float* psum, float* psq,
                                      int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int padH,
                                      int padW, void* stream) {
  ECGMM_CHECK(psum && psq, ECGMM_ERR_ARG, "conv2d_fwd_stats: null statistics buffer");
  return conv2d_fwd_impl(x, w, y, psum, psq, N, H, W, Cin, Cout, R, S, stride, padH, padW, stream);
}

// BatchNorm-backward sums from the data-gradient epilogue: rows of the [rows][Cin] partials, 0 = not offered.
extern "C" int ecgmm_conv2d_dgrad_reduce_rows(int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                                              int padH, int padW) {
  if (N <= 0 || check_conv_cfg(Cin, Cout, R, S, stride, padH, padW)) return 0;
  if (nt_stack_enabled() && nt_stack_supported(Cin, Cout, R, S, stride, W)) return nt_stack_stats_rows(N, H, W);
  return 0;
}

static int conv2d_dgrad_impl(const ecgmm_bf16* dy_, const ecgmm_bf16* wt_, ecgmm_bf16* dx_, int N, int H, int W,
                             int Cin, int Cout, int R, int S, int stride, int padH, int padW, int accumulate,
                             void* stream, const ecgmm_bf16* red_x, const uint8_t* red_mask, const float* red_mean,
                             const float* red_invstd, float* p1, float* p2);

extern "C" int ecgmm_conv2d_dgrad_reduce(const ecgmm_bf16* dy, const ecgmm_bf16* wt, ecgmm_bf16* dx,
                                         const ecgmm_bf16* bn_x, const uint8_t* bn_mask, const float* bn_mean,
                                         const float* bn_invstd, float* p1, float* p2, int N, int H, int W, int Cin,
                                         int Cout, int R, int S, int stride, int padH, int padW, int accumulate,
                                         void* stream) {
  ECGMM_CHECK(bn_x && bn_mean && bn_invstd && p1 && p2, ECGMM_ERR_ARG, "conv2d_dgrad_reduce: null pointer");
  ECGMM_CHECK(ecgmm_conv2d_dgrad_reduce_rows(N, H, W, Cin, Cout, R, S, stride, padH, padW) > 0 || N == 0,
              ECGMM_ERR_SHAPE, "conv2d_dgrad_reduce: not offered for this shape (see ecgmm_conv2d_dgrad_reduce_rows)");
  return conv2d_dgrad_impl(dy, wt, dx, N, H, W, Cin, Cout, R, S, stride, padH, padW, accumulate, stream, bn_x,
                           bn_mask, bn_mean, bn_invstd, p1, p2);
}

extern "C" int ecgmm_conv2d_dgrad(const ecgmm_bf16* dy_, const ecgmm_bf16* wt_, ecgmm_bf16* dx_, int N, int H,
                                  int W, int Cin, int Cout, int R, int S, int stride, int padH, int padW,
                                  int accumulate, void* stream) {
  return conv2d_dgrad_impl(dy_, wt_, dx_, N, H, W, Cin, Cout, R, S, stride, padH, padW, accumulate, stream, nullptr,
                           nullptr, nullptr, nullptr, nullptr, nullptr);
}

static int conv2d_dgrad_impl(const ecgmm_bf16* dy_, const ecgmm_bf16* wt_, ecgmm_bf16* dx_, int N, int H, int W,
                             int Cin, int Cout, int R, int S, int stride, int padH, int padW, int accumulate,
                             void* stream, const ecgmm_bf16* red_x, const uint8_t* red_mask, const float* red_mean,
                             const float* red_invstd, float* p1, float* p2) {
  ECGMM_CHECK(dy_ && wt_ && dx_, ECGMM_ERR_ARG, "conv2d_dgrad: null pointer");
  int rc = check_conv_cfg(Cin, Cout, R, S, stride, padH, padW);
  if (rc) return rc;
  if (N == 0) return ECGMM_OK;
  if (nt_stack_enabled() && nt_stack_supported(Cin, Cout, R, S, stride, W))
    return launch_nt_stack(reinterpret_cast<const __nv_bfloat16*>(dy_), reinterpret_cast<const __nv_bfloat16*>(wt_),
                           reinterpret_cast<__nv_bfloat16*>(dx_), N, H, W, 1, accumulate,
                           static_cast<cudaStream_t>(stream), p1, p2, reinterpret_cast<const __nv_bfloat16*>(red_x),
                           red_mask, red_mean, red_invstd);
  if (nt_halo_supported(Cin, Cout, R, S, stride, W) && !getenv("ECGMM_NT_LEGACY"))
    return launch_nt_halo(reinterpret_cast<const __nv_bfloat16*>(dy_), reinterpret_cast<const __nv_bfloat16*>(wt_),
                          reinterpret_cast<__nv_bfloat16*>(dx_), N, H, W, R, S, 1, accumulate,
                          static_cast<cudaStream_t>(stream));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(dy_);
  __nv_bfloat16* dx = reinterpret_cast<__nv_bfloat16*>(dx_);
  const int Ho = (H + 2 * padH - R) / stride + 1, Wo = (W + 2 * padW - S) / stride + 1;
  const uint64_t e = 2;
  const int bn = (Cin % 256 == 0) ? 256 : (Cin % 128 == 0 ? 128 : 64);

  // Input positions of parity (ph,pw) only receive taps with (ph+padH-r) and (pw+padW-s) even.
  const int nph = (stride == 2 && H > 1) ? 2 : 1, npw = (stride == 2 && W > 1) ? 2 : 1;
  bool need_zero = false;
  if (stride == 2 && !accumulate) {
    for (int ph = 0; ph < nph; ++ph)
      for (int pw = 0; pw < npw; ++pw) {
        int cnt = 0;
        for (int r = 0; r < R; ++r)
          for (int s = 0; s < S; ++s)
            if (((ph + padH - r) & 1) == 0 && ((pw + padW - s) & 1) == 0) ++cnt;
        if (cnt == 0) need_zero = true;
      }
    if (need_zero) ECGMM_CUDA(cudaMemsetAsync(dx, 0, (size_t)N * H * W * Cin * e, st));
  }

  for (int ph = 0; ph < nph; ++ph)
    for (int pw = 0; pw < npw; ++pw) {
      NtParams p;
      memset(&p, 0, sizeof(p));
      int n = 0;
      for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s) {
          Tap t;
          t.map = 0;
          t.id = r * S + s;
          t.wk = (r * S + s) * Cout;
          if (stride == 1) {
            t.dh = padH - r;
            t.dw = padW - s;
          } else {
            if (((ph + padH - r) & 1) || ((pw + padW - s) & 1)) continue;
            t.dh = (ph + padH - r) / 2;
            t.dw = (pw + padW - s) / 2;
          }
          p.taps[n++] = t;
        }
      if (n == 0) continue;
      p.ntaps = n;
      p.k_chunks = Cout / 64;
      const int Hc = (stride == 1) ? H : (H - ph + 1) / 2, Wc = (stride == 1) ? W : (W - pw + 1) / 2;
      if (Hc <= 0 || Wc <= 0) continue;
      set_tile_grid(p, N, Hc, Wc);
      rc = make_tmap_4d(&p.a_maps[0], dy, Cout, Wo, Ho, N, (uint64_t)Cout * e, (uint64_t)Wo * Cout * e,
                        (uint64_t)Ho * Wo * Cout * e, 64, p.TW, p.TH);
      if (rc) return rc;
      for (int i = 1; i < 4; ++i) p.a_maps[i] = p.a_maps[0];
      rc = make_tmap_2d(&p.b_map, wt_, (uint64_t)R * S * Cout, Cin, (uint64_t)R * S * Cout * 2, 64, bn);
      if (rc) return rc;
      if (bn >= 128) {
        rc = make_tmap_2d(&p.b_map_half, wt_, (uint64_t)R * S * Cout, Cin, (uint64_t)R * S * Cout * 2, 64, bn / 2);
        if (rc) return rc;
        p.pair_ok = 1;
      }
      p.out = dx + ((size_t)ph * W + pw) * Cin * (stride == 2 ? 1 : 0);
      p.out_sW = (long long)stride * Cin;
      p.out_sH = (long long)stride * W * Cin;
      p.out_sN = (long long)H * W * Cin;
      p.accumulate = accumulate;
      rc = launch_nt(p, Cin, st);
      if (rc) return rc;
    }
  return ECGMM_OK;
}

// =====================================================================================
// Weight gradient
// =====================================================================================
namespace ecgmm {

constexpr int kKPix = 32;          // output pixels (GEMM K) per pipeline stage
constexpr int kBox = kKPix * 128;  // one TMA box: 32 pixel rows x 64 channels = 4 KiB

struct alignas(64) TnParams {
  CUtensorMap x_maps[4];
  CUtensorMap dy_map;
  Tap taps[kMaxTaps];
  int ntaps;
  int cin_chunks;  // Cin / 64
  int n_atoms;     // ntaps * cin_chunks : 64-row pieces of the dW matrix
  int n_slots;     // ceil(n_atoms / 2)  : 128-row accumulators
  int slots_per_group, n_groups_m, n_tiles_n, ksplit;
  int TH, TW, tiles_h, tiles_w, n_img, total_kblocks;
  float* dw;
  float* ws;  // split-K partials [CTA][slot][128 rows][BN] (deterministic fold by tn_reduce_kernel); NULL: fp32 atomics
  int Cin, Cout, RS;
  int mode;  // 0: dw is OIHW [Cout][Cin][R][S];  1: ResNet stem (space-to-depth atoms -> [64][3][7][7])
};

template <int BN, int SMAX, int STAGES>
struct TnSmem {
  static constexpr int kNB = BN / 64;  // dY boxes per stage
  static constexpr int kStage = (kNB + 2 * SMAX) * kBox;
  static constexpr int kBarOff = STAGES * kStage;
  static constexpr int kBytes = kBarOff + 256 + 1024;
};

template <int BN, int SMAX, int STAGES>
__global__ void __launch_bounds__(192, 1) igemm_tn_kernel(const __grid_constant__ TnParams p) {
  using L = TnSmem<BN, SMAX, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.x_maps[i]);
    tma_prefetch_desc(&p.dy_map);
    for (int i = 0; i < STAGES; ++i) {
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

  // work decomposition: blockIdx -> (k-split id, slot group, N tile)
  const int ks = blockIdx.x % p.ksplit;
  const int g = blockIdx.x / p.ksplit;
  const int gm = g % p.n_groups_m;
  const int nt = g / p.n_groups_m;
  const int s0 = gm * p.slots_per_group;
  const int ns = min(p.slots_per_group, p.n_slots - s0);
  const int per = (p.total_kblocks + p.ksplit - 1) / p.ksplit;
  const int kb0 = ks * per;
  const int kb1 = min(p.total_kblocks, kb0 + per);
  const bool has_work = (kb0 < kb1) && (ns > 0);

  if (has_work) {
    if (warp == 0) {
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          int m = kb;
          const int twi = m % p.tiles_w;
          m /= p.tiles_w;
          const int thi = m % p.tiles_h;
          const int img = m / p.tiles_h;
          const int h0 = thi * p.TH, w0 = twi * p.TW;
          uint8_t* st = smem + stage * L::kStage;
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], (L::kNB + 2 * ns) * kBox);
#pragma unroll
          for (int j = 0; j < L::kNB; ++j)
            tma_load_4d(st + j * kBox, &p.dy_map, &full[stage], nt * BN + j * 64, w0, h0, img);
          for (int i = 0; i < 2 * ns; ++i) {
            int atom = 2 * s0 + i;
            if (atom >= p.n_atoms) atom = p.n_atoms - 1;  // odd atom count: dummy second half
            const int tap = atom / p.cin_chunks;
            const int cc = atom - tap * p.cin_chunks;
            const Tap tp = p.taps[tap];
            tma_load_4d(st + (L::kNB + i) * kBox, &p.x_maps[tp.map], &full[stage], cc * 64, w0 + tp.dw, h0 + tp.dh,
                        img);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
        const uint32_t s_addr = smem_u32(smem);
        // MN-major operands: 64-wide atoms kBox bytes apart (LBO), 8 K-rows = 1024 B (SBO)
        const uint64_t b_desc0 = make_sw128_desc(s_addr, kBox, 1024);
        const uint64_t a_desc0 = make_sw128_desc(s_addr + L::kNB * kBox, kBox, 1024);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t so = (uint64_t)(stage * (L::kStage >> 4));
#pragma unroll
          for (int i = 0; i < SMAX; ++i) {
            if (i < ns) {
#pragma unroll
              for (int k = 0; k < kKPix / 16; ++k)  // 16 K-rows = 2048 B further into the box
                umma_bf16(tmem_base + i * BN, a_desc0 + so + (uint64_t)(i * (2 * kBox >> 4) + k * 128),
                          b_desc0 + so + (uint64_t)(k * 128), idesc, (kb > kb0) || (k > 0));
            }
          }
          umma_commit(&empty[stage]);
          if (kb == kb1 - 1) umma_commit(tfull);
          if (++stage == STAGES) {
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
      for (int i = 0; i < ns; ++i) {
        const int atom = 2 * (s0 + i) + (m_row >> 6);
        const bool row_ok = atom < p.n_atoms;
        const int a_c = row_ok ? atom : 0;
        const int tap = a_c / p.cin_chunks;
        const int cc = a_c - tap * p.cin_chunks;
        const int e = m_row & 63;
        long long row_off = 0, col_stride = 0;
        bool ok = row_ok;
        if (p.mode == 0) {
          const int cin = cc * 64 + e;
          row_off = (long long)cin * p.RS + p.taps[tap].id;
          col_stride = (long long)p.Cin * p.RS;
        } else {
          const int ra = p.taps[tap].id, sa = e >> 4, ch = e & 15;
          const int dr = ch / 6, ds = (ch % 6) / 3, c = ch % 3;
          const int r = 2 * ra + dr, s = 2 * sa + ds;
          ok = ok && ch < 12 && r < 7 && s < 7;
          row_off = ((long long)c * 7 + r) * 7 + s;
          col_stride = 147;
        }
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + i * BN;
        if (p.ws) {  // deterministic path: this CTA's partial tile goes to the workspace as it is
          float4* dst = reinterpret_cast<float4*>(
              p.ws + (((size_t)blockIdx.x * p.slots_per_group + i) * 128 + m_row) * BN);
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(t_addr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[c * 8 + j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                           __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          }
          continue;
        }
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_addr + c * 32, r);
          tmem_ld_wait();
          if (ok) {
            float* dst = p.dw + row_off + (long long)(nt * BN + c * 32) * col_stride;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + j * col_stride, __uint_as_float(r[j]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Sum of one element's n_ks split-K partials in a FIXED order (both folds below use it: bit-identical results).
// Splits <= 8: four interleaved accumulators.  Larger splits (the 1x1 projections: one tile, up to 148 partials): 16
// loads in flight per thread instead of 4 -- the fold was a chain of ~37 L2 round trips (25 us for 8 K outputs).
__device__ __forceinline__ float tn_fold(const float* __restrict__ src, size_t stride, int n_ks) {
  if (n_ks > 8) {
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 0.f;
    int k = 0;
    for (; k + 15 < n_ks; k += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] += src[(size_t)(k + j) * stride];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (k + j < n_ks) a[j] += src[(size_t)(k + j) * stride];
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
      for (int j = 0; j < w; ++j) a[j] += a[j + w];
    return a[0];
  }
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int k = 0;
  for (; k + 3 < n_ks; k += 4) {
    a0 += src[(size_t)k * stride];
    a1 += src[(size_t)(k + 1) * stride];
    a2 += src[(size_t)(k + 2) * stride];
    a3 += src[(size_t)(k + 3) * stride];
  }
  for (; k < n_ks; ++k) a0 += src[(size_t)k * stride];
  return (a0 + a1) + (a2 + a3);
}

// Fixed-order fold of the split-K partials written by igemm_tn_kernel (mode 0): one thread per dW element of a slot
// group, partials added in k-split order, then ONE += into dw -- bit-reproducible, unlike the atomics it replaces.
__global__ void __launch_bounds__(256) tn_reduce_kernel(TnParams p, int BN) {
  const long long per_group = (long long)p.slots_per_group * 128 * BN;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_group * p.n_groups_m * p.n_tiles_n) return;
  const int g = (int)(idx / per_group);
  long long rem = idx - (long long)g * per_group;
  const int i = (int)(rem / (128 * BN));
  rem -= (long long)i * 128 * BN;
  const int m_row = (int)(rem / BN), col = (int)(rem - (long long)m_row * BN);
  const int gm = g % p.n_groups_m, nt = g / p.n_groups_m;
  const int s0 = gm * p.slots_per_group;
  if (i >= min(p.slots_per_group, p.n_slots - s0)) return;
  const int atom = 2 * (s0 + i) + (m_row >> 6);
  if (atom >= p.n_atoms) return;
  const int tap = atom / p.cin_chunks, cc = atom - tap * p.cin_chunks;
  const int cin = cc * 64 + (m_row & 63), cout = nt * BN + col;
  const int per = (p.total_kblocks + p.ksplit - 1) / p.ksplit;
  const int n_ks = min(p.ksplit, (p.total_kblocks + per - 1) / per);  // CTAs with ks >= n_ks had no pixel block
  const size_t stride = (size_t)p.slots_per_group * 128 * BN;
  const float* src = p.ws + (size_t)g * p.ksplit * stride + ((size_t)i * 128 + m_row) * BN + col;
  p.dw[((size_t)cout * p.Cin + cin) * p.RS + p.taps[tap].id] += tn_fold(src, stride, n_ks);
}

// The same fold with coalesced accesses (OIHW output): a CTA owns 16 output channels x one 64-wide Cin slice, gathers
// all RS taps of it from wherever their atoms sit in the workspace (atom = tap * cin_chunks + slice -> accumulator
// slot, 64-row half, slot group) into a [16][64 cin x RS] tile in shared memory laid out like dw, and adds RS * 64
// contiguous floats per output channel.  Reads run along the output channel (contiguous in the workspace), writes along
// (cin, tap) (contiguous in dw); the kernel above does a 32-byte sector read-modify-write per 4 useful bytes.
// Same per-element summation order -> bit-identical.
constexpr int kTnRedM = 16;
constexpr int kTnCinSplit = 4;  // a CTA folds 16 couts x 16 cins x RS taps (a whole 64-wide slice per CTA left the
constexpr int kTnCin = 64 / kTnCinSplit;  // 512-channel stride-2 layer with 128 CTAs)
__global__ void __launch_bounds__(256) tn_reduce_tile_kernel(TnParams p, int BN) {
  __shared__ float tile[kTnRedM][kTnCin * 9 + 1];
  const int mblocks = BN / kTnRedM;
  int rest = blockIdx.x;
  const int cs = rest % kTnCinSplit;
  rest /= kTnCinSplit;
  const int mb = rest % mblocks;
  rest /= mblocks;
  const int cc = rest % p.cin_chunks, nt = rest / p.cin_chunks;
  const int mi = threadIdx.x & (kTnRedM - 1), r0 = threadIdx.x / kTnRedM;  // 16 couts x 16 row lanes
  const int per = (p.total_kblocks + p.ksplit - 1) / p.ksplit;
  const int n_ks = min(p.ksplit, (p.total_kblocks + per - 1) / per);
  const size_t stride = (size_t)p.slots_per_group * 128 * BN;
  const int col = mb * kTnRedM + mi;
  for (int tap = 0; tap < p.ntaps; ++tap) {
    const int atom = tap * p.cin_chunks + cc;
    const int s = atom >> 1, half = atom & 1;
    const int gm = s / p.slots_per_group, i = s - gm * p.slots_per_group;
    const int g = nt * p.n_groups_m + gm;
    const float* base = p.ws + (size_t)g * p.ksplit * stride + ((size_t)i * 128 + half * 64 + cs * kTnCin) * BN + col;
    const int id = p.taps[tap].id;
    for (int e = r0; e < kTnCin; e += 256 / kTnRedM) {
      const float* src = base + (size_t)e * BN;
      tile[mi][e * p.RS + id] = tn_fold(src, stride, n_ks);
    }
  }
  __syncthreads();
  const int row_len = kTnCin * p.RS;
  for (int row = 0; row < kTnRedM; ++row) {
    float* dst = p.dw + ((size_t)(nt * BN + mb * kTnRedM + row) * p.Cin + cc * 64 + cs * kTnCin) * p.RS;
    for (int j = threadIdx.x; j < row_len; j += 256) dst[j] += tile[row][j];
  }
}

// Decomposition shared by the launch and the workspace query.
static void tn_plan(TnParams& p, int BN, int SMAX) {
  p.n_atoms = p.ntaps * p.cin_chunks;
  p.n_slots = (p.n_atoms + 1) / 2;
  p.n_groups_m = ceil_div(p.n_slots, SMAX);
  p.slots_per_group = ceil_div(p.n_slots, p.n_groups_m);
  p.n_tiles_n = p.Cout / BN;
  const int groups = p.n_groups_m * p.n_tiles_n;
  int ksplit = num_sms() / groups;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > p.total_kblocks) ksplit = p.total_kblocks;
  p.ksplit = ksplit;
}
static void tn_shape(int Cout, int* BN, int* SMAX) {
  if (Cout % 256 == 0) { *BN = 256; *SMAX = 2; }
  else if (Cout % 128 == 0) { *BN = 128; *SMAX = 4; }
  else { *BN = 64; *SMAX = 5; }
}
static size_t tn_workspace_floats(const TnParams& p, int BN) {
  return (size_t)p.n_groups_m * p.n_tiles_n * p.ksplit * p.slots_per_group * 128 * BN;
}

template <int BN, int SMAX, int STAGES>
static int launch_tn_t(TnParams& p, cudaStream_t s) {
  using L = TnSmem<BN, SMAX, STAGES>;
  static bool configured[kMaxDevices] = {};  // the shared-memory limit of a kernel is a per-device attribute
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(igemm_tn_kernel<BN, SMAX, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    L::kBytes));
    configured[ds] = true;
  }
  tn_plan(p, BN, SMAX);
  const int groups = p.n_groups_m * p.n_tiles_n;
  igemm_tn_kernel<BN, SMAX, STAGES><<<groups * p.ksplit, 192, L::kBytes, s>>>(p);
  int rc = check_launch("igemm_tn_kernel");
  if (rc || !p.ws) return rc;
  const char* etr = getenv("ECGMM_TN_REDUCE_LEGACY");
  const char* emk = getenv("ECGMM_TN_REDUCE_MAXKS");
  // few partials: the scattered writes of the per-element fold dominate; many (the 1x1 projections, up to 148): its
  // parallelism wins.  Whole step at batch 64 with the threshold at 0 / 6 / 16 / 40 / 200 partials: 16.24 / 16.19 / 16.15 /
  // 16.01 / 16.21 ms.
  const bool tiled = etr ? atoi(etr) == 0 : (p.ksplit <= (emk ? atoi(emk) : 40));
  if (p.mode == 0 && p.RS <= 9 && p.ntaps == p.RS && tiled) {
    tn_reduce_tile_kernel<<<(unsigned)(p.n_tiles_n * (BN / kTnRedM) * p.cin_chunks * kTnCinSplit), 256, 0, s>>>(p, BN);
    return check_launch("tn_reduce_tile_kernel");
  }
  const long long total = (long long)groups * p.slots_per_group * 128 * BN;
  tn_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p, BN);
  return check_launch("tn_reduce_kernel");
}

static int launch_tn(TnParams& p, cudaStream_t s) {
  if (p.total_kblocks == 0) return ECGMM_OK;
  if (p.Cout % 256 == 0) return launch_tn_t<256, 2, 6>(p, s);
  if (p.Cout % 128 == 0) return launch_tn_t<128, 4, 4>(p, s);
  return launch_tn_t<64, 5, 4>(p, s);
}

}  // namespace ecgmm

namespace ecgmm {
bool wgrad_halo_supported(int Cin, int Cout, int R, int S, int stride);
int launch_wgrad_halo(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, int N, int H, int W, int Cin,
                      int Cout, int R, int S, int padH, int padW, void* workspace, size_t ws_bytes, cudaStream_t st);
size_t wgrad_halo_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S, int padH, int padW);
}  // namespace ecgmm

// Pixel-block decomposition of the generic weight-gradient kernel (stride-2 and 1x1 shapes).
static void tn_pixel_blocks(TnParams& p, int N, int Ho, int Wo) {
  pick_tile(Ho, Wo, kKPix, &p.TH, &p.TW);
  p.tiles_h = ceil_div(Ho, p.TH);
  p.tiles_w = ceil_div(Wo, p.TW);
  p.n_img = N;
  p.total_kblocks = N * p.tiles_h * p.tiles_w;
}

extern "C" long long ecgmm_conv2d_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                                                  int padH, int padW) {
  if (N <= 0 || check_conv_cfg(Cin, Cout, R, S, stride, padH, padW)) return 0;
  if (wgrad_halo_supported(Cin, Cout, R, S, stride) && !getenv("ECGMM_WGRAD_LEGACY"))
    return (long long)wgrad_halo_workspace_bytes(N, H, W, Cin, Cout, R, S, padH, padW);
  // generic kernel: split-K partials for the deterministic fold (without a workspace it falls back to fp32 atomics)
  TnParams p;
  memset(&p, 0, sizeof(p));
  tn_pixel_blocks(p, N, (H + 2 * padH - R) / stride + 1, (W + 2 * padW - S) / stride + 1);
  if (p.total_kblocks == 0) return 0;
  p.ntaps = R * S;
  p.cin_chunks = Cin / 64;
  p.Cout = Cout;
  int bn, smax;
  tn_shape(Cout, &bn, &smax);
  tn_plan(p, bn, smax);
  return (long long)(tn_workspace_floats(p, bn) * sizeof(float));
}

extern "C" int ecgmm_conv2d_wgrad(const ecgmm_bf16* x_, const ecgmm_bf16* dy_, float* dw, int N, int H, int W,
                                  int Cin, int Cout, int R, int S, int stride, int padH, int padW, void* workspace,
                                  long long workspace_bytes, void* stream) {
  ECGMM_CHECK(x_ && dy_ && dw, ECGMM_ERR_ARG, "conv2d_wgrad: null pointer");
  int rc = check_conv_cfg(Cin, Cout, R, S, stride, padH, padW);
  if (rc) return rc;
  if (N == 0) return ECGMM_OK;
  if (wgrad_halo_supported(Cin, Cout, R, S, stride) && !getenv("ECGMM_WGRAD_LEGACY"))
    return launch_wgrad_halo(reinterpret_cast<const __nv_bfloat16*>(x_), reinterpret_cast<const __nv_bfloat16*>(dy_),
                             dw, N, H, W, Cin, Cout, R, S, padH, padW, workspace,
                             workspace_bytes > 0 ? (size_t)workspace_bytes : 0, static_cast<cudaStream_t>(stream));
  const int Ho = (H + 2 * padH - R) / stride + 1, Wo = (W + 2 * padW - S) / stride + 1;
  TnParams p;
  memset(&p, 0, sizeof(p));
  tn_pixel_blocks(p, N, Ho, Wo);
  bool used[4] = {false, false, false, false};
  p.ntaps = build_fwd_taps(p.taps, used, R, S, stride, padH, padW, Cin);
  p.cin_chunks = Cin / 64;
  p.Cout = Cout;
  {  // a workspace of the size ecgmm_conv2d_wgrad_workspace() reports selects the deterministic fold
    int bn, smax;
    tn_shape(Cout, &bn, &smax);
    tn_plan(p, bn, smax);
    const size_t need = tn_workspace_floats(p, bn) * sizeof(float);
    p.ws = (workspace && workspace_bytes > 0 && (size_t)workspace_bytes >= need && p.total_kblocks > 0)
               ? reinterpret_cast<float*>(workspace) : nullptr;
  }
  rc = make_input_maps(p.x_maps, used, reinterpret_cast<const __nv_bfloat16*>(x_), N, H, W, Cin, stride, p.TW, p.TH);
  if (rc) return rc;
  for (int i = 0; i < 4; ++i)
    if (!used[i]) p.x_maps[i] = p.x_maps[p.taps[0].map];
  rc = make_tmap_4d(&p.dy_map, dy_, Cout, Wo, Ho, N, (uint64_t)Cout * 2, (uint64_t)Wo * Cout * 2,
                    (uint64_t)Ho * Wo * Cout * 2, 64, p.TW, p.TH);
  if (rc) return rc;
  p.dw = dw;
  p.Cin = Cin;
  p.Cout = Cout;
  p.RS = R * S;
  p.mode = 0;
  return launch_tn(p, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------- ResNet stem
extern "C" void ecgmm_stem_s2d_dims(int H, int W, int* Hs, int* Ws) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  if (Hs) *Hs = Ho + 3;
  if (Ws) *Ws = Wo + 3;
}

static int stem_input_map(CUtensorMap* m, const ecgmm_bf16* xs, int N, int H, int W, int box_w, int box_h) {
  int Hs, Ws;
  ecgmm_stem_s2d_dims(H, W, &Hs, &Ws);
  const int Wo = Ws - 3;
  // column b of the view = the 64 contiguous elements (4 pixels x 16 ch) starting at pixel b:
  // consecutive columns overlap (stride 32 B), which turns the 4 horizontal taps into GEMM K.
  return make_tmap_4d(m, xs, 64, Wo, Hs, N, 32, (uint64_t)Ws * 32, (uint64_t)Hs * Ws * 32, 64, box_w, box_h);
}

namespace ecgmm {
int launch_stem_fwd_ring(const void* xs, const void* w_s2d, __nv_bfloat16* y, int N, int H, int W, cudaStream_t st,
                         float* psum = nullptr, float* psq = nullptr);
int stem_fwd_ring_stats_rows(int N, int H, int W);
int launch_stem_wgrad_ring(const void* xs, const void* dy, float* dw, int N, int H, int W, void* workspace,
                           size_t ws_bytes, cudaStream_t st);
size_t stem_wgrad_ring_workspace_bytes(int N, int H, int W);
}  // namespace ecgmm

extern "C" int ecgmm_stem_conv_fwd_stats_rows(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0 || getenv("ECGMM_STEM_LEGACY")) return 0;
  return stem_fwd_ring_stats_rows(N, H, W);
}

extern "C" int ecgmm_stem_conv_fwd_stats(const ecgmm_bf16* xs, const ecgmm_bf16* w_s2d, ecgmm_bf16* y, float* psum,
                                         float* psq, int N, int H, int W, void* stream) {
  ECGMM_CHECK(xs && w_s2d && y && psum && psq, ECGMM_ERR_ARG, "stem_conv_fwd_stats: null pointer");
  ECGMM_CHECK(!getenv("ECGMM_STEM_LEGACY"), ECGMM_ERR_SHAPE, "stem_conv_fwd_stats: not offered by the legacy kernel");
  if (N == 0) return ECGMM_OK;
  return launch_stem_fwd_ring(xs, w_s2d, reinterpret_cast<__nv_bfloat16*>(y), N, H, W,
                              static_cast<cudaStream_t>(stream), psum, psq);
}

extern "C" int ecgmm_stem_conv_fwd(const ecgmm_bf16* xs, const ecgmm_bf16* w_s2d, ecgmm_bf16* y, int N, int H,
                                   int W, void* stream) {
  ECGMM_CHECK(xs && w_s2d && y, ECGMM_ERR_ARG, "stem_conv_fwd: null pointer");
  if (N == 0) return ECGMM_OK;
  if (!getenv("ECGMM_STEM_LEGACY"))
    return launch_stem_fwd_ring(xs, w_s2d, reinterpret_cast<__nv_bfloat16*>(y), N, H, W,
                                static_cast<cudaStream_t>(stream));
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  NtParams p;
  memset(&p, 0, sizeof(p));
  set_tile_grid(p, N, Ho, Wo);
  p.ntaps = 4;
  for (int ra = 0; ra < 4; ++ra) p.taps[ra] = Tap{0, (int16_t)ra, 0, (int16_t)ra, ra * 64};
  p.k_chunks = 1;
  int rc = stem_input_map(&p.a_maps[0], xs, N, H, W, p.TW, p.TH);
  if (rc) return rc;
  for (int i = 1; i < 4; ++i) p.a_maps[i] = p.a_maps[0];
  rc = make_tmap_2d(&p.b_map, w_s2d, 256, 64, 512, 64, 64);
  if (rc) return rc;
  p.out = reinterpret_cast<__nv_bfloat16*>(y);
  p.out_sW = 64;
  p.out_sH = (long long)Wo * 64;
  p.out_sN = (long long)Ho * Wo * 64;
  return launch_nt(p, 64, static_cast<cudaStream_t>(stream));
}

extern "C" long long ecgmm_stem_conv_wgrad_workspace(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0 || getenv("ECGMM_STEM_LEGACY")) return 0;
  return (long long)stem_wgrad_ring_workspace_bytes(N, H, W);
}

extern "C" int ecgmm_stem_conv_wgrad(const ecgmm_bf16* xs, const ecgmm_bf16* dy, float* dw, int N, int H, int W,
                                     void* workspace, long long workspace_bytes, void* stream) {
  ECGMM_CHECK(xs && dy && dw, ECGMM_ERR_ARG, "stem_conv_wgrad: null pointer");
  if (N == 0) return ECGMM_OK;
  if (!getenv("ECGMM_STEM_LEGACY"))
    return launch_stem_wgrad_ring(xs, dy, dw, N, H, W, workspace, workspace_bytes > 0 ? (size_t)workspace_bytes : 0,
                                  static_cast<cudaStream_t>(stream));
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  TnParams p;
  memset(&p, 0, sizeof(p));
  pick_tile(Ho, Wo, kKPix, &p.TH, &p.TW);
  p.tiles_h = ceil_div(Ho, p.TH);
  p.tiles_w = ceil_div(Wo, p.TW);
  p.n_img = N;
  p.total_kblocks = N * p.tiles_h * p.tiles_w;
  p.ntaps = 4;
  for (int ra = 0; ra < 4; ++ra) p.taps[ra] = Tap{0, (int16_t)ra, 0, (int16_t)ra, ra * 64};
  p.cin_chunks = 1;
  int rc = stem_input_map(&p.x_maps[0], xs, N, H, W, p.TW, p.TH);
  if (rc) return rc;
  for (int i = 1; i < 4; ++i) p.x_maps[i] = p.x_maps[0];
  rc = make_tmap_4d(&p.dy_map, dy, 64, Wo, Ho, N, 128, (uint64_t)Wo * 128, (uint64_t)Ho * Wo * 128, 64, p.TW, p.TH);
  if (rc) return rc;
  p.dw = dw;
  p.Cin = 3;
  p.Cout = 64;
  p.RS = 49;
  p.mode = 1;
  return launch_tn(p, static_cast<cudaStream_t>(stream));
}
