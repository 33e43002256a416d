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
// MMA shape.  M = 128 rows = two (tap, 64-channel Cin slice) atoms, N = BN output channels, K = 16 pixels.
// BN = 64 is the default; BN = 128 (ECGMM_WG_BN=128) exists and is tested but measured no faster (see make_plan:
// the wider N that doubles the rate of the K-major forward MMAs, profiles/r01_mma_n64_vs_n128.txt, does not pay
// off for these MN-major operands).  A [2 taps x 64] x 128
// fp32 accumulator is 128 TMEM columns and TMEM has 512: four accumulators = 8 of the 9 taps.  The CTAs of a
// 3x3 layer therefore come in two TYPES that share one launch:
//   type A  (cin slice c, cout slice n): taps 0..7 as 4 accumulators                       -> 4 MMA groups per stage
//   type B  (cin slices 2c and 2c+1, cout slice n): tap 8 of BOTH slices as one accumulator -> 1 MMA group per stage
// (no half-empty accumulator: the 9 taps of two cin slices are exactly 9 M = 128 groups) and the pixel range
// (split-K) is divided 4 : 1 between the types so that all CTAs finish together.
// BN = 64 (Cout = 64, or no workspace): one type, ceil(R*S/2) accumulators, the last one half empty for 3x3.
//
// A stage may cover RPS consecutive output rows (RPS dy boxes + RPS+R-1 input rows).  Measured
// (profiles/r01_wgrad_shape_sweep.txt): staging fewer bytes per pixel changes nothing -- the kernel is bound by
// the MMA operand feed, not by L2 -- and short boxes (small KP) hurt, so KP only minimises the pixel slots wasted
// at the right edge of a row.
//
// Split-K partials go to a caller-provided workspace, one [accumulator][BN columns][128 rows] slice per CTA, and
// wgrad_halo_reduce_kernel folds them in a fixed order (deterministic; fp32 atomics only without workspace).
#include "common.h"
#include "ptx.cuh"

#include <stdlib.h>
#include <string.h>

namespace ecgmm {

constexpr int kHaloMaxStages = 6;

struct alignas(64) WgHaloParams {
  CUtensorMap x_map;   // x  [N][H][W][Cin],   box (64, KP+S-1, 1, 1)
  CUtensorMap dy_map;  // dy [N][Ho][Wo][Cout], box (64, KP, 1, 1)
  float* dw;           // [Cout][Cin][R][S]
  float* ws;           // split-K workspace (see launch_wgrad_halo); NULL -> atomics (BN = 64 only)
  int R, S, padH, padW;
  int KP, kmma;              // pixels per stage row, KP/16
  int rps;                   // output rows per stage
  int x_box_bytes, x_box_stride, dy_box_bytes, dy_box_stride, stage_bytes, stages;
  int tiles_w, Ho, row_groups, total_kblocks;  // row_groups = ceil(Ho / rps)
  int cin_chunks, cout_tiles;  // Cin / 64, Cout / BN
  int Cin, Cout;
  // CTA types: blocks [0, nA) are type A (groupsA x ksA), the rest type B (groupsB x ksB); typed == 0: type A only
  int typed, nA, ksA, ksB, cin_pairs, slotsA;
  long long wsB_off;  // float offset of the type-B slices in the workspace
  int tmode;          // wgrad_halo_kernel<128, true> (transposed GEMM, see there)
};

// T = true (the default for 3x3 layers with Cout % 128 == 0 since round 2; ECGMM_WG_T=0 selects the untransposed kernel):
// the TRANSPOSED GEMM
//     dW^T[cout][(r; s, cin)] = sum_pixels dY[pixel][cout] * X[pixel + (r, s)][cin]
// M = 128 output channels (dY is the A operand: two 64-wide atoms, as it is staged for BN = 128), N = 192 = the three
// HORIZONTAL taps of filter row r x a 64-wide Cin slice, all read from ONE staged input-row box as three MN-major atoms
// 128 B (one pixel) apart -- the overlapping-atom addressing this kernel already uses for its M atoms.  Why: an
// M128 x N64 MMA reads 6 KB of operands per 32 tensor clocks (192 B/clk against the 128 B/clk the shared memory
// delivers: capped at 0.67, where the N = 64 kernel is measured, profiles/r01_mma_shape_model.txt); M128 x N192 reads
// 10 KB per 96 clocks (107 B/clk).  An accumulator is 192 TMEM columns, two fit: type A CTAs hold filter rows 0 and 1
// of one Cin slice, type B CTAs hold filter row 2 of a PAIR of Cin slices -- equal MMA work per pixel block, so both
// types get the same split-K count.  Workspace slices are [slot][192 columns][128 rows].
template <int BN, bool T = false>
__global__ void __launch_bounds__(192, 1) wgrad_halo_kernel(const __grid_constant__ WgHaloParams p) {
  static_assert(!T || BN == 128, "the transposed kernel stages dY as two 64-channel atoms");
  constexpr int NCOL = T ? 192 : BN;  // TMEM / workspace columns of one accumulator
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* empty = full + kHaloMaxStages;
  uint64_t* tfull = empty + kHaloMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  constexpr int NB = BN / 64;  // dy boxes (64-channel atoms) per output row
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int RS = p.R * p.S;

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

  // blockIdx -> (type, k-split id, cin slice(s), cout slice)
  const bool typeB = p.typed && (int)blockIdx.x >= p.nA;
  const int bid = typeB ? (int)blockIdx.x - p.nA : (int)blockIdx.x;
  const int ksplit = typeB ? p.ksB : p.ksA;
  const int ks = bid % ksplit;
  const int g = bid / ksplit;
  const int gdiv = typeB ? p.cin_pairs : p.cin_chunks;
  const int cc = typeB ? 2 * (g % gdiv) : (g % gdiv);  // first (or only) 64-channel Cin slice
  const int nt = g / gdiv;                             // BN-channel Cout slice
  const int n_slots = T ? 2 : (typeB ? 1 : p.slotsA);
  const bool pair_ok = typeB && (cc + 1 < p.cin_chunks);  // odd slice count: the last type-B CTA has one slice
  const int per = (p.total_kblocks + ksplit - 1) / ksplit;
  const int kb0 = ks * per;
  const int kb1 = min(p.total_kblocks, kb0 + per);
  const bool has_work = kb0 < kb1;
  // this CTA's workspace slice: [n_slots][NCOL][128]
  float* ws_cta = nullptr;
  if (p.ws)
    ws_cta = typeB ? p.ws + p.wsB_off + (size_t)(g * ksplit + ks) * ((T ? 2 : 1) * NCOL * 128)
                   : p.ws + (size_t)(g * ksplit + ks) * p.slotsA * (NCOL * 128);
  const int n_dy = p.rps * NB;
  const int x_base = n_dy * p.dy_box_stride;
  // input rows staged per Cin slice (transposed type A: filter rows 0 and 1 only)
  const int n_xrows = typeB ? p.rps : (T ? p.rps + 1 : p.rps + p.R - 1);

  if (has_work) {
    if (warp == 0) {
      if (elect_one()) {  // single elected thread: no ELECT serialisation loops around UTMALDG
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = n_dy * p.dy_box_bytes + (typeB ? 2 : 1) * n_xrows * p.x_box_bytes;
        const int row_first = typeB ? p.R - 1 : 0;  // type B only needs the last filter row
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
            for (int a = 0; a < NB; ++a)
              tma_load_4d(st + (j * NB + a) * p.dy_box_stride, &p.dy_map, &full[stage], nt * BN + a * 64, w0, oh0 + j,
                          img);
          for (int r = 0; r < n_xrows; ++r)
            tma_load_4d(st + x_base + r * p.x_box_stride, &p.x_map, &full[stage], cc * 64, w0 - p.padW,
                        oh0 + row_first + r - p.padH, img);
          if (typeB)  // second Cin slice (the first one again when there is none: its rows are ignored)
            for (int r = 0; r < n_xrows; ++r)
              tma_load_4d(st + x_base + (n_xrows + r) * p.x_box_stride, &p.x_map, &full[stage],
                          (pair_ok ? cc + 1 : cc) * 64, w0 - p.padW, oh0 + row_first + r - p.padH, img);
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
        if constexpr (T) {
          constexpr uint32_t idesc_t = make_idesc_bf16(128, 192, 1, 1);
          const uint32_t s_addr = smem_u32(smem);
          // A = dY: 128 output channels = the two 64-wide atoms of a stage row, one dy box apart
          const uint64_t dy_rel = make_sw128_desc(s_addr, p.dy_box_stride, 1024);
          // B = one staged input row: taps s = 0, 1, 2 are three atoms one pixel (128 B) apart.  Slot i is filter row
          // i (type A) or filter row R-1 of Cin slice i (type B: the slices' row sets lie n_xrows boxes apart)
          uint64_t x_rel[2];
#pragma unroll
          for (int i = 0; i < 2; ++i)
            x_rel[i] = make_sw128_desc(s_addr + x_base + i * (typeB ? n_xrows : 1) * p.x_box_stride, 128, 1024);
          int stage = 0;
          uint32_t phase = 0;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t so = (uint64_t)((stage * p.stage_bytes) >> 4);
            for (int j = 0; j < p.rps; ++j) {
              const uint64_t ao = so + (uint64_t)((j * NB * p.dy_box_stride) >> 4);
              const uint64_t bo = so + (uint64_t)((j * p.x_box_stride) >> 4);
#pragma unroll
              for (int i = 0; i < 2; ++i)
                for (int k = 0; k < p.kmma; ++k)
                  umma_bf16(tmem_base + i * 192, dy_rel + ao + k * 128, x_rel[i] + bo + k * 128, idesc_t,
                            (kb > kb0) || (j > 0) || (k > 0));
            }
            umma_commit(&empty[stage]);
            if (kb == kb1 - 1) umma_commit(tfull);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        } else {
        constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
        const uint32_t s_addr = smem_u32(smem);
        // per-accumulator A descriptors relative to the stage base (output row 0 of the stage), built once
        constexpr int kMaxSlots = 8;
        uint64_t a_rel[kMaxSlots];
#pragma unroll
        for (int i = 0; i < kMaxSlots; ++i) {
          uint32_t a0, a1;
          if (typeB) {  // tap RS-1 of the two Cin slices: same position in two box sets n_xrows boxes apart
            a0 = x_base + ((RS - 1) % p.S) * 128;
            a1 = a0 + n_xrows * p.x_box_stride;
          } else {
            const int t0 = min(2 * i, RS - 1), t1 = min(2 * i + 1, RS - 1);
            a0 = x_base + (t0 / p.S) * p.x_box_stride + (t0 % p.S) * 128;
            a1 = x_base + (t1 / p.S) * p.x_box_stride + (t1 % p.S) * 128;
          }
          a_rel[i] = make_sw128_desc(s_addr + a0, a1 - a0, 1024);
        }
        const uint64_t b_rel = make_sw128_desc(s_addr, NB > 1 ? p.dy_box_stride : 0, 1024);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t so = (uint64_t)((stage * p.stage_bytes) >> 4);
          for (int j = 0; j < p.rps; ++j) {  // output row j of the stage: dy boxes j against input rows j .. j+R-1
            const uint64_t ao = so + (uint64_t)((j * p.x_box_stride) >> 4);
            const uint64_t bo = so + (uint64_t)((j * NB * p.dy_box_stride) >> 4);
#pragma unroll
            for (int i = 0; i < kMaxSlots; ++i) {
              if (i < n_slots) {
                const uint64_t a_desc = a_rel[i] + ao, b_desc = b_rel + bo;
                for (int k = 0; k < p.kmma; ++k)  // 16 pixel rows = 2048 B further into both boxes
                  umma_bf16(tmem_base + i * BN, a_desc + k * 128, b_desc + k * 128, idesc,
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
        }  // !T
      }
    } else {
      const int quad = warp & 3;
      const int m_row = quad * 32 + lane;
      mbar_wait(tfull, 0);
      tc_fence_after();
      for (int i = 0; i < n_slots; ++i) {
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + i * NCOL;
        if (ws_cta) {
          // split-K partial, column-major so that the 32 lanes (= rows) of a warp store 128 contiguous bytes
          float* dst0 = ws_cta + (size_t)i * (NCOL * 128) + m_row;
#pragma unroll 1
          for (int c = 0; c < NCOL / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(t_addr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) dst0[(c * 32 + j) * 128] = __uint_as_float(r[j]);
          }
        } else {  // BN = 64, one type (the host never launches BN = 128 without a workspace)
          const int tap = 2 * i + (m_row >> 6);
          const bool ok = tap < RS;
          const int cin = cc * 64 + (m_row & 63);
          float* dst0 = p.dw + ((size_t)(nt * BN) * p.Cin + cin) * RS + (ok ? tap : 0);
          const size_t col_stride = (size_t)p.Cin * RS;
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
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
  } else if (ws_cta && warp >= 2) {
    // a CTA without pixels still owns a workspace slice: the reduction kernel reads every slice
    const int m_row = (warp & 3) * 32 + lane;
    for (int i = 0; i < n_slots; ++i) {
      float* dst0 = ws_cta + (size_t)i * (NCOL * 128) + m_row;
      for (int c = 0; c < NCOL; ++c) dst0[c * 128] = 0.f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Second stage of the split-K reduction: dw[cout][cin][tap] += sum over the ksplit partials of the owning CTAs.
// Thread index = element of one (group, accumulator) result: type-A elements first, then type-B.
struct WgReduceParams {
  const float* ws;
  float* dw;
  int BN, RS, Cin, cin_chunks, cin_pairs;
  int slotsA, ksA, ksB;
  long long totalA, totalB, wsB_off;
  int S;  // transposed layout only (wgrad_halo_reduce_t_kernel); appended: the kernels below read the same offsets as before
};

// A CTA folds 64 consecutive result elements; its 256 threads are 64 elements x 4 interleaved k-groups (k = kg, kg+4,
// ...), each with 4 independent accumulators: 16 partial loads in flight per element instead of 4 (the fold of the
// 148 x 160 KB layer1 partials took 29 us, ~0.8 TB/s).  Fixed summation order -> deterministic.
__global__ void __launch_bounds__(256) wgrad_halo_reduce_kernel(const WgReduceParams p) {
  __shared__ float red[4][64];
  const int o = threadIdx.x & 63, kg = threadIdx.x >> 6;
  long long idx = blockIdx.x * 64LL + o;
  const bool typeB = idx >= p.totalA;  // totalA is a multiple of 64: a CTA never straddles the two regions
  if (typeB) idx -= p.totalA;
  const bool in_range = !typeB || idx < p.totalB;
  const int m = (int)(idx & 127);
  const long long rest = idx >> 7;
  const int c = (int)(rest % p.BN);
  const long long gs = rest / p.BN;
  const int n_slots = typeB ? 1 : p.slotsA;
  const int slot = (int)(gs % n_slots);
  const int g = (int)(gs / n_slots);
  const int ksplit = typeB ? p.ksB : p.ksA;
  int tap, chunk, nt;
  bool live = in_range;
  if (typeB) {
    tap = p.RS - 1;
    chunk = 2 * (g % p.cin_pairs) + (m >> 6);
    nt = g / p.cin_pairs;
    live = live && chunk < p.cin_chunks;
  } else {
    tap = 2 * slot + (m >> 6);
    chunk = g % p.cin_chunks;
    nt = g / p.cin_chunks;
    live = live && tap < p.RS;
  }
  float acc = 0.f;
  if (live) {
    const size_t slice = (size_t)n_slots * p.BN * 128;  // one CTA of the producer kernel
    const float* src =
        p.ws + (typeB ? p.wsB_off : 0) + (size_t)g * ksplit * slice + ((size_t)slot * p.BN + c) * 128 + m;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = kg;
    for (; k + 12 < ksplit; k += 16) {
      a0 += src[(size_t)k * slice];
      a1 += src[(size_t)(k + 4) * slice];
      a2 += src[(size_t)(k + 8) * slice];
      a3 += src[(size_t)(k + 12) * slice];
    }
    for (; k < ksplit; k += 4) a0 += src[(size_t)k * slice];
    acc = (a0 + a1) + (a2 + a3);
  }
  red[kg][o] = acc;
  __syncthreads();
  if (kg == 0 && live) {
    const float total = (red[0][o] + red[1][o]) + (red[2][o] + red[3][o]);
    const int cin = chunk * 64 + (m & 63), cout = nt * p.BN + c;
    p.dw[((size_t)cout * p.Cin + cin) * p.RS + tap] += total;
  }
}

// One thread per result element (few partials per element: layers with many (cin, cout) slice pairs).
__global__ void __launch_bounds__(256) wgrad_halo_reduce_flat_kernel(const WgReduceParams p) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool typeB = idx >= p.totalA;
  if (typeB) idx -= p.totalA;
  if (typeB && idx >= p.totalB) return;
  const int m = (int)(idx & 127);
  const long long rest = idx >> 7;
  const int c = (int)(rest % p.BN);
  const long long gs = rest / p.BN;
  const int n_slots = typeB ? 1 : p.slotsA;
  const int slot = (int)(gs % n_slots);
  const int g = (int)(gs / n_slots);
  const int ksplit = typeB ? p.ksB : p.ksA;
  int tap, chunk, nt;
  if (typeB) {
    tap = p.RS - 1;
    chunk = 2 * (g % p.cin_pairs) + (m >> 6);
    nt = g / p.cin_pairs;
    if (chunk >= p.cin_chunks) return;
  } else {
    tap = 2 * slot + (m >> 6);
    chunk = g % p.cin_chunks;
    nt = g / p.cin_chunks;
    if (tap >= p.RS) return;
  }
  const size_t slice = (size_t)n_slots * p.BN * 128;
  const float* src = p.ws + (typeB ? p.wsB_off : 0) + (size_t)g * ksplit * slice + ((size_t)slot * p.BN + c) * 128 + m;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int k = 0;
  for (; k + 3 < ksplit; k += 4) {
    a0 += src[(size_t)k * slice];
    a1 += src[(size_t)(k + 1) * slice];
    a2 += src[(size_t)(k + 2) * slice];
    a3 += src[(size_t)(k + 3) * slice];
  }
  for (; k < ksplit; ++k) a0 += src[(size_t)k * slice];
  const float acc = (a0 + a1) + (a2 + a3);
  const int cin = chunk * 64 + (m & 63), cout = nt * p.BN + c;
  p.dw[((size_t)cout * p.Cin + cin) * p.RS + tap] += acc;
}

// Reduction of the TRANSPOSED kernel's partials (wgrad_halo_kernel<128, true>, see there): workspace
// slices [2 slots][192 columns][128 rows]; row m = output channel, column c = (s, cin); type A slot = filter row r of
// Cin slice g % cin_chunks, type B slot = Cin slice 2 (g % cin_pairs) + slot with r = R - 1.  One thread per element.
// (A separate kernel so that the two reducers above stay exactly the code that has run on hardware.)
__global__ void __launch_bounds__(256) wgrad_halo_reduce_t_kernel(const WgReduceParams p) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool typeB = idx >= p.totalA;
  if (typeB) idx -= p.totalA;
  if (typeB && idx >= p.totalB) return;
  const int m = (int)(idx & 127);
  const long long rest = idx >> 7;
  const int c = (int)(rest % 192);
  const long long gs = rest / 192;
  const int slot = (int)(gs % 2);
  const int g = (int)(gs / 2);
  const int ksplit = typeB ? p.ksB : p.ksA;
  int r, chunk, nt;
  if (typeB) {
    r = p.RS / p.S - 1;
    chunk = 2 * (g % p.cin_pairs) + slot;
    nt = g / p.cin_pairs;
    if (chunk >= p.cin_chunks) return;
  } else {
    r = slot;
    chunk = g % p.cin_chunks;
    nt = g / p.cin_chunks;
  }
  const int tap = r * p.S + (c >> 6);
  const int cin = chunk * 64 + (c & 63), cout = nt * p.BN + m;
  const size_t slice = (size_t)2 * 192 * 128;
  const float* src = p.ws + (typeB ? p.wsB_off : 0) + (size_t)g * ksplit * slice + ((size_t)slot * 192 + c) * 128 + m;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int k = 0;
  for (; k + 3 < ksplit; k += 4) {
    a0 += src[(size_t)k * slice];
    a1 += src[(size_t)(k + 1) * slice];
    a2 += src[(size_t)(k + 2) * slice];
    a3 += src[(size_t)(k + 3) * slice];
  }
  for (; k < ksplit; ++k) a0 += src[(size_t)k * slice];
  p.dw[((size_t)cout * p.Cin + cin) * p.RS + tap] += (a0 + a1) + (a2 + a3);
}

// The same fold with COALESCED writes (3x3 only).  The kernel above is one thread per workspace element with the output
// channel fastest: neighbouring threads update dw elements Cin * 9 floats apart, a read-modify-write of a 32-byte sector
// per 4 useful bytes -- 50 us per launch at any batch size (profiles/r02_ncu_launches_gb64.txt: 2.6 % of the batch-64
// step).  Here a CTA owns 16 output channels x one 64-wide Cin slice, gathers the three filter rows of it (type A slots
// 0 / 1 of group (nt, chunk), type B slot chunk % 2 of group (nt, chunk / 2)) into a [16][64 cin x 9 taps] tile in
// shared memory in dw's own order, and adds 576 contiguous floats per output channel.  Same summation order per
// element (k-split index ascending in four interleaved accumulators) -> bit-identical results.
constexpr int kRtM = 16;
constexpr int kRtCinSplit = 4;                  // a CTA folds 16 output channels x 16 input channels x 9 taps
constexpr int kRtCin = 64 / kRtCinSplit;
__global__ void __launch_bounds__(256) wgrad_halo_reduce_t3_kernel(const WgReduceParams p) {
  // (a 64-wide Cin slice per CTA meant 256 CTAs for the 512-channel layers: 65 us for a 9 MB result; split 4 ways)
  __shared__ float tile[kRtM][kRtCin * 9 + 1];
  int b = blockIdx.x;
  const int cs = b % kRtCinSplit;
  b /= kRtCinSplit;
  const int mb = b % (128 / kRtM);
  const int gc = b / (128 / kRtM);  // (nt, chunk)
  const int chunk = gc % p.cin_chunks, nt = gc / p.cin_chunks;
  const int mi = threadIdx.x & (kRtM - 1), c0 = threadIdx.x / kRtM;  // 16 x 16
  const size_t slice = (size_t)2 * 192 * 128;
  const int m = mb * kRtM + mi;
  for (int r = 0; r < 3; ++r) {
    const bool typeB = (r == 2);
    const int g = typeB ? nt * p.cin_pairs + (chunk >> 1) : nt * p.cin_chunks + chunk;
    const int slot = typeB ? (chunk & 1) : r;
    const int ksplit = typeB ? p.ksB : p.ksA;
    const float* base = p.ws + (typeB ? p.wsB_off : 0) + (size_t)g * ksplit * slice + (size_t)slot * 192 * 128 + m;
    for (int cc = c0; cc < 3 * kRtCin; cc += 256 / kRtM) {  // column c = s * 64 + cin, cin in this CTA's quarter
      const int sft = cc / kRtCin, ci = cc - sft * kRtCin;
      const int c = sft * 64 + cs * kRtCin + ci;
      const float* src = base + (size_t)c * 128;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int k = 0;
      for (; k + 3 < ksplit; k += 4) {
        a0 += src[(size_t)k * slice];
        a1 += src[(size_t)(k + 1) * slice];
        a2 += src[(size_t)(k + 2) * slice];
        a3 += src[(size_t)(k + 3) * slice];
      }
      for (; k < ksplit; ++k) a0 += src[(size_t)k * slice];
      tile[mi][ci * 9 + r * 3 + sft] = (a0 + a1) + (a2 + a3);
    }
  }
  __syncthreads();
  for (int row = 0; row < kRtM; ++row) {
    float* dst = p.dw + ((size_t)(nt * 128 + mb * kRtM + row) * p.Cin + chunk * 64 + cs * kRtCin) * 9;
    for (int j = threadIdx.x; j < kRtCin * 9; j += 256) dst[j] += tile[row][j];
  }
}

// KP (multiple of 16, <= 128) minimises the pixel slots wasted at the right edge of a row (ties: larger KP); RPS = 2
// only where three stages still fit (measured: a small, free reduction of L2 traffic, no speed-up by itself).
static int round1k(int v) { return (v + 1023) & ~1023; }
static int halo_stage_bytes(int kp, int rps, int R, int S, int bn, bool typed, bool tmode = false) {
  const int xb = round1k((kp + S - 1) * 128);
  const int a = (tmode ? rps + 1 : rps + R - 1) * xb, b = typed ? 2 * rps * xb : 0;
  return rps * (bn / 64) * round1k(kp * 128) + (a > b ? a : b);
}
static void pick_shape(int Ho, int Wo, int R, int S, int bn, bool typed, int* kp_out, int* rps_out) {
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
  if (bn == 64 && Ho >= 2 && 3 * halo_stage_bytes(best_kp, 2, R, S, bn, typed) <= budget) best_rps = 2;
  *kp_out = best_kp;
  *rps_out = best_rps;
  // development knobs (tools/conv_bench.py sweeps): force a shape
  const char* ekp = getenv("ECGMM_WG_KP");
  const char* erps = getenv("ECGMM_WG_RPS");
  if (ekp && erps) {
    const int kp = atoi(ekp), rps = atoi(erps);
    if (kp >= 16 && kp <= 128 && kp % 16 == 0 && rps >= 1 && rps <= (Ho > 1 ? Ho : 1) &&
        2 * halo_stage_bytes(kp, rps, R, S, bn, typed) <= budget) {
      *kp_out = kp;
      *rps_out = rps;
    }
  }
}

bool wgrad_halo_supported(int Cin, int Cout, int R, int S, int stride) {
  return stride == 1 && Cin % 64 == 0 && Cout % 64 == 0 && R * S > 1 && R * S <= 15 && (R == 1 || R == 3) &&
         (S == 3);
}

// The whole decomposition of one problem; shared by the workspace query and the launch.
struct WgPlan {
  int bn, typed, tmode, slotsA, KP, rps, total_kb;
  int cin_chunks, cin_pairs, cout_tiles, groupsA, groupsB, ksA, ksB;
  size_t ws_floats, wsB_off;
};

static WgPlan make_plan(int N, int H, int W, int Cin, int Cout, int R, int S, int padH, int padW, bool have_ws) {
  WgPlan q;
  memset(&q, 0, sizeof(q));
  const int Ho = H + 2 * padH - R + 1, Wo = W + 2 * padW - S + 1;
  const int RS = R * S;
  // Measured (profiles/r01_wgrad_bn128.txt, batch 64): BN = 128 with the two CTA types does NOT beat BN = 64 on
  // layers 2 and 3 (0.260 / 0.243 ms against 0.217 / 0.228 ms) -- unlike the K-major forward kernels, the MN-major
  // operand feed does not get faster with the wider N.  On layer4 it wins with a cold L2 (0.257 against 0.281 ms)
  // but not inside the training step (0.77-0.81 ms per 3 launches against 0.74-0.76 ms), so BN = 64 is the default
  // everywhere; ECGMM_WG_BN=128 selects the two-type kernel (kept under test: tests/test_conv_gpu.py).
  const char* ebn = getenv("ECGMM_WG_BN");
  const bool want128 = ebn && atoi(ebn) == 128;
  // Transposed GEMM (wgrad_halo_kernel<128, true>: M = 128 output channels, N = 192 = three horizontal taps x 64
  // input channels) for every 3x3 layer with Cout % 128 == 0 when a workspace is given: measured on a B200 inside the
  // training step (profiles/r02a_*): layer2 1000 -> 1443, layer3 ~780 -> 1555, layer4 ~610 -> 1542 TFLOP/s at batch 512.
  // ECGMM_WG_T=0 selects the M = 128 (two taps) x N = 64 kernel again (kept under test).
  const char* et = getenv("ECGMM_WG_T");
  q.tmode = (!(et && atoi(et) == 0) && !want128 && R == 3 && S == 3 && Cout % 128 == 0 && have_ws) ? 1 : 0;
  q.bn = (Cout % 128 == 0 && have_ws && (want128 || q.tmode)) ? 128 : 64;
  const int slots_all = (RS + 1) / 2;
  q.typed = (q.bn == 128 && slots_all * q.bn > 512) ? 1 : 0;  // does not fit the 512 TMEM columns -> two CTA types
  q.slotsA = q.typed ? 512 / q.bn : slots_all;                // typed: 4 accumulators = taps 0..7, type B: tap 8
  if (q.tmode) {
    q.typed = 1;
    q.slotsA = 2;  // filter rows 0 and 1; type B: filter row 2 of a pair of Cin slices (also 2 accumulators)
  }
  pick_shape(Ho, Wo, R, S, q.bn, q.typed != 0, &q.KP, &q.rps);
  if (q.tmode) q.rps = 1;
  q.total_kb = N * ceil_div(Ho, q.rps) * ceil_div(Wo, q.KP);
  q.cin_chunks = Cin / 64;
  q.cin_pairs = (q.cin_chunks + 1) / 2;
  q.cout_tiles = Cout / q.bn;
  q.groupsA = q.cin_chunks * q.cout_tiles;
  q.groupsB = q.typed ? q.cin_pairs * q.cout_tiles : 0;
  const int sms = num_sms();
  if (q.tmode) {
    // two MMA groups per pixel block in both types: same split-K count
    q.ksA = q.ksB = sms / (q.groupsA + q.groupsB);
    if (q.ksA < 1) q.ksA = q.ksB = 1;
  } else if (q.typed) {
    // MMA groups per pixel block: slotsA per type-A group, 1 per type-B group; CTAs in proportion
    const double per_unit = (double)sms / (double)(q.groupsA * q.slotsA + q.groupsB);
    q.ksA = (int)(per_unit * q.slotsA);
    q.ksB = (int)per_unit;
    if (q.ksA < 1) q.ksA = 1;
    if (q.ksB < 1) q.ksB = 1;
  } else {
    q.ksA = sms / q.groupsA;
    if (q.ksA < 1) q.ksA = 1;
    q.ksB = 0;
  }
  const int cap = q.total_kb > 0 ? q.total_kb : 1;
  if (q.ksA > cap) q.ksA = cap;
  if (q.ksB > cap) q.ksB = cap;
  const size_t ncol = q.tmode ? 192 : q.bn, slotsB = q.tmode ? 2 : 1;
  q.wsB_off = (size_t)q.groupsA * q.ksA * q.slotsA * ncol * 128;
  q.ws_floats = q.wsB_off + (size_t)q.groupsB * q.ksB * slotsB * ncol * 128;
  return q;
}

size_t wgrad_halo_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S, int padH, int padW) {
  const WgPlan q = make_plan(N, H, W, Cin, Cout, R, S, padH, padW, true);
  return q.ws_floats * sizeof(float);
}

template <int BN, bool T = false>
static int launch_typed(const WgHaloParams& p, int grid, int smem, cudaStream_t st) {
  static bool configured[kMaxDevices] = {};  // the shared-memory limit of a kernel is a per-device attribute
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel<BN, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    configured[ds] = true;
  }
  wgrad_halo_kernel<BN, T><<<grid, 192, smem, st>>>(p);
  return check_launch("wgrad_halo_kernel");
}

int launch_wgrad_halo(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, int N, int H, int W, int Cin,
                      int Cout, int R, int S, int padH, int padW, void* workspace, size_t ws_bytes,
                      cudaStream_t st) {
  const int Ho = H + 2 * padH - R + 1, Wo = W + 2 * padW - S + 1;
  // the caller sized the workspace with wgrad_halo_workspace_bytes (the BN = 128 plan when Cout allows it); anything
  // smaller means "no workspace" and the BN = 64 kernel with atomics
  const bool have_ws = workspace && ws_bytes >= wgrad_halo_workspace_bytes(N, H, W, Cin, Cout, R, S, padH, padW);
  const WgPlan q = make_plan(N, H, W, Cin, Cout, R, S, padH, padW, have_ws);
  WgHaloParams p;
  memset(&p, 0, sizeof(p));
  p.R = R;
  p.S = S;
  p.padH = padH;
  p.padW = padW;
  p.KP = q.KP;
  p.rps = q.rps;
  p.kmma = p.KP / 16;
  const int xw = p.KP + S - 1;
  p.x_box_bytes = xw * 128;
  p.x_box_stride = round1k(p.x_box_bytes);
  p.dy_box_bytes = p.KP * 128;
  p.dy_box_stride = round1k(p.dy_box_bytes);
  p.stage_bytes = halo_stage_bytes(p.KP, p.rps, R, S, q.bn, q.typed != 0, q.tmode != 0);
  p.tmode = q.tmode;
  int stages = (220 * 1024) / p.stage_bytes;
  if (stages > kHaloMaxStages) stages = kHaloMaxStages;
  ECGMM_CHECK(stages >= 2, ECGMM_ERR_SHAPE, "wgrad_halo: stage of %d bytes does not fit twice", p.stage_bytes);
  p.stages = stages;
  p.tiles_w = ceil_div(Wo, p.KP);
  p.Ho = Ho;
  p.row_groups = ceil_div(Ho, p.rps);
  p.total_kblocks = q.total_kb;
  p.cin_chunks = q.cin_chunks;
  p.cout_tiles = q.cout_tiles;
  p.Cin = Cin;
  p.Cout = Cout;
  p.dw = dw;
  p.typed = q.typed;
  p.nA = q.groupsA * q.ksA;
  p.ksA = q.ksA;
  p.ksB = q.ksB > 0 ? q.ksB : 1;
  p.cin_pairs = q.cin_pairs;
  p.slotsA = q.slotsA;
  p.wsB_off = (long long)q.wsB_off;
  p.ws = have_ws ? reinterpret_cast<float*>(workspace) : nullptr;
  const uint64_t e = 2;
  int rc = make_tmap_4d(&p.x_map, x, Cin, W, H, N, (uint64_t)Cin * e, (uint64_t)W * Cin * e, (uint64_t)H * W * Cin * e,
                        64, xw, 1);
  if (rc) return rc;
  rc = make_tmap_4d(&p.dy_map, dy, Cout, Wo, Ho, N, (uint64_t)Cout * e, (uint64_t)Wo * Cout * e,
                    (uint64_t)Ho * Wo * Cout * e, 64, p.KP, 1);
  if (rc) return rc;
  const int smem = p.stages * p.stage_bytes + 256 + 1024;
  const int grid = q.groupsA * q.ksA + q.groupsB * q.ksB;
  rc = q.tmode ? launch_typed<128, true>(p, grid, smem, st)
               : (q.bn == 128 ? launch_typed<128>(p, grid, smem, st) : launch_typed<64>(p, grid, smem, st));
  if (rc || !p.ws) return rc;
  WgReduceParams r;
  memset(&r, 0, sizeof(r));
  r.ws = p.ws;
  r.dw = dw;
  r.BN = q.bn;
  r.RS = R * S;
  r.Cin = Cin;
  r.cin_chunks = q.cin_chunks;
  r.cin_pairs = q.cin_pairs;
  r.slotsA = q.slotsA;
  r.ksA = q.ksA;
  r.ksB = p.ksB;
  r.S = S;
  const long long ncol = q.tmode ? 192 : q.bn, slotsB = q.tmode ? 2 : 1;
  r.totalA = (long long)q.groupsA * q.slotsA * ncol * 128;
  r.totalB = (long long)q.groupsB * slotsB * ncol * 128;
  r.wsB_off = (long long)q.wsB_off;
  const long long total = r.totalA + r.totalB;  // both multiples of 64 (BN * 128 elements per accumulator)
  // (measured at batch 64: with few partials per element the scattered writes dominate and the tiled fold wins --
  //  512 channels, k-split 3: 0.638 -> 0.553 ms per 3 launches -- with many partials the one-thread-per-element kernel's
  //  parallelism wins: 128 channels, k-split 49: 0.420 vs 0.839 ms.  With the tiled fold's CTAs split 4 ways along Cin
  //  the crossover moved: whole step at batch 64 with the threshold at 0 / 6 / 12 / 24 / 64 partials:
  //  16.12 / 16.02 / 15.89 / 15.87 / 15.92 ms)
  const char* ert = getenv("ECGMM_WG_REDUCE_T_LEGACY");
  const char* emk = getenv("ECGMM_WG_REDUCE_T_MAXKS");
  const bool tiled = ert ? atoi(ert) == 0 : (q.ksA <= (emk ? atoi(emk) : 24));
  if (q.tmode && R == 3 && S == 3 && tiled)
    wgrad_halo_reduce_t3_kernel<<<(unsigned)(q.cout_tiles * q.cin_chunks * (128 / kRtM) * kRtCinSplit), 256, 0, st>>>(r);
  else if (q.tmode)
    wgrad_halo_reduce_t_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(r);
  else if (q.ksA >= 16)  // many partials per element: 4 k-groups per element
    wgrad_halo_reduce_kernel<<<(unsigned)((total + 63) / 64), 256, 0, st>>>(r);
  else
    wgrad_halo_reduce_flat_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(r);
  return check_launch("wgrad_halo_reduce_kernel");
}

}  // namespace ecgmm
