// Batched perturbation inference of the fusion head in ONE kernel (BASELINE.json configs[3]):
//     prob[s][v] = softmax( W2 relu( W1 (z[v] * e[s] + (1 - z[v]) * b) + b1 ) + b2 )[class]
// for V masked variants of every sample's fused embedding e[s] (shap_fusion_modal_balance.py:126-159,
// lime_fusion_modal_balance.py:126-131 drive fusion_classifier row by row through a wrapper).
//
// The three-kernel path of perturb.cu writes the S*V x D variants to HBM as bf16, reads them back in the GEMM and
// round-trips the hidden layer once more (head_tail).  Here the variants never exist outside shared memory:
//   warps 2..9  (producers) build the A operand of the GEMM directly in its SWIZZLE_128B K-major shared-memory layout.
//                           The masks are shared by all samples: they are packed once per call to one BIT per element
//                           (ecgmm_perturb_pack_masks: V x D / 8 bytes, L1/L2-resident) in an order chosen so that one
//                           shift puts four elements' bits on the sign bits of a register's four bytes and PRMT's
//                           sign-replicate mode expands them to the 16-bit select masks of two bf16 pairs.  The
//                           selection between the bf16 bit patterns of e[s] and b is exact (no arithmetic).  Two
//                           groups of four warps alternate over the K chunks (one group per stage of the ring) with
//                           the mask words prefetched two chunks ahead, so the L2 latency never sits on the ring;
//   warp 1                  tcgen05.mma M128 x N128 x K16 against W1, which stays RESIDENT in shared memory for the
//                           whole persistent CTA (D <= 768: 192 KB), accumulators double-buffered in TMEM;
//   warps 10..13 (epilogue) TMEM -> + b1 -> ReLU -> Linear(128, C) -> softmax, fp32, one thread per variant row.
// HBM traffic per variant: 4 (or 4 C) bytes out.  Bound: shared-memory bandwidth (MMA operand reads 128 B/clk + producer
// writes 64 B/clk of the 128 B/clk an SM moves), i.e. ~2/3 of the tensor pipe.
#include "common.h"
#include "ptx.cuh"

#include <string.h>

namespace ecgmm {

constexpr int kPfTile = 128;            // variants per tile (GEMM M)
constexpr int kPfHid = 128;             // hidden width (GEMM N)
constexpr int kPfChunk = 128 * 128;     // one 64-wide K chunk of an operand tile: 128 rows x 128 B
constexpr int kPfThreads = 448;         // TMA, MMA, 2 x 4 producer warps, 4 epilogue warps
constexpr int kPfMaxC = 8;

// prmt.b32 in its default mode: selector nibble 8 + i replicates the SIGN of byte i of `a` over the result byte
// (__byte_perm only honours the low three bits of each nibble).
__device__ __forceinline__ uint32_t prmt_sign(uint32_t a, uint32_t selector) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(selector));
  return d;
}

struct alignas(64) PerturbFusedParams {
  CUtensorMap w_map;           // W1 bf16 [128][D] (k contiguous), box (64, 128)
  const __nv_bfloat16* e;      // [S][D]
  const __nv_bfloat16* bg;     // [D]
  const uint32_t* bits;        // [V][D / 32] packed keep-masks (perturb_pack_masks_kernel's bit order)
  const float* b1;             // [128]
  const float* w2;             // [C][128]
  const float* b2;             // [C]
  float* out;                  // [S][V] (cls >= 0) or [S][V][C]
  int S, V, D, C, cls, kchunks, tiles_v, total_tiles;
};

template <int MAXC>  // 2: the binary head of the reference (the loops over classes unroll without predicates); 8: generic
__global__ void __launch_bounds__(kPfThreads, 1) perturb_fused_kernel(const __grid_constant__ PerturbFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sW = smem;                               // kchunks x 16 KB, resident
  uint8_t* sA = smem + p.kchunks * kPfChunk;        // 2 stages x 16 KB
  uint64_t* afull = reinterpret_cast<uint64_t*>(sA + 2 * kPfChunk);
  uint64_t* aempty = afull + 2;
  uint64_t* tfull = aempty + 2;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.w_map);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&afull[i], 4);   // one arrival per producer warp
      mbar_init(&aempty[i], 1);  // tcgen05.commit
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);  // one arrival per epilogue warp
    }
    mbar_init(wfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * kPfHid);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {  // W1 once per CTA
      mbar_expect_tx(wfull, p.kchunks * kPfChunk);
      for (int kc = 0; kc < p.kchunks; ++kc) tma_load_2d(sW + kc * kPfChunk, &p.w_map, wfull, kc * 64, 0);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, kPfHid, 0, 0);
      const uint64_t w_desc0 = make_sw128_desc(smem_u32(sW), 0, 1024);
      const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA), 0, 1024);
      mbar_wait(wfull, 0);
      tc_fence_after();
      uint32_t i = 0;  // running K-chunk counter: stage = i & 1, parity = (i >> 1) & 1
      int it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kPfHid;
        for (int kc = 0; kc < p.kchunks; ++kc, ++i) {
          const uint32_t stage = i & 1u;
          mbar_wait(&afull[stage], (i >> 1) & 1u);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + (uint64_t)(stage * (kPfChunk >> 4));
          const uint64_t w_desc = w_desc0 + (uint64_t)(kc * (kPfChunk >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_desc + 2 * k, w_desc + 2 * k, idesc, (kc | k) != 0);
          umma_commit(&aempty[stage]);
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ producers: the masked variants, straight into smem
    // Group g (warps 2..5 / 6..9) fills ring stage g: the running K-chunk counter i = it * kchunks + kc with i & 1 == g.
    const int g = (warp - 2) >> 2;
    const int row = ((warp - 2) & 3) * 32 + lane;
    const int words = p.D >> 5;
    const uint4* brow = reinterpret_cast<const uint4*>(p.bg);
    // load cursor (2 chunks of this group ahead of the build cursor)
    int lt = blockIdx.x, lkc = g;
    auto norm = [&](int& t, int& kc) {
      while (kc >= p.kchunks && t < p.total_tiles) {
        kc -= p.kchunks;
        t += gridDim.x;
      }
    };
    auto load_bits = [&](int t, int kc) -> uint2 {
      if (t >= p.total_tiles) return make_uint2(0u, 0u);
      const int s = t / p.tiles_v;
      const int v = (t - s * p.tiles_v) * kPfTile + row;
      return __ldg(reinterpret_cast<const uint2*>(p.bits + (size_t)(v < p.V ? v : 0) * words + kc * 2));
    };
    norm(lt, lkc);
    uint2 m0 = load_bits(lt, lkc);
    lkc += 2;
    norm(lt, lkc);
    uint2 m1 = load_bits(lt, lkc);
    lkc += 2;
    norm(lt, lkc);
    int t = blockIdx.x, kc = g;
    norm(t, kc);
    uint32_t n = 0;  // chunks this group has built: stage g is on its n-th use
    while (t < p.total_tiles) {
      const uint2 m = m0;
      m0 = m1;
      m1 = load_bits(lt, lkc);
      lkc += 2;
      norm(lt, lkc);
      const int s = t / p.tiles_v;
      const uint4* erow = reinterpret_cast<const uint4*>(p.e + (size_t)s * p.D);
      mbar_wait(&aempty[g], (n & 1u) ^ 1u);
      uint8_t* dst = sA + g * kPfChunk + row * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // 16-byte chunk j = elements 8j .. 8j+7 of the K chunk
        const uint4 ev = __ldg(erow + kc * 8 + j), bv = __ldg(brow + kc * 8 + j);  // same address in every lane
        const uint32_t w = (j < 4) ? m.x : m.y;
        const uint32_t r0 = w << (2 * (j & 3)), r1 = w << (2 * (j & 3) + 1);  // sign bits of bytes 0..3 = 4 elements
        const uint32_t s0 = prmt_sign(r0, 0x9988u), s1 = prmt_sign(r0, 0xBBAAu);
        const uint32_t s2 = prmt_sign(r1, 0x9988u), s3 = prmt_sign(r1, 0xBBAAu);
        uint4 o;
        o.x = (ev.x & s0) | (bv.x & ~s0);
        o.y = (ev.y & s1) | (bv.y & ~s1);
        o.z = (ev.z & s2) | (bv.z & ~s2);
        o.w = (ev.w & s3) | (bv.w & ~s3);
        *reinterpret_cast<uint4*>(dst + ((j ^ (row & 7)) << 4)) = o;
      }
      fence_proxy_async_smem();  // the tensor core reads shared memory through the async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&afull[g]);
      ++n;
      kc += 2;
      norm(t, kc);
    }
  } else {
    // ------------------------------------------------------------ epilogue: bias, ReLU, Linear(128, C), softmax
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int s = t / p.tiles_v;
      const int v = (t - s * p.tiles_v) * kPfTile + row;
      const int acc = it & 1;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kPfHid;
      float logit[MAXC];
#pragma unroll
      for (int c = 0; c < MAXC; ++c) logit[c] = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < kPfHid / 32; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32(t_addr + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b1 + cc * 32 + j));
          const float h0 = fmaxf(__uint_as_float(r[j]) + bb.x, 0.f), h1 = fmaxf(__uint_as_float(r[j + 1]) + bb.y, 0.f);
          const float h2 = fmaxf(__uint_as_float(r[j + 2]) + bb.z, 0.f), h3 = fmaxf(__uint_as_float(r[j + 3]) + bb.w, 0.f);
#pragma unroll
          for (int c = 0; c < MAXC; ++c) {
            if (c < p.C) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(p.w2 + c * kPfHid + cc * 32 + j));
              logit[c] = fmaf(h0, w.x, fmaf(h1, w.y, fmaf(h2, w.z, fmaf(h3, w.w, logit[c]))));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (v < p.V) {
        const size_t o = (size_t)s * p.V + v;
        if (p.cls < 0) {
#pragma unroll
          for (int c = 0; c < MAXC; ++c)
            if (c < p.C) p.out[o * p.C + c] = logit[c] + __ldg(p.b2 + c);
        } else {
          float mx = -INFINITY, den = 0.f, num = 0.f;
#pragma unroll
          for (int c = 0; c < MAXC; ++c)
            if (c < p.C) {
              logit[c] += __ldg(p.b2 + c);
              mx = fmaxf(mx, logit[c]);
            }
#pragma unroll
          for (int c = 0; c < MAXC; ++c)
            if (c < p.C) {
              const float ex = __expf(logit[c] - mx);
              den += ex;
              if (c == p.cls) num = ex;
            }
          p.out[o] = num / den;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * kPfHid);
}

// masks [V][D] bytes (nonzero = keep) -> one bit per element, 32 elements per word: element 4k + b of a word sits on bit
// 8b + 7 - k, so `word << k` carries elements 4k .. 4k+3 on the sign bits of its four bytes (see the producers above).
__global__ void perturb_pack_masks_kernel(const uint8_t* __restrict__ masks, uint32_t* __restrict__ bits, size_t nwords) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(masks + i * 32);
    const uint4 lo = __ldg(src), hi = __ldg(src + 1);
    const uint32_t q[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};  // q[k] = elements 4k .. 4k+3
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) w |= (__vcmpne4(q[k], 0u) & 0x80808080u) >> k;
    bits[i] = w;
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}

}  // namespace ecgmm

using namespace ecgmm;

extern "C" int ecgmm_f32_to_bf16(const float* x, ecgmm_bf16* y, long long n, void* stream) {
  ECGMM_CHECK(x && y, ECGMM_ERR_ARG, "f32_to_bf16: null pointer");
  if (n <= 0) return ECGMM_OK;
  size_t blocks = ((size_t)n + 255) / 256;
  if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
  f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), (size_t)n);
  return check_launch("f32_to_bf16_kernel");
}

// 1 when ecgmm_perturb_head_fused covers the shape: hidden width 128, D a multiple of 64 and <= 768 (W1 resident in
// shared memory), at most 8 classes.
extern "C" int ecgmm_perturb_head_fused_supported(int D, int HID, int C) {
  return (HID == kPfHid && D > 0 && D % 64 == 0 && D <= 768 && C >= 1 && C <= kPfMaxC) ? 1 : 0;
}

extern "C" int ecgmm_perturb_pack_masks(const uint8_t* masks, uint32_t* bits, int V, int D, void* stream) {
  ECGMM_CHECK(masks && bits, ECGMM_ERR_ARG, "perturb_pack_masks: null pointer");
  ECGMM_CHECK(V >= 0 && D > 0 && D % 32 == 0, ECGMM_ERR_SHAPE, "perturb_pack_masks: D=%d must be a positive multiple of 32", D);
  ECGMM_CHECK((reinterpret_cast<uintptr_t>(masks) & 15) == 0, ECGMM_ERR_ALIGN, "perturb_pack_masks: masks must be 16-byte aligned");
  const size_t nwords = (size_t)V * (D / 32);
  if (nwords == 0) return ECGMM_OK;
  size_t blocks = (nwords + 255) / 256;
  if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
  perturb_pack_masks_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(masks, bits, nwords);
  return check_launch("perturb_pack_masks_kernel");
}

extern "C" int ecgmm_perturb_head_fused(const ecgmm_bf16* e, const ecgmm_bf16* bg, const uint32_t* bits,
                                        const ecgmm_bf16* w1, const float* b1, const float* w2, const float* b2,
                                        float* out, long long S, int V, int D, int C, int cls, void* stream) {
  ECGMM_CHECK(e && bg && bits && w1 && b1 && w2 && b2 && out, ECGMM_ERR_ARG, "perturb_head_fused: null pointer");
  ECGMM_CHECK(ecgmm_perturb_head_fused_supported(D, kPfHid, C), ECGMM_ERR_SHAPE,
              "perturb_head_fused: D=%d C=%d not covered (D %% 64 == 0, D <= 768, C <= 8, hidden 128)", D, C);
  ECGMM_CHECK(cls < C, ECGMM_ERR_ARG, "perturb_head_fused: class index %d out of range", cls);
  ECGMM_CHECK(S >= 0 && V >= 0 && S <= 0x7fffffffLL / ((V + kPfTile - 1) / kPfTile + 1), ECGMM_ERR_SHAPE,
              "perturb_head_fused: extent");
  ECGMM_CHECK(((reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(bg)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(bits) & 7) == 0 &&
                  ((reinterpret_cast<uintptr_t>(b1) | reinterpret_cast<uintptr_t>(w2)) & 15) == 0,
              ECGMM_ERR_ALIGN, "perturb_head_fused: e / bg / b1 / w2 must be 16-byte aligned, bits 8-byte aligned");
  if (S == 0 || V == 0) return ECGMM_OK;
  PerturbFusedParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap_2d(&p.w_map, w1, (uint64_t)D, kPfHid, (uint64_t)D * 2, 64, kPfHid);
  if (rc) return rc;
  p.e = reinterpret_cast<const __nv_bfloat16*>(e);
  p.bg = reinterpret_cast<const __nv_bfloat16*>(bg);
  p.bits = bits;
  p.b1 = b1;
  p.w2 = w2;
  p.b2 = b2;
  p.out = out;
  p.S = (int)S;
  p.V = V;
  p.D = D;
  p.C = C;
  p.cls = cls;
  p.kchunks = D / 64;
  p.tiles_v = ceil_div(V, kPfTile);
  p.total_tiles = (int)S * p.tiles_v;
  const int smem = (p.kchunks + 2) * kPfChunk + 256 + 1024;
  static bool configured[kMaxDevices] = {};
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(perturb_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ECGMM_CUDA(cudaFuncSetAttribute(perturb_fused_kernel<kPfMaxC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[ds] = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (C <= 2)
    perturb_fused_kernel<2><<<grid, kPfThreads, smem, (cudaStream_t)stream>>>(p);
  else
    perturb_fused_kernel<kPfMaxC><<<grid, kPfThreads, smem, (cudaStream_t)stream>>>(p);
  return check_launch("perturb_fused_kernel");
}
