// Batched perturbation inference of the fusion head in ONE kernel (BASELINE.json configs[3]):
//     prob[s][v] = softmax( W2 relu( W1 (z[v] * e[s] + (1 - z[v]) * b) + b1 ) + b2 )[class]
// for V masked variants of every sample's fused embedding e[s] (shap_fusion_modal_balance.py:126-159,
// lime_fusion_modal_balance.py:126-131 drive fusion_classifier row by row through a wrapper).
//
// The three-kernel path of perturb.cu writes the S*V x D variants to HBM as bf16, reads them back in the GEMM and
// round-trips the hidden layer once more (head_tail).  Here the variants never exist outside shared memory:
//   producer warps          build the A operand of the GEMM directly in its SWIZZLE_128B K-major shared-memory layout.
//                           The masks are shared by all samples: they are packed once per call to one BIT per element
//                           (ecgmm_perturb_pack_masks: V x D / 8 bytes, chunk-major) in an order chosen so that one
//                           shift puts four elements' bits on the sign bits of a register's four bytes and PRMT's
//                           sign-replicate mode expands them to the 16-bit select masks of two bf16 pairs.  The
//                           selection between the bf16 bit patterns of e[s] and b is exact (no arithmetic).  Groups of
//                           four warps take turns over the K chunks (one ring stage each); a thread owns ONE 16-byte
//                           piece of e[s] / b and eight rows (see PfChunk), the next chunk's operands are loaded
//                           before the current one is built;
//   one MMA warp            tcgen05.mma (M128 or, on CTA pairs, M256) x N128 x K16 against W1, which stays RESIDENT in
//                           shared memory for the whole persistent CTA, accumulators double-buffered in TMEM;
//   4 epilogue warps        TMEM -> + b1 -> ReLU -> Linear(128, C) -> softmax, fp32, one thread per variant row, the
//                           head's operands in constant memory.
// Two kernels: perturb_fused_pair_kernel (default when a sample has more than 128 variants; cta_group::2, half of W1
// per CTA, 8-stage ring, 16 producer warps: 0.73 of the measured tensor peak at D = 768) and perturb_fused_kernel (single
// CTA, all of W1 resident = 192 KB, 2-stage ring: 0.54).  HBM traffic per variant: 4 (or 4 C) bytes out.
#include "common.h"
#include "ptx.cuh"

#include <stdlib.h>
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

// Cache policy of the global loads in these kernels.  The L1 left beside 227 KB of shared memory is ~28 KB: the mask words
// (streamed, each used once) bypass it, the e[s] / b pieces (re-read by every tile of a sample) are marked evict-last.
__device__ __forceinline__ uint4 ldg_hot_u4(const void* ptr) {
  uint4 v;
  asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(ptr));
  return v;
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

// ------------------------------------------------------------------------------------------------- producer helpers
// One producer group = 4 warps = 128 threads builds one K chunk (128 variant rows x 64 elements = 16 KB) of the A operand.
// Thread (piece j = tid & 7, row group rg = tid >> 3) owns the 16-byte piece j of rows rg, rg + 16, ..., rg + 112: its e / b
// pieces are loaded ONCE per chunk (2 loads per thread; a warp reads 8 distinct pieces = 128 contiguous bytes), the select
// masks of its 8 rows come from 8 mask words, and the 8 stores go to base + q * 2048 (row & 7 == rg & 7: the swizzled
// piece offset is a per-thread constant).  The first version gave every thread a whole ROW: 16 warp-uniform 128-bit loads
// per warp and chunk, which cost 4 LSU wavefronts each although all lanes read the same 16 bytes -- with the epilogue's
// uniform loads the LSU data pipe carried ~660 wavefronts per chunk and was the limiter whatever the ring depth, the
// number of producer warps or the CTA pairing (profiles/r02y_perturb_ablation.txt).
struct PfChunk {
  uint4 ev, bv;
  uint32_t mw[8];
};
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t* ptr) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(ptr));
  return v;
}
// s: sample, v0: first variant row of this CTA's tile, kc: K chunk
__device__ __forceinline__ void pf_load_chunk(const PerturbFusedParams& p, PfChunk& c, int s, int v0, int kc, int j, int rg) {
  c.ev = ldg_hot_u4(p.e + (size_t)s * p.D + kc * 64 + j * 8);
  c.bv = ldg_hot_u4(p.bg + kc * 64 + j * 8);
  const uint32_t* mb = p.bits + (size_t)kc * p.V * 2 + (j >> 2);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int v = v0 + rg + 16 * q;
    c.mw[q] = ldg_stream_u32(mb + (size_t)(v < p.V ? v : 0) * 2);
  }
}
// dst = stage base + rg * 128 + ((j ^ (rg & 7)) << 4); jj = j & 3
__device__ __forceinline__ void pf_store_chunk(const PfChunk& c, uint8_t* dst, int jj) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t r0 = c.mw[q] << (2 * jj), r1 = c.mw[q] << (2 * jj + 1);  // sign bits of bytes 0..3 = 4 elements
    const uint32_t s0 = prmt_sign(r0, 0x9988u), s1 = prmt_sign(r0, 0xBBAAu);
    const uint32_t s2 = prmt_sign(r1, 0x9988u), s3 = prmt_sign(r1, 0xBBAAu);
    uint4 o;
    o.x = (c.ev.x & s0) | (c.bv.x & ~s0);
    o.y = (c.ev.y & s1) | (c.bv.y & ~s1);
    o.z = (c.ev.z & s2) | (c.bv.z & ~s2);
    o.w = (c.ev.w & s3) | (c.bv.w & ~s3);
    *reinterpret_cast<uint4*>(dst + q * 2048) = o;
  }
}

// The head's small operands live in CONSTANT memory for the duration of a launch: b1 [128], W2 [C][128], b2 [C], copied
// device-to-device on the launch's stream by ecgmm_perturb_head_fused.  The epilogue then needs no load instruction at
// all for them (constant-bank operands of FADD / FFMA); as warp-uniform 128-bit global loads they were 96 load
// instructions and ~400 LSU wavefronts per tile and warp.  (One head per device at a time: calls with DIFFERENT heads
// must not run concurrently on two streams of the same device -- stated in include/ecgmm.h.)
__constant__ float c_pf_head[kPfHid + kPfMaxC * kPfHid + kPfMaxC];

// Epilogue arithmetic shared by the single-CTA and the CTA-pair kernel: one thread = one variant row of the accumulator
// at t_addr (128 fp32 columns in TMEM) -> + b1 -> ReLU -> Linear(128, C) (without b2).
template <int MAXC>
__device__ __forceinline__ void pf_head_tail(const PerturbFusedParams& p, uint32_t t_addr, float (&logit)[MAXC]) {
#pragma unroll
  for (int c = 0; c < MAXC; ++c) logit[c] = 0.f;
#pragma unroll
  for (int cc = 0; cc < kPfHid / 32; ++cc) {
    uint32_t r[32];
    tmem_ld_32x32(t_addr + cc * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float h = fmaxf(__uint_as_float(r[j]) + c_pf_head[cc * 32 + j], 0.f);
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (MAXC <= 2 || c < p.C) logit[c] = fmaf(h, c_pf_head[kPfHid + c * kPfHid + cc * 32 + j], logit[c]);
    }
  }
}

// + b2, then the logits (cls < 0) or softmax(logits)[cls] of row o.
template <int MAXC>
__device__ __forceinline__ void pf_store(const PerturbFusedParams& p, size_t o, float (&logit)[MAXC]) {
  if (p.cls < 0) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < p.C) p.out[o * p.C + c] = logit[c] + c_pf_head[kPfHid + kPfMaxC * kPfHid + c];
  } else {
    float mx = -INFINITY, den = 0.f, num = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < p.C) {
        logit[c] += c_pf_head[kPfHid + kPfMaxC * kPfHid + c];
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
    const int tid = ((warp - 2) & 3) * 32 + lane;
    const int j = tid & 7, rg = tid >> 3;
    const uint32_t dst_off = rg * 128 + ((j ^ (rg & 7)) << 4);
    auto norm = [&](int& t, int& kc) {
      while (kc >= p.kchunks && t < p.total_tiles) {
        kc -= p.kchunks;
        t += gridDim.x;
      }
    };
    auto load = [&](PfChunk& c, int t, int kc) {
      const int s = t / p.tiles_v;
      pf_load_chunk(p, c, s, (t - s * p.tiles_v) * kPfTile, kc, j, rg);
    };
    int t = blockIdx.x, kc = g;
    norm(t, kc);
    PfChunk cur, nxt;
    if (t < p.total_tiles) load(cur, t, kc);
    uint32_t n = 0;  // chunks this group has built: stage g is on its n-th use
    while (t < p.total_tiles) {
      int nt = t, nkc = kc + 2;
      norm(nt, nkc);
      if (nt < p.total_tiles) load(nxt, nt, nkc);  // in flight across the wait and the build of the current chunk
      mbar_wait(&aempty[g], (n & 1u) ^ 1u);
      pf_store_chunk(cur, sA + g * kPfChunk + dst_off, j & 3);
      fence_proxy_async_smem();  // the tensor core reads shared memory through the async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&afull[g]);
      ++n;
      cur = nxt;
      t = nt;
      kc = nkc;
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
      pf_head_tail<MAXC>(p, t_addr, logit);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (v < p.V) pf_store<MAXC>(p, (size_t)s * p.V + v, logit);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * kPfHid);
}

// ---------------------------------------------------------------------------------------------------------------------
// The same path on CTA PAIRS (tcgen05 cta_group::2, M = 256): the two CTAs of a cluster take two adjacent 128-variant
// tiles of the same sample; each keeps HALF of W1 resident (64 of the 128 hidden units: 96 KB at D = 768), which frees
// 128 KB for an 8-stage ring of A chunks, and each SM reads half of B (shared-memory traffic per chunk: 16 KB written +
// 24 KB read instead of 16 + 32).
// Barriers: every CTA owns aempty[] / tfull[] (rank 0's multicast tcgen05.commit arrives on both copies); afull[],
// tempty[] and wfull are rank 0's: the 4 producer warps of EACH CTA arrive on afull[stage] (after fence.proxy.async:
// the tensor core reads rank 1's shared memory on behalf of rank 0's MMA, whose thread waits with an acquire at cluster
// scope), all 8 epilogue warps on tempty[], and both CTAs' TMA bytes of W1 complete on wfull.
constexpr int kPpStages = 8;
constexpr int kPpWChunk = 64 * 128;  // one 64-wide K chunk of HALF of W1: 64 hidden units x 128 B
// Producer groups of 4 warps; group g builds the chunks i = g (mod kPpGroups): four chunks under construction at once,
// the ring has room for them.
constexpr int kPpGroups = 4;
constexpr int kPpThreads = (2 + 4 * kPpGroups + 4) * 32;

template <int MAXC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPpThreads, 1)
    perturb_fused_pair_kernel(const __grid_constant__ PerturbFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // identical in both CTAs
  uint8_t* sA = smem;                                  // kPpStages x 16 KB
  uint8_t* sW = smem + kPpStages * kPfChunk;           // kchunks x 8 KB, resident
  uint64_t* afull = reinterpret_cast<uint64_t*>(sW + p.kchunks * kPpWChunk);
  uint64_t* aempty = afull + kPpStages;
  uint64_t* tfull = aempty + kPpStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.w_map);
    for (int i = 0; i < kPpStages; ++i) {
      mbar_init(&afull[i], 8);   // 4 producer warps of each CTA
      mbar_init(&aempty[i], 1);  // multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);  // 4 epilogue warps of each CTA
    }
    mbar_init(wfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 2 * kPfHid);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {  // this CTA's half of W1, once; the bytes of both CTAs complete on rank 0's barrier
      if (rank == 0) mbar_expect_tx(wfull, 2 * p.kchunks * kPpWChunk);
      for (int kc = 0; kc < p.kchunks; ++kc)
        tma_load_2d_pair(sW + kc * kPpWChunk, &p.w_map, wfull, kc * 64, (int)rank * (kPfHid / 2));
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(256, kPfHid, 0, 0);
      const uint64_t w_desc0 = make_sw128_desc(smem_u32(sW), 0, 1024);
      const uint64_t a_desc0 = make_sw128_desc(smem_u32(sA), 0, 1024);
      mbar_wait(wfull, 0);
      tc_fence_after();
      uint32_t i = 0;  // running K-chunk counter: stage = i % 8, parity = (i / 8) & 1
      int it = 0;
      for (int q = cid; q < p.total_tiles; q += n_clusters, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kPfHid;
        for (int kc = 0; kc < p.kchunks; ++kc, ++i) {
          const uint32_t stage = i % kPpStages;
          mbar_wait_cluster(&afull[stage], (i / kPpStages) & 1u);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + (uint64_t)(stage * (kPfChunk >> 4));
          const uint64_t w_desc = w_desc0 + (uint64_t)(kc * (kPpWChunk >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_pair(d_tmem, a_desc + 2 * k, w_desc + 2 * k, idesc, (kc | k) != 0);
          umma_commit_pair(&aempty[stage]);
        }
        umma_commit_pair(&tfull[acc]);
      }
    }
  } else if (warp < 2 + 4 * kPpGroups) {
    // ------------------------------------------------------------ producers (both CTAs: own 128 variant rows)
    const int g = (warp - 2) >> 2;
    const int tid = ((warp - 2) & 3) * 32 + lane;
    const int j = tid & 7, rg = tid >> 3;
    const uint32_t dst_off = rg * 128 + ((j ^ (rg & 7)) << 4);
    auto norm = [&](int& t, int& kc) {
      while (kc >= p.kchunks && t < p.total_tiles) {
        kc -= p.kchunks;
        t += n_clusters;
      }
    };
    auto load = [&](PfChunk& c, int t, int kc) {
      const int s = t / p.tiles_v;
      pf_load_chunk(p, c, s, (t - s * p.tiles_v) * (2 * kPfTile) + (int)rank * kPfTile, kc, j, rg);
    };
    int t = cid, kc = g;
    norm(t, kc);
    PfChunk cur, nxt;
    if (t < p.total_tiles) load(cur, t, kc);
    uint32_t i = g;  // running K-chunk counter of the pair (this group: every kPpGroups-th chunk)
    while (t < p.total_tiles) {
      int nt = t, nkc = kc + kPpGroups;
      norm(nt, nkc);
      if (nt < p.total_tiles) load(nxt, nt, nkc);
      const uint32_t stage = i % kPpStages;
      mbar_wait(&aempty[stage], ((i / kPpStages) & 1u) ^ 1u);
      pf_store_chunk(cur, sA + stage * kPfChunk + dst_off, j & 3);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_rank0(&afull[stage]);  // plain remote arrival, as CUTLASS's 2-SM transform kernels do; a
      // release at cluster scope (MEMBAR + ERRBAR) was 15 % of all samples and is not needed after fence.proxy.async
      i += kPpGroups;
      cur = nxt;
      t = nt;
      kc = nkc;
    }
  } else {
    // ------------------------------------------------------------ epilogue (both CTAs: own 128 rows)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int it = 0;
    for (int q = cid; q < p.total_tiles; q += n_clusters, ++it) {
      const int s = q / p.tiles_v;
      const int v = (q - s * p.tiles_v) * (2 * kPfTile) + (int)rank * kPfTile + row;
      const int acc = it & 1;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kPfHid;
      float logit[MAXC];
      pf_head_tail<MAXC>(p, t_addr, logit);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_rank0(&tempty[acc]);
      if (v < p.V) pf_store<MAXC>(p, (size_t)s * p.V + v, logit);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may leave (or free TMEM) while the other can still signal or be read
  if (warp == 1) tmem_dealloc_pair(tmem_base, 2 * kPfHid);
}

// masks [V][D] bytes (nonzero = keep) -> one bit per element, 32 elements per word: element 4k + b of a word sits on bit
// 8b + 7 - k, so `word << k` carries elements 4k .. 4k+3 on the sign bits of its four bytes (see the producers above).
// Layout: K-chunk major, bits[kc][v][2] (the two words of row v's 64-wide chunk kc), so that the 32 rows a producer warp
// loads for one chunk are 256 contiguous bytes (row-major words cost 24 L1 tag requests per load instruction and
// 32-byte sectors for 8 useful bytes).
__global__ void perturb_pack_masks_kernel(const uint8_t* __restrict__ masks, uint32_t* __restrict__ bits, size_t nwords,
                                          int V, int words) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(masks + i * 32);
    const uint4 lo = __ldg(src), hi = __ldg(src + 1);
    const uint32_t q[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};  // q[k] = elements 4k .. 4k+3
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) w |= (__vcmpne4(q[k], 0u) & 0x80808080u) >> k;
    const size_t v = i / words;
    const int wi = (int)(i - v * words);
    bits[((size_t)(wi >> 1) * V + v) * 2 + (wi & 1)] = w;
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
  ECGMM_CHECK(V >= 0 && D > 0 && D % 64 == 0, ECGMM_ERR_SHAPE, "perturb_pack_masks: D=%d must be a positive multiple of 64", D);
  ECGMM_CHECK((reinterpret_cast<uintptr_t>(masks) & 15) == 0, ECGMM_ERR_ALIGN, "perturb_pack_masks: masks must be 16-byte aligned");
  const size_t nwords = (size_t)V * (D / 32);
  if (nwords == 0) return ECGMM_OK;
  size_t blocks = (nwords + 255) / 256;
  if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
  perturb_pack_masks_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(masks, bits, nwords, V, D / 32);
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
  // CTA pairs (M = 256 variants per tile) unless ECGMM_PERTURB_PAIR=0 or a sample has at most one 128-variant tile
  const char* ep = getenv("ECGMM_PERTURB_PAIR");
  const bool pair = (ep ? atoi(ep) != 0 : true) && V > kPfTile && num_sms() >= 2;
  PerturbFusedParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap_2d(&p.w_map, w1, (uint64_t)D, kPfHid, (uint64_t)D * 2, 64, pair ? kPfHid / 2 : kPfHid);
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
  p.tiles_v = ceil_div(V, pair ? 2 * kPfTile : kPfTile);
  p.total_tiles = (int)S * p.tiles_v;
  static bool configured[kMaxDevices] = {};
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(perturb_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ECGMM_CUDA(cudaFuncSetAttribute(perturb_fused_kernel<kPfMaxC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ECGMM_CUDA(cudaFuncSetAttribute(perturb_fused_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ECGMM_CUDA(cudaFuncSetAttribute(perturb_fused_pair_kernel<kPfMaxC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[ds] = true;
  }
  ECGMM_CUDA(cudaMemcpyToSymbolAsync(c_pf_head, b1, kPfHid * sizeof(float), 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  ECGMM_CUDA(cudaMemcpyToSymbolAsync(c_pf_head, w2, (size_t)C * kPfHid * sizeof(float), kPfHid * sizeof(float),
                                     cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  ECGMM_CUDA(cudaMemcpyToSymbolAsync(c_pf_head, b2, C * sizeof(float), (kPfHid + kPfMaxC * kPfHid) * sizeof(float),
                                     cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  if (pair) {
    const int smem = kPpStages * kPfChunk + p.kchunks * kPpWChunk + 512 + 1024;
    const int clusters = p.total_tiles < num_sms() / 2 ? p.total_tiles : num_sms() / 2;
    if (C <= 2)
      perturb_fused_pair_kernel<2><<<2 * clusters, kPpThreads, smem, (cudaStream_t)stream>>>(p);
    else
      perturb_fused_pair_kernel<kPfMaxC><<<2 * clusters, kPpThreads, smem, (cudaStream_t)stream>>>(p);
    return check_launch("perturb_fused_pair_kernel");
  }
  const int smem = (p.kchunks + 2) * kPfChunk + 256 + 1024;
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (C <= 2)
    perturb_fused_kernel<2><<<grid, kPfThreads, smem, (cudaStream_t)stream>>>(p);
  else
    perturb_fused_kernel<kPfMaxC><<<grid, kPfThreads, smem, (cudaStream_t)stream>>>(p);
  return check_launch("perturb_fused_kernel");
}
