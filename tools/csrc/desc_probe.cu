// Hardware probe (development only): can a SWIZZLE_128B shared-memory matrix descriptor
// start at an arbitrary 128-byte row of a TMA-written tile?  The halo-reuse convolution kernels
// depend on the answer (a filter tap becomes a row offset into ONE staged input tile instead of
// its own TMA load).  mode bit 0: 0 = K-major A (forward style), 1 = MN-major A (wgrad style);
// mode bit 1: set the descriptor's base-offset field to (start_address >> 7) & 7.
// mode bit 2 (round-2 question, DESIGN.md "known headroom" 3): may the 64-wide N atoms of an MN-major SWIZZLE_128B
// operand OVERLAP, i.e. leading-dimension byte offset = 128 B = one pixel row?  Then ONE staged input-row box is the
// B operand of an N = 192 MMA whose three atoms are the three horizontal filter taps (the same box read from pixel
// s, s+1, s+2), which is what a transposed weight-gradient GEMM (M = Cout, N = taps x Cin) needs to leave the
// shared-memory-bound M128 x N64 shape.  D[m][j*64 + c] = sum_{k<32} a[k][m] * a[k + shift + j][c]; out is [128][192].
// Not part of libecgmm.so: tools/desc_probe.py compiles this file (+ csrc/common.cu) into tools/build/libecgmm_probe.so.
#include "../../ecg-multimodal-model_b200/csrc/common.h"
#include "../../ecg-multimodal-model_b200/csrc/ptx.cuh"

namespace ecgmm {

constexpr int kProbeRows = 160;  // rows of 64 bf16 staged per atom

struct alignas(64) ProbeParams {
  CUtensorMap a_map;  // [rows][128] bf16, box (64, kProbeRows)
  CUtensorMap b_map;  // K-major: [64][64] (n, k); MN-major: [32][64] (k, n)
  float* out;         // [128][64]
  int shift, mode;
};

__global__ void __launch_bounds__(128, 1) desc_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int kAtom = kProbeRows * 128;  // 20 KiB, multiple of 1024
  uint8_t* sA = smem;                      // two atoms (channels 0..63 / 64..127)
  uint8_t* sB = smem + 2 * kAtom;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 8192);
  uint64_t* done = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool mn = p.mode & 1;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t bbytes = mn ? 32 * 128 : 64 * 128;
    mbar_expect_tx(bar, 2 * kAtom + bbytes);
    tma_load_2d(sA, &p.a_map, bar, 0, 0);
    tma_load_2d(sA + kAtom, &p.a_map, bar, 64, 0);
    tma_load_2d(sB, &p.b_map, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA) + p.shift * 128;
    uint64_t boff = 0;
    if (p.mode & 2) boff = static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    if (p.mode & 4) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 192, 1, 1);
      const uint64_t a_desc = make_sw128_desc(smem_u32(sA), kAtom, 1024);  // channels 0..63 | 64..127, rows 0..31
      const uint64_t b_desc = make_sw128_desc(a_addr, 128, 1024) | boff;   // three atoms one 128-byte row apart
      for (int k = 0; k < 2; ++k) umma_bf16(tmem, a_desc + k * 128, b_desc + k * 128, idesc, k != 0);
    } else if (!mn) {
      // D[m][n] = sum_k A[m + shift][k] * B[n][k], m < 128 (rows of atom 0), K = 64
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      const uint64_t a_desc = make_sw128_desc(a_addr, 0, 1024) | boff;
      const uint64_t b_desc = make_sw128_desc(smem_u32(sB), 0, 1024);
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
    } else {
      // D[m][n] = sum_{k<32} A[k + shift][m] * B[k][n], m < 128 (atom 0 | atom 1)
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      const uint64_t a_desc = make_sw128_desc(a_addr, kAtom, 1024) | boff;
      const uint64_t b_desc = make_sw128_desc(smem_u32(sB), 4096, 1024);
      for (int k = 0; k < 2; ++k) umma_bf16(tmem, a_desc + k * 128, b_desc + k * 128, idesc, k != 0);
    }
    umma_commit(done);
  }
  __syncwarp();
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t r[32];
  const int ncols = (p.mode & 4) ? 192 : 64;
  for (int c = 0; c < ncols / 32; ++c) {
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.out[(warp * 32 + lane) * ncols + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace ecgmm

using namespace ecgmm;

extern "C" int ecgmm_debug_desc_probe(const ecgmm_bf16* a, const ecgmm_bf16* b, float* out, int shift, int mode,
                                      void* stream) {
  ECGMM_CHECK(a && b && out, ECGMM_ERR_ARG, "desc_probe: null pointer");
  ECGMM_CHECK(shift >= 0 && shift <= 16, ECGMM_ERR_ARG, "desc_probe: shift %d", shift);
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap_2d(&p.a_map, a, 128, kProbeRows, 256, 64, kProbeRows);
  if (rc) return rc;
  if (mode & 1)
    rc = make_tmap_2d(&p.b_map, b, 64, 32, 128, 64, 32);
  else
    rc = make_tmap_2d(&p.b_map, b, 64, 64, 128, 64, 64);
  if (rc) return rc;
  p.out = out;
  p.shift = shift;
  p.mode = mode;
  const int smem = 2 * kProbeRows * 128 + 8192 + 64 + 1024;
  static bool configured = false;
  if (!configured) {
    ECGMM_CUDA(cudaFuncSetAttribute(desc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  desc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  return check_launch("desc_probe_kernel");
}
