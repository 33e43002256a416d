// 16-byte vector access helpers for bf16 channels-last activations (8 channels per thread).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ecgmm {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(b[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* b = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  return v;
}

// streaming (read-once) 16-byte load that does not pollute L1
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// ReLU bit mask of an 8-channel vector (bit j = channel j, as bn_apply writes it) applied to the four packed bf16
// words of a gradient vector: one multiply puts bit 2k on the sign of byte 1 and bit 2k+1 on the sign of byte 3 (the
// two shifted copies of the 8-bit mask do not overlap), prmt.b32's sign-replicate selectors (nibble 8 + i; __byte_perm
// only honours three bits per nibble) widen them to 16-bit lane masks.  3 instructions per 2 channels; a masked lane
// becomes +0.0, the value `dz = 0.f` packs to.
__device__ __forceinline__ uint32_t relu_lane_mask(uint32_t m8, int k) {
  const uint32_t t = m8 * ((1u << (15 - 2 * k)) + (1u << (30 - 2 * k)));
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(t), "r"(0u), "r"(0xBB99u));
  return d;
}
__device__ __forceinline__ void relu_mask_words(uint32_t m8, uint4& v) {
  v.x &= relu_lane_mask(m8, 0);
  v.y &= relu_lane_mask(m8, 1);
  v.z &= relu_lane_mask(m8, 2);
  v.w &= relu_lane_mask(m8, 3);
}

// Ampere-style asynchronous 16-byte copies global -> shared (LDGSTS, L2-only caching).  Used thread-privately by the
// streaming BatchNorm kernels: a thread copies its own future vectors into its own slots of a shared-memory ring and
// reads back only those, so cp.async.wait_group is all the synchronisation there is (no barrier, nothing to hang on);
// the bytes in flight live in shared memory instead of registers.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0];
  const float4 b = reinterpret_cast<const float4*>(p)[1];
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Column sums across a warp whose lane l holds row l of a 32-row x 32-column tile (v[j] = column j of its row):
// a butterfly reduce-scatter, 31 shuffles per quantity; afterwards lane j holds the sum over the 32 rows of
// column j.  Used by the convolution epilogues to produce BatchNorm batch statistics without a second pass.
template <int H>
__device__ __forceinline__ void colsum_step(float (&a)[32], int lane) {
  const bool up = (lane & H) != 0;
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float keep = up ? a[i + H] : a[i];
    const float send = up ? a[i] : a[i + H];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, H);
  }
}
__device__ __forceinline__ float warp_colsum32(float (&a)[32], int lane) {
  colsum_step<16>(a, lane);
  colsum_step<8>(a, lane);
  colsum_step<4>(a, lane);
  colsum_step<2>(a, lane);
  colsum_step<1>(a, lane);
  return a[0];
}
// v: the lane's 32 values (already rounded to what is stored); adds column `lane`'s sum / sum of squares.
__device__ __forceinline__ void warp_colstats32(const float (&v)[32], int lane, double& s, double& q) {
  float a[32], b[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    a[j] = v[j];
    b[j] = v[j] * v[j];
  }
  s += (double)warp_colsum32(a, lane);
  q += (double)warp_colsum32(b, lane);
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); every thread receives the result.
__device__ __forceinline__ float block_sum(float v, float* red /* >= 33 floats of shared memory */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = (lane < (int)((blockDim.x + 31) >> 5)) ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

}  // namespace ecgmm
