// Host-side helpers shared by every translation unit of libecgmm.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ecgmm.h"

namespace ecgmm {

// Last error text, readable through ecgmm_last_error(); thread-local so one host
// thread per GPU process never races with another.
void set_error(const char* fmt, ...);

#define ECGMM_CHECK(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::ecgmm::set_error(__VA_ARGS__); \
      return (code);                  \
    }                                 \
  } while (0)

#define ECGMM_CUDA(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::ecgmm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ECGMM_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

// Number of kernel launches issued by this library in this process (ecgmm_launch_count()).
extern unsigned long long g_launches;

inline int check_launch(const char* what, int n_kernels = 1) {
  g_launches += (unsigned long long)n_kernels;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return ECGMM_ERR_CUDA;
  }
  return ECGMM_OK;
}

int num_sms();

// Index of the current device for per-device one-time setup (cudaFuncSetAttribute): a process normally drives one
// GPU, but nothing here may silently depend on that.
constexpr int kMaxDevices = 64;
inline int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

// 4-D activation view for TMA: dims (C, W, H, N) innermost first, byte strides for W/H/N,
// box (box_c, box_w, box_h, 1), SWIZZLE_128B, zero fill out of bounds.
int make_tmap_4d(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N,
                 uint64_t strideW_bytes, uint64_t strideH_bytes, uint64_t strideN_bytes, uint32_t box_c,
                 uint32_t box_w, uint32_t box_h);
// 2-D row-major matrix [rows][cols] bf16 (cols contiguous); box (box_cols, box_rows), SWIZZLE_128B.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                 uint32_t box_cols, uint32_t box_rows);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Pick the (TH, TW) power-of-two factorisation of `npix` output pixels that wastes the fewest
// tile slots on an OH x OW output plane.
void pick_tile(int OH, int OW, int npix, int* TH, int* TW);

}  // namespace ecgmm
