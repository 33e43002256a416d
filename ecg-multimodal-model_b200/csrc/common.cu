// Error reporting, device query and TMA tensor-map construction for libecgmm.so.
#include "common.h"

#include <stdarg.h>
#include <string.h>

#include <mutex>

namespace ecgmm {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// cuTensorMapEncodeTiled is fetched through the runtime so that the library carries no
// link-time dependency on libcuda.so (it must dlopen on a machine without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode();
  ECGMM_CHECK(fn != nullptr, ECGMM_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  ECGMM_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, ECGMM_ERR_ALIGN, "TMA base %p not 16-byte aligned",
              base);
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ECGMM_CHECK(r == CUDA_SUCCESS, ECGMM_ERR_CUDA,
              "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu strides %llu %llu %llu box %u %u "
              "%u %u",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              (unsigned long long)strides[0], (unsigned long long)(rank > 2 ? strides[1] : 0),
              (unsigned long long)(rank > 3 ? strides[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0,
              rank > 3 ? box[3] : 0);
  return ECGMM_OK;
}

int make_tmap_4d(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N,
                 uint64_t sW, uint64_t sH, uint64_t sN, uint32_t box_c, uint32_t box_w, uint32_t box_h) {
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {sW, sH, sN};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  return encode(out, base, 4, dims, strides, box);
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                 uint32_t box_cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode(out, base, 2, dims, strides, box);
}

void pick_tile(int OH, int OW, int npix, int* TH, int* TW) {
  long best = -1;
  for (int tw = npix; tw >= 1; tw >>= 1) {
    int th = npix / tw;
    if (tw > 256 || th > 256) continue;
    long slots = (long)ceil_div(OH, th) * ceil_div(OW, tw);
    // prefer wide tiles on ties (longer contiguous runs in NHWC)
    if (best < 0 || slots < best) {
      best = slots;
      *TH = th;
      *TW = tw;
    }
  }
}

}  // namespace ecgmm

extern "C" {

int ecgmm_version(void) { return 107; }

unsigned long long ecgmm_launch_count(void) { return ecgmm::g_launches; }

const char* ecgmm_last_error(void) { return ecgmm::g_err; }

int ecgmm_check_device(void) {
  int dev = 0;
  ECGMM_CUDA(cudaGetDevice(&dev));
  int major = 0;
  ECGMM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  ECGMM_CHECK(major == 10, ECGMM_ERR_ARCH, "device %d has compute capability %d.x; libecgmm needs sm_100", dev,
              major);
  return ECGMM_OK;
}

}  // extern "C"
