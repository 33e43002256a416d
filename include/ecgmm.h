/*
 * ecgmm.h -- C ABI of libecgmm.so, the sm_100a kernel library behind the drop-in
 * ECGMultimodalModel (reference: /root/reference/multimodal_paper_modal_balance.py:197-354,
 * multimodal.py:333-469).
 *
 * The reference has no FFI layer: its hot path is stock torch.nn operators dispatched to
 * ATen -> cuDNN/cuBLAS.  Each entry point below therefore cites the torch operator call site
 * in the reference that it replaces.  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch caching allocator);
 *     the library never allocates, frees or synchronises, and launches only on `stream`;
 *   - return value 0 = success, negative = error (see ECGMM_ERR_*); the text of the last
 *     error on the calling thread is returned by ecgmm_last_error();
 *   - activations are NHWC (channels-last) bf16 unless stated; parameters, statistics,
 *     gradients of parameters and optimizer state are fp32; reductions accumulate in fp32/fp64;
 *   - `void* stream` is a cudaStream_t.
 */
#ifndef ECGMM_H_
#define ECGMM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECGMM_OK 0
#define ECGMM_ERR_SHAPE (-1) /* unsupported shape / size */
#define ECGMM_ERR_ALIGN (-2) /* pointer or stride not aligned as required */
#define ECGMM_ERR_CUDA (-3)  /* CUDA runtime / driver error */
#define ECGMM_ERR_ARCH (-4)  /* device is not sm_100 */
#define ECGMM_ERR_ARG (-5)   /* null pointer / invalid flag */

typedef uint16_t ecgmm_bf16; /* raw bfloat16 bits */

int ecgmm_version(void);
const char* ecgmm_last_error(void);
/* 0 when the current device can run the library (compute capability 10.x). */
int ecgmm_check_device(void);

/* ------------------------------------------------------------------ layout / precision */

/* NCHW fp32 -> NHWC bf16 (and back).  Replaces the implicit layout of every torch tensor the
 * reference feeds to Conv2d/Conv1d (multimodal_paper_modal_balance.py:325,328). */
int ecgmm_nchw_f32_to_nhwc_bf16(const float* x, ecgmm_bf16* y, int N, int C, int H, int W, void* stream);
int ecgmm_nhwc_bf16_to_nchw_f32(const ecgmm_bf16* x, float* y, int N, int C, int H, int W, void* stream);

/* Conv weight fp32 OIHW [O][I][R][S] -> bf16 shadows used by the implicit GEMMs:
 *   w_fwd   [O][R][S][I]  (B operand of the forward GEMM, K = (r,s,i) contiguous)
 *   w_dgrad [I][R][S][O]  (B operand of the data-gradient GEMM, K = (r,s,o) contiguous)
 * Either output may be NULL.  Master weights stay fp32 in the state_dict. */
int ecgmm_conv_weight_prep(const float* w_oihw, ecgmm_bf16* w_fwd, ecgmm_bf16* w_dgrad, int O, int I, int R, int S,
                           void* stream);

/* ------------------------------------------------------------------ ResNet18 stem
 * torchvision resnet.py:197 conv1 = Conv2d(3,64,k7,s2,p3,bias=False), reached from
 * multimodal_paper_modal_balance.py:325 `self.image_encoder(image)`.
 * The 7x7/s2 convolution over 3 channels is evaluated as a 4x4/s1 convolution over a
 * space-to-depth (2x2) view of the zero-padded image with 12 (padded to 16) channels. */

/* geometry of the space-to-depth buffer for an H x W image: rows, cols (16 channels each) */
void ecgmm_stem_s2d_dims(int H, int W, int* Hs, int* Ws);
/* image NCHW (3 channels; fp32 if x_is_bf16 == 0, else bf16) -> xs [N][Hs][Ws][16] bf16 */
int ecgmm_stem_s2d(const void* x, int x_is_bf16, ecgmm_bf16* xs, int N, int H, int W, void* stream);
/* w [64][3][7][7] fp32 -> w_s2d [64][4][4][16] bf16 */
int ecgmm_stem_weight_prep(const float* w, ecgmm_bf16* w_s2d, void* stream);
/* y [N][Ho][Wo][64] bf16, Ho = (H+6-7)/2+1 */
int ecgmm_stem_conv_fwd(const ecgmm_bf16* xs, const ecgmm_bf16* w_s2d, ecgmm_bf16* y, int N, int H, int W,
                        void* stream);
/* dw [64][3][7][7] fp32 += sum over pixels (atomic accumulation; caller zeroes dw) */
int ecgmm_stem_conv_wgrad(const ecgmm_bf16* xs, const ecgmm_bf16* dy, float* dw, int N, int H, int W,
                          void* stream);

/* ------------------------------------------------------------------ implicit-GEMM convolutions
 * Replaces nn.Conv2d in torchvision BasicBlock (resnet.py:59-104) and nn.Conv1d in
 * BasicBlock1D (multimodal_paper_modal_balance.py:67-93; 1-D = H == 1, R == 1).
 * Supported: Cin % 64 == 0, Cout % 64 == 0, (R,S,stride,pad) in
 *   {(3,3,1,1), (3,3,2,1), (1,1,1,0), (1,1,2,0), (1,3,1,1), (1,3,2,1)}.
 * x [N][H][W][Cin], y/dy [N][Ho][Wo][Cout], tcgen05.mma with fp32 accumulation in TMEM. */
int ecgmm_conv2d_fwd(const ecgmm_bf16* x, const ecgmm_bf16* w_fwd, ecgmm_bf16* y, int N, int H, int W, int Cin,
                     int Cout, int R, int S, int stride, int padH, int padW, void* stream);
/* dx = conv_transpose(dy, w).  accumulate != 0 adds into dx instead of overwriting. */
int ecgmm_conv2d_dgrad(const ecgmm_bf16* dy, const ecgmm_bf16* w_dgrad, ecgmm_bf16* dx, int N, int H, int W,
                       int Cin, int Cout, int R, int S, int stride, int padH, int padW, int accumulate,
                       void* stream);
/* dw (fp32, OIHW) += x^T * dy (atomic accumulation across CTAs; caller zeroes dw when needed). */
int ecgmm_conv2d_wgrad(const ecgmm_bf16* x, const ecgmm_bf16* dy, float* dw_oihw, int N, int H, int W, int Cin,
                       int Cout, int R, int S, int stride, int padH, int padW, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ECGMM_H_ */
