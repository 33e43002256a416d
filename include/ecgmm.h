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
/* kernels launched by this library in this process so far (bench.py reports the per-step delta) */
unsigned long long ecgmm_launch_count(void);
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
/* The same conversion for n weights in one launch (a HOST array of descriptors; O and I multiples of 32, R * S <= 9):
 * the 27 convolutions of the fusion model's two ResNets are refreshed after every optimizer step. */
typedef struct {
  const float* w;      /* fp32 [O][I][R][S] */
  ecgmm_bf16* w_fwd;   /* [O][R][S][I] or NULL */
  ecgmm_bf16* w_dgrad; /* [I][R][S][O] or NULL */
  int O, I, R, S;
} ecgmm_weight_prep_desc;
int ecgmm_conv_weight_prep_batch(const ecgmm_weight_prep_desc* descs, int n, void* stream);

/* ------------------------------------------------------------------ ResNet18 stem
 * torchvision resnet.py:197 conv1 = Conv2d(3,64,k7,s2,p3,bias=False), reached from
 * multimodal_paper_modal_balance.py:325 `self.image_encoder(image)`.
 * The 7x7/s2 convolution over 3 channels is evaluated as a 4x4/s1 convolution over a
 * space-to-depth (2x2) view of the zero-padded image with 12 (padded to 16) channels. */

/* geometry of the space-to-depth buffer for an H x W image: rows, cols (16 channels each) */
void ecgmm_stem_s2d_dims(int H, int W, int* Hs, int* Ws);
/* image NCHW (3 channels) -> xs [N][Hs][Ws][16] bf16.  x_dtype: 0 = fp32, 1 = bf16 (already normalised tensors,
 * what dataset.py:119-123 yields), 2 = uint8 raw pixels: torchvision ToTensor (u/255) + Normalize(0.5, 0.5) of
 * dataset.py:119-123 are applied on the fly in fp32, so a batch can cross PCIe at 1 byte per pixel value
 * (SURVEY.md section 8f rank 2). */
int ecgmm_stem_s2d(const void* x, int x_dtype, ecgmm_bf16* xs, int N, int H, int W, void* stream);
/* w [64][3][7][7] fp32 -> w_s2d [64][4][4][16] bf16 */
int ecgmm_stem_weight_prep(const float* w, ecgmm_bf16* w_s2d, void* stream);
/* y [N][Ho][Wo][64] bf16, Ho = (H+6-7)/2+1 */
int ecgmm_stem_conv_fwd(const ecgmm_bf16* xs, const ecgmm_bf16* w_s2d, ecgmm_bf16* y, int N, int H, int W,
                        void* stream);
/* The same convolution + the BatchNorm batch statistics of its (stored, bf16) output from the epilogue: psum / psq
 * [ecgmm_stem_conv_fwd_stats_rows()][64] fp32 partial sums for ecgmm_bn_finalize (replaces the statistics pass of
 * nn.BatchNorm2d(64) in training mode, torchvision resnet.py:197-198,268-269).  rows == 0: not offered. */
int ecgmm_stem_conv_fwd_stats_rows(int N, int H, int W);
int ecgmm_stem_conv_fwd_stats(const ecgmm_bf16* xs, const ecgmm_bf16* w_s2d, ecgmm_bf16* y, float* psum, float* psq,
                              int N, int H, int W, void* stream);
/* dw [64][3][7][7] fp32 += sum over pixels.  With a workspace of ecgmm_stem_conv_wgrad_workspace() bytes the per-CTA
 * partial sums are folded in a fixed order (bit-reproducible); workspace == NULL: fp32 atomics. */
long long ecgmm_stem_conv_wgrad_workspace(int N, int H, int W);
int ecgmm_stem_conv_wgrad(const ecgmm_bf16* xs, const ecgmm_bf16* dy, float* dw, int N, int H, int W,
                          void* workspace, long long workspace_bytes, void* stream);

/* ------------------------------------------------------------------ implicit-GEMM convolutions
 * Replaces nn.Conv2d in torchvision BasicBlock (resnet.py:59-104) and nn.Conv1d in
 * BasicBlock1D (multimodal_paper_modal_balance.py:67-93; 1-D = H == 1, R == 1).
 * Supported: Cin % 64 == 0, Cout % 64 == 0, (R,S,stride,pad) in
 *   {(3,3,1,1), (3,3,2,1), (1,1,1,0), (1,1,2,0), (1,3,1,1), (1,3,2,1)}.
 * x [N][H][W][Cin], y/dy [N][Ho][Wo][Cout], tcgen05.mma with fp32 accumulation in TMEM. */
int ecgmm_conv2d_fwd(const ecgmm_bf16* x, const ecgmm_bf16* w_fwd, ecgmm_bf16* y, int N, int H, int W, int Cin,
                     int Cout, int R, int S, int stride, int padH, int padW, void* stream);
/* Same convolution, and the BatchNorm batch statistics of its output produced by the epilogue (no second pass
 * over y): psum / psq [rows][Cout] fp32 receive, per (CTA, epilogue warp) of the kernel, the sum and the sum of
 * squares of every output channel over that warp's pixels -- of the bf16 values as stored.  rows =
 * ecgmm_conv2d_fwd_stats_rows(same shape), 0 when the library does not offer it for the shape (1x1 convolutions and
 * the 64->64 layers: their tiles hold too few MMAs to hide the reduction; use ecgmm_chan_stats); ecgmm_bn_finalize(psum, psq, rows, 1, Cout, N*Ho*Wo, ...) folds them
 * (replaces ecgmm_chan_stats behind torchvision resnet.py:92-97 bn1/bn2 in training mode). */
int ecgmm_conv2d_fwd_stats_rows(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int padH, int padW);
int ecgmm_conv2d_fwd_stats(const ecgmm_bf16* x, const ecgmm_bf16* w_fwd, ecgmm_bf16* y, float* psum, float* psq, int N,
                           int H, int W, int Cin, int Cout, int R, int S, int stride, int padH, int padW,
                           void* stream);
/* dx = conv_transpose(dy, w).  accumulate != 0 adds into dx instead of overwriting. */
int ecgmm_conv2d_dgrad(const ecgmm_bf16* dy, const ecgmm_bf16* w_dgrad, ecgmm_bf16* dx, int N, int H, int W,
                       int Cin, int Cout, int R, int S, int stride, int padH, int padW, int accumulate,
                       void* stream);
/* The same data gradient + the reduction pass of the BatchNorm(+ReLU) backward that consumes dx as its upstream
 * gradient (torch: batch_norm_backward's sum(dy), sum(dy * xhat)), computed in the epilogue while the dx tile is on
 * chip: bn_x [N][H][W][Cin] is that BatchNorm's input, bn_mask its ReLU decisions (one byte per 8 channels; NULL: no
 * ReLU), bn_mean / bn_invstd [Cin] its batch statistics; p1 / p2 [ecgmm_conv2d_dgrad_reduce_rows()][Cin] receive
 * partial sums of dz and dz * xhat (dz = dx * relu') for ecgmm_bn_bwd_finalize(count = N*H*W).  rows == 0: not
 * offered for the shape -- use ecgmm_conv2d_dgrad + ecgmm_bn_bwd_reduce. */
int ecgmm_conv2d_dgrad_reduce_rows(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int padH, int padW);
int ecgmm_conv2d_dgrad_reduce(const ecgmm_bf16* dy, const ecgmm_bf16* w_dgrad, ecgmm_bf16* dx, const ecgmm_bf16* bn_x,
                              const uint8_t* bn_mask, const float* bn_mean, const float* bn_invstd, float* p1,
                              float* p2, int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int padH,
                              int padW, int accumulate, void* stream);
/* dw (fp32, OIHW) += x^T * dy.  The pixel range is split across CTAs (split-K); with a workspace of at
 * least ecgmm_conv2d_wgrad_workspace() bytes the partials are combined by a second deterministic kernel,
 * otherwise (workspace NULL / too small / shape not covered, query returns 0) by fp32 atomics. */
long long ecgmm_conv2d_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int padH,
                                       int padW);
int ecgmm_conv2d_wgrad(const ecgmm_bf16* x, const ecgmm_bf16* dy, float* dw_oihw, int N, int H, int W, int Cin,
                       int Cout, int R, int S, int stride, int padH, int padW, void* workspace,
                       long long workspace_bytes, void* stream);

/* ------------------------------------------------------------------ BatchNorm / ReLU / pooling
 * Replaces nn.BatchNorm2d + ReLU + residual add + MaxPool2d of torchvision resnet18
 * (resnet.py:198-200, 92-104; reached from multimodal_paper_modal_balance.py:325) and
 * nn.BatchNorm1d + ReLU + MaxPool1d + SEBlock scaling of ResNet1D_SE
 * (multimodal_paper_modal_balance.py:49-64, 83-93, 99-104).
 * Activations [N][P][C] bf16 (P = H*W or L), C = 8 * 2^k.  Reductions are two-stage and
 * deterministic: a (split, N) grid writes partial rows [N][split][C], a finalize kernel folds
 * them in fp64.  Use ecgmm_reduce_split to size the partial buffers. */

/* number of pixel slabs per sample the reduction kernels should use for this shape */
int ecgmm_reduce_split(int N, int P, int C);
/* psum / psq [N][split][C] = per-slab sum and sum of squares of x */
int ecgmm_chan_stats(const ecgmm_bf16* x, float* psum, float* psq, int N, int P, int C, int split, void* stream);
/* Training-mode BatchNorm statistics from the partials (count = N*P elements per channel):
 * mean, invstd (saved for backward), scale = gamma*invstd, shift = beta - mean*scale,
 * running_mean/var update (momentum, unbiased variance; conv_bias, if given, is the bias of the
 * preceding convolution which the conv kernels do not add: it only moves running_mean),
 * num_batches += 1, and optionally nsum [N][C] = per-sample sums (the SE squeeze).
 * gamma/beta/conv_bias/running_x/num_batches/nsum may be NULL. */
int ecgmm_bn_finalize(const float* psum, const float* psq, int N, int split, int C, long long count,
                      const float* gamma, const float* beta, const float* conv_bias, float eps, float momentum,
                      float* running_mean, float* running_var, long long* num_batches, float* mean, float* invstd,
                      float* scale, float* shift, float* nsum, void* stream);
/* eval mode: scale = gamma/sqrt(running_var+eps), shift = beta + (conv_bias - running_mean)*scale */
int ecgmm_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* conv_bias,
                         const float* running_mean, const float* running_var, float eps, float* scale, float* shift,
                         void* stream);
/* y = act((x*scale[c] + shift[c]) * se[n][c] + res); se (fp32 [N][C]) and res may be NULL.
 * relu_mask (may be NULL; only with relu) receives one byte per 8 channels, bit j = (y[8g+j] > 0):
 * the backward kernels (mode 3) read it instead of y, 1/16 of the bytes. */
int ecgmm_bn_apply(const ecgmm_bf16* x, const float* scale, const float* shift, const float* se,
                   const ecgmm_bf16* res, ecgmm_bf16* y, uint8_t* relu_mask, int N, int P, int C, int relu,
                   void* stream);
/* stem: y [N][Ho][Wo][C] = maxpool3x3/s2/p1(relu(x*scale+shift)); argmax (may be NULL) [N][Ho][Wo][C] u8 holds the
 * window position 0..8 of the first maximum.  H == 1 gives MaxPool1d(3,2,1). */
int ecgmm_bn_relu_maxpool(const ecgmm_bf16* x, const float* scale, const float* shift, ecgmm_bf16* y,
                          uint8_t* argmax, int N, int H, int W, int C, void* stream);
/* Backward partials p1/p2 [N][split][C] = sum dz, sum dz*xhat with
 *   mode 0: dz = dy;  mode 1: dz = dy * (y > 0);  mode 2: stem -- dy is the gradient of the POOLED
 *   output [N][Ho][Wo][C], routed through argmax and gated by relu(x*scale+shift) > 0;
 *   mode 3: dz = dy * bit, the `argmax` argument carrying the relu_mask written by ecgmm_bn_apply;
 *   mode 4: the stem sums evaluated in the POOLED domain (4x fewer elements, no gather): x is the pooled
 *           output y = maxpool(relu(bn(.))), dy its gradient, H x W the pooled size, `mean` := beta and
 *           `invstd` := gamma, so that dz = dy * (y > 0) and xhat = (y - beta) / gamma is the normalised
 *           value of the one pre-pool element each pooled gradient is routed to. */
int ecgmm_bn_bwd_reduce(const ecgmm_bf16* x, const ecgmm_bf16* dy, const ecgmm_bf16* y, const uint8_t* argmax,
                        const float* mean, const float* invstd, const float* scale, const float* shift, float* p1,
                        float* p2, int N, int H, int W, int C, int split, int mode, void* stream);
/* Folds the partials into dgamma, dbeta and the coefficients of
 *   dx = A*se*dz + B*x + D + A*q.   se/q/nsum (all [N][C], NULL outside SE blocks) describe
 *   du = dz*se + q, the gradient reaching the BatchNorm output of an SE block. */
int ecgmm_bn_bwd_finalize(const float* p1, const float* p2, int N, int split, int C, long long per_sample,
                          const float* gamma, const float* mean, const float* invstd, const float* se,
                          const float* q, const float* nsum, float* dgamma, float* dbeta, float* coefA,
                          float* coefB, float* coefD, long long count, void* stream);
/* (p1 / p2 hold N * split rows; the mean is taken over count elements per channel, or over N * per_sample when count
 *  is 0 -- partials written by ecgmm_conv2d_dgrad_reduce have one row per CTA, not per sample: pass count = N*H*W) */
/* dx (gradient of the convolution output) and optionally dz_out = the masked upstream gradient
 * (what flows into the residual branch). */
int ecgmm_bn_bwd_apply(const ecgmm_bf16* x, const ecgmm_bf16* dy, const ecgmm_bf16* y, const uint8_t* argmax,
                       const float* coefA, const float* coefB, const float* coefD, const float* scale,
                       const float* shift, const float* se, const float* q, ecgmm_bf16* dx, ecgmm_bf16* dz_out,
                       int N, int H, int W, int C, int mode, void* stream);
/* AdaptiveAvgPool to 1x1 (resnet.py:203, multimodal_paper_modal_balance.py:110): [N][P][C] bf16 <-> [N][C] fp32 */
int ecgmm_avgpool_fwd(const ecgmm_bf16* x, float* out, int N, int P, int C, void* stream);
int ecgmm_avgpool_bwd(const float* dout, ecgmm_bf16* dx, int N, int P, int C, void* stream);

/* ------------------------------------------------------------------ 1-D ResNet-SE stem
 * nn.Conv1d(Cin, 64, 7, stride 2, padding 3) of ResNet1D_SE.initial
 * (multimodal_paper_modal_balance.py:99-100; 12-lead: train_signal_12_af.py:184).
 * x [B][Cin][L] fp32 (Cin <= 16), w [64][Cin][7] fp32, y/dy [B][Lo][64] bf16, Lo = (L-1)/2+1.
 * The bias is NOT added (see ecgmm_bn_finalize / ecgmm_bn_eval_coeffs). dw accumulates. */
int ecgmm_signal_stem_fwd(const float* x, const float* w, ecgmm_bf16* y, int B, int Cin, int L, void* stream);
/* The same layer through the tensor-core kernels: xs4 [B][Lq][64] bf16 groups 4 consecutive samples of all leads
 * (channel j*Cin + ci = x[b][ci][4q + j]); with w4 [128][1][3][64] (ecgmm_signal_stem_w4) the stem is
 * ecgmm_conv2d_fwd(xs4 as [B][1][Lq][64], w4, stride 1) -> [B][1][Lq][128] == [B][2 Lq][64], its weight gradient
 * ecgmm_conv2d_wgrad into dw4 [128][64][1][3] followed by ecgmm_signal_stem_dw4_fold (dw [64][Cin][7] += ...).
 * ecgmm_signal_s4d_len(L) = Lq = ceil(L/4) when 2*Lq equals the stem's output length (L mod 4 in {0,3}), else 0. */
int ecgmm_signal_s4d_len(int L);
int ecgmm_signal_s4d(const float* x, ecgmm_bf16* xs4, int B, int Cin, int L, void* stream);
int ecgmm_signal_stem_w4(const float* w, ecgmm_bf16* w4, int Cin, void* stream);
int ecgmm_signal_stem_dw4_fold(const float* dw4, float* dw, int Cin, void* stream);
/* workspace of ecgmm_signal_stem_wgrad_workspace() bytes: fixed-order fold of the per-CTA partial sums; NULL: atomics */
long long ecgmm_signal_stem_wgrad_workspace(int B, int Cin, int L);
int ecgmm_signal_stem_wgrad(const float* x, const ecgmm_bf16* dy, float* dw, int B, int Cin, int L, void* workspace,
                            long long workspace_bytes, void* stream);
/* SEBlock.fc (multimodal_paper_modal_balance.py:52-64) on pooled = scale*nsum/L + shift:
 * hid = relu(W1 pooled + b1) [N][R], gate = sigmoid(W2 hid + b2) [N][C];  w1 [R][C], w2 [C][R]. */
int ecgmm_se_fwd(const float* nsum, const float* scale, const float* shift, const float* w1, const float* b1,
                 const float* w2, const float* b2, float* pooled, float* hid, float* gate, int N, int C, int R, int L,
                 void* stream);
/* from the bn_bwd_reduce partials of the block output: q [N][C] (see ecgmm_bn_bwd_finalize) and the
 * pre-activation gradients dpre2 [N][C], dpre1 [N][R] whose outer products give the SE weight grads. */
int ecgmm_se_bwd(const float* p1, const float* p2, int split, const float* gamma, const float* beta,
                 const float* w1, const float* w2, const float* hid, const float* gate, float* q, float* dpre2,
                 float* dpre1, int N, int C, int R, int L, void* stream);

/* ------------------------------------------------------------------ dense tails and fusion head (fp32)
 * nn.Linear / nn.LayerNorm / AttentionFusion / var regulariser / losses:
 * multimodal_paper_modal_balance.py:31-46, 221-223, 256-289, 340-352; train.py:31,72,78;
 * signal_model.py:91-106 (FocalLoss), 203-206 (z-score). */
/* C[M][N] (+)= op(A) op(B) (+ bias[N]) (relu).  transA: A stored [K][M]; transB: B stored [N][K]. */
int ecgmm_sgemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int transA,
                int transB, int accumulate, int relu, void* stream);
/* out[n] (+)= sum_m X[m][n] */
int ecgmm_colsum(const float* X, float* out, int M, int N, int accumulate, void* stream);
int ecgmm_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                        int rows, int D, float eps, void* stream);
/* any of dx / dgamma / dbeta may be NULL */
int ecgmm_layernorm_bwd(const float* x, const float* dy, const float* gamma, const float* mean, const float* rstd,
                        float* dx, float* dgamma, float* dbeta, int rows, int D, int accumulate_dx, void* stream);
/* fused [B][D0+D1+D2] = cat(w0*f0, w1*f1, w2*f2), w = softmax(weights[3]) (also stored in soft_w) */
int ecgmm_fusion_gate_fwd(const float* f0, const float* f1, const float* f2, const float* weights, float* fused,
                          float* soft_w, int B, int D0, int D1, int D2, void* stream);
int ecgmm_fusion_gate_bwd(const float* dfused, const float* f0, const float* f1, const float* f2,
                          const float* weights, float* df0, float* df1, float* df2, float* dweights, int B, int D0,
                          int D1, int D2, int accumulate, void* stream);
/* loss = sum_{i<j} |v_i - v_j|, v_i = mean_b var_unbiased_d(f_i); row_mean [3][B], coef [3] = dloss/dv_i */
int ecgmm_var_loss_fwd(const float* f0, const float* f1, const float* f2, float* loss, float* row_mean, float* coef,
                       int B, int D0, int D1, int D2, void* stream);
/* one modality: df (+)= gout[0]*coef[0]*2(f-row_mean)/((D-1)B) */
int ecgmm_var_loss_bwd(const float* f, const float* row_mean, const float* coef, const float* gout, float* df, int B,
                       int D, int accumulate, void* stream);
/* mean softmax cross entropy (focal == 0) or focal loss alpha(1-pt)^gamma ce (focal == 1);
 * dlogits (may be NULL) = gscale * dloss/dlogits.  Rows labelled ignore_index are skipped as torch does (zero gradient;
 * CrossEntropyLoss averages over the remaining rows, the focal loss over all rows).  Any other label outside [0, C):
 * *bad_label (may be NULL) is set to 1, the row's gradient is zero and the loss is NaN (torch: device-side assert). */
int ecgmm_ce_loss(const float* logits, const long long* labels, float* loss, float* dlogits, int B, int C, int focal,
                  float alpha, float gamma, float gscale, long long ignore_index, int* bad_label, void* stream);
/* y = x*mask; mask_in given (0 or 1/(1-p)) or drawn from (seed, element index); mask_out may be NULL.
 * seed_dev (may be NULL): device word added to seed at run time -- the per-step offset of a captured CUDA graph */
int ecgmm_dropout_fwd(const float* x, const float* mask_in, float* y, float* mask_out, long long n, float p,
                      unsigned long long seed, const unsigned long long* seed_dev, void* stream);
/* dx = dy * mask * (y > 0); y and mask may be NULL */
int ecgmm_mask_bwd(const float* dy, const float* y, const float* mask, float* dx, long long n, void* stream);
/* BatchNorm1d (+ReLU) over a small fp32 matrix [B][C]: clinical MLP, multimodal_paper_modal_balance.py:258.
 * train != 0: batch statistics + running update (+ num_batches); else running statistics. */
int ecgmm_bn_rows_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      long long* num_batches, float* y, float* mean, float* invstd, int B, int C, float eps,
                      float momentum, int train, int relu, void* stream);
int ecgmm_bn_rows_bwd(const float* x, const float* dy, const float* y, const float* gamma, const float* mean,
                      const float* invstd, float* dx, float* dgamma, float* dbeta, int B, int C, int relu,
                      void* stream);
int ecgmm_zscore(const float* x, float* y, long long rows, int L, float eps, void* stream);

/* ------------------------------------------------------------------ signal preprocessing
 * What the reference does per sample on the host inside Dataset.__getitem__ (dataset.py:76-95; identical
 * copies signal_model.py:203-224, evaluation_signal.py:20-39), batched over `rows` signals of length L:
 *   window > 0 : remove_baseline_drift -- x - np.convolve(x, ones(window)/window, mode='same')   dataset.py:81-83
 *   order  > 0 : lowpass_filter -- scipy.signal.butter(order, wn, 'low') + filtfilt(b, a, x)
 *                (odd extension by 3*(order+1) samples, lfilter_zi initial state)              dataset.py:85-89
 *   zscore     : z_score_normalize -- (x - mean) / (std_population + eps)                      dataset.py:76-79
 * in that order, all in float64 like the reference; the float32 result is what dataset.py:68 builds.
 * x: [rows][L] float32 (x_is_f64 = 0) or float64 (1); y: [rows][L] float32; wn = cutoff / (0.5 * fs).
 * workspace: ecgmm_signal_preprocess_workspace(rows, L, order) bytes of device scratch (8-byte aligned).
 * ecgmm_butter_lowpass is the HOST-side filter design used by the launch (b, a: order+1 doubles, zi: order
 * doubles, host pointers); it touches no device and is exported for tests / callers that want the taps. */
int ecgmm_butter_lowpass(int order, double wn, double* b, double* a, double* zi);
long long ecgmm_signal_preprocess_workspace(long long rows, int L, int order);
int ecgmm_signal_preprocess(const void* x, int x_is_f64, float* y, void* workspace, long long workspace_bytes,
                            long long rows, int L, int window, int order, double wn, int zscore, double eps,
                            void* stream);

/* ------------------------------------------------------------------ batched perturbation inference
 * BASELINE.json configs[3] / SURVEY.md section 8d cfg4: V masked variants per sample of the fused embedding,
 *     variant[s][v][d] = masks[v][d] ? e[s][d] : bg[d]            (masks: V x D bytes, shared by all samples)
 * through fusion_classifier (multimodal_paper_modal_balance.py:283-289; driven by
 * shap_fusion_modal_balance.py:135,159 and lime_fusion_modal_balance.py:118-160) in eval mode.
 *   ecgmm_perturb_build : e [S][D] fp32, bg [D] fp32, masks [V][D] uint8 -> variants [S*V][D] bf16  (D % 8 == 0)
 *   (GEMM)              : ecgmm_conv2d_fwd(variants as N=1,H=1,W=S*V,Cin=D; w_fwd = Linear(D,HID).weight as
 *                         [HID][1][1][D] bf16) -> hidden [S*V][HID] bf16 on the tcgen05 implicit-GEMM kernel
 *   ecgmm_head_tail     : relu(hidden + b1) -> Linear(HID, C) (w2 [C][HID], b2 [C], fp32) ->
 *                         out[row] = softmax(logits)[cls]  (cls >= 0)   or   out[row][0..C) = logits (cls < 0) */
int ecgmm_perturb_build(const float* e, const float* bg, const uint8_t* masks, ecgmm_bf16* variants, long long S,
                        int V, int D, void* stream);
int ecgmm_head_tail(const ecgmm_bf16* hidden, const float* b1, const float* w2, const float* b2, float* out,
                    long long rows, int HID, int C, int cls, void* stream);
/* The same path in ONE kernel (csrc/perturb_fused.cu): the masked variants are built by producer warps straight into the
 * SWIZZLE_128B shared-memory operand of the tcgen05 GEMM (W1 resident in shared memory), the epilogue applies
 * bias / ReLU / Linear(HID, C) / softmax from TMEM: no variant or hidden tensor in HBM.  e and bg are bf16 here
 * (ecgmm_f32_to_bf16; the selection between them is exact); the masks are packed once per call to one bit per element
 * (ecgmm_perturb_pack_masks: masks [V][D] bytes, 16-byte aligned, D % 64 == 0 -> V * D / 32 uint32 words laid out
 * [D/64 chunks][V][2], in the bit order the kernel expands).  Covered: HID == 128, D % 64 == 0, D <= 768, C <= 8
 * (ecgmm_perturb_head_fused_supported returns 1); w1 [128][D] bf16, k contiguous.  b1 / w2 / b2 are copied to constant
 * memory on the launch's stream: calls with DIFFERENT heads must not run concurrently on two streams of one device.
 *   out [S][V] = softmax(logits)[cls] (cls >= 0)   or   out [S][V][C] = logits (cls < 0) */
int ecgmm_perturb_head_fused_supported(int D, int HID, int C);
int ecgmm_perturb_pack_masks(const uint8_t* masks, uint32_t* bits, int V, int D, void* stream);
int ecgmm_perturb_head_fused(const ecgmm_bf16* e, const ecgmm_bf16* bg, const uint32_t* bits, const ecgmm_bf16* w1,
                             const float* b1, const float* w2, const float* b2, float* out, long long S, int V, int D,
                             int C, int cls, void* stream);
int ecgmm_f32_to_bf16(const float* x, ecgmm_bf16* y, long long n, void* stream);

/* ------------------------------------------------------------------ expected gradients over the fusion head
 * SURVEY.md section 8f rank 3.  shap_fusion_modal_balance.py:135,159 explains FusionClassifierWrapper (the logits of
 * fusion_classifier, multimodal_paper_modal_balance.py:283-289) with shap.GradientExplainer = expected gradients;
 * `shap` is unpinned and absent, so the estimator is restated (ecgmm/explain.py, oracle/model.py) with an EXPLICIT
 * sampling plan: for sample s and draw k a background row idx[s*K+k] and an interpolation weight alpha[s*K+k]:
 *     phi[s][d][c] = 1/K * sum_k (e[s][d] - bg[j][d]) * dlogit_c/dx_d (bg[j] + alpha (e[s] - bg[j]))
 * Everything is fp32 (the sign of the hidden pre-activations decides the gradient).
 *   ecgmm_eg_points : e [S][D], bg [NB][D], idx [S*K] int32, alpha [S*K] -> points [S*K][D]          (D % 4 == 0)
 *   (GEMM)          : ecgmm_sgemm  hidden = relu(points W1^T + b1)  [S*K][HID]
 *   ecgmm_eg_gate   : gate[c][r][h] = hidden[r][h] > 0 ? w2[c][h] : 0      [C][S*K][HID]   (HID % 4 == 0, C <= 8)
 *   (GEMM)          : ecgmm_sgemm  grad[c] = gate[c] W1   [S*K][D] per class, stored [C][S*K][D]
 *   ecgmm_eg_reduce : phi [S][D][C] as above
 *   ecgmm_modality_share : share[s][c][m] = 100 * mean_{d in modality m} |phi[s][d][c]| / sum over the 3 modalities
 *                          (shap_fusion_modal_balance.py:177-200; 0 when the three means are all 0); use_sum: see below */
int ecgmm_eg_points(const float* e, const float* bg, const int* idx, const float* alpha, float* points, long long S,
                    int K, int D, int NB, void* stream);
int ecgmm_eg_gate(const float* hidden, const float* w2, float* gate, long long rows, int HID, int C, void* stream);
int ecgmm_eg_reduce(const float* e, const float* bg, const int* idx, const float* grad, float* phi, long long S, int K,
                    int D, int C, int NB, void* stream);
int ecgmm_modality_share(const float* phi, float* share, long long S, int C, int D0, int D1, int D2, int use_sum,
                         void* stream);

/* ------------------------------------------------------------------ local surrogate regression (LIME / KernelSHAP)
 * lime_fusion_modal_balance.py:158-160 fits, per explained instance, lime's default regressor -- sklearn
 * Ridge(alpha=1, fit_intercept=True) with the kernel weights as sample_weight -- to the model's responses on perturbed
 * rows; KernelSHAP is the same weighted least squares with Shapley-kernel weights.  On the binary keep-masks of
 * ecgmm_perturb_build the fit is linear in the responses f [V]:  (w, b) = R f  with
 *     zbar = sum pi z / sum pi,  Zc = Z - zbar,  w = (Zc^T diag(pi) Zc + alpha I)^-1 Zc^T diag(pi) f,
 *     b = pi.f / sum pi - zbar.w
 * ecgmm_ridge_operator designs R [(D+1)][V] (row D = intercept) ON THE HOST in float64 -- it depends only on the
 * sampling plan, like the filter taps of ecgmm_butter_lowpass; all pointers are HOST pointers, no device is touched.
 * Per sample the device then computes coefficients [S][D+1] = f [S][V] R^T with ecgmm_sgemm.
 * ecgmm_modality_share(use_sum = 1) aggregates |w| per modality by SUM (lime_fusion_modal_balance.py:163-175); use_sum
 * = 0 by MEAN (shap_fusion_modal_balance.py:189-200). */
int ecgmm_ridge_operator(const uint8_t* masks, const double* weights, int V, int D, double alpha, float* R);

/* ------------------------------------------------------------------ serving helpers (image-only endpoint, Grad-CAM)
 * the endpoint the mobile app posts to (Groove/components/SubmitButton.tsx:44-45) has no
 * server code in the reference; what is replaced is the eval-mode chain image_encoder -> image_norm ->
 * image_classifier (multimodal_paper_modal_balance.py:325-327,337) plus softmax / argmax, and the Grad-CAM map of
 * the last ResNet stage (artefacts under gpt/, generator not in the repository).
 *   ecgmm_softmax_rows : logits [rows][C] -> probs [rows][C] and/or argmax [rows] int32 (either may be NULL)
 *   ecgmm_gather_rows  : out[r][:] = table[idx[r]][:]   (table [NT][D] fp32, idx int32) -- the classifier row of each
 *                        sample's class is d logit / d feature
 *   ecgmm_gradcam      : act [N][P][C] bf16 (layer4 output, P = h*w pixels), g [N][C] fp32 (d logit / d pooled) ->
 *                        cam[n][p] = relu(scale * sum_c g[n][c] act[n][p][c]),  scale = 1/P for the average pool
 *                        (C % 8 == 0) */
/* Forward convolution with the folded-BatchNorm epilogue of the serving path (eval mode only: scale / shift come from
 * the running statistics, ecgmm_bn_eval_coeffs): replaces Conv2d -> BatchNorm2d [-> += identity] [-> ReLU] of a
 * torchvision BasicBlock (resnet.py:59-104) by ONE kernel,
 *     y[n][h][w][c] = act( conv(x, w)[n][h][w][c] * scale[c] + shift[c] + res[n][h][w][c] ),
 * with no bf16 rounding between the convolution and the affine map.  res may be NULL; relu != 0 applies ReLU.
 * Same shapes / kernels as ecgmm_conv2d_fwd; scale, shift [Cout] fp32 and res (layout of y) 16-byte aligned.
 * ecgmm.serve uses it by default (ECGMM_SERVE_FUSED=0 selects separate convolution and BatchNorm kernels). */
int ecgmm_conv2d_fwd_bn(const ecgmm_bf16* x, const ecgmm_bf16* w_fwd, ecgmm_bf16* y, const float* scale,
                        const float* shift, const ecgmm_bf16* res, int relu, int N, int H, int W, int Cin, int Cout,
                        int R, int S, int stride, int padH, int padW, void* stream);
int ecgmm_softmax_rows(const float* logits, float* probs, int* argmax, long long rows, int C, void* stream);
int ecgmm_gather_rows(const float* table, const int* idx, float* out, long long rows, int D, int NT, void* stream);
int ecgmm_gradcam(const ecgmm_bf16* act, const float* g, float* cam, int N, int P, int C, float scale, void* stream);

/* ------------------------------------------------------------------ optimizer
 * torch.optim.Adam step (train.py:43,81).  chunk_table: device array of
 * { float* p; const float* g; float* m; float* v; long long n; } (ecgmm_adam_chunk_bytes() each);
 * step is the 1-based step count; gradients are multiplied by grad_scale first. */
int ecgmm_adam_chunk_bytes(void);
int ecgmm_adam_step(const void* chunk_table, int n_chunks, float lr, float beta1, float beta2, float eps,
                    float weight_decay, long long step, float grad_scale, void* stream);
/* The same step with the learning rate and the step count read from DEVICE memory, for replay inside a CUDA graph
 * (kernel arguments are frozen at capture; train.py:158-161 and OneCycleLR change lr between steps).
 * ecgmm_step_advance(state): state[0] += 1 (the step count ecgmm_adam_step_dev reads), state[1] += odd constant (the
 * seed_dev word of ecgmm_dropout_fwd); captured at the head of the graph so that every replay sees fresh values. */
int ecgmm_step_advance(long long* state, void* stream);
int ecgmm_adam_step_dev(const void* chunk_table, int n_chunks, const float* lr_dev, float beta1, float beta2, float eps,
                        float weight_decay, const long long* step_dev, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ECGMM_H_ */
