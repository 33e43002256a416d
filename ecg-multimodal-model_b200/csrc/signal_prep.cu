// Per-signal ECG preprocessing on the GPU: what the reference does with
// numpy / scipy inside Dataset.__getitem__, one signal at a time on the host (dataset.py:76-95, identical
// copies in signal_model.py:203-224 and evaluation_signal.py:20-39):
//
//   remove_baseline_drift : x - moving_average_200(x)                  (np.convolve(..., mode='same'))
//   lowpass_filter        : Butterworth low-pass, zero-phase            (scipy.signal.butter + filtfilt)
//   z_score_normalize     : (x - mean) / (std_population + 1e-8)        (optional; commented out at dataset.py:92)
//
// All arithmetic is float64, like the reference.  An IIR filter is a serial recurrence along time, so the
// parallel axis is the signal: one thread per signal (B x leads signals per batch), 32 signals per CTA.
// The float64 scratch row of a signal lives TIME-MAJOR in the caller's workspace (ws[t][signal]), so the 32
// lanes of a warp touch 32 consecutive doubles at every time step (coalesced); the recurrence itself is two
// dependent DFMAs per sample (direct form II transposed); with ~1 warp per SM the L2 latency of the samples has to be
// hidden by software prefetch (iir_pass).
//
// Host side: the Butterworth design (analog prototype -> pre-warp -> bilinear transform -> polynomial
// expansion) and the steady-state initial condition of filtfilt (lfilter_zi) are computed here in double
// precision; no scipy on the product path.
#include "common.h"

#include <math.h>
#include <stdlib.h>

#include <complex>

namespace ecgmm {

constexpr int kMaxOrder = 8;

struct IirCoef {
  double b[kMaxOrder + 1];
  double a[kMaxOrder + 1];
  double zi[kMaxOrder];
};

// One DF2T step: y = b0*x + z0;  z[k] = b[k+1]*x + z[k+1] - a[k+1]*y
template <int ORDER>
__device__ __forceinline__ double iir_step(const IirCoef& c, double (&z)[kMaxOrder], double x) {
  const double y = fma(c.b[0], x, z[0]);
#pragma unroll
  for (int k = 0; k < ORDER - 1; ++k) z[k] = fma(-c.a[k + 1], y, fma(c.b[k + 1], x, z[k + 1]));
  z[ORDER - 1] = fma(-c.a[ORDER], y, c.b[ORDER] * x);
  return y;
}

// One filter pass over n samples of a time-major scratch row, forward (t0, t0+1, ...) or backward (t0, t0-1, ...).
// A signal is one thread and there is about one warp per SM, so nothing hides the ~700-clock L2 latency of the next
// sample unless it is in flight early: samples are fetched in chunks of CH, the NEXT chunk is issued before the
// current one is filtered (16 samples x 2 dependent DFMA each ~ one latency).  emit(t, y) consumes the output.
template <int ORDER, bool REVERSE, typename Emit>
__device__ __forceinline__ void iir_pass(const double* w, long long ld, int t0, int n, const IirCoef& c,
                                         double (&z)[kMaxOrder], Emit&& emit) {
  constexpr int CH = 16;
  double cur[CH], nxt[CH];
  auto at = [&](int k) { return REVERSE ? t0 - k : t0 + k; };
#pragma unroll
  for (int k = 0; k < CH; ++k) cur[k] = k < n ? w[(long long)at(k) * ld] : 0.0;
  for (int done = 0; done < n; done += CH) {
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int kk = done + CH + k;
      nxt[k] = kk < n ? w[(long long)at(kk) * ld] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < CH; ++k)
      if (done + k < n) emit(at(done + k), iir_step<ORDER>(c, z, cur[k]));
#pragma unroll
    for (int k = 0; k < CH; ++k) cur[k] = nxt[k];
  }
}

// ws: [L + 2*edge][ld] doubles, ld = number of signals rounded up to 32.
template <typename TIn, int ORDER>
__global__ void __launch_bounds__(32) signal_preprocess_kernel(const TIn* __restrict__ x, float* __restrict__ out,
                                                               double* __restrict__ ws, long long rows,
                                                               long long ld, int L, int window,
                                                               const __grid_constant__ IirCoef c, int zscore,
                                                               double eps) {
  const long long row = (long long)blockIdx.x * 32 + threadIdx.x;
  if (row >= rows) return;
  const TIn* xr = x + row * L;
  float* outr = out + row * L;
  double* w = ws + row;
  constexpr int edge = ORDER > 0 ? 3 * (ORDER + 1) : 0;  // filtfilt's default padlen = 3 * max(len(a), len(b))
  // ---- pass 1: baseline removal (or a plain copy) into w[edge .. edge+L)
  if (window > 0) {
    // np.convolve(x, ones(W)/W, 'same')[i] = (1/W) * sum x[i-lo .. i+hi], zeros outside, lo = W/2, hi = W-lo-1
    const int lo = window / 2, hi = window - lo - 1;
    const double inv = 1.0 / (double)window;
    double S = 0.0;
    for (int t = 0; t <= hi && t < L; ++t) S += (double)xr[t];
    constexpr int CH = 8;  // the running sum is the only dependence: 3 x CH independent loads per trip
    for (int i0 = 0; i0 < L; i0 += CH) {
      double xa[CH], xin[CH], xout[CH];
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int i = i0 + k;
        xa[k] = i < L ? (double)xr[i] : 0.0;
        xin[k] = (i + 1 + hi < L) ? (double)xr[i + 1 + hi] : 0.0;
        xout[k] = (i - lo >= 0 && i < L) ? (double)xr[i - lo] : 0.0;
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int i = i0 + k;
        if (i < L) {
          w[(long long)(edge + i) * ld] = xa[k] - S * inv;
          S += xin[k];
          S -= xout[k];
        }
      }
    }
  } else {
#pragma unroll 8
    for (int i = 0; i < L; ++i) w[(long long)(edge + i) * ld] = (double)xr[i];
  }
  double s1 = 0.0;
  if constexpr (ORDER > 0) {
    // ---- odd extension by `edge` samples on both sides
    const double y0 = w[(long long)edge * ld], yl = w[(long long)(edge + L - 1) * ld];
    for (int j = 1; j <= edge; ++j) {
      w[(long long)(edge - j) * ld] = 2.0 * y0 - w[(long long)(edge + j) * ld];
      w[(long long)(edge + L - 1 + j) * ld] = 2.0 * yl - w[(long long)(edge + L - 1 - j) * ld];
    }
    const int n = L + 2 * edge;
    double z[kMaxOrder];
    // ---- forward pass, in place
    const double x0 = w[0];
#pragma unroll
    for (int k = 0; k < ORDER; ++k) z[k] = c.zi[k] * x0;
    iir_pass<ORDER, false>(w, ld, 0, n, c, z, [&](int t, double y) { w[(long long)t * ld] = y; });
    // ---- backward pass from the end down to `edge`; the un-padded part is the result
    const double xl = w[(long long)(n - 1) * ld];
#pragma unroll
    for (int k = 0; k < ORDER; ++k) z[k] = c.zi[k] * xl;
    iir_pass<ORDER, true>(w, ld, n - 1, n - edge, c, z, [&](int t, double y) {
      if (t < edge + L) {
        if (zscore) {
          w[(long long)t * ld] = y;
          s1 += y;
        } else {
          outr[t - edge] = (float)y;
        }
      }
    });
  } else {
    if (!zscore) {
#pragma unroll 8
      for (int i = 0; i < L; ++i) outr[i] = (float)w[(long long)(edge + i) * ld];
    } else {
#pragma unroll 8
      for (int i = 0; i < L; ++i) s1 += w[(long long)(edge + i) * ld];
    }
  }
  if (zscore) {
    const double mean = s1 / (double)L;
    double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;  // independent chains: the loads, not the adds, should pace this
    int i = 0;
    for (; i + 3 < L; i += 4) {
      const double d0 = w[(long long)(edge + i) * ld] - mean, d1 = w[(long long)(edge + i + 1) * ld] - mean;
      const double d2 = w[(long long)(edge + i + 2) * ld] - mean, d3 = w[(long long)(edge + i + 3) * ld] - mean;
      q0 = fma(d0, d0, q0);
      q1 = fma(d1, d1, q1);
      q2 = fma(d2, d2, q2);
      q3 = fma(d3, d3, q3);
    }
    for (; i < L; ++i) {
      const double d = w[(long long)(edge + i) * ld] - mean;
      q0 = fma(d, d, q0);
    }
    const double inv = 1.0 / (sqrt(((q0 + q1) + (q2 + q3)) / (double)L) + eps);
#pragma unroll 8
    for (int i2 = 0; i2 < L; ++i2) outr[i2] = (float)((w[(long long)(edge + i2) * ld] - mean) * inv);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The default since round 2 (ECGMM_PREP_BLOCK=0 selects the one-thread-per-signal kernel above; 0.8 M -> 12.9 M signals/s):
// the same preprocessing, parallel ALONG TIME.  One CTA of 256 threads per signal, the signal in shared
// memory as float64; a filter pass over n samples is
//   A. every thread runs the DF2T recurrence over its own block of T consecutive samples from a ZERO state and keeps
//      the final state f_j (the recurrence is linear: end state = A^T * start state + f_j);
//   B. the block start states follow from a scan of the affine maps s -> M s + f_j, M = A^T (all blocks share M, so
//      the scan needs only the powers M^1..M^32): a Hillis-Steele scan inside each warp (5 steps with M^1, M^2, M^4,
//      M^8, M^16), 8 serial steps across the warps with M^32, one matrix-vector product per thread;
//   C. every thread reruns its block from its true start state and writes the outputs in place.
// Twice the recurrence arithmetic of the serial kernel, spread over 256 threads instead of one.  The state matrix is
// a companion matrix: for narrow-band / high-order designs its powers are badly conditioned and the block result
// drifts away from the serial recurrence (wn = 0.02, order 5: 1e-2 relative, emulated in float64 on the host), so
// the launcher selects this kernel only when max |A^k| stays below 1e3 (the reference's two designs, butter(5, 0.1)
// and butter(5, 0.32): 216 and 2.2; difference to the serial recurrence 1e-9 relative) and keeps the serial kernel
// for everything else.  tests/test_preprocess_cpu.py emulates exactly this schedule in numpy.
constexpr int kBlkThreads = 256;

struct BlockPrepParams {
  IirCoef c;
  double M[kMaxOrder * kMaxOrder];  // A^T, row-major ORDER x ORDER
  int T;                            // samples per thread in a filter pass (odd: conflict-free 64-bit smem strides)
  int chunk;                        // samples per thread in the baseline pass (odd)
};

__device__ __forceinline__ double block_sum_f64(double v, double* red /* >= 9 doubles */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kBlkThreads / 32; ++i) t += red[i];
    red[8] = t;
  }
  __syncthreads();
  return red[8];
}

// y = Mat * v  (Mat row-major ORDER x ORDER in shared memory)
template <int ORDER>
__device__ __forceinline__ void matvec(const double* Mat, const double (&v)[kMaxOrder], double (&y)[kMaxOrder]) {
#pragma unroll
  for (int r = 0; r < ORDER; ++r) {
    double a = 0.0;
#pragma unroll
    for (int q = 0; q < ORDER; ++q) a = fma(Mat[r * ORDER + q], v[q], a);
    y[r] = a;
  }
}

// One zero-phase half: filter buf[0..n) in place, forward (REVERSE = false) or from the end backwards.
template <int ORDER, bool REVERSE>
__device__ __forceinline__ void block_filter_pass(double* buf, int n, int T, const IirCoef& c, const double* P,
                                                  double* Fw, double* wsS) {
  constexpr int OO = ORDER * ORDER;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = (n + T - 1) / T;
  auto at = [&](int m) { return REVERSE ? n - 1 - m : m; };
  const int m0 = tid * T, len = tid < nb ? min(T, n - m0) : 0;
  const double x_first = buf[at(0)];  // read before anybody writes (pass C comes after two barriers)
  // ---- A: zero-state run
  double z[kMaxOrder];
#pragma unroll
  for (int k = 0; k < kMaxOrder; ++k) z[k] = 0.0;
  for (int k = 0; k < len; ++k) (void)iir_step<ORDER>(c, z, buf[at(m0 + k)]);
  // ---- B1: inclusive scan of s -> M s + f inside the warp
  double v[kMaxOrder], u[kMaxOrder], t[kMaxOrder];
#pragma unroll
  for (int k = 0; k < kMaxOrder; ++k) v[k] = k < ORDER ? z[k] : 0.0;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
    for (int k = 0; k < ORDER; ++k) u[k] = __shfl_up_sync(0xffffffffu, v[k], d);
    matvec<ORDER>(P + d * OO, u, t);
    if (lane >= d) {
#pragma unroll
      for (int k = 0; k < ORDER; ++k) v[k] += t[k];
    }
  }
  double lp[kMaxOrder];  // state at the start of this block if the warp had started from zero
#pragma unroll
  for (int k = 0; k < ORDER; ++k) {
    lp[k] = __shfl_up_sync(0xffffffffu, v[k], 1);
    if (lane == 0) lp[k] = 0.0;
  }
  if (lane == 31) {
#pragma unroll
    for (int k = 0; k < ORDER; ++k) Fw[warp * ORDER + k] = v[k];
  }
  __syncthreads();
  // ---- B2: across the warps (serial, 8 steps): start state of every warp
  if (tid == 0) {
    double s[kMaxOrder], sn[kMaxOrder];
#pragma unroll
    for (int k = 0; k < kMaxOrder; ++k) s[k] = k < ORDER ? c.zi[k] * x_first : 0.0;
    for (int w = 0; w < kBlkThreads / 32; ++w) {
#pragma unroll
      for (int k = 0; k < ORDER; ++k) wsS[w * ORDER + k] = s[k];
      matvec<ORDER>(P + 32 * OO, s, sn);
#pragma unroll
      for (int k = 0; k < ORDER; ++k) s[k] = sn[k] + Fw[w * ORDER + k];
    }
  }
  __syncthreads();
  // ---- B3 + C: true start state, rerun, write in place
#pragma unroll
  for (int k = 0; k < kMaxOrder; ++k) u[k] = k < ORDER ? wsS[warp * ORDER + k] : 0.0;
  matvec<ORDER>(P + lane * OO, u, t);
#pragma unroll
  for (int k = 0; k < ORDER; ++k) z[k] = t[k] + lp[k];
  for (int k = 0; k < len; ++k) {
    const int i = at(m0 + k);
    buf[i] = iir_step<ORDER>(c, z, buf[i]);
  }
  __syncthreads();
}

// dynamic shared memory (doubles): xs[L] (raw signal, only if window > 0) | buf[L + 2 edge] | P[33][ORDER^2] |
// Fw[8][ORDER] | wsS[8][ORDER] | red[9]
template <typename TIn, int ORDER>
__global__ void __launch_bounds__(kBlkThreads) signal_preprocess_block_kernel(
    const TIn* __restrict__ x, float* __restrict__ out, int L, int window, const __grid_constant__ BlockPrepParams prm,
    int zscore, double eps) {
  extern __shared__ double sm_d[];
  constexpr int edge = ORDER > 0 ? 3 * (ORDER + 1) : 0;
  constexpr int OO = ORDER * ORDER;
  const int n = L + 2 * edge;
  double* xs = sm_d;
  double* buf = xs + (window > 0 ? L : 0);
  double* P = buf + n;
  double* Fw = P + 33 * OO;
  double* wsS = Fw + 8 * (ORDER > 0 ? ORDER : 1);
  double* red = wsS + 8 * (ORDER > 0 ? ORDER : 1);
  const int tid = threadIdx.x;
  const TIn* xr = x + (size_t)blockIdx.x * L;
  float* outr = out + (size_t)blockIdx.x * L;
  // ---- load (coalesced), float64 from here on
  for (int i = tid; i < L; i += kBlkThreads) {
    const double v = (double)xr[i];
    if (window > 0)
      xs[i] = v;
    else
      buf[edge + i] = v;
  }
  // powers of M = A^T: P[0] = I, P[k] = M P[k-1]  (one thread per matrix entry, 32 dependent steps)
  if constexpr (ORDER > 0) {
    if (tid < OO) P[tid] = (tid / ORDER == tid % ORDER) ? 1.0 : 0.0;
    for (int k = 1; k <= 32; ++k) {
      __syncthreads();
      if (tid < OO) {
        const int r = tid / ORDER, q = tid % ORDER;
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < ORDER; ++j) a = fma(prm.M[r * ORDER + j], P[(k - 1) * OO + j * ORDER + q], a);
        P[k * OO + tid] = a;
      }
    }
  }
  __syncthreads();
  // ---- baseline removal: np.convolve(x, ones(W)/W, 'same')[i] = (1/W) sum x[i-lo .. i+hi], zeros outside
  if (window > 0) {
    const int lo = window / 2, hi = window - lo - 1;
    const double inv = 1.0 / (double)window;
    const int i0 = tid * prm.chunk, i1 = min(L, i0 + prm.chunk);
    if (i0 < L) {
      double S = 0.0;
      for (int t = max(0, i0 - lo); t <= min(L - 1, i0 + hi); ++t) S += xs[t];
      for (int i = i0; i < i1; ++i) {
        buf[edge + i] = xs[i] - S * inv;
        if (i + 1 + hi < L) S += xs[i + 1 + hi];
        if (i - lo >= 0) S -= xs[i - lo];
      }
    }
    __syncthreads();
  }
  if constexpr (ORDER > 0) {
    // ---- odd extension by `edge` samples on both sides
    if (tid >= 1 && tid <= edge) {
      buf[edge - tid] = 2.0 * buf[edge] - buf[edge + tid];
      buf[edge + L - 1 + tid] = 2.0 * buf[edge + L - 1] - buf[edge + L - 1 - tid];
    }
    __syncthreads();
    block_filter_pass<ORDER, false>(buf, n, prm.T, prm.c, P, Fw, wsS);
    block_filter_pass<ORDER, true>(buf, n, prm.T, prm.c, P, Fw, wsS);
  }
  // ---- output (float32), optionally z-scored: (y - mean) / (std_population + eps)
  if (zscore) {
    double s1 = 0.0;
    for (int i = tid; i < L; i += kBlkThreads) s1 += buf[edge + i];
    const double mean = block_sum_f64(s1, red) / (double)L;
    double q = 0.0;
    for (int i = tid; i < L; i += kBlkThreads) {
      const double d = buf[edge + i] - mean;
      q = fma(d, d, q);
    }
    const double inv = 1.0 / (sqrt(block_sum_f64(q, red) / (double)L) + eps);
    for (int i = tid; i < L; i += kBlkThreads) outr[i] = (float)((buf[edge + i] - mean) * inv);
  } else {
    for (int i = tid; i < L; i += kBlkThreads) outr[i] = (float)buf[edge + i];
  }
}

// Host side of the block kernel: M = A^T and the conditioning guard.  Returns false when the serial kernel has to run.
static bool block_plan(int order, const IirCoef& c, int L, int window, BlockPrepParams* prm, size_t* smem_bytes) {
  const int edge = order > 0 ? 3 * (order + 1) : 0;
  const int n = L + 2 * edge;
  int T = (n + kBlkThreads - 1) / kBlkThreads;
  T |= 1;
  int chunk = (L + kBlkThreads - 1) / kBlkThreads;
  chunk |= 1;
  prm->c = c;
  prm->T = T;
  prm->chunk = chunk;
  for (int i = 0; i < kMaxOrder * kMaxOrder; ++i) prm->M[i] = 0.0;
  const int oo = order > 0 ? order : 1;
  *smem_bytes = sizeof(double) * ((size_t)(window > 0 ? L : 0) + n + 33 * (size_t)order * order + 16 * (size_t)oo + 9);
  if (*smem_bytes > 200 * 1024) return false;
  if (order == 0) return true;
  // companion (DF2T) state matrix: z' = A z + B x with A[k][0] = -a[k+1], A[k][k+1] = 1
  long double A[kMaxOrder][kMaxOrder] = {}, Pw[kMaxOrder][kMaxOrder] = {}, Nx[kMaxOrder][kMaxOrder];
  for (int k = 0; k < order; ++k) {
    A[k][0] = -(long double)c.a[k + 1];
    if (k + 1 < order) A[k][k + 1] = 1.0L;
    Pw[k][k] = 1.0L;
  }
  long double worst = 0.0L;
  for (int t = 1; t <= 32 * T; ++t) {  // every power that the scan multiplies with lies on this path
    for (int r = 0; r < order; ++r)
      for (int q = 0; q < order; ++q) {
        long double a = 0.0L;
        for (int j = 0; j < order; ++j) a += A[r][j] * Pw[j][q];
        Nx[r][q] = a;
      }
    for (int r = 0; r < order; ++r)
      for (int q = 0; q < order; ++q) {
        Pw[r][q] = Nx[r][q];
        const long double m = Nx[r][q] < 0 ? -Nx[r][q] : Nx[r][q];
        if (m > worst) worst = m;
      }
    if (t == T)
      for (int r = 0; r < order; ++r)
        for (int q = 0; q < order; ++q) prm->M[r * order + q] = (double)Pw[r][q];
    if (!(worst < 1e3L)) return false;  // badly conditioned design (or unstable): serial kernel
  }
  return true;
}

// scipy.signal.butter(order, wn, 'low') + lfilter_zi, restated (see oracle/preprocess.py for the numpy form).
static int design_butter_lowpass(int order, double wn, double* b, double* a, double* zi) {
  typedef std::complex<double> cd;
  const double pi = 3.14159265358979323846;
  cd p[kMaxOrder];
  for (int k = 0; k < order; ++k) {
    const double m = -order + 1 + 2 * k;
    p[k] = -std::exp(cd(0.0, pi * m / (2.0 * order)));  // analog prototype (buttap)
  }
  const double warped = 4.0 * tan(pi * wn / 2.0);  // 2*fs*tan(pi*wn/fs) with fs = 2
  cd den(1.0, 0.0);
  cd pz[kMaxOrder];
  for (int k = 0; k < order; ++k) {
    p[k] *= warped;                    // lp2lp
    pz[k] = (4.0 + p[k]) / (4.0 - p[k]);  // bilinear, fs2 = 2*fs = 4
    den *= (4.0 - p[k]);
  }
  const double kz = pow(warped, order) * (cd(1.0, 0.0) / den).real();
  // numerator: kz * (1 + z^-1)^order ; denominator: prod (1 - pz_k z^-1)
  cd pa[kMaxOrder + 1];
  double pb[kMaxOrder + 1];
  for (int i = 0; i <= order; ++i) {
    pa[i] = cd(0.0, 0.0);
    pb[i] = 0.0;
  }
  pa[0] = cd(1.0, 0.0);
  pb[0] = 1.0;
  for (int k = 0; k < order; ++k) {
    for (int i = k + 1; i >= 1; --i) {
      pa[i] = pa[i] - pz[k] * pa[i - 1];
      pb[i] = pb[i] + pb[i - 1];
    }
  }
  for (int i = 0; i <= order; ++i) {
    b[i] = kz * pb[i];
    a[i] = pa[i].real();
  }
  // lfilter_zi (a[0] == 1): zi solves zi = A zi + B for the companion matrix of a
  double Bsum = 0.0, col0 = 1.0 + a[1];
  for (int k = 1; k <= order; ++k) Bsum += b[k] - a[k] * b[0];
  for (int k = 2; k <= order; ++k) col0 += a[k];
  zi[0] = Bsum / col0;
  double asum = 1.0, csum = 0.0;
  for (int k = 1; k < order; ++k) {
    asum += a[k];
    csum += b[k] - a[k] * b[0];
    zi[k] = asum * zi[0] - csum;
  }
  return ECGMM_OK;
}

template <typename TIn, int ORDER>
static int launch_block(const void* x, float* y, long long rows, int L, int window, const BlockPrepParams& prm,
                        int zscore, double eps, size_t smem, cudaStream_t st) {
  static bool configured[kMaxDevices] = {};  // the shared-memory limit of a kernel is a per-device attribute
  const int ds = device_slot();
  if (!configured[ds]) {
    ECGMM_CUDA(cudaFuncSetAttribute(signal_preprocess_block_kernel<TIn, ORDER>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured[ds] = true;
  }
  signal_preprocess_block_kernel<TIn, ORDER><<<(unsigned)rows, kBlkThreads, smem, st>>>(
      reinterpret_cast<const TIn*>(x), y, L, window, prm, zscore, eps);
  return ECGMM_OK;
}

}  // namespace ecgmm

using namespace ecgmm;

extern "C" int ecgmm_butter_lowpass(int order, double wn, double* b, double* a, double* zi) {
  ECGMM_CHECK(b && a && zi, ECGMM_ERR_ARG, "butter_lowpass: null pointer");
  ECGMM_CHECK(order >= 1 && order <= kMaxOrder, ECGMM_ERR_SHAPE, "butter_lowpass: order %d not in 1..%d", order,
              kMaxOrder);
  ECGMM_CHECK(wn > 0.0 && wn < 1.0, ECGMM_ERR_ARG, "butter_lowpass: normalised cutoff %g must be in (0, 1)", wn);
  return design_butter_lowpass(order, wn, b, a, zi);
}

extern "C" long long ecgmm_signal_preprocess_workspace(long long rows, int L, int order) {
  if (rows <= 0 || L <= 0 || order < 0 || order > kMaxOrder) return 0;
  const long long ld = (rows + 31) / 32 * 32;
  const int edge = order > 0 ? 3 * (order + 1) : 0;
  return ld * (long long)(L + 2 * edge) * (long long)sizeof(double);
}

extern "C" int ecgmm_signal_preprocess(const void* x, int x_is_f64, float* y, void* workspace,
                                       long long workspace_bytes, long long rows, int L, int window, int order,
                                       double wn, int zscore, double eps, void* stream) {
  ECGMM_CHECK(x && y, ECGMM_ERR_ARG, "signal_preprocess: null pointer");
  ECGMM_CHECK(order >= 0 && order <= kMaxOrder, ECGMM_ERR_SHAPE, "signal_preprocess: order %d not in 0..%d", order,
              kMaxOrder);
  ECGMM_CHECK(window >= 0, ECGMM_ERR_SHAPE, "signal_preprocess: window %d", window);
  ECGMM_CHECK(rows >= 0 && L > 0, ECGMM_ERR_SHAPE, "signal_preprocess: bad extent rows=%lld L=%d", rows, L);
  if (rows == 0) return ECGMM_OK;
  const int edge = order > 0 ? 3 * (order + 1) : 0;
  // np.convolve(mode='same') returns max(L, window) samples and scipy's filtfilt refuses len(x) <= padlen:
  ECGMM_CHECK(window == 0 || L >= window, ECGMM_ERR_SHAPE,
              "signal_preprocess: signal length %d shorter than the moving-average window %d", L, window);
  ECGMM_CHECK(L > edge, ECGMM_ERR_SHAPE, "signal_preprocess: signal length %d must exceed filtfilt's padlen %d", L,
              edge);
  const long long need = ecgmm_signal_preprocess_workspace(rows, L, order);
  ECGMM_CHECK(workspace && workspace_bytes >= need, ECGMM_ERR_ARG,
              "signal_preprocess: workspace of %lld bytes needed, %lld given", need, workspace_bytes);
  ECGMM_CHECK((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, ECGMM_ERR_ALIGN, "signal_preprocess: workspace alignment");
  IirCoef c;
  for (int i = 0; i <= kMaxOrder; ++i) c.b[i] = c.a[i] = 0.0;
  for (int i = 0; i < kMaxOrder; ++i) c.zi[i] = 0.0;
  if (order > 0) {
    ECGMM_CHECK(wn > 0.0 && wn < 1.0, ECGMM_ERR_ARG, "signal_preprocess: normalised cutoff %g must be in (0, 1)", wn);
    design_butter_lowpass(order, wn, c.b, c.a, c.zi);
  }
  cudaStream_t st = (cudaStream_t)stream;
  const char* eblk = getenv("ECGMM_PREP_BLOCK");  // "0": force the one-thread-per-signal kernel
  if (!(eblk && atoi(eblk) == 0) && rows <= 0x7fffffffLL) {  // time-parallel kernel when the design is well conditioned
    BlockPrepParams prm;
    size_t smem = 0;
    if (block_plan(order, c, L, window, &prm, &smem)) {
#define ECGMM_BLK_ORDER(O_)                                                                                     \
  case O_:                                                                                                      \
    rc = x_is_f64 ? launch_block<double, O_>(x, y, rows, L, window, prm, zscore, eps, smem, st)                 \
                  : launch_block<float, O_>(x, y, rows, L, window, prm, zscore, eps, smem, st);                 \
    break
      int rc = ECGMM_OK;
      switch (order) {
        ECGMM_BLK_ORDER(0);
        ECGMM_BLK_ORDER(1);
        ECGMM_BLK_ORDER(2);
        ECGMM_BLK_ORDER(3);
        ECGMM_BLK_ORDER(4);
        ECGMM_BLK_ORDER(5);
        ECGMM_BLK_ORDER(6);
        ECGMM_BLK_ORDER(7);
        default:
          ECGMM_BLK_ORDER(8);
      }
#undef ECGMM_BLK_ORDER
      if (rc) return rc;
      return check_launch("signal_preprocess_block_kernel");
    }
  }
  const long long ld = (rows + 31) / 32 * 32;
  const unsigned grid = (unsigned)(ld / 32);
  double* ws = reinterpret_cast<double*>(workspace);
#define ECGMM_PREP(T_, O_)                                                                                          \
  signal_preprocess_kernel<T_, O_><<<grid, 32, 0, st>>>(reinterpret_cast<const T_*>(x), y, ws, rows, ld, L, window, c, \
                                                         zscore, eps)
#define ECGMM_PREP_ORDER(O_)              \
  case O_:                                \
    if (x_is_f64)                         \
      ECGMM_PREP(double, O_);             \
    else                                  \
      ECGMM_PREP(float, O_);              \
    break
  switch (order) {
    ECGMM_PREP_ORDER(0);
    ECGMM_PREP_ORDER(1);
    ECGMM_PREP_ORDER(2);
    ECGMM_PREP_ORDER(3);
    ECGMM_PREP_ORDER(4);
    ECGMM_PREP_ORDER(5);
    ECGMM_PREP_ORDER(6);
    ECGMM_PREP_ORDER(7);
    default:
      ECGMM_PREP_ORDER(8);
  }
#undef ECGMM_PREP_ORDER
#undef ECGMM_PREP
  return check_launch("signal_preprocess_kernel");
}
