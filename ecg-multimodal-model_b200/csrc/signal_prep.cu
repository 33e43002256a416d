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
  const long long ld = (rows + 31) / 32 * 32;
  const unsigned grid = (unsigned)(ld / 32);
  double* ws = reinterpret_cast<double*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
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
