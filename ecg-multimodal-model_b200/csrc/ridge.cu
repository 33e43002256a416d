// Local-surrogate regression operator for the perturbation explainers: the weighted ridge
// fit that LIME runs per explained instance (lime_fusion_modal_balance.py:158-160 -> lime's default model_regressor,
// sklearn Ridge(alpha=1, fit_intercept=True) with the kernel weights as sample_weight) and that KernelSHAP runs with
// Shapley-kernel weights, on the BINARY keep-masks of ecgmm_perturb_build.
//
//     minimise over (w, b):   sum_v pi_v (f_v - b - z_v . w)^2 + alpha |w|^2
//     zbar = sum pi z / sum pi,  Zc = Z - zbar,  G = Zc^T diag(pi) Zc + alpha I
//     w = G^-1 Zc^T diag(pi) f,     b = (pi . f) / sum pi - zbar . w            (Zc^T pi = 0: f needs no centring)
//
// The fit is LINEAR in the responses f, and its operator R [(D+1) x V] depends only on the sampling plan (masks,
// weights, alpha) -- not on the model or the data.  It is therefore designed once per plan on the host, in float64, like
// the Butterworth taps of ecgmm_butter_lowpass (no device involved); the per-sample work -- V model evaluations
// (ecgmm_perturb_build + tensor-core GEMM + ecgmm_head_tail) and coefficients = f R^T (ecgmm_sgemm) -- runs on the GPU.
#include "common.h"

#include <math.h>

#include <vector>

using namespace ecgmm;

extern "C" int ecgmm_ridge_operator(const uint8_t* masks, const double* weights, int V, int D, double alpha, float* R) {
  ECGMM_CHECK(masks && weights && R, ECGMM_ERR_ARG, "ridge_operator: null pointer");
  ECGMM_CHECK(V >= 1 && D >= 1 && D <= 8192, ECGMM_ERR_SHAPE, "ridge_operator: bad extents V=%d D=%d", V, D);
  ECGMM_CHECK(alpha >= 0.0, ECGMM_ERR_ARG, "ridge_operator: alpha %g < 0", alpha);
  double sw = 0.0;
  for (int v = 0; v < V; ++v) {
    ECGMM_CHECK(weights[v] >= 0.0 && isfinite(weights[v]), ECGMM_ERR_ARG, "ridge_operator: weight %d is %g", v,
                weights[v]);
    sw += weights[v];
  }
  ECGMM_CHECK(sw > 0.0, ECGMM_ERR_ARG, "ridge_operator: all weights are zero");
  const size_t Ds = (size_t)D, Vs = (size_t)V;
  std::vector<double> zbar(Ds, 0.0), G(Ds * Ds, 0.0), X(Ds * Vs, 0.0);
  // zbar and the raw second moment sum pi z z^T (z binary: only the kept positions of a row contribute)
  std::vector<int> kept;
  kept.reserve(Ds);
  for (int v = 0; v < V; ++v) {
    const double pi = weights[v];
    if (pi == 0.0) continue;
    kept.clear();
    const uint8_t* z = masks + (size_t)v * Ds;
    for (int d = 0; d < D; ++d)
      if (z[d]) kept.push_back(d);
    for (size_t a = 0; a < kept.size(); ++a) {
      zbar[kept[a]] += pi;
      double* row = &G[(size_t)kept[a] * Ds];
      for (size_t b = 0; b <= a; ++b) row[kept[b]] += pi;  // lower triangle (kept[] ascends)
    }
  }
  for (int d = 0; d < D; ++d) zbar[d] /= sw;
  // G = sum pi z z^T - sw zbar zbar^T + alpha I   (lower triangle)
  for (int i = 0; i < D; ++i) {
    for (int j = 0; j <= i; ++j) G[(size_t)i * Ds + j] -= sw * zbar[i] * zbar[j];
    G[(size_t)i * Ds + i] += alpha;
  }
  // Cholesky G = L L^T in place (lower triangle)
  for (int j = 0; j < D; ++j) {
    double* Lj = &G[(size_t)j * Ds];
    double djj = Lj[j];
    for (int k = 0; k < j; ++k) djj -= Lj[k] * Lj[k];
    ECGMM_CHECK(djj > 0.0, ECGMM_ERR_ARG,
                "ridge_operator: normal matrix is not positive definite at column %d (alpha = 0 with constant or "
                "collinear mask columns?)", j);
    const double ljj = sqrt(djj);
    Lj[j] = ljj;
    for (int i = j + 1; i < D; ++i) {
      double* Li = &G[(size_t)i * Ds];
      double s = Li[j];
      for (int k = 0; k < j; ++k) s -= Li[k] * Lj[k];
      Li[j] = s / ljj;
    }
  }
  // right-hand sides B = Zc^T diag(pi): X[d][v] = pi_v (z_vd - zbar_d); then L Y = B, L^T X = Y for all V columns at once
  for (int d = 0; d < D; ++d) {
    double* x = &X[(size_t)d * Vs];
    for (int v = 0; v < V; ++v) x[v] = weights[v] * ((masks[(size_t)v * Ds + d] ? 1.0 : 0.0) - zbar[d]);
  }
  for (int i = 0; i < D; ++i) {
    double* xi = &X[(size_t)i * Vs];
    const double* Li = &G[(size_t)i * Ds];
    for (int k = 0; k < i; ++k) {
      const double l = Li[k];
      if (l == 0.0) continue;
      const double* xk = &X[(size_t)k * Vs];
      for (int v = 0; v < V; ++v) xi[v] -= l * xk[v];
    }
    const double inv = 1.0 / Li[i];
    for (int v = 0; v < V; ++v) xi[v] *= inv;
  }
  for (int i = D - 1; i >= 0; --i) {
    double* xi = &X[(size_t)i * Vs];
    for (int k = i + 1; k < D; ++k) {
      const double l = G[(size_t)k * Ds + i];
      if (l == 0.0) continue;
      const double* xk = &X[(size_t)k * Vs];
      for (int v = 0; v < V; ++v) xi[v] -= l * xk[v];
    }
    const double inv = 1.0 / G[(size_t)i * Ds + i];
    for (int v = 0; v < V; ++v) xi[v] *= inv;
  }
  // rows 0..D-1: the coefficients' operator; row D: the intercept's
  for (int d = 0; d < D; ++d)
    for (int v = 0; v < V; ++v) R[(size_t)d * Vs + v] = (float)X[(size_t)d * Vs + v];
  for (int v = 0; v < V; ++v) {
    double s = weights[v] / sw;
    for (int d = 0; d < D; ++d) s -= zbar[d] * X[(size_t)d * Vs + v];
    R[Ds * Vs + v] = (float)s;
  }
  return ECGMM_OK;
}
