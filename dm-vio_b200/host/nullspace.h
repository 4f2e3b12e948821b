// Gauge handling of the reduced BA solve, host fp64 like the reference:
//   windowNullspaces  = FrameHessian::setStateZero's numeric nullspaces (FullSystem/HessianBlocks.cpp:L74-126) assembled as
//                       FullSystem::getNullspaces does (FullSystem/FullSystemOptimize.cpp:L704-760): 6 pose + 1 scale directions;
//   orthogonalizeX    = EnergyFunctional::orthogonalize(&x, 0) (OptimizationBackend/EnergyFunctional.cpp:L784-838): x -= N N^+ x with the
//                       singular values below setting_solverModeDelta * max dropped (applied by solveSystemF from iteration 2 on in the
//                       default solver mode SOLVER_ORTHOGONALIZE_X_LATER, EnergyFunctional.cpp:L980-984).
// Header-only and free of device code so that the C glue can expose it to CPU tests.
#pragma once
#include "se3.h"
#include <algorithm>
#include <cmath>
#include <vector>

namespace dmvio_b200 {

// out: 7 vectors of size 8 nf + 4 (frame block f at 4 + 8 f: trans3, rot3, a, b); scale_trans / scale_rot = SCALE_XI_TRANS / SCALE_XI_ROT
inline std::vector<std::vector<double>> windowNullspaces(const std::vector<SE3>& evalPT, double scale_trans, double scale_rot) {
  const int n = (int)evalPT.size(), N = 8 * n + 4;
  std::vector<std::vector<double>> ns(7, std::vector<double>(N, 0.0));
  for (int f = 0; f < n; f++) {
    const SE3& T = evalPT[f];
    const SE3 Ti = T.inverse();
    double lp[6], lm[6];
    for (int i = 0; i < 6; i++) {
      double eps[6] = {0, 0, 0, 0, 0, 0}, epsm[6] = {0, 0, 0, 0, 0, 0};
      eps[i] = 1e-3; epsm[i] = -1e-3;
      ((T * SE3::exp(eps)) * Ti).log(lp);
      ((T * SE3::exp(epsm)) * Ti).log(lm);
      for (int k = 0; k < 6; k++) ns[i][4 + 8 * f + k] = (lp[k] - lm[k]) / (2e-3) * (k < 3 ? 1.0 / scale_trans : 1.0 / scale_rot);
    }
    SE3 P = T, M = T;
    for (int k = 0; k < 3; k++) { P.t[k] *= 1.00001; M.t[k] /= 1.00001; }
    (P * Ti).log(lp);
    (M * Ti).log(lm);
    for (int k = 0; k < 6; k++) ns[6][4 + 8 * f + k] = (lp[k] - lm[k]) / (2e-3) * (k < 3 ? 1.0 / scale_trans : 1.0 / scale_rot);
  }
  return ns;
}

namespace detail {
// eigen-decomposition of a small symmetric matrix (k x k row-major) by cyclic Jacobi rotations: A = V diag(w) V^T
inline void jacobiEigenSym(std::vector<double> A, int k, std::vector<double>& w, std::vector<double>& V) {
  V.assign((size_t)k * k, 0.0);
  for (int i = 0; i < k; i++) V[(size_t)i * k + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int p = 0; p < k; p++) for (int q = p + 1; q < k; q++) off += A[(size_t)p * k + q] * A[(size_t)p * k + q];
    if (off < 1e-300) break;
    for (int p = 0; p < k; p++)
      for (int q = p + 1; q < k; q++) {
        const double apq = A[(size_t)p * k + q];
        if (apq == 0.0) continue;
        const double theta = (A[(size_t)q * k + q] - A[(size_t)p * k + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < k; r++) {  // columns p, q
          const double a = A[(size_t)r * k + p], b = A[(size_t)r * k + q];
          A[(size_t)r * k + p] = c * a - s * b; A[(size_t)r * k + q] = s * a + c * b;
        }
        for (int r = 0; r < k; r++) {  // rows p, q
          const double a = A[(size_t)p * k + r], b = A[(size_t)q * k + r];
          A[(size_t)p * k + r] = c * a - s * b; A[(size_t)q * k + r] = s * a + c * b;
        }
        for (int r = 0; r < k; r++) {
          const double a = V[(size_t)r * k + p], b = V[(size_t)r * k + q];
          V[(size_t)r * k + p] = c * a - s * b; V[(size_t)r * k + q] = s * a + c * b;
        }
      }
  }
  w.resize(k);
  for (int i = 0; i < k; i++) w[i] = A[(size_t)i * k + i];
}
}  // namespace detail

// Orthonormal basis of span(ns) restricted to the singular directions above solverModeDelta * max: the columns u_a = N v_a / sigma_a of the
// normalised nullspace matrix (sigma^2, v from the 7 x 7 Gram matrix).  It depends on the frames' evaluation points only, which do not
// move during FullSystem::optimize: built once per window state and reused by every solve (gaugeProject).
inline std::vector<std::vector<double>> gaugeBasis(std::vector<std::vector<double>> ns, double solverModeDelta) {
  const int k = (int)ns.size(), N = k ? (int)ns[0].size() : 0;
  for (std::vector<double>& v : ns) {
    double nrm = 0;
    for (double e : v) nrm += e * e;
    nrm = std::sqrt(nrm);
    for (double& e : v) e /= nrm;
  }
  std::vector<double> G((size_t)k * k), w, V;
  for (int a = 0; a < k; a++)
    for (int b = 0; b < k; b++) { double d = 0; for (int i = 0; i < N; i++) d += ns[a][i] * ns[b][i]; G[(size_t)a * k + b] = d; }
  detail::jacobiEigenSym(G, k, w, V);
  double maxSv = 0;
  for (int a = 0; a < k; a++) maxSv = std::max(maxSv, std::sqrt(std::max(0.0, w[a])));
  std::vector<std::vector<double>> U;
  for (int a = 0; a < k; a++) {
    const double sv = std::sqrt(std::max(0.0, w[a]));
    if (!(sv > solverModeDelta * maxSv)) continue;
    std::vector<double> u(N);
    for (int i = 0; i < N; i++) { double d = 0; for (int b = 0; b < k; b++) d += ns[b][i] * V[(size_t)b * k + a]; u[i] = d / sv; }
    U.push_back(std::move(u));
  }
  return U;
}
inline void gaugeProject(std::vector<double>& x, const std::vector<std::vector<double>>& U) {
  const int N = (int)x.size();
  std::vector<double> proj(N, 0.0);
  for (const std::vector<double>& u : U) {
    double ux = 0;
    for (int i = 0; i < N; i++) ux += u[i] * x[i];
    for (int i = 0; i < N; i++) proj[i] += u[i] * ux;
  }
  for (int i = 0; i < N; i++) x[i] -= proj[i];
}
// x (size 8 nf + 4) is projected onto the orthogonal complement of span(ns); the singular values of the normalised nullspace matrix are
// the square roots of the eigenvalues of its 7 x 7 Gram matrix
inline void orthogonalizeX(std::vector<double>& x, std::vector<std::vector<double>> ns, double solverModeDelta) {
  gaugeProject(x, gaugeBasis(std::move(ns), solverModeDelta));
}

}  // namespace dmvio_b200
