// Minimal SE(3) / small dense algebra for the host adapters (the reference uses Sophus + Eigen, absent here).
// Conventions follow Sophus: tangent = (translation, rotation), exp(xi) * T is a LEFT increment
// (reference thirdparty/Sophus/sophus/se3.hpp:L131-139 Adj, L407-428 exp).  Rotation matrices, row-major, double.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace dmvio_b200 {

struct SE3 {
  double R[9];
  double t[3];
  SE3() { setIdentity(); }
  void setIdentity() {
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    t[0] = t[1] = t[2] = 0.0;
  }
  SE3 operator*(const SE3& o) const {
    SE3 r;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) r.R[i * 3 + j] = R[i * 3] * o.R[j] + R[i * 3 + 1] * o.R[3 + j] + R[i * 3 + 2] * o.R[6 + j];
      r.t[i] = R[i * 3] * o.t[0] + R[i * 3 + 1] * o.t[1] + R[i * 3 + 2] * o.t[2] + t[i];
    }
    return r;
  }
  SE3 inverse() const {
    SE3 r;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r.R[i * 3 + j] = R[j * 3 + i];
    for (int i = 0; i < 3; i++) r.t[i] = -(r.R[i * 3] * t[0] + r.R[i * 3 + 1] * t[1] + r.R[i * 3 + 2] * t[2]);
    return r;
  }
  // 6x6 adjoint, row-major: [R, hat(t) R; 0, R]
  void Adj(double A[36]) const {
    const double h[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
    std::memset(A, 0, sizeof(double) * 36);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        A[i * 6 + j] = R[i * 3 + j];
        A[(i + 3) * 6 + j + 3] = R[i * 3 + j];
        A[i * 6 + j + 3] = h[i * 3] * R[j] + h[i * 3 + 1] * R[3 + j] + h[i * 3 + 2] * R[6 + j];
      }
  }
  // Sophus SE3::log (se3.hpp): xi = (V^-1 t, omega), omega from the rotation matrix
  void log(double xi[6]) const {
    const double tr = R[0] + R[4] + R[8];
    const double c = std::min(1.0, std::max(-1.0, 0.5 * (tr - 1.0)));
    const double th = std::acos(c);
    const double v[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    const double f = (th < 1e-10) ? (0.5 + th * th / 12.0) : th / (2.0 * std::sin(th));
    const double om[3] = {v[0] * f, v[1] * f, v[2] * f};
    const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2], tho = std::sqrt(th2);
    const double W[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    double W2[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) W2[i * 3 + j] = W[i * 3] * W[j] + W[i * 3 + 1] * W[3 + j] + W[i * 3 + 2] * W[6 + j];
    const double k = (tho < 1e-10) ? (1.0 / 12.0 + th2 / 720.0) : (1.0 - tho * std::cos(0.5 * tho) / (2.0 * std::sin(0.5 * tho))) / th2;
    for (int i = 0; i < 3; i++) {
      double s = 0;
      for (int j = 0; j < 3; j++) s += (((i == j) ? 1.0 : 0.0) - 0.5 * W[i * 3 + j] + k * W2[i * 3 + j]) * t[j];
      xi[i] = s;
      xi[3 + i] = om[i];
    }
  }
  static SE3 exp(const double xi[6]) {
    const double wx = xi[3], wy = xi[4], wz = xi[5];
    const double th2 = wx * wx + wy * wy + wz * wz, th = std::sqrt(th2);
    const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double W2[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) W2[i * 3 + j] = W[i * 3] * W[j] + W[i * 3 + 1] * W[3 + j] + W[i * 3 + 2] * W[6 + j];
    double a, b, c;  // R = I + a W + b W^2 ; V = I + b W + c W^2
    if (th < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; c = 1.0 / 6.0 - th2 / 120.0; }
    else { a = std::sin(th) / th; b = (1.0 - std::cos(th)) / th2; c = (th - std::sin(th)) / (th2 * th); }
    SE3 r;
    double V[9];
    for (int i = 0; i < 9; i++) {
      const double I = (i % 4 == 0) ? 1.0 : 0.0;
      r.R[i] = I + a * W[i] + b * W2[i];
      V[i] = I + b * W[i] + c * W2[i];
    }
    for (int i = 0; i < 3; i++) r.t[i] = V[i * 3] * xi[0] + V[i * 3 + 1] * xi[1] + V[i * 3 + 2] * xi[2];
    return r;
  }
};

// Plain LDL^T solve of a symmetric positive (semi-)definite system; n x n row-major.  (Eigen's ldlt() pivots; rounding only.)
// dot product with four independent partial sums: the reduced systems are small (N <= 68), so the solve is bound by the latency of the
// dependent adds of a plain loop, not by flops
inline double dot4(const double* a, const double* b, int n) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int k = 0;
  for (; k + 3 < n; k += 4) { s0 += a[k] * b[k]; s1 += a[k + 1] * b[k + 1]; s2 += a[k + 2] * b[k + 2]; s3 += a[k + 3] * b[k + 3]; }
  for (; k < n; k++) s0 += a[k] * b[k];
  return (s0 + s1) + (s2 + s3);
}

// x = A^-1 b by LDL^T without pivoting (A symmetric positive definite after the Jacobi scaling of the caller), row-major
inline void ldlt_solve(int n, const double* A, const double* b, double* x) {
  std::vector<double> L((size_t)n * n, 0.0), D(n, 0.0), y(n), w(n);
  for (int j = 0; j < n; j++) {
    double* Lj = &L[(size_t)j * n];
    for (int k = 0; k < j; k++) w[k] = Lj[k] * D[k];
    const double dj = A[(size_t)j * n + j] - dot4(Lj, w.data(), j);
    D[j] = dj;
    Lj[j] = 1.0;
    const double inv = dj != 0.0 ? 1.0 / dj : 0.0;
    for (int i = j + 1; i < n; i++) L[(size_t)i * n + j] = (A[(size_t)i * n + j] - dot4(&L[(size_t)i * n], w.data(), j)) * inv;
  }
  for (int i = 0; i < n; i++) y[i] = b[i] - dot4(&L[(size_t)i * n], y.data(), i);
  for (int i = 0; i < n; i++) y[i] = D[i] != 0.0 ? y[i] / D[i] : 0.0;
  for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * x[k]; x[i] = s; }
}

}  // namespace dmvio_b200
