// TEST INFRASTRUCTURE ONLY — CPU oracle for the DM-VIO photometric hot path.
// Nothing under oracle/ is part of the product; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Small fixed-size linear algebra + SE(3) in double, written from scratch (the reference
// uses Eigen + Sophus, neither of which exists in this image).
//   SE3::exp / Adj follow thirdparty/Sophus/sophus/se3.hpp:L131-139 (Adj) and L407-428 (exp),
//   so3.hpp:L343-369 (expAndTheta), restated with plain rotation matrices.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include <cassert>

namespace orc {

template <class T, int R, int C>
struct Mat {
  T d[R * C];
  Mat() { for (int i = 0; i < R * C; i++) d[i] = T(0); }
  T& operator()(int r, int c) { return d[r * C + c]; }
  const T& operator()(int r, int c) const { return d[r * C + c]; }
  T& operator[](int i) { return d[i]; }
  const T& operator[](int i) const { return d[i]; }
  static Mat identity() { Mat m; for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = T(1); return m; }
  Mat<T, C, R> transpose() const { Mat<T, C, R> m; for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) m(c, r) = (*this)(r, c); return m; }
  template <class U> Mat<U, R, C> cast() const { Mat<U, R, C> m; for (int i = 0; i < R * C; i++) m.d[i] = (U)d[i]; return m; }
  Mat& operator+=(const Mat& o) { for (int i = 0; i < R * C; i++) d[i] += o.d[i]; return *this; }
  Mat& operator-=(const Mat& o) { for (int i = 0; i < R * C; i++) d[i] -= o.d[i]; return *this; }
  Mat& operator*=(T s) { for (int i = 0; i < R * C; i++) d[i] *= s; return *this; }
};
template <class T, int R, int C> Mat<T, R, C> operator+(Mat<T, R, C> a, const Mat<T, R, C>& b) { a += b; return a; }
template <class T, int R, int C> Mat<T, R, C> operator-(Mat<T, R, C> a, const Mat<T, R, C>& b) { a -= b; return a; }
template <class T, int R, int C> Mat<T, R, C> operator*(Mat<T, R, C> a, T s) { a *= s; return a; }
template <class T, int R, int K, int C>
Mat<T, R, C> operator*(const Mat<T, R, K>& a, const Mat<T, K, C>& b) {
  Mat<T, R, C> m;
  for (int r = 0; r < R; r++)
    for (int c = 0; c < C; c++) {
      T s = T(0);
      for (int k = 0; k < K; k++) s += a(r, k) * b(k, c);
      m(r, c) = s;
    }
  return m;
}

typedef Mat<double, 3, 3> Mat33;
typedef Mat<double, 3, 1> Vec3;
typedef Mat<double, 6, 1> Vec6;
typedef Mat<double, 6, 6> Mat66;
typedef Mat<double, 8, 8> Mat88;
typedef Mat<double, 8, 1> Vec8;
typedef Mat<double, 10, 1> Vec10;
typedef Mat<float, 3, 3> Mat33f;
typedef Mat<float, 3, 1> Vec3f;
typedef Mat<float, 8, 8> Mat88f;

inline Mat33 hat(const Vec3& w) {
  Mat33 m;
  m(0, 1) = -w[2]; m(0, 2) = w[1];
  m(1, 0) = w[2];  m(1, 2) = -w[0];
  m(2, 0) = -w[1]; m(2, 1) = w[0];
  return m;
}

// Rotation stored as a unit quaternion (w,x,y,z) like Sophus::SO3 so that products renormalise the same way.
struct SO3 {
  double qw = 1, qx = 0, qy = 0, qz = 0;
  Mat33 matrix() const {
    Mat33 R;
    const double w = qw, x = qx, y = qy, z = qz;
    R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - w * z);     R(0, 2) = 2 * (x * z + w * y);
    R(1, 0) = 2 * (x * y + w * z);     R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - w * x);
    R(2, 0) = 2 * (x * z - w * y);     R(2, 1) = 2 * (y * z + w * x);     R(2, 2) = 1 - 2 * (x * x + y * y);
    return R;
  }
  void normalize() { double n = std::sqrt(qw * qw + qx * qx + qy * qy + qz * qz); qw /= n; qx /= n; qy /= n; qz /= n; }
  SO3 operator*(const SO3& o) const {
    SO3 r;
    r.qw = qw * o.qw - qx * o.qx - qy * o.qy - qz * o.qz;
    r.qx = qw * o.qx + qx * o.qw + qy * o.qz - qz * o.qy;
    r.qy = qw * o.qy - qx * o.qz + qy * o.qw + qz * o.qx;
    r.qz = qw * o.qz + qx * o.qy - qy * o.qx + qz * o.qw;
    r.normalize();
    return r;
  }
  SO3 inverse() const { SO3 r; r.qw = qw; r.qx = -qx; r.qy = -qy; r.qz = -qz; return r; }
  // so3.hpp:L343-369
  static SO3 expAndTheta(const Vec3& omega, double* theta) {
    const double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
    *theta = std::sqrt(theta_sq);
    const double half_theta = 0.5 * (*theta);
    double imag_factor, real_factor;
    if (*theta < 1e-10) {
      const double theta_po4 = theta_sq * theta_sq;
      imag_factor = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
      real_factor = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * theta_po4;
    } else {
      imag_factor = std::sin(half_theta) / (*theta);
      real_factor = std::cos(half_theta);
    }
    SO3 r; r.qw = real_factor; r.qx = imag_factor * omega[0]; r.qy = imag_factor * omega[1]; r.qz = imag_factor * omega[2];
    r.normalize();
    return r;
  }
  // so3.hpp log(): rotation vector
  Vec3 log(double* theta_out = nullptr) const {
    const double squared_n = qx * qx + qy * qy + qz * qz;
    const double n = std::sqrt(squared_n);
    const double w = qw;
    double two_atan_nbyw_by_n;
    if (n < 1e-10) {
      const double squared_w = w * w;
      two_atan_nbyw_by_n = 2.0 / w - 2.0 * squared_n / (w * squared_w);
    } else {
      if (std::fabs(w) < 1e-10) {
        two_atan_nbyw_by_n = (w > 0 ? M_PI : -M_PI) / n;
      } else {
        two_atan_nbyw_by_n = 2.0 * std::atan(n / w) / n;
      }
    }
    if (theta_out) *theta_out = two_atan_nbyw_by_n * n;
    Vec3 r; r[0] = two_atan_nbyw_by_n * qx; r[1] = two_atan_nbyw_by_n * qy; r[2] = two_atan_nbyw_by_n * qz;
    return r;
  }
  static SO3 fromMatrix(const Mat33& R) {
    SO3 q;
    double tr = R(0, 0) + R(1, 1) + R(2, 2);
    if (tr > 0) {
      double s = std::sqrt(tr + 1.0) * 2;
      q.qw = 0.25 * s; q.qx = (R(2, 1) - R(1, 2)) / s; q.qy = (R(0, 2) - R(2, 0)) / s; q.qz = (R(1, 0) - R(0, 1)) / s;
    } else if (R(0, 0) > R(1, 1) && R(0, 0) > R(2, 2)) {
      double s = std::sqrt(1.0 + R(0, 0) - R(1, 1) - R(2, 2)) * 2;
      q.qw = (R(2, 1) - R(1, 2)) / s; q.qx = 0.25 * s; q.qy = (R(0, 1) + R(1, 0)) / s; q.qz = (R(0, 2) + R(2, 0)) / s;
    } else if (R(1, 1) > R(2, 2)) {
      double s = std::sqrt(1.0 + R(1, 1) - R(0, 0) - R(2, 2)) * 2;
      q.qw = (R(0, 2) - R(2, 0)) / s; q.qx = (R(0, 1) + R(1, 0)) / s; q.qy = 0.25 * s; q.qz = (R(1, 2) + R(2, 1)) / s;
    } else {
      double s = std::sqrt(1.0 + R(2, 2) - R(0, 0) - R(1, 1)) * 2;
      q.qw = (R(1, 0) - R(0, 1)) / s; q.qx = (R(0, 2) + R(2, 0)) / s; q.qy = (R(1, 2) + R(2, 1)) / s; q.qz = 0.25 * s;
    }
    q.normalize();
    return q;
  }
};

struct SE3 {
  SO3 so3;
  Vec3 t;
  Mat33 rotationMatrix() const { return so3.matrix(); }
  const Vec3& translation() const { return t; }
  SE3 operator*(const SE3& o) const {
    SE3 r; r.so3 = so3 * o.so3;
    Vec3 rt = so3.matrix() * o.t;
    r.t = t + rt;
    return r;
  }
  SE3 inverse() const {
    SE3 r; r.so3 = so3.inverse();
    Vec3 v = r.so3.matrix() * t;
    r.t = v * (-1.0);
    return r;
  }
  // se3.hpp:L131-139 — tangent ordering (translation, rotation)
  Mat66 Adj() const {
    Mat33 R = so3.matrix();
    Mat33 tR = hat(t) * R;
    Mat66 A;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        A(i, j) = R(i, j);
        A(i + 3, j + 3) = R(i, j);
        A(i, j + 3) = tR(i, j);
      }
    return A;
  }
  // se3.hpp:L407-428
  static SE3 exp(const Vec6& a) {
    Vec3 omega; omega[0] = a[3]; omega[1] = a[4]; omega[2] = a[5];
    Vec3 ups; ups[0] = a[0]; ups[1] = a[1]; ups[2] = a[2];
    double theta;
    SO3 so3 = SO3::expAndTheta(omega, &theta);
    Mat33 Omega = hat(omega);
    Mat33 Omega_sq = Omega * Omega;
    Mat33 V;
    if (theta < 1e-10) {
      V = so3.matrix();
    } else {
      double theta_sq = theta * theta;
      V = Mat33::identity() + Omega * ((1.0 - std::cos(theta)) / theta_sq) + Omega_sq * ((theta - std::sin(theta)) / (theta_sq * theta));
    }
    SE3 r; r.so3 = so3; r.t = V * ups;
    return r;
  }
  // se3.hpp log()
  Vec6 log() const {
    double theta;
    Vec3 omega = so3.log(&theta);
    Mat33 Omega = hat(omega);
    Vec3 ups;
    if (std::fabs(theta) < 1e-10) {
      Mat33 V_inv = Mat33::identity() - Omega * 0.5 + (Omega * Omega) * (1.0 / 12.0);
      ups = V_inv * t;
    } else {
      double half_theta = 0.5 * theta;
      Mat33 V_inv = Mat33::identity() - Omega * 0.5 +
                    (Omega * Omega) * ((1.0 - theta * std::cos(half_theta) / (2.0 * std::sin(half_theta))) / (theta * theta));
      ups = V_inv * t;
    }
    Vec6 r; r[0] = ups[0]; r[1] = ups[1]; r[2] = ups[2]; r[3] = omega[0]; r[4] = omega[1]; r[5] = omega[2];
    return r;
  }
  static SE3 fromRt(const double* R9_rowmajor, const double* t3) {
    Mat33 R; for (int i = 0; i < 9; i++) R.d[i] = R9_rowmajor[i];
    SE3 r; r.so3 = SO3::fromMatrix(R); r.t[0] = t3[0]; r.t[1] = t3[1]; r.t[2] = t3[2];
    return r;
  }
};

// Dense dynamic matrix (row-major, double) for the reduced system.
struct MatX {
  int rows = 0, cols = 0;
  std::vector<double> d;
  MatX() {}
  MatX(int r, int c) : rows(r), cols(c), d((size_t)r * c, 0.0) {}
  double& operator()(int r, int c) { return d[(size_t)r * cols + c]; }
  const double& operator()(int r, int c) const { return d[(size_t)r * cols + c]; }
  void setZero() { std::fill(d.begin(), d.end(), 0.0); }
};
typedef std::vector<double> VecX;

// Plain LDL^T (no pivoting) solve of a symmetric system; Eigen's ldlt() pivots, which only changes rounding.
inline bool ldlt_solve(const MatX& A, const VecX& b, VecX& x) {
  const int n = A.rows;
  std::vector<double> L((size_t)n * n, 0.0), D(n, 0.0);
  for (int j = 0; j < n; j++) {
    double dj = A(j, j);
    for (int k = 0; k < j; k++) dj -= L[(size_t)j * n + k] * L[(size_t)j * n + k] * D[k];
    D[j] = dj;
    L[(size_t)j * n + j] = 1.0;
    for (int i = j + 1; i < n; i++) {
      double s = A(i, j);
      for (int k = 0; k < j; k++) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k] * D[k];
      L[(size_t)i * n + j] = (dj != 0.0) ? s / dj : 0.0;
    }
  }
  x.assign(n, 0.0);
  std::vector<double> y(n);
  for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[(size_t)i * n + k] * y[k]; y[i] = s; }
  for (int i = 0; i < n; i++) y[i] = (D[i] != 0.0) ? y[i] / D[i] : 0.0;
  for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * x[k]; x[i] = s; }
  return true;
}


// Fixed-size 3x3 inverse the way Eigen evaluates it (cofactors, determinant from the first column, ONE reciprocal): the reference
// calls K.inverse() on float 3x3 pinhole matrices (HessianBlocks.cpp:L217, CoarseTracker.cpp:L128, globalCalib.cpp:L82) and the
// rounding of that inverse is visible downstream; pinned against the compiled reference by tests/test_ref_pin.py.
template <class T>
inline Mat<T, 3, 3> inverse3_cofactor(const Mat<T, 3, 3>& K) {
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return K(i1, j1) * K(i2, j2) - K(i1, j2) * K(i2, j1);
  };
  const T c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  const T det = (c0 * K(0, 0) + c1 * K(1, 0)) + c2 * K(2, 0);
  const T invdet = T(1) / det;
  Mat<T, 3, 3> R;
  R(0, 0) = c0 * invdet; R(0, 1) = c1 * invdet; R(0, 2) = c2 * invdet;
  R(1, 0) = cof(0, 1) * invdet; R(1, 1) = cof(1, 1) * invdet; R(1, 2) = cof(2, 1) * invdet;
  R(2, 0) = cof(0, 2) * invdet; R(2, 1) = cof(1, 2) * invdet; R(2, 2) = cof(2, 2) * invdet;
  return R;
}

}  // namespace orc
