// TEST INFRASTRUCTURE ONLY — minimal stand-in for the part of Sophus (SO3d / SE3d / Sim3d) the reference's hot-path
// translation units touch, on top of oracle/shim/Eigen.  The reference vendors Sophus under thirdparty/, but that copy needs
// the real Eigen.  Conventions follow Sophus: tangent = (upsilon, omega), exp/log with the usual closed forms, Adj = [R, t^ R; 0, R].
#pragma once
#include "Eigen/Core"

namespace Sophus {
typedef Eigen::Matrix<double, 3, 3> M3;
typedef Eigen::Matrix<double, 3, 1> V3;
typedef Eigen::Matrix<double, 6, 1> V6;
typedef Eigen::Matrix<double, 6, 6> M6;

inline M3 hat3(const V3& w) { M3 m; m << 0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0; return m; }

class SO3d {
 public:
  M3 R;
  SO3d() { R.setIdentity(); }
  explicit SO3d(const M3& r) : R(r) {}
  static SO3d exp(const V3& w) {
    const double th2 = w.squaredNorm(), th = std::sqrt(th2);
    const M3 W = hat3(w);
    double A, B;
    if (th < 1e-10) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; } else { A = std::sin(th) / th; B = (1.0 - std::cos(th)) / th2; }
    SO3d r;
    r.R = M3::Identity() + W * A + W * W * B;
    return r;
  }
  V3 log() const {
    const double tr = R.trace();
    double c = 0.5 * (tr - 1.0);
    c = std::min(1.0, std::max(-1.0, c));
    const double th = std::acos(c);
    V3 v(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
    if (th < 1e-10) return v * (0.5 + th * th / 12.0);
    return v * (th / (2.0 * std::sin(th)));
  }
  const M3& matrix() const { return R; }
  SO3d inverse() const { return SO3d(M3(R.transpose())); }
  SO3d operator*(const SO3d& o) const { return SO3d(M3(R * o.R)); }
  V3 operator*(const V3& p) const { return V3(R * p); }
};

class SE3d {
 public:
  M3 R; V3 t;
  SE3d() { R.setIdentity(); }
  SE3d(const M3& r, const V3& tt) : R(r), t(tt) {}
  SE3d(const SO3d& r, const V3& tt) : R(r.R), t(tt) {}
  static SE3d exp(const V6& xi) {
    const V3 ups(xi[0], xi[1], xi[2]), om(xi[3], xi[4], xi[5]);
    const double th2 = om.squaredNorm(), th = std::sqrt(th2);
    const M3 W = hat3(om), W2 = M3(W * W);
    double A, B, Cc;
    if (th < 1e-10) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; Cc = 1.0 / 6.0 - th2 / 120.0; }
    else { A = std::sin(th) / th; B = (1.0 - std::cos(th)) / th2; Cc = (th - std::sin(th)) / (th2 * th); }
    SE3d T;
    T.R = M3::Identity() + W * A + W2 * B;
    const M3 V = M3::Identity() + W * B + W2 * Cc;
    T.t = V * ups;
    return T;
  }
  V6 log() const {
    const V3 om = SO3d(R).log();
    const double th2 = om.squaredNorm(), th = std::sqrt(th2);
    const M3 W = hat3(om), W2 = M3(W * W);
    double k;
    if (th < 1e-10) k = 1.0 / 12.0 + th2 / 720.0;
    else k = (1.0 - th * std::cos(0.5 * th) / (2.0 * std::sin(0.5 * th))) / th2;
    const M3 Vinv = M3::Identity() - W * 0.5 + W2 * k;
    const V3 ups = Vinv * t;
    V6 xi; xi << ups[0], ups[1], ups[2], om[0], om[1], om[2];
    return xi;
  }
  SE3d inverse() const { const M3 Rt = R.transpose(); return SE3d(Rt, V3(-(Rt * t))); }
  SE3d operator*(const SE3d& o) const { return SE3d(M3(R * o.R), V3(R * o.t + t)); }
  V3 operator*(const V3& p) const { return V3(R * p + t); }
  SE3d& operator*=(const SE3d& o) { *this = (*this) * o; return *this; }
  const M3& rotationMatrix() const { return R; }
  const V3& translation() const { return t; }
  V3& translation() { return t; }
  SO3d so3() const { return SO3d(R); }
  void setRotationMatrix(const M3& r) { R = r; }
  Eigen::Matrix<double, 3, 4> matrix3x4() const { Eigen::Matrix<double, 3, 4> m; m.block<3, 3>(0, 0) = R; m.col(3) = t; return m; }
  Eigen::Matrix<double, 4, 4> matrix() const { Eigen::Matrix<double, 4, 4> m; m.setIdentity(); m.block<3, 3>(0, 0) = R; m.block<3, 1>(0, 3) = t; return m; }
  M6 Adj() const {
    M6 a; a.setZero();
    a.block<3, 3>(0, 0) = R;
    a.block<3, 3>(3, 3) = R;
    a.block<3, 3>(0, 3) = hat3(t) * R;
    return a;
  }
};

class Sim3d {  // only named by typedefs on the hot path
 public:
  SE3d se3; double s = 1.0;
  Sim3d() {}
};
typedef SE3d SE3;
typedef SO3d SO3;
typedef Sim3d Sim3;
}  // namespace Sophus
