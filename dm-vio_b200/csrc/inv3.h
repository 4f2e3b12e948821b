// 3x3 float inverse with the rounding of the reference: Eigen evaluates a fixed-size 3x3 inverse by cofactors, determinant from
// the first column and ONE reciprocal (every cofactor is multiplied by it).  The reference calls K.inverse() on float pinhole
// matrices (FullSystem/HessianBlocks.cpp:L217, FullSystem/CoarseTracker.cpp:L128); the rounding is visible in K*R*K^-1 and R*K^-1,
// so the host side reproduces it exactly (pinned against the compiled reference by tests/test_ref_pin.py).  Row-major 3x3.
#pragma once
namespace dmv {
inline void inv3_cofactor(const float K[9], float Ki[9]) {
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return K[i1 * 3 + j1] * K[i2 * 3 + j2] - K[i1 * 3 + j2] * K[i2 * 3 + j1];
  };
  const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  const float det = (c0 * K[0] + c1 * K[3]) + c2 * K[6];
  const float invdet = 1.0f / det;
  Ki[0] = c0 * invdet; Ki[1] = c1 * invdet; Ki[2] = c2 * invdet;
  Ki[3] = cof(0, 1) * invdet; Ki[4] = cof(1, 1) * invdet; Ki[5] = cof(2, 1) * invdet;
  Ki[6] = cof(0, 2) * invdet; Ki[7] = cof(1, 2) * invdet; Ki[8] = cof(2, 2) * invdet;
}
}  // namespace dmv
