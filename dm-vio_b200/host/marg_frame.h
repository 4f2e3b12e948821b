// EnergyFunctional::marginalizeFrame, the visual branch (OptimizationBackend/EnergyFunctional.cpp:L569-631): Schur complement of the
// marginalisation prior HM/bM with respect to one keyframe's 8 variables.  Host fp64 like the reference — (8 nf + 4)^2 doubles, once per
// marginalised keyframe; no device work.  Header-only so that the C glue can expose it to CPU tests without a GPU handle.
#pragma once
#include <cmath>
#include <utility>
#include <vector>

namespace dmvio_b200 {

namespace detail {
// 8x8 inverse, LU with partial pivoting (Eigen's fixed-size inverse() above 4x4)
inline void inverse8(const double A[8][8], double Ainv[8][8]) {
  double M[8][16];
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) { M[i][j] = A[i][j]; M[i][8 + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < 8; c++) {
    int piv = c;
    for (int r = c + 1; r < 8; r++) if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
    if (piv != c) for (int j = 0; j < 16; j++) std::swap(M[c][j], M[piv][j]);
    for (int r = c + 1; r < 8; r++) {
      const double f = M[r][c] / M[c][c];
      if (f != 0.0) for (int j = c; j < 16; j++) M[r][j] -= f * M[c][j];
    }
  }
  for (int c = 7; c >= 0; c--)
    for (int j = 8; j < 16; j++) {
      double v = M[c][j];
      for (int k = c + 1; k < 8; k++) v -= M[c][k] * M[k][j];
      M[c][j] = v / M[c][c];
    }
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) Ainv[i][j] = M[i][8 + j];
}
}  // namespace detail

// HM: odim x odim row-major, bM: odim, odim = 8 nFrames + 4; on return both hold the ndim = odim - 8 system.
// prior / delta_prior: EFFrame::prior, EFFrame::delta_prior of the frame (EnergyFunctionalStructs.cpp:L52-64).
inline void marginalizeFrameHM(std::vector<double>& HM, std::vector<double>& bM, int nFrames, int idx, const double prior[8], const double delta_prior[8]) {
  const int CP = 4, odim = nFrames * 8 + CP, ndim = odim - 8, io = idx * 8 + CP;
  // L572-592: move the frame's rows / columns to the end
  std::vector<int> perm;
  perm.reserve(odim);
  for (int i = 0; i < odim; i++) if (i < io || i >= io + 8) perm.push_back(i);
  for (int k = 0; k < 8; k++) perm.push_back(io + k);
  std::vector<double> H((size_t)odim * odim), b(odim);
  for (int i = 0; i < odim; i++) {
    b[i] = bM[perm[i]];
    for (int j = 0; j < odim; j++) H[(size_t)i * odim + j] = HM[(size_t)perm[i] * odim + perm[j]];
  }
  // L595-596
  for (int k = 0; k < 8; k++) { H[(size_t)(ndim + k) * odim + ndim + k] += prior[k]; b[ndim + k] += prior[k] * delta_prior[k]; }
  // L603-612: scale by 1 / sqrt(|diag| + 10)
  std::vector<double> S(odim), SI(odim);
  for (int i = 0; i < odim; i++) { S[i] = std::sqrt(std::fabs(H[(size_t)i * odim + i]) + 10.0); SI[i] = 1.0 / S[i]; }
  for (int i = 0; i < odim; i++) {
    for (int j = 0; j < odim; j++) H[(size_t)i * odim + j] = (SI[i] * H[(size_t)i * odim + j]) * SI[j];
    b[i] = SI[i] * b[i];
  }
  // L615-618
  double blk[8][8], hpi[8][8];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) blk[i][j] = H[(size_t)(ndim + i) * odim + ndim + j];
  detail::inverse8(blk, hpi);
  // L621-623: bli = bottomLeft^T * hpi; top -= bli * bottomLeft; b_top -= bli * b_bottom
  std::vector<double> bli((size_t)ndim * 8);
  for (int i = 0; i < ndim; i++)
    for (int k = 0; k < 8; k++) {
      double v = 0;
      for (int m = 0; m < 8; m++) v += H[(size_t)(ndim + m) * odim + i] * hpi[m][k];
      bli[(size_t)i * 8 + k] = v;
    }
  std::vector<double> Hn((size_t)ndim * ndim), bn(ndim);
  for (int i = 0; i < ndim; i++) {
    for (int j = 0; j < ndim; j++) {
      double v = 0;
      for (int k = 0; k < 8; k++) v += bli[(size_t)i * 8 + k] * H[(size_t)(ndim + k) * odim + j];
      Hn[(size_t)i * ndim + j] = H[(size_t)i * odim + j] - v;
    }
    double v = 0;
    for (int k = 0; k < 8; k++) v += bli[(size_t)i * 8 + k] * b[ndim + k];
    bn[i] = b[i] - v;
  }
  // L626-631: unscale, symmetrise
  for (int i = 0; i < ndim; i++) {
    for (int j = 0; j < ndim; j++) Hn[(size_t)i * ndim + j] = (S[i] * Hn[(size_t)i * ndim + j]) * S[j];
    bn[i] = S[i] * bn[i];
  }
  HM.assign((size_t)ndim * ndim, 0.0);
  for (int i = 0; i < ndim; i++)
    for (int j = 0; j < ndim; j++) HM[(size_t)i * ndim + j] = 0.5 * (Hn[(size_t)i * ndim + j] + Hn[(size_t)j * ndim + i]);
  bM = bn;
}

// EnergyFunctional::insertFrame (EnergyFunctional.cpp:L453-496): the prior grows by 8 zero rows / columns for the new keyframe
inline void growHM(std::vector<double>& HM, std::vector<double>& bM, int nFramesOld) {
  const int o = nFramesOld * 8 + 4, n = o + 8;
  std::vector<double> H((size_t)n * n, 0.0), b(n, 0.0);
  if ((int)bM.size() == o && (int)HM.size() == o * o)
    for (int i = 0; i < o; i++) {
      b[i] = bM[i];
      for (int j = 0; j < o; j++) H[(size_t)i * n + j] = HM[(size_t)i * o + j];
    }
  HM.swap(H);
  bM.swap(b);
}

}  // namespace dmvio_b200
