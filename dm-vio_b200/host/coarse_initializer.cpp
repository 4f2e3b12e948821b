// See coarse_initializer.h.  Reference: src/dso/FullSystem/CoarseInitializer.cpp (line ranges cited per function).
#include "coarse_initializer.h"

#include <algorithm>
#include <cmath>
#include <cstdio>

namespace dmvio_b200 {

namespace {
// Hl.ldlt().solve(b) on the float 6x6 / 8x8 system of trackFrame (CoarseInitializer.cpp:L164-176): LDL^T with diagonal pivoting like Eigen's
void ldltSolveFloat(int n, const float* Ain, const float* b, float* x) {
  std::vector<float> L(Ain, Ain + (size_t)n * n), D(n, 0.f), y(n);
  std::vector<int> perm(n);
  for (int i = 0; i < n; i++) perm[i] = i;
  auto at = [&](int i, int j) -> float& { return L[(size_t)i * n + j]; };
  for (int k = 0; k < n; k++) {
    int piv = k;
    for (int i = k + 1; i < n; i++)
      if (std::fabs(at(i, i)) > std::fabs(at(piv, piv))) piv = i;
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(at(k, j), at(piv, j));
      for (int i = 0; i < n; i++) std::swap(at(i, k), at(i, piv));
      std::swap(perm[k], perm[piv]);
    }
    const float d = at(k, k);
    D[k] = d;
    if (d == 0.f) continue;
    for (int i = k + 1; i < n; i++) at(i, k) /= d;
    for (int j = k + 1; j < n; j++) {
      const float ljk = at(j, k) * d;
      for (int i = j; i < n; i++) { at(i, j) -= at(i, k) * ljk; at(j, i) = at(i, j); }
    }
  }
  for (int i = 0; i < n; i++) y[i] = b[perm[i]];
  for (int i = 0; i < n; i++)
    for (int k = 0; k < i; k++) y[i] -= at(i, k) * y[k];
  for (int i = 0; i < n; i++) y[i] = (D[i] != 0.f) ? y[i] / D[i] : 0.f;
  for (int i = n - 1; i >= 0; i--)
    for (int k = i + 1; k < n; k++) y[i] -= at(k, i) * y[k];
  for (int i = 0; i < n; i++) x[perm[i]] = y[i];
}
}  // namespace

CoarseInitializer::CoarseInitializer(int w, int h, int levels, int max_points, int device) : levels_(levels) {  // L49-73
  for (int l = 0; l < DMV_MAX_PYR_LEVELS; l++) { w_[l] = w >> l; h_[l] = h >> l; fx_[l] = fy_[l] = 1; cx_[l] = cy_[l] = 0; }
  wM_[0] = wM_[1] = wM_[2] = SCALE_XI_ROT;
  wM_[3] = wM_[4] = wM_[5] = SCALE_XI_TRANS;
  wM_[6] = SCALE_A;
  wM_[7] = SCALE_B;
  dmv_ci_config cfg{w, h, levels, max_points, device};
  if (dmv_ci_create(&cfg, &ci_) != DMV_OK) { err_ = dmv_last_error(); ci_ = nullptr; }
}
CoarseInitializer::~CoarseInitializer() { if (ci_) dmv_ci_destroy(ci_); }
bool CoarseInitializer::fail(const char* what) {
  err_ = std::string(what) + ": " + dmv_last_error();
  return false;
}

void CoarseInitializer::makeK(const CalibHessian& HCalib) {  // L967-999
  fx_[0] = HCalib.value_scaledf[0]; fy_[0] = HCalib.value_scaledf[1]; cx_[0] = HCalib.value_scaledf[2]; cy_[0] = HCalib.value_scaledf[3];
  for (int l = 1; l < levels_; l++) {
    fx_[l] = fx_[l - 1] * 0.5;
    fy_[l] = fy_[l - 1] * 0.5;
    cx_[l] = (cx_[0] + 0.5) / ((int)1 << l) - 0.5;
    cy_[l] = (cy_[0] + 0.5) / ((int)1 << l) - 0.5;
  }
  for (int l = 0; l < levels_; l++) {  // K.inverse() of the upper-triangular 3x3 in double: cofactors / determinant
    const double det = fx_[l] * fy_[l], id = 1.0 / det;
    double* Ki = Ki_[l];
    Ki[0] = fy_[l] * id; Ki[1] = 0; Ki[2] = (0.0 * cy_[l] - cx_[l] * fy_[l]) * id;
    Ki[3] = 0; Ki[4] = fx_[l] * id; Ki[5] = -(fx_[l] * cy_[l] - cx_[l] * 0.0) * id;
    Ki[6] = 0; Ki[7] = 0; Ki[8] = det * id;
    dmv_ci_set_K(ci_, l, (float)fx_[l], (float)fy_[l], (float)cx_[l], (float)cy_[l]);
  }
}

bool CoarseInitializer::setFirst(const float* const* dIp, float ab_exposure) {  // L804-889 (points[lvl] filled by the caller)
  err_.clear();
  if (!ci_) return false;
  first_exposure_ = ab_exposure;
  size_t maxn = 0;
  for (int l = 0; l < levels_; l++) {
    if (dmv_ci_upload_first(ci_, l, dIp[l]) != DMV_OK) return fail("dmv_ci_upload_first");
    maxn = std::max(maxn, points[l].size());
    std::vector<float> u(points[l].size()), v(points[l].size()), th(points[l].size());
    for (size_t i = 0; i < points[l].size(); i++) {
      Pnt& p = points[l][i];
      p.idepth = 1; p.iR = 1; p.isGood = true; p.energy[0] = p.energy[1] = 0; p.lastHessian = 0; p.lastHessian_new = 0;
      p.outlierTH = patternNum * s.setting_outlierTH;
      u[i] = p.u; v[i] = p.v; th[i] = p.outlierTH;
    }
    if (dmv_ci_set_points(ci_, l, (int)points[l].size(), u.data(), v.data(), th.data()) != DMV_OK) return fail("dmv_ci_set_points");
  }
  JbBuffer_.assign(maxn, std::array<float, 10>());
  JbBuffer_new_.assign(maxn, std::array<float, 10>());
  thisToNext = SE3();
  thisToNext_aff = AffLight();
  snapped = false;
  frameID = snappedAt = 0;
  points_uploaded_ = true;
  return true;
}

bool CoarseInitializer::calcResAndGS(int lvl, System& out, const SE3& refToNew, AffLight refToNew_aff, float res3[3]) {  // L333-625
  const int n = (int)points[lvl].size();
  std::vector<float> idn(n), en((size_t)2 * n), iR(n), en_new((size_t)2 * n), mstep(n), lastH(n), jb((size_t)10 * n);
  std::vector<uint8_t> good(n), good_new(n);
  for (int i = 0; i < n; i++) {
    const Pnt& p = points[lvl][i];
    idn[i] = p.idepth_new; good[i] = p.isGood; en[2 * i] = p.energy[0]; en[2 * i + 1] = p.energy[1]; iR[i] = p.iR;
  }
  dmv_ci_eval_args a;
  std::memset(&a, 0, sizeof(a));
  a.level = lvl;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      a.RKi[3 * i + j] = (float)(refToNew.R[3 * i] * Ki_[lvl][j] + refToNew.R[3 * i + 1] * Ki_[lvl][3 + j] + refToNew.R[3 * i + 2] * Ki_[lvl][6 + j]);
  double lg[6];
  refToNew.log(lg);
  for (int i = 0; i < 3; i++) { a.t_d[i] = refToNew.t[i]; a.t_log[i] = lg[i]; }
  a.r2new_aff[0] = (float)std::exp(refToNew_aff.a); a.r2new_aff[1] = (float)refToNew_aff.b;
  a.huberTH = s.setting_huberTH; a.alphaK = alphaK; a.alphaW = alphaW; a.couplingWeight = couplingWeight;
  a.weightZeroPriorX = weightZeroPriorDSOInitX; a.weightZeroPriorY = weightZeroPriorDSOInitY;
  a.idepth_new = idn.data(); a.isGood = good.data(); a.energy2 = en.data(); a.iR = iR.data();
  a.isGood_new = good_new.data(); a.energy_new2 = en_new.data(); a.maxstep = mstep.data(); a.lastHessian_new = lastH.data(); a.JbBuffer_new10 = jb.data();
  dmv_ci_eval_result r;
  if (dmv_ci_calc_res_and_gs(ci_, &a, &r) != DMV_OK) return fail("dmv_ci_calc_res_and_gs");
  evaluations++;
  for (int i = 0; i < n; i++) {  // what processPointsForReduce and the Schur pass leave in the Pnt array / JbBuffer_new (L369-586)
    Pnt& p = points[lvl][i];
    p.isGood_new = good_new[i] != 0;
    p.energy_new[0] = en_new[2 * i]; p.energy_new[1] = en_new[2 * i + 1];
    p.maxstep = p.isGood ? mstep[i] : 1e10f;
    if (p.isGood_new) {
      p.lastHessian_new = lastH[i];
      for (int k = 0; k < 10; k++) JbBuffer_new_[i][k] = jb[(size_t)10 * i + k];
    }
  }
  std::memcpy(out.H, r.H, sizeof(out.H)); std::memcpy(out.b, r.b, sizeof(out.b));
  std::memcpy(out.Hsc, r.Hsc, sizeof(out.Hsc)); std::memcpy(out.bsc, r.bsc, sizeof(out.bsc));
  for (int i = 0; i < 3; i++) res3[i] = r.res3[i];
  return true;
}

void CoarseInitializer::calcEC(int lvl, float out3[3]) {  // L650-670; AccumulatorX<2> (MatrixAccumulators.h:L177-244): float sums with 1k / 1M tiers
  if (!snapped) { out3[0] = 0; out3[1] = 0; out3[2] = (float)points[lvl].size(); return; }
  float A[2] = {0, 0}, A1k[2] = {0, 0}, A1m[2] = {0, 0};
  float n1 = 0, n1k = 0, n1m = 0;
  auto shiftUp = [&](bool force) {
    if (n1 > 1000 || force) { for (int k = 0; k < 2; k++) { A1k[k] += A[k]; A[k] = 0; } n1k += n1; n1 = 0; }
    if (n1k > 1000 || force) { for (int k = 0; k < 2; k++) { A1m[k] += A1k[k]; A1k[k] = 0; } n1m += n1k; n1k = 0; }
  };
  for (const Pnt& p : points[lvl]) {
    if (!p.isGood_new) continue;
    const float rOld = p.idepth - p.iR, rNew = p.idepth_new - p.iR;
    A[0] += rOld * rOld; A[1] += rNew * rNew;
    n1++;
    shiftUp(false);
  }
  shiftUp(true);
  out3[0] = couplingWeight * A1m[0]; out3[1] = couplingWeight * A1m[1]; out3[2] = (float)(size_t)(n1 + n1k + n1m);
}

void CoarseInitializer::optReg(int lvl) {  // L671-706: median of the neighbours' regularised depths
  if (!snapped) return;
  std::vector<Pnt>& pts = points[lvl];
  for (Pnt& p : pts) {
    if (!p.isGood) continue;
    float idnn[10];
    int nnn = 0;
    for (int j = 0; j < 10; j++) {
      if (p.neighbours[j] == -1) continue;
      const Pnt& o = pts[p.neighbours[j]];
      if (!o.isGood) continue;
      idnn[nnn++] = o.iR;
    }
    if (nnn > 2) {
      std::nth_element(idnn, idnn + nnn / 2, idnn + nnn);
      p.iR = (1 - regWeight) * p.idepth + regWeight * idnn[nnn / 2];
    }
  }
}

void CoarseInitializer::propagateUp(int srcLvl) {  // L708-747
  std::vector<Pnt>& src = points[srcLvl];
  std::vector<Pnt>& dst = points[srcLvl + 1];
  for (Pnt& parent : dst) { parent.iR = 0; parent.iRSumNum = 0; }
  for (const Pnt& p : src) {
    if (!p.isGood) continue;
    Pnt& parent = dst[p.parent];
    parent.iR += p.iR * p.lastHessian;
    parent.iRSumNum += p.lastHessian;
  }
  for (Pnt& parent : dst)
    if (parent.iRSumNum > 0) {
      parent.idepth = parent.iR = parent.iR / parent.iRSumNum;
      parent.isGood = true;
    }
  optReg(srcLvl + 1);
}

void CoarseInitializer::propagateDown(int srcLvl) {  // L749-777
  const std::vector<Pnt>& src = points[srcLvl];
  for (Pnt& p : points[srcLvl - 1]) {
    const Pnt& parent = src[p.parent];
    if (!parent.isGood || parent.lastHessian < 0.1) continue;
    if (!p.isGood) {
      p.iR = p.idepth = p.idepth_new = parent.iR;
      p.isGood = true;
      p.lastHessian = 0;
    } else {
      const float newiR = (p.iR * p.lastHessian * 2 + parent.iR * parent.lastHessian) / (p.lastHessian * 2 + parent.lastHessian);
      p.iR = p.idepth = p.idepth_new = newiR;
    }
  }
  optReg(srcLvl - 1);
}

void CoarseInitializer::resetPoints(int lvl) {  // L891-917
  std::vector<Pnt>& pts = points[lvl];
  for (Pnt& p : pts) {
    p.energy[0] = p.energy[1] = 0;
    p.idepth_new = p.idepth;
    if (lvl == levels_ - 1 && !p.isGood) {
      float snd = 0, sn = 0;
      for (int n = 0; n < 10; n++) {
        if (p.neighbours[n] == -1 || !pts[p.neighbours[n]].isGood) continue;
        snd += pts[p.neighbours[n]].iR;
        sn += 1;
      }
      if (sn > 0) {
        p.isGood = true;
        p.iR = p.idepth = p.idepth_new = snd / sn;
      }
    }
  }
}

void CoarseInitializer::doStep(int lvl, float lambda, const float inc[8]) {  // L919-946
  const float maxPixelStep = 0.25f, idMaxStep = 1e10f;
  std::vector<Pnt>& pts = points[lvl];
  for (size_t i = 0; i < pts.size(); i++) {
    Pnt& p = pts[i];
    if (!p.isGood) continue;
    float dot = 0;
    for (int k = 0; k < 8; k++) dot += JbBuffer_[i][k] * inc[k];
    const float b = JbBuffer_[i][8] + dot;
    float step = -b * JbBuffer_[i][9] / (1 + lambda);
    float maxstep = maxPixelStep * p.maxstep;
    if (maxstep > idMaxStep) maxstep = idMaxStep;
    step = std::min(maxstep, std::max(-maxstep, step));
    float newIdepth = p.idepth + step;
    if (newIdepth < 1e-3f) newIdepth = 1e-3f;
    if (newIdepth > 50) newIdepth = 50;
    p.idepth_new = newIdepth;
  }
}

void CoarseInitializer::applyStep(int lvl) {  // L948-965
  for (Pnt& p : points[lvl]) {
    if (!p.isGood) {
      p.idepth = p.idepth_new = p.iR;
      continue;
    }
    p.energy[0] = p.energy_new[0]; p.energy[1] = p.energy_new[1];
    p.isGood = p.isGood_new;
    p.idepth = p.idepth_new;
    p.lastHessian = p.lastHessian_new;
  }
  std::swap(JbBuffer_, JbBuffer_new_);
}

bool CoarseInitializer::trackFrame(const float* const* dIp, float ab_exposure) {  // L85-282
  err_.clear();
  if (!ci_ || !points_uploaded_) { err_ = "setFirst first"; return false; }
  for (int l = 0; l < levels_; l++)
    if (dmv_ci_upload_new(ci_, l, dIp[l]) != DMV_OK) return fail("dmv_ci_upload_new");
  new_exposure_ = ab_exposure;
  const int maxIterations[] = {5, 5, 10, 30, 50, 50};
  alphaK = 2.5f * 2.5f;
  alphaW = 150 * 150;
  regWeight = 0.8f;
  couplingWeight = 1;
  if (!snapped) {
    thisToNext.t[0] = thisToNext.t[1] = thisToNext.t[2] = 0;
    for (int lvl = 0; lvl < levels_; lvl++)
      for (Pnt& p : points[lvl]) { p.iR = 1; p.idepth_new = 1; p.lastHessian = 0; }
  }
  SE3 refToNew_current = thisToNext;
  AffLight refToNew_aff_current = thisToNext_aff;
  if (first_exposure_ > 0 && new_exposure_ > 0) {
    refToNew_aff_current.a = logf(new_exposure_ / first_exposure_);  // coarse approximation
    refToNew_aff_current.b = 0;
  }
  for (int lvl = levels_ - 1; lvl >= 0; lvl--) {
    if (lvl < levels_ - 1) propagateDown(lvl + 1);
    System S;
    resetPoints(lvl);
    float resOld[3];
    if (!calcResAndGS(lvl, S, refToNew_current, refToNew_aff_current, resOld)) return false;
    applyStep(lvl);
    float lambda = 0.1f;
    const float eps = 1e-4f;
    int fails = 0, iteration = 0;
    while (true) {
      float Hl[64], bl[8];
      for (int i = 0; i < 64; i++) Hl[i] = S.H[i];
      for (int i = 0; i < 8; i++) Hl[i * 8 + i] *= (1 + lambda);
      for (int i = 0; i < 64; i++) Hl[i] -= S.Hsc[i] * (1 / (1 + lambda));
      for (int i = 0; i < 8; i++) bl[i] = S.b[i] - S.bsc[i] * (1 / (1 + lambda));
      const float sc = 0.01f / (w_[lvl] * h_[lvl]);
      for (int i = 0; i < 8; i++) {
        for (int j = 0; j < 8; j++) Hl[i * 8 + j] = ((wM_[i] * Hl[i * 8 + j]) * wM_[j]) * sc;
        bl[i] = (wM_[i] * bl[i]) * sc;
      }
      float inc[8];
      if (fixAffine) {
        float H6[36], x6[6];
        for (int i = 0; i < 6; i++)
          for (int j = 0; j < 6; j++) H6[i * 6 + j] = Hl[i * 8 + j];
        ldltSolveFloat(6, H6, bl, x6);
        for (int i = 0; i < 6; i++) inc[i] = -(wM_[i] * x6[i]);
        inc[6] = inc[7] = 0;
      } else {
        float x8[8];
        ldltSolveFloat(8, Hl, bl, x8);
        for (int i = 0; i < 8; i++) inc[i] = -(wM_[i] * x8[i]);
      }
      float n2 = 0;
      for (int i = 0; i < 8; i++) n2 += inc[i] * inc[i];
      const double incNorm = std::sqrt(n2);
      double e6[6];
      for (int i = 0; i < 6; i++) e6[i] = (double)inc[i];
      const SE3 refToNew_new = SE3::exp(e6) * refToNew_current;
      AffLight refToNew_aff_new = refToNew_aff_current;
      refToNew_aff_new.a += inc[6];
      refToNew_aff_new.b += inc[7];
      doStep(lvl, lambda, inc);
      System Sn;
      float resNew[3], regEnergy[3];
      if (!calcResAndGS(lvl, Sn, refToNew_new, refToNew_aff_new, resNew)) return false;
      calcEC(lvl, regEnergy);
      const float eTotalNew = resNew[0] + resNew[1] + regEnergy[1];
      const float eTotalOld = resOld[0] + resOld[1] + regEnergy[0];
      if (eTotalOld > eTotalNew) {  // accept
        if (resNew[1] == alphaK * (int)points[lvl].size()) snapped = true;
        S = Sn;
        for (int i = 0; i < 3; i++) resOld[i] = resNew[i];
        refToNew_aff_current = refToNew_aff_new;
        refToNew_current = refToNew_new;
        applyStep(lvl);
        optReg(lvl);
        lambda *= 0.5f;
        fails = 0;
        if (lambda < 0.0001f) lambda = 0.0001f;
      } else {
        fails++;
        lambda *= 4;
        if (lambda > 10000) lambda = 10000;
      }
      if (!(incNorm > eps) || iteration >= maxIterations[lvl] || fails >= 2) break;
      iteration++;
    }
  }
  thisToNext = refToNew_current;
  thisToNext_aff = refToNew_aff_current;
  for (int i = 0; i < levels_ - 1; i++) propagateUp(i);
  frameID++;
  if (!snapped) snappedAt = 0;
  if (snapped && snappedAt == 0) snappedAt = frameID;
  return snapped && frameID > snappedAt + 5;
}

}  // namespace dmvio_b200
