// TEST INFRASTRUCTURE ONLY — see orc_init.h.  Every function cites the reference lines it restates (FullSystem/CoarseInitializer.cpp,
// OptimizationBackend/MatrixAccumulators.h).  float arithmetic follows the reference operation by operation (the build has no FMA).
#include "orc_init.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace orc {

namespace {

// MatrixAccumulators.h:L91-175 (Accumulator11): 4 float lanes, 1 / 1k / 1M tiers
struct Acc11 {
  float A = 0;
  size_t num = 0;
  float d[4], d1k[4], d1m[4];
  float numIn1 = 0, numIn1k = 0, numIn1m = 0;
  void initialize() {
    A = 0;
    std::memset(d, 0, sizeof(d)); std::memset(d1k, 0, sizeof(d1k)); std::memset(d1m, 0, sizeof(d1m));
    num = 0; numIn1 = numIn1k = numIn1m = 0;
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int k = 0; k < 4; k++) d1k[k] = d[k] + d1k[k];
      numIn1k += numIn1; numIn1 = 0;
      std::memset(d, 0, sizeof(d));
    }
    if (numIn1k > 1000 || force) {
      for (int k = 0; k < 4; k++) d1m[k] = d1k[k] + d1m[k];
      numIn1m += numIn1k; numIn1k = 0;
      std::memset(d1k, 0, sizeof(d1k));
    }
  }
  void updateSingle(float val) { d[0] += val; num++; numIn1++; shiftUp(false); }
  void finish() { shiftUp(true); A = d1m[0] + d1m[1] + d1m[2] + d1m[3]; }
};

// MatrixAccumulators.h:L982-1345 (Accumulator9): 45 upper-triangular entries x 4 lanes, 1 / 1k / 1M tiers
struct Acc9 {
  float H[9][9];
  size_t num = 0;
  float d[45][4], d1k[45][4], d1m[45][4];
  float numIn1 = 0, numIn1k = 0, numIn1m = 0;
  void initialize() {
    std::memset(H, 0, sizeof(H));
    std::memset(d, 0, sizeof(d)); std::memset(d1k, 0, sizeof(d1k)); std::memset(d1m, 0, sizeof(d1m));
    num = 0; numIn1 = numIn1k = numIn1m = 0;
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < 45; i++) for (int k = 0; k < 4; k++) d1k[i][k] = d[i][k] + d1k[i][k];
      numIn1k += numIn1; numIn1 = 0;
      std::memset(d, 0, sizeof(d));
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < 45; i++) for (int k = 0; k < 4; k++) d1m[i][k] = d1k[i][k] + d1m[i][k];
      numIn1m += numIn1k; numIn1k = 0;
      std::memset(d1k, 0, sizeof(d1k));
    }
  }
  // updateSSE (L1022-1076): J[r] are 4-lane vectors
  void updateSSE(const float J[9][4]) {
    int idx = 0;
    for (int r = 0; r < 9; r++)
      for (int c = r; c < 9; c++) {
        for (int k = 0; k < 4; k++) d[idx][k] = d[idx][k] + J[r][k] * J[c][k];
        idx++;
      }
    num += 4; numIn1++;
    shiftUp(false);
  }
  // updateSingleWeighted (L1251-1318), lane 0
  void updateSingleWeighted(float J[9], float w) {
    int idx = 0;
    for (int r = 0; r < 9; r++) {
      d[idx][0] += J[r] * J[r] * w; idx++;
      J[r] *= w;
      for (int c = r + 1; c < 9; c++) { d[idx][0] += J[c] * J[r]; idx++; }
    }
    num++; numIn1++;
    shiftUp(false);
  }
  void finish() {  // L1001-1017
    std::memset(H, 0, sizeof(H));
    shiftUp(true);
    int idx = 0;
    for (int r = 0; r < 9; r++)
      for (int c = r; c < 9; c++) {
        const float v = d1m[idx][0] + d1m[idx][1] + d1m[idx][2] + d1m[idx][3];
        H[r][c] = H[c][r] = v;
        idx++;
      }
  }
};

inline float interp31(const float* mat, float x, float y, int width) {  // util/globalFuncs.h:L160-174 (channel 0 of the [I,dx,dy] AoS)
  const int ix = (int)x, iy = (int)y;
  const float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  return dxdy * bp[3 * (1 + width)] + (dy - dxdy) * bp[3 * width] + (dx - dxdy) * bp[3] + (1 - dx - dy + dxdy) * bp[0];
}
inline void interp33(const float* mat, float x, float y, int width, float out[3]) {  // util/globalFuncs.h:L103-118
  const int ix = (int)x, iy = (int)y;
  const float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  for (int c = 0; c < 3; c++)
    out[c] = dxdy * bp[3 * (1 + width) + c] + (dy - dxdy) * bp[3 * width + c] + (dx - dxdy) * bp[3 + c] + (1 - dx - dy + dxdy) * bp[c];
}

// the pivoted LDL^T of the stand-in matrix headers the reference is compiled against (oracle/shim/Eigen/Core), in float: x = A^-1 b
void ldltSolveF(int n, const float* Ain, const float* b, float* x) {
  std::vector<float> L(Ain, Ain + (size_t)n * n), Dg(n, 0.f), y(n);
  std::vector<int> perm(n);
  for (int i = 0; i < n; i++) perm[i] = i;
  auto at = [&](int i, int j) -> float& { return L[(size_t)i * n + j]; };
  for (int k = 0; k < n; k++) {
    int piv = k;
    for (int i = k + 1; i < n; i++) if (std::abs(at(i, i)) > std::abs(at(piv, piv))) piv = i;
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(at(k, j), at(piv, j));
      for (int i = 0; i < n; i++) std::swap(at(i, k), at(i, piv));
      std::swap(perm[k], perm[piv]);
    }
    const float dd = at(k, k);
    Dg[k] = dd;
    if (dd == 0.f) continue;
    for (int i = k + 1; i < n; i++) at(i, k) /= dd;
    for (int j = k + 1; j < n; j++) {
      const float ljk = at(j, k) * dd;
      for (int i = j; i < n; i++) { at(i, j) -= at(i, k) * ljk; at(j, i) = at(i, j); }
    }
  }
  for (int i = 0; i < n; i++) y[i] = b[perm[i]];
  for (int i = 0; i < n; i++) for (int k = 0; k < i; k++) y[i] -= at(i, k) * y[k];
  for (int i = 0; i < n; i++) y[i] = (Dg[i] != 0.f) ? y[i] / Dg[i] : 0.f;
  for (int i = n - 1; i >= 0; i--) for (int k = i + 1; k < n; k++) y[i] -= at(k, i) * y[k];
  for (int i = 0; i < n; i++) x[perm[i]] = y[i];
}

}  // namespace

CoarseInit::CoarseInit() {  // CoarseInitializer.cpp:L49-73
  for (int l = 0; l < PYR_LEVELS; l++) { dIFirst[l] = dINew[l] = nullptr; w[l] = h[l] = 0; }
  wM[0] = wM[1] = wM[2] = SCALE_XI_ROT;
  wM[3] = wM[4] = wM[5] = SCALE_XI_TRANS;
  wM[6] = SCALE_A;
  wM[7] = SCALE_B;
}

void CoarseInit::makeK(int w0, int h0, double fx0, double fy0, double cx0, double cy0, int forceLevels) {  // L967-999
  GlobalCalib g;
  g.set(w0, h0, (float)fx0, (float)fy0, (float)cx0, (float)cy0, forceLevels);
  levels = g.pyrLevelsUsed;
  w[0] = w0; h[0] = h0;
  fx[0] = fx0; fy[0] = fy0; cx[0] = cx0; cy[0] = cy0;
  for (int level = 1; level < levels; ++level) {
    w[level] = w[0] >> level;
    h[level] = h[0] >> level;
    fx[level] = fx[level - 1] * 0.5;
    fy[level] = fy[level - 1] * 0.5;
    cx[level] = (cx[0] + 0.5) / ((int)1 << level) - 0.5;
    cy[level] = (cy[0] + 0.5) / ((int)1 << level) - 0.5;
  }
  for (int level = 0; level < levels; ++level) {
    Mat33 K;
    K(0, 0) = fx[level]; K(0, 1) = 0; K(0, 2) = cx[level];
    K(1, 0) = 0; K(1, 1) = fy[level]; K(1, 2) = cy[level];
    K(2, 0) = 0; K(2, 1) = 0; K(2, 2) = 1;
    Ki[level] = inverse3_cofactor<double>(K);
  }
}

void CoarseInit::setFirst(const float* const* dIp, float exposure) {  // L804-889 (points[lvl] filled by the caller)
  for (int l = 0; l < levels; l++) dIFirst[l] = dIp[l];
  first_exposure = exposure;
  size_t maxn = 0;
  for (int l = 0; l < levels; l++) {
    maxn = std::max(maxn, points[l].size());
    for (InitPnt& p : points[l]) {
      p.idepth = 1; p.iR = 1; p.isGood = true; p.energy[0] = p.energy[1] = 0; p.lastHessian = 0; p.lastHessian_new = 0;
      p.outlierTH = PATTERN_NUM * s.outlierTH;
    }
  }
  JbBuffer.assign(maxn, std::array<float, 10>());
  JbBuffer_new.assign(maxn, std::array<float, 10>());
  thisToNext = SE3();
  thisToNext_aff = AffLight();
  snapped = false;
  frameID = snappedAt = 0;
}

void CoarseInit::calcResAndGS(int lvl, InitSystem& out, const SE3& refToNew, AffLight refToNew_aff, float res3[3]) {  // L333-625
  const int wl = w[lvl], hl = h[lvl];
  const float* colorRef = dIFirst[lvl];
  const float* colorNew = dINew[lvl];
  const Mat33 RKid = refToNew.rotationMatrix() * Ki[lvl];
  float RKi[9], t[3];
  for (int i = 0; i < 9; i++) RKi[i] = (float)RKid.d[i];
  for (int i = 0; i < 3; i++) t[i] = (float)refToNew.t[i];
  const float r2new_aff[2] = {(float)std::exp(refToNew_aff.a), (float)refToNew_aff.b};
  const float fxl = (float)fx[lvl], fyl = (float)fy[lvl], cxl = (float)cx[lvl], cyl = (float)cy[lvl];

  Acc9 acc9;
  Acc11 E;
  acc9.initialize();
  E.initialize();
  const int npts = (int)points[lvl].size();
  InitPnt* ptsl = points[lvl].data();

  for (int i = 0; i < npts; i++) {  // processPointsForReduce (L369-504), single worker
    InitPnt* point = ptsl + i;
    point->maxstep = 1e10;
    if (!point->isGood) {
      E.updateSingle((float)(point->energy[0]));
      point->energy_new[0] = point->energy[0]; point->energy_new[1] = point->energy[1];
      point->isGood_new = false;
      continue;
    }
    float dp[8][8], dd[8], r[8];  // dp[k][idx]
    std::array<float, 10>& Jb = JbBuffer_new[i];
    Jb.fill(0.f);
    bool isGood = true;
    float energy = 0;
    for (int idx = 0; idx < PATTERN_NUM; idx++) {
      const int dx = patternP[idx][0], dy = patternP[idx][1];
      const float px = point->u + dx, py = point->v + dy;
      float pt[3];
      for (int k = 0; k < 3; k++) pt[k] = (RKi[3 * k] * px + RKi[3 * k + 1] * py + RKi[3 * k + 2] * 1.0f) + t[k] * point->idepth_new;
      const float u = pt[0] / pt[2], v = pt[1] / pt[2];
      const float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
      const float new_idepth = point->idepth_new / pt[2];
      if (!(Ku > 1 && Kv > 1 && Ku < wl - 2 && Kv < hl - 2 && new_idepth > 0)) { isGood = false; break; }
      float hitColor[3];
      interp33(colorNew, Ku, Kv, wl, hitColor);
      const float rlR = interp31(colorRef, point->u + dx, point->v + dy, wl);
      if (!std::isfinite(rlR) || !std::isfinite((float)hitColor[0])) { isGood = false; break; }
      const float residual = hitColor[0] - r2new_aff[0] * rlR - r2new_aff[1];
      float hw = std::fabs(residual) < s.huberTH ? 1 : s.huberTH / std::fabs(residual);
      energy += hw * residual * residual * (2 - hw);
      const float dxdd = (t[0] - t[2] * u) / pt[2];
      const float dydd = (t[1] - t[2] * v) / pt[2];
      if (hw < 1) hw = sqrtf(hw);
      const float dxInterp = hw * hitColor[1] * fxl;
      const float dyInterp = hw * hitColor[2] * fyl;
      dp[0][idx] = new_idepth * dxInterp;
      dp[1][idx] = new_idepth * dyInterp;
      dp[2][idx] = -new_idepth * (u * dxInterp + v * dyInterp);
      dp[3][idx] = -u * v * dxInterp - (1 + v * v) * dyInterp;
      dp[4][idx] = (1 + u * u) * dxInterp + u * v * dyInterp;
      dp[5][idx] = -v * dxInterp + u * dyInterp;
      dp[6][idx] = -hw * r2new_aff[0] * rlR;
      dp[7][idx] = -hw * 1;
      dd[idx] = dxInterp * dxdd + dyInterp * dydd;
      r[idx] = hw * residual;
      const float mx = dxdd * fxl, my = dydd * fyl;
      const float maxstep = 1.0f / std::sqrt(mx * mx + my * my);
      if (maxstep < point->maxstep) point->maxstep = maxstep;
      for (int k = 0; k < 8; k++) Jb[k] += dp[k][idx] * dd[idx];
      Jb[8] += r[idx] * dd[idx];
      Jb[9] += dd[idx] * dd[idx];
    }
    if (!isGood || energy > point->outlierTH * 20) {
      E.updateSingle((float)(point->energy[0]));
      point->isGood_new = false;
      point->energy_new[0] = point->energy[0]; point->energy_new[1] = point->energy[1];
      continue;
    }
    E.updateSingle(energy);
    point->isGood_new = true;
    point->energy_new[0] = energy;
    for (int i4 = 0; i4 + 3 < PATTERN_NUM; i4 += 4) {  // L478-489: lanes = pixels i4 .. i4+3
      float J[9][4];
      for (int k = 0; k < 8; k++) for (int l = 0; l < 4; l++) J[k][l] = dp[k][i4 + l];
      for (int l = 0; l < 4; l++) J[8][l] = r[i4 + l];
      acc9.updateSSE(J);
    }
  }
  acc9.finish();
  E.finish();

  // L520-542: the alpha energy.  The reference adds these terms to accE[0] AFTER its finish() (dso issue #52): E.A is unchanged, E.num grows.
  Acc11 EAlpha;
  EAlpha.initialize();
  for (int i = 0; i < npts; i++) {
    InitPnt* point = ptsl + i;
    if (!point->isGood_new) {
      E.updateSingle((float)(point->energy[1]));
    } else {
      point->energy_new[1] = (point->idepth_new - 1) * (point->idepth_new - 1);
      E.updateSingle((float)(point->energy_new[1]));
    }
  }
  EAlpha.finish();
  const double tsq = refToNew.t[0] * refToNew.t[0] + refToNew.t[1] * refToNew.t[1] + refToNew.t[2] * refToNew.t[2];
  float alphaEnergy = alphaW * (EAlpha.A + tsq * npts);
  float alphaOpt;
  if (alphaEnergy > alphaK * npts) { alphaOpt = 0; alphaEnergy = alphaK * npts; }
  else alphaOpt = alphaW;

  Acc9 acc9SC;
  acc9SC.initialize();
  for (int i = 0; i < npts; i++) {  // L562-586
    InitPnt* point = ptsl + i;
    if (!point->isGood_new) continue;
    std::array<float, 10>& Jb = JbBuffer_new[i];
    point->lastHessian_new = Jb[9];
    Jb[8] += alphaOpt * (point->idepth_new - 1);
    Jb[9] += alphaOpt;
    if (alphaOpt == 0) {
      Jb[8] += couplingWeight * (point->idepth_new - point->iR);
      Jb[9] += couplingWeight;
    }
    Jb[9] = 1 / (1 + Jb[9]);
    float J[9] = {Jb[0], Jb[1], Jb[2], Jb[3], Jb[4], Jb[5], Jb[6], Jb[7], Jb[8]};
    acc9SC.updateSingleWeighted(J, Jb[9]);
  }
  acc9SC.finish();

  for (int a = 0; a < 8; a++) {
    for (int b = 0; b < 8; b++) { out.H[a * 8 + b] = 0.f + acc9.H[a][b]; out.Hsc[a * 8 + b] = acc9SC.H[a][b]; }
    out.b[a] = 0.f + acc9.H[a][8];
    out.bsc[a] = acc9SC.H[a][8];
  }
  out.H[0] += alphaOpt * npts;
  out.H[9] += alphaOpt * npts;
  out.H[18] += alphaOpt * npts;
  const Vec6 lg = refToNew.log();
  const float tlog[3] = {(float)lg[0], (float)lg[1], (float)lg[2]};
  out.b[0] += tlog[0] * alphaOpt * npts;
  out.b[1] += tlog[1] * alphaOpt * npts;
  out.b[2] += tlog[2] * alphaOpt * npts;
  // L606-611: zero prior on the translation (weights default to 0)
  out.H[9] = (float)(out.H[9] + weightZeroPriorDSOInitY);
  out.b[1] = (float)(out.b[1] + weightZeroPriorDSOInitY * refToNew.t[1]);
  out.H[0] = (float)(out.H[0] + weightZeroPriorDSOInitX);
  out.b[0] = (float)(out.b[0] + weightZeroPriorDSOInitX * refToNew.t[0]);

  double A = 0;
  int num = 0;
  A += E.A;
  num += (int)E.num;
  res3[0] = (float)A; res3[1] = alphaEnergy; res3[2] = (float)num;
}

void CoarseInit::calcEC(int lvl, float out3[3]) {  // L650-670 (AccumulatorX<2>, MatrixAccumulators.h:L177-244)
  if (!snapped) { out3[0] = 0; out3[1] = 0; out3[2] = (float)points[lvl].size(); return; }
  float A[2] = {0, 0}, A1k[2] = {0, 0}, A1m[2] = {0, 0};
  float numIn1 = 0, numIn1k = 0, numIn1m = 0;
  auto shiftUp = [&](bool force) {
    if (numIn1 > 1000 || force) { for (int k = 0; k < 2; k++) { A1k[k] += A[k]; A[k] = 0; } numIn1k += numIn1; numIn1 = 0; }
    if (numIn1k > 1000 || force) { for (int k = 0; k < 2; k++) { A1m[k] += A1k[k]; A1k[k] = 0; } numIn1m += numIn1k; numIn1k = 0; }
  };
  for (const InitPnt& p : points[lvl]) {
    if (!p.isGood_new) continue;
    const float rOld = (p.idepth - p.iR);
    const float rNew = (p.idepth_new - p.iR);
    A[0] += rOld * rOld; A[1] += rNew * rNew;
    numIn1++;
    shiftUp(false);
  }
  shiftUp(true);
  const size_t num = (size_t)(numIn1 + numIn1k + numIn1m);
  out3[0] = couplingWeight * A1m[0]; out3[1] = couplingWeight * A1m[1]; out3[2] = (float)num;
}

void CoarseInit::optReg(int lvl) {  // L671-706
  if (!snapped) return;
  InitPnt* ptsl = points[lvl].data();
  const int npts = (int)points[lvl].size();
  for (int i = 0; i < npts; i++) {
    InitPnt* point = ptsl + i;
    if (!point->isGood) continue;
    float idnn[10];
    int nnn = 0;
    for (int j = 0; j < 10; j++) {
      if (point->neighbours[j] == -1) continue;
      InitPnt* other = ptsl + point->neighbours[j];
      if (!other->isGood) continue;
      idnn[nnn] = other->iR;
      nnn++;
    }
    if (nnn > 2) {
      std::nth_element(idnn, idnn + nnn / 2, idnn + nnn);
      point->iR = (1 - regWeight) * point->idepth + regWeight * idnn[nnn / 2];
    }
  }
}

void CoarseInit::propagateUp(int srcLvl) {  // L708-747
  std::vector<InitPnt>& ptss = points[srcLvl];
  std::vector<InitPnt>& ptst = points[srcLvl + 1];
  for (InitPnt& parent : ptst) { parent.iR = 0; parent.iRSumNum = 0; }
  for (InitPnt& point : ptss) {
    if (!point.isGood) continue;
    InitPnt& parent = ptst[point.parent];
    parent.iR += point.iR * point.lastHessian;
    parent.iRSumNum += point.lastHessian;
  }
  for (InitPnt& parent : ptst) {
    if (parent.iRSumNum > 0) {
      parent.idepth = parent.iR = (parent.iR / parent.iRSumNum);
      parent.isGood = true;
    }
  }
  optReg(srcLvl + 1);
}

void CoarseInit::propagateDown(int srcLvl) {  // L749-777
  std::vector<InitPnt>& ptss = points[srcLvl];
  std::vector<InitPnt>& ptst = points[srcLvl - 1];
  for (InitPnt& point : ptst) {
    InitPnt& parent = ptss[point.parent];
    if (!parent.isGood || parent.lastHessian < 0.1) continue;
    if (!point.isGood) {
      point.iR = point.idepth = point.idepth_new = parent.iR;
      point.isGood = true;
      point.lastHessian = 0;
    } else {
      const float newiR = (point.iR * point.lastHessian * 2 + parent.iR * parent.lastHessian) / (point.lastHessian * 2 + parent.lastHessian);
      point.iR = point.idepth = point.idepth_new = newiR;
    }
  }
  optReg(srcLvl - 1);
}

void CoarseInit::resetPoints(int lvl) {  // L891-917
  std::vector<InitPnt>& pts = points[lvl];
  for (InitPnt& p : pts) {
    p.energy[0] = p.energy[1] = 0;
    p.idepth_new = p.idepth;
    if (lvl == levels - 1 && !p.isGood) {
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

void CoarseInit::doStep(int lvl, float lambda, const float inc[8]) {  // L919-946
  const float maxPixelStep = 0.25;
  const float idMaxStep = 1e10;
  std::vector<InitPnt>& pts = points[lvl];
  for (size_t i = 0; i < pts.size(); i++) {
    if (!pts[i].isGood) continue;
    float dot = 0;
    for (int k = 0; k < 8; k++) dot += JbBuffer[i][k] * inc[k];
    const float b = JbBuffer[i][8] + dot;
    float step = -b * JbBuffer[i][9] / (1 + lambda);
    float maxstep = maxPixelStep * pts[i].maxstep;
    if (maxstep > idMaxStep) maxstep = idMaxStep;
    if (step > maxstep) step = maxstep;
    if (step < -maxstep) step = -maxstep;
    float newIdepth = pts[i].idepth + step;
    if (newIdepth < 1e-3) newIdepth = 1e-3;
    if (newIdepth > 50) newIdepth = 50;
    pts[i].idepth_new = newIdepth;
  }
}

void CoarseInit::applyStep(int lvl) {  // L948-965
  for (InitPnt& p : points[lvl]) {
    if (!p.isGood) {
      p.idepth = p.idepth_new = p.iR;
      continue;
    }
    p.energy[0] = p.energy_new[0]; p.energy[1] = p.energy_new[1];
    p.isGood = p.isGood_new;
    p.idepth = p.idepth_new;
    p.lastHessian = p.lastHessian_new;
  }
  std::swap(JbBuffer, JbBuffer_new);
}

bool CoarseInit::trackFrame(const float* const* dIp, float exposure) {  // L85-282
  for (int l = 0; l < levels; l++) dINew[l] = dIp[l];
  new_exposure = exposure;
  const int maxIterations[] = {5, 5, 10, 30, 50};
  alphaK = 2.5 * 2.5;
  alphaW = 150 * 150;
  regWeight = 0.8;
  couplingWeight = 1;
  if (!snapped) {
    thisToNext.t[0] = thisToNext.t[1] = thisToNext.t[2] = 0;
    for (int lvl = 0; lvl < levels; lvl++)
      for (InitPnt& p : points[lvl]) { p.iR = 1; p.idepth_new = 1; p.lastHessian = 0; }
  }
  SE3 refToNew_current = thisToNext;
  AffLight refToNew_aff_current = thisToNext_aff;
  if (first_exposure > 0 && new_exposure > 0) {
    refToNew_aff_current.a = logf(new_exposure / first_exposure);  // coarse approximation
    refToNew_aff_current.b = 0;
  }
  float latestRes[3] = {0, 0, 0};
  for (int lvl = levels - 1; lvl >= 0; lvl--) {
    if (lvl < levels - 1) propagateDown(lvl + 1);
    InitSystem S;
    resetPoints(lvl);
    float resOld[3];
    calcResAndGS(lvl, S, refToNew_current, refToNew_aff_current, resOld);
    applyStep(lvl);
    float lambda = 0.1;
    const float eps = 1e-4;
    int fails = 0;
    int iteration = 0;
    while (true) {
      float Hl[64], bl[8];
      for (int i = 0; i < 64; i++) Hl[i] = S.H[i];
      for (int i = 0; i < 8; i++) Hl[i * 8 + i] *= (1 + lambda);
      for (int i = 0; i < 64; i++) Hl[i] -= S.Hsc[i] * (1 / (1 + lambda));
      for (int i = 0; i < 8; i++) bl[i] = S.b[i] - S.bsc[i] * (1 / (1 + lambda));
      const float sc = (0.01f / (w[lvl] * h[lvl]));
      for (int i = 0; i < 8; i++) {
        for (int j = 0; j < 8; j++) Hl[i * 8 + j] = ((wM[i] * Hl[i * 8 + j]) * wM[j]) * sc;
        bl[i] = (wM[i] * bl[i]) * sc;
      }
      float inc[8];
      if (fixAffine) {
        float H6[36], x6[6];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H6[i * 6 + j] = Hl[i * 8 + j];
        ldltSolveF(6, H6, bl, x6);
        for (int i = 0; i < 6; i++) {  // -(wM.toDenseMatrix().topLeftCorner<6,6>() * x): a full 6x6 product with the (diagonal) dense matrix
          float acc = 0;
          for (int k = 0; k < 6; k++) acc += ((i == k) ? wM[i] : 0.f) * x6[k];
          inc[i] = -acc;
        }
        inc[6] = inc[7] = 0;
      } else {
        float x8[8];
        ldltSolveF(8, Hl, bl, x8);
        for (int i = 0; i < 8; i++) inc[i] = -(wM[i] * x8[i]);
      }
      float n2 = 0;
      for (int i = 0; i < 8; i++) n2 += inc[i] * inc[i];
      const double incNorm = std::sqrt(n2);
      Vec6 e6;
      for (int i = 0; i < 6; i++) e6[i] = (double)inc[i];
      const SE3 refToNew_new = SE3::exp(e6) * refToNew_current;
      AffLight refToNew_aff_new = refToNew_aff_current;
      refToNew_aff_new.a += inc[6];
      refToNew_aff_new.b += inc[7];
      doStep(lvl, lambda, inc);
      InitSystem Sn;
      float resNew[3], regEnergy[3];
      calcResAndGS(lvl, Sn, refToNew_new, refToNew_aff_new, resNew);
      calcEC(lvl, regEnergy);
      const float eTotalNew = (resNew[0] + resNew[1] + regEnergy[1]);
      const float eTotalOld = (resOld[0] + resOld[1] + regEnergy[0]);
      const bool accept = eTotalOld > eTotalNew;
      if (accept) {
        if (resNew[1] == alphaK * (int)points[lvl].size()) snapped = true;
        S = Sn;
        for (int i = 0; i < 3; i++) resOld[i] = resNew[i];
        refToNew_aff_current = refToNew_aff_new;
        refToNew_current = refToNew_new;
        applyStep(lvl);
        optReg(lvl);
        lambda *= 0.5;
        fails = 0;
        if (lambda < 0.0001) lambda = 0.0001;
      } else {
        fails++;
        lambda *= 4;
        if (lambda > 10000) lambda = 10000;
      }
      bool quitOpt = false;
      if (!(incNorm > eps) || iteration >= maxIterations[lvl] || fails >= 2) quitOpt = true;
      if (quitOpt) break;
      iteration++;
    }
    for (int i = 0; i < 3; i++) latestRes[i] = resOld[i];
  }
  thisToNext = refToNew_current;
  thisToNext_aff = refToNew_aff_current;
  for (int i = 0; i < levels - 1; i++) propagateUp(i);
  frameID++;
  if (!snapped) snappedAt = 0;
  if (snapped && snappedAt == 0) snappedAt = frameID;
  return snapped && frameID > snappedAt + 5;
}

}  // namespace orc
