// TEST INFRASTRUCTURE ONLY — see orc_coarse.h.
#include "orc_coarse.h"
#include <algorithm>
#include <cstdio>

namespace orc {

void GlobalCalib::set(int w, int h, float fx, float fy, float cx, float cy, int forceLevels) {
  // util/globalCalib.cpp:L45-105
  int wlvl = w, hlvl = h;
  pyrLevelsUsed = 1;
  while (wlvl % 2 == 0 && hlvl % 2 == 0 && wlvl * hlvl > 5000 && pyrLevelsUsed < PYR_LEVELS) {
    wlvl /= 2; hlvl /= 2; pyrLevelsUsed++;
  }
  if (forceLevels > 0) pyrLevelsUsed = forceLevels;
  wG[0] = w; hG[0] = h; fxG[0] = fx; fyG[0] = fy; cxG[0] = cx; cyG[0] = cy;
  for (int level = 1; level < pyrLevelsUsed; ++level) {
    wG[level] = w >> level;
    hG[level] = h >> level;
    fxG[level] = fxG[level - 1] * 0.5;
    fyG[level] = fyG[level - 1] * 0.5;
    cxG[level] = (cxG[0] + 0.5) / ((int)1 << level) - 0.5;
    cyG[level] = (cyG[0] + 0.5) / ((int)1 << level) - 0.5;
  }
}

void makeImages(const GlobalCalib& g, const float* color, float* const* dIp, float* const* absSquaredGrad) {
  // HessianBlocks.cpp:L128-191
  int w = g.wG[0], h = g.hG[0];
  for (int lvl = 0; lvl < g.pyrLevelsUsed; lvl++) std::fill(dIp[lvl], dIp[lvl] + (size_t)g.wG[lvl] * g.hG[lvl] * 3, 0.0f);
  for (int i = 0; i < w * h; i++) dIp[0][3 * i] = color[i];
  for (int lvl = 0; lvl < g.pyrLevelsUsed; lvl++) {
    int wl = g.wG[lvl], hl = g.hG[lvl];
    float* dI_l = dIp[lvl];
    if (lvl > 0) {
      int wlm1 = g.wG[lvl - 1];
      const float* dI_lm = dIp[lvl - 1];
      for (int y = 0; y < hl; y++)
        for (int x = 0; x < wl; x++)
          dI_l[3 * (x + y * wl)] = 0.25f * (dI_lm[3 * (2 * x + 2 * y * wlm1)] + dI_lm[3 * (2 * x + 1 + 2 * y * wlm1)] +
                                            dI_lm[3 * (2 * x + 2 * y * wlm1 + wlm1)] + dI_lm[3 * (2 * x + 1 + 2 * y * wlm1 + wlm1)]);
    }
    for (int idx = wl; idx < wl * (hl - 1); idx++) {
      float dx = 0.5f * (dI_l[3 * (idx + 1)] - dI_l[3 * (idx - 1)]);
      float dy = 0.5f * (dI_l[3 * (idx + wl)] - dI_l[3 * (idx - wl)]);
      if (!std::isfinite(dx)) dx = 0;
      if (!std::isfinite(dy)) dy = 0;
      dI_l[3 * idx + 1] = dx;
      dI_l[3 * idx + 2] = dy;
      if (absSquaredGrad && absSquaredGrad[lvl]) absSquaredGrad[lvl][idx] = dx * dx + dy * dy;
    }
  }
}

bool initPointColorWeights(const float* dI, int w, float u, float v, float outlierTHSumComponent, float* color8, float* weights8) {
  // ImmaturePoint.cpp:L36-62 with util/globalFuncs.h:L203-226 (getInterpolatedElement33BiLin)
  for (int idx = 0; idx < PATTERN_NUM; idx++) {
    float x = u + patternP[idx][0], y = v + patternP[idx][1];
    int ix = (int)x, iy = (int)y;
    const float* bp = dI + 3 * (ix + iy * w);
    float tl = bp[0], tr = bp[3], bl = bp[3 * w], br = bp[3 * w + 3];
    float dx = x - ix, dy = y - iy;
    float topInt = dx * tr + (1 - dx) * tl;
    float botInt = dx * br + (1 - dx) * bl;
    float leftInt = dy * bl + (1 - dy) * tl;
    float rightInt = dy * br + (1 - dy) * tr;
    float ptc[3] = {dx * rightInt + (1 - dx) * leftInt, rightInt - leftInt, botInt - topInt};
    color8[idx] = ptc[0];
    if (!std::isfinite(color8[idx])) return false;
    weights8[idx] = sqrtf(outlierTHSumComponent / (outlierTHSumComponent + (ptc[1] * ptc[1] + ptc[2] * ptc[2])));
  }
  return true;
}

void CoarseTracker::makeK(const GlobalCalib& g) {  // CoarseTracker.cpp:L105-134
  levels = g.pyrLevelsUsed;
  for (int l = 0; l < levels; l++) {
    w[l] = g.wG[l]; h[l] = g.hG[l]; fx[l] = g.fxG[l]; fy[l] = g.fyG[l]; cx[l] = g.cxG[l]; cy[l] = g.cyG[l];
    Mat33f K;
    K(0, 0) = fx[l]; K(0, 2) = cx[l]; K(1, 1) = fy[l]; K(1, 2) = cy[l]; K(2, 2) = 1;
    Ki[l] = inverse3_cofactor(K);  // K[level].inverse() in float, as Eigen evaluates it (CoarseTracker.cpp:L128)
  }
}

void CoarseTracker::makeCoarseDepthL0(int n, const float* Ku, const float* Kv, const float* new_id, const float* HdiF, const float* const* refdIp) {
  // CoarseTracker.cpp:L138-295
  std::vector<float> idepth[PYR_LEVELS], weightSums[PYR_LEVELS], weightSums_bak[PYR_LEVELS];
  for (int l = 0; l < levels; l++) {
    idepth[l].assign((size_t)w[l] * h[l], 0.f);
    weightSums[l].assign((size_t)w[l] * h[l], 0.f);
    weightSums_bak[l].assign((size_t)w[l] * h[l], 0.f);
  }
  for (int i = 0; i < n; i++) {
    int u = Ku[i] + 0.5f;
    int v = Kv[i] + 0.5f;
    float weight = sqrtf(1e-3 / (HdiF[i] + 1e-12));
    idepth[0][u + w[0] * v] += new_id[i] * weight;
    weightSums[0][u + w[0] * v] += weight;
  }
  for (int lvl = 1; lvl < levels; lvl++) {
    int lvlm1 = lvl - 1;
    int wl = w[lvl], hl = h[lvl], wlm1 = w[lvlm1];
    for (int y = 0; y < hl; y++)
      for (int x = 0; x < wl; x++) {
        int bidx = 2 * x + 2 * y * wlm1;
        idepth[lvl][x + y * wl] = idepth[lvlm1][bidx] + idepth[lvlm1][bidx + 1] + idepth[lvlm1][bidx + wlm1] + idepth[lvlm1][bidx + wlm1 + 1];
        weightSums[lvl][x + y * wl] =
            weightSums[lvlm1][bidx] + weightSums[lvlm1][bidx + 1] + weightSums[lvlm1][bidx + wlm1] + weightSums[lvlm1][bidx + wlm1 + 1];
      }
  }
  for (int lvl = 0; lvl < std::min(2, levels); lvl++) {  // dilate by 1 (diagonal neighbours)
    int wh = w[lvl] * h[lvl] - w[lvl];
    int wl = w[lvl];
    float* ws = weightSums[lvl].data();
    float* bak = weightSums_bak[lvl].data();
    std::copy(ws, ws + (size_t)w[lvl] * h[lvl], bak);
    float* idl = idepth[lvl].data();
    for (int i = w[lvl] + 1; i < wh - 1; i++) {
      if (bak[i] <= 0) {
        float sum = 0, num = 0, numn = 0;
        if (bak[i + 1 + wl] > 0) { sum += idl[i + 1 + wl]; num += bak[i + 1 + wl]; numn++; }
        if (bak[i - 1 - wl] > 0) { sum += idl[i - 1 - wl]; num += bak[i - 1 - wl]; numn++; }
        if (bak[i + wl - 1] > 0) { sum += idl[i + wl - 1]; num += bak[i + wl - 1]; numn++; }
        if (bak[i - wl + 1] > 0) { sum += idl[i - wl + 1]; num += bak[i - wl + 1]; numn++; }
        if (numn > 0) { idl[i] = sum / numn; ws[i] = num / numn; }
      }
    }
  }
  for (int lvl = 2; lvl < levels; lvl++) {  // dilate by 1 (4-neighbourhood)
    int wh = w[lvl] * h[lvl] - w[lvl];
    int wl = w[lvl];
    float* ws = weightSums[lvl].data();
    float* bak = weightSums_bak[lvl].data();
    std::copy(ws, ws + (size_t)w[lvl] * h[lvl], bak);
    float* idl = idepth[lvl].data();
    for (int i = w[lvl] + 1; i < wh - 1; i++) {
      if (bak[i] <= 0) {
        float sum = 0, num = 0, numn = 0;
        if (bak[i + 1] > 0) { sum += idl[i + 1]; num += bak[i + 1]; numn++; }
        if (bak[i - 1] > 0) { sum += idl[i - 1]; num += bak[i - 1]; numn++; }
        if (bak[i + wl] > 0) { sum += idl[i + wl]; num += bak[i + wl]; numn++; }
        if (bak[i - wl] > 0) { sum += idl[i - wl]; num += bak[i - wl]; numn++; }
        if (numn > 0) { idl[i] = sum / numn; ws[i] = num / numn; }
      }
    }
  }
  for (int lvl = 0; lvl < levels; lvl++) {
    float* ws = weightSums[lvl].data();
    float* idl = idepth[lvl].data();
    const float* dIRefl = refdIp[lvl];
    int wl = w[lvl], hl = h[lvl];
    pc_u[lvl].clear(); pc_v[lvl].clear(); pc_idepth[lvl].clear(); pc_color[lvl].clear();
    for (int y = 2; y < hl - 2; y++)
      for (int x = 2; x < wl - 2; x++) {
        int i = x + y * wl;
        if (ws[i] > 0) {
          idl[i] /= ws[i];
          float col = dIRefl[3 * i];
          if (!std::isfinite(col) || !(idl[i] > 0)) { idl[i] = -1; continue; }
          pc_u[lvl].push_back((float)x); pc_v[lvl].push_back((float)y); pc_idepth[lvl].push_back(idl[i]); pc_color[lvl].push_back(col);
        } else
          idl[i] = -1;
        ws[i] = 1;
      }
    pc_n[lvl] = (int)pc_u[lvl].size();
  }
}

void CoarseTracker::calcRes(int lvl, const SE3& refToNew, AffLight aff_g2l, float cutoffTH, double out[6]) {
  // CoarseTracker.cpp:L361-517: the operands (L377-379), then the point loop
  Mat33f RKi = refToNew.rotationMatrix().cast<float>() * Ki[lvl];
  Vec3f t = refToNew.translation().cast<float>();
  double aff2[2];
  AffLight::fromToVecExposure(lastRef_ab_exposure, newFrame_ab_exposure, lastRef_aff_g2l, aff_g2l, aff2);
  float affLL[2] = {(float)aff2[0], (float)aff2[1]};
  calcResRaw(lvl, RKi, t, affLL, cutoffTH, out);
}

void CoarseTracker::calcResRaw(int lvl, const Mat33f& RKi, const Vec3f& t, const float affLL[2], float cutoffTH, double out[6]) {
  float E = 0;
  int numTermsInE = 0, numTermsInWarped = 0, numSaturated = 0;
  int wl = w[lvl], hl = h[lvl];
  const float* dINewl = newFrame_dIp[lvl];
  float fxl = fx[lvl], fyl = fy[lvl], cxl = cx[lvl], cyl = cy[lvl];
  float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
  float maxEnergy = 2 * s.huberTH * cutoffTH - s.huberTH * s.huberTH;
  int nl = pc_n[lvl];
  const float* lpc_u = pc_u[lvl].data();
  const float* lpc_v = pc_v[lvl].data();
  const float* lpc_idepth = pc_idepth[lvl].data();
  const float* lpc_color = pc_color[lvl].data();
  size_t cap = (size_t)nl + 4;
  buf_warped_idepth.resize(cap); buf_warped_u.resize(cap); buf_warped_v.resize(cap); buf_warped_dx.resize(cap); buf_warped_dy.resize(cap);
  buf_warped_residual.resize(cap); buf_warped_weight.resize(cap); buf_warped_refColor.resize(cap);
  const Mat33f& Kil = Ki[lvl];
  for (int i = 0; i < nl; i++) {
    float id = lpc_idepth[i], x = lpc_u[i], y = lpc_v[i];
    float pt[3];
    for (int k = 0; k < 3; k++) pt[k] = RKi(k, 0) * x + RKi(k, 1) * y + RKi(k, 2) + t[k] * id;
    float u = pt[0] / pt[2], v = pt[1] / pt[2];
    float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
    float new_idepth = id / pt[2];
    if (lvl == 0 && i % 32 == 0) {
      float ptT[3], ptT2[3], pt3[3];
      for (int k = 0; k < 3; k++) {
        float kp = Kil(k, 0) * x + Kil(k, 1) * y + Kil(k, 2);
        ptT[k] = kp + t[k] * id;
        ptT2[k] = kp - t[k] * id;
        pt3[k] = RKi(k, 0) * x + RKi(k, 1) * y + RKi(k, 2) - t[k] * id;
      }
      float KuT = fxl * (ptT[0] / ptT[2]) + cxl, KvT = fyl * (ptT[1] / ptT[2]) + cyl;
      float KuT2 = fxl * (ptT2[0] / ptT2[2]) + cxl, KvT2 = fyl * (ptT2[1] / ptT2[2]) + cyl;
      float Ku3 = fxl * (pt3[0] / pt3[2]) + cxl, Kv3 = fyl * (pt3[1] / pt3[2]) + cyl;
      sumSquaredShiftT += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      sumSquaredShiftT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      sumSquaredShiftRT += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      sumSquaredShiftRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      sumSquaredShiftNum += 2;
    }
    if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;
    float refColor = lpc_color[i];
    float hitColor[3];
    {
      int ix = (int)Ku, iy = (int)Kv;
      float dx = Ku - ix, dy = Kv - iy, dxdy = dx * dy;
      const float* bp = dINewl + 3 * (ix + iy * wl);
      for (int c = 0; c < 3; c++)
        hitColor[c] = dxdy * bp[3 * (1 + wl) + c] + (dy - dxdy) * bp[3 * wl + c] + (dx - dxdy) * bp[3 + c] + (1 - dx - dy + dxdy) * bp[c];
    }
    if (!std::isfinite(hitColor[0])) continue;
    float residual = hitColor[0] - (float)(affLL[0] * refColor + affLL[1]);
    float hw = fabs(residual) < s.huberTH ? 1 : s.huberTH / fabs(residual);
    if (fabs(residual) > cutoffTH) {
      E += maxEnergy;
      numTermsInE++;
      numSaturated++;
    } else {
      E += hw * residual * residual * (2 - hw);
      numTermsInE++;
      buf_warped_idepth[numTermsInWarped] = new_idepth;
      buf_warped_u[numTermsInWarped] = u;
      buf_warped_v[numTermsInWarped] = v;
      buf_warped_dx[numTermsInWarped] = hitColor[1];
      buf_warped_dy[numTermsInWarped] = hitColor[2];
      buf_warped_residual[numTermsInWarped] = residual;
      buf_warped_weight[numTermsInWarped] = hw;
      buf_warped_refColor[numTermsInWarped] = lpc_color[i];
      numTermsInWarped++;
    }
  }
  while (numTermsInWarped % 4 != 0) {
    buf_warped_idepth[numTermsInWarped] = 0; buf_warped_u[numTermsInWarped] = 0; buf_warped_v[numTermsInWarped] = 0;
    buf_warped_dx[numTermsInWarped] = 0; buf_warped_dy[numTermsInWarped] = 0; buf_warped_residual[numTermsInWarped] = 0;
    buf_warped_weight[numTermsInWarped] = 0; buf_warped_refColor[numTermsInWarped] = 0;
    numTermsInWarped++;
  }
  buf_warped_n = numTermsInWarped;
  out[0] = E;
  out[1] = numTermsInE;
  out[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1);
  out[3] = 0;
  out[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1);
  out[5] = numSaturated / (float)numTermsInE;
}

template <class T>
static void gsAccumulate(const CoarseTracker& ct, int lvl, float a, float b0, Mat<double, 9, 9>& Hout) {
  // CoarseTracker.cpp:L299-340 + MatrixAccumulators.h:L982-1345 (Accumulator9, 4 SSE lanes, 1/1k/1M tiers)
  const int n = ct.buf_warped_n;
  const float fxl = ct.fx[lvl], fyl = ct.fy[lvl];
  static thread_local std::vector<T> D, D1k, D1m;
  D.assign(4 * 45, 0); D1k.assign(4 * 45, 0); D1m.assign(4 * 45, 0);
  float numIn1 = 0, numIn1k = 0;
  auto shiftUp = [&](bool force) {
    if (numIn1 > 1000 || force) { for (int i = 0; i < 180; i++) { D1k[i] += D[i]; D[i] = 0; } numIn1k += numIn1; numIn1 = 0; }
    if (numIn1k > 1000 || force) { for (int i = 0; i < 180; i++) { D1m[i] += D1k[i]; D1k[i] = 0; } numIn1k = 0; }
  };
  for (int i = 0; i < n; i += 4) {
    for (int l = 0; l < 4; l++) {
      const int k = i + l;
      float dx = ct.buf_warped_dx[k] * fxl, dy = ct.buf_warped_dy[k] * fyl;
      float u = ct.buf_warped_u[k], v = ct.buf_warped_v[k], id = ct.buf_warped_idepth[k];
      float J[9];
      J[0] = id * dx;
      J[1] = id * dy;
      J[2] = 0 - id * (u * dx + v * dy);
      J[3] = 0 - ((u * v) * dx + dy * (1 + v * v));
      J[4] = (u * v) * dy + dx * (1 + u * u);
      J[5] = u * dy - v * dx;
      J[6] = a * (b0 - ct.buf_warped_refColor[k]);
      J[7] = -1;
      J[8] = ct.buf_warped_residual[k];
      float w = ct.buf_warped_weight[k];
      int e = 0;
      for (int r = 0; r < 9; r++) {
        T Jw = (T)J[r] * (T)w;
        for (int c = r; c < 9; c++) { D[4 * e + l] += Jw * (T)J[c]; e++; }
      }
    }
    numIn1++;
    shiftUp(false);
  }
  shiftUp(true);
  int idx = 0;
  for (int r = 0; r < 9; r++)
    for (int c = r; c < 9; c++) {
      T d = D1m[idx + 0] + D1m[idx + 1] + D1m[idx + 2] + D1m[idx + 3];
      Hout(r, c) = Hout(c, r) = (double)d;
      idx += 4;
    }
}

void CoarseTracker::calcGSSSE(int lvl, Mat88& H_out, Vec8& b_out, const SE3& /*refToNew*/, AffLight aff_g2l, int precision) {
  double aff2[2];
  AffLight::fromToVecExposure(lastRef_ab_exposure, newFrame_ab_exposure, lastRef_aff_g2l, aff_g2l, aff2);
  calcGSRaw(lvl, H_out, b_out, (float)aff2[0], (float)lastRef_aff_g2l.b, precision);
}

void CoarseTracker::calcGSRaw(int lvl, Mat88& H_out, Vec8& b_out, float a, float b0, int precision) {
  Mat<double, 9, 9> H;
  if (precision == 0) gsAccumulate<float>(*this, lvl, a, b0, H);
  else gsAccumulate<double>(*this, lvl, a, b0, H);
  const int n = buf_warped_n;
  const double inv = (double)(1.0f / n);
  for (int i = 0; i < 8; i++) {
    for (int j = 0; j < 8; j++) H_out(i, j) = H(i, j) * inv;
    b_out[i] = H(i, 8) * inv;
  }
  // L344-355 scaling (SCALE_XI_ROT/TRANS = 1; a,b columns and rows)
  for (int i = 0; i < 8; i++) { H_out(i, 6) *= SCALE_A; H_out(i, 7) *= SCALE_B; }
  for (int j = 0; j < 8; j++) { H_out(6, j) *= SCALE_A; H_out(7, j) *= SCALE_B; }
  b_out[6] *= SCALE_A;
  b_out[7] *= SCALE_B;
}

bool CoarseTracker::trackNewestCoarse(SE3& lastToNew_out, AffLight& aff_g2l_out, int coarsestLvl, const double minResForAbort[5], int precision,
                                      int* totalIterations) {
  // CoarseTracker.cpp:L539-770, visual-only branch (L639-683)
  for (int i = 0; i < 5; i++) lastResiduals[i] = NAN;
  for (int i = 0; i < 3; i++) lastFlowIndicators[i] = 1000;
  int maxIterations[] = {10, 20, 50, 50, 50};
  float lambdaExtrapolationLimit = 0.001;
  SE3 refToNew_current = lastToNew_out;
  AffLight aff_g2l_current = aff_g2l_out;
  bool haveRepeated = false;
  int its = 0;
  Mat88 H; Vec8 b;
  for (int lvl = coarsestLvl; lvl >= 0; lvl--) {
    float levelCutoffRepeat = 1;
    double resOld[6];
    calcRes(lvl, refToNew_current, aff_g2l_current, s.coarseCutoffTH * levelCutoffRepeat, resOld);
    while (resOld[5] > 0.6 && (levelCutoffRepeat < 50 || resOld[5] > 0.99)) {
      levelCutoffRepeat *= 2;
      calcRes(lvl, refToNew_current, aff_g2l_current, s.coarseCutoffTH * levelCutoffRepeat, resOld);
    }
    calcGSSSE(lvl, H, b, refToNew_current, aff_g2l_current, precision);
    float lambda = 0.01;
    for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
      its++;
      Mat88 Hl = H;
      for (int i = 0; i < 8; i++) Hl(i, i) *= (1 + lambda);
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = sqrt(sqrt(lambdaExtrapolationLimit / lambda));
      double inc[8];
      {
        int n = 8;
        if (s.affineOptModeA < 0 && s.affineOptModeB < 0) n = 6;
        else if (!(s.affineOptModeA < 0) && s.affineOptModeB < 0) n = 7;
        MatX A(n, n);
        VecX rhs(n), x;
        if (s.affineOptModeA < 0 && !(s.affineOptModeB < 0)) {  // fix a: stitch b into slot 6
          n = 7;
          A = MatX(7, 7); rhs.assign(7, 0.0);
          int map[7] = {0, 1, 2, 3, 4, 5, 7};
          for (int i = 0; i < 7; i++) { for (int j = 0; j < 7; j++) A(i, j) = Hl(map[i], map[j]); rhs[i] = -b[map[i]]; }
          ldlt_solve(A, rhs, x);
          for (int i = 0; i < 8; i++) inc[i] = 0;
          for (int i = 0; i < 6; i++) inc[i] = x[i];
          inc[7] = x[6];
        } else {
          for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) A(i, j) = Hl(i, j); rhs[i] = -b[i]; }
          ldlt_solve(A, rhs, x);
          for (int i = 0; i < 8; i++) inc[i] = (i < n) ? x[i] : 0.0;
        }
      }
      for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
      double incScaled[8];
      for (int i = 0; i < 8; i++) incScaled[i] = inc[i];
      incScaled[6] *= SCALE_A;
      incScaled[7] *= SCALE_B;
      double ssum = 0;
      for (int i = 0; i < 8; i++) ssum += incScaled[i];
      if (!std::isfinite(ssum)) for (int i = 0; i < 8; i++) incScaled[i] = 0;
      Vec6 xi; for (int i = 0; i < 6; i++) xi[i] = incScaled[i];
      SE3 refToNew_new = SE3::exp(xi) * refToNew_current;
      AffLight aff_g2l_new = aff_g2l_current;
      aff_g2l_new.a += incScaled[6];
      aff_g2l_new.b += incScaled[7];
      double incNorm = 0;
      for (int i = 0; i < 8; i++) incNorm += inc[i] * inc[i];
      incNorm = std::sqrt(incNorm);
      double resNew[6];
      calcRes(lvl, refToNew_new, aff_g2l_new, s.coarseCutoffTH * levelCutoffRepeat, resNew);
      bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
      if (accept) {
        calcGSSSE(lvl, H, b, refToNew_new, aff_g2l_new, precision);
        for (int i = 0; i < 6; i++) resOld[i] = resNew[i];
        aff_g2l_current = aff_g2l_new;
        refToNew_current = refToNew_new;
        lambda *= 0.5;
      } else {
        lambda *= 4;
        if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
      }
      if (!(incNorm > 1e-3)) break;
    }
    lastResiduals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
    lastFlowIndicators[0] = resOld[2]; lastFlowIndicators[1] = resOld[3]; lastFlowIndicators[2] = resOld[4];
    if (std::isnan(lastResiduals[lvl])) { if (totalIterations) *totalIterations = its; return false; }
    if (lastResiduals[lvl] > 1.5 * minResForAbort[lvl]) { if (totalIterations) *totalIterations = its; return false; }
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
  }
  lastToNew_out = refToNew_current;
  aff_g2l_out = aff_g2l_current;
  bool trackingGood = true;
  if ((s.affineOptModeA != 0 && (fabsf((float)aff_g2l_out.a) > 1.2)) || (s.affineOptModeB != 0 && (fabsf((float)aff_g2l_out.b) > 200)))
    trackingGood = false;
  double rel[2];
  AffLight::fromToVecExposure(lastRef_ab_exposure, newFrame_ab_exposure, lastRef_aff_g2l, aff_g2l_out, rel);
  if ((s.affineOptModeA == 0 && (fabsf(logf((float)rel[0])) > 1.5)) || (s.affineOptModeB == 0 && (fabsf((float)rel[1]) > 200)))
    trackingGood = false;
  if (s.affineOptModeA < 0) aff_g2l_out.a = 0;
  if (s.affineOptModeB < 0) aff_g2l_out.b = 0;
  if (totalIterations) *totalIterations = its;
  return trackingGood;
}

}  // namespace orc
