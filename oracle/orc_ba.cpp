// TEST INFRASTRUCTURE ONLY — see orc_ba.h.  CPU restatement of the BA hot path; pinned bit-exact against the compiled reference (oracle/_ref, tests/test_ref_pin.py).
#include "orc_ba.h"
#include "orc_threads.h"
#include <algorithm>
#include <cstdio>

namespace orc {

// ------------------------------------------------------------------------------------------------
// small state containers
// ------------------------------------------------------------------------------------------------
void Calib::setValue(const double* v) {  // HessianBlocks.h:L356-371
  for (int i = 0; i < 4; i++) value[i] = v[i];
  value_scaled[0] = SCALE_F * value[0];
  value_scaled[1] = SCALE_F * value[1];
  value_scaled[2] = SCALE_C * value[2];
  value_scaled[3] = SCALE_C * value[3];
  for (int i = 0; i < 4; i++) value_scaledf[i] = (float)value_scaled[i];
  value_scaledi[0] = 1.0f / value_scaledf[0];
  value_scaledi[1] = 1.0f / value_scaledf[1];
  value_scaledi[2] = -value_scaledf[2] / value_scaledf[0];
  value_scaledi[3] = -value_scaledf[3] / value_scaledf[1];
  for (int i = 0; i < 4; i++) value_minus_value_zero[i] = value[i] - value_zero[i];
}
void Calib::setValueScaled(const double* vs) {  // HessianBlocks.h:L373-387
  double v[4] = {vs[0] / SCALE_F, vs[1] / SCALE_F, vs[2] / SCALE_C, vs[3] / SCALE_C};
  for (int i = 0; i < 4; i++) value_scaled[i] = vs[i];
  for (int i = 0; i < 4; i++) value[i] = v[i];
  for (int i = 0; i < 4; i++) value_scaledf[i] = (float)value_scaled[i];
  value_scaledi[0] = 1.0f / value_scaledf[0];
  value_scaledi[1] = 1.0f / value_scaledf[1];
  value_scaledi[2] = -value_scaledf[2] / value_scaledf[0];
  value_scaledi[3] = -value_scaledf[3] / value_scaledf[1];
  for (int i = 0; i < 4; i++) value_minus_value_zero[i] = value[i] - value_zero[i];
}

void AffLight::fromToVecExposure(float exposureF, float exposureT, AffLight g2F, AffLight g2T, double out[2]) {
  // util/NumType.h:L174-186
  if (exposureF == 0 || exposureT == 0) exposureT = exposureF = 1;
  double a = std::exp(g2T.a - g2F.a) * exposureT / exposureF;
  double b = g2T.b - a * g2F.b;
  out[0] = a;
  out[1] = b;
}

void Frame::setState(const Vec10& s) {  // HessianBlocks.h:L172-186
  state = s;
  for (int i = 0; i < 3; i++) state_scaled[i] = SCALE_XI_TRANS * s[i];
  for (int i = 3; i < 6; i++) state_scaled[i] = SCALE_XI_ROT * s[i];
  state_scaled[6] = SCALE_A * s[6];
  state_scaled[7] = SCALE_B * s[7];
  state_scaled[8] = SCALE_A * s[8];
  state_scaled[9] = SCALE_B * s[9];
  Vec6 eps; for (int i = 0; i < 6; i++) eps[i] = state_scaled[i];
  PRE_worldToCam = SE3::exp(eps) * worldToCam_evalPT;
  PRE_camToWorld = PRE_worldToCam.inverse();
}
void Frame::setStateScaled(const Vec10& ss) {  // HessianBlocks.h:L187-202
  Vec10 s;
  for (int i = 0; i < 3; i++) s[i] = ss[i] / SCALE_XI_TRANS;
  for (int i = 3; i < 6; i++) s[i] = ss[i] / SCALE_XI_ROT;
  s[6] = ss[6] / SCALE_A; s[7] = ss[7] / SCALE_B; s[8] = ss[8] / SCALE_A; s[9] = ss[9] / SCALE_B;
  setState(s);
  state_scaled = ss;
}

// ------------------------------------------------------------------------------------------------
// precalc / adjoints / deltas (host-side quantities in the reference)
// ------------------------------------------------------------------------------------------------
void Window::setPrecalcValues() {
  const int n = nf();
  precalc.assign((size_t)n * n, FramePrecalc());
  for (int hI = 0; hI < n; hI++)
    for (int tI = 0; tI < n; tI++) {
      // HessianBlocks.cpp:L193-223 (FrameFramePrecalc::set)
      const Frame& host = frames[hI];
      const Frame& target = frames[tI];
      FramePrecalc& p = precalc[(size_t)hI * n + tI];
      SE3 leftToLeft_0 = target.worldToCam_evalPT * host.worldToCam_evalPT.inverse();
      p.PRE_RTll_0 = leftToLeft_0.rotationMatrix().cast<float>();
      p.PRE_tTll_0 = leftToLeft_0.translation().cast<float>();
      SE3 leftToLeft = target.PRE_worldToCam * host.PRE_camToWorld;
      p.PRE_RTll = leftToLeft.rotationMatrix().cast<float>();
      p.PRE_tTll = leftToLeft.translation().cast<float>();
      const Vec3& tt = leftToLeft.translation();
      p.distanceLL = (float)std::sqrt(tt[0] * tt[0] + tt[1] * tt[1] + tt[2] * tt[2]);
      Mat33f K;
      K(0, 0) = calib.fxl(); K(1, 1) = calib.fyl(); K(0, 2) = calib.cxl(); K(1, 2) = calib.cyl(); K(2, 2) = 1;
      // K.inverse(): Eigen evaluates a fixed-size 3x3 inverse in float by cofactors (det from the first column, ONE reciprocal,
      // every cofactor times that reciprocal), so e.g. Ki(0,2) = -(cx*fy) * (1/(fy*fx)), not -cx/fx; the rounding is visible
      // in PRE_KRKiTll (pinned against the compiled reference, tests/test_ref_pin.py)
      const Mat33f Ki = inverse3_cofactor(K);
      p.PRE_KRKiTll = K * p.PRE_RTll * Ki;
      p.PRE_RKiTll = p.PRE_RTll * Ki;
      p.PRE_KtTll = K * p.PRE_tTll;
      double aff[2];
      AffLight::fromToVecExposure(host.ab_exposure, target.ab_exposure, host.aff_g2l(), target.aff_g2l(), aff);
      p.PRE_aff_mode[0] = (float)aff[0];
      p.PRE_aff_mode[1] = (float)aff[1];
      p.PRE_b0_mode = (float)host.aff_g2l_0().b;
    }
  setDeltaF();
}

void Window::setAdjointsF() {  // EnergyFunctional.cpp:L48-108
  const int n = nf();
  adHost.assign((size_t)n * n, Mat88());
  adTarget.assign((size_t)n * n, Mat88());
  adHostF.assign((size_t)n * n, Mat88f());
  adTargetF.assign((size_t)n * n, Mat88f());
  for (int h = 0; h < n; h++)
    for (int t = 0; t < n; t++) {
      const Frame& host = frames[h];
      const Frame& target = frames[t];
      SE3 hostToTarget = target.worldToCam_evalPT * host.worldToCam_evalPT.inverse();
      Mat88 AH = Mat88::identity();
      Mat88 AT = Mat88::identity();
      Mat66 adj = hostToTarget.Adj();
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) AH(i, j) = -adj(j, i);  // -Adj^T
      double aff[2];
      AffLight::fromToVecExposure(host.ab_exposure, target.ab_exposure, host.aff_g2l_0(), target.aff_g2l_0(), aff);
      float affLL0 = (float)aff[0];
      AT(6, 6) = -affLL0;
      AH(6, 6) = affLL0;
      AT(7, 7) = -1;
      AH(7, 7) = affLL0;
      for (int c = 0; c < 8; c++) {
        for (int r = 0; r < 3; r++) { AH(r, c) *= SCALE_XI_TRANS; AT(r, c) *= SCALE_XI_TRANS; }
        for (int r = 3; r < 6; r++) { AH(r, c) *= SCALE_XI_ROT; AT(r, c) *= SCALE_XI_ROT; }
        AH(6, c) *= SCALE_A; AT(6, c) *= SCALE_A;
        AH(7, c) *= SCALE_B; AT(7, c) *= SCALE_B;
      }
      adHost[h + (size_t)t * n] = AH;
      adTarget[h + (size_t)t * n] = AT;
      adHostF[h + (size_t)t * n] = AH.cast<float>();
      adTargetF[h + (size_t)t * n] = AT.cast<float>();
    }
  for (int i = 0; i < 4; i++) cPrior[i] = s.initialCalibHessian;
}

static Vec10 framePrior(const Frame& f, const Settings& s) {  // HessianBlocks.h:L262-298 (getPrior)
  Vec10 p;
  if (f.frameID == 0) {
    for (int i = 0; i < 3; i++) p[i] = s.initialTransPrior;
    for (int i = 3; i < 6; i++) p[i] = s.initialRotPrior;
    p[6] = s.initialAffAPrior;
    p[7] = s.initialAffBPrior;
  } else {
    p[6] = (s.affineOptModeA < 0) ? s.initialAffAPrior : s.affineOptModeA;
    p[7] = (s.affineOptModeB < 0) ? s.initialAffBPrior : s.affineOptModeB;
  }
  p[8] = s.initialAffAPrior;
  p[9] = s.initialAffBPrior;
  if (f.addCamPrior) {
    for (int i = 0; i < 3; i++) p[i] = s.initialTransPrior;
    for (int i = 3; i < 6; i++) p[i] = s.initialRotPrior;
  }
  return p;
}

void Window::takeDataFrames() {  // EnergyFunctionalStructs.cpp:L52-85
  for (Frame& f : frames) {
    Vec10 p = framePrior(f, s);
    Vec10 d = f.get_state_minus_stateZero();
    for (int i = 0; i < 8; i++) { f.prior[i] = p[i]; f.delta[i] = d[i]; f.delta_prior[i] = f.state[i]; }
  }
  for (Point& p : points) {
    p.priorF = p.hasDepthPrior ? s.idepthFixPrior * SCALE_IDEPTH * SCALE_IDEPTH : 0;
    p.deltaF = p.idepth - p.idepth_zero;
  }
}

void Window::setDeltaF() {  // EnergyFunctional.cpp:L175-198
  const int n = nf();
  adHTdeltaF.assign((size_t)n * n, Mat<float, 1, 8>());
  if (adHostF.size() != (size_t)n * n) setAdjointsF();
  for (int h = 0; h < n; h++)
    for (int t = 0; t < n; t++) {
      int idx = h + t * n;
      Vec10 dh = frames[h].get_state_minus_stateZero();
      Vec10 dt = frames[t].get_state_minus_stateZero();
      Mat<float, 1, 8> r;
      for (int c = 0; c < 8; c++) {
        float sacc = 0;
        for (int k = 0; k < 8; k++) sacc += (float)dh[k] * adHostF[idx](k, c);
        float tacc = 0;
        for (int k = 0; k < 8; k++) tacc += (float)dt[k] * adTargetF[idx](k, c);
        r(0, c) = sacc + tacc;
      }
      adHTdeltaF[idx] = r;
    }
  for (int i = 0; i < 4; i++) cDeltaF[i] = (float)calib.value_minus_value_zero[i];
  for (Frame& f : frames) {
    Vec10 d = f.get_state_minus_stateZero();
    for (int i = 0; i < 8; i++) { f.delta[i] = d[i]; f.delta_prior[i] = f.state[i]; }  // getPriorZero() == 0
  }
  for (Point& p : points) p.deltaF = p.idepth - p.idepth_zero;
}

// ------------------------------------------------------------------------------------------------
// residual linearisation  (Residuals.cpp:L78-274)
// ------------------------------------------------------------------------------------------------
template <class T>
static inline bool projectPointFull(const Window& W, T u_pt, T v_pt, T idepth, const Mat33f& R, const Vec3f& t,
                                    T& drescale, T& u, T& v, T& Ku, T& Kv, T KliP[3], T& new_idepth) {
  // ResidualProjections.h:L62-87 (dx = dy = 0)
  const Calib& C = W.calib;
  KliP[0] = (u_pt - (T)C.cxl()) * (T)C.fxli();
  KliP[1] = (v_pt - (T)C.cyl()) * (T)C.fyli();
  KliP[2] = 1;
  T ptp[3];
  for (int i = 0; i < 3; i++) ptp[i] = (T)R(i, 0) * KliP[0] + (T)R(i, 1) * KliP[1] + (T)R(i, 2) * KliP[2] + (T)t[i] * idepth;
  drescale = T(1) / ptp[2];
  new_idepth = idepth * drescale;
  if (!(drescale > 0)) return false;
  u = ptp[0] * drescale;
  v = ptp[1] * drescale;
  Ku = u * (T)C.fxl() + (T)C.cxl();
  Kv = v * (T)C.fyl() + (T)C.cyl();
  return Ku > (T)1.1f && Kv > (T)1.1f && Ku < (T)W.wM3G() && Kv < (T)W.hM3G();
}

template <class T>
static inline bool projectPointSimple(const Window& W, T u_pt, T v_pt, T idepth, const Mat33f& KRKi, const Vec3f& Kt, T& Ku, T& Kv) {
  // ResidualProjections.h:L47-57
  T ptp[3];
  for (int i = 0; i < 3; i++) ptp[i] = (T)KRKi(i, 0) * u_pt + (T)KRKi(i, 1) * v_pt + (T)KRKi(i, 2) * T(1) + (T)Kt[i] * idepth;
  Ku = ptp[0] / ptp[2];
  Kv = ptp[1] / ptp[2];
  return Ku > (T)1.1f && Kv > (T)1.1f && Ku < (T)W.wM3G() && Kv < (T)W.hM3G();
}

template <class T>
static inline void interp33(const float* mat, T x, T y, int width, T out[3]) {
  // util/globalFuncs.h:L103-118 (getInterpolatedElement33)
  int ix = (int)x;
  int iy = (int)y;
  T dx = x - ix;
  T dy = y - iy;
  T dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  for (int c = 0; c < 3; c++)
    out[c] = dxdy * (T)bp[3 * (1 + width) + c] + (dy - dxdy) * (T)bp[3 * width + c] + (dx - dxdy) * (T)bp[3 + c] +
             (1 - dx - dy + dxdy) * (T)bp[c];
}

template <class T>
double linearizeOne(const Window& W, Residual& r, RawJ* J) {
  const Settings& S = W.s;
  const Calib& HCalib = W.calib;
  r.state_NewEnergyWithOutlier = -1;
  if (r.state_state == RS_OOB) { r.state_NewState = RS_OOB; return r.state_energy; }

  const Point& point = W.points[r.point];
  const Frame& host = W.frames[r.host];
  const Frame& target = W.frames[r.target];
  const FramePrecalc& precalc = W.precalc[(size_t)r.host * W.nf() + r.target];
  T energyLeft = 0;
  const float* dIl = target.dI;
  const float* color = point.color;
  const float* weights = point.weights;
  T affLL[2] = {(T)precalc.PRE_aff_mode[0], (T)precalc.PRE_aff_mode[1]};
  T b0 = (T)precalc.PRE_b0_mode;

  T d_xi_x[6], d_xi_y[6], d_C_x[4], d_C_y[4], d_d_x, d_d_y;
  {
    T drescale, u, v, new_idepth, Ku, Kv, KliP[3];
    if (!projectPointFull<T>(W, (T)point.u, (T)point.v, (T)point.idepth_zero, precalc.PRE_RTll_0, precalc.PRE_tTll_0, drescale, u, v,
                             Ku, Kv, KliP, new_idepth)) {
      r.state_NewState = RS_OOB;
      return r.state_energy;
    }
    r.centerProjectedTo[0] = (float)Ku; r.centerProjectedTo[1] = (float)Kv; r.centerProjectedTo[2] = (float)new_idepth;
    const Mat33f& R0 = precalc.PRE_RTll_0;
    const Vec3f& t0 = precalc.PRE_tTll_0;
    const T fx = (T)HCalib.fxl(), fy = (T)HCalib.fyl(), fxi = (T)HCalib.fxli(), fyi = (T)HCalib.fyli();
    d_d_x = drescale * ((T)t0[0] - (T)t0[2] * u) * (T)SCALE_IDEPTH * fx;
    d_d_y = drescale * ((T)t0[1] - (T)t0[2] * v) * (T)SCALE_IDEPTH * fy;

    d_C_x[2] = drescale * ((T)R0(2, 0) * u - (T)R0(0, 0));
    d_C_x[3] = fx * drescale * ((T)R0(2, 1) * u - (T)R0(0, 1)) * fyi;
    d_C_x[0] = KliP[0] * d_C_x[2];
    d_C_x[1] = KliP[1] * d_C_x[3];

    d_C_y[2] = fy * drescale * ((T)R0(2, 0) * v - (T)R0(1, 0)) * fxi;
    d_C_y[3] = drescale * ((T)R0(2, 1) * v - (T)R0(1, 1));
    d_C_y[0] = KliP[0] * d_C_y[2];
    d_C_y[1] = KliP[1] * d_C_y[3];

    d_C_x[0] = (d_C_x[0] + u) * (T)SCALE_F;
    d_C_x[1] *= (T)SCALE_F;
    d_C_x[2] = (d_C_x[2] + 1) * (T)SCALE_C;
    d_C_x[3] *= (T)SCALE_C;

    d_C_y[0] *= (T)SCALE_F;
    d_C_y[1] = (d_C_y[1] + v) * (T)SCALE_F;
    d_C_y[2] *= (T)SCALE_C;
    d_C_y[3] = (d_C_y[3] + 1) * (T)SCALE_C;

    d_xi_x[0] = new_idepth * fx;
    d_xi_x[1] = 0;
    d_xi_x[2] = -new_idepth * u * fx;
    d_xi_x[3] = -u * v * fx;
    d_xi_x[4] = (1 + u * u) * fx;
    d_xi_x[5] = -v * fx;

    d_xi_y[0] = 0;
    d_xi_y[1] = new_idepth * fy;
    d_xi_y[2] = -new_idepth * v * fy;
    d_xi_y[3] = -(1 + v * v) * fy;
    d_xi_y[4] = u * v * fy;
    d_xi_y[5] = u * fy;
  }
  for (int i = 0; i < 6; i++) { J->Jpdxi[0][i] = (float)d_xi_x[i]; J->Jpdxi[1][i] = (float)d_xi_y[i]; }
  for (int i = 0; i < 4; i++) { J->Jpdc[0][i] = (float)d_C_x[i]; J->Jpdc[1][i] = (float)d_C_y[i]; }
  J->Jpdd[0] = (float)d_d_x;
  J->Jpdd[1] = (float)d_d_y;

  T JIdxJIdx_00 = 0, JIdxJIdx_11 = 0, JIdxJIdx_10 = 0;
  T JabJIdx_00 = 0, JabJIdx_01 = 0, JabJIdx_10 = 0, JabJIdx_11 = 0;
  T JabJab_00 = 0, JabJab_01 = 0, JabJab_11 = 0;
  T wJI2_sum = 0;

  for (int idx = 0; idx < PATTERN_NUM; idx++) {
    T Ku, Kv;
    if (!projectPointSimple<T>(W, (T)(point.u + patternP[idx][0]), (T)(point.v + patternP[idx][1]), (T)point.idepth,
                               precalc.PRE_KRKiTll, precalc.PRE_KtTll, Ku, Kv)) {
      r.state_NewState = RS_OOB;
      return r.state_energy;
    }
    r.projectedTo[idx][0] = (float)Ku;
    r.projectedTo[idx][1] = (float)Kv;

    T hitColor[3];
    interp33<T>(dIl, Ku, Kv, W.w, hitColor);
    T residual = hitColor[0] - (T)(affLL[0] * (T)color[idx] + affLL[1]);
    T drdA = ((T)color[idx] - b0);
    if (!std::isfinite((float)hitColor[0])) { r.state_NewState = RS_OOB; return r.state_energy; }

    T w = std::sqrt((T)S.outlierTHSumComponent / ((T)S.outlierTHSumComponent + (hitColor[1] * hitColor[1] + hitColor[2] * hitColor[2])));
    w = T(0.5) * (w + (T)weights[idx]);

    T hw = std::fabs(residual) < (T)S.huberTH ? T(1) : (T)S.huberTH / std::fabs(residual);
    energyLeft += w * w * hw * residual * residual * (2 - hw);
    {
      if (hw < 1) hw = std::sqrt(hw);
      hw = hw * w;
      hitColor[1] *= hw;
      hitColor[2] *= hw;
      J->resF[idx] = (float)(residual * hw);
      J->JIdx[0][idx] = (float)hitColor[1];
      J->JIdx[1][idx] = (float)hitColor[2];
      J->JabF[0][idx] = (float)(drdA * hw);
      J->JabF[1][idx] = (float)hw;

      JIdxJIdx_00 += hitColor[1] * hitColor[1];
      JIdxJIdx_11 += hitColor[2] * hitColor[2];
      JIdxJIdx_10 += hitColor[1] * hitColor[2];

      JabJIdx_00 += drdA * hw * hitColor[1];
      JabJIdx_01 += drdA * hw * hitColor[2];
      JabJIdx_10 += hw * hitColor[1];
      JabJIdx_11 += hw * hitColor[2];

      JabJab_00 += drdA * drdA * hw * hw;
      JabJab_01 += drdA * hw * hw;
      JabJab_11 += hw * hw;

      wJI2_sum += hw * hw * (hitColor[1] * hitColor[1] + hitColor[2] * hitColor[2]);

      if (S.affineOptModeA < 0) J->JabF[0][idx] = 0;
      if (S.affineOptModeB < 0) J->JabF[1][idx] = 0;
    }
  }
  J->JIdx2[0][0] = (float)JIdxJIdx_00; J->JIdx2[0][1] = (float)JIdxJIdx_10; J->JIdx2[1][0] = (float)JIdxJIdx_10; J->JIdx2[1][1] = (float)JIdxJIdx_11;
  J->JabJIdx[0][0] = (float)JabJIdx_00; J->JabJIdx[0][1] = (float)JabJIdx_01; J->JabJIdx[1][0] = (float)JabJIdx_10; J->JabJIdx[1][1] = (float)JabJIdx_11;
  J->Jab2[0][0] = (float)JabJab_00; J->Jab2[0][1] = (float)JabJab_01; J->Jab2[1][0] = (float)JabJab_01; J->Jab2[1][1] = (float)JabJab_11;

  r.state_NewEnergyWithOutlier = energyLeft;
  const float TH = std::max<float>(host.frameEnergyTH, target.frameEnergyTH);
  if (energyLeft > (T)TH || wJI2_sum < 2) {
    energyLeft = TH;
    r.state_NewState = RS_OUTLIER;
  } else {
    r.state_NewState = RS_IN;
  }
  r.state_NewEnergy = energyLeft;
  return energyLeft;
}
template double linearizeOne<float>(const Window&, Residual&, RawJ*);
template double linearizeOne<double>(const Window&, Residual&, RawJ*);

static void takeDataF(Residual& r) {  // EnergyFunctionalStructs.cpp:L39-49
  std::swap(r.Jef, r.Jnew);
  const RawJ& J = r.Jef;
  float JI_JI_Jd[2] = {J.JIdx2[0][0] * J.Jpdd[0] + J.JIdx2[0][1] * J.Jpdd[1], J.JIdx2[1][0] * J.Jpdd[0] + J.JIdx2[1][1] * J.Jpdd[1]};
  for (int i = 0; i < 6; i++) r.JpJdF[i] = J.Jpdxi[0][i] * JI_JI_Jd[0] + J.Jpdxi[1][i] * JI_JI_Jd[1];
  r.JpJdF[6] = J.JabJIdx[0][0] * J.Jpdd[0] + J.JabJIdx[0][1] * J.Jpdd[1];
  r.JpJdF[7] = J.JabJIdx[1][0] * J.Jpdd[0] + J.JabJIdx[1][1] * J.Jpdd[1];
}

void applyRes(Residual& r) {  // Residuals.cpp:L306-328 (copyJacobians = true)
  if (r.state_state == RS_OOB) return;  // can never go back from OOB
  if (r.state_NewState == RS_IN) {
    r.isActiveAndIsGoodNEW = true;
    takeDataF(r);
  } else {
    r.isActiveAndIsGoodNEW = false;
  }
  r.state_state = r.state_NewState;
  r.state_energy = r.state_NewEnergy;
}

void Window::setNewFrameEnergyTH() {  // FullSystemOptimize.cpp:L96-149 (no IMU cap)
  std::vector<float> allResVec;
  allResVec.reserve(residuals.size());
  const int newest = nf() - 1;
  for (const Residual& r : residuals)
    if (!r.isLinearized && !r.dropped && r.state_NewEnergyWithOutlier >= 0 && r.target == newest) allResVec.push_back((float)r.state_NewEnergyWithOutlier);
  Frame& newFrame = frames.back();
  if (allResVec.empty()) { newFrame.frameEnergyTH = 12 * 12 * PATTERN_NUM; return; }
  int nthIdx = (int)(s.frameEnergyTHN * allResVec.size());
  std::nth_element(allResVec.begin(), allResVec.begin() + nthIdx, allResVec.end());
  float nthElement = sqrtf(allResVec[nthIdx]);
  newFrame.frameEnergyTH = nthElement * s.frameEnergyTHFacMedian;
  newFrame.frameEnergyTH = 26.0f * s.frameEnergyTHConstWeight + newFrame.frameEnergyTH * (1 - s.frameEnergyTHConstWeight);
  newFrame.frameEnergyTH = newFrame.frameEnergyTH * newFrame.frameEnergyTH;
  newFrame.frameEnergyTH *= s.overallEnergyTHWeight * s.overallEnergyTHWeight;
}

double Window::linearizeAll(bool fixLinearization, std::vector<int>* toRemove, bool updateEnergyTH) {
  // FullSystemOptimize.cpp:L55-88 (linearizeAll_Reductor) + L150-218
  const int nres = (int)residuals.size();
  auto body = [&](int lo, int hi, double* stats, int /*tid*/, std::vector<int>* rem) {
    for (int k = lo; k < hi; k++) {
      Residual& r = residuals[k];
      if (r.isLinearized || r.dropped) continue;  // activeResiduals = all !isLinearized (L431-448)
      stats[0] += linearizeOne<float>(*this, r, &r.Jnew);
      if (fixLinearization) {
        applyRes(r);
        if (r.isActive()) {
          if (r.isNew) {
            Point& p = points[r.point];
            const FramePrecalc& pc = precalc[(size_t)r.host * nf() + r.target];
            float inf3[3], ptp[3];
            for (int i = 0; i < 3; i++) inf3[i] = pc.PRE_KRKiTll(i, 0) * p.u + pc.PRE_KRKiTll(i, 1) * p.v + pc.PRE_KRKiTll(i, 2);
            for (int i = 0; i < 3; i++) ptp[i] = inf3[i] + pc.PRE_KtTll[i] * p.idepth;
            float dx = inf3[0] / inf3[2] - ptp[0] / ptp[2];
            float dy = inf3[1] / inf3[2] - ptp[1] / ptp[2];
            float relBS = 0.01 * std::sqrt(dx * dx + dy * dy);
            if (relBS > p.maxRelBaseline) p.maxRelBaseline = relBS;
            p.numGoodResiduals++;
          }
        } else if (rem) {
          rem->push_back(k);
        }
      }
    }
  };
  double E = 0;
  if (nthreads > 1 && pool && !fixLinearization) {
    pool->reduce([&](int lo, int hi, double* st, int tid) { body(lo, hi, st, tid, nullptr); }, 0, nres, 0);
    E = pool->stats[0];
  } else {
    double st[10] = {0};
    body(0, nres, st, 0, toRemove);
    E = st[0];
  }
  if (updateEnergyTH) setNewFrameEnergyTH();
  return E;
}

void Window::applyResAll() {
  for (Residual& r : residuals)
    if (!r.isLinearized && !r.dropped) applyRes(r);
}

// ------------------------------------------------------------------------------------------------
// accumulators (OptimizationBackend/MatrixAccumulators.h)
// ------------------------------------------------------------------------------------------------
template <class T, int N>
struct Tiered {  // the 1 / 1k / 1M tiering shared by AccumulatorApprox (L595-972), AccumulatorXX (L36-89), AccumulatorX (L177-237)
  T A[N], A1k[N], A1m[N];
  float numIn1 = 0, numIn1k = 0, numIn1m = 0;
  size_t num = 0;
  void initialize() {
    for (int i = 0; i < N; i++) A[i] = A1k[i] = A1m[i] = 0;
    numIn1 = numIn1k = numIn1m = 0;
    num = 0;
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < N; i++) { A1k[i] += A[i]; A[i] = 0; }
      numIn1k += numIn1;
      numIn1 = 0;
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < N; i++) { A1m[i] += A1k[i]; A1k[i] = 0; }
      numIn1m += numIn1k;
      numIn1k = 0;
    }
  }
  void finish() { shiftUp(true); num = (size_t)(numIn1 + numIn1k + numIn1m); }
};

template <class T>
struct AccApprox {  // MatrixAccumulators.h:L595-972 — 55 (10x10 upper) + 30 (10x3) + 6 (3x3 upper)
  Tiered<T, 91> t;
  void initialize() { t.initialize(); }
  // update(): L754-847
  void update(const float* x4, const float* x6, const float* y4, const float* y6, T a, T b, T c) {
    T x[10], y[10];
    for (int i = 0; i < 4; i++) { x[i] = x4[i]; y[i] = y4[i]; }
    for (int i = 0; i < 6; i++) { x[4 + i] = x6[i]; y[4 + i] = y6[i]; }
    int idx = 0;
    for (int r = 0; r < 10; r++)
      for (int cidx = r; cidx < 10; cidx++) {
        t.A[idx] += a * x[cidx] * x[r] + c * y[cidx] * y[r] + b * (x[cidx] * y[r] + y[cidx] * x[r]);
        idx++;
      }
    t.num++;
    t.numIn1++;
    t.shiftUp(false);
  }
  // updateTopRight(): L850-899
  void updateTopRight(const float* x4, const float* x6, const float* y4, const float* y6, T TR00, T TR10, T TR01, T TR11, T TR02, T TR12) {
    T x[10], y[10];
    for (int i = 0; i < 4; i++) { x[i] = x4[i]; y[i] = y4[i]; }
    for (int i = 0; i < 6; i++) { x[4 + i] = x6[i]; y[4 + i] = y6[i]; }
    for (int r = 0; r < 10; r++) {
      t.A[55 + 3 * r + 0] += x[r] * TR00 + y[r] * TR10;
      t.A[55 + 3 * r + 1] += x[r] * TR01 + y[r] * TR11;
      t.A[55 + 3 * r + 2] += x[r] * TR02 + y[r] * TR12;
    }
  }
  // updateBotRight(): L901-915
  void updateBotRight(T a00, T a01, T a02, T a11, T a12, T a22) {
    t.A[85] += a00; t.A[86] += a01; t.A[87] += a02; t.A[88] += a11; t.A[89] += a12; t.A[90] += a22;
  }
  // finish(): L619-651 -> 13x13 symmetric
  void finish(Mat<double, 13, 13>& H) {
    t.shiftUp(true);
    int idx = 0;
    for (int r = 0; r < 10; r++)
      for (int c = r; c < 10; c++) { H(r, c) = H(c, r) = (double)t.A1m[idx]; idx++; }
    idx = 55;
    for (int r = 0; r < 10; r++)
      for (int c = 0; c < 3; c++) { H(r, c + 10) = H(c + 10, r) = (double)t.A1m[idx]; idx++; }
    H(10, 10) = t.A1m[85];
    H(10, 11) = H(11, 10) = t.A1m[86];
    H(10, 12) = H(12, 10) = t.A1m[87];
    H(11, 11) = t.A1m[88];
    H(11, 12) = H(12, 11) = t.A1m[89];
    H(12, 12) = t.A1m[90];
    t.num = (size_t)(t.numIn1 + t.numIn1k + t.numIn1m);
  }
};

template <class T>
struct AccSet {  // one worker's accumulators: AccumulatedTopHessian.h:L146, AccumulatedSCHessian.h:L136-140
  int nf = 0;
  std::vector<AccApprox<T>> top;          // [h + t*nf]
  std::vector<Tiered<T, 64>> accD;        // [h + t1*nf + t2*nf*nf]
  std::vector<Tiered<T, 32>> accE;        // [h + t*nf]  (8x4)
  std::vector<Tiered<T, 8>> accEB;        // [h + t*nf]
  Tiered<T, 16> accHcc;
  Tiered<T, 4> accbc;
  int nres = 0;
  void setZero(int n) {
    nf = n;
    top.resize((size_t)n * n); accD.resize((size_t)n * n * n); accE.resize((size_t)n * n); accEB.resize((size_t)n * n);
    for (auto& a : top) a.initialize();
    for (auto& a : accD) a.initialize();
    for (auto& a : accE) a.initialize();
    for (auto& a : accEB) a.initialize();
    accHcc.initialize();
    accbc.initialize();
    nres = 0;
  }
};

// AccumulatedTopHessian.cpp:L39-159  addPoint<mode>  (mode 0 = active, 1 = linearized, 2 = marginalize)
template <class T>
static void topAddPoint(Window& W, AccSet<T>& acc, Point& p, int mode) {
  const int nf = W.nf();
  float dd = p.deltaF;
  T bd_acc = 0, Hdd_acc = 0, Hcd_acc[4] = {0, 0, 0, 0};
  for (int ri : p.residuals) {
    Residual& r = W.residuals[ri];
    if (mode == 0) { if (r.isLinearized || !r.isActive()) continue; }
    if (mode == 1) { if (!r.isLinearized || !r.isActive()) continue; }
    if (mode == 2) { if (!r.isActive()) continue; }
    const RawJ* rJ = &r.Jef;
    int htIDX = r.host + r.target * nf;
    const Mat<float, 1, 8>& dp = W.adHTdeltaF[htIDX];
    float resApprox[8];
    if (mode == 0) for (int i = 0; i < 8; i++) resApprox[i] = rJ->resF[i];
    if (mode == 2) for (int i = 0; i < 8; i++) resApprox[i] = r.res_toZeroF[i];
    if (mode == 1) {
      float Jp_delta_x = 0, Jp_delta_y = 0;
      for (int i = 0; i < 6; i++) { Jp_delta_x += rJ->Jpdxi[0][i] * dp[i]; Jp_delta_y += rJ->Jpdxi[1][i] * dp[i]; }
      float cx = 0, cy = 0;
      for (int i = 0; i < 4; i++) { cx += rJ->Jpdc[0][i] * W.cDeltaF[i]; cy += rJ->Jpdc[1][i] * W.cDeltaF[i]; }
      Jp_delta_x = Jp_delta_x + cx + rJ->Jpdd[0] * dd;
      Jp_delta_y = Jp_delta_y + cy + rJ->Jpdd[1] * dd;
      float delta_a = dp[6], delta_b = dp[7];
      for (int i = 0; i < 8; i++) {
        float rtz = r.res_toZeroF[i];
        rtz += rJ->JIdx[0][i] * Jp_delta_x;
        rtz += rJ->JIdx[1][i] * Jp_delta_y;
        rtz += rJ->JabF[0][i] * delta_a;
        rtz += rJ->JabF[1][i] * delta_b;
        resApprox[i] = rtz;
      }
    }
    T JI_r[2] = {0, 0}, Jab_r[2] = {0, 0}, rr = 0;
    for (int i = 0; i < 8; i++) {
      JI_r[0] += (T)resApprox[i] * (T)rJ->JIdx[0][i];
      JI_r[1] += (T)resApprox[i] * (T)rJ->JIdx[1][i];
      Jab_r[0] += (T)resApprox[i] * (T)rJ->JabF[0][i];
      Jab_r[1] += (T)resApprox[i] * (T)rJ->JabF[1][i];
      rr += (T)resApprox[i] * (T)resApprox[i];
    }
    AccApprox<T>& A = acc.top[htIDX];
    A.update(rJ->Jpdc[0], rJ->Jpdxi[0], rJ->Jpdc[1], rJ->Jpdxi[1], (T)rJ->JIdx2[0][0], (T)rJ->JIdx2[0][1], (T)rJ->JIdx2[1][1]);
    A.updateBotRight((T)rJ->Jab2[0][0], (T)rJ->Jab2[0][1], Jab_r[0], (T)rJ->Jab2[1][1], Jab_r[1], rr);
    A.updateTopRight(rJ->Jpdc[0], rJ->Jpdxi[0], rJ->Jpdc[1], rJ->Jpdxi[1], (T)rJ->JabJIdx[0][0], (T)rJ->JabJIdx[0][1],
                     (T)rJ->JabJIdx[1][0], (T)rJ->JabJIdx[1][1], JI_r[0], JI_r[1]);
    T Ji2_Jpdd[2] = {(T)rJ->JIdx2[0][0] * (T)rJ->Jpdd[0] + (T)rJ->JIdx2[0][1] * (T)rJ->Jpdd[1],
                     (T)rJ->JIdx2[1][0] * (T)rJ->Jpdd[0] + (T)rJ->JIdx2[1][1] * (T)rJ->Jpdd[1]};
    bd_acc += JI_r[0] * (T)rJ->Jpdd[0] + JI_r[1] * (T)rJ->Jpdd[1];
    Hdd_acc += Ji2_Jpdd[0] * (T)rJ->Jpdd[0] + Ji2_Jpdd[1] * (T)rJ->Jpdd[1];
    for (int i = 0; i < 4; i++) Hcd_acc[i] += (T)rJ->Jpdc[0][i] * Ji2_Jpdd[0] + (T)rJ->Jpdc[1][i] * Ji2_Jpdd[1];
    acc.nres++;
  }
  if (mode == 0) {
    p.Hdd_accAF = (float)Hdd_acc; p.bd_accAF = (float)bd_acc;
    for (int i = 0; i < 4; i++) p.Hcd_accAF[i] = (float)Hcd_acc[i];
  }
  if (mode == 1 || mode == 2) {
    p.Hdd_accLF = (float)Hdd_acc; p.bd_accLF = (float)bd_acc;
    for (int i = 0; i < 4; i++) p.Hcd_accLF[i] = (float)Hcd_acc[i];
  }
  if (mode == 2) {
    for (int i = 0; i < 4; i++) p.Hcd_accAF[i] = 0;
    p.Hdd_accAF = 0;
    p.bd_accAF = 0;
  }
}

// AccumulatedSCHessian.cpp:L34-77
template <class T>
static void scAddPoint(Window& W, AccSet<T>& acc, Point& p, bool shiftPriorToZero) {
  const int nf = W.nf();
  int ngoodres = 0;
  for (int ri : p.residuals) if (W.residuals[ri].isActive()) ngoodres++;
  if (ngoodres == 0) {
    p.HdiF = 0; p.bdSumF = 0; p.idepth_hessian = 0; p.maxRelBaseline = 0;
    return;
  }
  float H = p.Hdd_accAF + p.Hdd_accLF + p.priorF;
  if (H < 1e-10) H = 1e-10;
  p.idepth_hessian = H;
  p.HdiF = 1.0 / H;
  p.bdSumF = p.bd_accAF + p.bd_accLF;
  if (shiftPriorToZero) p.bdSumF += p.priorF * p.deltaF;
  float Hcd[4];
  for (int i = 0; i < 4; i++) Hcd[i] = p.Hcd_accAF[i] + p.Hcd_accLF[i];
  // accHcc.update(Hcd,Hcd,HdiF) : A += w*L*R^T ;  accbc.update(Hcd, bdSumF*HdiF)
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) acc.accHcc.A[i * 4 + j] += (T)(p.HdiF * Hcd[i]) * (T)Hcd[j];
  acc.accHcc.numIn1++; acc.accHcc.shiftUp(false);
  for (int i = 0; i < 4; i++) acc.accbc.A[i] += (T)(p.bdSumF * p.HdiF) * (T)Hcd[i];
  acc.accbc.numIn1++; acc.accbc.shiftUp(false);

  const int nFrames2 = nf * nf;
  for (int r1i : p.residuals) {
    Residual& r1 = W.residuals[r1i];
    if (!r1.isActive()) continue;
    int r1ht = r1.host + r1.target * nf;
    for (int r2i : p.residuals) {
      Residual& r2 = W.residuals[r2i];
      if (!r2.isActive()) continue;
      Tiered<T, 64>& D = acc.accD[r1ht + r2.target * nFrames2];
      for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) D.A[i * 8 + j] += (T)(p.HdiF * r1.JpJdF[i]) * (T)r2.JpJdF[j];
      D.numIn1++; D.shiftUp(false);
    }
    Tiered<T, 32>& E = acc.accE[r1ht];
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 4; j++) E.A[i * 4 + j] += (T)(p.HdiF * r1.JpJdF[i]) * (T)Hcd[j];
    E.numIn1++; E.shiftUp(false);
    Tiered<T, 8>& EB = acc.accEB[r1ht];
    for (int i = 0; i < 8; i++) EB.A[i] += (T)(p.HdiF * p.bdSumF) * (T)r1.JpJdF[i];
    EB.numIn1++; EB.shiftUp(false);
  }
}

static inline void addBlock88(MatX& H, int r0, int c0, const Mat88& M) {
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) H(r0 + i, c0 + j) += M(i, j);
}

// AccumulatedTopHessian.cpp:L241-303 (stitchDoubleInternal) + AccumulatedTopHessian.h:L91-139 (stitchDoubleMT)
template <class T>
static void stitchTopRange(Window& W, std::vector<AccSet<T>>& accs, MatX& H, VecX& b, bool usePrior, int kmin, int kmax) {
  // AccumulatedTopHessian.cpp:L241-303 (stitchDoubleInternal): one worker's share [kmin,kmax) of the nf*nf pair blocks
  const int nf = W.nf();
  for (int k = kmin; k < kmax; k++) {
    int h = k % nf, t = k / nf;
    int hIdx = CPARS + h * 8, tIdx = CPARS + t * 8, aidx = h + nf * t;
    Mat<double, 13, 13> accH;
    for (auto& as : accs) {
      Mat<double, 13, 13> Hk;
      as.top[aidx].finish(Hk);
      if (as.top[aidx].t.num == 0) continue;
      accH += Hk;
    }
    Mat88 P; Mat<double, 8, 4> Q; Vec8 pv;
    for (int i = 0; i < 8; i++) {
      for (int j = 0; j < 8; j++) P(i, j) = accH(CPARS + i, CPARS + j);
      for (int j = 0; j < 4; j++) Q(i, j) = accH(CPARS + i, j);
      pv[i] = accH(CPARS + i, CPARS + 8);
    }
    const Mat88& AH = W.adHost[aidx];
    const Mat88& AT = W.adTarget[aidx];
    addBlock88(H, hIdx, hIdx, AH * P * AH.transpose());
    addBlock88(H, tIdx, tIdx, AT * P * AT.transpose());
    addBlock88(H, hIdx, tIdx, AH * P * AT.transpose());
    Mat<double, 8, 4> hq = AH * Q, tq = AT * Q;
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) { H(hIdx + i, j) += hq(i, j); H(tIdx + i, j) += tq(i, j); }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) H(i, j) += accH(i, j);
    Vec8 hb = AH * pv, tb = AT * pv;
    for (int i = 0; i < 8; i++) { b[hIdx + i] += hb[i]; b[tIdx + i] += tb[i]; }
    for (int i = 0; i < 4; i++) b[i] += accH(i, CPARS + 8);
  }
  if (kmin == 0 && usePrior) {  // only do this on one thread (L290-302)
    for (int i = 0; i < 4; i++) { H(i, i) += W.cPrior[i]; b[i] += W.cPrior[i] * (double)W.cDeltaF[i]; }
    for (int h = 0; h < nf; h++)
      for (int i = 0; i < 8; i++) {
        H(CPARS + h * 8 + i, CPARS + h * 8 + i) += W.frames[h].prior[i];
        b[CPARS + h * 8 + i] += W.frames[h].prior[i] * W.frames[h].delta_prior[i];
      }
  }
}

template <class T>
static void stitchTop(Window& W, std::vector<AccSet<T>>& accs, MatX& H, VecX& b, bool usePrior, int* nresOut) {
  // AccumulatedTopHessian.h:L91-139 (stitchDoubleMT): per-thread H/b summed afterwards
  const int nf = W.nf();
  const int N = nf * 8 + CPARS;
  if (W.nthreads > 1 && W.pool) {
    const int nt = W.pool->size();
    std::vector<MatX> Hs(nt, MatX(N, N));
    std::vector<VecX> bs(nt, VecX(N, 0.0));
    W.pool->reduce([&](int lo, int hi, double*, int tid) { if (lo != hi) stitchTopRange<T>(W, accs, Hs[tid], bs[tid], usePrior, lo, hi); }, 0, nf * nf, 0);
    H = Hs[0]; b = bs[0];
    for (int i = 1; i < nt; i++) { for (size_t a = 0; a < H.d.size(); a++) H.d[a] += Hs[i].d[a]; for (int a = 0; a < N; a++) b[a] += bs[i][a]; }
  } else {
    H = MatX(N, N);
    b.assign(N, 0.0);
    stitchTopRange<T>(W, accs, H, b, usePrior, 0, nf * nf);
  }
  // make diagonal by copying over parts (AccumulatedTopHessian.h:L125-138)
  for (int h = 0; h < nf; h++) {
    int hIdx = CPARS + h * 8;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) H(i, hIdx + j) = H(hIdx + j, i);
    for (int t = h + 1; t < nf; t++) {
      int tIdx = CPARS + t * 8;
      for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) H(hIdx + i, tIdx + j) += H(tIdx + j, hIdx + i);
      for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) H(tIdx + i, hIdx + j) = H(hIdx + j, tIdx + i);
    }
  }
  int nres = 0;
  for (auto& as : accs) nres += as.nres;
  if (nresOut) *nresOut = nres;
}

// AccumulatedSCHessian.cpp:L78-157 + AccumulatedSCHessian.h:L93-133
template <class T>
static void stitchSCRange(Window& W, std::vector<AccSet<T>>& accs, MatX& H, VecX& b, int kmin, int kmax) {
  // AccumulatedSCHessian.cpp:L78-157 (stitchDoubleInternal)
  const int nf = W.nf();
  const int nframes2 = nf * nf;
  for (int k0 = kmin; k0 < kmax; k0++) {
    int i = k0 % nf, j = k0 / nf;
    int iIdx = CPARS + i * 8, jIdx = CPARS + j * 8, ijIdx = i + nf * j;
    Mat<double, 8, 4> Hpc; Vec8 bp;
    for (auto& as : accs) {
      as.accE[ijIdx].finish();
      as.accEB[ijIdx].finish();
      for (int a = 0; a < 32; a++) Hpc.d[a] += (double)as.accE[ijIdx].A1m[a];
      for (int a = 0; a < 8; a++) bp[a] += (double)as.accEB[ijIdx].A1m[a];
    }
    Mat<double, 8, 4> hq = W.adHost[ijIdx] * Hpc, tq = W.adTarget[ijIdx] * Hpc;
    for (int a = 0; a < 8; a++) for (int c = 0; c < 4; c++) { H(iIdx + a, c) += hq(a, c); H(jIdx + a, c) += tq(a, c); }
    Vec8 hb = W.adHost[ijIdx] * bp, tb = W.adTarget[ijIdx] * bp;
    for (int a = 0; a < 8; a++) { b[iIdx + a] += hb[a]; b[jIdx + a] += tb[a]; }
    for (int k = 0; k < nf; k++) {
      int kIdx = CPARS + k * 8, ijkIdx = ijIdx + k * nframes2, ikIdx = i + nf * k;
      Mat88 accDM;
      for (auto& as : accs) {
        as.accD[ijkIdx].finish();
        if (as.accD[ijkIdx].num == 0) continue;
        for (int a = 0; a < 64; a++) accDM.d[a] += (double)as.accD[ijkIdx].A1m[a];
      }
      addBlock88(H, iIdx, iIdx, W.adHost[ijIdx] * accDM * W.adHost[ikIdx].transpose());
      addBlock88(H, jIdx, kIdx, W.adTarget[ijIdx] * accDM * W.adTarget[ikIdx].transpose());
      addBlock88(H, jIdx, iIdx, W.adTarget[ijIdx] * accDM * W.adHost[ikIdx].transpose());
      addBlock88(H, iIdx, kIdx, W.adHost[ijIdx] * accDM * W.adTarget[ikIdx].transpose());
    }
  }
  if (kmin == 0) {
    for (auto& as : accs) {
      as.accHcc.finish();
      as.accbc.finish();
      for (int a = 0; a < 4; a++) for (int c = 0; c < 4; c++) H(a, c) += (double)as.accHcc.A1m[a * 4 + c];
      for (int a = 0; a < 4; a++) b[a] += (double)as.accbc.A1m[a];
    }
  }
}

template <class T>
static void stitchSC(Window& W, std::vector<AccSet<T>>& accs, MatX& H, VecX& b) {
  // AccumulatedSCHessian.h:L93-133 (stitchDoubleMT)
  const int nf = W.nf();
  const int N = nf * 8 + CPARS;
  if (W.nthreads > 1 && W.pool) {
    const int nt = W.pool->size();
    std::vector<MatX> Hs(nt, MatX(N, N));
    std::vector<VecX> bs(nt, VecX(N, 0.0));
    W.pool->reduce([&](int lo, int hi, double*, int tid) { if (lo != hi) stitchSCRange<T>(W, accs, Hs[tid], bs[tid], lo, hi); }, 0, nf * nf, 0);
    H = Hs[0]; b = bs[0];
    for (int i = 1; i < nt; i++) { for (size_t a = 0; a < H.d.size(); a++) H.d[a] += Hs[i].d[a]; for (int a = 0; a < N; a++) b[a] += bs[i][a]; }
  } else {
    H = MatX(N, N);
    b.assign(N, 0.0);
    stitchSCRange<T>(W, accs, H, b, 0, nf * nf);
  }
  for (int h = 0; h < nf; h++) {
    int hIdx = CPARS + h * 8;
    for (int a = 0; a < 4; a++) for (int c = 0; c < 8; c++) H(a, hIdx + c) = H(hIdx + c, a);
  }
}

template <class T>
static void accumulateT(Window& W, ReducedSystem& sys) {
  const int nf = W.nf();
  const int npts = (int)W.points.size();
  const bool MT = W.nthreads > 1 && W.pool;
  const int nacc = MT ? W.pool->size() : 1;
  sys.N = nf * 8 + CPARS;
  // --- accumulateAF_MT (EnergyFunctional.cpp:L201-220)
  {
    std::vector<AccSet<T>> accs(nacc);
    if (MT) {
      W.pool->reduce([&](int, int, double*, int tid) { accs[tid].setZero(nf); }, 0, 0, 0);
      W.pool->reduce([&](int lo, int hi, double*, int tid) { for (int i = lo; i < hi; i++) topAddPoint<T>(W, accs[tid], W.points[i], 0); }, 0, npts, 50);
    } else {
      accs[0].setZero(nf);
      for (int i = 0; i < npts; i++) topAddPoint<T>(W, accs[0], W.points[i], 0);
    }
    stitchTop<T>(W, accs, sys.HA, sys.bA, false, &sys.resInA);
  }
  // --- accumulateLF_MT (L223-242): linearized residuals + priors
  {
    std::vector<AccSet<T>> accs(nacc);
    if (MT) {
      W.pool->reduce([&](int, int, double*, int tid) { accs[tid].setZero(nf); }, 0, 0, 0);
      W.pool->reduce([&](int lo, int hi, double*, int tid) { for (int i = lo; i < hi; i++) topAddPoint<T>(W, accs[tid], W.points[i], 1); }, 0, npts, 50);
    } else {
      accs[0].setZero(nf);
      for (int i = 0; i < npts; i++) topAddPoint<T>(W, accs[0], W.points[i], 1);
    }
    stitchTop<T>(W, accs, sys.HL, sys.bL, true, &sys.resInL);
  }
  // --- accumulateSCF_MT (L248-265)
  {
    std::vector<AccSet<T>> accs(nacc);
    if (MT) {
      W.pool->reduce([&](int, int, double*, int tid) { accs[tid].setZero(nf); }, 0, 0, 0);
      W.pool->reduce([&](int lo, int hi, double*, int tid) { for (int i = lo; i < hi; i++) scAddPoint<T>(W, accs[tid], W.points[i], true); }, 0, npts, 50);
    } else {
      accs[0].setZero(nf);
      for (int i = 0; i < npts; i++) scAddPoint<T>(W, accs[0], W.points[i], true);
    }
    stitchSC<T>(W, accs, sys.Hsc, sys.bsc);
  }
}

// EFResidual::fixLinearizationF (EnergyFunctionalStructs.cpp:L88-114): res_toZeroF = resF - [JIdx*Jp | JabF] * delta
static void fixLinearizationF(const Window& W, Residual& r, const Point& p) {
  const RawJ& J = r.Jef;
  const Mat<float, 1, 8>& dp = W.adHTdeltaF[r.host + W.nf() * r.target];
  float d6x = 0, d6y = 0, d4x = 0, d4y = 0;
  for (int i = 0; i < 6; i++) { d6x += J.Jpdxi[0][i] * dp[i]; d6y += J.Jpdxi[1][i] * dp[i]; }
  for (int i = 0; i < 4; i++) { d4x += J.Jpdc[0][i] * W.cDeltaF[i]; d4y += J.Jpdc[1][i] * W.cDeltaF[i]; }
  const float Jp_delta_x = d6x + d4x + J.Jpdd[0] * p.deltaF;
  const float Jp_delta_y = d6y + d4y + J.Jpdd[1] * p.deltaF;
  const float delta_a = dp[6], delta_b = dp[7];
  for (int i = 0; i < PATTERN_NUM; i++) {
    float rtz = J.resF[i];
    rtz = rtz - J.JIdx[0][i] * Jp_delta_x;
    rtz = rtz - J.JIdx[1][i] * Jp_delta_y;
    rtz = rtz - J.JabF[0][i] * delta_a;
    rtz = rtz - J.JabF[1][i] * delta_b;
    r.res_toZeroF[i] = rtz;
  }
  r.isLinearized = true;
}

std::vector<int> Window::fixLinearization(const std::vector<int>& pts) {  // FullSystem.cpp:L826-838
  std::vector<int> ngood(pts.size(), 0);
  for (size_t k = 0; k < pts.size(); k++) {
    Point& p = points[pts[k]];
    for (int ri : p.residuals) {
      Residual& r = residuals[ri];
      r.state_NewEnergy = r.state_energy = 0;  // resetOOB (Residuals.h:L91-98)
      r.state_NewState = RS_OUTLIER;
      r.state_state = RS_IN;
      linearizeOne<float>(*this, r, &r.Jnew);
      r.isLinearized = false;
      applyRes(r);
      if (r.isActive()) { fixLinearizationF(*this, r, p); ngood[k]++; }
    }
  }
  return ngood;
}

template <class T>
static void marginalizeT(Window& W, const std::vector<int>& pts, ReducedSystem& sys) {
  const int nf = W.nf();
  sys.N = nf * 8 + CPARS;
  std::vector<AccSet<T>> top(1), bot(1);
  top[0].setZero(nf);
  bot[0].setZero(nf);
  for (int i : pts) {
    Point& p = W.points[i];
    p.priorF *= W.s.idepthFixPriorMargFac;
    topAddPoint<T>(W, top[0], p, 2);
    scAddPoint<T>(W, bot[0], p, false);
  }
  stitchTop<T>(W, top, sys.HA, sys.bA, false, &sys.resInA);
  stitchSC<T>(W, bot, sys.Hsc, sys.bsc);
}

void Window::marginalizePoints(const std::vector<int>& pts, int precision, ReducedSystem& sys) {  // EnergyFunctional.cpp:L678-742
  if (precision == 0) marginalizeT<float>(*this, pts, sys);
  else marginalizeT<double>(*this, pts, sys);
  const int N = sys.N;
  if ((int)HM.rows != N) { HM = MatX(N, N); bM.assign(N, 0.0); }
  for (int i = 0; i < N; i++) {
    for (int j = 0; j < N; j++) HM(i, j) += (double)s.margWeightFac * (sys.HA(i, j) - sys.Hsc(i, j));
    bM[i] += (double)s.margWeightFac * (sys.bA[i] - sys.bsc[i]);
  }
}

void Window::accumulate(ReducedSystem& sys, int precision) {
  if (precision == 0) accumulateT<float>(*this, sys);
  else accumulateT<double>(*this, sys);
}

// EnergyFunctional.cpp:L267-321
void Window::resubstitute(const VecX& x) {
  const int n = nf();
  std::vector<float> xF(x.size());
  for (size_t i = 0; i < x.size(); i++) xF[i] = (float)x[i];
  for (int i = 0; i < 4; i++) calib.step[i] = -x[i];
  std::vector<Mat<float, 1, 8>> xAd((size_t)n * n);
  float cstep[4] = {xF[0], xF[1], xF[2], xF[3]};
  for (int h = 0; h < n; h++) {
    for (int i = 0; i < 8; i++) frames[h].step[i] = -x[CPARS + 8 * h + i];
    frames[h].step[8] = frames[h].step[9] = 0;
    for (int t = 0; t < n; t++) {
      Mat<float, 1, 8> r;
      for (int c = 0; c < 8; c++) {
        float a = 0, bb = 0;
        for (int k = 0; k < 8; k++) a += xF[CPARS + 8 * h + k] * adHostF[h + n * t](k, c);
        for (int k = 0; k < 8; k++) bb += xF[CPARS + 8 * t + k] * adTargetF[h + n * t](k, c);
        r(0, c) = a + bb;
      }
      xAd[(size_t)n * h + t] = r;
    }
  }
  auto body = [&](int lo, int hi, double*, int) {
    for (int k = lo; k < hi; k++) {
      Point& p = points[k];
      int ngoodres = 0;
      for (int ri : p.residuals) if (residuals[ri].isActive()) ngoodres++;
      if (ngoodres == 0) { p.step = 0; continue; }
      float b = p.bdSumF;
      float dotc = 0;
      for (int i = 0; i < 4; i++) dotc += cstep[i] * (p.Hcd_accAF[i] + p.Hcd_accLF[i]);
      b -= dotc;
      for (int ri : p.residuals) {
        const Residual& r = residuals[ri];
        if (!r.isActive()) continue;
        float d = 0;
        for (int i = 0; i < 8; i++) d += xAd[(size_t)r.host * n + r.target](0, i) * r.JpJdF[i];
        b -= d;
      }
      p.step = -b * p.HdiF;
    }
  };
  if (nthreads > 1 && pool) pool->reduce(body, 0, (int)points.size(), 50);
  else body(0, (int)points.size(), nullptr, 0);
}

std::vector<VecX> Window::getNullspaces() const {
  const int n = nf(), N = 8 * n + CPARS;
  std::vector<VecX> ns(7, VecX(N));
  for (VecX& v : ns) std::fill(v.begin(), v.end(), 0.0);
  for (int f = 0; f < n; f++) {
    const SE3& T = frames[f].worldToCam_evalPT;
    const SE3 Ti = T.inverse();
    for (int i = 0; i < 6; i++) {  // HessianBlocks.cpp:L81-89
      Vec6 eps; for (int k = 0; k < 6; k++) eps[k] = 0; eps[i] = 1e-3;
      Vec6 epsm; for (int k = 0; k < 6; k++) epsm[k] = -eps[k];
      const Vec6 lp = ((T * SE3::exp(eps)) * Ti).log(), lm = ((T * SE3::exp(epsm)) * Ti).log();
      for (int k = 0; k < 6; k++) {
        double v = (lp[k] - lm[k]) / (2e-3);
        v *= (k < 3) ? (1.0 / SCALE_XI_TRANS) : (1.0 / SCALE_XI_ROT);  // FullSystemOptimize.cpp:L725-726
        ns[i][CPARS + 8 * f + k] = v;
      }
    }
    // HessianBlocks.cpp:L94-100: global scale change
    SE3 P = T, M = T;
    for (int k = 0; k < 3; k++) { P.t[k] *= 1.00001; M.t[k] /= 1.00001; }
    const Vec6 lp = (P * Ti).log(), lm = (M * Ti).log();
    for (int k = 0; k < 6; k++) {
      double v = (lp[k] - lm[k]) / (2e-3);
      v *= (k < 3) ? (1.0 / SCALE_XI_TRANS) : (1.0 / SCALE_XI_ROT);
      ns[6][CPARS + 8 * f + k] = v;
    }
  }
  return ns;
}

// symmetric eigen-decomposition by cyclic Jacobi rotations (k x k, k <= 9): A = V diag(w) V^T
static void jacobiEigen(std::vector<double>& A, int k, std::vector<double>& w, std::vector<double>& V) {
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
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int r = 0; r < k; r++) {
          const double arp = A[(size_t)r * k + p], arq = A[(size_t)r * k + q];
          A[(size_t)r * k + p] = c * arp - sn * arq;
          A[(size_t)r * k + q] = sn * arp + c * arq;
        }
        for (int r = 0; r < k; r++) {
          const double apr = A[(size_t)p * k + r], aqr = A[(size_t)q * k + r];
          A[(size_t)p * k + r] = c * apr - sn * aqr;
          A[(size_t)q * k + r] = sn * apr + c * aqr;
        }
        for (int r = 0; r < k; r++) {
          const double vrp = V[(size_t)r * k + p], vrq = V[(size_t)r * k + q];
          V[(size_t)r * k + p] = c * vrp - sn * vrq;
          V[(size_t)r * k + q] = sn * vrp + c * vrq;
        }
      }
  }
  w.resize(k);
  for (int i = 0; i < k; i++) w[i] = A[(size_t)i * k + i];
}

void Window::orthogonalize(VecX& x) const {  // EnergyFunctional.cpp:L784-838 (b only)
  std::vector<VecX> ns = getNullspaces();
  const int k = (int)ns.size(), N = (int)x.size();
  for (VecX& v : ns) {  // L806-807: normalised columns
    double nrm = 0;
    for (double e : v) nrm += e * e;
    nrm = std::sqrt(nrm);
    for (double& e : v) e /= nrm;
  }
  // the singular values of N are the square roots of the eigenvalues of N^T N; U = N V / sigma
  std::vector<double> G((size_t)k * k), w, V;
  for (int a = 0; a < k; a++)
    for (int b = 0; b < k; b++) { double d = 0; for (int i = 0; i < N; i++) d += ns[a][i] * ns[b][i]; G[(size_t)a * k + b] = d; }
  jacobiEigen(G, k, w, V);
  double maxSv = 0;
  for (int a = 0; a < k; a++) maxSv = std::max(maxSv, std::sqrt(std::max(0.0, w[a])));
  VecX proj(N);
  std::fill(proj.begin(), proj.end(), 0.0);
  for (int a = 0; a < k; a++) {
    const double sv = std::sqrt(std::max(0.0, w[a]));
    if (!(sv > s.solverModeDelta * maxSv)) continue;  // L821-822
    VecX u(N);
    for (int i = 0; i < N; i++) { double d = 0; for (int b = 0; b < k; b++) d += ns[b][i] * V[(size_t)b * k + a]; u[i] = d / sv; }
    double ux = 0;
    for (int i = 0; i < N; i++) ux += u[i] * x[i];
    for (int i = 0; i < N; i++) proj[i] += u[i] * ux;
  }
  for (int i = 0; i < N; i++) x[i] -= proj[i];  // L828: b -= N N^+ b
}

// EnergyFunctional.cpp:L841-996 — default solver mode (SOLVER_ORTHOGONALIZE_X_LATER only), no GTSAM (L971-973)
void Window::solveSystem(int iteration, double lambda, int precision, ReducedSystem* sysOut, MatX* HFinalOut, VecX* bFinalOut) {
  ReducedSystem sys;
  accumulate(sys, precision);
  const int N = sys.N;
  if ((int)HM.rows != N) { HM = MatX(N, N); bM.assign(N, 0.0); }
  // getStitchedDeltaF (L1019-1024)
  VecX delta(N);
  for (int i = 0; i < 4; i++) delta[i] = (double)cDeltaF[i];
  for (int h = 0; h < nf(); h++) for (int i = 0; i < 8; i++) delta[CPARS + 8 * h + i] = frames[h].delta[i];
  VecX bM_top(N);
  for (int i = 0; i < N; i++) { double sacc = bM[i]; for (int j = 0; j < N; j++) sacc += HM(i, j) * delta[j]; bM_top[i] = sacc; }
  MatX HFinal(N, N);
  VecX bFinal(N);
  for (int i = 0; i < N; i++) {
    for (int j = 0; j < N; j++) HFinal(i, j) = sys.HL(i, j) + HM(i, j) + sys.HA(i, j);
    bFinal[i] = sys.bL[i] + bM_top[i] + sys.bA[i] - sys.bsc[i];
  }
  for (int i = 0; i < N; i++) HFinal(i, i) *= (1 + lambda);
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) HFinal(i, j) -= sys.Hsc(i, j) * (1.0 / (1 + lambda));  // L916: (1.0f/(1+lambda)) is a double division
  if (sysOut) *sysOut = sys;
  if (HFinalOut) *HFinalOut = HFinal;
  if (bFinalOut) *bFinalOut = bFinal;
  // Jacobi-preconditioned LDLT (L971-973)
  VecX SVecI(N);
  for (int i = 0; i < N; i++) SVecI[i] = 1.0 / std::sqrt(HFinal(i, i) + 10);
  MatX Hs(N, N);
  VecX bs(N);
  for (int i = 0; i < N; i++) { for (int j = 0; j < N; j++) Hs(i, j) = SVecI[i] * HFinal(i, j) * SVecI[j]; bs[i] = SVecI[i] * bFinal[i]; }
  VecX xs;
  ldlt_solve(Hs, bs, xs);
  VecX x(N);
  for (int i = 0; i < N; i++) x[i] = SVecI[i] * xs[i];
  if (iteration >= 2 && s.orthogonalizeXLater) orthogonalize(x);  // L980-984 (SOLVER_ORTHOGONALIZE_X_LATER)
  lastX = x;
  resubstitute(x);
}

double Window::calcLEnergy() {  // EnergyFunctional.cpp:L349-431 (steady state: no linearized residuals -> priors only)
  double E = 0;
  for (const Frame& f : frames) for (int i = 0; i < 8; i++) E += f.delta_prior[i] * f.prior[i] * f.delta_prior[i];
  float ec = 0;
  for (int i = 0; i < 4; i++) ec += cDeltaF[i] * (float)cPrior[i] * cDeltaF[i];
  E += ec;
  float A = 0;
  for (const Point& p : points) {
    for (int ri : p.residuals) {
      const Residual& r = residuals[ri];
      if (!r.isLinearized || !r.isActive()) continue;
      const Mat<float, 1, 8>& dp = adHTdeltaF[r.host + nf() * r.target];
      const RawJ* rJ = &r.Jef;
      float Jp_delta_x_1 = 0, Jp_delta_y_1 = 0;
      for (int i = 0; i < 6; i++) { Jp_delta_x_1 += rJ->Jpdxi[0][i] * dp[i]; Jp_delta_y_1 += rJ->Jpdxi[1][i] * dp[i]; }
      for (int i = 0; i < 4; i++) { Jp_delta_x_1 += rJ->Jpdc[0][i] * cDeltaF[i]; Jp_delta_y_1 += rJ->Jpdc[1][i] * cDeltaF[i]; }
      Jp_delta_x_1 += rJ->Jpdd[0] * p.deltaF;
      Jp_delta_y_1 += rJ->Jpdd[1] * p.deltaF;
      for (int i = 0; i < 8; i++) {
        float Jdelta = rJ->JIdx[0][i] * Jp_delta_x_1 + rJ->JIdx[1][i] * Jp_delta_y_1 + rJ->JabF[0][i] * dp[6] + rJ->JabF[1][i] * dp[7];
        A += Jdelta * (Jdelta + 2 * r.res_toZeroF[i]);
      }
    }
    A += p.deltaF * p.deltaF * p.priorF;
  }
  return E + A;
}

double Window::calcMEnergy() {  // EnergyFunctional.cpp:L324-346 (firstVal)
  const int N = nf() * 8 + CPARS;
  if ((int)HM.rows != N) return 0;
  VecX delta(N);
  for (int i = 0; i < 4; i++) delta[i] = (double)cDeltaF[i];
  for (int h = 0; h < nf(); h++) for (int i = 0; i < 8; i++) delta[CPARS + 8 * h + i] = frames[h].delta[i];
  double v = 0;
  for (int i = 0; i < N; i++) {
    double sacc = 2 * bM[i];
    for (int j = 0; j < N; j++) sacc += HM(i, j) * delta[j];
    v += delta[i] * sacc;
  }
  return v;
}

void Window::backupState() {  // FullSystemOptimize.cpp:L322-370 (no momentum)
  for (int i = 0; i < 4; i++) calib.value_backup[i] = calib.value[i];
  for (Frame& f : frames) f.state_backup = f.state;
  for (Point& p : points) p.idepth_backup = p.idepth;
}

bool Window::doStepFromBackup() {  // FullSystemOptimize.cpp:L224-317 (stepfac = 1, no momentum)
  float sumA = 0, sumB = 0, sumT = 0, sumR = 0, sumID = 0, numID = 0, sumNID = 0;
  double nv[4];
  for (int i = 0; i < 4; i++) nv[i] = calib.value_backup[i] + calib.step[i];
  calib.setValue(nv);
  for (Frame& f : frames) {
    f.setState(f.state_backup + f.step);
    sumA += f.step[6] * f.step[6];
    sumB += f.step[7] * f.step[7];
    sumT += f.step[0] * f.step[0] + f.step[1] * f.step[1] + f.step[2] * f.step[2];
    sumR += f.step[3] * f.step[3] + f.step[4] * f.step[4] + f.step[5] * f.step[5];
  }
  for (Point& p : points) {
    p.idepth = p.idepth_backup + p.step;
    sumID += p.step * p.step;
    sumNID += fabsf(p.idepth_backup);
    numID++;
    p.idepth_zero = p.idepth_backup + p.step;  // setIdepthZero (DM-VIO: L268)
  }
  sumA /= frames.size(); sumB /= frames.size(); sumR /= frames.size(); sumT /= frames.size();
  sumID /= numID; sumNID /= numID;
  setPrecalcValues();
  return sqrtf(sumA) < 0.0005 * s.thOptIterations && sqrtf(sumB) < 0.00005 * s.thOptIterations && sqrtf(sumR) < 0.00005 * s.thOptIterations &&
         sqrtf(sumT) * sumNID < 0.00005 * s.thOptIterations;
}

void Window::loadStateBackup() {  // FullSystemOptimize.cpp:L371-388
  calib.setValue(calib.value_backup);
  for (Frame& f : frames) f.setState(f.state_backup);
  for (Point& p : points) { p.idepth = p.idepth_backup; p.idepth_zero = p.idepth_backup; }
  setPrecalcValues();
}

// 8x8 inverse by LU with partial pivoting (what Eigen's inverse() does for fixed sizes above 4x4)
static void inverse8(const double A[8][8], double Ainv[8][8]) {
  double M[8][16];
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) { M[i][j] = A[i][j]; M[i][8 + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < 8; c++) {
    int piv = c;
    for (int r = c + 1; r < 8; r++) if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
    if (piv != c) for (int j = 0; j < 16; j++) std::swap(M[c][j], M[piv][j]);
    const double d = M[c][c];
    for (int r = c + 1; r < 8; r++) {
      const double f = M[r][c] / d;
      if (f == 0.0) continue;
      for (int j = c; j < 16; j++) M[r][j] -= f * M[c][j];
    }
  }
  for (int c = 7; c >= 0; c--) {  // back substitution on the 8 right-hand sides
    for (int j = 8; j < 16; j++) {
      double v = M[c][j];
      for (int k = c + 1; k < 8; k++) v -= M[c][k] * M[k][j];
      M[c][j] = v / M[c][c];
    }
  }
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) Ainv[i][j] = M[i][8 + j];
}

void Window::marginalizeFrame(int idx) {  // EnergyFunctional.cpp:L522-675
  const int n = nf(), odim = n * 8 + CPARS, ndim = odim - 8;
  for (const Point& p : points) if (p.host == idx && !p.residuals.empty()) { fprintf(stderr, "orc: marginalizeFrame: frame still hosts points\n"); abort(); }
  if ((int)HM.rows != odim) { HM = MatX(odim, odim); bM.assign(odim, 0.0); }
  // L572-592: move the frame's block to the end
  std::vector<int> perm;
  const int io = idx * 8 + CPARS;
  for (int i = 0; i < odim; i++) if (i < io || i >= io + 8) perm.push_back(i);
  for (int k = 0; k < 8; k++) perm.push_back(io + k);
  MatX H(odim, odim);
  VecX b(odim);
  for (int i = 0; i < odim; i++) {
    b[i] = bM[perm[i]];
    for (int j = 0; j < odim; j++) H(i, j) = HM(perm[i], perm[j]);
  }
  // L595-596: the frame's own prior
  const Frame& f = frames[idx];
  for (int k = 0; k < 8; k++) { H(ndim + k, ndim + k) += f.prior[k]; b[ndim + k] += f.prior[k] * f.delta_prior[k]; }
  // L603-612: diagonal scaling
  VecX SVec(odim), SVecI(odim);
  for (int i = 0; i < odim; i++) { SVec[i] = std::sqrt(std::fabs(H(i, i)) + 10.0); SVecI[i] = 1.0 / SVec[i]; }
  for (int i = 0; i < odim; i++) {
    for (int j = 0; j < odim; j++) H(i, j) = (SVecI[i] * H(i, j)) * SVecI[j];
    b[i] = SVecI[i] * b[i];
  }
  // L615-618: hpi = inverse of the bottom-right block (the 0.5f*(hpi+hpi) lines of the reference are the identity)
  double blk[8][8], hpi[8][8];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) blk[i][j] = H(ndim + i, ndim + j);
  inverse8(blk, hpi);
  // L621-623: Schur complement
  MatX bli(ndim, 8);
  for (int i = 0; i < ndim; i++)
    for (int k = 0; k < 8; k++) {
      double v = 0;
      for (int m = 0; m < 8; m++) v += H(ndim + m, i) * hpi[m][k];
      bli(i, k) = v;
    }
  MatX Hn(ndim, ndim);
  VecX bn(ndim);
  for (int i = 0; i < ndim; i++) {
    for (int j = 0; j < ndim; j++) {
      double v = 0;
      for (int k = 0; k < 8; k++) v += bli(i, k) * H(ndim + k, j);
      Hn(i, j) = H(i, j) - v;
    }
    double v = 0;
    for (int k = 0; k < 8; k++) v += bli(i, k) * b[ndim + k];
    bn[i] = b[i] - v;
  }
  // L626-631: unscale, symmetrise
  for (int i = 0; i < ndim; i++) {
    for (int j = 0; j < ndim; j++) Hn(i, j) = (SVec[i] * Hn(i, j)) * SVec[j];
    bn[i] = SVec[i] * bn[i];
  }
  HM = MatX(ndim, ndim);
  bM.assign(ndim, 0.0);
  for (int i = 0; i < ndim; i++) {
    for (int j = 0; j < ndim; j++) HM(i, j) = 0.5 * (Hn(i, j) + Hn(j, i));
    bM[i] = bn[i];
  }
  // L642-649: the frame leaves the window; later frames shift down
  for (Residual& r : residuals) {
    if ((r.target == idx || r.host == idx) && !r.dropped) { fprintf(stderr, "orc: marginalizeFrame: live residual still references the frame\n"); abort(); }
    if (r.target > idx) r.target--;
    if (r.host > idx) r.host--;
  }
  for (Point& p : points) if (p.host > idx) p.host--;
  frames.erase(frames.begin() + idx);
}

double Window::finishOptimize(std::vector<int>* toRemove) {  // FullSystemOptimize.cpp:L591-609
  Frame& newest = frames.back();
  Vec10 newStateZero;
  for (int i = 0; i < 10; i++) newStateZero[i] = 0;
  newStateZero[6] = newest.state[6];
  newStateZero[7] = newest.state[7];
  newest.worldToCam_evalPT = newest.PRE_worldToCam;  // FrameHessian::setEvalPT (HessianBlocks.h:L237-245)
  newest.setState(newStateZero);
  newest.state_zero = newStateZero;
  setAdjointsF();
  setPrecalcValues();
  std::vector<int> rem;
  const double E = linearizeAll(true, &rem, true);
  for (int k : rem) {  // FullSystemOptimize.cpp:L196-214: the residual leaves its point and the energy functional
    Residual& r = residuals[k];
    r.dropped = true;
    r.isActiveAndIsGoodNEW = false;
    std::vector<int>& list = points[r.point].residuals;
    list.erase(std::remove(list.begin(), list.end(), k), list.end());
  }
  if (toRemove) *toRemove = rem;
  return E;
}

int Window::optimize(int mnumOptIts, int precision, std::vector<double>* energyLog) {
  // FullSystemOptimize.cpp:L417-647 without IMU/GTSAM, logging and the final linearizeAll(true) bookkeeping
  if (nf() < 2) return 0;
  if (nf() < 3) mnumOptIts = 20;
  if (nf() < 4) mnumOptIts = 15;
  for (Residual& r : residuals) if (!r.isLinearized && !r.dropped) { r.state_state = RS_IN; r.state_NewState = RS_OUTLIER; }  // resetOOB (Residuals.h:L91-98)
  for (Residual& r : residuals) if (!r.isLinearized && !r.dropped) { r.state_energy = 0; r.state_NewEnergy = 0; }
  double lastEnergy = linearizeAll(false, nullptr);
  double lastEnergyL = calcLEnergy();
  double lastEnergyM = calcMEnergy();
  applyResAll();
  if (energyLog) energyLog->push_back(lastEnergy);
  double lambda = 1e-5;
  const double minLambda = 1e-5;
  int numIterations = 0;
  for (int iteration = 0; iteration < mnumOptIts; iteration++) {
    backupState();
    solveSystem(iteration, lambda, precision);
    bool canbreak = doStepFromBackup();
    double newEnergy = linearizeAll(false, nullptr);
    double newEnergyL = calcLEnergy();
    double newEnergyM = calcMEnergy();
    if (newEnergy + newEnergyL + newEnergyM < lastEnergy + lastEnergyL + lastEnergyM) {
      applyResAll();
      lastEnergy = newEnergy; lastEnergyL = newEnergyL; lastEnergyM = newEnergyM;
      lambda *= 0.25;
      lambda = std::max(lambda, minLambda);
    } else {
      loadStateBackup();
      lastEnergy = linearizeAll(false, nullptr);
      lastEnergyL = calcLEnergy();
      lastEnergyM = calcMEnergy();
      lambda *= 1e2;
    }
    if (energyLog) energyLog->push_back(lastEnergy);
    numIterations++;
    if (canbreak && iteration >= s.minOptIterations) break;
  }
  return numIterations;
}

}  // namespace orc
