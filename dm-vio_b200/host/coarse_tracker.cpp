// Host-side C++ mirror of CoarseTracker on top of the C ABI — see coarse_tracker.h.
#include "coarse_tracker.h"
#include "../csrc/inv3.h"
#include <algorithm>
#include <cmath>

namespace dmvio_b200 {

CoarseTracker::CoarseTracker(int w, int h, int levels, int max_points, int device) : levels_(levels) {
  dmv_ct_config cfg = {w, h, levels, max_points, device};
  if (dmv_ct_create(&cfg, &ct_) != DMV_OK) { err_ = dmv_last_error(); ct_ = nullptr; }
  for (int l = 0; l < DMV_MAX_PYR_LEVELS; l++) { w_[l] = w >> l; h_[l] = h >> l; pc_n[l] = 0; fx_[l] = fy_[l] = 1; cx_[l] = cy_[l] = 0; }
  for (int i = 0; i < 5; i++) lastResiduals[i] = NAN;
  for (int i = 0; i < 3; i++) lastFlowIndicators[i] = 1000;
}
CoarseTracker::~CoarseTracker() { if (ct_) dmv_ct_destroy(ct_); }

void CoarseTracker::makeK(const CalibHessian& HCalib) {  // CoarseTracker.cpp:L105-134
  fx_[0] = HCalib.value_scaledf[0]; fy_[0] = HCalib.value_scaledf[1]; cx_[0] = HCalib.value_scaledf[2]; cy_[0] = HCalib.value_scaledf[3];
  for (int l = 1; l < levels_; l++) {
    fx_[l] = fx_[l - 1] * 0.5;
    fy_[l] = fy_[l - 1] * 0.5;
    cx_[l] = (cx_[0] + 0.5) / ((int)1 << l) - 0.5;
    cy_[l] = (cy_[0] + 0.5) / ((int)1 << l) - 0.5;
  }
  if (ct_) {
    dmv_ct_set_huber(ct_, s.setting_huberTH);
    for (int l = 0; l < levels_; l++) dmv_ct_set_K(ct_, l, fx_[l], fy_[l], cx_[l], cy_[l]);
  }
}

void CoarseTracker::setCoarseTrackingRef(int n, const float* Ku, const float* Kv, const float* new_idepth, const float* HdiF,
                                         const float* const* refdIp, AffLight ref_aff, float ref_exposure) {
  // makeCoarseDepthL0 (CoarseTracker.cpp:L138-295): weighted splat on level 0, 2x2 sum pooling, one dilation step per level
  // (diagonal neighbours on levels 0-1, 4-neighbourhood above), normalisation and compaction into pc_* lists
  lastRef_aff_g2l_ = ref_aff;
  lastRef_ab_exposure_ = ref_exposure;
  std::vector<std::vector<float>> idepth(levels_), wsum(levels_);
  for (int l = 0; l < levels_; l++) { idepth[l].assign((size_t)w_[l] * h_[l], 0.f); wsum[l].assign((size_t)w_[l] * h_[l], 0.f); }
  for (int i = 0; i < n; i++) {
    const int u = Ku[i] + 0.5f, v = Kv[i] + 0.5f;
    const float weight = sqrtf(1e-3 / (HdiF[i] + 1e-12));
    idepth[0][u + w_[0] * v] += new_idepth[i] * weight;
    wsum[0][u + w_[0] * v] += weight;
  }
  for (int l = 1; l < levels_; l++) {
    const int wl = w_[l], hl = h_[l], wm = w_[l - 1];
    for (int y = 0; y < hl; y++)
      for (int x = 0; x < wl; x++) {
        const int b = 2 * x + 2 * y * wm;
        idepth[l][x + y * wl] = idepth[l - 1][b] + idepth[l - 1][b + 1] + idepth[l - 1][b + wm] + idepth[l - 1][b + wm + 1];
        wsum[l][x + y * wl] = wsum[l - 1][b] + wsum[l - 1][b + 1] + wsum[l - 1][b + wm] + wsum[l - 1][b + wm + 1];
      }
  }
  for (int l = 0; l < levels_; l++) {
    const int wl = w_[l], wh = w_[l] * h_[l] - w_[l];
    const int nb[2][4] = {{1 + wl, -1 - wl, wl - 1, -wl + 1}, {1, -1, wl, -wl}};
    const int* off = nb[l < 2 ? 0 : 1];
    const std::vector<float> bak = wsum[l];
    float* idl = idepth[l].data();
    float* ws = wsum[l].data();
    for (int i = wl + 1; i < wh - 1; i++) {
      if (bak[i] > 0) continue;
      float sum = 0, num = 0, numn = 0;
      for (int k = 0; k < 4; k++)
        if (bak[i + off[k]] > 0) { sum += idl[i + off[k]]; num += bak[i + off[k]]; numn++; }
      if (numn > 0) { idl[i] = sum / numn; ws[i] = num / numn; }
    }
  }
  for (int l = 0; l < levels_; l++) {
    const int wl = w_[l], hl = h_[l];
    pc_u[l].clear(); pc_v[l].clear(); pc_idepth[l].clear(); pc_color[l].clear();
    for (int y = 2; y < hl - 2; y++)
      for (int x = 2; x < wl - 2; x++) {
        const int i = x + y * wl;
        if (!(wsum[l][i] > 0)) continue;
        const float id = idepth[l][i] / wsum[l][i];
        const float col = refdIp[l][3 * (size_t)i];
        if (!std::isfinite(col) || !(id > 0)) continue;
        pc_u[l].push_back((float)x); pc_v[l].push_back((float)y); pc_idepth[l].push_back(id); pc_color[l].push_back(col);
      }
    pc_n[l] = (int)pc_u[l].size();
    if (ct_ && dmv_ct_set_ref(ct_, l, pc_n[l], pc_u[l].data(), pc_v[l].data(), pc_idepth[l].data(), pc_color[l].data()) != DMV_OK)
      err_ = dmv_last_error();
  }
}

bool CoarseTracker::setNewFrame(const float* image, float ab_exposure) {
  newFrame_ab_exposure_ = ab_exposure;
  if (!ct_ || dmv_ct_upload_new_image(ct_, image) != DMV_OK) { err_ = dmv_last_error(); return false; }
  return true;
}
bool CoarseTracker::setNewFramePyramid(const float* const* dIp, float ab_exposure) {
  newFrame_ab_exposure_ = ab_exposure;
  for (int l = 0; l < levels_; l++)
    if (!ct_ || dmv_ct_upload_new(ct_, l, dIp[l]) != DMV_OK) { err_ = dmv_last_error(); return false; }
  return true;
}

bool CoarseTracker::setCoarseTrackingRefOnDevice(int n, const float* Ku, const float* Kv, const float* new_idepth, const float* HdiF,
                                                 const float* ref_image_wh, AffLight ref_aff, float ref_exposure) {
  lastRef_aff_g2l_ = ref_aff;
  lastRef_ab_exposure_ = ref_exposure;
  if (!ct_ || dmv_ct_upload_new_image(ct_, ref_image_wh) != DMV_OK) { err_ = dmv_last_error(); return false; }
  int32_t cnt[DMV_MAX_PYR_LEVELS] = {0};
  if (dmv_ct_make_coarse_depth(ct_, n, Ku, Kv, new_idepth, HdiF, cnt) != DMV_OK) { err_ = dmv_last_error(); return false; }
  for (int l = 0; l < levels_; l++) { pc_n[l] = cnt[l]; pc_u[l].clear(); pc_v[l].clear(); pc_idepth[l].clear(); pc_color[l].clear(); }
  return true;
}

bool CoarseTracker::eval(int lvl, const SE3& refToNew, AffLight aff_g2l, float cutoffTH, bool wantGS, double res6[6], double H[64], double b[8]) {
  // operands of calcRes (CoarseTracker.cpp:L377-379): RKi = R * Ki[lvl] in float, t in float, affLL in float
  float R[9], t[3], RKi[9];
  for (int i = 0; i < 9; i++) R[i] = (float)refToNew.R[i];
  for (int i = 0; i < 3; i++) t[i] = (float)refToNew.t[i];
  const float Kl[9] = {fx_[lvl], 0.f, cx_[lvl], 0.f, fy_[lvl], cy_[lvl], 0.f, 0.f, 1.f};
  float Ki[9];
  dmv::inv3_cofactor(Kl, Ki);  // Ki[lvl] = K[lvl].inverse() with the reference's rounding (CoarseTracker.cpp:L128)
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) RKi[i * 3 + j] = R[i * 3] * Ki[j] + R[i * 3 + 1] * Ki[3 + j] + R[i * 3 + 2] * Ki[6 + j];
  double aff[2];
  AffLight::fromToVecExposure(lastRef_ab_exposure_, newFrame_ab_exposure_, lastRef_aff_g2l_, aff_g2l, aff);
  const float affLL[2] = {(float)aff[0], (float)aff[1]};
  int nw = 0;
  evaluations++;
  if (dmv_ct_calc_res_gs(ct_, lvl, RKi, t, affLL, (float)lastRef_aff_g2l_.b, cutoffTH, wantGS ? 1 : 0, res6, H, b, &nw) != DMV_OK) {
    err_ = dmv_last_error();
    return false;
  }
  return true;
}

bool CoarseTracker::trackNewestCoarse(SE3& lastToNew_out, AffLight& aff_g2l_out, int coarsestLvl, const double minResForAbort[5]) {
  // CoarseTracker.cpp:L539-770, visual-only branch.  calcRes and calcGSSSE are one launch: every evaluation returns the
  // residual statistics AND the Gauss-Newton system at that pose; H,b are adopted only when the step is accepted.
  if (useDeviceLM && !computeCoarseUpdate) {  // a host consumer needs H, b on the host every iteration: host loop
    dmv_ct_track_args in;
    for (int i = 0; i < 9; i++) in.R[i] = lastToNew_out.R[i];
    for (int i = 0; i < 3; i++) in.t[i] = lastToNew_out.t[i];
    in.a = aff_g2l_out.a; in.b = aff_g2l_out.b;
    in.ref_a = lastRef_aff_g2l_.a; in.ref_b = lastRef_aff_g2l_.b;
    in.ref_exposure = lastRef_ab_exposure_; in.new_exposure = newFrame_ab_exposure_;
    in.coarseCutoffTH = s.setting_coarseCutoffTH; in.affineOptModeA = s.setting_affineOptModeA; in.affineOptModeB = s.setting_affineOptModeB;
    in.coarsestLvl = coarsestLvl;
    for (int i = 0; i < 5; i++) in.minResForAbort[i] = minResForAbort[i];
    dmv_ct_track_result r;
    if (dmv_ct_track(ct_, &in, &r) != DMV_OK) { err_ = dmv_last_error(); return false; }
    for (int i = 0; i < 5; i++) lastResiduals[i] = r.lastResiduals[i];
    for (int i = 0; i < 3; i++) lastFlowIndicators[i] = r.flowIndicators[i];
    iterations = r.iterations;
    evaluations = r.evaluations;
    dmv_ct_last_point_evaluations(ct_, &pointEvaluations);
    if (r.status != 0) return false;  // aborted inside the level loop: outputs untouched, like the reference's early return
    for (int i = 0; i < 9; i++) lastToNew_out.R[i] = r.R[i];
    for (int i = 0; i < 3; i++) lastToNew_out.t[i] = r.t[i];
    aff_g2l_out.a = r.a; aff_g2l_out.b = r.b;
    return r.trackingGood != 0;
  }
  for (int i = 0; i < 5; i++) lastResiduals[i] = NAN;
  for (int i = 0; i < 3; i++) lastFlowIndicators[i] = 1000;
  const int maxIterations[] = {10, 20, 50, 50, 50};
  const float lambdaExtrapolationLimit = 0.001;
  SE3 refToNew_current = lastToNew_out;
  AffLight aff_g2l_current = aff_g2l_out;
  bool haveRepeated = false;
  iterations = 0;
  evaluations = 0;
  double H[64], b[8], Hn[64], bn[8];
  for (int lvl = coarsestLvl; lvl >= 0; lvl--) {
    float levelCutoffRepeat = 1;
    double resOld[6];
    if (!eval(lvl, refToNew_current, aff_g2l_current, s.setting_coarseCutoffTH * levelCutoffRepeat, true, resOld, H, b)) return false;
    while (resOld[5] > 0.6 && (levelCutoffRepeat < 50 || resOld[5] > 0.99)) {
      levelCutoffRepeat *= 2;
      if (!eval(lvl, refToNew_current, aff_g2l_current, s.setting_coarseCutoffTH * levelCutoffRepeat, true, resOld, H, b)) return false;
    }
    float lambda = 0.01;
    for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
      iterations++;
      double Hl[64];
      for (int i = 0; i < 64; i++) Hl[i] = H[i];
      for (int i = 0; i < 8; i++) Hl[i * 8 + i] *= (1 + lambda);
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = sqrt(sqrt(lambdaExtrapolationLimit / lambda));
      SE3 refToNew_new;
      AffLight aff_g2l_new = aff_g2l_current;
      double incNorm = 0;
      if (computeCoarseUpdate) {  // CoarseTracker.cpp:L616-637: the IMU integration forms the step from the photometric H, b
        double incA = 0, incB = 0;
        refToNew_new = computeCoarseUpdate(Hl, b, extrapFac, lambda, incA, incB, incNorm);
        aff_g2l_new.a += incA;
        aff_g2l_new.b += incB;
      } else {
        double inc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        {
          // Vec8 inc = Hl.ldlt().solve(-b) with the fixed-a / fixed-b variants (L639-665)
          int map[8] = {0, 1, 2, 3, 4, 5, 6, 7};
          int n = 8;
          const bool fixA = s.setting_affineOptModeA < 0, fixB = s.setting_affineOptModeB < 0;
          if (fixA && fixB) n = 6;
          else if (!fixA && fixB) n = 7;
          else if (fixA && !fixB) { n = 7; map[6] = 7; }
          double A[64], rhs[8], x[8];
          for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) A[i * n + j] = Hl[map[i] * 8 + map[j]]; rhs[i] = -b[map[i]]; }
          ldlt_solve(n, A, rhs, x);
          for (int i = 0; i < n; i++) inc[map[i]] = x[i];
        }
        for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
        double incScaled[8];
        for (int i = 0; i < 8; i++) incScaled[i] = inc[i];
        incScaled[6] *= SCALE_A;
        incScaled[7] *= SCALE_B;
        double ssum = 0;
        for (int i = 0; i < 8; i++) ssum += incScaled[i];
        if (!std::isfinite(ssum)) for (int i = 0; i < 8; i++) incScaled[i] = 0;
        refToNew_new = SE3::exp(incScaled) * refToNew_current;
        aff_g2l_new.a += incScaled[6];
        aff_g2l_new.b += incScaled[7];
        for (int i = 0; i < 8; i++) incNorm += inc[i] * inc[i];
        incNorm = std::sqrt(incNorm);
      }
      double resNew[6];
      if (!eval(lvl, refToNew_new, aff_g2l_new, s.setting_coarseCutoffTH * levelCutoffRepeat, true, resNew, Hn, bn)) return false;
      const bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
      if (accept) {
        for (int i = 0; i < 64; i++) H[i] = Hn[i];
        for (int i = 0; i < 8; i++) b[i] = bn[i];
        for (int i = 0; i < 6; i++) resOld[i] = resNew[i];
        aff_g2l_current = aff_g2l_new;
        refToNew_current = refToNew_new;
        if (acceptCoarseUpdate) acceptCoarseUpdate();
        lambda *= 0.5;
      } else {
        lambda *= 4;
        if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
      }
      if (!(incNorm > 1e-3)) break;
    }
    lastResiduals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
    lastFlowIndicators[0] = resOld[2]; lastFlowIndicators[1] = resOld[3]; lastFlowIndicators[2] = resOld[4];
    if (std::isnan(lastResiduals[lvl])) return false;
    if (lastResiduals[lvl] > 1.5 * minResForAbort[lvl]) return false;
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
  }
  lastToNew_out = refToNew_current;
  aff_g2l_out = aff_g2l_current;
  bool trackingGood = true;
  if ((s.setting_affineOptModeA != 0 && (fabsf((float)aff_g2l_out.a) > 1.2)) || (s.setting_affineOptModeB != 0 && (fabsf((float)aff_g2l_out.b) > 200)))
    trackingGood = false;
  double rel[2];
  AffLight::fromToVecExposure(lastRef_ab_exposure_, newFrame_ab_exposure_, lastRef_aff_g2l_, aff_g2l_out, rel);
  if ((s.setting_affineOptModeA == 0 && (fabsf(logf((float)rel[0])) > 1.5)) || (s.setting_affineOptModeB == 0 && (fabsf((float)rel[1]) > 200)))
    trackingGood = false;
  if (s.setting_affineOptModeA < 0) aff_g2l_out.a = 0;
  if (s.setting_affineOptModeB < 0) aff_g2l_out.b = 0;
  return trackingGood;
}

}  // namespace dmvio_b200
