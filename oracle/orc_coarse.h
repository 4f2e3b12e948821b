// TEST INFRASTRUCTURE ONLY — CPU oracle (restatement) of DM-VIO's coarse direct image alignment
// (FullSystem/CoarseTracker.cpp) and of the image pyramid construction (FullSystem/HessianBlocks.cpp
// makeImages).  Pinned bit-exact against the compiled reference (see orc_ba.h, tests/test_ref_pin.py).
#pragma once
#include "orc_ba.h"

namespace orc {

constexpr int PYR_LEVELS = 6;  // util/settings.h:L52

// util/globalCalib.cpp:L45-105 : pyramid level rule + per-level intrinsics
struct GlobalCalib {
  int pyrLevelsUsed = 1;
  int wG[PYR_LEVELS], hG[PYR_LEVELS];
  float fxG[PYR_LEVELS], fyG[PYR_LEVELS], cxG[PYR_LEVELS], cyG[PYR_LEVELS];
  void set(int w, int h, float fx, float fy, float cx, float cy, int forceLevels = 0);
};

// HessianBlocks.cpp:L128-191 (makeImages, without the gamma-weighted absSquaredGrad which only feeds the pixel selector).
// out[lvl] must hold wG[lvl]*hG[lvl]*3 floats.  First/last row gradients are left 0 (the reference leaves them uninitialised).
void makeImages(const GlobalCalib& g, const float* color, float* const* dIp_out, float* const* absSquaredGrad_out);

// ImmaturePoint.cpp:L36-62 : per-point pattern colours and gradient weights from the host frame
bool initPointColorWeights(const float* dI, int w, float u, float v, float outlierTHSumComponent, float* color8, float* weights8);

struct CoarseTracker {
  Settings s;
  int levels = 1;
  int w[PYR_LEVELS], h[PYR_LEVELS];
  float fx[PYR_LEVELS], fy[PYR_LEVELS], cx[PYR_LEVELS], cy[PYR_LEVELS];
  Mat33f Ki[PYR_LEVELS];
  // reference frame point cloud (pc_*), CoarseTracker.cpp:L249-293
  std::vector<float> pc_u[PYR_LEVELS], pc_v[PYR_LEVELS], pc_idepth[PYR_LEVELS], pc_color[PYR_LEVELS];
  int pc_n[PYR_LEVELS];
  const float* newFrame_dIp[PYR_LEVELS];
  float lastRef_ab_exposure = 1, newFrame_ab_exposure = 1;
  AffLight lastRef_aff_g2l;
  // warped buffers
  std::vector<float> buf_warped_idepth, buf_warped_u, buf_warped_v, buf_warped_dx, buf_warped_dy, buf_warped_residual, buf_warped_weight,
      buf_warped_refColor;
  int buf_warped_n = 0;
  double lastResiduals[5];
  double lastFlowIndicators[3];

  void makeK(const GlobalCalib& g);  // CoarseTracker.cpp:L105-134
  // CoarseTracker.cpp:L138-295 — splat / pool / dilate / compact.  Inputs: per-point (Ku, Kv, new_idepth, HdiF) of IN residuals targeting the ref.
  void makeCoarseDepthL0(int n, const float* Ku, const float* Kv, const float* new_idepth, const float* HdiF, const float* const* refdIp);
  void calcRes(int lvl, const SE3& refToNew, AffLight aff_g2l, float cutoffTH, double out6[6]);            // L361-517
  void calcGSSSE(int lvl, Mat88& H_out, Vec8& b_out, const SE3& refToNew, AffLight aff_g2l, int precision);  // L299-356
  // the same two functions behind their operand set-up (RKi, t, affLL / a, b0 given): what the C ABI's dmv_ct_calc_res_gs takes
  void calcResRaw(int lvl, const Mat33f& RKi, const Vec3f& t, const float affLL[2], float cutoffTH, double out6[6]);
  void calcGSRaw(int lvl, Mat88& H_out, Vec8& b_out, float a, float b0, int precision);
  // L539-770 (no IMU branch).  returns trackingGood; iterations per level are logged in itsOut (optional)
  bool trackNewestCoarse(SE3& lastToNew_out, AffLight& aff_g2l_out, int coarsestLvl, const double minResForAbort[5], int precision,
                         int* totalIterations = nullptr);
};

}  // namespace orc
