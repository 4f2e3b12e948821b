// Host-side C++ mirror of the reference's CoarseInitializer surface (src/dso/FullSystem/CoarseInitializer.{h,cpp}) on top of the C ABI:
// trackFrame's pyramid loop and LM control, doStep / applyStep / calcEC / optReg / propagateUp / propagateDown / resetPoints on the host (scalar
// per-point code, once per sequence), calcResAndGS as ONE CUDA launch per evaluation (dmv_ci_calc_res_and_gs).  Pixel selection and the kd-tree
// (setFirst's PixelSelector + makeNN, CoarseInitializer.cpp:L804-889, L1001-1072) are inputs: points, parents and neighbour lists are given.
#pragma once
#include "window_ba.h"

#include <array>

namespace dmvio_b200 {

struct Pnt {  // CoarseInitializer.h:L44-82
  float u = 0, v = 0;
  float idepth = 1;
  bool isGood = true;
  float energy[2] = {0, 0};
  bool isGood_new = false;
  float idepth_new = 1;
  float energy_new[2] = {0, 0};
  float iR = 1, iRSumNum = 0;
  float lastHessian = 0, lastHessian_new = 0;
  float maxstep = 0;
  int parent = -1;
  int neighbours[10] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
  float my_type = 1, outlierTH = 0;
};

class CoarseInitializer {
 public:
  CoarseInitializer(int w, int h, int levels, int max_points, int device = 0);
  ~CoarseInitializer();
  CoarseInitializer(const CoarseInitializer&) = delete;
  CoarseInitializer& operator=(const CoarseInitializer&) = delete;
  bool ok() const { return ci_ != nullptr; }
  const std::string& error() const { return err_; }

  Settings s;
  int levels() const { return levels_; }
  int width(int l) const { return w_[l]; }
  int height(int l) const { return h_[l]; }
  void makeK(const CalibHessian& HCalib);   // CoarseInitializer.cpp:L967-999
  // CoarseInitializer.cpp:L804-889 without the pixel selector / makeNN: points[lvl] (u, v, my_type, parent, neighbours) are filled by the caller,
  // dIp[l] = firstFrame->dIp[l] (w_l*h_l*3 floats)
  bool setFirst(const float* const* dIp, float ab_exposure);
  // CoarseInitializer.cpp:L85-282
  bool trackFrame(const float* const* dIp, float ab_exposure);

  std::vector<Pnt> points[DMV_MAX_PYR_LEVELS];
  SE3 thisToNext;
  AffLight thisToNext_aff;
  bool snapped = false, fixAffine = true;
  int snappedAt = 0, frameID = -1;
  float alphaK = 2.5f * 2.5f, alphaW = 150 * 150, regWeight = 0.8f, couplingWeight = 1;
  double weightZeroPriorDSOInitX = 0, weightZeroPriorDSOInitY = 0;   // util/settings.cpp:L40-41
  long long evaluations = 0;   // calcResAndGS launches

  // the pieces of trackFrame (public like the adapters' other members: the CPU tests drive them one by one)
  struct System { float H[64], b[8], Hsc[64], bsc[8]; };
  bool calcResAndGS(int lvl, System& out, const SE3& refToNew, AffLight refToNew_aff, float res3[3]);   // L333-625 -> dmv_ci_calc_res_and_gs
  void calcEC(int lvl, float out3[3]);       // L650-670
  void optReg(int lvl);                      // L671-706
  void propagateUp(int srcLvl);              // L708-747
  void propagateDown(int srcLvl);            // L749-777
  void resetPoints(int lvl);                 // L891-917
  void doStep(int lvl, float lambda, const float inc[8]);   // L919-946
  void applyStep(int lvl);                   // L948-965

 private:
  bool fail(const char* what);
  dmv_ci* ci_ = nullptr;
  int w_[DMV_MAX_PYR_LEVELS], h_[DMV_MAX_PYR_LEVELS], levels_ = 1;
  double fx_[DMV_MAX_PYR_LEVELS], fy_[DMV_MAX_PYR_LEVELS], cx_[DMV_MAX_PYR_LEVELS], cy_[DMV_MAX_PYR_LEVELS];
  double Ki_[DMV_MAX_PYR_LEVELS][9];
  float first_exposure_ = 1, new_exposure_ = 1;
  float wM_[8];
  std::vector<std::array<float, 10>> JbBuffer_, JbBuffer_new_;
  bool points_uploaded_ = false;
  std::string err_;
};

}  // namespace dmvio_b200
