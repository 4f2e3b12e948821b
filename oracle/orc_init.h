// TEST INFRASTRUCTURE ONLY — CPU oracle (restatement) of DM-VIO's CoarseInitializer (FullSystem/CoarseInitializer.cpp): the two-frame
// direct initialiser that optimises a relative pose, an affine brightness pair and one inverse depth per selected point (eliminated by a
// Schur complement) over the pyramid, with neighbourhood regularisation of the depths.  SURVEY.md section 8f-4 lists calcResAndGS as a
// later row; this restatement (pinned against the reference's compiled code by tests/test_ref_pin.py) is the groundwork for it.
// Pixel selection (setFirst's PixelSelector) and the kd-tree of makeNN are inputs here: points, parents and neighbour lists are given.
#pragma once
#include "orc_ba.h"
#include "orc_coarse.h"

#include <array>
#include <vector>

namespace orc {

struct InitPnt {  // CoarseInitializer.h:L44-82 (Pnt)
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

struct InitSystem { float H[64], b[8], Hsc[64], bsc[8]; };

struct CoarseInit {
  Settings s;
  int levels = 1;
  int w[PYR_LEVELS], h[PYR_LEVELS];
  double fx[PYR_LEVELS], fy[PYR_LEVELS], cx[PYR_LEVELS], cy[PYR_LEVELS];
  Mat33 Ki[PYR_LEVELS];
  std::vector<InitPnt> points[PYR_LEVELS];
  const float* dIFirst[PYR_LEVELS];
  const float* dINew[PYR_LEVELS];
  float first_exposure = 1, new_exposure = 1;
  SE3 thisToNext;
  AffLight thisToNext_aff;
  bool snapped = false, fixAffine = true;
  int snappedAt = 0, frameID = -1;
  float wM[8];
  float alphaK = 2.5f * 2.5f, alphaW = 150 * 150, regWeight = 0.8f, couplingWeight = 1;
  double weightZeroPriorDSOInitX = 0, weightZeroPriorDSOInitY = 0;  // util/settings.cpp:L40-41
  std::vector<std::array<float, 10>> JbBuffer, JbBuffer_new;

  CoarseInit();                                                          // CoarseInitializer.cpp:L49-73
  void makeK(int w0, int h0, double fx0, double fy0, double cx0, double cy0, int forceLevels = 0);  // L967-999 (+ setGlobalCalib's level rule)
  void setFirst(const float* const* dIp, float exposure);                // L804-889 without the pixel selector / makeNN (points are given)
  bool trackFrame(const float* const* dIp, float exposure);              // L85-282
  void calcResAndGS(int lvl, InitSystem& out, const SE3& refToNew, AffLight refToNew_aff, float res3[3]);  // L333-625
  void calcEC(int lvl, float out3[3]);                                   // L650-670
  void optReg(int lvl);                                                  // L671-706
  void propagateUp(int srcLvl);                                          // L708-747
  void propagateDown(int srcLvl);                                        // L749-777
  void resetPoints(int lvl);                                             // L891-917
  void doStep(int lvl, float lambda, const float inc[8]);                // L919-946
  void applyStep(int lvl);                                               // L948-965
};

}  // namespace orc
