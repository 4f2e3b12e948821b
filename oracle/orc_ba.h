// TEST INFRASTRUCTURE ONLY — CPU oracle (restatement) of DM-VIO's sliding-window photometric
// bundle-adjustment hot path.  The reference has no tests/golden vectors on this path and its build
// (Eigen3/Boost/GTSAM) cannot run in this image; this restatement is validated by finite differences, closed-form cases
// and invariants (tests/test_oracle_*.py).
//
// Every function cites the reference file:line (relative to /root/reference/src/dso) it follows.
// PINNED: bit-exact against the reference's own translation units compiled into oracle/_ref (oracle/ref_build.sh,
// oracle/ref_harness.cpp, tests/test_ref_pin.py; DESIGN.md §2 lists the substituted third-party headers).
#pragma once
#include "orc_math.h"
#include <cstdint>
#include <vector>

namespace orc {

constexpr int PATTERN_NUM = 8;   // util/settings.h:L227-229 (patternNum 8)
constexpr int CPARS = 4;         // util/NumType.h:L54
constexpr int NUM_THREADS = 6;   // util/NumType.h:L42
// util/settings.cpp:L296 (staticPattern[8])
static const int patternP[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

// FullSystem/HessianBlocks.h:L60-68
constexpr float SCALE_IDEPTH = 1.0f, SCALE_XI_ROT = 1.0f, SCALE_XI_TRANS = 1.0f;
constexpr float SCALE_F = 50.0f, SCALE_C = 50.0f, SCALE_A = 10.0f, SCALE_B = 1000.0f;

enum ResState { RS_IN = 0, RS_OOB = 1, RS_OUTLIER = 2 };  // FullSystem/Residuals.h:L43

// util/settings.cpp:L60-160
struct Settings {
  float huberTH = 9;
  float outlierTH = 12 * 12;
  float outlierTHSumComponent = 50 * 50;
  float overallEnergyTHWeight = 1;
  float coarseCutoffTH = 20;
  float affineOptModeA = 1e12f, affineOptModeB = 1e8f;
  float idepthFixPrior = 50 * 50;
  float idepthFixPriorMargFac = 600 * 600;
  float initialRotPrior = 1e11f, initialTransPrior = 1e10f, initialAffAPrior = 1e14f, initialAffBPrior = 1e14f;
  float initialCalibHessian = 5e9f;
  float frameEnergyTHConstWeight = 0.5f, frameEnergyTHN = 0.7f, frameEnergyTHFacMedian = 1.5f;
  float margWeightFac = 0.25f;
  float minIdepthH_marg = 50;               // util/settings.cpp:L89
  float thOptIterations = 1.2f;
  int minOptIterations = 1;
  double solverModeDelta = 0.00001;
  bool orthogonalizeXLater = true;  // SOLVER_ORTHOGONALIZE_X_LATER (settings.cpp:L81)
};

struct Calib {  // FullSystem/HessianBlocks.h:L309-409 (CalibHessian)
  double value[4], value_zero[4], value_scaled[4], value_backup[4], step[4];
  float value_scaledf[4], value_scaledi[4];
  double value_minus_value_zero[4];
  float fxl() const { return value_scaledf[0]; }
  float fyl() const { return value_scaledf[1]; }
  float cxl() const { return value_scaledf[2]; }
  float cyl() const { return value_scaledf[3]; }
  float fxli() const { return value_scaledi[0]; }
  float fyli() const { return value_scaledi[1]; }
  float cxli() const { return value_scaledi[2]; }
  float cyli() const { return value_scaledi[3]; }
  void setValue(const double* v);
  void setValueScaled(const double* vs);
};

struct AffLight {  // util/NumType.h:L166-192
  double a = 0, b = 0;
  static void fromToVecExposure(float exposureF, float exposureT, AffLight g2F, AffLight g2T, double out[2]);
};

struct FramePrecalc {  // FullSystem/HessianBlocks.h:L80-107
  Mat33f PRE_RTll, PRE_KRKiTll, PRE_RKiTll, PRE_RTll_0;
  float PRE_aff_mode[2];
  float PRE_b0_mode;
  Vec3f PRE_tTll, PRE_KtTll, PRE_tTll_0;
  float distanceLL;
};

struct Frame {  // FullSystem/HessianBlocks.h:L113-307 (FrameHessian) + EnergyFunctionalStructs.h:L140-165 (EFFrame)
  SE3 worldToCam_evalPT;
  Vec10 state_zero, state_scaled, state, step, state_backup;
  SE3 PRE_worldToCam, PRE_camToWorld;
  float frameEnergyTH = 8 * 8 * PATTERN_NUM;
  float ab_exposure = 1;
  int frameID = 0;
  bool addCamPrior = false;
  const float* dI = nullptr;  // [h*w][3] = I, dx, dy   (level 0)
  Vec8 prior, delta_prior, delta;  // EFFrame
  void setState(const Vec10& s);
  void setStateScaled(const Vec10& s);
  AffLight aff_g2l() const { AffLight l; l.a = state_scaled[6]; l.b = state_scaled[7]; return l; }
  AffLight aff_g2l_0() const { AffLight l; l.a = state_zero[6] * SCALE_A; l.b = state_zero[7] * SCALE_B; return l; }
  Vec10 get_state_minus_stateZero() const { return state - state_zero; }
};

struct RawJ {  // OptimizationBackend/RawResidualJacobian.h:L32-61
  float resF[8];
  float Jpdxi[2][6];
  float Jpdc[2][4];
  float Jpdd[2];
  float JIdx[2][8];
  float JabF[2][8];
  float JIdx2[2][2];
  float JabJIdx[2][2];
  float Jab2[2][2];
};
constexpr int RAWJ_FLOATS = sizeof(RawJ) / sizeof(float);  // 74

struct Residual {  // FullSystem/Residuals.h:L53-110 (PointFrameResidual) + EnergyFunctionalStructs.h:L51-100 (EFResidual)
  int point = -1, host = -1, target = -1;
  int state_state = RS_IN, state_NewState = RS_OUTLIER;
  double state_energy = 0, state_NewEnergy = 0, state_NewEnergyWithOutlier = 0;
  float centerProjectedTo[3] = {0, 0, 0};
  float projectedTo[8][2];
  bool isNew = true;
  RawJ Jnew;  // PointFrameResidual::J  (written by linearize)
  RawJ Jef;   // EFResidual::J          (swapped in by takeDataF)
  float res_toZeroF[8] = {0};
  float JpJdF[8] = {0};
  bool isLinearized = false;
  bool dropped = false;  // deleted by linearizeAll(true) (FullSystemOptimize.cpp:L196-214); kept in the vector so that indices stay stable
  bool isActiveAndIsGoodNEW = false;
  bool isActive() const { return isActiveAndIsGoodNEW; }
};

struct Point {  // HessianBlocks.h:L413-508 (PointHessian) + EnergyFunctionalStructs.h:L104-137 (EFPoint)
  int host = 0;
  float u = 0, v = 0;
  float idepth = 0, idepth_zero = 0, idepth_backup = 0, step = 0;
  float color[8], weights[8];
  bool hasDepthPrior = false;
  float maxRelBaseline = 0;
  int numGoodResiduals = 0;
  float idepth_hessian = 0;
  std::vector<int> residuals;  // indices into Window::residuals (residualsAll order)
  // EFPoint
  float priorF = 0, deltaF = 0;
  float bdSumF = 0, HdiF = 0;
  float Hdd_accLF = 0, Hcd_accLF[4] = {0, 0, 0, 0}, bd_accLF = 0;
  float Hdd_accAF = 0, Hcd_accAF[4] = {0, 0, 0, 0}, bd_accAF = 0;
};

struct ReducedSystem {
  int N = 0;
  MatX HA, HL, Hsc;
  VecX bA, bL, bsc;
  int resInA = 0, resInL = 0;
};

class ThreadPool;

// accumulate precision: 0 = faithful fp32 with the reference's 1/1k/1M tiering, 1 = fp64 accumulators
struct Window {
  int w = 0, h = 0;
  Settings s;
  Calib calib;
  std::vector<Frame> frames;
  std::vector<Point> points;        // ordered by host frame like EnergyFunctional::allPoints (makeIDX, EnergyFunctional.cpp:L998-1016)
  std::vector<Residual> residuals;
  std::vector<FramePrecalc> precalc;     // [h*nf + t]  (host->targetPrecalc[target])
  std::vector<Mat88> adHost, adTarget;   // [h + t*nf]
  std::vector<Mat88f> adHostF, adTargetF;
  std::vector<Mat<float, 1, 8>> adHTdeltaF;  // [h + t*nf]
  float cDeltaF[4] = {0, 0, 0, 0};
  double cPrior[4];
  MatX HM;
  VecX bM;
  VecX lastX;
  int nthreads = 1;  // 1 = the reference's nomt path, 6 = IndexThreadReduce
  ThreadPool* pool = nullptr;

  int nf() const { return (int)frames.size(); }
  float wM3G() const { return (float)(w - 3); }
  float hM3G() const { return (float)(h - 3); }

  // FullSystem.cpp:L1670-1680 (setPrecalcValues) -> HessianBlocks.cpp:L193-223 ; EnergyFunctional.cpp:L175-198
  void setPrecalcValues();
  // EnergyFunctional.cpp:L48-108
  void setAdjointsF();
  void setDeltaF();
  void takeDataFrames();  // EFFrame::takeData / EFPoint::takeData (EnergyFunctionalStructs.cpp:L52-85)

  // FullSystemOptimize.cpp:L150-218 (linearizeAll) incl. setNewFrameEnergyTH (L96-149)
  // returns sum of energies; toRemove (if fixLinearization) gets residual indices to drop
  double linearizeAll(bool fixLinearization, std::vector<int>* toRemove, bool updateEnergyTH = true);
  void applyResAll();  // applyRes_Reductor(true)
  void setNewFrameEnergyTH();

  // EnergyFunctional.cpp:L201-265
  void accumulate(ReducedSystem& sys, int precision);
  // Gauge nullspaces of the window at the frames' evaluation points: FrameHessian::setStateZero (HessianBlocks.cpp:L74-126: numeric
  // derivatives of a global left perturbation / a global scale change, seen as left increments of each frame) assembled by
  // FullSystem::getNullspaces (FullSystemOptimize.cpp:L704-760).  Returns 7 vectors of size 8 nf + 4: 6 pose + 1 scale.
  std::vector<VecX> getNullspaces() const;
  // EnergyFunctional::orthogonalize (EnergyFunctional.cpp:L784-838) for a vector: x -= N N^+ x with N = normalised [pose | scale]
  // nullspaces and singular values below solverModeDelta * max dropped.
  void orthogonalize(VecX& x) const;
  // EnergyFunctional.cpp:L841-996 (default solver mode, no GTSAM branch L971-973) — fills lastX, steps
  void solveSystem(int iteration, double lambda, int precision, ReducedSystem* sysOut = nullptr, MatX* HFinal = nullptr, VecX* bFinal = nullptr);
  void resubstitute(const VecX& x);  // EnergyFunctional.cpp:L267-321
  double calcLEnergy();              // EnergyFunctional.cpp:L349-431
  double calcMEnergy();              // EnergyFunctional.cpp:L324-346 (no GTSAM)

  // Marginalisation of a point subset (makeKeyFrame's flagPointsForRemoval + marginalizePointsF):
  //   fixLinearization(): the compute part of FullSystem::flagPointsForRemoval (FullSystem.cpp:L826-838) for every listed point:
  //     resetOOB, linearize, applyRes(true) and EFResidual::fixLinearizationF (EnergyFunctionalStructs.cpp:L88-114) per residual;
  //     returns ngoodRes per point.
  //   marginalizePoints(): EnergyFunctional::marginalizePointsF (EnergyFunctional.cpp:L678-742) restricted to the listed points
  //     (the caller's PS_MARGINALIZE set): priorF *= idepthFixPriorMargFac, addPoint<2> + SC addPoint(shiftPriorToZero = false),
  //     stitch without priors; sys gets M/Mb (HA/bA) and Msc/Mbsc; HM += margWeightFac (M - Msc), bM likewise. Points are not erased.
  std::vector<int> fixLinearization(const std::vector<int>& pts);
  void marginalizePoints(const std::vector<int>& pts, int precision, ReducedSystem& sys);

  // EnergyFunctional::marginalizeFrame (EnergyFunctional.cpp:L522-675, visual branch L569-631): the frame's 8 rows/columns of HM, bM are
  // moved to the end, its prior is added, the system is diagonally scaled, the 8x8 block is inverted and Schur-complemented away, the
  // result is unscaled and symmetrised.  The frame must not host points or be targeted by residuals any more; it is erased from `frames`
  // (adjoints / precalc / deltas must be rebuilt by the caller, as the reference invalidates them).
  void marginalizeFrame(int idx);

  // FullSystemOptimize.cpp:L224-388
  void backupState();
  bool doStepFromBackup();
  void loadStateBackup();
  // FullSystemOptimize.cpp:L417-647 (without IMU / GTSAM / logging); returns number of iterations run
  int optimize(int mnumOptIts, int precision, std::vector<double>* energyLog = nullptr);
  // the tail of FullSystem::optimize (FullSystemOptimize.cpp:L591-609): the newest frame's evaluation point moves to its estimate
  // (FrameHessian::setEvalPT, HessianBlocks.h:L237-245), setAdjointsF, setPrecalcValues, then linearizeAll(true): applyRes, maxRelBaseline /
  // numGoodResiduals of the active residuals (L55-88), setNewFrameEnergyTH, and the list of residuals the reference now deletes (L186-215).
  // Returns the energy; `toRemove` gets the residual indices (they are marked `dropped` and skipped from then on).
  double finishOptimize(std::vector<int>* toRemove);
};

// Residuals.cpp:L78-274. T = float follows the reference's arithmetic; T = double is used for finite-difference tests.
template <class T>
double linearizeOne(const Window& W, Residual& r, RawJ* Jout);
// Residuals.cpp:L306-328 + EnergyFunctionalStructs.cpp:L39-49
void applyRes(Residual& r);

}  // namespace orc
