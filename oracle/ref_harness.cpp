// TEST INFRASTRUCTURE ONLY — C harness around the REFERENCE's own hot-path code, compiled from where it lies under
// /root/reference into oracle/_ref/libdso_ref.so by oracle/ref_build.sh (nothing of the reference is copied into this repo).
//
// Compiled reference translation units (unmodified): OptimizationBackend/{AccumulatedTopHessian,AccumulatedSCHessian,
// EnergyFunctional,EnergyFunctionalStructs}.cpp, FullSystem/{Residuals,HessianBlocks,ImmaturePoint,CoarseTracker}.cpp,
// util/{settings,globalCalib}.cpp, src/util/TimeMeasurement.cpp.  Third-party headers the image lacks are replaced by the
// stand-ins in oracle/shim/ (Eigen, Sophus, Boost.Thread); the reference headers FullSystem/FullSystem.h and
// IMU/IMUIntegration.hpp are shadowed there because they drag in GTSAM / yaml-cpp (DESIGN.md §2 lists exactly what is
// shadowed).  This file only builds the object graph the reference's functions expect (FrameHessian / PointHessian /
// PointFrameResidual / EnergyFunctional / CoarseTracker) from flat arrays and calls them; it exports the same C entry points as
// oracle/orc_capi.h with the prefix ref_ so that tests/test_ref_pin.py can run the oracle and the reference side by side.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>
#include <deque>
#include <fstream>
#include <iostream>
#include <functional>
#include <thread>
#include <mutex>
#include <condition_variable>

#define private public
#define protected public
#include "FullSystem/FullSystem.h"
#include "FullSystem/CoarseTracker.h"
#include "FullSystem/ImmaturePoint.h"
#define private public  /* the harness drives CoarseInitializer::calcResAndGS & co. directly (test infrastructure) */
#include "FullSystem/CoarseInitializer.h"
#undef private
#include "OptimizationBackend/EnergyFunctional.h"
#include "OptimizationBackend/EnergyFunctionalStructs.h"
#include "OptimizationBackend/AccumulatedTopHessian.h"
#include "OptimizationBackend/AccumulatedSCHessian.h"
#include "util/globalCalib.h"
#include "util/settings.h"
#undef private
#undef protected

using namespace dso;

// Drop-in build (oracle/ref_build.sh dropin): the hot-path members are REPLACED at link time by oracle/dropin_stubs.cpp, which forwards them
// to the CUDA library through include/dmvio_b200.h.  The stubs cache device copies keyed by object; the harness rewrites frames in place
// (the reference allocates new ones), so it announces that here.  No-ops in the plain reference build.
#ifdef DMV_DROPIN
extern "C" { void dropin_invalidate(); void dropin_set_calib(void* hcalib); void dropin_release(const void* owner); }
#else
static inline void dropin_invalidate() {}
static inline void dropin_set_calib(void*) {}
static inline void dropin_release(const void*) {}
#endif

// ---- link-time leftovers of translation units that are NOT compiled (FullSystem.cpp, IOWrapper): never executed
namespace dso {
int PointHessian::instanceCounter = 0;  // FullSystem.cpp:L66-68
int FrameHessian::instanceCounter = 0;
int CalibHessian::instanceCounter = 0;
namespace IOWrap {
void displayImage(const char*, const MinimalImageB3*, bool) {}
int waitKey(int) { return 0; }
void writeImage(std::string, MinimalImageB3*) {}
}  // namespace IOWrap
}  // namespace dso

namespace {

struct RefWin {
  int w, h, nf;
  CalibHessian* Hcalib = nullptr;
  dmvio::BAGTSAMIntegration gtsam;
  EnergyFunctional* ef = nullptr;
  std::vector<FrameHessian*> frames;
  std::vector<PointHessian*> points;
  std::vector<PointFrameResidual*> residuals;
  bool prepared = false;
  // staged inputs
  struct Pt { int host; float u, v, idepth, idepth_zero; float color[8], weights[8]; unsigned char prior; };
  struct Rs { int point, target, state; float energy; };
  std::vector<Pt> pts;
  std::vector<Rs> res;
};

void set_calib_globals(int w, int h, const double K[4]) {
  Eigen::Matrix3f Km;
  Km.setZero();
  Km(0, 0) = (float)K[0]; Km(1, 1) = (float)K[1]; Km(0, 2) = (float)K[2]; Km(1, 2) = (float)K[3]; Km(2, 2) = 1.f;
  setGlobalCalib(w, h, Km);
  setting_useIMU = false;               // visual-only branches of the reference (no GTSAM in this image)
  setting_useGTSAMIntegration = false;
  multiThreading = false;
  setting_debugout_runquiet = true;
}

template <class M> void copy_rowmajor(const M& m, int r, int c, float* out) { for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) out[i * c + j] = (float)m(i, j); }

}  // namespace

extern "C" {

RefWin* ref_win_create(int w, int h, int nf, const double calib_value_scaled[4], int nthreads) {
  set_calib_globals(w, h, calib_value_scaled);
  RefWin* W = new RefWin();
  W->w = w; W->h = h; W->nf = nf;
  W->Hcalib = new CalibHessian();
  W->ef = new EnergyFunctional(W->gtsam);
  W->frames.assign(nf, nullptr);
  // the reference's own worker pool (NUM_THREADS = 6 workers, util/IndexThreadReduce.h); FullSystem always attaches it
  // (FullSystem.cpp:L170) and calcLEnergyF_MT uses it unconditionally; `multiThreading` selects the MT accumulation paths
  W->ef->red = new IndexThreadReduce<Vec10>();
  multiThreading = nthreads > 1;
  return W;
}

void ref_win_destroy(RefWin* W) {
  if (!W) return;
  dropin_release(W->ef);
  // tear down in the reference's ownership order (residuals -> EF mirrors -> points -> frames)
  for (PointFrameResidual* r : W->residuals) { if (r->efResidual) { delete r->efResidual; r->efResidual = 0; } }
  for (PointHessian* p : W->points) { if (p->efPoint) { delete p->efPoint; p->efPoint = 0; } }
  for (FrameHessian* f : W->frames) if (f) { if (f->efFrame) { delete f->efFrame; f->efFrame = 0; } }
  W->ef->frames.clear();
  W->ef->allPoints.clear();
  if (W->ef->red) { delete W->ef->red; W->ef->red = 0; }
  for (FrameHessian* f : W->frames) { FrameShell* sh = f ? f->shell : 0; delete f; delete sh; }  // a frame deletes its pointHessians, which delete their residuals
  delete W->ef;
  delete W->Hcalib;
  delete W;
}

void ref_win_set_setting(RefWin*, const char* name, double v) {
  std::string n(name);
  if (n == "huberTH") setting_huberTH = (float)v;
  else if (n == "outlierTHSumComponent") setting_outlierTHSumComponent = (float)v;
  else if (n == "affineOptModeA") setting_affineOptModeA = (float)v;
  else if (n == "affineOptModeB") setting_affineOptModeB = (float)v;
  else if (n == "solverMode") setting_solverMode = (int)v;
  else fprintf(stderr, "ref_win_set_setting: unknown setting %s\n", name);
}

void ref_win_set_frame(RefWin* W, int idx, const double R[9], const double t[3], const double state[10], const double state_zero[10], float ab_exposure,
                       float frameEnergyTH, int frameID, const float* dI) {
  FrameHessian* fh = new FrameHessian();
  fh->shell = new FrameShell();  // EnergyFunctional::dropResidual counts into host->data->shell (EnergyFunctional.cpp:L510-513)
  fh->shell->id = frameID;
  fh->idx = idx;
  fh->frameID = frameID;
  fh->ab_exposure = ab_exposure;
  fh->frameEnergyTH = frameEnergyTH;
  Mat33 Rm; Vec3 tv;
  for (int i = 0; i < 3; i++) { tv[i] = t[i]; for (int j = 0; j < 3; j++) Rm(i, j) = R[3 * i + j]; }
  fh->worldToCam_evalPT = SE3(Rm, tv);
  Vec10 s, s0;
  for (int i = 0; i < 10; i++) { s[i] = state[i]; s0[i] = state_zero[i]; }
  fh->setStateZero(s0);
  fh->setState(s);
  fh->step.setZero(); fh->step_backup.setZero(); fh->state_backup = s;
  for (int l = 0; l < PYR_LEVELS; l++) { fh->dIp[l] = 0; fh->absSquaredGrad[l] = 0; }
  for (int l = 0; l < pyrLevelsUsed; l++) {
    fh->dIp[l] = new Eigen::Vector3f[wG[l] * hG[l]];
    fh->absSquaredGrad[l] = new float[wG[l] * hG[l]];
  }
  for (int i = 0; i < W->w * W->h; i++) fh->dIp[0][i] = Eigen::Vector3f(dI[3 * i], dI[3 * i + 1], dI[3 * i + 2]);
  fh->dI = fh->dIp[0];
  W->frames[idx] = fh;
}

void ref_win_set_points(RefWin* W, int npts, const int32_t* host, const float* u, const float* v, const float* idepth, const float* idepth_zero,
                        const float* color8, const float* weights8, const uint8_t* hasDepthPrior) {
  W->pts.resize(npts);
  for (int i = 0; i < npts; i++) {
    RefWin::Pt& p = W->pts[i];
    p.host = host[i]; p.u = u[i]; p.v = v[i]; p.idepth = idepth[i]; p.idepth_zero = idepth_zero[i];
    std::memcpy(p.color, color8 + 8 * i, 32); std::memcpy(p.weights, weights8 + 8 * i, 32);
    p.prior = hasDepthPrior ? hasDepthPrior[i] : 0;
  }
}

void ref_win_set_residuals(RefWin* W, int nres, const int32_t* point, const int32_t* target, const int32_t* state_state, const float* state_energy,
                           const uint8_t*) {
  W->res.resize(nres);
  for (int i = 0; i < nres; i++) {
    W->res[i].point = point[i]; W->res[i].target = target[i];
    W->res[i].state = state_state ? state_state[i] : 0;
    W->res[i].energy = state_energy ? state_energy[i] : 0.f;
  }
}

void ref_win_prepare(RefWin* W) {
  // FullSystem::makeKeyFrame order: frames into the energy functional, then points, then residuals
  for (FrameHessian* fh : W->frames) W->ef->insertFrame(fh, W->Hcalib);
  for (const RefWin::Pt& p : W->pts) {
    FrameHessian* host = W->frames[p.host];
    ImmaturePoint ipt((int)p.u, (int)p.v, host, 1.f, W->Hcalib);  // samples colour / weights from the host image (ImmaturePoint.cpp:L34-63)
    ipt.idepth_min = ipt.idepth_max = p.idepth;
    PointHessian* ph = new PointHessian(&ipt, W->Hcalib);
    ph->u = p.u; ph->v = p.v;
    std::memcpy(ph->color, p.color, 32);      // identical inputs on both sides of the comparison
    std::memcpy(ph->weights, p.weights, 32);
    ph->setIdepthScaled(p.idepth);
    ph->setIdepthZero(p.idepth_zero);
    ph->hasDepthPrior = p.prior != 0;
    ph->setPointStatus(PointHessian::ACTIVE);
    ph->step = ph->step_backup = 0; ph->idepth_backup = p.idepth;
    ph->idx = (int)host->pointHessians.size();
    host->pointHessians.push_back(ph);
    W->ef->insertPoint(ph);
    W->points.push_back(ph);
  }
  for (const RefWin::Rs& r : W->res) {
    PointHessian* ph = W->points[r.point];
    PointFrameResidual* pr = new PointFrameResidual(ph, ph->host, W->frames[r.target]);
    pr->setState((ResState)r.state);
    pr->state_energy = r.energy;
    ph->residuals.push_back(pr);
    W->ef->insertResidual(pr);
    W->residuals.push_back(pr);
  }
  W->ef->setAdjointsF(W->Hcalib);
  W->ef->makeIDX();
  // FullSystem::setPrecalcValues (FullSystem.cpp:L1670-1680)
  for (FrameHessian* fh : W->frames) {
    fh->targetPrecalc.resize(W->frames.size());
    for (unsigned int i = 0; i < W->frames.size(); i++) fh->targetPrecalc[i].set(fh, W->frames[i], W->Hcalib);
  }
  W->ef->setDeltaF(W->Hcalib);
  W->prepared = true;
}

void ref_win_get_RT(RefWin* W, float* out) {
  const int nf = W->nf;
  for (int h = 0; h < nf; h++)
    for (int t = 0; t < nf; t++) {
      const FrameFramePrecalc& p = W->frames[h]->targetPrecalc[t];
      float* o = out + (size_t)(h * nf + t) * 12;
      copy_rowmajor(p.PRE_RTll, 3, 3, o);
      for (int k = 0; k < 3; k++) o[9 + k] = p.PRE_tTll[k];
    }
}

// FullSystem::optimizeImmaturePoint (FullSystemOptPoint.cpp:L51-205; a FullSystem member, so its driver loop is mirrored here) around the
// reference's own ImmaturePoint::linearizeResidual (ImmaturePoint.cpp:L498-565).  status: 1 activated, 0 skipped, -1 outlier.
void ref_win_activate(RefWin* W, int n, const int32_t* host, const float* u, const float* v, const float* color8, const float* weights8,
                      const float* energyTH, const float* idepth_min, const float* idepth_max, int minObs, int32_t* status, float* idepth,
                      int32_t* res_state) {
  const int nf = W->nf;
  for (int i = 0; i < n; i++) {
    FrameHessian* hostF = W->frames[host[i]];
    ImmaturePoint ipt(8, 8, hostF, 1.f, W->Hcalib);
    ipt.u = u[i]; ipt.v = v[i];
    std::memcpy(ipt.color, color8 + 8 * i, 32); std::memcpy(ipt.weights, weights8 + 8 * i, 32);
    ipt.energyTH = energyTH[i]; ipt.idepth_min = idepth_min[i]; ipt.idepth_max = idepth_max[i];
    ImmaturePoint* point = &ipt;
    ImmaturePointTemporaryResidual residuals[16];
    int32_t* rs = res_state + (size_t)i * nf;
    for (int f = 0; f < nf; f++) rs[f] = 255;
    int nres = 0;
    for (FrameHessian* fh : W->frames) {
      if (fh != point->host) {
        residuals[nres].state_NewEnergy = residuals[nres].state_energy = 0;
        residuals[nres].state_NewState = ResState::OUTLIER;
        residuals[nres].state_state = ResState::IN;
        residuals[nres].target = fh;
        nres++;
      }
    }
    float lastEnergy = 0, lastHdd = 0, lastbd = 0;
    float currentIdepth = (point->idepth_max + point->idepth_min) * 0.5f;
    for (int k = 0; k < nres; k++) {
      lastEnergy += point->linearizeResidual(W->Hcalib, 1000, residuals + k, lastHdd, lastbd, currentIdepth);
      residuals[k].state_state = residuals[k].state_NewState;
      residuals[k].state_energy = residuals[k].state_NewEnergy;
    }
    idepth[i] = currentIdepth;
    if (!std::isfinite(lastEnergy) || lastHdd < setting_minIdepthH_act) { status[i] = 0; continue; }
    float lambda = 0.1;
    bool skipped = false;
    for (int iteration = 0; iteration < setting_GNItsOnPointActivation; iteration++) {
      float H = lastHdd;
      H *= 1 + lambda;
      float step = (1.0 / H) * lastbd;
      float newIdepth = currentIdepth - step;
      float newHdd = 0; float newbd = 0; float newEnergy = 0;
      for (int k = 0; k < nres; k++) newEnergy += point->linearizeResidual(W->Hcalib, 1, residuals + k, newHdd, newbd, newIdepth);
      if (!std::isfinite(lastEnergy) || newHdd < setting_minIdepthH_act) { skipped = true; break; }
      if (newEnergy < lastEnergy) {
        currentIdepth = newIdepth;
        lastHdd = newHdd; lastbd = newbd; lastEnergy = newEnergy;
        for (int k = 0; k < nres; k++) { residuals[k].state_state = residuals[k].state_NewState; residuals[k].state_energy = residuals[k].state_NewEnergy; }
        lambda *= 0.5;
      } else {
        lambda *= 5;
      }
      if (fabsf(step) < 0.0001 * currentIdepth) break;
    }
    idepth[i] = currentIdepth;
    if (skipped) { status[i] = 0; continue; }
    if (!std::isfinite(currentIdepth)) { status[i] = -1; continue; }
    int numGoodRes = 0;
    for (int k = 0; k < nres; k++) {
      rs[residuals[k].target->idx] = (int)residuals[k].state_state;
      if (residuals[k].state_state == ResState::IN) numGoodRes++;
    }
    if (numGoodRes < minObs) { status[i] = -1; continue; }
    if (!std::isfinite(point->energyTH)) { status[i] = -1; continue; }
    status[i] = 1;
  }
}

int ref_win_nres(RefWin* W) { return (int)W->residuals.size(); }
int ref_win_npts(RefWin* W) { return (int)W->points.size(); }
int ref_win_nf(RefWin* W) { return W->nf; }

void ref_win_get_precalc(RefWin* W, float* out) {
  const int nf = W->nf;
  for (int h = 0; h < nf; h++)
    for (int t = 0; t < nf; t++) {
      float* o = out + (size_t)(h * nf + t) * 32;
      std::memset(o, 0, 32 * sizeof(float));
      const FrameFramePrecalc& p = W->frames[h]->targetPrecalc[t];
      copy_rowmajor(p.PRE_KRKiTll, 3, 3, o);
      for (int k = 0; k < 3; k++) o[9 + k] = p.PRE_KtTll[k];
      copy_rowmajor(p.PRE_RTll_0, 3, 3, o + 12);
      for (int k = 0; k < 3; k++) o[21 + k] = p.PRE_tTll_0[k];
      o[24] = p.PRE_aff_mode[0]; o[25] = p.PRE_aff_mode[1]; o[26] = p.PRE_b0_mode;
    }
}

void ref_win_get_adjoints(RefWin* W, double* adHost, double* adTarget) {
  const int nf = W->nf;
  for (int k = 0; k < nf * nf; k++)
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 8; j++) { adHost[(size_t)k * 64 + i * 8 + j] = W->ef->adHost[k](i, j); adTarget[(size_t)k * 64 + i * 8 + j] = W->ef->adTarget[k](i, j); }
}

void ref_win_get_adHTdeltaF(RefWin* W, float* out) {
  for (int k = 0; k < W->nf * W->nf; k++) for (int i = 0; i < 8; i++) out[k * 8 + i] = W->ef->adHTdeltaF[k][i];
}

void ref_win_get_frame_tables(RefWin* W, double* prior8, double* delta_prior8, double* delta8, float* TH) {
  for (int f = 0; f < W->nf; f++) {
    EFFrame* e = W->frames[f]->efFrame;
    for (int i = 0; i < 8; i++) { prior8[f * 8 + i] = e->prior[i]; delta_prior8[f * 8 + i] = e->delta_prior[i]; delta8[f * 8 + i] = e->delta[i]; }
    TH[f] = W->frames[f]->frameEnergyTH;
  }
}

void ref_win_get_calib(RefWin* W, float* k8, float* cDeltaF4, double* cPrior4) {
  for (int i = 0; i < 4; i++) { k8[i] = W->Hcalib->value_scaledf[i]; k8[4 + i] = W->Hcalib->value_scaledi[i]; cDeltaF4[i] = W->ef->cDeltaF[i]; cPrior4[i] = W->ef->cPrior[i]; }
}

// FullSystem::linearizeAll(false) without setNewFrameEnergyTH (FullSystemOptimize.cpp:L55-88, L150-172): map linearize, sum the energies
double ref_win_linearize_all(RefWin* W, int fixLinearization, int updateEnergyTH) {
  if (fixLinearization || updateEnergyTH) { fprintf(stderr, "ref harness: fixLinearization / setNewFrameEnergyTH live in FullSystem (not compiled)\n"); abort(); }
  double E = 0;
  for (PointFrameResidual* r : W->residuals) E += r->linearize(W->Hcalib);
  return E;
}

void ref_win_apply_res(RefWin* W) { for (PointFrameResidual* r : W->residuals) r->applyRes(true); }

void ref_win_get_res_outputs(RefWin* W, int32_t* newState, float* newEnergy, float* newEnergyWithOutlier, float* cpt3, float* J74, int32_t* state_state,
                             uint8_t* isActive, float* JpJdF8) {
  for (size_t i = 0; i < W->residuals.size(); i++) {
    PointFrameResidual* r = W->residuals[i];
    if (newState) newState[i] = (int)r->state_NewState;
    if (newEnergy) newEnergy[i] = (float)r->state_NewEnergy;
    if (newEnergyWithOutlier) newEnergyWithOutlier[i] = (float)r->state_NewEnergyWithOutlier;
    if (cpt3) for (int k = 0; k < 3; k++) cpt3[3 * i + k] = r->centerProjectedTo[k];
    if (J74) {
      float* o = J74 + 74 * i;
      const RawResidualJacobian* J = r->J;
      for (int k = 0; k < 8; k++) o[k] = J->resF[k];
      for (int a = 0; a < 2; a++) for (int k = 0; k < 6; k++) o[8 + a * 6 + k] = J->Jpdxi[a][k];
      for (int a = 0; a < 2; a++) for (int k = 0; k < 4; k++) o[20 + a * 4 + k] = J->Jpdc[a][k];
      o[28] = J->Jpdd[0]; o[29] = J->Jpdd[1];
      for (int a = 0; a < 2; a++) for (int k = 0; k < 8; k++) { o[30 + a * 8 + k] = J->JIdx[a][k]; o[46 + a * 8 + k] = J->JabF[a][k]; }
      for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) { o[62 + a * 2 + b] = J->JIdx2(a, b); o[66 + a * 2 + b] = J->JabJIdx(a, b); o[70 + a * 2 + b] = J->Jab2(a, b); }
    }
    if (state_state) state_state[i] = (int)r->state_state;
    if (isActive) isActive[i] = r->efResidual->isActiveAndIsGoodNEW ? 1 : 0;
    if (JpJdF8) for (int k = 0; k < 8; k++) JpJdF8[8 * i + k] = r->efResidual->JpJdF[k];
  }
}

static void out_mat(const MatXX& M, double* out) { if (out) for (int i = 0; i < M.rows(); i++) for (int j = 0; j < M.cols(); j++) out[(size_t)i * M.cols() + j] = M(i, j); }
static void out_vec(const VecX& v, double* out) { if (out) for (int i = 0; i < v.size(); i++) out[i] = v[i]; }

// the accumulate half of EnergyFunctional::solveSystemF (EnergyFunctional.cpp:L853-860)
void ref_win_accumulate(RefWin* W, int, double* HA, double* bA, double* HL, double* bL, double* Hsc, double* bsc, int* resInA) {
  dropin_set_calib(W->Hcalib);
  MatXX HA_top, HL_top, H_sc;
  VecX bA_top, bL_top, b_sc;
  W->ef->accumulateAF_MT(HA_top, bA_top, multiThreading);
  W->ef->accumulateLF_MT(HL_top, bL_top, multiThreading);
  W->ef->accumulateSCF_MT(H_sc, b_sc, multiThreading);
  out_mat(HA_top, HA); out_vec(bA_top, bA); out_mat(HL_top, HL); out_vec(bL_top, bL); out_mat(H_sc, Hsc); out_vec(b_sc, bsc);
  if (resInA) *resInA = W->ef->resInA;
}

void ref_win_get_point_outputs(RefWin* W, float* Hdd, float* bd, float* Hcd4, float* HdiF, float* bdSumF, float* step, float* idepth, float* maxRelBaseline) {
  for (size_t i = 0; i < W->points.size(); i++) {
    PointHessian* p = W->points[i];
    EFPoint* e = p->efPoint;
    if (Hdd) Hdd[i] = e->Hdd_accAF;
    if (bd) bd[i] = e->bd_accAF;
    if (Hcd4) for (int k = 0; k < 4; k++) Hcd4[4 * i + k] = e->Hcd_accAF[k];
    if (HdiF) HdiF[i] = e->HdiF;
    if (bdSumF) bdSumF[i] = e->bdSumF;
    if (step) step[i] = p->step;
    if (idepth) idepth[i] = p->idepth;
    if (maxRelBaseline) maxRelBaseline[i] = p->maxRelBaseline;
  }
}

// FullSystem::getNullspaces (FullSystemOptimize.cpp:L704-760; a FullSystem member, mirrored) from the reference's own
// FrameHessian::nullspaces_pose / nullspaces_scale (computed by its setStateZero, HessianBlocks.cpp:L74-126)
static void fill_nullspaces(RefWin* W) {
  EnergyFunctional* ef = W->ef;
  ef->lastNullspaces_pose.clear(); ef->lastNullspaces_scale.clear(); ef->lastNullspaces_affA.clear(); ef->lastNullspaces_affB.clear();
  const int n = CPARS + (int)W->frames.size() * 8;
  for (int i = 0; i < 6; i++) {
    VecX nullspace_x0(n);
    nullspace_x0.setZero();
    for (FrameHessian* fh : W->frames) {
      nullspace_x0.segment<6>(CPARS + fh->idx * 8) = fh->nullspaces_pose.col(i);
      nullspace_x0.segment<3>(CPARS + fh->idx * 8) *= SCALE_XI_TRANS_INVERSE;
      nullspace_x0.segment<3>(CPARS + fh->idx * 8 + 3) *= SCALE_XI_ROT_INVERSE;
    }
    ef->lastNullspaces_pose.push_back(nullspace_x0);
  }
  VecX nullspace_x0(n);
  nullspace_x0.setZero();
  for (FrameHessian* fh : W->frames) {
    nullspace_x0.segment<6>(CPARS + fh->idx * 8) = fh->nullspaces_scale;
    nullspace_x0.segment<3>(CPARS + fh->idx * 8) *= SCALE_XI_TRANS_INVERSE;
    nullspace_x0.segment<3>(CPARS + fh->idx * 8 + 3) *= SCALE_XI_ROT_INVERSE;
  }
  ef->lastNullspaces_scale.push_back(nullspace_x0);
}
void ref_win_get_nullspaces(RefWin* W, double* out) {
  fill_nullspaces(W);
  const int N = CPARS + (int)W->frames.size() * 8;
  for (int a = 0; a < 6; a++) for (int i = 0; i < N; i++) out[(size_t)a * N + i] = W->ef->lastNullspaces_pose[a][i];
  for (int i = 0; i < N; i++) out[(size_t)6 * N + i] = W->ef->lastNullspaces_scale[0][i];
}
// the reference's EnergyFunctional::orthogonalize (EnergyFunctional.cpp:L784-838) on a vector
void ref_win_orthogonalize(RefWin* W, double* x) {
  fill_nullspaces(W);
  const int N = CPARS + (int)W->frames.size() * 8;
  VecX v(N);
  for (int i = 0; i < N; i++) v[i] = x[i];
  W->ef->orthogonalize(&v, 0);
  for (int i = 0; i < N; i++) x[i] = v[i];
}

// EnergyFunctional::solveSystemF (EnergyFunctional.cpp:L841-996), no-GTSAM branch, including resubstituteF_MT
void ref_win_solve(RefWin* W, int iteration, double lambda, int, double* x_out, double* HFinal, double* bFinal) {
  fill_nullspaces(W);  // FullSystem::solveSystem (FullSystemOptimize.cpp:L655-661) refreshes them before every solve
  dropin_set_calib(W->Hcalib);
  W->ef->solveSystemF(iteration, lambda, W->Hcalib);
  out_vec(W->ef->lastX, x_out);
  out_mat(W->ef->lastHS, HFinal);
  out_vec(W->ef->lastbS, bFinal);
}

// One GN iteration of the hot path WITHOUT the dense solve, as timed by bench.py (same sequence as orc_win_hot_iteration):
// backup, accumulateAF/LF/SCF (+stitch), resubstituteF_MT(x), doStepFromBackup (frames, points, setPrecalcValues), linearizeAll, applyRes,
// restore.  The pieces that live in FullSystem (not compiled) are restated with the reference's own setters; linearize / applyRes run on
// the reference's IndexThreadReduce pool with its split rules (static n/NUM_THREADS for linearize, chunks of 50 for applyRes).
double ref_win_hot_iteration(RefWin* W, const double* x, int) {
  EnergyFunctional* ef = W->ef;
  const int N = W->nf * 8 + CPARS;
  dropin_set_calib(W->Hcalib);
  // FullSystem::backupState (FullSystemOptimize.cpp:L322-370)
  W->Hcalib->value_backup = W->Hcalib->value;
  for (FrameHessian* fh : W->frames) fh->state_backup = fh->get_state();
  for (PointHessian* ph : W->points) ph->idepth_backup = ph->idepth;
  MatXX HA, HL, Hsc;
  VecX bA, bL, bsc;
  ef->accumulateAF_MT(HA, bA, multiThreading);
  ef->accumulateLF_MT(HL, bL, multiThreading);
  ef->accumulateSCF_MT(Hsc, bsc, multiThreading);
  VecX xv(N);
  for (int i = 0; i < N; i++) xv[i] = x[i];
  ef->resubstituteF_MT(xv, W->Hcalib, multiThreading);
  // FullSystem::doStepFromBackup (L224-317), stepfac = 1
  W->Hcalib->setValue(W->Hcalib->value_backup + W->Hcalib->step);
  for (FrameHessian* fh : W->frames) {
    fh->setState(fh->state_backup + fh->step);
    for (PointHessian* ph : fh->pointHessians) { ph->setIdepth(ph->idepth_backup + ph->step); ph->setIdepthZero(ph->idepth_backup + ph->step); }
  }
  EFDeltaValid = false;
  for (FrameHessian* fh : W->frames)
    for (unsigned int i = 0; i < W->frames.size(); i++) fh->targetPrecalc[i].set(fh, W->frames[i], W->Hcalib);
  ef->setDeltaF(W->Hcalib);
  // FullSystem::linearizeAll(false) (L150-172) + applyRes_Reductor (L90-94)
  double E = 0;
  const int n = (int)W->residuals.size();
  if (multiThreading) {
    ef->red->reduce([W](int a, int b, Vec10* stats, int) { for (int k = a; k < b; k++) (*stats)[0] += W->residuals[k]->linearize(W->Hcalib); }, 0, n, 0);
    E = ef->red->stats[0];
    ef->red->reduce([W](int a, int b, Vec10*, int) { for (int k = a; k < b; k++) W->residuals[k]->applyRes(true); }, 0, n, 50);
  } else {
    for (PointFrameResidual* r : W->residuals) E += r->linearize(W->Hcalib);
    for (PointFrameResidual* r : W->residuals) r->applyRes(true);
  }
  // FullSystem::loadSateBackup (L371-388)
  W->Hcalib->setValue(W->Hcalib->value_backup);
  for (FrameHessian* fh : W->frames) {
    fh->setState(fh->state_backup);
    for (PointHessian* ph : fh->pointHessians) { ph->setIdepth(ph->idepth_backup); ph->setIdepthZero(ph->idepth_backup); }
  }
  EFDeltaValid = false;
  for (FrameHessian* fh : W->frames)
    for (unsigned int i = 0; i < W->frames.size(); i++) fh->targetPrecalc[i].set(fh, W->frames[i], W->Hcalib);
  ef->setDeltaF(W->Hcalib);
  return E;
}

// Marginalisation of the listed points.  The per-point loop is the body of FullSystem::flagPointsForRemoval (FullSystem.cpp:L826-838; a
// FullSystem member, so the loop is mirrored here around the reference's own resetOOB / linearize / applyRes / fixLinearizationF), then the
// reference's EnergyFunctional::marginalizePointsF (EnergyFunctional.cpp:L678-742) runs unchanged on the flagged points.  It ERASES the
// points from the energy functional: call this last on a window.  rtz8 / isLin are indexed like the residual list of ref_win_set_residuals.
void ref_win_marginalize(RefWin* W, int n, const int32_t* pts, int, double* M, double* Mb, double* Msc, double* Mbsc, double* HM, double* bM,
                         int* resInM, int32_t* ngood, float* rtz8, uint8_t* isLin) {
  std::unordered_map<PointFrameResidual*, size_t> index;
  for (size_t i = 0; i < W->residuals.size(); i++) index[W->residuals[i]] = i;
  if (isLin) std::memset(isLin, 0, W->residuals.size());
  if (rtz8) std::memset(rtz8, 0, sizeof(float) * 8 * W->residuals.size());
  for (int k = 0; k < n; k++) {
    PointHessian* ph = W->points[pts[k]];
    int ngoodRes = 0;
    for (PointFrameResidual* r : ph->residuals) {
      r->resetOOB();
      r->linearize(W->Hcalib);
      r->efResidual->isLinearized = false;
      r->applyRes(true);
      if (r->efResidual->isActive()) {
        r->efResidual->fixLinearizationF(W->ef);
        ngoodRes++;
      }
      const size_t i = index[r];
      if (isLin) isLin[i] = r->efResidual->isLinearized ? 1 : 0;
      if (rtz8 && r->efResidual->isLinearized) for (int c = 0; c < 8; c++) rtz8[8 * i + c] = r->efResidual->res_toZeroF[c];
    }
    if (ngood) ngood[k] = ngoodRes;
    ph->efPoint->stateFlag = EFPointStatus::PS_MARGINALIZE;
  }
  const int resBefore = W->ef->resInM;
  // the two stitched halves are locals of marginalizePointsF: recover M - Msc from the change of HM (margWeightFac is a global setting)
  MatXX HM0 = W->ef->HM; VecX bM0 = W->ef->bM;
  W->ef->marginalizePointsF();
  if (resInM) *resInM = W->ef->resInM - resBefore;
  MatXX dH = (W->ef->HM - HM0) * (1.0 / setting_margWeightFac);
  VecX db = (W->ef->bM - bM0) * (1.0 / setting_margWeightFac);
  out_mat(dH, M); out_vec(db, Mb);  // M := M - Msc, Mb := Mb - Mbsc (Msc / Mbsc are reported as zero)
  if (Msc) std::memset(Msc, 0, sizeof(double) * dH.rows() * dH.cols());
  if (Mbsc) std::memset(Mbsc, 0, sizeof(double) * db.size());
  out_mat(W->ef->HM, HM); out_vec(W->ef->bM, bM);
  // the flagged points are gone from the energy functional (efPoint / efResidual are null now; ref_win_destroy copes)
}

// The reference's EnergyFunctional::marginalizeFrame (EnergyFunctional.cpp:L522-675) on a given prior.  The frame must host no points and
// be the target of no residuals (frames-only window).  The frame object stays in the harness list for ref_win_destroy (efFrame is null).
void ref_win_marginalize_frame(RefWin* W, int idx, const double* HM_in, const double* bM_in, double* HM_out, double* bM_out) {
  const int odim = W->nf * 8 + CPARS;
  W->ef->HM = MatXX::Zero(odim, odim);
  W->ef->bM = VecX::Zero(odim);
  for (int i = 0; i < odim; i++) {
    W->ef->bM[i] = bM_in[i];
    for (int j = 0; j < odim; j++) W->ef->HM(i, j) = HM_in[(size_t)i * odim + j];
  }
  W->ef->HMForGTSAM = MatXX::Zero(odim, odim);
  W->ef->bMForGTSAM = VecX::Zero(odim);
  W->ef->marginalizeFrame(W->frames[idx]->efFrame);
  out_mat(W->ef->HM, HM_out); out_vec(W->ef->bM, bM_out);
  W->nf--;
}

double ref_win_calc_LEnergy(RefWin* W) { return W->ef->calcLEnergyF_MT(); }
double ref_win_calc_MEnergy(RefWin* W) { return W->ef->calcMEnergyF(false); }

// ---- images
int ref_pyr_levels(int w, int h, const double K[4]) { set_calib_globals(w, h, K); return pyrLevelsUsed; }

// FrameHessian::makeImages (HessianBlocks.cpp:L128-191): out = concatenated levels of [I,dx,dy]; returns floats written
int64_t ref_make_images(int w, int h, const double K[4], const float* color, float* dIp_out, float* absSqGrad_out) {
  set_calib_globals(w, h, K);
  CalibHessian Hc;
  FrameHessian* fh = new FrameHessian();
  for (int l = 0; l < PYR_LEVELS; l++) { fh->dIp[l] = 0; fh->absSquaredGrad[l] = 0; }
  std::vector<float> img(color, color + (size_t)w * h);
  fh->makeImages(img.data(), &Hc);
  int64_t n = 0, m = 0;
  for (int l = 0; l < pyrLevelsUsed; l++)
    for (int i = 0; i < wG[l] * hG[l]; i++) {
      for (int k = 0; k < 3; k++) dIp_out[n++] = fh->dIp[l][i][k];
      if (absSqGrad_out) absSqGrad_out[m++] = fh->absSquaredGrad[l][i];
    }
  delete fh;
  return n;
}

// ImmaturePoint constructor (ImmaturePoint.cpp:L34-63): colour and gradient weights of the 8-pattern in the host image
int ref_init_point(const float* dI, int w, int h, const double K[4], float u, float v, float* color8, float* weights8) {
  set_calib_globals(w, h, K);
  CalibHessian Hc;
  FrameHessian* fh = new FrameHessian();
  for (int l = 0; l < PYR_LEVELS; l++) { fh->dIp[l] = 0; fh->absSquaredGrad[l] = 0; }
  for (int l = 0; l < pyrLevelsUsed; l++) { fh->dIp[l] = new Eigen::Vector3f[wG[l] * hG[l]]; fh->absSquaredGrad[l] = new float[wG[l] * hG[l]]; }
  for (int i = 0; i < w * h; i++) fh->dIp[0][i] = Eigen::Vector3f(dI[3 * i], dI[3 * i + 1], dI[3 * i + 2]);
  fh->dI = fh->dIp[0];
  ImmaturePoint ipt((int)u, (int)v, fh, 1.f, &Hc);
  std::memcpy(color8, ipt.color, 32); std::memcpy(weights8, ipt.weights, 32);
  const int ok = std::isfinite(ipt.energyTH) ? 1 : 0;
  delete fh;
  return ok;
}

}  // extern "C"

// =====================================================================================================================
// Coarse tracker: the reference's CoarseTracker driven through its own setCoarseTrackingRef / calcRes / calcGSSSE /
// trackNewestCoarse (FullSystem/CoarseTracker.cpp), visual-only branch.
// =====================================================================================================================
namespace {
struct RefCT {
  int w, h;
  CalibHessian* Hcalib = nullptr;
  dmvio::IMUIntegration imu;
  CoarseTracker* ct = nullptr;
  FrameHessian* host = nullptr;     // hosts the points whose residuals target lastRef
  FrameHessian* lastRef = nullptr;
  FrameHessian* newFrame = nullptr;
  std::vector<PointHessian*> points;
  std::vector<PointFrameResidual*> residuals;
  SE3 lastPose;                     // pose of the last calcRes (calcGSSSE takes it as an argument but only uses the warped buffers)
};

FrameHessian* blank_frame(int id) {
  FrameHessian* fh = new FrameHessian();
  fh->shell = new FrameShell();
  fh->shell->id = id;
  fh->frameID = id; fh->idx = id;
  fh->ab_exposure = 1.f;
  fh->worldToCam_evalPT = SE3();
  Vec10 z = Vec10::Zero();
  fh->setStateZero(z);
  fh->setState(z);
  for (int l = 0; l < PYR_LEVELS; l++) { fh->dIp[l] = 0; fh->absSquaredGrad[l] = 0; }
  for (int l = 0; l < pyrLevelsUsed; l++) {
    fh->dIp[l] = new Eigen::Vector3f[wG[l] * hG[l]];
    fh->absSquaredGrad[l] = new float[wG[l] * hG[l]];
    for (int i = 0; i < wG[l] * hG[l]; i++) { fh->dIp[l][i] = Eigen::Vector3f(0, 0, 0); fh->absSquaredGrad[l][i] = 0; }
  }
  fh->dI = fh->dIp[0];
  return fh;
}
void load_pyramid(FrameHessian* fh, const float* concat) {
  size_t off = 0;
  for (int l = 0; l < pyrLevelsUsed; l++)
    for (int i = 0; i < wG[l] * hG[l]; i++, off += 3) fh->dIp[l][i] = Eigen::Vector3f(concat[off], concat[off + 1], concat[off + 2]);
}
void free_frame(FrameHessian* fh) {
  if (!fh) return;
  if (fh->efFrame) { delete fh->efFrame; fh->efFrame = 0; }
  FrameShell* s = fh->shell;
  delete fh;
  delete s;
}
}  // namespace

extern "C" {

RefCT* ref_ct_create(int w, int h, const double K[4]) {
  set_calib_globals(w, h, K);
  RefCT* C = new RefCT();
  C->w = w; C->h = h;
  C->Hcalib = new CalibHessian();
  C->ct = new CoarseTracker(w, h, C->imu);
  C->ct->makeK(C->Hcalib);
  C->host = blank_frame(0);
  C->lastRef = blank_frame(1);
  C->newFrame = blank_frame(2);
  C->host->efFrame = new EFFrame(C->host);
  C->lastRef->efFrame = new EFFrame(C->lastRef);
  return C;
}

void ref_ct_destroy(RefCT* C) {
  if (!C) return;
  dropin_release(C->ct);
  for (PointFrameResidual* r : C->residuals) { delete r->efResidual; r->efResidual = 0; }
  for (PointHessian* p : C->points) { delete p->efPoint; p->efPoint = 0; }
  free_frame(C->host);  // deletes its pointHessians and their residuals
  free_frame(C->lastRef);
  free_frame(C->newFrame);
  delete C->ct;
  delete C->Hcalib;
  delete C;
}

int ref_ct_levels(RefCT*) { return pyrLevelsUsed; }

// setCoarseTrackingRef (CoarseTracker.cpp:L524-538) -> makeCoarseDepthL0 (L138-295) from per-residual (Ku, Kv, new_idepth) and per-point HdiF
int ref_ct_make_coarse_depth(RefCT* C, int n, const float* Ku, const float* Kv, const float* new_idepth, const float* HdiF, const float* ref_dIp_concat) {
  load_pyramid(C->lastRef, ref_dIp_concat);
  for (int i = 0; i < n; i++) {
    ImmaturePoint ipt(8, 8, C->host, 1.f, C->Hcalib);
    ipt.idepth_min = ipt.idepth_max = 1.f;
    PointHessian* ph = new PointHessian(&ipt, C->Hcalib);
    ph->setPointStatus(PointHessian::ACTIVE);
    C->host->pointHessians.push_back(ph);
    ph->efPoint = new EFPoint(ph, C->host->efFrame);
    ph->efPoint->HdiF = HdiF[i];
    PointFrameResidual* r = new PointFrameResidual(ph, C->host, C->lastRef);
    r->efResidual = new EFResidual(r, ph->efPoint, C->host->efFrame, C->lastRef->efFrame);
    r->efResidual->isActiveAndIsGoodNEW = true;
    r->centerProjectedTo = Vec3f(Ku[i], Kv[i], new_idepth[i]);
    r->setState(ResState::IN);
    ph->residuals.push_back(r);
    ph->lastResiduals[0] = std::pair<PointFrameResidual*, ResState>(r, ResState::IN);
    ph->lastResiduals[1] = std::pair<PointFrameResidual*, ResState>(0, ResState::OOB);
    C->points.push_back(ph);
    C->residuals.push_back(r);
  }
  std::vector<FrameHessian*> fhs;
  fhs.push_back(C->host);
  fhs.push_back(C->lastRef);
  C->ct->setCoarseTrackingRef(fhs);
  dropin_invalidate();
  return C->ct->pc_n[0];
}

int ref_ct_get_ref_points(RefCT* C, int lvl, float* u, float* v, float* idepth, float* color) {
  const int n = C->ct->pc_n[lvl];
  for (int i = 0; i < n; i++) {
    if (u) u[i] = C->ct->pc_u[lvl][i];
    if (v) v[i] = C->ct->pc_v[lvl][i];
    if (idepth) idepth[i] = C->ct->pc_idepth[lvl][i];
    if (color) color[i] = C->ct->pc_color[lvl][i];
  }
  return n;
}

void ref_ct_set_new_frame(RefCT* C, const float* dIp_concat, float ref_exposure, float new_exposure, double ref_a, double ref_b) {
  load_pyramid(C->newFrame, dIp_concat);
  C->newFrame->ab_exposure = new_exposure;
  C->lastRef->ab_exposure = ref_exposure;
  C->ct->lastRef_aff_g2l = AffLight(ref_a, ref_b);
  C->ct->newFrame = C->newFrame;
  dropin_invalidate();
}

void ref_ct_get_K(RefCT* C, int lvl, float* k4, int* wh) {
  k4[0] = C->ct->fx[lvl]; k4[1] = C->ct->fy[lvl]; k4[2] = C->ct->cx[lvl]; k4[3] = C->ct->cy[lvl];
  wh[0] = C->ct->w[lvl]; wh[1] = C->ct->h[lvl];
}

static SE3 pose_from(const double R[9], const double t[3]) {
  Mat33 Rm; Vec3 tv;
  for (int i = 0; i < 3; i++) { tv[i] = t[i]; for (int j = 0; j < 3; j++) Rm(i, j) = R[3 * i + j]; }
  return SE3(Rm, tv);
}

void ref_ct_calc_res(RefCT* C, int lvl, const double R[9], const double t[3], double a, double b, float cutoffTH, double out6[6]) {
  C->lastPose = pose_from(R, t);
  Vec6 r = C->ct->calcRes(lvl, C->lastPose, AffLight(a, b), cutoffTH);
  for (int i = 0; i < 6; i++) out6[i] = r[i];
}

int ref_ct_get_warped(RefCT* C, float* buf8xn) {
  const int n = C->ct->buf_warped_n;
  if (buf8xn) {
    const float* src[8] = {C->ct->buf_warped_idepth, C->ct->buf_warped_u, C->ct->buf_warped_v, C->ct->buf_warped_dx, C->ct->buf_warped_dy,
                           C->ct->buf_warped_residual, C->ct->buf_warped_weight, C->ct->buf_warped_refColor};
    for (int k = 0; k < 8; k++) std::memcpy(buf8xn + (size_t)k * n, src[k], sizeof(float) * n);
  }
  return n;
}

void ref_ct_calc_gs(RefCT* C, int lvl, double a, double b, int, double H64[64], double b8[8]) {
  Mat88 H; Vec8 bb;
  C->ct->calcGSSSE(lvl, H, bb, C->lastPose, AffLight(a, b));
  for (int i = 0; i < 8; i++) { b8[i] = bb[i]; for (int j = 0; j < 8; j++) H64[i * 8 + j] = H(i, j); }
}

int ref_ct_track(RefCT* C, double R[9], double t[3], double* a, double* b, int coarsestLvl, const double minRes[5], int, double lastRes[5],
                 double flow3[3], int* iterations) {
  SE3 T = pose_from(R, t);
  AffLight aff(*a, *b);
  Vec5 mr;
  for (int i = 0; i < 5; i++) mr[i] = minRes[i];
  const bool good = C->ct->trackNewestCoarse(C->newFrame, T, aff, coarsestLvl, mr, 0);
  for (int i = 0; i < 3; i++) { t[i] = T.translation()[i]; for (int j = 0; j < 3; j++) R[3 * i + j] = T.rotationMatrix()(i, j); }
  *a = aff.a; *b = aff.b;
  for (int i = 0; i < 5; i++) lastRes[i] = C->ct->lastResiduals[i];
  for (int i = 0; i < 3; i++) flow3[i] = C->ct->lastFlowIndicators[i];
  if (iterations) *iterations = -1;  // not exposed by the reference
  return good ? 1 : 0;
}

}  // extern "C"

// =====================================================================================================================
// Immature points: the reference's ImmaturePoint constructor and ImmaturePoint::traceOn (FullSystem/ImmaturePoint.cpp)
// =====================================================================================================================
namespace {
FrameHessian* frame_with_dI(const float* dI, int w, int h) {
  FrameHessian* fh = blank_frame(0);
  for (int i = 0; i < w * h; i++) fh->dIp[0][i] = Eigen::Vector3f(dI[3 * i], dI[3 * i + 1], dI[3 * i + 2]);
  return fh;
}
}  // namespace

extern "C" {

int ref_ip_init(int n, const float* dI_host, int w, int h, const double K[4], const int32_t* u, const int32_t* v, float* color8, float* weights8,
                float* gradH4, float* energyTH, uint8_t* ok) {
  set_calib_globals(w, h, K);
  CalibHessian Hc;
  FrameHessian* host = frame_with_dI(dI_host, w, h);
  int good = 0;
  for (int i = 0; i < n; i++) {
    ImmaturePoint ipt(u[i], v[i], host, 1.f, &Hc);
    std::memcpy(color8 + 8 * i, ipt.color, 32); std::memcpy(weights8 + 8 * i, ipt.weights, 32);
    gradH4[4 * i] = ipt.gradH(0, 0); gradH4[4 * i + 1] = ipt.gradH(0, 1); gradH4[4 * i + 2] = ipt.gradH(1, 0); gradH4[4 * i + 3] = ipt.gradH(1, 1);
    energyTH[i] = ipt.energyTH;
    ok[i] = std::isfinite(ipt.energyTH) ? 1 : 0;
    good += ok[i];
  }
  free_frame(host);
  return good;
}

void ref_ip_trace(int n, const float* dI, int w, int h, const double K[4], const float* KRKi9, const float* Kt3, const float* aff2, const float* u,
                  const float* v, const float* color8, const float* weights8, const float* gradH4, const float* energyTH, float* idepth_min,
                  float* idepth_max, float* quality, int32_t* status, float* uv2, float* interval) {
  set_calib_globals(w, h, K);
  CalibHessian Hc;
  FrameHessian* frame = frame_with_dI(dI, w, h);
  FrameHessian* host = blank_frame(1);   // the constructor samples it; every sampled field is overwritten below
  Mat33f KRKi; Vec3f Kt; Vec2f aff;
  for (int i = 0; i < 3; i++) { Kt[i] = Kt3[i]; for (int j = 0; j < 3; j++) KRKi(i, j) = KRKi9[3 * i + j]; }
  aff[0] = aff2[0]; aff[1] = aff2[1];
  for (int i = 0; i < n; i++) {
    ImmaturePoint ipt(8, 8, host, 1.f, &Hc);
    ipt.u = u[i]; ipt.v = v[i];
    std::memcpy(ipt.color, color8 + 8 * i, 32); std::memcpy(ipt.weights, weights8 + 8 * i, 32);
    ipt.gradH(0, 0) = gradH4[4 * i]; ipt.gradH(0, 1) = gradH4[4 * i + 1]; ipt.gradH(1, 0) = gradH4[4 * i + 2]; ipt.gradH(1, 1) = gradH4[4 * i + 3];
    ipt.energyTH = energyTH[i]; ipt.idepth_min = idepth_min[i]; ipt.idepth_max = idepth_max[i]; ipt.quality = quality[i];
    ipt.lastTraceStatus = (ImmaturePointStatus)status[i];
    ipt.lastTraceUV = Vec2f(uv2[2 * i], uv2[2 * i + 1]); ipt.lastTracePixelInterval = interval[i];
    ipt.traceOn(frame, KRKi, Kt, aff, &Hc, false);
    idepth_min[i] = ipt.idepth_min; idepth_max[i] = ipt.idepth_max; quality[i] = ipt.quality; status[i] = (int)ipt.lastTraceStatus;
    uv2[2 * i] = ipt.lastTraceUV[0]; uv2[2 * i + 1] = ipt.lastTraceUV[1]; interval[i] = ipt.lastTracePixelInterval;
  }
  free_frame(host);
  free_frame(frame);
}

}  // extern "C"

// =====================================================================================================================
// Coarse initialiser: the reference's CoarseInitializer (FullSystem/CoarseInitializer.cpp) driven function by function.
// The pixel selection of setFirst and the kd-tree of makeNN are replaced by caller-provided points / parents / neighbours
// (everything else of setFirst, L804-889, is mirrored); trackFrame & co. are the reference's compiled code.
// Note: IndexThreadReduce::reduce always fans out to its worker threads with dynamic chunks of 50 points
// (util/IndexThreadReduce.h:L77-135), so the summed H / b of calcResAndGS are reproducible only up to float summation order
// once a level has more than 50 points; the per-point results are deterministic.
// =====================================================================================================================
namespace {
struct RefCI {
  CalibHessian* Hcalib = nullptr;
  CoarseInitializer* ci = nullptr;
  FrameHessian* first = nullptr;
  FrameHessian* cur = nullptr;
  std::vector<IOWrap::Output3DWrapper*> wraps;
};
}  // namespace

extern "C" {

RefCI* ref_ci_create(int w, int h, const double K[4]) {
  set_calib_globals(w, h, K);
  RefCI* C = new RefCI();
  C->Hcalib = new CalibHessian();
  C->ci = new CoarseInitializer(w, h);
  C->ci->printDebug = false;
  return C;
}
void ref_ci_destroy(RefCI* C) {
  if (!C) return;
  dropin_release(C->ci);
  delete C->ci;
  free_frame(C->first);
  free_frame(C->cur);
  delete C->Hcalib;
  delete C;
}
int ref_ci_levels(RefCI*) { return pyrLevelsUsed; }

void ref_ci_set_first(RefCI* C, const float* dIp_concat, float exposure, const int32_t* n, const float* u, const float* v, const float* type,
                      const int32_t* parent, const int32_t* neighbours10) {
  CoarseInitializer* ci = C->ci;
  free_frame(C->first);
  C->first = blank_frame(0);
  load_pyramid(C->first, dIp_concat);
  C->first->ab_exposure = exposure;
  ci->makeK(C->Hcalib);
  ci->firstFrame = C->first;
  size_t off = 0;
  for (int lvl = 0; lvl < pyrLevelsUsed; lvl++) {
    if (ci->points[lvl] != 0) delete[] ci->points[lvl];
    ci->points[lvl] = new Pnt[n[lvl] > 0 ? n[lvl] : 1];
    Pnt* pl = ci->points[lvl];
    for (int i = 0; i < n[lvl]; i++, off++) {  // CoarseInitializer.cpp:L843-872
      pl[i].u = u[off];
      pl[i].v = v[off];
      pl[i].idepth = 1;
      pl[i].iR = 1;
      pl[i].isGood = true;
      pl[i].energy.setZero();
      pl[i].lastHessian = 0;
      pl[i].lastHessian_new = 0;
      pl[i].my_type = type[off];
      pl[i].outlierTH = patternNum * setting_outlierTH;
      pl[i].parent = parent[off];
      pl[i].parentDist = 0;
      pl[i].isGood_new = false; pl[i].idepth_new = 1; pl[i].energy_new.setZero(); pl[i].iRSumNum = 0; pl[i].maxstep = 0;
      for (int k = 0; k < 10; k++) { pl[i].neighbours[k] = neighbours10[10 * off + k]; pl[i].neighboursDist[k] = 0; }
    }
    ci->numPoints[lvl] = n[lvl];
  }
  ci->thisToNext = SE3();
  ci->thisToNext_aff = AffLight(0, 0);
  ci->snapped = false;
  ci->frameID = ci->snappedAt = 0;
  for (int i = 0; i < pyrLevelsUsed; i++) ci->dGrads[i].setZero();
  // trackFrame sets these at its start (CoarseInitializer.cpp:L94-97); the function-by-function tests call calcResAndGS before any trackFrame
  ci->alphaK = 2.5 * 2.5;
  ci->alphaW = 150 * 150;
  ci->regWeight = 0.8;
  ci->couplingWeight = 1;
  dropin_invalidate();
}
void ref_ci_set_new(RefCI* C, const float* dIp_concat, float exposure) {
  free_frame(C->cur);
  C->cur = blank_frame(1);
  load_pyramid(C->cur, dIp_concat);
  C->cur->ab_exposure = exposure;
  C->ci->newFrame = C->cur;
  dropin_invalidate();
}
void ref_ci_calc(RefCI* C, int lvl, const double R[9], const double t[3], double a, double b, float* H64, float* b8, float* Hsc64, float* bsc8, float* res3) {
  Mat88f H, Hsc; Vec8f bb, bsc;
  Vec3f r = C->ci->calcResAndGS(lvl, H, bb, Hsc, bsc, pose_from(R, t), AffLight(a, b), false);
  for (int i = 0; i < 8; i++) {
    for (int j = 0; j < 8; j++) { H64[i * 8 + j] = H(i, j); Hsc64[i * 8 + j] = Hsc(i, j); }
    b8[i] = bb[i]; bsc8[i] = bsc[i];
  }
  for (int i = 0; i < 3; i++) res3[i] = r[i];
}
void ref_ci_apply_step(RefCI* C, int lvl) { C->ci->applyStep(lvl); }
void ref_ci_do_step(RefCI* C, int lvl, float lambda, const float* inc8) {
  Vec8f inc;
  for (int i = 0; i < 8; i++) inc[i] = inc8[i];
  C->ci->doStep(lvl, lambda, inc);
}
void ref_ci_calc_ec(RefCI* C, int lvl, float* out3) { Vec3f r = C->ci->calcEC(lvl); for (int i = 0; i < 3; i++) out3[i] = r[i]; }
void ref_ci_opt_reg(RefCI* C, int lvl) { C->ci->optReg(lvl); }
void ref_ci_propagate_up(RefCI* C, int lvl) { C->ci->propagateUp(lvl); }
void ref_ci_propagate_down(RefCI* C, int lvl) { C->ci->propagateDown(lvl); }
void ref_ci_reset_points(RefCI* C, int lvl) { C->ci->resetPoints(lvl); }
void ref_ci_set_snapped(RefCI* C, int snapped) { C->ci->snapped = snapped != 0; }
int ref_ci_npts(RefCI* C, int lvl) { return C->ci->numPoints[lvl]; }
void ref_ci_get_points(RefCI* C, int lvl, float* out12) {
  for (int i = 0; i < C->ci->numPoints[lvl]; i++) {
    const Pnt& p = C->ci->points[lvl][i];
    float* q = out12 + 12 * i;
    q[0] = p.idepth; q[1] = p.idepth_new; q[2] = p.iR; q[3] = p.energy[0]; q[4] = p.energy[1]; q[5] = p.energy_new[0]; q[6] = p.energy_new[1];
    q[7] = p.lastHessian; q[8] = p.lastHessian_new; q[9] = p.maxstep; q[10] = p.isGood ? 1.f : 0.f; q[11] = p.isGood_new ? 1.f : 0.f;
  }
}
void ref_ci_set_points(RefCI* C, int lvl, const float* in5) {
  for (int i = 0; i < C->ci->numPoints[lvl]; i++) {
    Pnt& p = C->ci->points[lvl][i];
    const float* q = in5 + 5 * i;
    p.idepth = q[0]; p.idepth_new = q[1]; p.iR = q[2]; p.lastHessian = q[3]; p.isGood = q[4] != 0.f;
  }
}
int ref_ci_track(RefCI* C, const float* dIp_concat, float exposure, double* R9, double* t3, double* ab2, int32_t* state3) {
  ref_ci_set_new(C, dIp_concat, exposure);
  const bool ok = C->ci->trackFrame(C->cur, C->wraps);
  const Mat33 Rm = C->ci->thisToNext.rotationMatrix();
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R9[i * 3 + j] = Rm(i, j); t3[i] = C->ci->thisToNext.translation()[i]; }
  ab2[0] = C->ci->thisToNext_aff.a; ab2[1] = C->ci->thisToNext_aff.b;
  state3[0] = C->ci->snapped ? 1 : 0; state3[1] = C->ci->snappedAt; state3[2] = C->ci->frameID;
  return ok ? 1 : 0;
}

}  // extern "C"

