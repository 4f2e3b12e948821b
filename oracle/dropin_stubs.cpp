// TEST INFRASTRUCTURE ONLY — the forwarding stubs of INTEGRATION.md §2, COMPILED against the reference's own headers and linked with the
// reference's own translation units into oracle/_ref/libdso_ref_dropin.so (oracle/ref_build.sh dropin).  This is the proof that the C ABI
// of include/dmvio_b200.h is a drop-in for the reference's hot path: the member functions below REPLACE the reference's definitions of the
// same symbols (this object precedes the reference's objects on the link line, -Wl,--allow-multiple-definition), everything else — the
// object graph, EnergyFunctional::solveSystemF with its dense solve, CoarseTracker::trackNewestCoarse with its Levenberg-Marquardt loop —
// is the reference's unmodified code calling into them.  tests/test_gpu_dropin.py runs this library next to the unmodified
// oracle/_ref/libdso_ref.so on the same inputs.
//
//   CoarseTracker::calcRes / calcGSSSE            (FullSystem/CoarseTracker.cpp:L299-517)  -> dmv_ct_calc_res_gs
//   EnergyFunctional::accumulateAF_MT / accumulateSCF_MT  (OptimizationBackend/EnergyFunctional.cpp:L201-265) -> dmv_ba_linearize + dmv_ba_accumulate
//   EnergyFunctional::resubstituteF_MT            (EnergyFunctional.cpp:L267-289)          -> dmv_ba_resubstitute
//   CoarseInitializer::calcResAndGS               (FullSystem/CoarseInitializer.cpp:L333-625) -> dmv_ci_calc_res_and_gs
// A maintainer keeps the handles as members (EnergyFunctional::gpu, CoarseTracker::gpu); here they live in side tables keyed by `this`.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
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
#include "FullSystem/CoarseInitializer.h"
#include "OptimizationBackend/EnergyFunctional.h"
#include "OptimizationBackend/EnergyFunctionalStructs.h"
#include "util/globalCalib.h"
#include "util/settings.h"
#undef private
#undef protected

#include "../include/dmvio_b200.h"

#define DMV_CHECK(call)                                                                  \
  do {                                                                                   \
    if ((call) != DMV_OK) { fprintf(stderr, "dropin: %s failed: %s\n", #call, dmv_last_error()); abort(); } \
  } while (0)

namespace {
unsigned long long g_generation = 1;   // bumped by the harness whenever it rewrites frame data in place (the reference would allocate new frames)
dso::CalibHessian* g_calib = nullptr;  // solveSystemF receives HCalib; accumulateAF_MT (stubbed below) does not: the harness announces it

struct CtState {
  dmv_ct* ct = nullptr;
  unsigned long long gen = 0;
  double H[64], b[8];
};
std::map<const dso::CoarseTracker*, CtState> g_ct;

struct BaState {
  dmv_ba* ba = nullptr;
  int max_points = 0;
  std::vector<double> Hsc, bsc;
  int nFrames = 0;
};
std::map<const dso::EnergyFunctional*, BaState> g_ba;

struct CiState {
  dmv_ci* ci = nullptr;
  unsigned long long gen = 0;
};
std::map<const dso::CoarseInitializer*, CiState> g_ci;
}  // namespace

extern "C" void dropin_invalidate() { g_generation++; }
extern "C" void dropin_set_calib(void* hcalib) { g_calib = static_cast<dso::CalibHessian*>(hcalib); }
extern "C" void dropin_release(const void* owner) {
  auto a = g_ct.find(static_cast<const dso::CoarseTracker*>(owner));
  if (a != g_ct.end()) { dmv_ct_destroy(a->second.ct); g_ct.erase(a); }
  auto c = g_ci.find(static_cast<const dso::CoarseInitializer*>(owner));
  if (c != g_ci.end()) { dmv_ci_destroy(c->second.ci); g_ci.erase(c); }
  auto b = g_ba.find(static_cast<const dso::EnergyFunctional*>(owner));
  if (b != g_ba.end()) { dmv_ba_destroy(b->second.ba); g_ba.erase(b); }
}

namespace dso {

// ---------------------------------------------------------------------------------------------------------------- coarse tracker
static CtState& ct_sync(CoarseTracker* self) {
  CtState& s = g_ct[self];
  if (!s.ct) {
    dmv_ct_config cfg{self->w[0], self->h[0], pyrLevelsUsed, 65536, 0};
    DMV_CHECK(dmv_ct_create(&cfg, &s.ct));
    DMV_CHECK(dmv_ct_set_huber(s.ct, setting_huberTH));
  }
  if (s.gen != g_generation) {  // reference point lists (setCoarseTrackingRef) and the new frame's pyramid, as the reference holds them
    for (int l = 0; l < pyrLevelsUsed; l++) {
      DMV_CHECK(dmv_ct_set_K(s.ct, l, self->fx[l], self->fy[l], self->cx[l], self->cy[l]));
      DMV_CHECK(dmv_ct_set_ref(s.ct, l, self->pc_n[l], self->pc_u[l], self->pc_v[l], self->pc_idepth[l], self->pc_color[l]));
      DMV_CHECK(dmv_ct_upload_new(s.ct, l, reinterpret_cast<const float*>(self->newFrame->dIp[l])));
    }
    s.gen = g_generation;
  }
  return s;
}

Vec6 CoarseTracker::calcRes(int lvl, const SE3& refToNew, AffLight aff_g2l, float cutoffTH) {
  CtState& s = ct_sync(this);
  const Mat33f RKi = (refToNew.rotationMatrix().cast<float>() * Ki[lvl]);
  const Vec3f t = (refToNew.translation()).cast<float>();
  const Vec2f affLL = AffLight::fromToVecExposure(lastRef->ab_exposure, newFrame->ab_exposure, lastRef_aff_g2l, aff_g2l).cast<float>();
  float RKi_rm[9], tt[3] = {t[0], t[1], t[2]}, aff[2] = {affLL[0], affLL[1]};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) RKi_rm[3 * i + j] = RKi(i, j);
  double r6[6];
  int nw = 0;
  DMV_CHECK(dmv_ct_calc_res_gs(s.ct, lvl, RKi_rm, tt, aff, (float)lastRef_aff_g2l.b, cutoffTH, 1, r6, s.H, s.b, &nw));
  buf_warped_n = nw;
  Vec6 rs;
  for (int i = 0; i < 6; i++) rs[i] = r6[i];
  return rs;
}

void CoarseTracker::calcGSSSE(int lvl, Mat88& H_out, Vec8& b_out, const SE3&, AffLight) {
  (void)lvl;
  CtState& s = g_ct[this];   // produced by the fused launch of the calcRes call that precedes every calcGSSSE (CoarseTracker.cpp:L583, L700)
  for (int i = 0; i < 8; i++) {
    b_out[i] = s.b[i];
    for (int j = 0; j < 8; j++) H_out(i, j) = s.H[i * 8 + j];
  }
}


// ---------------------------------------------------------------------------------------------------------------- coarse initialiser
Vec3f CoarseInitializer::calcResAndGS(int lvl, Mat88f& H_out, Vec8f& b_out, Mat88f& H_out_sc, Vec8f& b_out_sc, const SE3& refToNew, AffLight refToNew_aff,
                                      bool plot) {
  (void)plot;
  CiState& s = g_ci[this];
  if (!s.ci) {
    int maxn = 1;
    for (int l = 0; l < pyrLevelsUsed; l++) maxn = std::max(maxn, numPoints[l]);
    dmv_ci_config cfg{w[0], h[0], pyrLevelsUsed, maxn, 0};
    DMV_CHECK(dmv_ci_create(&cfg, &s.ci));
  }
  if (s.gen != g_generation) {   // frames and the constant point fields, as setFirst / trackFrame hold them
    for (int l = 0; l < pyrLevelsUsed; l++) {
      DMV_CHECK(dmv_ci_set_K(s.ci, l, (float)fx[l], (float)fy[l], (float)cx[l], (float)cy[l]));
      DMV_CHECK(dmv_ci_upload_first(s.ci, l, reinterpret_cast<const float*>(firstFrame->dIp[l])));
      DMV_CHECK(dmv_ci_upload_new(s.ci, l, reinterpret_cast<const float*>(newFrame->dIp[l])));
      const int n = numPoints[l];
      std::vector<float> u(n), v(n), th(n);
      for (int i = 0; i < n; i++) { u[i] = points[l][i].u; v[i] = points[l][i].v; th[i] = points[l][i].outlierTH; }
      DMV_CHECK(dmv_ci_set_points(s.ci, l, n, u.data(), v.data(), th.data()));
    }
    s.gen = g_generation;
  }
  const int n = numPoints[lvl];
  Pnt* pts = points[lvl];
  std::vector<float> idn(n), en((size_t)2 * n), iR(n), en_new((size_t)2 * n), mstep(n), lastH(n), jb((size_t)10 * n);
  std::vector<uint8_t> good(n), good_new(n);
  for (int i = 0; i < n; i++) { idn[i] = pts[i].idepth_new; good[i] = pts[i].isGood; en[2 * i] = pts[i].energy[0]; en[2 * i + 1] = pts[i].energy[1]; iR[i] = pts[i].iR; }
  dmv_ci_eval_args a;
  std::memset(&a, 0, sizeof(a));
  a.level = lvl;
  const Mat33f RKi = (refToNew.rotationMatrix() * Ki[lvl]).cast<float>();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) a.RKi[3 * i + j] = RKi(i, j);
  const Vec6 lg = refToNew.log();
  for (int i = 0; i < 3; i++) { a.t_d[i] = refToNew.translation()[i]; a.t_log[i] = lg[i]; }
  const Eigen::Vector2f r2new_aff((float)std::exp(refToNew_aff.a), (float)refToNew_aff.b);
  a.r2new_aff[0] = r2new_aff[0]; a.r2new_aff[1] = r2new_aff[1];
  a.huberTH = setting_huberTH; a.alphaK = alphaK; a.alphaW = alphaW; a.couplingWeight = couplingWeight;
  a.weightZeroPriorX = setting_weightZeroPriorDSOInitX; a.weightZeroPriorY = setting_weightZeroPriorDSOInitY;
  a.idepth_new = idn.data(); a.isGood = good.data(); a.energy2 = en.data(); a.iR = iR.data();
  a.isGood_new = good_new.data(); a.energy_new2 = en_new.data(); a.maxstep = mstep.data(); a.lastHessian_new = lastH.data(); a.JbBuffer_new10 = jb.data();
  dmv_ci_eval_result r;
  DMV_CHECK(dmv_ci_calc_res_and_gs(s.ci, &a, &r));
  for (int i = 0; i < n; i++) {   // what processPointsForReduce and the Schur pass leave in the Pnt array / JbBuffer_new (L369-586)
    Pnt& p = pts[i];
    p.isGood_new = good_new[i] != 0;
    p.energy_new[0] = en_new[2 * i]; p.energy_new[1] = en_new[2 * i + 1];
    if (p.isGood) p.maxstep = mstep[i]; else p.maxstep = 1e10;
    if (p.isGood_new) {
      p.lastHessian_new = lastH[i];
      for (int k = 0; k < 10; k++) JbBuffer_new[i][k] = jb[(size_t)10 * i + k];
    }
  }
  for (int i = 0; i < 8; i++) {
    b_out[i] = r.b[i]; b_out_sc[i] = r.bsc[i];
    for (int j = 0; j < 8; j++) { H_out(i, j) = r.H[i * 8 + j]; H_out_sc(i, j) = r.Hsc[i * 8 + j]; }
  }
  return Vec3f(r.res3[0], r.res3[1], r.res3[2]);
}

// ---------------------------------------------------------------------------------------------------------------- energy functional
// the window as the reference holds it -> the BA handle; one linearisation at the reference's current state
static BaState& ba_sync_and_linearize(EnergyFunctional* ef) {
  BaState& s = g_ba[ef];
  const int nf = ef->nFrames, npts = (int)ef->allPoints.size();
  if (!g_calib) { fprintf(stderr, "dropin: dropin_set_calib() first\n"); abort(); }
  if (!s.ba || s.max_points < npts) {
    if (s.ba) dmv_ba_destroy(s.ba);
    s.max_points = std::max(npts, 4096);
    dmv_ba_config cfg{wG[0], hG[0], DMV_MAX_FRAMES, s.max_points, 0, 0};
    DMV_CHECK(dmv_ba_create(&cfg, &s.ba));
  }
  dmv_ba_params prm;
  dmv_ba_default_params(&prm);
  prm.huberTH = setting_huberTH; prm.outlierTHSumComponent = setting_outlierTHSumComponent;
  prm.affineOptModeA = setting_affineOptModeA; prm.affineOptModeB = setting_affineOptModeB;
  DMV_CHECK(dmv_ba_set_params(s.ba, &prm));
  for (int f = 0; f < nf; f++) DMV_CHECK(dmv_ba_upload_frame(s.ba, f, reinterpret_cast<const float*>(ef->frames[f]->data->dI)));
  DMV_CHECK(dmv_ba_set_window(s.ba, nf, nullptr));
  // allPoints order (EnergyFunctional::makeIDX, EnergyFunctional.cpp:L998-1016) is by host frame
  std::vector<int32_t> host(npts);
  std::vector<float> u(npts), v(npts), id(npts), idz(npts), col((size_t)8 * npts), wgt((size_t)8 * npts), prior(npts);
  std::map<const EFPoint*, int> pidx;
  for (int i = 0; i < npts; i++) {
    const EFPoint* p = ef->allPoints[i];
    const PointHessian* ph = p->data;
    pidx[p] = i;
    host[i] = p->host->idx; u[i] = ph->u; v[i] = ph->v; id[i] = ph->idepth_scaled; idz[i] = ph->idepth_zero_scaled; prior[i] = p->priorF;
    std::memcpy(&col[(size_t)8 * i], ph->color, 32); std::memcpy(&wgt[(size_t)8 * i], ph->weights, 32);
  }
  DMV_CHECK(dmv_ba_set_points(s.ba, npts, host.data(), u.data(), v.data(), id.data(), idz.data(), col.data(), wgt.data(), prior.data()));
  std::vector<int32_t> rp, rt, rs;
  std::vector<float> re;
  for (int i = 0; i < npts; i++)
    for (const EFResidual* r : ef->allPoints[i]->residualsAll) {
      if (r->isLinearized) continue;
      rp.push_back(i); rt.push_back(r->targetIDX);
      rs.push_back((int)r->data->state_state); re.push_back((float)r->data->state_energy);
    }
  DMV_CHECK(dmv_ba_set_residuals(s.ba, (int)rp.size(), rp.data(), rt.data(), rs.data(), re.data()));
  // adjoints (EnergyFunctional::setAdjointsF keeps them as Mat88 [h + t*nFrames]): row-major copies
  std::vector<double> adH((size_t)nf * nf * 64), adT((size_t)nf * nf * 64);
  for (int k = 0; k < nf * nf; k++)
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 8; j++) { adH[(size_t)k * 64 + i * 8 + j] = ef->adHost[k](i, j); adT[(size_t)k * 64 + i * 8 + j] = ef->adTarget[k](i, j); }
  DMV_CHECK(dmv_ba_set_adjoints(s.ba, adH.data(), adT.data()));
  // per-iteration tables straight from the reference's FrameFramePrecalc objects (HessianBlocks.h:L80-107)
  std::vector<float> pc((size_t)nf * nf * DMV_PRECALC_FLOATS, 0.f), th(nf);
  for (int h = 0; h < nf; h++) {
    const FrameHessian* fh = ef->frames[h]->data;
    th[h] = fh->frameEnergyTH;
    for (int t = 0; t < nf; t++) {
      const FrameFramePrecalc& q = fh->targetPrecalc[t];
      float* o = &pc[(size_t)(h * nf + t) * DMV_PRECALC_FLOATS];
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) { o[3 * i + j] = q.PRE_KRKiTll(i, j); o[12 + 3 * i + j] = q.PRE_RTll_0(i, j); }
      for (int i = 0; i < 3; i++) { o[9 + i] = q.PRE_KtTll[i]; o[21 + i] = q.PRE_tTll_0[i]; }
      o[24] = q.PRE_aff_mode[0]; o[25] = q.PRE_aff_mode[1]; o[26] = q.PRE_b0_mode;
    }
  }
  dmv_ba_state st;
  std::memset(&st, 0, sizeof(st));
  for (int i = 0; i < 4; i++) { st.calib[i] = g_calib->value_scaledf[i]; st.calib[4 + i] = g_calib->value_scaledi[i]; }
  st.precalc = pc.data(); st.frameEnergyTH = th.data();
  DMV_CHECK(dmv_ba_set_state(s.ba, &st));
  dmv_ba_lin_result lr;
  DMV_CHECK(dmv_ba_linearize(s.ba, &lr));
  DMV_CHECK(dmv_ba_apply_res(s.ba));
  s.nFrames = nf;
  return s;
}

void EnergyFunctional::accumulateAF_MT(MatXX& H, VecX& b, bool MT) {
  (void)MT;
  BaState& s = ba_sync_and_linearize(this);
  const int N = CPARS + 8 * nFrames, npts = (int)allPoints.size();
  std::vector<double> HA((size_t)N * N), bA(N);
  s.Hsc.assign((size_t)N * N, 0.0); s.bsc.assign(N, 0.0);
  int nres = 0;
  DMV_CHECK(dmv_ba_accumulate(s.ba, HA.data(), bA.data(), s.Hsc.data(), s.bsc.data(), &nres));
  H = MatXX::Zero(N, N);
  b = VecX::Zero(N);
  for (int i = 0; i < N; i++) { b[i] = bA[i]; for (int j = 0; j < N; j++) H(i, j) = HA[(size_t)i * N + j]; }
  resInA = nres;
  // what AccumulatedTopHessianSSE::addPoint / AccumulatedSCHessianSSE::addPoint leave in the EFPoint (read by resubstituteFPt and makeCoarseDepthL0)
  std::vector<float> Hdd(npts), bd(npts), Hcd((size_t)4 * npts), HdiF(npts), bdSum(npts);
  DMV_CHECK(dmv_ba_get_point_outputs(s.ba, Hdd.data(), bd.data(), Hcd.data(), HdiF.data(), bdSum.data()));
  for (int i = 0; i < npts; i++) {
    EFPoint* p = allPoints[i];
    p->Hdd_accAF = Hdd[i]; p->bd_accAF = bd[i];
    for (int k = 0; k < 4; k++) p->Hcd_accAF[k] = Hcd[(size_t)4 * i + k];
    p->HdiF = HdiF[i]; p->bdSumF = bdSum[i];
    p->data->idepth_hessian = HdiF[i] > 0 ? 1.0f / HdiF[i] : 0.f;
  }
}

void EnergyFunctional::accumulateSCF_MT(MatXX& H, VecX& b, bool MT) {
  (void)MT;
  BaState& s = g_ba[this];   // produced by the same launch as the top system (solveSystemF calls accumulateAF_MT first, EnergyFunctional.cpp:L853-860)
  const int N = CPARS + 8 * nFrames;
  H = MatXX::Zero(N, N);
  b = VecX::Zero(N);
  for (int i = 0; i < N; i++) { b[i] = s.bsc[i]; for (int j = 0; j < N; j++) H(i, j) = s.Hsc[(size_t)i * N + j]; }
}

void EnergyFunctional::resubstituteF_MT(VecX x, CalibHessian* HCalib, bool MT) {
  (void)MT;
  BaState& s = g_ba[this];
  const int N = CPARS + 8 * nFrames, npts = (int)allPoints.size();
  HCalib->step = -x.head<CPARS>();
  for (EFFrame* h : frames) {   // frame part as in the reference (L278-279)
    h->data->step.head<8>() = -x.segment<8>(CPARS + 8 * h->idx);
    h->data->step.tail<2>().setZero();
  }
  std::vector<double> xs(N);
  for (int i = 0; i < N; i++) xs[i] = x[i];
  std::vector<float> step(npts);
  double sums[3];
  DMV_CHECK(dmv_ba_resubstitute(s.ba, xs.data(), step.data(), /*apply*/ 0, sums));
  for (int i = 0; i < npts; i++) allPoints[i]->data->step = step[i];   // EnergyFunctional::resubstituteFPt (L295-321)
}

}  // namespace dso
