#include "host_capi.h"
#include <string>
#include "coarse_tracker.h"
#include "coarse_initializer.h"
#include "window_ba.h"
#include "marg_frame.h"
#include "nullspace.h"
#include <cstring>

using namespace dmvio_b200;

static SE3 mkSE3(const double R[9], const double t[3]) {
  SE3 T;
  for (int i = 0; i < 9; i++) T.R[i] = R[i];
  for (int i = 0; i < 3; i++) T.t[i] = t[i];
  return T;
}
static void initCalib(CalibHessian& C, const double vs[4]) {
  const double v[4] = {vs[0] / SCALE_F, vs[1] / SCALE_F, vs[2] / SCALE_C, vs[3] / SCALE_C};
  for (int i = 0; i < 4; i++) { C.value_zero[i] = v[i]; C.step[i] = 0; C.value_backup[i] = v[i]; }
  C.setValueScaled(vs);
}

extern "C" {

void* dmvh_window_create(int w, int h, int max_frames, int max_points, int device, const double cvs[4]) {
  WindowBA* W = new WindowBA(w, h, max_frames, max_points, device);
  initCalib(W->Hcalib, cvs);
  return W;
}
void dmvh_window_destroy(void* p) { delete static_cast<WindowBA*>(p); }
const char* dmvh_window_error(void* p) { return static_cast<WindowBA*>(p)->error().c_str(); }

int dmvh_window_add_frame(void* p, const float* data, int is_image, const double R[9], const double t[3], const double state[10],
                          const double state_zero[10], float ab_exposure, int frameID) {
  WindowBA* W = static_cast<WindowBA*>(p);
  if (!W->ok()) return -1;
  const SE3 T = mkSE3(R, t);
  return is_image ? W->insertFrame(data, T, state, state_zero, ab_exposure, frameID) : W->insertFrameDI(data, T, state, state_zero, ab_exposure, frameID);
}
void dmvh_window_drop_frame(void* p, int idx) { static_cast<WindowBA*>(p)->dropFrame(idx); }
int dmvh_window_marginalize_points(void* p, int nmarg, const int32_t* marg, int ndrop, const int32_t* drop, double* HM, double* bM, int* npts_left,
                                   int* nres_left) {
  WindowBA* W = static_cast<WindowBA*>(p);
  const int rc = W->marginalizePointsF(std::vector<int>(marg, marg + nmarg), std::vector<int>(drop, drop + ndrop));
  if (HM) std::memcpy(HM, W->HM.data(), sizeof(double) * W->HM.size());
  if (bM) std::memcpy(bM, W->bM.data(), sizeof(double) * W->bM.size());
  if (npts_left) *npts_left = (int)W->points.size();
  if (nres_left) *nres_left = (int)W->activeResiduals.size();
  return rc;
}
int dmvh_window_set_points(void* p, int n, const int32_t* host, const float* u, const float* v, const float* idepth, const float* idepth_zero,
                           const float* color8, const float* weights8, const uint8_t* hdp) {
  static_cast<WindowBA*>(p)->insertPoints(n, host, u, v, idepth, idepth_zero, color8, weights8, hdp);
  return 0;
}
int dmvh_window_set_points_carry(void* p, int n, const int32_t* host, const float* u, const float* v, const float* idepth, const float* idepth_zero,
                                 const float* color8, const float* weights8, const uint8_t* hdp, const int32_t* carry_from) {
  static_cast<WindowBA*>(p)->insertPoints(n, host, u, v, idepth, idepth_zero, color8, weights8, hdp, carry_from);
  return 0;
}
int dmvh_window_set_residuals(void* p, int n, const int32_t* point, const int32_t* target) {
  static_cast<WindowBA*>(p)->insertResiduals(n, point, target);
  return 0;
}
void dmvh_window_set_ba_update_hook(void* p, dmvh_ba_update_cb cb, void* user) {
  WindowBA* W = static_cast<WindowBA*>(p);
  if (!cb) { W->computeBAUpdate = nullptr; return; }
  W->computeBAUpdate = [cb, user](const std::vector<double>& H, const std::vector<double>& b, double lambda, int nFrames, const std::vector<double>& HNoLambda) {
    std::vector<double> x(b.size());
    cb(H.data(), b.data(), lambda, nFrames, HNoLambda.data(), x.data(), user);
    return x;
  };
}
int dmvh_window_set_sharding(void* p, int rank, int nranks, dmvh_allgather_cb cb, void* user) {
  WindowBA* W = static_cast<WindowBA*>(p);
  if (cb) W->allgather = [cb, user](const void* send, void* recv, size_t bytes) { cb(send, recv, bytes, user); };
  else W->allgather = nullptr;
  return W->setSharding(rank, nranks) ? 0 : -1;
}
int dmvh_window_p2p_setup(void* p) { return static_cast<WindowBA*>(p)->p2pSetup() ? 0 : -1; }
int dmvh_window_comm_init(void* p, const void* uid) { return static_cast<WindowBA*>(p)->commInit(uid) ? 0 : -1; }
int dmvh_window_get_idepths(void* p, float* idepth) {
  WindowBA* W = static_cast<WindowBA*>(p);
  W->getIdepths(idepth);
  return W->error().empty() ? 0 : -1;
}
void dmvh_window_profile(void* p, double out5[5], int reset) {
  WindowBA* W = static_cast<WindowBA*>(p);
  for (int i = 0; i < 5; i++) { out5[i] = W->profile_us[i]; if (reset) W->profile_us[i] = 0; }
}
int dmvh_window_set_setting(void* p, const char* name, double value) {
  Settings& s = static_cast<WindowBA*>(p)->s;
  const std::string n(name);
  if (n == "minOptIterations") s.setting_minOptIterations = (int)value;
  else if (n == "thOptIterations") s.setting_thOptIterations = (float)value;
  else if (n == "margWeightFac") s.setting_margWeightFac = (float)value;
  else if (n == "huberTH") s.setting_huberTH = (float)value;
  else if (n == "minIdepthH_marg") s.setting_minIdepthH_marg = (float)value;
  else return -1;
  return 0;
}
int dmvh_window_prepare(void* p) {
  WindowBA* W = static_cast<WindowBA*>(p);
  if (!W->makeIDX()) return -1;
  W->setAdjointsF();
  W->setPrecalcValues();
  return W->error().empty() ? 0 : -1;
}
double dmvh_window_linearize(void* p, int fix) { return static_cast<WindowBA*>(p)->linearizeAll(fix != 0); }
void dmvh_window_apply(void* p) { static_cast<WindowBA*>(p)->applyRes_Reductor(); }
int dmvh_window_solve(void* p, int iteration, double lambda, double* x) {
  WindowBA* W = static_cast<WindowBA*>(p);
  W->solveSystemF(iteration, lambda);
  if (x) std::memcpy(x, W->lastX.data(), sizeof(double) * W->lastX.size());
  return (int)W->lastX.size();
}
int dmvh_window_marginalize_frame(void* p, int idx, double* HM, double* bM, int* nf_left, int* nres_left) {
  WindowBA* W = static_cast<WindowBA*>(p);
  const bool ok = W->marginalizeFrame(idx);
  if (HM) std::memcpy(HM, W->HM.data(), sizeof(double) * W->HM.size());
  if (bM) std::memcpy(bM, W->bM.data(), sizeof(double) * W->bM.size());
  if (nf_left) *nf_left = (int)W->frameHessians.size();
  if (nres_left) *nres_left = (int)W->activeResiduals.size();
  return ok ? 0 : -1;
}
void dmvh_nullspaces_orthogonalize(int nf, const double* evalPT12, double* ns_out, double* x, double solverModeDelta) {
  std::vector<SE3> T(nf);
  for (int f = 0; f < nf; f++) T[f] = mkSE3(evalPT12 + 12 * f, evalPT12 + 12 * f + 9);
  const std::vector<std::vector<double>> ns = windowNullspaces(T, SCALE_XI_TRANS, SCALE_XI_ROT);
  const int N = 8 * nf + 4;
  if (ns_out) for (int a = 0; a < 7; a++) std::memcpy(ns_out + (size_t)a * N, ns[a].data(), sizeof(double) * N);
  if (x) {
    std::vector<double> v(x, x + N);
    orthogonalizeX(v, ns, solverModeDelta);
    std::memcpy(x, v.data(), sizeof(double) * N);
  }
}
void dmvh_marginalize_frame_hm(double* HM, double* bM, int nFrames, int idx, const double prior8[8], const double delta_prior8[8]) {
  const int odim = 8 * nFrames + 4;
  std::vector<double> H(HM, HM + (size_t)odim * odim), b(bM, bM + odim);
  marginalizeFrameHM(H, b, nFrames, idx, prior8, delta_prior8);
  std::memcpy(HM, H.data(), sizeof(double) * H.size());
  std::memcpy(bM, b.data(), sizeof(double) * b.size());
}
double dmvh_window_finish_optimize(void* p, int32_t* removed, int cap, int* nremoved, int* nres_left) {
  WindowBA* W = static_cast<WindowBA*>(p);
  const double E = W->finishOptimize();
  const std::vector<int>& r = W->lastRemovedResiduals;
  for (int i = 0; i < cap && i < (int)r.size(); i++) removed[i] = r[i];
  if (nremoved) *nremoved = (int)r.size();
  if (nres_left) *nres_left = (int)W->activeResiduals.size();
  return E;
}
void dmvh_window_get_point_stats(void* p, float* maxRelBaseline, int32_t* numGoodResiduals) {
  WindowBA* W = static_cast<WindowBA*>(p);
  for (size_t i = 0; i < W->points.size(); i++) {
    if (maxRelBaseline) maxRelBaseline[i] = W->points[i].maxRelBaseline;
    if (numGoodResiduals) numGoodResiduals[i] = W->points[i].numGoodResiduals;
  }
}
void dmvh_window_set_last_residuals(void* p, const int32_t* target2, const int32_t* state2) {
  WindowBA* W = static_cast<WindowBA*>(p);
  for (size_t i = 0; i < W->points.size(); i++)
    for (int k = 0; k < 2; k++) { W->points[i].lastResiduals_target[k] = target2[2 * i + k]; W->points[i].lastResiduals_state[k] = state2[2 * i + k]; }
}
void dmvh_window_flag_points(void* p, int nflagged, const int32_t* flagged, int32_t* marg, int* nmarg, int32_t* drop, int* ndrop) {
  WindowBA* W = static_cast<WindowBA*>(p);
  std::vector<int> m, d;
  W->flagPointsForRemoval(std::vector<int>(flagged, flagged + nflagged), &m, &d);
  for (size_t i = 0; i < m.size(); i++) marg[i] = m[i];
  for (size_t i = 0; i < d.size(); i++) drop[i] = d[i];
  *nmarg = (int)m.size(); *ndrop = (int)d.size();
}
int dmvh_window_optimize(void* p, int its, double* log, int cap) {
  std::vector<double> e;
  const int n = static_cast<WindowBA*>(p)->optimize(its, &e, false);
  for (int i = 0; i < cap; i++) log[i] = i < (int)e.size() ? e[i] : -1.0;
  return n;
}
void dmvh_window_get_tables(void* p, float* precalc, double* adH, double* adT) {
  WindowBA* W = static_cast<WindowBA*>(p);
  if (precalc) std::memcpy(precalc, W->precalc.data(), sizeof(float) * W->precalc.size());
  if (adH) std::memcpy(adH, W->adHost.data(), sizeof(double) * W->adHost.size());
  if (adT) std::memcpy(adT, W->adTarget.data(), sizeof(double) * W->adTarget.size());
}
void dmvh_window_get_system(void* p, double* HA, double* bA, double* Hsc, double* bsc, double* HS, double* bS) {
  WindowBA* W = static_cast<WindowBA*>(p);
  auto cp = [](const std::vector<double>& v, double* o) { if (o && !v.empty()) std::memcpy(o, v.data(), sizeof(double) * v.size()); };
  cp(W->last_HA, HA); cp(W->last_bA, bA); cp(W->last_Hsc, Hsc); cp(W->last_bsc, bsc); cp(W->lastHS, HS); cp(W->lastbS, bS);
}
void dmvh_window_get_states(void* p, double* st, float* idepth, float* th) {
  WindowBA* W = static_cast<WindowBA*>(p);
  if (st) W->getFrameStates(st);
  if (idepth) W->getIdepths(idepth);
  if (th) for (size_t f = 0; f < W->frameHessians.size(); f++) th[f] = W->frameHessians[f].frameEnergyTH;
}
double dmvh_window_energy_L(void* p) { return static_cast<WindowBA*>(p)->calcLEnergyF_MT(); }
double dmvh_window_energy_M(void* p) { return static_cast<WindowBA*>(p)->calcMEnergyF(); }

void* dmvh_ct_create(int w, int h, int levels, int max_points, int device, const double cvs[4]) {
  CoarseTracker* C = new CoarseTracker(w, h, levels, max_points, device);
  CalibHessian H;
  initCalib(H, cvs);
  C->makeK(H);
  return C;
}
void dmvh_ct_destroy(void* p) { delete static_cast<CoarseTracker*>(p); }
int dmvh_ct_set_ref(void* p, int n, const float* Ku, const float* Kv, const float* nid, const float* HdiF, const float* ref, double ra, double rb,
                    float rexp) {
  CoarseTracker* C = static_cast<CoarseTracker*>(p);
  if (!C->ok()) return -1;
  const float* lv[DMV_MAX_PYR_LEVELS];
  size_t off = 0;
  for (int l = 0; l < C->levels(); l++) { lv[l] = ref + off; off += (size_t)C->levelPixels(l) * 3; }
  AffLight a; a.a = ra; a.b = rb;
  C->setCoarseTrackingRef(n, Ku, Kv, nid, HdiF, lv, a, rexp);
  return C->error().empty() ? 0 : -1;
}
int dmvh_ct_pc_n(void* p, int lvl) { return static_cast<CoarseTracker*>(p)->pc_n[lvl]; }
int dmvh_ct_set_new_image(void* p, const float* image, float exposure) { return static_cast<CoarseTracker*>(p)->setNewFrame(image, exposure) ? 0 : -1; }
int dmvh_ct_set_ref_device(void* p, int n, const float* Ku, const float* Kv, const float* nid, const float* HdiF, const float* ref_image, double ref_a,
                           double ref_b, float ref_exposure) {
  AffLight a; a.a = ref_a; a.b = ref_b;
  return static_cast<CoarseTracker*>(p)->setCoarseTrackingRefOnDevice(n, Ku, Kv, nid, HdiF, ref_image, a, ref_exposure) ? 0 : -1;
}
void dmvh_ct_set_device_lm(void* p, int on) { static_cast<CoarseTracker*>(p)->useDeviceLM = on != 0; }
double dmvh_ct_point_evaluations(void* p) { return static_cast<CoarseTracker*>(p)->pointEvaluations; }
int dmvh_ct_track(void* p, double R[9], double t[3], double* a, double* b, int coarsestLvl, const double minRes[5], double lastRes[5], double flow[3],
                  int* iterations, long long* evaluations) {
  CoarseTracker* C = static_cast<CoarseTracker*>(p);
  SE3 T = mkSE3(R, t);
  AffLight aff; aff.a = *a; aff.b = *b;
  const bool good = C->trackNewestCoarse(T, aff, coarsestLvl, minRes);
  for (int i = 0; i < 9; i++) R[i] = T.R[i];
  for (int i = 0; i < 3; i++) t[i] = T.t[i];
  *a = aff.a; *b = aff.b;
  for (int i = 0; i < 5; i++) lastRes[i] = C->lastResiduals[i];
  for (int i = 0; i < 3; i++) flow[i] = C->lastFlowIndicators[i];
  if (iterations) *iterations = C->iterations;
  if (evaluations) *evaluations = C->evaluations;
  return good ? 1 : 0;
}


// ---- CoarseInitializer adapter
void* dmvh_ci_create(int w, int h, int levels, int max_points, int device, const double cvs[4]) {
  CoarseInitializer* C = new CoarseInitializer(w, h, levels, max_points, device);
  CalibHessian H;
  initCalib(H, cvs);
  C->makeK(H);
  return C;
}
void dmvh_ci_destroy(void* p) { delete static_cast<CoarseInitializer*>(p); }
const char* dmvh_ci_error(void* p) { return static_cast<CoarseInitializer*>(p)->error().c_str(); }
// points of all levels concatenated (n[l] per level): u, v, my_type, parent, neighbours (10 per point); dIp: the first frame's pyramid, concatenated
int dmvh_ci_set_first(void* p, const float* dIp, float exposure, const int32_t* n, const float* u, const float* v, const float* type, const int32_t* parent,
                      const int32_t* neighbours10) {
  CoarseInitializer* C = static_cast<CoarseInitializer*>(p);
  if (!C->ok()) return -1;
  const float* lv[DMV_MAX_PYR_LEVELS];
  size_t off = 0, q = 0;
  for (int l = 0; l < C->levels(); l++) {
    lv[l] = dIp + off;
    off += (size_t)(C->width(l) * C->height(l)) * 3;
    C->points[l].assign(n[l], Pnt());
    for (int i = 0; i < n[l]; i++, q++) {
      Pnt& pt = C->points[l][i];
      pt.u = u[q]; pt.v = v[q]; pt.my_type = type[q]; pt.parent = parent[q];
      for (int k = 0; k < 10; k++) pt.neighbours[k] = neighbours10[10 * q + k];
    }
  }
  return C->setFirst(lv, exposure) ? 0 : -1;
}
// CoarseInitializer::trackFrame; out: R[9] t[3] (thisToNext), ab[2], state[3] = snapped, snappedAt, frameID; returns its bool (or -1 on error)
int dmvh_ci_track(void* p, const float* dIp, float exposure, double* R9, double* t3, double* ab2, int32_t* state3) {
  CoarseInitializer* C = static_cast<CoarseInitializer*>(p);
  const float* lv[DMV_MAX_PYR_LEVELS];
  size_t off = 0;
  for (int l = 0; l < C->levels(); l++) { lv[l] = dIp + off; off += (size_t)(C->width(l) * C->height(l)) * 3; }
  const bool ok = C->trackFrame(lv, exposure);
  if (!C->error().empty()) return -1;
  for (int i = 0; i < 9; i++) R9[i] = C->thisToNext.R[i];
  for (int i = 0; i < 3; i++) t3[i] = C->thisToNext.t[i];
  ab2[0] = C->thisToNext_aff.a; ab2[1] = C->thisToNext_aff.b;
  state3[0] = C->snapped ? 1 : 0; state3[1] = C->snappedAt; state3[2] = C->frameID;
  return ok ? 1 : 0;
}
int dmvh_ci_npts(void* p, int lvl) { return (int)static_cast<CoarseInitializer*>(p)->points[lvl].size(); }
// per-point state, 12 floats: idepth idepth_new iR energy0 energy1 energy_new0 energy_new1 lastHessian lastHessian_new maxstep isGood isGood_new
void dmvh_ci_get_points(void* p, int lvl, float* out12) {
  const std::vector<Pnt>& pts = static_cast<CoarseInitializer*>(p)->points[lvl];
  for (size_t i = 0; i < pts.size(); i++) {
    const Pnt& q = pts[i];
    float* o = out12 + 12 * i;
    o[0] = q.idepth; o[1] = q.idepth_new; o[2] = q.iR; o[3] = q.energy[0]; o[4] = q.energy[1]; o[5] = q.energy_new[0]; o[6] = q.energy_new[1];
    o[7] = q.lastHessian; o[8] = q.lastHessian_new; o[9] = q.maxstep; o[10] = q.isGood ? 1.f : 0.f; o[11] = q.isGood_new ? 1.f : 0.f;
  }
}
long long dmvh_ci_evaluations(void* p) { return static_cast<CoarseInitializer*>(p)->evaluations; }
}  // extern "C"
