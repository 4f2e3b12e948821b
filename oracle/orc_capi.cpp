// TEST INFRASTRUCTURE ONLY — C API glue of the CPU oracle.
#include "orc_capi.h"
#include "orc_init.h"
#include "orc_ba.h"
#include "orc_coarse.h"
#include "orc_threads.h"
#include <cstring>
#include <string>

using namespace orc;

struct OrcWin {
  Window W;
  ThreadPool* pool = nullptr;
  ReducedSystem lastSys;
};
struct OrcCT {
  CoarseTracker ct;
  GlobalCalib g;
  std::vector<const float*> lv;
};

static void setSetting(Settings& s, const char* name, double v) {
  std::string n(name);
  if (n == "huberTH") s.huberTH = (float)v;
  else if (n == "outlierTH") s.outlierTH = (float)v;
  else if (n == "outlierTHSumComponent") s.outlierTHSumComponent = (float)v;
  else if (n == "coarseCutoffTH") s.coarseCutoffTH = (float)v;
  else if (n == "affineOptModeA") s.affineOptModeA = (float)v;
  else if (n == "affineOptModeB") s.affineOptModeB = (float)v;
  else if (n == "idepthFixPrior") s.idepthFixPrior = (float)v;
  else if (n == "thOptIterations") s.thOptIterations = (float)v;
  else if (n == "minOptIterations") s.minOptIterations = (int)v;
  else if (n == "initialCalibHessian") s.initialCalibHessian = (float)v;
  else if (n == "orthogonalizeXLater") s.orthogonalizeXLater = v != 0;
}

extern "C" {

OrcWin* orc_win_create(int w, int h, int nf, const double cvs[4], int nthreads) {
  OrcWin* o = new OrcWin();
  Window& W = o->W;
  W.w = w; W.h = h;
  W.frames.resize(nf);
  double v[4] = {cvs[0] / SCALE_F, cvs[1] / SCALE_F, cvs[2] / SCALE_C, cvs[3] / SCALE_C};
  for (int i = 0; i < 4; i++) { W.calib.value_zero[i] = v[i]; W.calib.step[i] = 0; W.calib.value_backup[i] = v[i]; }
  W.calib.setValueScaled(cvs);
  W.nthreads = nthreads;
  if (nthreads > 1) { o->pool = new ThreadPool(nthreads); W.pool = o->pool; }
  return o;
}
void orc_win_destroy(OrcWin* o) {
  if (!o) return;
  delete o->pool;
  delete o;
}
void orc_win_set_setting(OrcWin* o, const char* name, double value) { setSetting(o->W.s, name, value); }

void orc_win_set_frame(OrcWin* o, int idx, const double R[9], const double t[3], const double state[10], const double state_zero[10],
                       float ab_exposure, float frameEnergyTH, int frameID, const float* dI) {
  Frame& f = o->W.frames[idx];
  f.worldToCam_evalPT = SE3::fromRt(R, t);
  for (int i = 0; i < 10; i++) f.state_zero[i] = state_zero[i];
  Vec10 s; for (int i = 0; i < 10; i++) s[i] = state[i];
  f.ab_exposure = ab_exposure;
  f.frameEnergyTH = frameEnergyTH;
  f.frameID = frameID;
  f.dI = dI;
  f.setState(s);
}

void orc_win_set_points(OrcWin* o, int npts, const int32_t* host, const float* u, const float* v, const float* idepth, const float* idepth_zero,
                        const float* color8, const float* weights8, const uint8_t* hasDepthPrior) {
  Window& W = o->W;
  W.points.assign(npts, Point());
  for (int i = 0; i < npts; i++) {
    Point& p = W.points[i];
    p.host = host[i]; p.u = u[i]; p.v = v[i]; p.idepth = idepth[i]; p.idepth_zero = idepth_zero[i]; p.idepth_backup = idepth[i];
    for (int k = 0; k < 8; k++) { p.color[k] = color8[8 * i + k]; p.weights[k] = weights8[8 * i + k]; }
    p.hasDepthPrior = hasDepthPrior ? hasDepthPrior[i] != 0 : false;
  }
}

void orc_win_set_residuals(OrcWin* o, int nres, const int32_t* point, const int32_t* target, const int32_t* state_state, const float* state_energy,
                           const uint8_t* isNew) {
  Window& W = o->W;
  W.residuals.assign(nres, Residual());
  for (Point& p : W.points) p.residuals.clear();
  for (int i = 0; i < nres; i++) {
    Residual& r = W.residuals[i];
    r.point = point[i];
    r.host = W.points[point[i]].host;
    r.target = target[i];
    r.state_state = state_state ? state_state[i] : RS_IN;
    r.state_energy = state_energy ? state_energy[i] : 0;
    r.isNew = isNew ? isNew[i] != 0 : true;
    W.points[point[i]].residuals.push_back(i);
  }
}

void orc_win_set_marg_prior(OrcWin* o, const double* HM, const double* bM) {
  Window& W = o->W;
  const int N = W.nf() * 8 + CPARS;
  W.HM = MatX(N, N);
  W.bM.assign(N, 0.0);
  if (HM) for (int i = 0; i < N * N; i++) W.HM.d[i] = HM[i];
  if (bM) for (int i = 0; i < N; i++) W.bM[i] = bM[i];
}

void orc_win_prepare(OrcWin* o) {
  Window& W = o->W;
  W.setAdjointsF();
  W.takeDataFrames();
  W.setPrecalcValues();
}

int orc_win_nres(OrcWin* o) { return (int)o->W.residuals.size(); }
int orc_win_npts(OrcWin* o) { return (int)o->W.points.size(); }
int orc_win_nf(OrcWin* o) { return o->W.nf(); }

void orc_win_get_precalc(OrcWin* o, float* out) {
  Window& W = o->W;
  const int n = W.nf();
  for (int i = 0; i < n * n; i++) {
    const FramePrecalc& p = W.precalc[i];
    float* q = out + 32 * i;
    for (int k = 0; k < 9; k++) q[k] = p.PRE_KRKiTll.d[k];
    for (int k = 0; k < 3; k++) q[9 + k] = p.PRE_KtTll[k];
    for (int k = 0; k < 9; k++) q[12 + k] = p.PRE_RTll_0.d[k];
    for (int k = 0; k < 3; k++) q[21 + k] = p.PRE_tTll_0[k];
    q[24] = p.PRE_aff_mode[0]; q[25] = p.PRE_aff_mode[1]; q[26] = p.PRE_b0_mode;
    for (int k = 27; k < 32; k++) q[k] = 0;
  }
}
// PRE_RTll (row-major) | PRE_tTll per pair [h*nf+t]: the CURRENT relative pose (FrameFramePrecalc::set, HessianBlocks.cpp:L204-206), read by
// ImmaturePoint::linearizeResidual
void orc_win_get_RT(OrcWin* o, float* out) {
  Window& W = o->W;
  const int n = W.nf();
  for (int i = 0; i < n * n; i++) {
    const FramePrecalc& p = W.precalc[i];
    for (int k = 0; k < 9; k++) out[12 * i + k] = p.PRE_RTll.d[k];
    for (int k = 0; k < 3; k++) out[12 * i + 9 + k] = p.PRE_tTll[k];
  }
}
void orc_win_get_adjoints(OrcWin* o, double* adHost, double* adTarget) {
  Window& W = o->W;
  const int n = W.nf();
  for (int i = 0; i < n * n; i++)
    for (int k = 0; k < 64; k++) { adHost[64 * i + k] = W.adHost[i].d[k]; adTarget[64 * i + k] = W.adTarget[i].d[k]; }
}
void orc_win_get_adHTdeltaF(OrcWin* o, float* out) {
  Window& W = o->W;
  const int n = W.nf();
  for (int i = 0; i < n * n; i++) for (int k = 0; k < 8; k++) out[8 * i + k] = W.adHTdeltaF[i].d[k];
}
void orc_win_get_frame_tables(OrcWin* o, double* prior8, double* delta_prior8, double* delta8, float* frameEnergyTH) {
  Window& W = o->W;
  for (int f = 0; f < W.nf(); f++) {
    for (int k = 0; k < 8; k++) {
      if (prior8) prior8[8 * f + k] = W.frames[f].prior[k];
      if (delta_prior8) delta_prior8[8 * f + k] = W.frames[f].delta_prior[k];
      if (delta8) delta8[8 * f + k] = W.frames[f].delta[k];
    }
    if (frameEnergyTH) frameEnergyTH[f] = W.frames[f].frameEnergyTH;
  }
}
void orc_win_get_calib(OrcWin* o, float* k8, float* cDeltaF4, double* cPrior4) {
  Window& W = o->W;
  if (k8) { for (int i = 0; i < 4; i++) { k8[i] = W.calib.value_scaledf[i]; k8[4 + i] = W.calib.value_scaledi[i]; } }
  if (cDeltaF4) for (int i = 0; i < 4; i++) cDeltaF4[i] = W.cDeltaF[i];
  if (cPrior4) for (int i = 0; i < 4; i++) cPrior4[i] = W.cPrior[i];
}

double orc_win_linearize_all(OrcWin* o, int fix, int updateEnergyTH) {
  std::vector<int> rem;
  return o->W.linearizeAll(fix != 0, &rem, updateEnergyTH != 0);
}
void orc_win_apply_res(OrcWin* o) { o->W.applyResAll(); }

double orc_win_override_new_states(OrcWin* o, const int32_t* newState, int* changed, int* unfixable) {
  Window& W = o->W;
  int nch = 0, nbad = 0;
  double E = 0;
  for (size_t i = 0; i < W.residuals.size(); i++) {
    Residual& r = W.residuals[i];
    if (r.isLinearized || r.dropped) continue;
    const int want = newState[i];
    if (want != r.state_NewState) {
      if (r.state_NewState == RS_OOB) {
        nbad++;  // left through an OOB exit: no Jacobian / energy to reuse
      } else if (want == RS_OOB) {
        r.state_NewState = RS_OOB;
        r.state_NewEnergy = r.state_energy;  // OOB exits return the old energy (Residuals.cpp:L82-83)
        nch++;
      } else {
        const float TH = std::max(W.frames[r.host].frameEnergyTH, W.frames[r.target].frameEnergyTH);
        r.state_NewState = want;
        r.state_NewEnergy = (want == RS_IN) ? r.state_NewEnergyWithOutlier : (double)TH;  // Residuals.cpp:L262-271
        nch++;
      }
    }
    E += (r.state_NewState == RS_OOB) ? r.state_energy : r.state_NewEnergy;  // OOB exits return the old energy and leave state_NewEnergy alone
  }
  if (changed) *changed = nch;
  if (unfixable) *unfixable = nbad;
  return E;
}

void orc_win_get_res_outputs(OrcWin* o, int32_t* newState, float* newEnergy, float* newEnergyWithOutlier, float* cpt3, float* Jnew74,
                             int32_t* state_state, uint8_t* isActive, float* JpJdF8) {
  Window& W = o->W;
  for (size_t i = 0; i < W.residuals.size(); i++) {
    const Residual& r = W.residuals[i];
    if (newState) newState[i] = r.state_NewState;
    if (newEnergy) newEnergy[i] = (float)r.state_NewEnergy;
    if (newEnergyWithOutlier) newEnergyWithOutlier[i] = (float)r.state_NewEnergyWithOutlier;
    if (cpt3) for (int k = 0; k < 3; k++) cpt3[3 * i + k] = r.centerProjectedTo[k];
    if (Jnew74) std::memcpy(Jnew74 + (size_t)RAWJ_FLOATS * i, &r.Jnew, sizeof(RawJ));
    if (state_state) state_state[i] = r.state_state;
    if (isActive) isActive[i] = r.isActiveAndIsGoodNEW ? 1 : 0;
    if (JpJdF8) for (int k = 0; k < 8; k++) JpJdF8[8 * i + k] = r.JpJdF[k];
  }
}

void orc_win_accumulate(OrcWin* o, int precision, double* HA, double* bA, double* HL, double* bL, double* Hsc, double* bsc, int* resInA) {
  ReducedSystem& sys = o->lastSys;
  o->W.accumulate(sys, precision);
  const int N = sys.N;
  auto cp = [&](const MatX& M, double* out) { if (out) std::memcpy(out, M.d.data(), sizeof(double) * N * N); };
  auto cv = [&](const VecX& v, double* out) { if (out) std::memcpy(out, v.data(), sizeof(double) * N); };
  cp(sys.HA, HA); cv(sys.bA, bA); cp(sys.HL, HL); cv(sys.bL, bL); cp(sys.Hsc, Hsc); cv(sys.bsc, bsc);
  if (resInA) *resInA = sys.resInA;
}

// fixLinearization + marginalizePoints on the listed points (FullSystem.cpp:L826-838, EnergyFunctional.cpp:L678-742); rtz8 / isLin are
// indexed like the residual list
void orc_win_marginalize(OrcWin* o, int n, const int32_t* pts, int precision, double* M, double* Mb, double* Msc, double* Mbsc, double* HM,
                         double* bM, int* resInM, int32_t* ngood, float* rtz8, uint8_t* isLin) {
  Window& W = o->W;
  std::vector<int> list(pts, pts + n);
  std::vector<int> ng = W.fixLinearization(list);
  if (ngood) for (int k = 0; k < n; k++) ngood[k] = ng[k];
  for (size_t i = 0; i < W.residuals.size(); i++) {
    const Residual& r = W.residuals[i];
    if (isLin) isLin[i] = r.isLinearized ? 1 : 0;
    if (rtz8) for (int c = 0; c < 8; c++) rtz8[8 * i + c] = r.isLinearized ? r.res_toZeroF[c] : 0.f;
  }
  ReducedSystem sys;
  W.marginalizePoints(list, precision, sys);
  const int N = sys.N;
  auto cp = [&](const MatX& A, double* out) { if (out) std::memcpy(out, A.d.data(), sizeof(double) * N * N); };
  auto cv = [&](const VecX& v, double* out) { if (out) std::memcpy(out, v.data(), sizeof(double) * N); };
  cp(sys.HA, M); cv(sys.bA, Mb); cp(sys.Hsc, Msc); cv(sys.bsc, Mbsc); cp(W.HM, HM); cv(W.bM, bM);
  if (resInM) *resInM = sys.resInA;
}

void orc_win_get_point_outputs(OrcWin* o, float* Hdd, float* bd, float* Hcd4, float* HdiF, float* bdSumF, float* step, float* idepth,
                               float* maxRelBaseline) {
  Window& W = o->W;
  for (size_t i = 0; i < W.points.size(); i++) {
    const Point& p = W.points[i];
    if (Hdd) Hdd[i] = p.Hdd_accAF;
    if (bd) bd[i] = p.bd_accAF;
    if (Hcd4) for (int k = 0; k < 4; k++) Hcd4[4 * i + k] = p.Hcd_accAF[k];
    if (HdiF) HdiF[i] = p.HdiF;
    if (bdSumF) bdSumF[i] = p.bdSumF;
    if (step) step[i] = p.step;
    if (idepth) idepth[i] = p.idepth;
    if (maxRelBaseline) maxRelBaseline[i] = p.maxRelBaseline;
  }
}

void orc_win_solve(OrcWin* o, int iteration, double lambda, int precision, double* x_out, double* HFinal, double* bFinal) {
  MatX HF; VecX bF;
  o->W.solveSystem(iteration, lambda, precision, &o->lastSys, &HF, &bF);
  const int N = o->lastSys.N;
  if (x_out) std::memcpy(x_out, o->W.lastX.data(), sizeof(double) * N);
  if (HFinal) std::memcpy(HFinal, HF.d.data(), sizeof(double) * N * N);
  if (bFinal) std::memcpy(bFinal, bF.data(), sizeof(double) * N);
}
void orc_win_resubstitute(OrcWin* o, const double* x) {
  const int N = o->W.nf() * 8 + CPARS;
  VecX xv(x, x + N);
  o->W.resubstitute(xv);
}
double orc_win_calc_LEnergy(OrcWin* o) { return o->W.calcLEnergy(); }
double orc_win_calc_MEnergy(OrcWin* o) { return o->W.calcMEnergy(); }
int orc_win_optimize(OrcWin* o, int mnumOptIts, int precision, double* energyLog, int cap) {
  std::vector<double> log;
  int n = o->W.optimize(mnumOptIts, precision, &log);
  for (int i = 0; i < cap && i < (int)log.size(); i++) energyLog[i] = log[i];
  for (int i = (int)log.size(); i < cap; i++) energyLog[i] = -1;
  return n;
}
// tail of FullSystem::optimize (FullSystemOptimize.cpp:L591-609); returns the energy, writes up to cap removed residual indices
double orc_win_finish_optimize(OrcWin* o, int32_t* removed, int cap, int* nremoved) {
  std::vector<int> rem;
  const double E = o->W.finishOptimize(&rem);
  for (int i = 0; i < cap && i < (int)rem.size(); i++) removed[i] = rem[i];
  if (nremoved) *nremoved = (int)rem.size();
  return E;
}
void orc_win_get_point_stats(OrcWin* o, float* maxRelBaseline, int32_t* numGoodResiduals) {
  for (size_t i = 0; i < o->W.points.size(); i++) {
    if (maxRelBaseline) maxRelBaseline[i] = o->W.points[i].maxRelBaseline;
    if (numGoodResiduals) numGoodResiduals[i] = o->W.points[i].numGoodResiduals;
  }
}
// EnergyFunctional::marginalizeFrame on the given prior (HM_in: odim*odim, bM_in: odim); outputs ndim = odim - 8.  The window must hold no
// points hosted in / residuals targeting the frame (use a frames-only window).
void orc_win_marginalize_frame(OrcWin* o, int idx, const double* HM_in, const double* bM_in, double* HM_out, double* bM_out) {
  Window& W = o->W;
  const int odim = W.nf() * 8 + CPARS, ndim = odim - 8;
  W.HM = MatX(odim, odim);
  W.bM.assign(odim, 0.0);
  std::memcpy(W.HM.d.data(), HM_in, sizeof(double) * odim * odim);
  std::memcpy(W.bM.data(), bM_in, sizeof(double) * odim);
  W.marginalizeFrame(idx);
  std::memcpy(HM_out, W.HM.d.data(), sizeof(double) * ndim * ndim);
  std::memcpy(bM_out, W.bM.data(), sizeof(double) * ndim);
}
// getNullspaces: out[7][N]; orthogonalize: x (N) in place
void orc_win_get_nullspaces(OrcWin* o, double* out) {
  std::vector<VecX> ns = o->W.getNullspaces();
  const int N = 8 * o->W.nf() + CPARS;
  for (int a = 0; a < 7; a++) std::memcpy(out + (size_t)a * N, ns[a].data(), sizeof(double) * N);
}
void orc_win_orthogonalize(OrcWin* o, double* x) {
  const int N = 8 * o->W.nf() + CPARS;
  VecX v(x, x + N);
  o->W.orthogonalize(v);
  std::memcpy(x, v.data(), sizeof(double) * N);
}
void orc_win_get_frame_states(OrcWin* o, double* state10) {
  for (int f = 0; f < o->W.nf(); f++) for (int k = 0; k < 10; k++) state10[10 * f + k] = o->W.frames[f].state[k];
}
double orc_win_gn_iteration(OrcWin* o, double lambda, int precision, int do_step) {
  Window& W = o->W;
  W.backupState();
  W.solveSystem(0, lambda, precision);
  if (do_step) W.doStepFromBackup();
  double e = W.linearizeAll(false, nullptr);
  if (do_step) W.loadStateBackup();  // keep the window stationary so that the bench can repeat the iteration
  W.applyResAll();
  return e;
}

double orc_win_hot_iteration(OrcWin* o, const double* x, int precision) {
  // the measured path of one GN iteration WITHOUT the dense host solve: accumulate A/L/SC + stitch,
  // resubstitute(x), doStepFromBackup (incl. setPrecalcValues), linearizeAll, applyRes; the window is restored afterwards.
  Window& W = o->W;
  const int N = W.nf() * 8 + CPARS;
  W.backupState();
  W.accumulate(o->lastSys, precision);
  VecX xv(x, x + N);
  W.resubstitute(xv);
  W.doStepFromBackup();
  double e = W.linearizeAll(false, nullptr);
  W.applyResAll();
  W.loadStateBackup();
  return e;
}

int orc_win_eval_raw_double(OrcWin* o, int ri, const double dsh[8], const double dst[8], double didepth, const double dcalib[4], double r_raw[8]) {
  // independent first-principles evaluation in double: r_i = I_t(pi(K T_th pi^-1(p_i, idepth))) - (a * color_i + b)
  Window& W = o->W;
  const Residual& r = W.residuals[ri];
  const Point& p = W.points[r.point];
  Frame fh = W.frames[r.host], ft = W.frames[r.target];
  Vec10 sh = fh.state, st = ft.state;
  for (int i = 0; i < 8; i++) { sh[i] += dsh[i]; st[i] += dst[i]; }
  fh.setState(sh);
  ft.setState(st);
  double cv[4];
  for (int i = 0; i < 4; i++) cv[i] = W.calib.value[i] + dcalib[i];
  double fx = cv[0] * SCALE_F, fy = cv[1] * SCALE_F, cx = cv[2] * SCALE_C, cy = cv[3] * SCALE_C;
  SE3 Tth = ft.PRE_worldToCam * fh.PRE_camToWorld;
  Mat33 R = Tth.rotationMatrix();
  Vec3 t = Tth.translation();
  double aff[2];
  AffLight::fromToVecExposure(fh.ab_exposure, ft.ab_exposure, fh.aff_g2l(), ft.aff_g2l(), aff);
  double id = (double)p.idepth + didepth;
  for (int k = 0; k < 8; k++) {
    double x = ((double)p.u + patternP[k][0] - cx) / fx, y = ((double)p.v + patternP[k][1] - cy) / fy;
    double q[3];
    for (int i = 0; i < 3; i++) q[i] = R(i, 0) * x + R(i, 1) * y + R(i, 2) + t[i] * id;
    double Ku = fx * q[0] / q[2] + cx, Kv = fy * q[1] / q[2] + cy;
    if (!(Ku > 1.1 && Kv > 1.1 && Ku < W.w - 3 && Kv < W.h - 3)) return 0;
    int ix = (int)Ku, iy = (int)Kv;
    double dx = Ku - ix, dy = Kv - iy;
    const float* bp = ft.dI + 3 * (ix + iy * W.w);
    double I = dx * dy * bp[3 * (1 + W.w)] + (dy - dx * dy) * bp[3 * W.w] + (dx - dx * dy) * bp[3] + (1 - dx - dy + dx * dy) * bp[0];
    r_raw[k] = I - (aff[0] * (double)p.color[k] + aff[1]);
  }
  return 1;
}

/* ---- images ---- */
int orc_pyr_levels(int w, int h, int forceLevels) {
  GlobalCalib g; g.set(w, h, 1, 1, 0, 0, forceLevels);
  return g.pyrLevelsUsed;
}
int64_t orc_make_images(int w, int h, int levels, float fx, float fy, float cx, float cy, const float* color, float* out, float* absOut) {
  GlobalCalib g; g.set(w, h, fx, fy, cx, cy, levels);
  float* lv[PYR_LEVELS]; float* av[PYR_LEVELS];
  int64_t off = 0, aoff = 0;
  for (int l = 0; l < g.pyrLevelsUsed; l++) {
    lv[l] = out + off; off += (int64_t)g.wG[l] * g.hG[l] * 3;
    av[l] = absOut ? absOut + aoff : nullptr; aoff += (int64_t)g.wG[l] * g.hG[l];
    if (av[l]) std::memset(av[l], 0, sizeof(float) * g.wG[l] * g.hG[l]);
  }
  makeImages(g, color, lv, absOut ? av : nullptr);
  return off;
}
int orc_init_point(const float* dI, int w, float u, float v, float* color8, float* weights8) {
  return initPointColorWeights(dI, w, u, v, 50 * 50, color8, weights8) ? 1 : 0;
}

/* ---- coarse tracker ---- */
OrcCT* orc_ct_create(int w, int h, int levels, float fx, float fy, float cx, float cy) {
  OrcCT* o = new OrcCT();
  o->g.set(w, h, fx, fy, cx, cy, levels);
  o->ct.makeK(o->g);
  for (int l = 0; l < PYR_LEVELS; l++) { o->ct.pc_n[l] = 0; o->ct.newFrame_dIp[l] = nullptr; }
  return o;
}
void orc_ct_destroy(OrcCT* o) { delete o; }
void orc_ct_set_setting(OrcCT* o, const char* name, double v) { setSetting(o->ct.s, name, v); }
void orc_ct_set_ref_points(OrcCT* o, int lvl, int n, const float* u, const float* v, const float* idepth, const float* color) {
  CoarseTracker& c = o->ct;
  c.pc_u[lvl].assign(u, u + n); c.pc_v[lvl].assign(v, v + n); c.pc_idepth[lvl].assign(idepth, idepth + n); c.pc_color[lvl].assign(color, color + n);
  c.pc_n[lvl] = n;
}
int orc_ct_make_coarse_depth(OrcCT* o, int n, const float* Ku, const float* Kv, const float* nid, const float* HdiF, const float* ref) {
  const float* lv[PYR_LEVELS];
  int64_t off = 0;
  for (int l = 0; l < o->ct.levels; l++) { lv[l] = ref + off; off += (int64_t)o->ct.w[l] * o->ct.h[l] * 3; }
  o->ct.makeCoarseDepthL0(n, Ku, Kv, nid, HdiF, lv);
  int tot = 0;
  for (int l = 0; l < o->ct.levels; l++) tot += o->ct.pc_n[l];
  return tot;
}
int orc_ct_get_ref_points(OrcCT* o, int lvl, float* u, float* v, float* idepth, float* color) {
  CoarseTracker& c = o->ct;
  int n = c.pc_n[lvl];
  if (u) std::memcpy(u, c.pc_u[lvl].data(), 4 * n);
  if (v) std::memcpy(v, c.pc_v[lvl].data(), 4 * n);
  if (idepth) std::memcpy(idepth, c.pc_idepth[lvl].data(), 4 * n);
  if (color) std::memcpy(color, c.pc_color[lvl].data(), 4 * n);
  return n;
}
void orc_ct_set_new_frame(OrcCT* o, const float* dIp, float refExp, float newExp, double ref_a, double ref_b) {
  int64_t off = 0;
  for (int l = 0; l < o->ct.levels; l++) { o->ct.newFrame_dIp[l] = dIp + off; off += (int64_t)o->ct.w[l] * o->ct.h[l] * 3; }
  o->ct.lastRef_ab_exposure = refExp;
  o->ct.newFrame_ab_exposure = newExp;
  o->ct.lastRef_aff_g2l.a = ref_a;
  o->ct.lastRef_aff_g2l.b = ref_b;
}
void orc_ct_get_K(OrcCT* o, int lvl, float* k4, int* wh) {
  k4[0] = o->ct.fx[lvl]; k4[1] = o->ct.fy[lvl]; k4[2] = o->ct.cx[lvl]; k4[3] = o->ct.cy[lvl];
  wh[0] = o->ct.w[lvl]; wh[1] = o->ct.h[lvl];
}
void orc_ct_calc_res(OrcCT* o, int lvl, const double R[9], const double t[3], double a, double b, float cutoffTH, double out6[6]) {
  AffLight aff; aff.a = a; aff.b = b;
  o->ct.calcRes(lvl, SE3::fromRt(R, t), aff, cutoffTH, out6);
}
int orc_ct_get_warped(OrcCT* o, float* buf) {
  CoarseTracker& c = o->ct;
  int n = c.buf_warped_n;
  if (buf) {
    const std::vector<float>* v[8] = {&c.buf_warped_idepth, &c.buf_warped_u, &c.buf_warped_v, &c.buf_warped_dx, &c.buf_warped_dy,
                                      &c.buf_warped_residual, &c.buf_warped_weight, &c.buf_warped_refColor};
    for (int k = 0; k < 8; k++) std::memcpy(buf + (size_t)k * n, v[k]->data(), 4 * (size_t)n);
  }
  return n;
}
void orc_ct_calc_gs(OrcCT* o, int lvl, double a, double b, int precision, double H64[64], double b8[8]) {
  AffLight aff; aff.a = a; aff.b = b;
  Mat88 H; Vec8 bb;
  SE3 dummy;
  o->ct.calcGSSSE(lvl, H, bb, dummy, aff, precision);
  std::memcpy(H64, H.d, sizeof(double) * 64);
  std::memcpy(b8, bb.d, sizeof(double) * 8);
}
int orc_ct_track(OrcCT* o, double R[9], double t[3], double* a, double* b, int coarsestLvl, const double minRes[5], int precision,
                 double lastRes[5], double flow3[3], int* iterations) {
  SE3 T = SE3::fromRt(R, t);
  AffLight aff; aff.a = *a; aff.b = *b;
  bool good = o->ct.trackNewestCoarse(T, aff, coarsestLvl, minRes, precision, iterations);
  Mat33 Rm = T.rotationMatrix();
  for (int i = 0; i < 9; i++) R[i] = Rm.d[i];
  for (int i = 0; i < 3; i++) t[i] = T.t[i];
  *a = aff.a; *b = aff.b;
  for (int i = 0; i < 5; i++) lastRes[i] = o->ct.lastResiduals[i];
  for (int i = 0; i < 3; i++) flow3[i] = o->ct.lastFlowIndicators[i];
  return good ? 1 : 0;
}

void orc_se3_exp(const double xi[6], double R[9], double t[3]) {
  Vec6 v; for (int i = 0; i < 6; i++) v[i] = xi[i];
  SE3 T = SE3::exp(v);
  Mat33 Rm = T.rotationMatrix();
  for (int i = 0; i < 9; i++) R[i] = Rm.d[i];
  for (int i = 0; i < 3; i++) t[i] = T.t[i];
}
void orc_se3_log(const double R[9], const double t[3], double xi[6]) {
  Vec6 v = SE3::fromRt(R, t).log();
  for (int i = 0; i < 6; i++) xi[i] = v[i];
}
void orc_se3_mul(const double Ra[9], const double ta[3], const double Rb[9], const double tb[3], double R[9], double t[3]) {
  SE3 T = SE3::fromRt(Ra, ta) * SE3::fromRt(Rb, tb);
  Mat33 Rm = T.rotationMatrix();
  for (int i = 0; i < 9; i++) R[i] = Rm.d[i];
  for (int i = 0; i < 3; i++) t[i] = T.t[i];
}
void orc_se3_inv(const double Ra[9], const double ta[3], double R[9], double t[3]) {
  SE3 T = SE3::fromRt(Ra, ta).inverse();
  Mat33 Rm = T.rotationMatrix();
  for (int i = 0; i < 9; i++) R[i] = Rm.d[i];
  for (int i = 0; i < 3; i++) t[i] = T.t[i];
}

}  // extern "C"

/* ---- coarse initialiser (orc_init.h) ---- */
extern "C" {
struct OrcCI {
  orc::CoarseInit ci;
  std::vector<std::vector<float>> first, cur;  // pyramids [lvl][w*h*3]
};
static void ci_load(OrcCI* o, std::vector<std::vector<float>>& dst, const float* concat) {
  dst.resize(o->ci.levels);
  size_t off = 0;
  for (int l = 0; l < o->ci.levels; l++) {
    const size_t n = (size_t)o->ci.w[l] * o->ci.h[l] * 3;
    dst[l].assign(concat + off, concat + off + n);
    off += n;
  }
}
OrcCI* orc_ci_create(int w, int h, const double K[4]) {
  OrcCI* o = new OrcCI();
  o->ci.makeK(w, h, K[0], K[1], K[2], K[3]);
  return o;
}
void orc_ci_destroy(OrcCI* o) { delete o; }
int orc_ci_levels(OrcCI* o) { return o->ci.levels; }
// points of all levels concatenated: n[lvl]; u, v, type (floats); parent (index in lvl+1 or -1); neighbours 10 per point
void orc_ci_set_first(OrcCI* o, const float* dIp_concat, float exposure, const int32_t* n, const float* u, const float* v, const float* type,
                      const int32_t* parent, const int32_t* neighbours10) {
  ci_load(o, o->first, dIp_concat);
  size_t off = 0;
  for (int l = 0; l < o->ci.levels; l++) {
    o->ci.points[l].assign(n[l], orc::InitPnt());
    for (int i = 0; i < n[l]; i++, off++) {
      orc::InitPnt& p = o->ci.points[l][i];
      p.u = u[off]; p.v = v[off]; p.my_type = type[off]; p.parent = parent[off];
      for (int k = 0; k < 10; k++) p.neighbours[k] = neighbours10[10 * off + k];
    }
  }
  const float* lv[orc::PYR_LEVELS];
  for (int l = 0; l < o->ci.levels; l++) lv[l] = o->first[l].data();
  o->ci.setFirst(lv, exposure);
}
void orc_ci_set_new(OrcCI* o, const float* dIp_concat, float exposure) {
  ci_load(o, o->cur, dIp_concat);
  for (int l = 0; l < o->ci.levels; l++) o->ci.dINew[l] = o->cur[l].data();
  o->ci.new_exposure = exposure;
}
void orc_ci_calc(OrcCI* o, int lvl, const double R[9], const double t[3], double a, double b, float* H64, float* b8, float* Hsc64, float* bsc8, float* res3) {
  orc::InitSystem S;
  orc::AffLight aff; aff.a = a; aff.b = b;
  o->ci.calcResAndGS(lvl, S, orc::SE3::fromRt(R, t), aff, res3);
  std::memcpy(H64, S.H, sizeof(S.H)); std::memcpy(b8, S.b, sizeof(S.b)); std::memcpy(Hsc64, S.Hsc, sizeof(S.Hsc)); std::memcpy(bsc8, S.bsc, sizeof(S.bsc));
}
void orc_ci_apply_step(OrcCI* o, int lvl) { o->ci.applyStep(lvl); }
void orc_ci_do_step(OrcCI* o, int lvl, float lambda, const float* inc8) { o->ci.doStep(lvl, lambda, inc8); }
void orc_ci_calc_ec(OrcCI* o, int lvl, float* out3) { o->ci.calcEC(lvl, out3); }
void orc_ci_opt_reg(OrcCI* o, int lvl) { o->ci.optReg(lvl); }
void orc_ci_propagate_up(OrcCI* o, int lvl) { o->ci.propagateUp(lvl); }
void orc_ci_propagate_down(OrcCI* o, int lvl) { o->ci.propagateDown(lvl); }
void orc_ci_reset_points(OrcCI* o, int lvl) { o->ci.resetPoints(lvl); }
void orc_ci_set_snapped(OrcCI* o, int snapped) { o->ci.snapped = snapped != 0; }
int orc_ci_npts(OrcCI* o, int lvl) { return (int)o->ci.points[lvl].size(); }
// per-point state, 12 floats: idepth idepth_new iR energy0 energy1 energy_new0 energy_new1 lastHessian lastHessian_new maxstep isGood isGood_new
void orc_ci_get_points(OrcCI* o, int lvl, float* out12) {
  for (size_t i = 0; i < o->ci.points[lvl].size(); i++) {
    const orc::InitPnt& p = o->ci.points[lvl][i];
    float* q = out12 + 12 * i;
    q[0] = p.idepth; q[1] = p.idepth_new; q[2] = p.iR; q[3] = p.energy[0]; q[4] = p.energy[1]; q[5] = p.energy_new[0]; q[6] = p.energy_new[1];
    q[7] = p.lastHessian; q[8] = p.lastHessian_new; q[9] = p.maxstep; q[10] = p.isGood ? 1.f : 0.f; q[11] = p.isGood_new ? 1.f : 0.f;
  }
}
// constant per-point fields u, v, outlierTH and the JbBuffer_new rows of the last calcResAndGS (10 per point): inputs / expected outputs of the
// CUDA kernel's parity test (tests/test_gpu_init.py)
void orc_ci_get_static(OrcCI* o, int lvl, float* u, float* v, float* outlierTH) {
  for (size_t i = 0; i < o->ci.points[lvl].size(); i++) {
    const orc::InitPnt& p = o->ci.points[lvl][i];
    u[i] = p.u; v[i] = p.v; outlierTH[i] = p.outlierTH;
  }
}
void orc_ci_get_K(OrcCI* o, int lvl, double* k4, int* wh) {
  k4[0] = o->ci.fx[lvl]; k4[1] = o->ci.fy[lvl]; k4[2] = o->ci.cx[lvl]; k4[3] = o->ci.cy[lvl];
  wh[0] = o->ci.w[lvl]; wh[1] = o->ci.h[lvl];
}
void orc_ci_get_jb(OrcCI* o, int lvl, float* out10) {
  const size_t n = o->ci.points[lvl].size();
  for (size_t i = 0; i < n && i < o->ci.JbBuffer_new.size(); i++)
    for (int k = 0; k < 10; k++) out10[10 * i + k] = o->ci.JbBuffer_new[i][k];
}
// sets idepth, idepth_new, iR, lastHessian, isGood (5 floats per point) — lets the tests start from arbitrary states
void orc_ci_set_points(OrcCI* o, int lvl, const float* in5) {
  for (size_t i = 0; i < o->ci.points[lvl].size(); i++) {
    orc::InitPnt& p = o->ci.points[lvl][i];
    const float* q = in5 + 5 * i;
    p.idepth = q[0]; p.idepth_new = q[1]; p.iR = q[2]; p.lastHessian = q[3]; p.isGood = q[4] != 0.f;
  }
}
// CoarseInitializer::trackFrame; out: R[9] t[3] a b, state[3] = snapped, snappedAt, frameID; returns its bool
int orc_ci_track(OrcCI* o, const float* dIp_concat, float exposure, double* R9, double* t3, double* ab2, int32_t* state3) {
  ci_load(o, o->cur, dIp_concat);
  const float* lv[orc::PYR_LEVELS];
  for (int l = 0; l < o->ci.levels; l++) lv[l] = o->cur[l].data();
  const bool ok = o->ci.trackFrame(lv, exposure);
  const orc::Mat33 Rm = o->ci.thisToNext.rotationMatrix();
  for (int i = 0; i < 9; i++) R9[i] = Rm.d[i];
  for (int i = 0; i < 3; i++) t3[i] = o->ci.thisToNext.t[i];
  ab2[0] = o->ci.thisToNext_aff.a; ab2[1] = o->ci.thisToNext_aff.b;
  state3[0] = o->ci.snapped ? 1 : 0; state3[1] = o->ci.snappedAt; state3[2] = o->ci.frameID;
  return ok ? 1 : 0;
}
}  // extern "C"

