// TEST INFRASTRUCTURE ONLY — never shipped, never linked into libdmvio_b200.so.
//
// A CPU stand-in for the part of the C ABI (include/dmvio_b200.h) that the C++ host adapter dm-vio_b200/host/window_ba.cpp calls,
// implemented on top of the oracle (orc_ba).  oracle/Makefile links it with the UNMODIFIED host adapter sources into
// oracle/libhost_on_oracle.so so that tests/test_host_on_oracle.py can run the adapter's own control flow — FullSystem::optimize's LM
// loop, its tail (linearizeAll(true)), flagPointsForRemoval, marginalizePointsF, marginalizeFrame, solveSystemF with the gauge
// projection — in the CPU test tier, where no GPU exists.  What it checks is the HOST LOGIC (call order, index bookkeeping, table
// plumbing, priors, step handling); the device arithmetic is covered by the `-m gpu` parity tests.  The semantics mirrored here are the
// documented ones of the header: tentative vs committed linearisation, fused resubstitute + point step in dmv_ba_gn_step, ping-pong
// depth backup, residual slots that can be dropped, a marginalisation launch that leaves the committed linearisation alone.
#include "../include/dmvio_b200.h"
#include "orc_ba.h"
#include "orc_coarse.h"
#include "orc_init.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace orc;

// Multi-rank emulation for the CPU tier: the CUDA library sums the linearised system over the ranks inside the launch; here the test
// installs a host all-reduce (sum over ranks of a double buffer, e.g. torch.distributed / gloo) and the stand-in calls it at the same places.
typedef void (*mock_allreduce_cb)(double* buf, int n, void* user);
static mock_allreduce_cb g_allreduce = nullptr;
static void* g_allreduce_user = nullptr;
extern "C" void mock_set_allreduce(mock_allreduce_cb cb, void* user) { g_allreduce = cb; g_allreduce_user = user; }

struct dmv_ba {
  int nranks = 1;
  Window W;
  dmv_ba_config cfg;
  std::vector<std::vector<float>> slot_dI;  // per image slot: level-0 [I, dx, dy] AoS
  std::vector<int> slots;                   // window frame -> slot
  bool have_tentative = false, have_committed = false, have_adj = false, have_state = false;
  ReducedSystem sys;                        // system of the committed linearisation
  std::vector<float> hdi_solve;             // EFPoint::HdiF at the last dmv_ba_accumulate
};
struct dmv_ci {
  CoarseInit ci;
  dmv_ci_config cfg;
  std::vector<std::vector<float>> first, next;   // level planes [I, dx, dy] AoS
};
struct dmv_ct {
  CoarseTracker ct;
  std::vector<std::vector<float>> pyr;  // pyramid of the frame uploaded last
};

static void mock_allreduce(dmv_ba* b, std::vector<double>& v) {
  if (b->nranks > 1 && g_allreduce) g_allreduce(v.data(), (int)v.size(), g_allreduce_user);
}
// energy + counters of a linearisation, summed over the ranks like the counters slot of the CUDA library's result blob
static void mock_reduce_lin(dmv_ba* b, double& E, dmv_ba_lin_result* out, double s3[3]) {
  if (b->nranks <= 1) return;
  std::vector<double> v = {E, out ? (double)out->n_in : 0.0, out ? (double)out->n_oob : 0.0, out ? (double)out->n_outlier : 0.0, s3 ? s3[0] : 0.0, s3 ? s3[1] : 0.0, s3 ? s3[2] : 0.0};
  mock_allreduce(b, v);
  E = v[0];
  if (out) { out->energy = v[0]; out->n_in = (int)v[1]; out->n_oob = (int)v[2]; out->n_outlier = (int)v[3]; }
  if (s3) { s3[0] = v[4]; s3[1] = v[5]; s3[2] = v[6]; }
}
static void mock_reduce_system(dmv_ba* b, ReducedSystem& sys) {
  if (b->nranks <= 1) return;
  const int N = sys.N;
  std::vector<double> v;
  v.insert(v.end(), sys.HA.d.begin(), sys.HA.d.end());
  v.insert(v.end(), sys.bA.begin(), sys.bA.end());
  v.insert(v.end(), sys.Hsc.d.begin(), sys.Hsc.d.end());
  v.insert(v.end(), sys.bsc.begin(), sys.bsc.end());
  v.push_back((double)sys.resInA);
  mock_allreduce(b, v);
  size_t o = 0;
  std::copy(v.begin() + o, v.begin() + o + (size_t)N * N, sys.HA.d.begin()); o += (size_t)N * N;
  std::copy(v.begin() + o, v.begin() + o + N, sys.bA.begin()); o += N;
  std::copy(v.begin() + o, v.begin() + o + (size_t)N * N, sys.Hsc.d.begin()); o += (size_t)N * N;
  std::copy(v.begin() + o, v.begin() + o + N, sys.bsc.begin()); o += N;
  sys.resInA = (int)v[o];
}
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

extern "C" {

const char* dmv_last_error(void) { return g_err.c_str(); }
const char* dmv_version(void) { return "dmvio_b200 host-logic mock over the CPU oracle (test infrastructure)"; }
int dmv_device_count(void) { return 0; }

void dmv_ba_default_params(dmv_ba_params* p) { p->huberTH = 9.f; p->outlierTHSumComponent = 2500.f; p->affineOptModeA = 1e12f; p->affineOptModeB = 1e8f; }

int dmv_ba_create(const dmv_ba_config* cfg, dmv_ba** out) {
  if (!cfg || !out) return fail(DMV_ERR_INVALID, "null argument");
  dmv_ba* b = new dmv_ba();
  b->cfg = *cfg;
  b->W.w = cfg->w; b->W.h = cfg->h;
  b->slot_dI.resize(cfg->max_frames);
  *out = b;
  return DMV_OK;
}
int dmv_ba_destroy(dmv_ba* b) { delete b; return DMV_OK; }
int dmv_ba_set_params(dmv_ba* b, const dmv_ba_params* p) {
  b->W.s.huberTH = p->huberTH; b->W.s.outlierTHSumComponent = p->outlierTHSumComponent;
  b->W.s.affineOptModeA = p->affineOptModeA; b->W.s.affineOptModeB = p->affineOptModeB;
  return DMV_OK;
}
int dmv_ba_upload_frame(dmv_ba* b, int slot, const float* dI) {
  if (slot < 0 || slot >= b->cfg.max_frames) return fail(DMV_ERR_INVALID, "slot out of range");
  b->slot_dI[slot].assign(dI, dI + (size_t)b->cfg.w * b->cfg.h * 3);
  return DMV_OK;
}
int dmv_ba_upload_image(dmv_ba* b, int slot, const float* image) {
  if (slot < 0 || slot >= b->cfg.max_frames) return fail(DMV_ERR_INVALID, "slot out of range");
  GlobalCalib g;
  g.set(b->cfg.w, b->cfg.h, 1.f, 1.f, 0.f, 0.f, 1);
  std::vector<float> lvl0((size_t)b->cfg.w * b->cfg.h * 3);
  float* lv[1] = {lvl0.data()};
  makeImages(g, image, lv, nullptr);
  b->slot_dI[slot] = lvl0;
  return DMV_OK;
}
int dmv_ba_set_window(dmv_ba* b, int nf, const int* slots) {
  if (nf < 2 || nf > b->cfg.max_frames) return fail(DMV_ERR_INVALID, "bad window size");
  b->W.frames.assign(nf, Frame());
  b->slots.resize(nf);
  for (int f = 0; f < nf; f++) b->slots[f] = slots ? slots[f] : f;   // NULL = identity, like the CUDA library
  for (int f = 0; f < nf; f++) b->W.frames[f].dI = b->slot_dI[b->slots[f]].data();
  b->W.points.clear(); b->W.residuals.clear();
  b->have_tentative = b->have_committed = b->have_adj = false;
  return DMV_OK;
}
int dmv_ba_set_points(dmv_ba* b, int npts, const int32_t* host, const float* u, const float* v, const float* idepth, const float* idepth_zero,
                      const float* color8, const float* weights8, const float* priorF) {
  Window& W = b->W;
  W.points.assign(npts, Point());
  b->hdi_solve.clear();
  for (int i = 0; i < npts; i++) {
    Point& p = W.points[i];
    if (i > 0 && host[i] < host[i - 1]) return fail(DMV_ERR_INVALID, "points must be ordered by host frame");
    p.host = host[i]; p.u = u[i]; p.v = v[i]; p.idepth = idepth[i]; p.idepth_zero = idepth_zero ? idepth_zero[i] : idepth[i];
    p.idepth_backup = p.idepth;
    std::memcpy(p.color, color8 + 8 * i, 32); std::memcpy(p.weights, weights8 + 8 * i, 32);
    p.priorF = priorF ? priorF[i] : 0.f;
    p.deltaF = p.idepth - p.idepth_zero;
  }
  W.residuals.clear();
  b->have_tentative = b->have_committed = false;
  return DMV_OK;
}
int dmv_ba_set_residuals(dmv_ba* b, int nres, const int32_t* point, const int32_t* target, const int32_t* st, const float* en) {
  Window& W = b->W;
  W.residuals.assign(nres, Residual());
  for (Point& p : W.points) p.residuals.clear();
  for (int i = 0; i < nres; i++) {
    if (point[i] < 0 || point[i] >= (int)W.points.size() || target[i] < 0 || target[i] >= W.nf()) return fail(DMV_ERR_INVALID, "residual %d out of range", i);
    Residual& r = W.residuals[i];
    r.point = point[i]; r.host = W.points[point[i]].host; r.target = target[i];
    if (r.target == r.host) return fail(DMV_ERR_INVALID, "residual %d targets its own host frame", i);
    r.state_state = st ? st[i] : RS_IN;
    r.state_energy = en ? en[i] : 0;
    W.points[point[i]].residuals.push_back(i);
  }
  b->have_tentative = b->have_committed = false;
  return DMV_OK;
}
int dmv_ba_set_adjoints(dmv_ba* b, const double* adHost, const double* adTarget) {
  Window& W = b->W;
  const int n = W.nf();
  W.adHost.assign((size_t)n * n, Mat88()); W.adTarget.assign((size_t)n * n, Mat88());
  W.adHostF.assign((size_t)n * n, Mat88f()); W.adTargetF.assign((size_t)n * n, Mat88f());
  for (int i = 0; i < n * n; i++)
    for (int r = 0; r < 8; r++)
      for (int c = 0; c < 8; c++) {
        W.adHost[i](r, c) = adHost[(size_t)i * 64 + r * 8 + c]; W.adTarget[i](r, c) = adTarget[(size_t)i * 64 + r * 8 + c];
        W.adHostF[i](r, c) = (float)adHost[(size_t)i * 64 + r * 8 + c]; W.adTargetF[i](r, c) = (float)adTarget[(size_t)i * 64 + r * 8 + c];
      }
  W.adHTdeltaF.assign((size_t)n * n, Mat<float, 1, 8>());
  b->have_adj = true;
  return DMV_OK;
}

static void take_state(dmv_ba* b, const dmv_ba_state* st) {
  Window& W = b->W;
  const int n = W.nf();
  for (int i = 0; i < 4; i++) { W.calib.value_scaledf[i] = st->calib[i]; W.calib.value_scaledi[i] = st->calib[4 + i]; }
  W.precalc.assign((size_t)n * n, FramePrecalc());
  for (int i = 0; i < n * n; i++) {
    const float* q = st->precalc + (size_t)32 * i;
    FramePrecalc& p = W.precalc[i];
    for (int k = 0; k < 9; k++) p.PRE_KRKiTll.d[k] = q[k];
    for (int k = 0; k < 3; k++) p.PRE_KtTll[k] = q[9 + k];
    for (int k = 0; k < 9; k++) p.PRE_RTll_0.d[k] = q[12 + k];
    for (int k = 0; k < 3; k++) p.PRE_tTll_0[k] = q[21 + k];
    p.PRE_aff_mode[0] = q[24]; p.PRE_aff_mode[1] = q[25]; p.PRE_b0_mode = q[26];
  }
  for (int f = 0; f < n; f++) W.frames[f].frameEnergyTH = st->frameEnergyTH[f];
  if (st->idepth) for (size_t i = 0; i < W.points.size(); i++) W.points[i].idepth = st->idepth[i];
  if (st->idepth_zero) for (size_t i = 0; i < W.points.size(); i++) W.points[i].idepth_zero = st->idepth_zero[i];
  for (Point& p : W.points) p.deltaF = p.idepth - p.idepth_zero;
}

int dmv_ba_gn_step(dmv_ba* b, const double* x, const dmv_ba_state* st, dmv_ba_lin_result* out, double sums[3]) {
  if (!b || !st) return fail(DMV_ERR_INVALID, "null argument");
  if (!b->have_adj) return fail(DMV_ERR_STATE, "dmv_ba_set_adjoints first");
  if (x && !b->have_committed) return fail(DMV_ERR_STATE, "no committed linearisation to resubstitute");
  Window& W = b->W;
  double s3[3] = {0, 0, 0};
  if (x) {  // EnergyFunctional::resubstituteF_MT + the point part of doStepFromBackup, on the committed linearisation
    VecX xv(x, x + 8 * W.nf() + CPARS);
    W.resubstitute(xv);
    for (Point& p : W.points) {
      s3[0] += (double)p.step * p.step; s3[1] += std::fabs(p.idepth_backup); s3[2] += 1;
      p.idepth = p.idepth_backup + p.step;
      p.idepth_zero = p.idepth;
    }
  }
  take_state(b, st);
  b->have_state = true;
  const double E = W.linearizeAll(false, nullptr, false);
  if (out) {
    out->energy = E; out->n_in = out->n_oob = out->n_outlier = 0;
    for (const Residual& r : W.residuals) {
      if (r.dropped) continue;
      out->n_in += r.state_NewState == RS_IN; out->n_oob += r.state_NewState == RS_OOB; out->n_outlier += r.state_NewState == RS_OUTLIER;
    }
  }
  {
    double Er = E;
    dmv_ba_lin_result tmp;
    if (!out) { tmp.energy = E; tmp.n_in = tmp.n_oob = tmp.n_outlier = 0; }
    mock_reduce_lin(b, Er, out ? out : &tmp, s3);
  }
  if (sums) { sums[0] = s3[0]; sums[1] = s3[1]; sums[2] = s3[2]; }
  b->have_tentative = true;
  return DMV_OK;
}
int dmv_ba_set_state(dmv_ba* b, const dmv_ba_state* st) {
  if (!b || !st) return fail(DMV_ERR_INVALID, "null argument");
  take_state(b, st);
  b->have_state = true;
  return DMV_OK;
}
int dmv_ba_linearize(dmv_ba* b, dmv_ba_lin_result* out) {
  if (!b) return fail(DMV_ERR_INVALID, "null argument");
  if (!b->have_adj || !b->have_state) return fail(DMV_ERR_STATE, "dmv_ba_set_adjoints + dmv_ba_set_state first");
  Window& W = b->W;
  const double E = W.linearizeAll(false, nullptr, false);
  if (out) {
    out->energy = E; out->n_in = out->n_oob = out->n_outlier = 0;
    for (const Residual& r : W.residuals) {
      if (r.dropped) continue;
      out->n_in += r.state_NewState == RS_IN; out->n_oob += r.state_NewState == RS_OOB; out->n_outlier += r.state_NewState == RS_OUTLIER;
    }
  }
  {
    double Er = E;
    dmv_ba_lin_result tmp;
    tmp.energy = E; tmp.n_in = tmp.n_oob = tmp.n_outlier = 0;
    mock_reduce_lin(b, Er, out ? out : &tmp, nullptr);
  }
  b->have_tentative = true;
  return DMV_OK;
}
int dmv_ba_resubstitute(dmv_ba* b, const double* x, float* step_out, int apply, double sums[3]) {
  if (!b || !x) return fail(DMV_ERR_INVALID, "null argument");
  if (!b->have_committed) return fail(DMV_ERR_STATE, "no committed linearisation to resubstitute");
  Window& W = b->W;
  VecX xv(x, x + 8 * W.nf() + CPARS);
  W.resubstitute(xv);
  double s3[3] = {0, 0, 0};
  for (size_t i = 0; i < W.points.size(); i++) {
    Point& p = W.points[i];
    if (step_out) step_out[i] = p.step;
    s3[0] += (double)p.step * p.step; s3[1] += std::fabs(p.idepth_backup); s3[2] += 1;
    if (apply) { p.idepth = p.idepth_backup + p.step; p.idepth_zero = p.idepth; }
  }
  if (sums) { sums[0] = s3[0]; sums[1] = s3[1]; sums[2] = s3[2]; }
  return DMV_OK;
}
// the CPU stand-in has no ranks: the exchange set-up entry points exist for the adapter's link, and refuse
int dmv_ba_p2p_export(dmv_ba*, void* h64) {
  if (!g_allreduce) return fail(DMV_ERR_STATE, "host-logic mock: no multi-rank exchange (mock_set_allreduce first)");
  std::memset(h64, 0, 64);
  return DMV_OK;
}
int dmv_ba_p2p_import(dmv_ba* b, int nranks, int, const void*) {
  if (!g_allreduce) return fail(DMV_ERR_STATE, "host-logic mock: no multi-rank exchange (mock_set_allreduce first)");
  b->nranks = nranks;
  return DMV_OK;
}
int dmv_ba_comm_init(dmv_ba* b, int nranks, int, const void*) {
  if (!g_allreduce) return fail(DMV_ERR_STATE, "host-logic mock: no multi-rank exchange (mock_set_allreduce first)");
  b->nranks = nranks;
  return DMV_OK;
}

int dmv_ba_apply_res(dmv_ba* b) {
  if (!b->have_tentative) return fail(DMV_ERR_STATE, "no tentative linearisation to commit");
  b->W.applyResAll();
  b->W.accumulate(b->sys, 1);  // the device accumulates while it linearises: the committed system belongs to the committed state
  mock_reduce_system(b, b->sys);
  b->have_committed = true; b->have_tentative = false;
  return DMV_OK;
}
int dmv_ba_accumulate(dmv_ba* b, double* HA, double* bA, double* Hsc, double* bsc, int* resInA) {
  if (!b->have_committed) return fail(DMV_ERR_STATE, "no committed linearisation (linearize + apply_res first)");
  const int N = b->sys.N;
  if (HA) std::memcpy(HA, b->sys.HA.d.data(), sizeof(double) * N * N);
  if (bA) std::memcpy(bA, b->sys.bA.data(), sizeof(double) * N);
  if (Hsc) std::memcpy(Hsc, b->sys.Hsc.d.data(), sizeof(double) * N * N);
  if (bsc) std::memcpy(bsc, b->sys.bsc.data(), sizeof(double) * N);
  if (resInA) *resInA = b->sys.resInA;
  b->hdi_solve.resize(b->W.points.size());  // what AccumulatedSCHessian::addPoint left in EFPoint::HdiF for THIS accumulation
  for (size_t i = 0; i < b->W.points.size(); i++) b->hdi_solve[i] = b->W.points[i].HdiF;
  return DMV_OK;
}
int dmv_ba_get_solve_HdiF(dmv_ba* b, float* HdiF) {
  if (b->hdi_solve.size() != b->W.points.size()) return fail(DMV_ERR_STATE, "no dmv_ba_accumulate since the points were set");
  for (size_t i = 0; i < b->hdi_solve.size(); i++) HdiF[i] = b->hdi_solve[i];
  return DMV_OK;
}
int dmv_ba_get_residual_outputs(dmv_ba* b, int32_t* ns, float* ne, float* nw, float* cpt3, float* jp8) {
  if (!b->have_tentative && !b->have_committed) return fail(DMV_ERR_STATE, "linearize first");
  int i = 0;
  for (const Residual& r : b->W.residuals) {
    if (r.dropped) continue;
    // after the commit the residual's state IS the committed new state (applyRes); an OOB residual reports OOB / its old energy
    const bool tent = b->have_tentative;
    if (ns) ns[i] = tent ? r.state_NewState : r.state_state;
    if (ne) ne[i] = (float)(tent ? r.state_NewEnergy : r.state_energy);
    if (nw) nw[i] = (float)r.state_NewEnergyWithOutlier;
    if (cpt3) for (int k = 0; k < 3; k++) cpt3[3 * i + k] = r.centerProjectedTo[k];
    if (jp8) for (int k = 0; k < 8; k++) jp8[8 * i + k] = r.JpJdF[k];
    i++;
  }
  return DMV_OK;
}
int dmv_ba_get_target_energies(dmv_ba* b, int target, float* out, int cap, int* n) {
  if (!b->have_tentative && !b->have_committed) return fail(DMV_ERR_STATE, "linearize first");
  int c = 0;
  // point order, like the device (slot arrays [target][point])
  for (const Point& p : b->W.points)
    for (int ri : p.residuals) {
      const Residual& r = b->W.residuals[ri];
      if (r.target == target && !r.dropped && r.state_NewEnergyWithOutlier >= 0 && c < cap) out[c++] = (float)r.state_NewEnergyWithOutlier;
    }
  *n = c;
  return DMV_OK;
}
int dmv_ba_get_point_outputs(dmv_ba* b, float* Hdd, float* bd, float* Hcd4, float* HdiF, float* bdSum) {
  if (!b->have_tentative && !b->have_committed) return fail(DMV_ERR_STATE, "linearize first");
  for (size_t i = 0; i < b->W.points.size(); i++) {
    const Point& p = b->W.points[i];
    if (Hdd) Hdd[i] = p.Hdd_accAF;
    if (bd) bd[i] = p.bd_accAF;
    if (Hcd4) for (int k = 0; k < 4; k++) Hcd4[4 * i + k] = p.Hcd_accAF[k];
    if (HdiF) HdiF[i] = p.HdiF;
    if (bdSum) bdSum[i] = p.bdSumF;
  }
  return DMV_OK;
}
int dmv_ba_backup_points(dmv_ba* b) { for (Point& p : b->W.points) p.idepth_backup = p.idepth; return DMV_OK; }
int dmv_ba_restore_points(dmv_ba* b) {
  for (Point& p : b->W.points) { p.idepth = p.idepth_backup; p.idepth_zero = p.idepth_backup; p.deltaF = 0; }
  return DMV_OK;
}
int dmv_ba_get_idepth(dmv_ba* b, float* idepth, float* idepth_zero) {
  for (size_t i = 0; i < b->W.points.size(); i++) {
    if (idepth) idepth[i] = b->W.points[i].idepth;
    if (idepth_zero) idepth_zero[i] = b->W.points[i].idepth_zero;
  }
  return DMV_OK;
}
int dmv_ba_last_timing(dmv_ba*, float ms[4]) { ms[0] = ms[1] = ms[2] = ms[3] = 0; return DMV_OK; }

int dmv_ba_reset_oob(dmv_ba* b) {
  for (Residual& r : b->W.residuals)
    if (!r.dropped) { r.state_state = RS_IN; r.state_NewState = RS_OUTLIER; r.state_energy = r.state_NewEnergy = 0; }
  b->have_tentative = b->have_committed = false;
  return DMV_OK;
}
int dmv_ba_drop_residuals(dmv_ba* b, int n, const int32_t* idx) {
  // indices refer to the CURRENT (compacted) residual order: map them onto the oracle's stable storage
  std::vector<int> live;
  for (int i = 0; i < (int)b->W.residuals.size(); i++) if (!b->W.residuals[i].dropped) live.push_back(i);
  for (int k = 0; k < n; k++) {
    if (idx[k] < 0 || idx[k] >= (int)live.size()) return fail(DMV_ERR_INVALID, "res_idx[%d] out of range", k);
    Residual& r = b->W.residuals[live[idx[k]]];
    r.dropped = true; r.isActiveAndIsGoodNEW = false;
    std::vector<int>& list = b->W.points[r.point].residuals;
    for (size_t j = 0; j < list.size(); j++) if (list[j] == live[idx[k]]) { list.erase(list.begin() + j); break; }
  }
  return DMV_OK;
}
int dmv_ba_marginalize_points(dmv_ba* b, const dmv_ba_marg_args* a) {
  if (!b || !a || (a->n > 0 && !a->point) || !a->adHTdeltaF) return fail(DMV_ERR_INVALID, "null argument");
  // works on a COPY of the window: the launch must leave the committed linearisation alone
  Window W = b->W;
  const int n = W.nf(), N = 8 * n + CPARS;
  for (int i = 0; i < n * n; i++) for (int c = 0; c < 8; c++) W.adHTdeltaF[i](0, c) = a->adHTdeltaF[(size_t)i * 8 + c];
  for (int i = 0; i < 4; i++) W.cDeltaF[i] = a->cDeltaF[i];
  W.s.idepthFixPriorMargFac = a->idepthFixPriorMargFac;
  W.HM = MatX(N, N); W.bM.assign(N, 0.0);
  std::vector<int> pts(a->point, a->point + a->n);
  for (int p : pts) if (p < 0 || p >= (int)W.points.size()) return fail(DMV_ERR_INVALID, "point out of range");
  std::vector<int> good = W.fixLinearization(pts);
  ReducedSystem sys;
  W.marginalizePoints(pts, 1, sys);
  mock_reduce_system(b, sys);
  if (a->M) std::memcpy(a->M, sys.HA.d.data(), sizeof(double) * N * N);
  if (a->Mb) std::memcpy(a->Mb, sys.bA.data(), sizeof(double) * N);
  if (a->Msc) std::memcpy(a->Msc, sys.Hsc.d.data(), sizeof(double) * N * N);
  if (a->Mbsc) std::memcpy(a->Mbsc, sys.bsc.data(), sizeof(double) * N);
  if (a->resInM) *a->resInM = sys.resInA;
  if (a->ngoodRes) for (int i = 0; i < a->n; i++) a->ngoodRes[i] = good[i];
  int k = 0;
  for (const Residual& r : W.residuals) {
    if (r.dropped) continue;
    if (a->isLinearized) a->isLinearized[k] = r.isLinearized ? 1 : 0;
    if (a->res_toZeroF) for (int c = 0; c < 8; c++) a->res_toZeroF[(size_t)8 * k + c] = r.isLinearized ? r.res_toZeroF[c] : 0.f;
    k++;
  }
  b->have_tentative = false;
  return DMV_OK;
}

// ---- coarse tracker handle on the oracle's CoarseTracker: the frame uploaded last is "the new frame"; dmv_ct_make_coarse_depth treats it as
// the reference keyframe (the reference builds the point lists when a keyframe becomes the tracking reference)
int dmv_ct_create(const dmv_ct_config* cfg, dmv_ct** out) {
  if (!cfg || !out) return fail(DMV_ERR_INVALID, "null argument");
  dmv_ct* c = new dmv_ct();
  c->ct.levels = cfg->levels;
  for (int l = 0; l < cfg->levels; l++) { c->ct.w[l] = cfg->w >> l; c->ct.h[l] = cfg->h >> l; c->ct.pc_n[l] = 0; }
  c->pyr.resize(cfg->levels);
  *out = c;
  return DMV_OK;
}
int dmv_ct_destroy(dmv_ct* c) { delete c; return DMV_OK; }
int dmv_ct_set_K(dmv_ct* c, int l, float fx, float fy, float cx, float cy) {
  c->ct.fx[l] = fx; c->ct.fy[l] = fy; c->ct.cx[l] = cx; c->ct.cy[l] = cy;
  Mat33f K;
  K(0, 0) = fx; K(0, 2) = cx; K(1, 1) = fy; K(1, 2) = cy; K(2, 2) = 1;
  c->ct.Ki[l] = inverse3_cofactor(K);
  return DMV_OK;
}
int dmv_ct_set_huber(dmv_ct* c, float huberTH) { c->ct.s.huberTH = huberTH; return DMV_OK; }
int dmv_ct_set_ref(dmv_ct* c, int l, int n, const float* u, const float* v, const float* id, const float* col) {
  c->ct.pc_u[l].assign(u, u + n); c->ct.pc_v[l].assign(v, v + n); c->ct.pc_idepth[l].assign(id, id + n); c->ct.pc_color[l].assign(col, col + n);
  c->ct.pc_n[l] = n;
  return DMV_OK;
}
int dmv_ct_upload_new(dmv_ct* c, int l, const float* dIp) {
  c->pyr[l].assign(dIp, dIp + (size_t)c->ct.w[l] * c->ct.h[l] * 3);
  c->ct.newFrame_dIp[l] = c->pyr[l].data();
  return DMV_OK;
}
int dmv_ct_upload_new_image(dmv_ct* c, const float* image) {
  GlobalCalib g;
  g.set(c->ct.w[0], c->ct.h[0], 1.f, 1.f, 0.f, 0.f, c->ct.levels);
  float* lv[PYR_LEVELS];
  for (int l = 0; l < c->ct.levels; l++) { c->pyr[l].assign((size_t)c->ct.w[l] * c->ct.h[l] * 3, 0.f); lv[l] = c->pyr[l].data(); }
  makeImages(g, image, lv, nullptr);
  for (int l = 0; l < c->ct.levels; l++) c->ct.newFrame_dIp[l] = c->pyr[l].data();
  return DMV_OK;
}
int dmv_ct_make_coarse_depth(dmv_ct* c, int n, const float* Ku, const float* Kv, const float* nid, const float* HdiF, int32_t* pc_n_out) {
  const float* lv[PYR_LEVELS];
  for (int l = 0; l < c->ct.levels; l++) lv[l] = c->pyr[l].data();
  c->ct.makeCoarseDepthL0(n, Ku, Kv, nid, HdiF, lv);
  if (pc_n_out) for (int l = 0; l < c->ct.levels; l++) pc_n_out[l] = c->ct.pc_n[l];
  return DMV_OK;
}
int dmv_ct_get_ref(dmv_ct* c, int l, int* n, float* u, float* v, float* id, float* col) {
  const int k = c->ct.pc_n[l];
  if (n) *n = k;
  if (u) std::memcpy(u, c->ct.pc_u[l].data(), sizeof(float) * k);
  if (v) std::memcpy(v, c->ct.pc_v[l].data(), sizeof(float) * k);
  if (id) std::memcpy(id, c->ct.pc_idepth[l].data(), sizeof(float) * k);
  if (col) std::memcpy(col, c->ct.pc_color[l].data(), sizeof(float) * k);
  return DMV_OK;
}
int dmv_ct_calc_res_gs(dmv_ct* c, int lvl, const float RKi[9], const float t[3], const float affLL[2], float b0, float cutoffTH, int want_gs,
                       double res6[6], double H[64], double b[8], int* n_warped) {
  Mat33f M;
  for (int i = 0; i < 9; i++) M.d[i] = RKi[i];
  Vec3f tv;
  for (int i = 0; i < 3; i++) tv[i] = t[i];
  c->ct.calcResRaw(lvl, M, tv, affLL, cutoffTH, res6);
  if (n_warped) *n_warped = c->ct.buf_warped_n;
  if (want_gs) {
    Mat88 Hm; Vec8 bm;
    c->ct.calcGSRaw(lvl, Hm, bm, affLL[0], b0, 1);
    for (int i = 0; i < 8; i++) { for (int j = 0; j < 8; j++) H[i * 8 + j] = Hm(i, j); b[i] = bm[i]; }
  }
  return DMV_OK;
}
int dmv_ct_last_point_evaluations(dmv_ct*, double* n) { *n = 0; return DMV_OK; }
int dmv_ct_track(dmv_ct* c, const dmv_ct_track_args* in, dmv_ct_track_result* out) {
  CoarseTracker& ct = c->ct;
  ct.lastRef_aff_g2l.a = in->ref_a; ct.lastRef_aff_g2l.b = in->ref_b;
  ct.lastRef_ab_exposure = in->ref_exposure; ct.newFrame_ab_exposure = in->new_exposure;
  ct.s.coarseCutoffTH = in->coarseCutoffTH; ct.s.affineOptModeA = in->affineOptModeA; ct.s.affineOptModeB = in->affineOptModeB;
  SE3 T = SE3::fromRt(in->R, in->t);
  AffLight aff; aff.a = in->a; aff.b = in->b;
  const SE3 T0 = T; const AffLight aff0 = aff;
  int its = 0;
  const bool good = ct.trackNewestCoarse(T, aff, in->coarsestLvl, in->minResForAbort, 1, &its);
  // the reference returns early (outputs untouched) when a level's residual is NaN or above 1.5 x minResForAbort (CoarseTracker.cpp:L731-735)
  bool aborted = false;
  for (int l = 0; l <= in->coarsestLvl; l++)  // levels run from coarsestLvl down to 0: an unfinished (NaN) or over-threshold level means the early return
    if (ct.lastResiduals[l] != ct.lastResiduals[l] || ct.lastResiduals[l] > 1.5 * in->minResForAbort[l]) aborted = true;
  if (aborted) { T = T0; aff = aff0; }
  const Mat33 Rm = T.rotationMatrix();
  for (int i = 0; i < 9; i++) out->R[i] = Rm.d[i];
  for (int i = 0; i < 3; i++) out->t[i] = T.t[i];
  out->a = aff.a; out->b = aff.b;
  for (int i = 0; i < 5; i++) out->lastResiduals[i] = ct.lastResiduals[i];
  for (int i = 0; i < 3; i++) out->flowIndicators[i] = ct.lastFlowIndicators[i];
  out->trackingGood = good ? 1 : 0; out->iterations = its; out->evaluations = 0; out->status = aborted ? 2 : 0;
  return DMV_OK;
}


// ---- CoarseInitializer::calcResAndGS (dmv_ci_*) on the restatement oracle/orc_init.cpp
int dmv_ci_create(const dmv_ci_config* cfg, dmv_ci** out) {
  if (!cfg || !out || cfg->levels < 1 || cfg->levels > PYR_LEVELS) return fail(DMV_ERR_INVALID, "bad dmv_ci_config");
  dmv_ci* c = new dmv_ci();
  c->cfg = *cfg;
  c->ci.levels = cfg->levels;
  c->first.resize(cfg->levels); c->next.resize(cfg->levels);
  for (int l = 0; l < cfg->levels; l++) { c->ci.w[l] = cfg->w >> l; c->ci.h[l] = cfg->h >> l; }
  *out = c;
  return DMV_OK;
}
int dmv_ci_destroy(dmv_ci* c) { delete c; return DMV_OK; }
int dmv_ci_set_K(dmv_ci* c, int l, float fx, float fy, float cx, float cy) {
  c->ci.fx[l] = fx; c->ci.fy[l] = fy; c->ci.cx[l] = cx; c->ci.cy[l] = cy;
  return DMV_OK;
}
int dmv_ci_upload_first(dmv_ci* c, int l, const float* dIp) {
  c->first[l].assign(dIp, dIp + (size_t)3 * c->ci.w[l] * c->ci.h[l]);
  c->ci.dIFirst[l] = c->first[l].data();
  return DMV_OK;
}
int dmv_ci_upload_new(dmv_ci* c, int l, const float* dIp) {
  c->next[l].assign(dIp, dIp + (size_t)3 * c->ci.w[l] * c->ci.h[l]);
  c->ci.dINew[l] = c->next[l].data();
  return DMV_OK;
}
int dmv_ci_set_points(dmv_ci* c, int l, int n, const float* u, const float* v, const float* outlierTH) {
  c->ci.points[l].assign(n, InitPnt());
  for (int i = 0; i < n; i++) { InitPnt& p = c->ci.points[l][i]; p.u = u[i]; p.v = v[i]; p.outlierTH = outlierTH[i]; }
  if ((int)c->ci.JbBuffer_new.size() < n) { c->ci.JbBuffer.assign(n, std::array<float, 10>()); c->ci.JbBuffer_new.assign(n, std::array<float, 10>()); }
  return DMV_OK;
}
int dmv_ci_calc_res_and_gs(dmv_ci* c, const dmv_ci_eval_args* a, dmv_ci_eval_result* r) {
  const int l = a->level, n = (int)c->ci.points[l].size();
  if (!c->ci.dIFirst[l] || !c->ci.dINew[l] || n < 1) return fail(DMV_ERR_STATE, "frames / points first");
  for (int i = 0; i < n; i++) {
    InitPnt& p = c->ci.points[l][i];
    p.idepth_new = a->idepth_new[i]; p.isGood = a->isGood[i] != 0; p.energy[0] = a->energy2[2 * i]; p.energy[1] = a->energy2[2 * i + 1]; p.iR = a->iR[i];
  }
  CoarseInit& ci = c->ci;
  ci.s.huberTH = a->huberTH; ci.alphaK = a->alphaK; ci.alphaW = a->alphaW; ci.couplingWeight = a->couplingWeight;
  ci.weightZeroPriorDSOInitX = a->weightZeroPriorX; ci.weightZeroPriorDSOInitY = a->weightZeroPriorY;
  // the restatement takes a pose and forms R * Ki itself: hand it the identity rotation and the given product as "Ki"
  for (int i = 0; i < 9; i++) ci.Ki[l].d[i] = a->RKi[i];
  SE3 T;
  for (int i = 0; i < 3; i++) T.t[i] = a->t_d[i];
  AffLight aff;
  aff.a = std::log((double)a->r2new_aff[0]); aff.b = a->r2new_aff[1];
  InitSystem sys;
  ci.calcResAndGS(l, sys, T, aff, r->res3);
  std::memcpy(r->H, sys.H, sizeof(sys.H)); std::memcpy(r->b, sys.b, sizeof(sys.b));
  std::memcpy(r->Hsc, sys.Hsc, sizeof(sys.Hsc)); std::memcpy(r->bsc, sys.bsc, sizeof(sys.bsc));
  const double tsq = a->t_d[0] * a->t_d[0] + a->t_d[1] * a->t_d[1] + a->t_d[2] * a->t_d[2];
  r->alphaOpt = (ci.alphaW * (float)(tsq * n) > ci.alphaK * n) ? 0.f : ci.alphaW;
  for (int k = 0; k < 3; k++) r->b[k] += (float)(a->t_log[k] - a->t_d[k]) * r->alphaOpt * n;   // log of the real pose vs the pure translation used above
  r->n_good_new = 0;
  for (int i = 0; i < n; i++) {
    const InitPnt& p = ci.points[l][i];
    r->n_good_new += p.isGood_new;
    if (a->isGood_new) a->isGood_new[i] = p.isGood_new;
    if (a->energy_new2) { a->energy_new2[2 * i] = p.energy_new[0]; a->energy_new2[2 * i + 1] = p.energy_new[1]; }
    if (a->maxstep) a->maxstep[i] = p.maxstep;
    if (a->lastHessian_new) a->lastHessian_new[i] = p.lastHessian_new;
    if (a->JbBuffer_new10) for (int k = 0; k < 10; k++) a->JbBuffer_new10[10 * i + k] = p.isGood_new ? ci.JbBuffer_new[i][k] : 0.f;
  }
  return DMV_OK;
}
int dmv_ci_kernel_launch_count(dmv_ci*, long long* n) { *n = 0; return DMV_OK; }

}  // extern "C"
