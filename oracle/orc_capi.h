/* TEST INFRASTRUCTURE ONLY — C API of the CPU oracle (ctypes-friendly).  See orc_ba.h / orc_coarse.h.
 * Index conventions: pair index of adjoint-like tables is h + t*nf (reference acc convention,
 * AccumulatedTopHessian.cpp:L71); precalc tables are [h*nf + t] (host->targetPrecalc[target]). */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcWin OrcWin;
typedef struct OrcCT OrcCT;

/* ---- window construction ---- */
OrcWin* orc_win_create(int w, int h, int nf, const double calib_value_scaled[4], int nthreads);
void orc_win_destroy(OrcWin*);
void orc_win_set_setting(OrcWin*, const char* name, double value);
/* R: row-major 3x3, t: 3 — worldToCam_evalPT ; state/state_zero: 10 doubles (unscaled) ; dI: w*h*3 floats (borrowed pointer!) */
void orc_win_set_frame(OrcWin*, int idx, const double R[9], const double t[3], const double state[10], const double state_zero[10],
                       float ab_exposure, float frameEnergyTH, int frameID, const float* dI);
void orc_win_set_points(OrcWin*, int npts, const int32_t* host, const float* u, const float* v, const float* idepth, const float* idepth_zero,
                        const float* color8, const float* weights8, const uint8_t* hasDepthPrior);
void orc_win_set_residuals(OrcWin*, int nres, const int32_t* point, const int32_t* target, const int32_t* state_state, const float* state_energy,
                           const uint8_t* isNew);
void orc_win_set_marg_prior(OrcWin*, const double* HM, const double* bM); /* N*N, N (may be NULL => zero) */
void orc_win_prepare(OrcWin*); /* setAdjointsF + takeData + setPrecalcValues(+setDeltaF) */

/* ---- host-side tables (for feeding / checking the product) ---- */
int orc_win_nres(OrcWin*);
int orc_win_npts(OrcWin*);
int orc_win_nf(OrcWin*);
/* 32 floats per pair [h*nf+t]: KRKi[9] Kt[3] R0[9] t0[3] aff[2] b0 pad[5] */
void orc_win_get_precalc(OrcWin*, float* out);
void orc_win_get_RT(OrcWin*, float* out);                               /* nf*nf*12: PRE_RTll row-major | PRE_tTll, [h*nf+t] */
void orc_win_get_adjoints(OrcWin*, double* adHost, double* adTarget); /* nf*nf*64 each, [h+t*nf], row-major 8x8 */
void orc_win_get_adHTdeltaF(OrcWin*, float* out);                     /* nf*nf*8 */
void orc_win_get_frame_tables(OrcWin*, double* prior8, double* delta_prior8, double* delta8, float* frameEnergyTH); /* nf*8 ... */
void orc_win_get_calib(OrcWin*, float* fxfycxcy_and_inv8, float* cDeltaF4, double* cPrior4);

/* ---- hot path ---- */
double orc_win_linearize_all(OrcWin*, int fixLinearization, int updateEnergyTH);
void orc_win_apply_res(OrcWin*);
/* Test helper for threshold ties: impose the classification another implementation reached (IN / OOB / OUTLIER per residual) on the
 * tentative linearisation, so that the committed systems of both sides cover the same residual set.  IN <-> OUTLIER flips reuse the
 * Jacobian linearize() already produced (it is computed before the classification, Residuals.cpp:L260-273); -> OOB drops the residual.
 * A residual that left through an OOB exit here but not there cannot be fixed (no Jacobian): counted in *unfixable.
 * Returns the energy sum of the imposed classification (what linearizeAll would have returned). */
double orc_win_override_new_states(OrcWin*, const int32_t* newState, int* changed, int* unfixable);
/* per-residual outputs of the last linearize */
void orc_win_get_res_outputs(OrcWin*, int32_t* newState, float* newEnergy, float* newEnergyWithOutlier, float* centerProjectedTo3,
                             float* Jnew74 /* may be NULL */, int32_t* state_state, uint8_t* isActive, float* JpJdF8);
void orc_win_accumulate(OrcWin*, int precision, double* HA, double* bA, double* HL, double* bL, double* Hsc, double* bsc, int* resInA);
/* per-point outputs of the last accumulate */
void orc_win_get_point_outputs(OrcWin*, float* Hdd_accAF, float* bd_accAF, float* Hcd_accAF4, float* HdiF, float* bdSumF, float* step,
                               float* idepth, float* maxRelBaseline);
void orc_win_solve(OrcWin*, int iteration, double lambda, int precision, double* x_out, double* HFinal, double* bFinal);
void orc_win_resubstitute(OrcWin*, const double* x);
double orc_win_calc_LEnergy(OrcWin*);
double orc_win_calc_MEnergy(OrcWin*);
int orc_win_optimize(OrcWin*, int mnumOptIts, int precision, double* energyLog, int energyLogCap);
void orc_win_get_frame_states(OrcWin*, double* state10);
/* one full GN iteration as timed by the bench: solveSystem (accumulate A/L/SC + stitch + solve + resubstitute) + doStep + linearizeAll */
double orc_win_gn_iteration(OrcWin*, double lambda, int precision, int do_step);

/* the measured path of one GN iteration without the dense solve: accumulate+stitch, resubstitute(x), step, linearizeAll, applyRes */
double orc_win_hot_iteration(OrcWin*, const double* x, int precision);

/* finite-difference helper: raw (un-weighted) residuals of one residual, evaluated from first principles in double,
 * after adding dstate (unscaled, 8) to the host / target frame state, didepth to the point and dcalib (unscaled) to the calib */
int orc_win_eval_raw_double(OrcWin*, int res_idx, const double dstate_host[8], const double dstate_target[8], double didepth,
                            const double dcalib[4], double r_raw[8]);

/* ---- images ---- */
int orc_pyr_levels(int w, int h, int forceLevels);
/* out: concatenated levels, each w_l*h_l*3 floats; returns total floats written */
int64_t orc_make_images(int w, int h, int levels, float fx, float fy, float cx, float cy, const float* color, float* dIp_out, float* absSqGrad_out);
int orc_init_point(const float* dI, int w, float u, float v, float* color8, float* weights8);

/* ---- coarse tracker ---- */
OrcCT* orc_ct_create(int w, int h, int levels, float fx, float fy, float cx, float cy);
void orc_ct_destroy(OrcCT*);
void orc_ct_set_setting(OrcCT*, const char* name, double value);
void orc_ct_set_ref_points(OrcCT*, int lvl, int n, const float* u, const float* v, const float* idepth, const float* color);
int orc_ct_make_coarse_depth(OrcCT*, int n, const float* Ku, const float* Kv, const float* new_idepth, const float* HdiF, const float* ref_dIp_concat);
int orc_ct_get_ref_points(OrcCT*, int lvl, float* u, float* v, float* idepth, float* color); /* returns n; NULL pointers allowed */
void orc_ct_set_new_frame(OrcCT*, const float* dIp_concat /* borrowed */, float ref_exposure, float new_exposure, double ref_a, double ref_b);
void orc_ct_get_K(OrcCT*, int lvl, float* fxfycxcy, int* wh);
/* refToNew: R(9 row-major), t(3) */
void orc_ct_calc_res(OrcCT*, int lvl, const double R[9], const double t[3], double aff_a, double aff_b, float cutoffTH, double out6[6]);
int orc_ct_get_warped(OrcCT*, float* buf8xn /* idepth,u,v,dx,dy,residual,weight,refColor each n */);
void orc_ct_calc_gs(OrcCT*, int lvl, double aff_a, double aff_b, int precision, double H64[64], double b8[8]);
int orc_ct_track(OrcCT*, double R[9], double t[3], double* aff_a, double* aff_b, int coarsestLvl, const double minResForAbort[5], int precision,
                 double lastResiduals5[5], double flow3[3], int* iterations);

/* SE3 helpers (so python tests need no own Lie algebra) */
void orc_se3_exp(const double xi[6], double R[9], double t[3]);
void orc_se3_log(const double R[9], const double t[3], double xi[6]);
void orc_se3_mul(const double Ra[9], const double ta[3], const double Rb[9], const double tb[3], double R[9], double t[3]);
void orc_se3_inv(const double Ra[9], const double ta[3], double R[9], double t[3]);

#ifdef __cplusplus
}
#endif
