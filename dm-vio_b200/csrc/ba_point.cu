// ba_point_kernel — the fused residual / Jacobian / Hessian / Schur kernel of the BA hot path.
//
// One CTA = one chunk of P points of ONE host frame h.  The CTA is wide on purpose (nf * P/(4*ITER) warps, e.g. 28 warps for
// nf = 7, P = 16): the work per point-residual is a ~600-instruction dependent chain, so the only way to keep an SM busy is
// many independent warps; every warp evaluates ITER quads of 4 residuals (8 lanes per residual = the 8 pattern pixels).
//
//   prologue  cp.async staging of the chunk's point records / residual states / the host's adjoint blocks;
//             optional fused EnergyFunctional::resubstituteFPt + point step (EnergyFunctional.cpp:L295-321, FullSystemOptimize.cpp:L264-272)
//   phase A   PointFrameResidual::linearize (Residuals.cpp:L78-274): project, 4-tap float4 gather from the target plane,
//             Huber residual, 8-lane butterfly sums; rows of the pair's 13x13 block (AccumulatorApprox::update*,
//             MatrixAccumulators.h:L754-915) written per 8-lane group to shared memory — no RawResidualJacobian is stored
//   phase A'  sum of the group partials -> one fp64 RED per (pair, entry) into the global accumulator
//   phase B   per point: Hdd/bd/Hcd, HdiF, bdSum (AccumulatedSCHessian.cpp:L36-58) and the Schur vector in ABSOLUTE frame
//             coordinates  w_p = [Hcd | adHost/adTarget * JpJdF ... | bdSum]
//   phase C   weighted Gram  sum_p HdiF w_p w_p^T  in 4x4 register tiles (replaces the nf^3 accD blocks + stitch of
//             AccumulatedSCHessian.cpp:L34-157), 4 lanes per tile split the chunk's points, fp64 RED to global
#include "ba_common.cuh"

namespace dmv {

template <int P>
struct PointSmem {
  float part[MAXF][P / 4][4][TOP_PART];  // [target][quad][8-lane group][rows 0..9 upper-triangular (85) | 6 | pad]
  float rec[P][MAXF][REC];               // per (point,target): JpJdF[8] Hdd bd Hcd[4] active pad
  float Wv[P][8 * MAXF + 8];             // Schur vectors
  float hdi[P];
  float adH[MAXF][64];
  float adT[MAXF][8];
  float id[P], idz[P];                   // inverse depths the residuals were evaluated at (written by the first target's groups)
  float misc[MAXF * (P / 4)][4];
};

// geometric Jacobians of the centre pixel (Residuals.cpp:L113-156): x = d(Ku)/d[C4|xi6], y = d(Kv)/d[C4|xi6], dd = d(Ku,Kv)/d(idepth)
__device__ __forceinline__ void geo_jac(const float* pc, float Kl0, float Kl1, float cu, float cv, float drescale, float new_idepth, float fx,
                                        float fy, float fxi, float fyi, float x[10], float y[10], float& ddx, float& ddy) {
  const float dCx2 = drescale * (pc[18] * cu - pc[12]);
  const float dCx3 = fx * drescale * (pc[19] * cu - pc[13]) * fyi;
  const float dCy2 = fy * drescale * (pc[18] * cv - pc[15]) * fxi;
  const float dCy3 = drescale * (pc[19] * cv - pc[16]);
  x[0] = (Kl0 * dCx2 + cu) * 50.0f; x[1] = (Kl1 * dCx3) * 50.0f; x[2] = (dCx2 + 1.f) * 50.0f; x[3] = dCx3 * 50.0f;
  y[0] = (Kl0 * dCy2) * 50.0f; y[1] = (Kl1 * dCy3 + cv) * 50.0f; y[2] = dCy2 * 50.0f; y[3] = (dCy3 + 1.f) * 50.0f;
  x[4] = new_idepth * fx; x[5] = 0.f; x[6] = -new_idepth * cu * fx; x[7] = -cu * cv * fx; x[8] = (1.f + cu * cu) * fx; x[9] = -cv * fx;
  y[4] = 0.f; y[5] = new_idepth * fy; y[6] = -new_idepth * cv * fy; y[7] = -(1.f + cv * cv) * fy; y[8] = cu * cv * fy; y[9] = cu * fy;
  ddx = drescale * (pc[21] - pc[23] * cu) * fx;  // Jpdd (SCALE_IDEPTH = 1)
  ddy = drescale * (pc[22] - pc[23] * cv) * fy;
}

// MARG = true is the marginalisation launch (dmv_ba_marginalize_points): only the points flagged in W.marg_mask take part, their
// residuals are re-linearised from scratch (PointFrameResidual::resetOOB; FullSystem.cpp:L826-838), EFResidual::fixLinearizationF
// (EnergyFunctionalStructs.cpp:L88-114) turns resF into res_toZeroF, and the accumulation is AccumulatedTopHessian::addPoint<2> +
// AccumulatedSCHessian::addPoint(p, shiftPriorToZero = false) with priorF * idepthFixPriorMargFac (EnergyFunctional.cpp:L678-742).
template <int P, int ITER, bool MARG>
__global__ void __launch_bounds__(32 * MAXF * (P / (4 * ITER)), 1)
    ba_point_kernel(const __grid_constant__ BAWinDev W, const __grid_constant__ BAIter it) {
  constexpr int WQ = P / (4 * ITER);  // warps per target
  // Programmatic dependent launch: the stitch kernel of this iteration may be scheduled on idle SMs right away; it blocks in
  // griddepcontrol.wait until every CTA of this grid has finished and its global writes / REDs are visible.
  asm volatile("griddepcontrol.launch_dependents;");
  if (W.dbg & 8) return;              // experiment: launch + event overhead only
  STAMP(0);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PointSmem<P>& S = *reinterpret_cast<PointSmem<P>*>(smem_raw);

  const int nf = W.nf;
  int h = 0;  // host frame of this chunk: branch-free so that the 7 constant-bank loads are independent
#pragma unroll
  for (int k = 1; k < MAXF; k++) h += ((int)blockIdx.x >= W.chunk_beg[k]) ? 1 : 0;
  h = min(h, nf - 1);
  const int ch_start = W.host_start[h] + ((int)blockIdx.x - W.chunk_beg[h]) * P;
  const int ch_count = min(P, W.host_start[h + 1] - ch_start);
  STAMP(8);
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int mp = W.mp;
  double* __restrict__ acc = W.acc;
  double* __restrict__ acc_misc = acc + (size_t)nf * nf * TOP_PART + (size_t)W.ntiles * 16;

  // ---------------------------------------------------------------- prologue: only the adjoint blocks are staged (needed from
  // phase B on); everything phase A needs is loaded straight into the registers of the lanes that use it, so the first
  // block-wide barrier comes after phase A
  {
    const BAAdj* __restrict__ A = W.adj;
    for (int i = tid; i < nf * 16; i += nthreads) cp_async16(&S.adH[i >> 4][(i & 15) * 4], &A->adHostF[h * nf + (i >> 4)][(i & 15) * 4]);
    for (int i = tid; i < nf * 2; i += nthreads) cp_async16(&S.adT[i >> 1][(i & 1) * 4], &A->adTdiagF[h * nf + (i >> 1)][(i & 1) * 4]);
    asm volatile("cp.async.commit_group;" ::: "memory");
    STAMP(9);
  }
  // the thread that finalises point pl_b in phase B fetches its prior now
  const int pl_b = nthreads - 1 - tid;
  float prior_b = (pl_b < ch_count) ? __ldg(W.priorF + ch_start + pl_b) : 0.f;
  bool masked_b = true;
  if constexpr (MARG) {
    masked_b = (pl_b < ch_count) && __ldg(W.marg_mask + ch_start + pl_b) != 0;
    prior_b *= __ldg(&W.marg->priorFac);
  }
  STAMP(6);

  // ---------------------------------------------------------------- phase A
  const int t = warp / WQ;  // target frame of this warp
  const int q0 = warp - t * WQ;
  if (t < nf && t != h) {
    const int g = lane >> 3, j = lane & 7;
    const float* pc = it.precalc[h * nf + t];
    const float fx = it.calib[0], fy = it.calib[1], cx = it.calib[2], cy = it.calib[3];
    const float fxi = it.calib[4], fyi = it.calib[5];
    const float TH = fmaxf(it.TH[h], it.TH[t]);
    const float wM3 = (float)(W.w - 3), hM3 = (float)(W.h - 3);
    const float4* __restrict__ img = W.img[t];
    const int iw = W.w;
    const float huber = W.huberTH, oth = W.outlierTHSum;
    const int pdx = c_pattern[j][0], pdy = c_pattern[j][1];
    float e_sum = 0.f;
    int n_in = 0, n_oob = 0, n_outl = 0;
    const int t0 = (h == 0) ? 1 : 0;  // first target frame: its groups publish the per-point step / depths
    float rs_step2 = 0.f, rs_nid = 0.f, rs_cnt = 0.f;

#pragma unroll 1
    for (int pass = 0; pass < ITER; pass++) {
      const int quad = q0 + pass * WQ;
      float* part = S.part[t][quad][g];
      if (quad * 4 >= ch_count) {  // nothing to do for this quad: its partial slot must still read as zero
        for (int c = j; c < TOP_PART; c += 8) part[c] = 0.f;
        continue;
      }
      const int pl = min(quad * 4 + g, ch_count - 1);
      const bool valid = quad * 4 + g < ch_count;
      const int p = ch_start + pl;
      const int slot = t * mp + p;
      // ---- direct loads (all independent: one memory round trip)
      int st = valid ? (int)__ldg(W.st_in + slot) : RES_NONE;
      float en_old = __ldg(W.en_in + slot);
      bool masked = true;
      if constexpr (MARG) {  // resetOOB: every existing residual of a flagged point starts as IN with zero energy; other points sit out
        masked = valid && __ldg(W.marg_mask + p) != 0;
        st = (masked && st != RES_NONE) ? RES_IN : RES_NONE;
        en_old = 0.f;
      }
      const float2 uv = __ldg(W.uv + p);
      const float col = __ldg(W.color + (size_t)p * 8 + j);
      const float wgt = __ldg(W.weights + (size_t)p * 8 + j);
      float idepth, idz;
      if (it.have_x) {
        // fused EnergyFunctional::resubstituteFPt (EnergyFunctional.cpp:L295-321) + point step (FullSystemOptimize.cpp:L264-272):
        // lane j of the group takes target frame j of the point's committed residuals; every group of the point recomputes the
        // same step (the loads hit L1/L2), the group of the FIRST target publishes it
        float d = 0.f;
        bool good = false;
        if (j < nf && j != h) {
          const int cs = j * mp + p;
          const int stc = __ldg(W.c_st + cs);
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(W.c_jpjd + (size_t)cs * 8));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(W.c_jpjd + (size_t)cs * 8) + 1);
          const float* xa = it.xAd[h * nf + j];
          const float dot = xa[0] * a0.x + xa[1] * a0.y + xa[2] * a0.z + xa[3] * a0.w + xa[4] * a1.x + xa[5] * a1.y + xa[6] * a1.z + xa[7] * a1.w;
          good = (stc == RES_IN);
          d = good ? dot : 0.f;
        }
        const float4 po0 = __ldg(reinterpret_cast<const float4*>(W.c_pout + (size_t)p * 8));
        const float4 po1 = __ldg(reinterpret_cast<const float4*>(W.c_pout + (size_t)p * 8) + 1);
        const float idb = __ldg(W.idepth_backup + p);
        const unsigned gb = __ballot_sync(0xffffffffu, good);
        const int ngood = __popc((gb >> (g * 8)) & 0xffu);
        const float dsum = group_sum8(d);
        const float bsum = po1.w - (it.xc[0] * po0.z + it.xc[1] * po0.w + it.xc[2] * po1.x + it.xc[3] * po1.y) - dsum;
        const float step = ngood > 0 ? -bsum * po1.z : 0.f;
        idepth = idb + step;
        idz = idepth;  // DM-VIO: idepth_zero follows (setIdepthZero in doStepFromBackup); the host aliases the pointers
        if (t == t0 && j == 0 && valid) {
          W.step[p] = step;
          W.idepth_out[p] = idepth;
          rs_step2 += step * step;
          rs_nid += fabsf(idb);
          rs_cnt += 1.f;
        }
      } else {
        idepth = __ldg(W.idepth + p);
        idz = __ldg(W.idepth_zero + p);
      }
      if (t == t0 && j == 0 && valid) { S.id[pl] = idepth; S.idz[pl] = idz; }
      bool live = (st != RES_NONE) && (st != RES_OOB);

      // ---- centre pixel at the FEJ point (ResidualProjections.h:L62-87, Residuals.cpp:L108-157)
      const float Kl0 = (uv.x - cx) * fxi, Kl1 = (uv.y - cy) * fyi;
      const float q2 = pc[18] * Kl0 + pc[19] * Kl1 + pc[20] + pc[23] * idz;
      const float drescale = 1.0f / q2;
      const float new_idepth = idz * drescale;
      const float cu = (pc[12] * Kl0 + pc[13] * Kl1 + pc[14] + pc[21] * idz) * drescale;
      const float cv = (pc[15] * Kl0 + pc[16] * Kl1 + pc[17] + pc[22] * idz) * drescale;
      const float cKu = cu * fx + cx, cKv = cv * fy + cy;
      live = live && (drescale > 0.f) && cKu > 1.1f && cKv > 1.1f && cKu < wM3 && cKv < hM3;

      // ---- own pattern pixel at the current state (ResidualProjections.h:L47-57)
      const float pu = uv.x + (float)pdx, pv = uv.y + (float)pdy;
      const float r2 = pc[6] * pu + pc[7] * pv + pc[8] + pc[11] * idepth;
      const float Ku = (pc[0] * pu + pc[1] * pv + pc[2] + pc[9] * idepth) / r2;
      const float Kv = (pc[3] * pu + pc[4] * pv + pc[5] + pc[10] * idepth) / r2;
      const bool px_ok = Ku > 1.1f && Kv > 1.1f && Ku < wM3 && Kv < hM3;
      unsigned bal = __ballot_sync(0xffffffffu, px_ok);
      live = live && (((bal >> (g * 8)) & 0xffu) == 0xffu);

      float h0 = 0.f, h1 = 0.f, h2 = 0.f;
      if (live && !(W.dbg & 4)) {  // getInterpolatedElement33 (util/globalFuncs.h:L103-118)
        const int ix = (int)Ku, iy = (int)Kv;
        const float dx = Ku - ix, dy = Kv - iy, dxdy = dx * dy;
        const float4* bp = img + (size_t)iy * iw + ix;
        const float4 tl = __ldg(bp), tr = __ldg(bp + 1), bl = __ldg(bp + iw), brr = __ldg(bp + iw + 1);
        const float w11 = dxdy, w10 = dy - dxdy, w01 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
        h0 = w11 * brr.x + w10 * bl.x + w01 * tr.x + w00 * tl.x;
        h1 = w11 * brr.y + w10 * bl.y + w01 * tr.y + w00 * tl.y;
        h2 = w11 * brr.z + w10 * bl.z + w01 * tr.z + w00 * tl.z;
      }
      bal = __ballot_sync(0xffffffffu, isfinite(h0));
      live = live && (((bal >> (g * 8)) & 0xffu) == 0xffu);

      // ---- photometric residual, gradient weight, Huber (Residuals.cpp:L194-258)
      const float residual = h0 - (pc[24] * col + pc[25]);
      const float drdA = col - pc[26];
      float w = sqrtf(oth / (oth + (h1 * h1 + h2 * h2)));
      w = 0.5f * (w + wgt);
      const float ar = fabsf(residual);
      float hw = ar < huber ? 1.f : huber / ar;
      float e_px = w * w * hw * residual * residual * (2.f - hw);
      if (hw < 1.f) hw = sqrtf(hw);
      hw = hw * w;
      if (!live) { hw = 0.f; e_px = 0.f; }
      const float gx = h1 * hw, gy = h2 * hw;
      const float resF = live ? residual * hw : 0.f;
      const float ja = drdA * hw, jb = hw;
      const float jaF = W.zeroA ? 0.f : ja, jbF = W.zeroB ? 0.f : jb;

      const float JI00 = group_sum8(gx * gx), JI11 = group_sum8(gy * gy), JI10 = group_sum8(gx * gy);
      const float JabJI00 = group_sum8(ja * gx), JabJI01 = group_sum8(ja * gy), JabJI10 = group_sum8(jb * gx), JabJI11 = group_sum8(jb * gy);
      const float Jab00 = group_sum8(ja * ja), Jab01 = group_sum8(ja * jb), Jab11 = group_sum8(jb * jb);
      float resAcc = resF;  // what the right-hand sides are built from: resF, or res_toZeroF when marginalising
      float x[10], y[10], ddx, ddy;
      if constexpr (MARG) {
        geo_jac(pc, Kl0, Kl1, cu, cv, drescale, new_idepth, fx, fy, fxi, fyi, x, y, ddx, ddy);
        const float* dp = W.marg->adHTdelta[h * nf + t];
        const float* cD = W.marg->cDelta;
        const float dlt = idepth - idz;  // EFPoint::deltaF
        const float jpx = (x[4] * dp[0] + x[5] * dp[1] + x[6] * dp[2] + x[7] * dp[3] + x[8] * dp[4] + x[9] * dp[5]) +
                          (x[0] * cD[0] + x[1] * cD[1] + x[2] * cD[2] + x[3] * cD[3]) + ddx * dlt;
        const float jpy = (y[4] * dp[0] + y[5] * dp[1] + y[6] * dp[2] + y[7] * dp[3] + y[8] * dp[4] + y[9] * dp[5]) +
                          (y[0] * cD[0] + y[1] * cD[1] + y[2] * cD[2] + y[3] * cD[3]) + ddy * dlt;
        resAcc = live ? (((resF - gx * jpx) - gy * jpy) - jaF * dp[6]) - jbF * dp[7] : 0.f;
      }
      const float JIr0 = group_sum8(resAcc * gx), JIr1 = group_sum8(resAcc * gy);
      const float Jabr0 = group_sum8(resAcc * jaF), Jabr1 = group_sum8(resAcc * jbF);
      const float rr = group_sum8(resAcc * resAcc);
      const float energy = group_sum8(e_px);
      // the reference sums hw*hw*(hitColor[1]^2+hitColor[2]^2) with hitColor already multiplied by hw (Residuals.cpp:L217-244)
      const float wJI2 = group_sum8(hw * hw * (gx * gx + gy * gy));

      // ---- classification (Residuals.cpp:L260-273) and per-residual outputs
      int newState;
      float newEnergy;
      if (st == RES_NONE) {
        newState = RES_NONE; newEnergy = 0.f;
      } else if (!live) {
        newState = RES_OOB; newEnergy = en_old;  // OOB exits return the old state_energy
      } else if (energy > TH || wJI2 < 2.f) {
        newState = RES_OUTLIER; newEnergy = TH;
      } else {
        newState = RES_IN; newEnergy = energy;
      }
      const bool in = (newState == RES_IN);
      if (j == 0 && st != RES_NONE) {
        e_sum += newEnergy;
        n_in += in; n_oob += (newState == RES_OOB); n_outl += (newState == RES_OUTLIER);
      }
      if constexpr (MARG) {
        if (masked) W.marg_rtz[(size_t)slot * 8 + j] = in ? resAcc : 0.f;
      }
      if (valid && j == 0 && masked) {
        W.st_new[slot] = (uint8_t)newState;
        W.en_new[slot] = newEnergy;
        W.en_wo[slot] = (st == RES_NONE || !live) ? -1.f : energy;
        const size_t plane = (size_t)MAXF * mp;
        W.cpt[slot] = cKu; W.cpt[plane + slot] = cKv; W.cpt[2 * plane + slot] = new_idepth;
      }

      if (in) {
        if constexpr (!MARG) geo_jac(pc, Kl0, Kl1, cu, cv, drescale, new_idepth, fx, fy, fxi, fyi, x, y, ddx, ddy);
        // EFResidual::takeDataF (EnergyFunctionalStructs.cpp:L39-49) and the per-point terms of addPoint (AccumulatedTopHessian.cpp:L131-135)
        const float J0 = JI00 * ddx + JI10 * ddy, J1 = JI10 * ddx + JI11 * ddy;  // JIdx2 * Jpdd
        if (j == 0) {
          float* rec = S.rec[pl][t];
          float4 v0, v1, v2, v3;
          v0.x = x[4] * J0 + y[4] * J1; v0.y = x[5] * J0 + y[5] * J1; v0.z = x[6] * J0 + y[6] * J1; v0.w = x[7] * J0 + y[7] * J1;
          v1.x = x[8] * J0 + y[8] * J1; v1.y = x[9] * J0 + y[9] * J1;
          v1.z = JabJI00 * ddx + JabJI01 * ddy; v1.w = JabJI10 * ddx + JabJI11 * ddy;
          v2.x = J0 * ddx + J1 * ddy;            // Hdd
          v2.y = JIr0 * ddx + JIr1 * ddy;        // bd
          v2.z = x[0] * J0 + y[0] * J1; v2.w = x[1] * J0 + y[1] * J1;  // Hcd[0..1]
          v3.x = x[2] * J0 + y[2] * J1; v3.y = x[3] * J0 + y[3] * J1;  // Hcd[2..3]
          v3.z = 1.f; v3.w = 0.f;                                      // active flag
          reinterpret_cast<float4*>(rec)[0] = v0;
          reinterpret_cast<float4*>(rec)[1] = v1;
          reinterpret_cast<float4*>(rec)[2] = v2;
          reinterpret_cast<float4*>(rec)[3] = v3;
          float4* gj = reinterpret_cast<float4*>(W.jpjd + (size_t)slot * 8);
          gj[0] = v0; gj[1] = v1;
        }
        // rows of the pair's 13x13 block: lane j owns row j, lanes (j&1) own rows 8/9 (AccumulatorApprox::update/updateTopRight)
        const float xr1 = pick8(x, j), yr1 = pick8(y, j);
        const float xr2 = (j & 1) ? x[9] : x[8], yr2 = (j & 1) ? y[9] : y[8];
        const float al1 = JI00 * xr1 + JI10 * yr1, be1 = JI10 * xr1 + JI11 * yr1;
        const float al2 = JI00 * xr2 + JI10 * yr2, be2 = JI10 * xr2 + JI11 * yr2;
        // packed upper-triangular storage: row r keeps columns r..12 only (the lower triangle is the mirror image)
        float* row1 = part + top_off(j) - j;
        const int r2 = 8 + (j & 1);
        float* row2 = part + top_off(r2) - r2;
#pragma unroll
        for (int c = 0; c < 10; c++) {
          const float v1 = al1 * x[c] + be1 * y[c];
          if (c >= j) row1[c] = v1;
          if (c >= 8 && j < 2 && c >= r2) row2[c] = al2 * x[c] + be2 * y[c];
        }
        row1[10] = xr1 * JabJI00 + yr1 * JabJI01;
        row1[11] = xr1 * JabJI10 + yr1 * JabJI11;
        row1[12] = xr1 * JIr0 + yr1 * JIr1;
        if (j < 2) {
          row2[10] = xr2 * JabJI00 + yr2 * JabJI01;
          row2[11] = xr2 * JabJI10 + yr2 * JabJI11;
          row2[12] = xr2 * JIr0 + yr2 * JIr1;
        }
        if (j < 6) part[TOP_TRI + j] = (j == 0) ? Jab00 : (j == 1) ? Jab01 : (j == 2) ? Jabr0 : (j == 3) ? Jab11 : (j == 4) ? Jabr1 : rr;
      } else {
        for (int c = j; c < TOP_PART; c += 8) part[c] = 0.f;
        if (j == 0 && valid) S.rec[pl][t][14] = 0.f;  // "no active residual" flag (shared memory is not pre-zeroed)
      }
    }
    float es = e_sum, fin = (float)n_in, foob = (float)n_oob, fout = (float)n_outl;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      es += __shfl_xor_sync(0xffffffffu, es, m);
      fin += __shfl_xor_sync(0xffffffffu, fin, m);
      foob += __shfl_xor_sync(0xffffffffu, foob, m);
      fout += __shfl_xor_sync(0xffffffffu, fout, m);
    }
    if (lane == 0) { S.misc[warp][0] = es; S.misc[warp][1] = fin; S.misc[warp][2] = foob; S.misc[warp][3] = fout; }
    if (it.have_x && t == t0) {  // sum step^2, sum |idepth_backup|, #points: feed only the convergence test of doStepFromBackup
      rs_step2 += __shfl_xor_sync(0xffffffffu, rs_step2, 8); rs_step2 += __shfl_xor_sync(0xffffffffu, rs_step2, 16);
      rs_nid += __shfl_xor_sync(0xffffffffu, rs_nid, 8); rs_nid += __shfl_xor_sync(0xffffffffu, rs_nid, 16);
      rs_cnt += __shfl_xor_sync(0xffffffffu, rs_cnt, 8); rs_cnt += __shfl_xor_sync(0xffffffffu, rs_cnt, 16);
      if (lane == 0) {
        RED_ADD(acc_misc + 4, (double)rs_step2);
        RED_ADD(acc_misc + 5, (double)rs_nid);
        RED_ADD(acc_misc + 6, (double)rs_cnt);
      }
    }
  } else if (lane < 4) {
    S.misc[warp][lane] = 0.f;  // idle warps (t == h or t >= nf) still own a counter slot
  }
  cp_async_wait_all();  // the adjoint blocks staged in the prologue
  __syncthreads();
  STAMP(3);

  // ---------------------------------------------------------------- phase A': fold the group partials, RED to the pair accumulators
  for (int e0 = tid; e0 < nf * TOP_PART; e0 += nthreads) {
    const int e = (e0 + (int)blockIdx.x * 53) % (nf * TOP_PART);  // staggered start per CTA (RED address spreading)
    const int tt = e / TOP_PART, k = e - tt * TOP_PART;
    if (tt == h || k >= TOP_USED) continue;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < P / 4; q++)
#pragma unroll
      for (int gg = 0; gg < 4; gg++) s += S.part[tt][q][gg][k];
    RED_ADD(acc + (size_t)(h * nf + tt) * TOP_PART + k, (double)s);
  }
  if (tid < 4) {
    double s = 0.0;
    for (int wv = 0; wv < nf * WQ; wv++) s += (double)S.misc[wv][tid];
    RED_ADD(acc_misc + tid, s);
  }

  // ---------------------------------------------------------------- phase B
  const int N = W.N;
  // (a) the point finalisation (AccumulatedSCHessian.cpp:L36-58) runs on the LAST warps' lanes so that it overlaps the vector tasks
  {
    const int pl = pl_b;
    if (pl < ch_count) {
      const int p = ch_start + pl;
      float Hdd = 0.f, bd = 0.f, Hcd0 = 0.f, Hcd1 = 0.f, Hcd2 = 0.f, Hcd3 = 0.f;
      int ngood = 0;
      for (int tt = 0; tt < nf; tt++) {
        if (tt == h) continue;
        const float* rec = S.rec[pl][tt];
        if (rec[14] != 0.f) {
          ngood++;
          Hdd += rec[8]; bd += rec[9]; Hcd0 += rec[10]; Hcd1 += rec[11]; Hcd2 += rec[12]; Hcd3 += rec[13];
        }
      }
      float HdiF = 0.f, bdSum = 0.f;
      float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
      if (ngood > 0) {
        const float prior = prior_b;
        float H = Hdd + prior;
        if (H < 1e-10f) H = 1e-10f;
        HdiF = 1.0f / H;
        bdSum = MARG ? bd : bd + prior * (S.id[pl] - S.idz[pl]);  // shiftPriorToZero (AccumulatedSCHessian.cpp:L47-50)
        w0 = Hcd0; w1 = Hcd1; w2 = Hcd2; w3 = Hcd3;
      }
      S.Wv[pl][0] = w0; S.Wv[pl][1] = w1; S.Wv[pl][2] = w2; S.Wv[pl][3] = w3;
      S.Wv[pl][N] = bdSum;
      for (int c = N + 1; c < W.NW; c++) S.Wv[pl][c] = 0.f;  // padding columns of the last 4x4 tiles
      S.hdi[pl] = HdiF;
      if (masked_b) {
        float4* po = reinterpret_cast<float4*>(W.pout + (size_t)p * 8);
        po[0] = make_float4(Hdd, bd, Hcd0, Hcd1);
        po[1] = make_float4(Hcd2, Hcd3, HdiF, bdSum);
      }
    }
  }
  // (b) Schur vector entries: (point, frame, k)
  for (int idx = tid; idx < ch_count * nf * 8; idx += nthreads) {
    const int k = idx & 7;
    const int pf = idx >> 3;
    const int pl = pf / nf, f = pf - pl * nf;
    float val = 0.f;
    if (f == h) {
      for (int tt = 0; tt < nf; tt++) {
        if (tt == h) continue;
        const float* rec = S.rec[pl][tt];
        if (rec[14] != 0.f) {
          const float* A = &S.adH[tt][k * 8];
#pragma unroll
          for (int c = 0; c < 8; c++) val += A[c] * rec[c];
        }
      }
    } else {
      const float* rec = S.rec[pl][f];
      if (rec[14] != 0.f) val = S.adT[f][k] * rec[k];
    }
    S.Wv[pl][4 + 8 * f + k] = val;
  }
  __syncthreads();
  STAMP(4);

  // ---------------------------------------------------------------- phase C: 4 lanes per 4x4 tile, each takes every 4th point
  if (!(W.dbg & 2)) {
    const int T = W.T;
    double* scp = acc + (size_t)nf * nf * TOP_PART;
    const int ntask = W.ntiles * 4;
    for (int task = tid; task < ((ntask + 31) & ~31); task += nthreads) {  // whole warps iterate together (shuffles below)
      // every CTA walks the tiles from a different start so that concurrent CTAs do not hammer the same L2 addresses
      const int tile = (min(task >> 2, W.ntiles - 1) + (int)blockIdx.x * 29) % W.ntiles, pg = task & 3;
      const bool tvalid = task < ntask;
      int ti = 0, rem = tile;
      while (rem >= T - ti) { rem -= T - ti; ti++; }
      const int tj = ti + rem;
      float a[4][4];
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) a[r][c] = 0.f;
      for (int pl = pg; pl < ch_count; pl += 4) {
        const float s = S.hdi[pl];
        const float4 wi = *reinterpret_cast<const float4*>(&S.Wv[pl][4 * ti]);
        const float4 wj = *reinterpret_cast<const float4*>(&S.Wv[pl][4 * tj]);
        const float si[4] = {s * wi.x, s * wi.y, s * wi.z, s * wi.w};
        const float vj[4] = {wj.x, wj.y, wj.z, wj.w};
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
          for (int c = 0; c < 4; c++) a[r][c] += si[r] * vj[c];
      }
      // butterfly over the 4 lanes of the tile; lane pg then owns row pg of the 4x4 tile
      float mine[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
          float v = a[r][c];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (r == pg) mine[c] = v;
        }
      if (tvalid) {
        double* o = scp + (size_t)tile * 16 + pg * 4;
#pragma unroll
        for (int c = 0; c < 4; c++) RED_ADD(o + c, (double)mine[c]);
      }
    }
  }
  STAMP(5);

}

template <int P, int ITER, bool MARG>
static void launch_cfg(const BAWinDev& W, const BAIter& it, cudaStream_t s) {
  static bool configured = false;
  const int smem = (int)sizeof(PointSmem<P>);
  if (!configured) {
    cudaFuncSetAttribute(ba_point_kernel<P, ITER, MARG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    configured = true;
  }
  const int WQ = P / (4 * ITER);
  dim3 grid(W.nchunks), block(32 * WQ * (W.nf < 2 ? 2 : W.nf));
  ba_point_kernel<P, ITER, MARG><<<grid, block, smem, s>>>(W, it);
}

void launch_point_kernel(const BAWinDev& W, const BAIter& it, cudaStream_t s) {
  if (W.P == 8) launch_cfg<8, 1, false>(W, it, s);
  else if (W.P == 32) launch_cfg<32, 2, false>(W, it, s);
  else if (W.iter2) launch_cfg<16, 2, false>(W, it, s);
  else launch_cfg<16, 1, false>(W, it, s);
}

// marginalisation launch: same chunking as the production kernel, W.marg / W.marg_mask / W.marg_rtz set, it.have_x = 0
void launch_point_kernel_marg(const BAWinDev& W, const BAIter& it, cudaStream_t s) {
  if (W.P == 8) launch_cfg<8, 1, true>(W, it, s);
  else if (W.P == 32) launch_cfg<32, 2, true>(W, it, s);
  else launch_cfg<16, 1, true>(W, it, s);
}

}  // namespace dmv
