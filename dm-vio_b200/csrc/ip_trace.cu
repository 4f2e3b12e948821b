// ip_trace_kernel — ImmaturePoint::traceOn (reference FullSystem/ImmaturePoint.cpp:L77-437) for all immature points of one host
// frame against the newest frame, as FullSystem::traceNewCoarse does on every tracked frame (FullSystem.cpp:L541-584).
//
// One thread per point: the epipolar search is a sequential, data-dependent scalar chain (<= 99 steps x 8 bilinear samples, then <= 3
// Gauss-Newton steps), embarrassingly parallel over the ~1500 candidates of a frame.  The new frame's level-0 plane is the float4
// texel plane already resident in the coarse-tracker handle (uploaded once per frame for tracking).
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: every expression keeps the reference's operation order and rounding (no FMA
// contraction; IEEE division and sqrt are nvcc's defaults), so the integer decisions (best step, status, GN accept/reject) and the
// float outputs are bit-identical to the reference's CPU code — tests/test_gpu_trace.py asserts equality, not a tolerance.
#include "../../include/dmvio_b200.h"
#include "common_host.h"
#include "ip_trace.h"
#include <cuda_runtime.h>
#include <math.h>

namespace dmv {

__device__ __constant__ int c_ip_pattern[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

// util/globalFuncs.h:L160-175 getInterpolatedElement31
__device__ __forceinline__ float ip_interp31(const float4* __restrict__ mat, float x, float y, int width) {
  const int ix = (int)x, iy = (int)y;
  const float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float4* bp = mat + ix + iy * width;
  return dxdy * __ldg(bp + 1 + width).x + (dy - dxdy) * __ldg(bp + width).x + (dx - dxdy) * __ldg(bp + 1).x + (1 - dx - dy + dxdy) * __ldg(bp).x;
}
// util/globalFuncs.h:L103-118 getInterpolatedElement33
__device__ __forceinline__ void ip_interp33(const float4* __restrict__ mat, float x, float y, int width, float out[3]) {
  const int ix = (int)x, iy = (int)y;
  const float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float4* bp = mat + ix + iy * width;
  const float4 br = __ldg(bp + 1 + width), bl = __ldg(bp + width), tr = __ldg(bp + 1), tl = __ldg(bp);
  out[0] = dxdy * br.x + (dy - dxdy) * bl.x + (dx - dxdy) * tr.x + (1 - dx - dy + dxdy) * tl.x;
  out[1] = dxdy * br.y + (dy - dxdy) * bl.y + (dx - dxdy) * tr.y + (1 - dx - dy + dxdy) * tl.y;
  out[2] = dxdy * br.z + (dy - dxdy) * bl.z + (dx - dxdy) * tr.z + (1 - dx - dy + dxdy) * tl.z;
}

__global__ void __launch_bounds__(128) ip_trace_kernel(const __grid_constant__ IPTraceArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const dmv_ip_settings& s = A.s;
  const int w = A.w, h = A.h;
  const float4* __restrict__ dI = A.img;
  const float* KRKi = A.KRKi;
  const float* Kt = A.Kt;
  int status = A.status[i];
  if (status == 1) return;  // IPS_OOB stays OOB (L79)
  const float u = A.u[i], v = A.v[i];
  float idepth_min = A.idepth_min[i], idepth_max = A.idepth_max[i];
  float color[8], weights[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { color[k] = A.color[8 * i + k]; weights[k] = A.weights[8 * i + k]; }
  const float g0 = A.gradH[4 * i], g1 = A.gradH[4 * i + 1], g2 = A.gradH[4 * i + 2], g3 = A.gradH[4 * i + 3];
  const float energyTH = A.energyTH[i];
  float uvx, uvy, interval;
#define IP_RETURN(st, ux, uy, itv) do { A.status[i] = (st); A.uv[2 * i] = (ux); A.uv[2 * i + 1] = (uy); A.interval[i] = (itv); return; } while (0)
  const float maxPixSearch = (w + h) * s.maxPixSearch;
  // ---- project min and max (L98-176)
  float pr[3];
#pragma unroll
  for (int k = 0; k < 3; k++) pr[k] = (KRKi[3 * k] * u + KRKi[3 * k + 1] * v) + KRKi[3 * k + 2] * 1.0f;
  float ptpMin[3];
#pragma unroll
  for (int k = 0; k < 3; k++) ptpMin[k] = pr[k] + Kt[k] * idepth_min;
  const float uMin = ptpMin[0] / ptpMin[2], vMin = ptpMin[1] / ptpMin[2];
  int maxRotPatX = 0, maxRotPatY = 0;
  float rot[8][2];
#pragma unroll
  for (int idx = 0; idx < 8; idx++) {
    const float px = (float)c_ip_pattern[idx][0], py = (float)c_ip_pattern[idx][1];
    rot[idx][0] = KRKi[0] * px + KRKi[1] * py;
    rot[idx][1] = KRKi[3] * px + KRKi[4] * py;
    maxRotPatX = max((int)fabsf(rot[idx][0]), maxRotPatX);
    maxRotPatY = max((int)fabsf(rot[idx][1]), maxRotPatY);
  }
  const int boundU = max(4, maxRotPatX + 2), boundV = max(4, maxRotPatY + 2);
  if (!(uMin > boundU && vMin > boundV && uMin < w - boundU - 1 && vMin < h - boundV - 1)) IP_RETURN(1, -1.f, -1.f, 0.f);
  float dist, uMax, vMax, ptpMax[3];
  if (isfinite(idepth_max)) {
#pragma unroll
    for (int k = 0; k < 3; k++) ptpMax[k] = pr[k] + Kt[k] * idepth_max;
    uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
    if (!(uMax > boundU && vMax > boundV && uMax < w - boundU - 1 && vMax < h - boundV - 1)) IP_RETURN(1, -1.f, -1.f, 0.f);
    dist = (uMin - uMax) * (uMin - uMax) + (vMin - vMax) * (vMin - vMax);
    dist = sqrtf(dist);
    if (dist < s.trace_slackInterval) IP_RETURN(3, (uMax + uMin) * 0.5f, (vMax + vMin) * 0.5f, dist);
  } else {
    dist = maxPixSearch;
#pragma unroll
    for (int k = 0; k < 3; k++) ptpMax[k] = pr[k] + Kt[k] * 0.01f;
    uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
    const float ddx = uMax - uMin, ddy = vMax - vMin;
    const float d = 1.0f / sqrtf(ddx * ddx + ddy * ddy);
    uMax = uMin + dist * ddx * d;
    vMax = vMin + dist * ddy * d;
    if (!(uMax > boundU && vMax > boundV && uMax < w - boundU - 1 && vMax < h - boundV - 1)) IP_RETURN(1, -1.f, -1.f, 0.f);
  }
  if (!(idepth_min < 0 || ((double)ptpMin[2] > 0.75 && (double)ptpMin[2] < 1.5))) IP_RETURN(1, -1.f, -1.f, 0.f);  // L179-185
  // ---- error bounds (L188-206)
  float dx = s.trace_stepsize * (uMax - uMin);
  float dy = s.trace_stepsize * (vMax - vMin);
  const float a = (dx * g0 + dy * g2) * dx + (dx * g1 + dy * g3) * dy;
  const float b = (dy * g0 + (-dx) * g2) * dy + (dy * g1 + (-dx) * g3) * (-dx);
  float errorInPixel = 0.2f + 0.2f * (a + b) / a;
  if (errorInPixel * s.trace_minImprovementFactor > dist && isfinite(idepth_max)) IP_RETURN(4, (uMax + uMin) * 0.5f, (vMax + vMin) * 0.5f, dist);
  if (errorInPixel > 10) errorInPixel = 10;
  // ---- discrete search (L210-277)
  dx /= dist;
  dy /= dist;
  if (dist > maxPixSearch) {
    uMax = uMin + maxPixSearch * dx;
    vMax = vMin + maxPixSearch * dy;
    dist = maxPixSearch;
  }
  int numSteps = 1.9999f + dist / s.trace_stepsize;
  const float randShift = uMin * 1000 - floorf(uMin * 1000);
  float ptx = uMin - randShift * dx;
  float pty = vMin - randShift * dy;
  if (!isfinite(dx) || !isfinite(dy)) IP_RETURN(1, -1.f, -1.f, 0.f);
  float errors[100];
  float bestU = 0, bestV = 0, bestEnergy = 1e10f;
  int bestIdx = -1;
  if (numSteps >= 100) numSteps = 99;
  for (int st = 0; st < numSteps; st++) {
    float energy = 0;
#pragma unroll
    for (int idx = 0; idx < 8; idx++) {
      const float hitColor = ip_interp31(dI, (float)(ptx + rot[idx][0]), (float)(pty + rot[idx][1]), w);
      if (!isfinite(hitColor)) { energy = (float)((double)energy + 1e5); continue; }
      const float residual = hitColor - (float)(A.aff[0] * color[idx] + A.aff[1]);
      const float hw = fabsf(residual) < s.huberTH ? 1 : s.huberTH / fabsf(residual);
      energy += hw * residual * residual * (2 - hw);
    }
    errors[st] = energy;
    if (energy < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = energy; bestIdx = st; }
    ptx += dx;
    pty += dy;
  }
  float secondBest = 1e10f;
  for (int st = 0; st < numSteps; st++)
    if ((st < bestIdx - s.minTraceTestRadius || st > bestIdx + s.minTraceTestRadius) && errors[st] < secondBest) secondBest = errors[st];
  const float newQuality = secondBest / bestEnergy;
  float quality = A.quality[i];
  if (newQuality < quality || numSteps > 10) quality = newQuality;
  A.quality[i] = quality;
  // ---- GN refinement along the line (L280-353)
  float uBak = bestU, vBak = bestV, stepBack = 0;
  const float gnstepsize = 1;
  if (s.trace_GNIterations > 0) bestEnergy = 1e5f;
  for (int it = 0; it < s.trace_GNIterations; it++) {
    float H = 1, bb = 0, energy = 0;
    for (int idx = 0; idx < 8; idx++) {
      const float posU = (float)(bestU + rot[idx][0]);
      const float posV = (float)(bestV + rot[idx][1]);
      if (posU < 0 || posV < 0 || posU >= w - 1 || posV >= h - 1) IP_RETURN(1, -1.f, -1.f, 0.f);
      float hit[3];
      ip_interp33(dI, posU, posV, w, hit);
      if (!isfinite(hit[0])) { energy = (float)((double)energy + 1e5); continue; }
      const float residual = hit[0] - (A.aff[0] * color[idx] + A.aff[1]);
      const float dResdDist = dx * hit[1] + dy * hit[2];
      const float hw = fabsf(residual) < s.huberTH ? 1 : s.huberTH / fabsf(residual);
      H += hw * dResdDist * dResdDist;
      bb += hw * residual * dResdDist;
      energy += weights[idx] * weights[idx] * hw * residual * residual * (2 - hw);
    }
    if (energy > bestEnergy) {
      stepBack = (float)((double)stepBack * 0.5);
      bestU = uBak + stepBack * dx;
      bestV = vBak + stepBack * dy;
    } else {
      float step = -gnstepsize * bb / H;
      if (step < -0.5f) step = -0.5f;
      else if (step > 0.5f) step = 0.5f;
      if (!isfinite(step)) step = 0;
      uBak = bestU;
      vBak = bestV;
      stepBack = step;
      bestU += step * dx;
      bestV += step * dy;
      bestEnergy = energy;
    }
    if (fabsf(stepBack) < s.trace_GNThreshold) break;
  }
  // ---- energy-based outlier (L360-376)
  if (!(bestEnergy < energyTH * s.trace_extraSlackOnTH)) IP_RETURN(status == 2 ? 1 : 2, -1.f, -1.f, 0.f);
  // ---- new interval (L380-402)
  if (dx * dx > dy * dy) {
    idepth_min = (pr[2] * (bestU - errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU - errorInPixel * dx));
    idepth_max = (pr[2] * (bestU + errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU + errorInPixel * dx));
  } else {
    idepth_min = (pr[2] * (bestV - errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV - errorInPixel * dy));
    idepth_max = (pr[2] * (bestV + errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV + errorInPixel * dy));
  }
  if (idepth_min > idepth_max) { const float tmp = idepth_min; idepth_min = idepth_max; idepth_max = tmp; }
  A.idepth_min[i] = idepth_min;   // the reference assigns the members before the final validity test
  A.idepth_max[i] = idepth_max;
  if (!isfinite(idepth_min) || !isfinite(idepth_max) || (idepth_max < 0)) IP_RETURN(2, -1.f, -1.f, 0.f);
  uvx = bestU; uvy = bestV; interval = 2 * errorInPixel;
  IP_RETURN(0, uvx, uvy, interval);
#undef IP_RETURN
}

// ImmaturePoint::ImmaturePoint (ImmaturePoint.cpp:L34-63): pattern colours, gradient weights, gradH and energyTH from the HOST frame's level-0
// plane (getInterpolatedElement33BiLin, util/globalFuncs.h:L203-226) for integer pixels (u, v); ok = 0 where a colour is not finite
// (the reference then leaves energyTH = NaN and the caller drops the point)
__global__ void __launch_bounds__(128) ip_init_kernel(const __grid_constant__ IPInitArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const int u = A.u[i], v = A.v[i];
  float g0 = 0, g1 = 0, g2 = 0, g3 = 0;
  bool ok = true;
  float col[8], wgt[8];
#pragma unroll
  for (int idx = 0; idx < 8; idx++) {
    col[idx] = 0.f; wgt[idx] = 0.f;
    if (!ok) continue;
    const float x = (float)(u + c_ip_pattern[idx][0]), y = (float)(v + c_ip_pattern[idx][1]);
    const int ix = (int)x, iy = (int)y;
    const float4* bp = A.img + ix + iy * A.w;
    const float tl = __ldg(bp).x, tr = __ldg(bp + 1).x, bl = __ldg(bp + A.w).x, br = __ldg(bp + A.w + 1).x;
    const float dx = x - ix, dy = y - iy;
    const float topInt = dx * tr + (1 - dx) * tl;
    const float botInt = dx * br + (1 - dx) * bl;
    const float leftInt = dy * bl + (1 - dy) * tl;
    const float rightInt = dy * br + (1 - dy) * tr;
    const float c0 = dx * rightInt + (1 - dx) * leftInt, c1 = rightInt - leftInt, c2 = botInt - topInt;
    col[idx] = c0;
    if (!isfinite(c0)) { ok = false; continue; }
    g0 += c1 * c1; g1 += c1 * c2; g2 += c2 * c1; g3 += c2 * c2;
    wgt[idx] = sqrtf(A.outlierTHSumComponent / (A.outlierTHSumComponent + (c1 * c1 + c2 * c2)));
  }
#pragma unroll
  for (int idx = 0; idx < 8; idx++) { A.color[8 * i + idx] = col[idx]; A.weights[8 * i + idx] = wgt[idx]; }
  A.gradH[4 * i] = g0; A.gradH[4 * i + 1] = g1; A.gradH[4 * i + 2] = g2; A.gradH[4 * i + 3] = g3;
  float eth = 8 * A.outlierTH;
  eth *= A.overallEnergyTHWeight * A.overallEnergyTHWeight;
  A.energyTH[i] = ok ? eth : __int_as_float(0x7fc00000);
  A.ok[i] = ok ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// ip_activate_kernel — FullSystem::optimizeImmaturePoint (FullSystem/FullSystemOptPoint.cpp:L51-205) with ImmaturePoint::linearizeResidual
// (FullSystem/ImmaturePoint.cpp:L498-565), projectPoint / derive_idepth (FullSystem/ResidualProjections.h:L36-87): one thread per immature
// point, residuals against every other keyframe of the window, <= 3 damped Gauss-Newton steps on the inverse depth.  Same -fmad=false
// rule as the tracing kernel: bit-identical to the CPU code (float/double promotions of the reference kept explicitly).
// ---------------------------------------------------------------------------------------------------------------------
struct IPTmpRes { int state_state, state_NewState, target; double state_energy, state_NewEnergy; };

__device__ double ip_linearize_residual(const IPActArgs& A, int host, float pu, float pv, const float* color, const float* weights, float energyTH,
                                        float outlierTHSlack, IPTmpRes* tmp, float& Hdd, float& bd, float idepth) {
  if (tmp->state_state == 1) { tmp->state_NewState = 1; return tmp->state_energy; }
  const float* R = A.RT + (size_t)(host * A.nf + tmp->target) * 12;
  const float* t = R + 9;
  const float* affLL = A.aff + (size_t)(host * A.nf + tmp->target) * 2;
  const float4* __restrict__ dIl = A.img[tmp->target];
  const float wM3G = (float)(A.w - 3), hM3G = (float)(A.h - 3);
  float energyLeft = 0;
  for (int idx = 0; idx < 8; idx++) {
    const int dx = c_ip_pattern[idx][0], dy = c_ip_pattern[idx][1];
    const float K0 = (pu + dx - A.cxl) * A.fxli, K1 = (pv + dy - A.cyl) * A.fyli, K2 = 1;
    float ptp[3];
#pragma unroll
    for (int i = 0; i < 3; i++) ptp[i] = ((R[3 * i] * K0 + R[3 * i + 1] * K1) + R[3 * i + 2] * K2) + t[i] * idepth;
    const float drescale = 1.0f / ptp[2];
    if (!(drescale > 0)) { tmp->state_NewState = 1; return tmp->state_energy; }
    const float u = ptp[0] * drescale, v = ptp[1] * drescale;
    const float Ku = u * A.fxl + A.cxl, Kv = v * A.fyl + A.cyl;
    if (!(Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G)) { tmp->state_NewState = 1; return tmp->state_energy; }
    float hit[3];
    ip_interp33(dIl, Ku, Kv, A.w, hit);
    if (!isfinite(hit[0])) { tmp->state_NewState = 1; return tmp->state_energy; }
    const float residual = hit[0] - (affLL[0] * color[idx] + affLL[1]);
    float hw = fabsf(residual) < A.huberTH ? 1 : A.huberTH / fabsf(residual);
    energyLeft += weights[idx] * weights[idx] * hw * residual * residual * (2 - hw);
    const float dxInterp = hit[1] * A.fxl, dyInterp = hit[2] * A.fyl;
    const float d_idepth = (dxInterp * drescale * (t[0] - t[2] * u) + dyInterp * drescale * (t[1] - t[2] * v)) * 1.0f;
    hw *= weights[idx] * weights[idx];
    Hdd += (hw * d_idepth) * d_idepth;
    bd += (hw * residual) * d_idepth;
  }
  if (energyLeft > energyTH * outlierTHSlack) { energyLeft = energyTH * outlierTHSlack; tmp->state_NewState = 2; }
  else tmp->state_NewState = 0;
  tmp->state_NewEnergy = energyLeft;
  return energyLeft;
}

__global__ void __launch_bounds__(128) ip_activate_kernel(const __grid_constant__ IPActArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const int nf = A.nf, host = A.host[i];
  const float pu = A.u[i], pv = A.v[i], energyTH = A.energyTH[i];
  float color[8], weights[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { color[k] = A.color[8 * i + k]; weights[k] = A.weights[8 * i + k]; }
  int* rs = A.res_state + (size_t)i * nf;
  IPTmpRes res[DMV_MAX_FRAMES];
  int nres = 0;
  for (int f = 0; f < nf; f++) {
    rs[f] = 255;
    if (f == host) continue;
    res[nres].state_NewEnergy = res[nres].state_energy = 0;
    res[nres].state_NewState = 2;
    res[nres].state_state = 0;
    res[nres].target = f;
    nres++;
  }
  float lastEnergy = 0, lastHdd = 0, lastbd = 0;
  float currentIdepth = (A.idepth_max[i] + A.idepth_min[i]) * 0.5f;
  for (int k = 0; k < nres; k++) {
    lastEnergy = (float)((double)lastEnergy + ip_linearize_residual(A, host, pu, pv, color, weights, energyTH, 1000.f, res + k, lastHdd, lastbd, currentIdepth));
    res[k].state_state = res[k].state_NewState;
    res[k].state_energy = res[k].state_NewEnergy;
  }
  A.idepth[i] = currentIdepth;
  if (!isfinite(lastEnergy) || lastHdd < A.minIdepthH_act) { A.status[i] = 0; return; }
  float lambda = 0.1f;
  for (int iteration = 0; iteration < A.GNIts; iteration++) {
    float H = lastHdd;
    H *= 1 + lambda;
    const float step = (float)((1.0 / (double)H) * (double)lastbd);
    const float newIdepth = currentIdepth - step;
    float newHdd = 0, newbd = 0, newEnergy = 0;
    for (int k = 0; k < nres; k++)
      newEnergy = (float)((double)newEnergy + ip_linearize_residual(A, host, pu, pv, color, weights, energyTH, 1.f, res + k, newHdd, newbd, newIdepth));
    if (!isfinite(lastEnergy) || newHdd < A.minIdepthH_act) { A.idepth[i] = currentIdepth; A.status[i] = 0; return; }
    if (newEnergy < lastEnergy) {
      currentIdepth = newIdepth;
      lastHdd = newHdd; lastbd = newbd; lastEnergy = newEnergy;
      for (int k = 0; k < nres; k++) { res[k].state_state = res[k].state_NewState; res[k].state_energy = res[k].state_NewEnergy; }
      lambda = (float)((double)lambda * 0.5);
    } else {
      lambda *= 5;
    }
    if ((double)fabsf(step) < 0.0001 * (double)currentIdepth) break;
  }
  A.idepth[i] = currentIdepth;
  if (!isfinite(currentIdepth)) { A.status[i] = -1; return; }
  int numGoodRes = 0;
  for (int k = 0; k < nres; k++) {
    rs[res[k].target] = res[k].state_state;
    if (res[k].state_state == 0) numGoodRes++;
  }
  if (numGoodRes < A.minObs || !isfinite(energyTH)) { A.status[i] = -1; return; }
  A.status[i] = 1;
}

void launch_ip_activate(const IPActArgs& A, cudaStream_t s) { ip_activate_kernel<<<(A.n + 127) / 128, 128, 0, s>>>(A); }

void launch_ip_init(const IPInitArgs& A, cudaStream_t s) { ip_init_kernel<<<(A.n + 127) / 128, 128, 0, s>>>(A); }

void launch_ip_trace(const IPTraceArgs& A, cudaStream_t s) { ip_trace_kernel<<<(A.n + 127) / 128, 128, 0, s>>>(A); }

}  // namespace dmv
