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

void launch_ip_trace(const IPTraceArgs& A, cudaStream_t s) { ip_trace_kernel<<<(A.n + 127) / 128, 128, 0, s>>>(A); }

}  // namespace dmv
