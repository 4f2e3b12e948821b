// TEST INFRASTRUCTURE ONLY — see orc_trace.h.
#include "orc_trace.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <utility>

namespace orc {

static const int patternP[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};  // util/settings.cpp:L296 (pattern 8)

// util/globalFuncs.h:L203-226 getInterpolatedElement33BiLin: [interpolated I, dx of the bilinear patch, dy of the bilinear patch]
static inline void interp33BiLin(const float* mat, float x, float y, int width, float out[3]) {
  const int ix = (int)x, iy = (int)y;
  const float* bp = mat + 3 * (ix + iy * width);
  const float tl = bp[0], tr = bp[3], bl = bp[3 * width], br = bp[3 * width + 3];
  const float dx = x - ix, dy = y - iy;
  const float topInt = dx * tr + (1 - dx) * tl;
  const float botInt = dx * br + (1 - dx) * bl;
  const float leftInt = dy * bl + (1 - dy) * tl;
  const float rightInt = dy * br + (1 - dy) * tr;
  out[0] = dx * rightInt + (1 - dx) * leftInt;
  out[1] = rightInt - leftInt;
  out[2] = botInt - topInt;
}
// util/globalFuncs.h:L160-175
static inline float interp31(const float* mat, float x, float y, int width) {
  const int ix = (int)x, iy = (int)y;
  const float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  return dxdy * bp[3 * (1 + width)] + (dy - dxdy) * bp[3 * width] + (dx - dxdy) * bp[3] + (1 - dx - dy + dxdy) * bp[0];
}
// util/globalFuncs.h:L103-118
static inline void interp33(const float* mat, float x, float y, int width, float out[3]) {
  const int ix = (int)x, iy = (int)y;
  const float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  for (int c = 0; c < 3; c++)
    out[c] = dxdy * bp[3 * (1 + width) + c] + (dy - dxdy) * bp[3 * width + c] + (dx - dxdy) * bp[3 + c] + (1 - dx - dy + dxdy) * bp[c];
}

bool initImmature(ImmaturePt& p, const float* dI, int w, int u, int v, const TraceSettings& s) {  // ImmaturePoint.cpp:L34-63
  p.u = (float)u; p.v = (float)v;
  p.idepth_min = 0; p.idepth_max = NAN; p.lastTraceStatus = IPS_UNINITIALIZED;
  p.gradH[0] = p.gradH[1] = p.gradH[2] = p.gradH[3] = 0;
  p.quality = 10000;
  p.lastTraceUV[0] = p.lastTraceUV[1] = 0; p.lastTracePixelInterval = 0;
  for (int idx = 0; idx < 8; idx++) {
    float ptc[3];
    interp33BiLin(dI, (float)(u + patternP[idx][0]), (float)(v + patternP[idx][1]), w, ptc);
    p.color[idx] = ptc[0];
    if (!std::isfinite(p.color[idx])) { p.energyTH = NAN; return false; }
    p.gradH[0] += ptc[1] * ptc[1]; p.gradH[1] += ptc[1] * ptc[2]; p.gradH[2] += ptc[2] * ptc[1]; p.gradH[3] += ptc[2] * ptc[2];
    p.weights[idx] = sqrtf(s.outlierTHSumComponent / (s.outlierTHSumComponent + (ptc[1] * ptc[1] + ptc[2] * ptc[2])));
  }
  p.energyTH = 8 * s.outlierTH;
  p.energyTH *= s.overallEnergyTHWeight * s.overallEnergyTHWeight;
  return true;
}

int traceOn(ImmaturePt& P, const float* dI, int w, int h, const float KRKi[9], const float Kt[3], const float aff[2], const TraceSettings& s) {
  if (P.lastTraceStatus == IPS_OOB) return P.lastTraceStatus;  // L79
  const float maxPixSearch = (w + h) * s.maxPixSearch;
  auto oob = [&]() { P.lastTraceUV[0] = -1; P.lastTraceUV[1] = -1; P.lastTracePixelInterval = 0; return P.lastTraceStatus = IPS_OOB; };
  // ---- project min and max (L98-176)
  float pr[3];
  for (int i = 0; i < 3; i++) pr[i] = (KRKi[3 * i] * P.u + KRKi[3 * i + 1] * P.v) + KRKi[3 * i + 2] * 1.0f;
  float ptpMin[3];
  for (int i = 0; i < 3; i++) ptpMin[i] = pr[i] + Kt[i] * P.idepth_min;
  const float uMin = ptpMin[0] / ptpMin[2], vMin = ptpMin[1] / ptpMin[2];
  int maxRotPatX = 0, maxRotPatY = 0;
  float rot[8][2];
  for (int idx = 0; idx < 8; idx++) {
    const float px = (float)patternP[idx][0], py = (float)patternP[idx][1];
    rot[idx][0] = KRKi[0] * px + KRKi[1] * py;
    rot[idx][1] = KRKi[3] * px + KRKi[4] * py;
    const int absX = (int)std::abs(rot[idx][0]), absY = (int)std::abs(rot[idx][1]);
    maxRotPatX = std::max(absX, maxRotPatX);
    maxRotPatY = std::max(absY, maxRotPatY);
  }
  const int boundU = std::max(4, maxRotPatX + 2), boundV = std::max(4, maxRotPatY + 2);
  if (!(uMin > boundU && vMin > boundV && uMin < w - boundU - 1 && vMin < h - boundV - 1)) return oob();
  float dist, uMax, vMax, ptpMax[3];
  if (std::isfinite(P.idepth_max)) {
    for (int i = 0; i < 3; i++) ptpMax[i] = pr[i] + Kt[i] * P.idepth_max;
    uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
    if (!(uMax > boundU && vMax > boundV && uMax < w - boundU - 1 && vMax < h - boundV - 1)) return oob();
    dist = (uMin - uMax) * (uMin - uMax) + (vMin - vMax) * (vMin - vMax);
    dist = sqrtf(dist);
    if (dist < s.trace_slackInterval) {
      P.lastTraceUV[0] = (uMax + uMin) * 0.5f; P.lastTraceUV[1] = (vMax + vMin) * 0.5f;
      P.lastTracePixelInterval = dist;
      return P.lastTraceStatus = IPS_SKIPPED;
    }
  } else {
    dist = maxPixSearch;
    for (int i = 0; i < 3; i++) ptpMax[i] = pr[i] + Kt[i] * 0.01f;
    uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
    const float ddx = uMax - uMin, ddy = vMax - vMin;
    const float d = 1.0f / sqrtf(ddx * ddx + ddy * ddy);
    uMax = uMin + dist * ddx * d;
    vMax = vMin + dist * ddy * d;
    if (!(uMax > boundU && vMax > boundV && uMax < w - boundU - 1 && vMax < h - boundV - 1)) return oob();
  }
  if (!(P.idepth_min < 0 || (ptpMin[2] > 0.75 && ptpMin[2] < 1.5))) return oob();  // L179-185 (double literals: compared in double)
  // ---- error bounds (L188-206)
  float dx = s.trace_stepsize * (uMax - uMin);
  float dy = s.trace_stepsize * (vMax - vMin);
  const float* g = P.gradH;
  const float a = (dx * g[0] + dy * g[2]) * dx + (dx * g[1] + dy * g[3]) * dy;
  const float b = (dy * g[0] + (-dx) * g[2]) * dy + (dy * g[1] + (-dx) * g[3]) * (-dx);
  float errorInPixel = 0.2f + 0.2f * (a + b) / a;
  if (errorInPixel * s.trace_minImprovementFactor > dist && std::isfinite(P.idepth_max)) {
    P.lastTraceUV[0] = (uMax + uMin) * 0.5f; P.lastTraceUV[1] = (vMax + vMin) * 0.5f;
    P.lastTracePixelInterval = dist;
    return P.lastTraceStatus = IPS_BADCONDITION;
  }
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
  if (!std::isfinite(dx) || !std::isfinite(dy)) { P.lastTracePixelInterval = 0; P.lastTraceUV[0] = -1; P.lastTraceUV[1] = -1; return P.lastTraceStatus = IPS_OOB; }
  float errors[100];
  float bestU = 0, bestV = 0, bestEnergy = 1e10;
  int bestIdx = -1;
  if (numSteps >= 100) numSteps = 99;
  for (int i = 0; i < numSteps; i++) {
    float energy = 0;
    for (int idx = 0; idx < 8; idx++) {
      const float hitColor = interp31(dI, (float)(ptx + rot[idx][0]), (float)(pty + rot[idx][1]), w);
      if (!std::isfinite(hitColor)) { energy += 1e5; continue; }
      const float residual = hitColor - (float)(aff[0] * P.color[idx] + aff[1]);
      const float hw = fabs(residual) < s.huberTH ? 1 : s.huberTH / fabs(residual);
      energy += hw * residual * residual * (2 - hw);
    }
    errors[i] = energy;
    if (energy < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = energy; bestIdx = i; }
    ptx += dx;
    pty += dy;
  }
  float secondBest = 1e10;
  for (int i = 0; i < numSteps; i++)
    if ((i < bestIdx - s.minTraceTestRadius || i > bestIdx + s.minTraceTestRadius) && errors[i] < secondBest) secondBest = errors[i];
  const float newQuality = secondBest / bestEnergy;
  if (newQuality < P.quality || numSteps > 10) P.quality = newQuality;
  // ---- GN refinement along the line (L280-353)
  float uBak = bestU, vBak = bestV, gnstepsize = 1, stepBack = 0;
  if (s.trace_GNIterations > 0) bestEnergy = 1e5;
  for (int it = 0; it < s.trace_GNIterations; it++) {
    float H = 1, bb = 0, energy = 0;
    for (int idx = 0; idx < 8; idx++) {
      const float posU = (float)(bestU + rot[idx][0]);
      const float posV = (float)(bestV + rot[idx][1]);
      if (posU < 0 || posV < 0 || posU >= w - 1 || posV >= h - 1) return oob();
      float hit[3];
      interp33(dI, posU, posV, w, hit);
      if (!std::isfinite((float)hit[0])) { energy += 1e5; continue; }
      const float residual = hit[0] - (aff[0] * P.color[idx] + aff[1]);
      const float dResdDist = dx * hit[1] + dy * hit[2];
      const float hw = fabs(residual) < s.huberTH ? 1 : s.huberTH / fabs(residual);
      H += hw * dResdDist * dResdDist;
      bb += hw * residual * dResdDist;
      energy += P.weights[idx] * P.weights[idx] * hw * residual * residual * (2 - hw);
    }
    if (energy > bestEnergy) {
      stepBack *= 0.5;
      bestU = uBak + stepBack * dx;
      bestV = vBak + stepBack * dy;
    } else {
      float step = -gnstepsize * bb / H;
      if (step < -0.5) step = -0.5;
      else if (step > 0.5) step = 0.5;
      if (!std::isfinite(step)) step = 0;
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
  if (!(bestEnergy < P.energyTH * s.trace_extraSlackOnTH)) {
    P.lastTracePixelInterval = 0;
    P.lastTraceUV[0] = -1; P.lastTraceUV[1] = -1;
    if (P.lastTraceStatus == IPS_OUTLIER) return P.lastTraceStatus = IPS_OOB;
    return P.lastTraceStatus = IPS_OUTLIER;
  }
  // ---- new interval (L380-402)
  if (dx * dx > dy * dy) {
    P.idepth_min = (pr[2] * (bestU - errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU - errorInPixel * dx));
    P.idepth_max = (pr[2] * (bestU + errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU + errorInPixel * dx));
  } else {
    P.idepth_min = (pr[2] * (bestV - errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV - errorInPixel * dy));
    P.idepth_max = (pr[2] * (bestV + errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV + errorInPixel * dy));
  }
  if (P.idepth_min > P.idepth_max) std::swap(P.idepth_min, P.idepth_max);
  if (!std::isfinite(P.idepth_min) || !std::isfinite(P.idepth_max) || (P.idepth_max < 0)) {
    P.lastTracePixelInterval = 0;
    P.lastTraceUV[0] = -1; P.lastTraceUV[1] = -1;
    return P.lastTraceStatus = IPS_OUTLIER;
  }
  P.lastTracePixelInterval = 2 * errorInPixel;
  P.lastTraceUV[0] = bestU; P.lastTraceUV[1] = bestV;
  return P.lastTraceStatus = IPS_GOOD;
}

}  // namespace orc

using namespace orc;
extern "C" {
int orc_ip_init(int n, const float* dI, int w, int h, const int32_t* u, const int32_t* v, float* color8, float* weights8, float* gradH4, float* energyTH,
                uint8_t* ok) {
  (void)h;
  TraceSettings s;
  int good = 0;
  for (int i = 0; i < n; i++) {
    ImmaturePt p;
    const bool o = initImmature(p, dI, w, u[i], v[i], s);
    std::memcpy(color8 + 8 * i, p.color, 32); std::memcpy(weights8 + 8 * i, p.weights, 32); std::memcpy(gradH4 + 4 * i, p.gradH, 16);
    energyTH[i] = p.energyTH;
    ok[i] = o; good += o;
  }
  return good;
}
void orc_ip_trace(int n, const float* dI, int w, int h, const float* KRKi, const float* Kt, const float* aff, const float* u, const float* v,
                  const float* color8, const float* weights8, const float* gradH4, const float* energyTH, float* idepth_min, float* idepth_max,
                  float* quality, int32_t* status, float* uv2, float* interval) {
  TraceSettings s;
  for (int i = 0; i < n; i++) {
    ImmaturePt p;
    p.u = u[i]; p.v = v[i];
    std::memcpy(p.color, color8 + 8 * i, 32); std::memcpy(p.weights, weights8 + 8 * i, 32); std::memcpy(p.gradH, gradH4 + 4 * i, 16);
    p.energyTH = energyTH[i]; p.idepth_min = idepth_min[i]; p.idepth_max = idepth_max[i]; p.quality = quality[i];
    p.lastTraceStatus = status[i]; p.lastTraceUV[0] = uv2[2 * i]; p.lastTraceUV[1] = uv2[2 * i + 1]; p.lastTracePixelInterval = interval[i];
    traceOn(p, dI, w, h, KRKi, Kt, aff, s);
    idepth_min[i] = p.idepth_min; idepth_max[i] = p.idepth_max; quality[i] = p.quality; status[i] = p.lastTraceStatus;
    uv2[2 * i] = p.lastTraceUV[0]; uv2[2 * i + 1] = p.lastTraceUV[1]; interval[i] = p.lastTracePixelInterval;
  }
}
}
