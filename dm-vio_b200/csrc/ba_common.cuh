// Device helpers shared by the BA kernels.
#pragma once
#include "ba_device.cuh"
#include <math.h>

namespace dmv {

static __device__ __constant__ int c_pattern[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

__device__ __forceinline__ float pick8(const float* v, int j) {
  float a = (j & 1) ? v[1] : v[0];
  float b = (j & 1) ? v[3] : v[2];
  float c = (j & 1) ? v[5] : v[4];
  float d = (j & 1) ? v[7] : v[6];
  float e = (j & 2) ? b : a;
  float f = (j & 2) ? d : c;
  return (j & 4) ? f : e;
}
__device__ __forceinline__ float group_sum8(float v) {  // all-reduce inside aligned groups of 8 lanes
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float cross_group_sum(float v) {  // sum over the 4 groups of a warp
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(g));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }  // result unused -> RED.E.ADD.F64
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }



// EnergyFunctional::resubstituteFPt for one point (EnergyFunctional.cpp:L295-321)
__device__ __forceinline__ float resub_point(const BAWinDev& W, const BAIter& it, int p, int h) {
  const int nf = W.nf, mp = W.mp;
  const float4 po0 = __ldg(reinterpret_cast<const float4*>(W.c_pout + (size_t)p * 8));
  const float4 po1 = __ldg(reinterpret_cast<const float4*>(W.c_pout + (size_t)p * 8) + 1);
  float b = po1.w;  // bdSumF
  b -= it.xc[0] * po0.z + it.xc[1] * po0.w + it.xc[2] * po1.x + it.xc[3] * po1.y;
  int ngood = 0;
  for (int t = 0; t < nf; t++) {
    if (t == h) continue;
    const int slot = t * mp + p;
    if (W.c_st[slot] != RES_IN) continue;
    ngood++;
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(W.c_jpjd + (size_t)slot * 8));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(W.c_jpjd + (size_t)slot * 8) + 1);
    const float* xa = it.xAd[h * nf + t];
    b -= xa[0] * a0.x + xa[1] * a0.y + xa[2] * a0.z + xa[3] * a0.w + xa[4] * a1.x + xa[5] * a1.y + xa[6] * a1.z + xa[7] * a1.w;
  }
  return ngood > 0 ? -b * po1.z : 0.f;  // step = -b * HdiF
}


}  // namespace dmv
