// argument block of ip_trace_kernel (ip_trace.cu), shared with the C ABI in ct_kernels.cu
#pragma once
#include "../../include/dmvio_b200.h"
#include <cuda_runtime.h>
namespace dmv {
struct IPTraceArgs {
  int n, w, h;
  float KRKi[9], Kt[3], aff[2];
  dmv_ip_settings s;
  const float *u, *v, *color, *weights, *gradH, *energyTH;
  float *idepth_min, *idepth_max, *quality;
  int* status;
  float *uv, *interval;
  const float4* img;
};
void launch_ip_trace(const IPTraceArgs& A, cudaStream_t s);
}  // namespace dmv
