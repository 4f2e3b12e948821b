// device-side makeCoarseDepthL0 (ct_depth.cu): per-level working planes and output lists
#pragma once
#include "../../include/dmvio_b200.h"
#include <cuda_runtime.h>
namespace dmv {
struct CDLevels {
  int levels, cap;
  int w[DMV_MAX_PYR_LEVELS], h[DMV_MAX_PYR_LEVELS];
  float *idepth[DMV_MAX_PYR_LEVELS], *ws[DMV_MAX_PYR_LEVELS], *ws2[DMV_MAX_PYR_LEVELS];
  const float4* img[DMV_MAX_PYR_LEVELS];                       // the reference frame's planes (I, dx, dy, 0)
  float *pc_u[DMV_MAX_PYR_LEVELS], *pc_v[DMV_MAX_PYR_LEVELS], *pc_id[DMV_MAX_PYR_LEVELS], *pc_col[DMV_MAX_PYR_LEVELS];
  int *rowcnt, *rowoff, *totals;                               // scratch rows; totals[l] = pc_n[l] (device-visible pinned host memory)
};
void cd_launch(const CDLevels& L, int n_unique, const int* d_pix, const float* d_idw, const float* d_wsum, cudaStream_t s);
}  // namespace dmv
