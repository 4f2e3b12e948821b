// Coarse-tracking reference construction on the device: CoarseTracker::makeCoarseDepthL0
// (reference FullSystem/CoarseTracker.cpp:L138-295) — weighted splat of the keyframe's IN residuals, 2x2 sum pooling to the coarser
// levels, one dilation step per level (diagonal neighbours on levels 0-1, 4-neighbourhood above), normalisation and ORDER-PRESERVING
// compaction into the pc_u / pc_v / pc_idepth / pc_color lists that calcRes reads.  Bit-identical to the CPU code:
//   * the splat is the only order-dependent float sum (points that hit the same pixel): the host folds colliding points in input order
//     (exactly the reference's sequential +=) and the device scatters collision-free entries;
//   * dilation reads only originally-valid pixels and writes only originally-empty ones, so the in-place sweep of the reference is
//     order-independent; pooling / dilation keep the reference's operand order; there is no multiply-add to contract;
//   * compaction = per-row counts, an exclusive scan over the rows, ballot-prefix writes: row-major order like the reference's loops.
#include "../../include/dmvio_b200.h"
#include "common_host.h"
#include "ct_depth.h"

namespace dmv {

__global__ void cd_scatter_kernel(int n, const int* __restrict__ pix, const float* __restrict__ idw, const float* __restrict__ wsum, float* idepth0, float* ws0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { idepth0[pix[i]] = idw[i]; ws0[pix[i]] = wsum[i]; }
}

__global__ void cd_pool_kernel(const float* __restrict__ id_lm, const float* __restrict__ ws_lm, float* id_l, float* ws_l, int wl, int hl, int wlm1) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= wl || y >= hl) return;
  const int b = 2 * x + 2 * y * wlm1;
  id_l[x + y * wl] = id_lm[b] + id_lm[b + 1] + id_lm[b + wlm1] + id_lm[b + wlm1 + 1];
  ws_l[x + y * wl] = ws_lm[b] + ws_lm[b + 1] + ws_lm[b + wlm1] + ws_lm[b + wlm1 + 1];
}

// ws_bak = weights before the sweep (read only), ws_out = weights after it, idl updated in place (see the file header)
__global__ void cd_dilate_kernel(const float* __restrict__ ws_bak, float* __restrict__ ws_out, float* idl, int wl, int hl, int diagonal) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int npx = wl * hl;
  if (i >= npx) return;
  float wv = ws_bak[i];
  const int wh = npx - wl;
  if (i >= wl + 1 && i < wh - 1 && wv <= 0) {
    const int o0 = diagonal ? 1 + wl : 1, o1 = diagonal ? -1 - wl : -1, o2 = diagonal ? wl - 1 : wl, o3 = diagonal ? -wl + 1 : -wl;
    float sum = 0, num = 0, numn = 0;
    if (ws_bak[i + o0] > 0) { sum += idl[i + o0]; num += ws_bak[i + o0]; numn++; }
    if (ws_bak[i + o1] > 0) { sum += idl[i + o1]; num += ws_bak[i + o1]; numn++; }
    if (ws_bak[i + o2] > 0) { sum += idl[i + o2]; num += ws_bak[i + o2]; numn++; }
    if (ws_bak[i + o3] > 0) { sum += idl[i + o3]; num += ws_bak[i + o3]; numn++; }
    if (numn > 0) { idl[i] = sum / numn; wv = num / numn; }
  }
  ws_out[i] = wv;
}

__device__ __forceinline__ bool cd_valid(const float* ws, const float* idl, const float4* img, int i, float& q, float& col) {
  const float w = ws[i];
  if (!(w > 0)) return false;
  q = idl[i] / w;
  col = img[i].x;
  return isfinite(col) && (q > 0);
}

// one warp per row y in [2, hl-2): number of list entries of the row
__global__ void cd_rowcount_kernel(const float* __restrict__ ws, const float* __restrict__ idl, const float4* __restrict__ img, int wl, int hl, int* rowcnt) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int y = row + 2;
  if (y >= hl - 2) return;
  int cnt = 0;
  for (int x0 = 2; x0 < wl - 2; x0 += 32) {
    const int x = x0 + lane;
    float q, col;
    const bool v = (x < wl - 2) && cd_valid(ws, idl, img, x + y * wl, q, col);
    cnt += __popc(__ballot_sync(0xffffffffu, v));
  }
  if (lane == 0) rowcnt[row] = cnt;
}

// exclusive scan of the row counts by one block; total -> *total
__global__ void cd_scan_kernel(const int* __restrict__ rowcnt, int* __restrict__ rowoff, int nrows, int* total) {
  __shared__ int s[1024];
  const int tid = threadIdx.x;
  int carry = 0;
  for (int base = 0; base < nrows; base += 1024) {
    const int v = (base + tid < nrows) ? rowcnt[base + tid] : 0;
    s[tid] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      const int t = (tid >= d) ? s[tid - d] : 0;
      __syncthreads();
      s[tid] += t;
      __syncthreads();
    }
    if (base + tid < nrows) rowoff[base + tid] = carry + s[tid] - v;
    carry += s[1023];
    __syncthreads();
  }
  if (tid == 0) *total = carry;
}

__global__ void cd_write_kernel(const float* __restrict__ ws, const float* __restrict__ idl, const float4* __restrict__ img, int wl, int hl,
                                const int* __restrict__ rowoff, int cap, float* pc_u, float* pc_v, float* pc_id, float* pc_col) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int y = row + 2;
  if (y >= hl - 2) return;
  int off = rowoff[row];
  for (int x0 = 2; x0 < wl - 2; x0 += 32) {
    const int x = x0 + lane;
    float q = 0, col = 0;
    const bool v = (x < wl - 2) && cd_valid(ws, idl, img, x + y * wl, q, col);
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (v) {
      const int o = off + __popc(m & ((1u << lane) - 1u));
      if (o < cap) { pc_u[o] = (float)x; pc_v[o] = (float)y; pc_id[o] = q; pc_col[o] = col; }
    }
    off += __popc(m);
  }
}

void cd_launch(const CDLevels& L, int n_unique, const int* d_pix, const float* d_idw, const float* d_wsum, cudaStream_t s) {
  cudaMemsetAsync(L.idepth[0], 0, sizeof(float) * L.w[0] * L.h[0], s);
  cudaMemsetAsync(L.ws[0], 0, sizeof(float) * L.w[0] * L.h[0], s);
  if (n_unique > 0) cd_scatter_kernel<<<(n_unique + 255) / 256, 256, 0, s>>>(n_unique, d_pix, d_idw, d_wsum, L.idepth[0], L.ws[0]);
  for (int l = 1; l < L.levels; l++) {
    dim3 grid((L.w[l] + 127) / 128, L.h[l]);
    cd_pool_kernel<<<grid, 128, 0, s>>>(L.idepth[l - 1], L.ws[l - 1], L.idepth[l], L.ws[l], L.w[l], L.h[l], L.w[l - 1]);
  }
  for (int l = 0; l < L.levels; l++) {
    const int npx = L.w[l] * L.h[l];
    cd_dilate_kernel<<<(npx + 255) / 256, 256, 0, s>>>(L.ws[l], L.ws2[l], L.idepth[l], L.w[l], L.h[l], l < 2 ? 1 : 0);
    const int nrows = L.h[l] - 4;
    if (nrows > 0) {
      cd_rowcount_kernel<<<(nrows + 7) / 8, 256, 0, s>>>(L.ws2[l], L.idepth[l], L.img[l], L.w[l], L.h[l], L.rowcnt);
      cd_scan_kernel<<<1, 1024, 0, s>>>(L.rowcnt, L.rowoff, nrows, L.totals + l);
      cd_write_kernel<<<(nrows + 7) / 8, 256, 0, s>>>(L.ws2[l], L.idepth[l], L.img[l], L.w[l], L.h[l], L.rowoff, L.cap, L.pc_u[l], L.pc_v[l], L.pc_id[l], L.pc_col[l]);
    } else {
      cudaMemsetAsync(L.totals + l, 0, sizeof(int), s);
    }
  }
}

}  // namespace dmv
