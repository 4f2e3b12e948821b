// sm_100a kernels + C-ABI of the coarse direct-image-alignment path.
//
//   ct_res_gs_kernel   CoarseTracker::calcRes (CoarseTracker.cpp:L361-517) fused with calcGSSSE (L299-356): one thread per
//                      reference point: project, 4-tap float4 gather from the new frame's level plane, Huber residual,
//                      energy / saturation counters / flow indicators and the 45 unique entries of the weighted 9x9
//                      outer product.  Warp-shuffle + shared-memory block reduction, per-block fp64 partials, and the last
//                      block to finish (threadfence + ticket) folds the partials in a fixed order -> deterministic, one launch.
//                      The buf_warped_* lists of the reference are never materialised.
//   pyr_down_kernel / grad_kernel   FrameHessian::makeImages (HessianBlocks.cpp:L128-191) on the device.
#include "../../include/dmvio_b200.h"
#include "common_host.h"
#include "inv3.h"
#include <algorithm>
#include <cstring>
#include <vector>

namespace dmv {

constexpr int CT_NRED = 45 + 8;  // 45 outer-product entries + E, nE, nSat, nWarped, shiftT, shiftRT, shiftNum, pad
constexpr int CT_THREADS = 256;

struct CTParams {
  float RKi[9], t[3], Ki[9];
  float fx, fy, cx, cy;
  float affa, affb, a_gs, b0;
  float cutoff, huber, maxEnergy;
  int w, h, n, lvl, want_gs;
};

__global__ void __launch_bounds__(CT_THREADS) ct_res_gs_kernel(CTParams P, const float* __restrict__ pc_u, const float* __restrict__ pc_v,
                                                               const float* __restrict__ pc_id, const float* __restrict__ pc_col,
                                                               const float4* __restrict__ img, double* __restrict__ partial,
                                                               unsigned int* __restrict__ ticket, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v[CT_NRED];
#pragma unroll
  for (int k = 0; k < CT_NRED; k++) v[k] = 0.f;
  if (i < P.n) {
    const float id = pc_id[i], x = pc_u[i], y = pc_v[i];
    const float p0 = P.RKi[0] * x + P.RKi[1] * y + P.RKi[2] + P.t[0] * id;
    const float p1 = P.RKi[3] * x + P.RKi[4] * y + P.RKi[5] + P.t[1] * id;
    const float p2 = P.RKi[6] * x + P.RKi[7] * y + P.RKi[8] + P.t[2] * id;
    const float u = p0 / p2, vv = p1 / p2;
    const float Ku = P.fx * u + P.cx, Kv = P.fy * vv + P.cy;
    const float new_idepth = id / p2;
    if (P.lvl == 0 && (i & 31) == 0) {  // flow indicators (L416-447)
      const float k0 = P.Ki[0] * x + P.Ki[1] * y + P.Ki[2], k1 = P.Ki[3] * x + P.Ki[4] * y + P.Ki[5], k2 = P.Ki[6] * x + P.Ki[7] * y + P.Ki[8];
      const float T0 = k0 + P.t[0] * id, T1 = k1 + P.t[1] * id, T2 = k2 + P.t[2] * id;
      const float M0 = k0 - P.t[0] * id, M1 = k1 - P.t[1] * id, M2 = k2 - P.t[2] * id;
      const float q0 = P.RKi[0] * x + P.RKi[1] * y + P.RKi[2] - P.t[0] * id;
      const float q1 = P.RKi[3] * x + P.RKi[4] * y + P.RKi[5] - P.t[1] * id;
      const float q2 = P.RKi[6] * x + P.RKi[7] * y + P.RKi[8] - P.t[2] * id;
      const float KuT = P.fx * (T0 / T2) + P.cx, KvT = P.fy * (T1 / T2) + P.cy;
      const float KuT2 = P.fx * (M0 / M2) + P.cx, KvT2 = P.fy * (M1 / M2) + P.cy;
      const float Ku3 = P.fx * (q0 / q2) + P.cx, Kv3 = P.fy * (q1 / q2) + P.cy;
      v[49] = (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y) + (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      v[50] = (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y) + (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      v[51] = 2.f;
    }
    if (Ku > 2.f && Kv > 2.f && Ku < (float)(P.w - 3) && Kv < (float)(P.h - 3) && new_idepth > 0.f) {
      const int ix = (int)Ku, iy = (int)Kv;
      const float dx = Ku - ix, dy = Kv - iy, dxdy = dx * dy;
      const float4* bp = img + (size_t)iy * P.w + ix;
      const float4 tl = __ldg(bp), tr = __ldg(bp + 1), bl = __ldg(bp + P.w), br = __ldg(bp + P.w + 1);
      const float w11 = dxdy, w10 = dy - dxdy, w01 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
      const float h0 = w11 * br.x + w10 * bl.x + w01 * tr.x + w00 * tl.x;
      const float h1 = w11 * br.y + w10 * bl.y + w01 * tr.y + w00 * tl.y;
      const float h2 = w11 * br.z + w10 * bl.z + w01 * tr.z + w00 * tl.z;
      if (isfinite(h0)) {
        const float refColor = pc_col[i];
        const float residual = h0 - (P.affa * refColor + P.affb);
        const float ar = fabsf(residual);
        const float hw = ar < P.huber ? 1.f : P.huber / ar;
        if (ar > P.cutoff) {
          v[45] = P.maxEnergy; v[46] = 1.f; v[47] = 1.f;
        } else {
          v[45] = hw * residual * residual * (2.f - hw); v[46] = 1.f; v[48] = 1.f;
          if (P.want_gs) {  // calcGSSSE row (L316-336) and Accumulator9::updateSSE_eighted (MatrixAccumulators.h:L1091-1166)
            const float gx = h1 * P.fx, gy = h2 * P.fy;
            float J[9];
            J[0] = new_idepth * gx;
            J[1] = new_idepth * gy;
            J[2] = 0.f - new_idepth * (u * gx + vv * gy);
            J[3] = 0.f - ((u * vv) * gx + gy * (1.f + vv * vv));
            J[4] = (u * vv) * gy + gx * (1.f + u * u);
            J[5] = u * gy - vv * gx;
            J[6] = P.a_gs * (P.b0 - refColor);
            J[7] = -1.f;
            J[8] = residual;
            int e = 0;
#pragma unroll
            for (int r = 0; r < 9; r++) {
              const float Jw = J[r] * hw;
#pragma unroll
              for (int c = r; c < 9; c++) v[e++] = Jw * J[c];
            }
          }
        }
      }
    }
  }
  // ---- block reduction: warp shuffles, then 8 warp partials through shared memory
  __shared__ float s_red[CT_THREADS / 32][CT_NRED];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < CT_NRED; k++) {
    float a = v[k];
    a += __shfl_xor_sync(0xffffffffu, a, 16);
    a += __shfl_xor_sync(0xffffffffu, a, 8);
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    if (lane == 0) s_red[warp][k] = a;
  }
  __syncthreads();
  if (threadIdx.x < CT_NRED) {
    double s = 0.0;
#pragma unroll
    for (int wv = 0; wv < CT_THREADS / 32; wv++) s += (double)s_red[wv][threadIdx.x];
    partial[(size_t)blockIdx.x * CT_NRED + threadIdx.x] = s;
  }
  // ---- last block folds the per-block partials in block order
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int tk = atomicAdd(ticket, 1u);
    s_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (threadIdx.x < CT_NRED) {
      double s = 0.0;
      for (unsigned int bk = 0; bk < gridDim.x; bk++) s += __ldcg(&partial[(size_t)bk * CT_NRED + threadIdx.x]);
      out[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

__global__ void ct_repack_kernel(const float* __restrict__ src, float4* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
}
// level l+1 intensity = 2x2 box mean of level l (HessianBlocks.cpp:L159-166)
__global__ void pyr_down_kernel(const float4* __restrict__ src, float* __restrict__ dst, int wl, int hl, int wlm1) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= wl || y >= hl) return;
  const float4* p = src + (size_t)2 * y * wlm1 + 2 * x;
  dst[(size_t)y * wl + x] = 0.25f * (p[0].x + p[1].x + p[wlm1].x + p[wlm1 + 1].x);
}
__global__ void grad_kernel(const float* __restrict__ img, float4* __restrict__ dst, int w, int h) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w * h) return;
  float dx = 0.f, dy = 0.f;
  if (idx >= w && idx < w * (h - 1)) {
    dx = 0.5f * (img[idx + 1] - img[idx - 1]);
    dy = 0.5f * (img[idx + w] - img[idx - w]);
    if (!isfinite(dx)) dx = 0.f;
    if (!isfinite(dy)) dy = 0.f;
  }
  dst[idx] = make_float4(img[idx], dx, dy, 0.f);
}

}  // namespace dmv

using namespace dmv;

struct dmv_ct {
  dmv_ct_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool timing = false;  // CUDA-event timing of every evaluation (dmv_ct_set_timing); off by default: two event records per launch
  int w[DMV_MAX_PYR_LEVELS], h[DMV_MAX_PYR_LEVELS];
  float fx[DMV_MAX_PYR_LEVELS], fy[DMV_MAX_PYR_LEVELS], cx[DMV_MAX_PYR_LEVELS], cy[DMV_MAX_PYR_LEVELS];
  bool haveK[DMV_MAX_PYR_LEVELS];
  float4* d_img[DMV_MAX_PYR_LEVELS] = {nullptr};
  float* d_gray[DMV_MAX_PYR_LEVELS] = {nullptr};
  float* d_stage = nullptr;
  float *d_u[DMV_MAX_PYR_LEVELS] = {nullptr}, *d_v[DMV_MAX_PYR_LEVELS] = {nullptr}, *d_id[DMV_MAX_PYR_LEVELS] = {nullptr},
        *d_col[DMV_MAX_PYR_LEVELS] = {nullptr};
  int n[DMV_MAX_PYR_LEVELS];
  double* d_partial = nullptr;
  unsigned int* d_ticket = nullptr;
  double* d_out = nullptr;
  double* h_out = nullptr;
  float* h_scratch = nullptr;
  float huber = 9.f;
  long long launches = 0;
  float last_ms[4] = {0, 0, 0, 0};
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) return dmv::set_error(DMV_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); \
  } while (0)

extern "C" {

int dmv_ct_create(const dmv_ct_config* cfg, dmv_ct** out) {
  if (!cfg || !out) return set_error(DMV_ERR_INVALID, "null argument");
  if (cfg->levels < 1 || cfg->levels > DMV_MAX_PYR_LEVELS || cfg->w < 16 || cfg->h < 16 || cfg->max_points < 1)
    return set_error(DMV_ERR_INVALID, "bad config");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(DMV_ERR_NO_DEVICE, "no CUDA device: dmvio_b200 has no CPU path");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return set_error(DMV_ERR_INVALID, "device out of range");
  CK(cudaSetDevice(cfg->device));
  dmv_ct* c = new dmv_ct();
  c->cfg = *cfg;
  c->device = cfg->device;
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&c->ev[0]));
  CK(cudaEventCreate(&c->ev[1]));
  for (int l = 0; l < cfg->levels; l++) {
    c->w[l] = cfg->w >> l; c->h[l] = cfg->h >> l; c->haveK[l] = false; c->n[l] = 0;
    const size_t npx = (size_t)c->w[l] * c->h[l];
    CK(cudaMalloc(&c->d_img[l], npx * sizeof(float4)));
    CK(cudaMalloc(&c->d_gray[l], npx * sizeof(float)));
    CK(cudaMalloc(&c->d_u[l], sizeof(float) * cfg->max_points));
    CK(cudaMalloc(&c->d_v[l], sizeof(float) * cfg->max_points));
    CK(cudaMalloc(&c->d_id[l], sizeof(float) * cfg->max_points));
    CK(cudaMalloc(&c->d_col[l], sizeof(float) * cfg->max_points));
  }
  const size_t npx0 = (size_t)cfg->w * cfg->h;
  CK(cudaMalloc(&c->d_stage, npx0 * 3 * sizeof(float)));
  const int maxBlocks = (cfg->max_points + CT_THREADS - 1) / CT_THREADS;
  CK(cudaMalloc(&c->d_partial, sizeof(double) * CT_NRED * maxBlocks));
  CK(cudaMalloc(&c->d_ticket, sizeof(unsigned int)));
  CK(cudaMemset(c->d_ticket, 0, sizeof(unsigned int)));
  CK(cudaMalloc(&c->d_out, sizeof(double) * 64));
  CK(cudaMallocHost(&c->h_out, sizeof(double) * 64));
  CK(cudaMallocHost(&c->h_scratch, sizeof(float) * std::max(npx0 * 3, (size_t)cfg->max_points * 4)));
  *out = c;
  return DMV_OK;
}

int dmv_ct_destroy(dmv_ct* c) {
  if (!c) return DMV_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int l = 0; l < DMV_MAX_PYR_LEVELS; l++) {
    cudaFree(c->d_img[l]); cudaFree(c->d_gray[l]); cudaFree(c->d_u[l]); cudaFree(c->d_v[l]); cudaFree(c->d_id[l]); cudaFree(c->d_col[l]);
  }
  cudaFree(c->d_stage); cudaFree(c->d_partial); cudaFree(c->d_ticket); cudaFree(c->d_out);
  cudaFreeHost(c->h_out); cudaFreeHost(c->h_scratch);
  cudaEventDestroy(c->ev[0]); cudaEventDestroy(c->ev[1]);
  cudaStreamDestroy(c->stream);
  delete c;
  return DMV_OK;
}

int dmv_ct_set_K(dmv_ct* c, int l, float fx, float fy, float cx, float cy) {
  if (!c || l < 0 || l >= c->cfg.levels) return set_error(DMV_ERR_INVALID, "bad level");
  c->fx[l] = fx; c->fy[l] = fy; c->cx[l] = cx; c->cy[l] = cy; c->haveK[l] = true;
  return DMV_OK;
}
int dmv_ct_set_huber(dmv_ct* c, float huberTH) {
  if (!c) return set_error(DMV_ERR_INVALID, "null handle");
  c->huber = huberTH;
  return DMV_OK;
}

int dmv_ct_set_ref(dmv_ct* c, int l, int n, const float* u, const float* v, const float* id, const float* col) {
  if (!c || l < 0 || l >= c->cfg.levels || n < 0 || n > c->cfg.max_points) return set_error(DMV_ERR_INVALID, "bad level / count");
  if (n > 0 && (!u || !v || !id || !col)) return set_error(DMV_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  c->n[l] = n;
  if (n > 0) {
    float* s = c->h_scratch;
    std::memcpy(s, u, 4 * (size_t)n); std::memcpy(s + n, v, 4 * (size_t)n); std::memcpy(s + 2 * (size_t)n, id, 4 * (size_t)n);
    std::memcpy(s + 3 * (size_t)n, col, 4 * (size_t)n);
    CK(cudaMemcpyAsync(c->d_u[l], s, 4 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_v[l], s + n, 4 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_id[l], s + 2 * (size_t)n, 4 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_col[l], s + 3 * (size_t)n, 4 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return DMV_OK;
}

int dmv_ct_upload_new(dmv_ct* c, int l, const float* dIp) {
  if (!c || !dIp || l < 0 || l >= c->cfg.levels) return set_error(DMV_ERR_INVALID, "bad level / pointer");
  CK(cudaSetDevice(c->device));
  const size_t npx = (size_t)c->w[l] * c->h[l];
  std::memcpy(c->h_scratch, dIp, npx * 3 * sizeof(float));
  CK(cudaMemcpyAsync(c->d_stage, c->h_scratch, npx * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  ct_repack_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, c->stream>>>(c->d_stage, c->d_img[l], (int)npx);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  return DMV_OK;
}

int dmv_ct_upload_new_image(dmv_ct* c, const float* image) {
  if (!c || !image) return set_error(DMV_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  const size_t npx = (size_t)c->w[0] * c->h[0];
  std::memcpy(c->h_scratch, image, npx * sizeof(float));
  CK(cudaMemcpyAsync(c->d_gray[0], c->h_scratch, npx * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  for (int l = 0; l < c->cfg.levels; l++) {
    const int wl = c->w[l], hl = c->h[l];
    if (l > 0) {
      dim3 grid((wl + 127) / 128, hl);
      pyr_down_kernel<<<grid, 128, 0, c->stream>>>(c->d_img[l - 1], c->d_gray[l], wl, hl, c->w[l - 1]);
      c->launches++;
    }
    grad_kernel<<<(wl * hl + 255) / 256, 256, 0, c->stream>>>(c->d_gray[l], c->d_img[l], wl, hl);
    c->launches++;
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  return DMV_OK;
}

int dmv_ct_calc_res_gs(dmv_ct* c, int l, const float RKi[9], const float t[3], const float affLL[2], float b0, float cutoffTH, int want_gs,
                       double res6[6], double H[64], double b[8], int* n_warped) {
  if (!c || !RKi || !t || !affLL || !res6 || l < 0 || l >= c->cfg.levels) return set_error(DMV_ERR_INVALID, "bad argument");
  if (!c->haveK[l]) return set_error(DMV_ERR_STATE, "dmv_ct_set_K(level) first");
  if (want_gs && (!H || !b)) return set_error(DMV_ERR_INVALID, "H/b required with want_gs");
  CK(cudaSetDevice(c->device));
  CTParams P;
  std::memcpy(P.RKi, RKi, sizeof(P.RKi));
  std::memcpy(P.t, t, sizeof(P.t));
  // Ki[lvl] = K^-1 (CoarseTracker.cpp:L126-133)
  {
    const float Kl[9] = {c->fx[l], 0.f, c->cx[l], 0.f, c->fy[l], c->cy[l], 0.f, 0.f, 1.f};
    inv3_cofactor(Kl, P.Ki);
  }
  P.fx = c->fx[l]; P.fy = c->fy[l]; P.cx = c->cx[l]; P.cy = c->cy[l];
  P.affa = affLL[0]; P.affb = affLL[1]; P.a_gs = affLL[0]; P.b0 = b0;
  P.cutoff = cutoffTH; P.huber = c->huber;
  P.maxEnergy = 2 * c->huber * cutoffTH - c->huber * c->huber;
  P.w = c->w[l]; P.h = c->h[l]; P.n = c->n[l]; P.lvl = l; P.want_gs = want_gs;
  const int nb = std::max(1, (c->n[l] + CT_THREADS - 1) / CT_THREADS);
  if (c->timing) CK(cudaEventRecord(c->ev[0], c->stream));
  // the last block writes the 53 reduced doubles straight into the pinned host buffer (zero-copy): no D2H copy node per evaluation
  ct_res_gs_kernel<<<nb, CT_THREADS, 0, c->stream>>>(P, c->d_u[l], c->d_v[l], c->d_id[l], c->d_col[l], c->d_img[l], c->d_partial, c->d_ticket,
                                                      c->h_out);
  c->launches++;
  if (c->timing) CK(cudaEventRecord(c->ev[1], c->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  if (c->timing) cudaEventElapsedTime(&c->last_ms[0], c->ev[0], c->ev[1]);
  const double* o = c->h_out;
  const double E = o[45], nE = o[46], nSat = o[47], nW = o[48];
  res6[0] = E;
  res6[1] = nE;
  res6[2] = (double)((float)o[49] / ((float)o[51] + 0.1f));
  res6[3] = 0;
  res6[4] = (double)((float)o[50] / ((float)o[51] + 0.1f));
  res6[5] = (double)((float)nSat / (float)nE);
  const int npad = ((int)nW + 3) & ~3;
  if (n_warped) *n_warped = npad;
  if (want_gs) {
    // acc.H (9x9) -> H_out = H[0:8,0:8] * (1/n), b_out = H[0:8,8] * (1/n), then SCALE_* (CoarseTracker.cpp:L341-355)
    double M[9][9];
    int e = 0;
    for (int r = 0; r < 9; r++)
      for (int cc = r; cc < 9; cc++) { M[r][cc] = M[cc][r] = o[e]; e++; }
    const double inv = (double)(1.0f / (float)npad);
    const double sc[8] = {1, 1, 1, 1, 1, 1, 10.0, 1000.0};  // SCALE_XI_ROT/TRANS = 1, SCALE_A, SCALE_B
    for (int r = 0; r < 8; r++) {
      for (int cc = 0; cc < 8; cc++) H[r * 8 + cc] = M[r][cc] * inv * sc[r] * sc[cc];
      b[r] = M[r][8] * inv * sc[r];
    }
  }
  return DMV_OK;
}

int dmv_ct_set_timing(dmv_ct* c, int enable) {
  if (!c) return set_error(DMV_ERR_INVALID, "null handle");
  c->timing = enable != 0;
  return DMV_OK;
}
int dmv_ct_last_timing(dmv_ct* c, float ms[4]) {
  if (!c || !ms) return set_error(DMV_ERR_INVALID, "null argument");
  for (int i = 0; i < 4; i++) ms[i] = c->last_ms[i];
  return DMV_OK;
}
int dmv_ct_kernel_launch_count(dmv_ct* c, long long* n) {
  if (!c || !n) return set_error(DMV_ERR_INVALID, "null argument");
  *n = c->launches;
  return DMV_OK;
}

}  // extern "C"
