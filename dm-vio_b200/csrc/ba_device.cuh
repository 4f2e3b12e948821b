// Device-side data layout of one sliding window.  Everything the BA kernels touch is reachable from
// one BAWinDev descriptor that lives in device memory, so that (a) a CUDA graph can be replayed without patching
// kernel arguments and (b) several independent windows can be processed by one launch (blockIdx.y = window).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmv {

constexpr int MAXF = 8;             // DMV_MAX_FRAMES
constexpr int TOP_ROWS = 10;        // geometric rows [C4 | xi6] of the 13x13 pair block
constexpr int TOP_COLS = 13;
constexpr int TOP_PART = TOP_ROWS * TOP_COLS + 6;  // 136 floats: 10 full rows + 6 bottom-right (a,b,r) entries
constexpr int REC = 16;             // per (point,target) record kept in shared memory
constexpr int RES_NONE = 255, RES_IN = 0, RES_OOB = 1, RES_OUTLIER = 2;

// per-iteration parameter block (host -> device every GN iteration; ~11 KB at nf = 8)
struct BAIter {
  float calib[8];                 // fxl fyl cxl cyl fxli fyli cxli cyli
  float TH[MAXF];                 // frameEnergyTH
  float precalc[MAXF * MAXF][32]; // [h*nf + t] : KRKi[9] Kt[3] R0[9] t0[3] aff[2] b0
  float xAd[MAXF * MAXF][8];      // [h*nf + t] : x_h^T adHostF + x_t^T adTargetF   (resubstitute)
  float xc[4];                    // x.head<4>() as float
  int have_x;                     // 0: skip the resubstitute/step prologue
  int pad[3];
};

// adjoints (host -> device once per linearisation point)
struct BAAdj {
  float adHostF[MAXF * MAXF][64];   // [h*nf + t] row-major 8x8
  float adTdiagF[MAXF * MAXF][8];   // [h*nf + t] diagonal of adTarget
  double adHost[MAXF * MAXF][64];   // [h*nf + t]
  double adTdiag[MAXF * MAXF][8];
};

struct BAChunk { int start, count, host, pad; };

struct BAWinDev {
  int nf, npts, nchunks, w, h, N, NW, T, ntiles, mp;  // mp = capacity (row pitch of the [target][point] slot arrays)
  float huberTH, outlierTHSum;
  int zeroA, zeroB;
  const float4* img[MAXF];   // per window frame index: level-0 texels (I, dx, dy, 0)
  const BAIter* it;
  const BAAdj* adj;
  const BAChunk* chunks;
  int chunk_beg[MAXF + 1];   // chunks of host h are [chunk_beg[h], chunk_beg[h+1])
  // points
  const float2* uv;
  float* idepth;
  float* idepth_zero;
  float* idepth_backup;
  const float* color;        // [p][8]
  const float* weights;      // [p][8]
  const float* priorF;
  // residual slots, [t*mp + p]
  const uint8_t* st_in;      // state_state
  const float* en_in;        // state_energy
  uint8_t* st_new;           // tentative outputs of this linearisation
  float* en_new;
  float* en_wo;
  float* cpt;                // [slot][3]  (stored as 3 planes: [k][slot])
  float* jpjd;               // [slot][8]
  float* pout;               // [p][8]: Hdd bd Hcd[4] HdiF bdSum
  // committed copies (read by resubstitute)
  const uint8_t* c_st;
  const float* c_jpjd;
  const float* c_pout;
  float* step;               // [p]
  // partials / sums / result
  float* top_part;           // [chunk][t][TOP_PART]
  float* sc_part;            // [chunk][tile][16]
  float* misc_part;          // [chunk][t][4] : energy, n_in, n_oob, n_outlier
  double* step_part;         // [chunk][2] : sum step^2, sum |idepth_backup|
  double* top_sum;           // [h*nf + t][TOP_PART]
  double* sc_sum;            // [tile][16]
  double* result;            // H_top N*N | b_top N | H_sc N*N | b_sc N | energy, n_in, n_oob, n_outl, sum step^2, sum |id_backup|, npts, pad
};

inline __host__ __device__ int result_doubles(int N) { return 2 * (N * N + N) + 8; }

}  // namespace dmv
