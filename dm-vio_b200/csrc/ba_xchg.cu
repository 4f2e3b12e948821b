// ba_xchg_kernel — the exchange step of the sharded BA over NVLink/NVSwitch peer memory.
//
// Every rank holds the stitched system of ITS points (46.7 KB of doubles at nf = 7).  One launch per GN iteration, chained
// behind the stitch kernel by programmatic dependent launch, all-reduces it without NCCL:
//   push   each thread loads one 16-byte vector of the local result and stores it into slot [parity][my rank] of EVERY
//          rank's inbox (peer-mapped device memory, CUDA IPC; the stores travel over NVLink, nobody reads remotely)
//   flag   __threadfence_system, then one st.release.sys per destination of the iteration number into [parity][my rank][cta]
//   wait   ld.acquire.sys spin until the nranks flags of this CTA's slice carry the iteration number
//   sum    the nranks slices are added in RANK ORDER (so every rank computes bit-identical sums) and written in place
// CTA c of rank A only depends on CTA c of the other ranks; there is no inter-CTA dependency inside a rank.  The inbox is
// double-buffered by iteration parity: a peer can run at most one exchange ahead (it needs my flag of iteration k+1, which my
// stream issues only after my iteration-k kernel has retired), so slot k%2 is never overwritten while it is still being read.
#include "ba_device.cuh"
#include "common_host.h"

namespace dmv {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(XCHG_THREADS) ba_xchg_kernel(const __grid_constant__ XchgDev X) {
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the stitch kernel has written X.buf
  const int tid = threadIdx.x, c = blockIdx.x;
  const int par = (int)(X.seq & 1ull);
  const int per = (X.nvec + XCHG_CTAS - 1) / XCHG_CTAS;
  const int i = c * per + tid;  // XCHG_THREADS >= per (checked on the host): at most one vector per thread
  const bool mine = tid < per && i < X.nvec;
  const size_t slot = (size_t)(par * XCHG_MAXR + X.rank) * X.pitch;
  if (mine) {
    const double2 v = X.buf[i];
#pragma unroll 1
    for (int k = 0; k < X.nranks; k++) {
      const int r = (X.rank + 1 + k) % X.nranks;  // start with the neighbour: spreads the NVSwitch ports
      X.inbox[r][slot + i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tid < X.nranks) {
    st_release_sys(X.flags[tid] + (size_t)(par * XCHG_MAXR + X.rank) * XCHG_CTAS + c, X.seq);
    const unsigned long long* f = X.flags[X.rank] + (size_t)(par * XCHG_MAXR + tid) * XCHG_CTAS + c;
    while (ld_acquire_sys(f) < X.seq) { /* spin: the peer's slice is on its way */ }
  }
  __syncthreads();
  if (mine) {
    const double2* in = X.inbox[X.rank] + (size_t)par * XCHG_MAXR * X.pitch + i;
    double2 s = make_double2(0.0, 0.0);
    for (int r = 0; r < X.nranks; r++) {
      const double2 v = __ldcg(in + (size_t)r * X.pitch);  // L2: the line was written by a remote GPU
      s.x += v.x; s.y += v.y;
    }
    X.buf[i] = s;
  }
}

void launch_xchg_kernel(const XchgDev& X, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(XCHG_CTAS);
  cfg.blockDim = dim3(XCHG_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, ba_xchg_kernel, X);
}

}  // namespace dmv
