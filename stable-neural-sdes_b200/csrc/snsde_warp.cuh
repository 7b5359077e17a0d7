// Warp-resident kernel for hidden sizes <= 32 (snsde_warp.cu): host-visible program form.
#pragma once
#include <cuda_runtime.h>

#include "snsde_common.cuh"

namespace snsde {

constexpr int kWarpMaxMv = 6;             // register budget: 32 weights per mat-vec per lane
constexpr int kWarpDstDrift = kNumRowBufs; // pseudo-destination of the final drift op

// One register-resident mat-vec of at most 32 x 32:  acc (+)= sum_{k<K} act_src[k] * Wt[k][lane].
struct WarpMv {
  int src;          // activation buffer id read (BUF_Y .. BUF_Q)
  int K, N;         // active inputs / outputs
  int w_off;        // offset of the transposed [K][N] image in the FMA weight image
  int first;        // starts an output: acc = bias + time term
  int last;         // completes it: activation, write to dst
  int dst;          // buffer id, or kWarpDstDrift
  int act, tmode;
  int b_off, tw_off;
};

struct WarpProg {
  int n_mv;
  WarpMv mv[kWarpMaxMv];
};

// Flattens the per-row ops of `pg` into mat-vecs; false when the model / method is outside the kernel's envelope
// (hidden or control width above 32, more than kWarpMaxMv mat-vecs, SRK, Milstein through a noise network, LatentSDE).
bool warp_plan(const Program& pg, int method, WarpProg& wp);
cudaError_t warp_launch(const FmaParams& p, const WarpProg& wp, int num_sms, cudaStream_t stream);

}  // namespace snsde
