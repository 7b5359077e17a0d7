// Warp-owned kernel for hidden sizes <= 32 (snsde_warp.cu): host-visible program form.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "snsde_common.cuh"

namespace snsde {

constexpr int kWarpMaxMv = 16;             // mat-vecs per step (the descriptors travel as a kernel parameter: every field is a
                                           // constant-bank operand of the unrolled mat-vec slot that uses it)
constexpr int kWarpMaxRows = 4096;         // rows per launch up to which the warp-owned kernel beats the interpreter (snsde_warp.cu)
constexpr int kWarpDstDrift = kNumRowBufs; // pseudo-destination of the final drift op
enum : int { kMvFirst = 1, kMvLast = 2, kMvSinCos = 4, kMvDiff = 8 };

// One mat-vec of at most 32 x 32:  acc (+)= sum_k act_src[k] * W[lane][k], K padded to 8 * n8 with zero weights.
struct WarpMv {
  int src;          // activation row read (BUF_Y .. BUF_Q)
  int n8;           // chunks of 8 inputs
  int N;            // outputs
  int stride;       // floats between weight rows: 8 * n8 + 4 (conflict-free 16-byte loads, lane = row)
  int w_off;        // offset of the [N][stride] rows in the warp image
  int flags;        // kMvFirst: starts an output (acc = bias + time term); kMvLast: completes it (activation, write to
                    // dst); kMvSinCos: the output has time-feature weights; kMvDiff: part of the diffusion (the SRK stages evaluate drift and
                    // diffusion at different states)
  int dst;          // activation row written, or kWarpDstDrift
  int act;
  int b_off, tw_off; // 32-float bias row; sin row followed by the cos row (time features)
};

struct WarpProg {
  int n_mv;
  int n_emits;      // entries of the emit table (filled in per launch)
  int coef_off;     // 32-float row of the per-feature diffusion coefficient (CO_IMG)
  int pad;
  WarpMv mv[kWarpMaxMv];
};

// Flattens the per-row ops of `pg` into mat-vecs and builds their weight image from the nn.Linear blob; false when the
// model / method is outside the kernel's envelope (hidden or control width above 32, more than kWarpMaxMv mat-vecs,
// Milstein through a noise network).  `fma_img`: the interpreter's host image (per-feature coefficient).
bool warp_build(const Program& pg, int method, const float* blob, const float* fma_img, WarpProg& wp, std::vector<float>& img);
size_t warp_smem_bytes(int img_floats, int pairs, int R, int S, int n_emits, bool tables, bool srk);
cudaError_t warp_launch(const FmaParams& p, const WarpProg& wp, int method, int n_emits, int num_sms, int smem_optin, cudaStream_t stream);

}  // namespace snsde
