// Reverse sweep through the Euler-Maruyama solve: host-visible interface.
//
// The reference trains through torchsde.sdeint with plain autograd
// (/root/reference/benchmark_classification/common_sde.py:156-162: `pred_y = model(...); loss.backward()`), i.e.
// it back-propagates through every solver step.  Here the forward solve stores the solver states y_s
// ([S+1, B, H]); the reverse sweep recomputes the activations of step s from y_s, pulls the state cotangent
// lambda_{s+1} back to lambda_s, and writes, per dense op, the pre-activation cotangents D_op [S*B, N] and the op's
// input activations P [S*B, K]; the parameter gradients are then plain library GEMMs  dW = D^T P  (cuBLAS).
#pragma once
#include <cuda_runtime.h>
#include "snsde_common.cuh"

namespace snsde {

struct BwdParams {
  Program prog;
  const float* wimg; int wimg_floats; int smem_w_floats;   // forward image ([in][out], as the FMA kernels read it)
  const float* blob;                                        // raw nn.Linear blob on device ([out][in]): transposed products
  const float* coeffs; long long coeff_row_stride;
  int B, S;
  const snsde_step* steps;
  const float* states;        // [S+1][B][H] solver states saved by the forward solve
  const float* grad_states;   // [S+1][B][H] dL/dy_s from the output selection (zeros where a state is not emitted)
  const float* dW;            // explicit increments [S][B][H] or null (Philox replay)
  unsigned long long seed, row_offset;
  const float* vtab;          // [S][H] row-independent diffusion coefficient of the forward pass (CO_VBUF) or null
  float* grad_y0;             // [B][H]
  float* grad_blob;           // zero-filled by the host; theta / sigma / sigma_diag land here by atomics
  float* gvtab;               // [S][H] cotangent of vtab (atomics), or null
  float* xbuf;                // [S*B][C] X(t_s) per row, or null
  float* dbuf[kMaxOps];       // per program op: D [S*B][N]   (null for vec ops)
  float* pbuf[kMaxOps];       // per program op: activated output [S*B][N] (null when no later op reads it)
  int n_rops;                 // number of per-row ops
  int rop[kMaxOps];           // their program indices, in program order
  int groups, nw;
  int has_lipswish;           // keep pre-activations too (tutorial family)
  // ---- method 'srk' (snsde_bwd_srk.cu): D / P / xbuf carry one block of S*B rows per evaluation site of the op's part
  // (drift sites f0, f1, f2; diffusion sites g0, g1, g2, g3); vtab / gvtab are [S][kSrkGPoints][H]
  const snsde_point* points;  // [S][kSrkPoints]
  const float* dU;            // explicit space-time Levy integrals [S][B][H] (with dW) or null
  float* pstate_f;            // [3*S*B][H] states the drift sites were evaluated at (GEMM operand of ops that read the state)
  float* pstate_g;            // [4*S*B][H] states of the diffusion sites, or null (no per-row diffusion network)
};

size_t bwd_group_smem_floats(const Program& pg, int n_rops, int R, int has_lipswish);
cudaError_t bwd_launch(const BwdParams& p, int R, size_t smem, cudaStream_t stream);
// aux[i] = (sin t_s, cos t_s, 1) for i = s*B + b: the "activations" behind the time-feature columns and the biases
cudaError_t bwd_fill_aux(const snsde_step* steps, int S, int B, float* aux, cudaStream_t stream);
// backward of the row-independent coefficient networks (vec ops): one CTA per step, atomics into grad_blob
// (points != null: method 'srk', npg = kSrkGPoints coefficient rows per step at t0, t0+h/4, t0+h)
cudaError_t vec_bwd_launch(const Program& pg, const float* wimg, const float* blob, const snsde_step* steps,
                           const snsde_point* points, int S, int npg, const float* gvtab, float* grad_blob, cudaStream_t stream);
// method 'srk'
size_t bwd_srk_group_smem_floats(const Program& pg, int n_rops, int R, int has_lipswish);
cudaError_t bwd_srk_launch(const BwdParams& p, int R, size_t smem, cudaStream_t stream);
cudaError_t bwd_srk_fill_aux(const snsde_point* points, int S, int B, float* aux_f, float* aux_g, cudaStream_t stream);

}  // namespace snsde
