// Shared host/device definitions of the SDE engine (internal; the public surface is include/snsde.h).
#pragma once
#include <stdint.h>
#include "../../include/snsde.h"

namespace snsde {

constexpr int kMaxOps = 24;
constexpr int kMaxRowsPerCta = 16;

// Activation buffers of the FMA interpreter kernel.  Row buffers are [R][ld]; vector
// buffers (row-independent work such as noise_t(time_features), neuralsde.py:272-282) are [ld].
enum Buf : int { BUF_NONE = -1, BUF_Y = 0, BUF_X, BUF_U, BUF_A, BUF_B, BUF_P, BUF_Q, BUF_V0, BUF_V1, BUF_COUNT };
constexpr int kNumRowBufs = BUF_V0;          // Y..Q
constexpr int kNumVecBufs = BUF_COUNT - BUF_V0;

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LIPSWISH = 2 };
enum TimeMode : int { TM_NONE = 0, TM_SINCOS = 1, TM_RAW = 2 };

// dst[r][j] = act( b[j] + sum_k src[r][k] Wt[k][j] + sum_k src2[r][k] Wt2[k][j] + time_term[j] )
// Wt is the TRANSPOSE of the nn.Linear weight ([in][out]) so that threads j read it coalesced.
struct DenseOp {
  int dst, src, src2;
  int K, K2, N;
  int w_off, w2_off, b_off, tw_off;   // float offsets into the weight image (-1: absent)
  int tmode;                          // TimeMode: tw is [2][N] (sin,cos rows) or [1][N] (raw t)
  int act;
  int vec;                            // 1: row-independent (one row per evaluation time: tabulated by the
                                      //    pre-kernel, snsde_fma.cu vec_tables_kernel), 0: per batch row
  int final_drift;                    // 1: result stays in registers and feeds the SDE update
  int part;                           // 0: drift f, 1: diffusion g (the SRK stages evaluate them at different states)
  // ---- backward pass (snsde_bwd.cu): where the op's parameters sit in the flat weight blob, and who feeds it
  int g_w, g_ldw, g_col, g_col2;      // nn.Linear weight [N][g_ldw] at blob offset g_w; src multiplies columns
                                      // [g_col, g_col+K), src2 columns [g_col2, g_col2+K2); time features columns [0,2)
  int g_b;                            // bias offset in the blob (-1: none)
  int src_op, src2_op;                // op that produced src / src2 (SRC_STATE, SRC_CONTROL for the inputs)
};
enum : int { SRC_STATE = -1, SRC_CONTROL = -2, SRC_ABSENT = -3 };

// Elementwise diffusion g (neuralsde.py:233-307) and the state update.
enum CoefSrc : int { CO_NONE = 0, CO_SCALAR, CO_IMG, CO_VBUF, CO_RBUF };
enum Mult : int { MU_ONE = 0, MU_T, MU_Y, MU_TY };
enum Special : int { SP_NONE = 0, SP_ZERO, SP_SQRT, SP_CUBE, SP_SIGMOID, SP_RELU };

struct TailOp {
  int geometric;      // drift *= tanh(y)           (input options 5,6; neuralsde.py:219-225)
  int clip_drift;     // drift = tanh(drift)        (neuralsde.py:227-231)
  int coef_src;       // CoefSrc
  int coef_ref;       // image offset / buffer id
  float coef_scalar;  // exp(sigma) for options 1-3
  int mult;           // Mult
  int special;        // Special (options 0,7,8,9,10)
  int bounded;        // g = tanh(sigmoid(theta) * nan_to_num(raw)) (benchmark) vs g = raw (tutorial)
  float s_theta;      // sigmoid(theta)
  int milstein;       // method
  // Milstein with a state-dependent noise NETWORK (options 14,15,18,19): g_i depends on every y_j, so torchsde's
  // vjp_y(g; g*(dW^2-h)) is a full transposed product through noise_y.  0: none (diagonal closed form),
  // 1: q = Linear(H+2,H)([tf,y]);  2: q = relu(Linear(H,H)(relu(Linear(H+2,H)([tf,y])))).   (FMA kernel only)
  int vjp_kind;
  int vjp_w1, vjp_w2; // float offsets of the transposed images [in][out] of noise_y.0 (y columns) / noise_y.2
  int vjp_h1;         // buffer holding relu(noise_y.0(...)) (kind 2)
  // backward pass: blob offsets of theta / sigma / sigma_diag (-1: absent) and the op producing a CO_RBUF coefficient
  int g_theta, g_sigma, coef_op;
  // LatentSDE augmented system (latent_sde.py:77-90): features 0..H-2 carry the posterior drift f and the constant
  // diffusion sigma; feature H-1 has drift 0.5 * sum_j u_j^2 with u = (f - theta (mu - y)) / stable(sigma) and no noise.
  int latent;
  float lat_theta, lat_mu, lat_div;   // prior drift theta (mu - y); lat_div = the _stable_division denominator (:24-26)
};

struct Program {
  int n_ops;
  int uses_control;   // spline X(t) is read (input options 0,2,4,6 / tutorial)
  int C, H, HH;
  int ld;             // leading dimension (floats) of every activation buffer
  DenseOp ops[kMaxOps];
  TailOp tail;
};

// Evaluation points per step: Euler/Milstein evaluate f and g at t0 only; SRK (SRID2) evaluates f at
// t0 + {0, 1, 1/2} h and g at t0 + {0, 1/4, 1} h.  The point table of an SRK step is [t0, t0+h/4, t0+h/2, t0+h].
constexpr int kSrkPoints = 4;
constexpr int kSrkGPoints = 3;            // rows of the row-independent coefficient table per SRK step: t0, t0+h/4, t0+h

struct FmaParams {
  Program prog;
  const float* wimg;          // weight image (global)
  int wimg_floats;
  int smem_w_floats;          // leading floats of the image staged into shared memory
  const float* coeffs; long long coeff_row_stride;
  const float* y0; int B;
  const snsde_step* steps; int S;
  const snsde_point* points;  // SRK: [S][kSrkPoints]; null otherwise
  const snsde_emit* emits; int n_init_emits; int n_out;
  const int* row_slot;
  const float* dW; const float* dU;
  unsigned long long seed, row_offset;
  const float* vtab;          // [S][1 or kSrkGPoints][H] row-independent diffusion coefficient (CO_VBUF), or null
  float* out;
  int groups;                 // row groups per CTA
  int nw;                     // warps per row group (32*nw >= max(H, HH))
};

}  // namespace snsde
