/* snsde.h - C ABI of the B200-native Neural-SDE integration engine.
 *
 * Drop-in boundary for ONE reference path: the fixed-step SDE solve that
 *   /root/reference/benchmark_classification/models_sde/neuralsde.py:71-82
 * (= benchmark_forecasting/models_sde/neuralsde.py:71-82,145-156 and
 *    torch-ists/torch_ists/diff_module/NSDE/nsde_model.py:63-74)
 * delegates to `torchsde.sdeint(sde=func, y0=z0, ts=ts, dt=dt, method=...)`, with the
 * vector field `Diffusion_model.f/g` (neuralsde.py:123-307) evaluated inside the step.
 *
 * The reference is pure Python and has no FFI; the binding a maintainer adds is the
 * ctypes stub shown in INTEGRATION.md (it replaces the body of `_solve_sde_path`).
 *
 * Conventions
 *  - plain C, no torch types; all device pointers are raw CUDA pointers owned by the caller;
 *  - every entry point returns 0 on success or a negative snsde_status; nothing is written
 *    to `out` on a negative return that is detected before launch; no exception crosses;
 *  - `snsde_last_error()` returns a thread-local, NUL-terminated description of the last failure;
 *  - every entry point restores the caller's current CUDA device before returning;
 *  - entry points only ENQUEUE work on `stream` (a cudaStream_t passed as void*); they never
 *    synchronise the device, except `snsde_plan_set_weights` with `on_device=1`
 *    (one blocking D2H copy of <= a few MB) ;
 *  - a plan is not thread-safe; distinct plans are independent.
 */
#ifndef SNSDE_H_
#define SNSDE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNSDE_ABI_VERSION 4

typedef enum {
  SNSDE_OK = 0,
  SNSDE_ERR_BAD_ARG = -1,        /* NULL pointer, non-positive size, inconsistent plan/step tables  */
  SNSDE_ERR_UNSUPPORTED = -2,    /* option pair / method / precision the engine does not implement  */
  SNSDE_ERR_CUDA = -3,           /* CUDA runtime error (allocation, launch); text in last_error      */
  SNSDE_ERR_NO_WEIGHTS = -4,     /* forward called before snsde_plan_set_weights                     */
  SNSDE_ERR_INTERNAL = -5        /* host allocation failure / unexpected C++ exception (never thrown) */
} snsde_status;

enum { SNSDE_FAMILY_BENCHMARK = 0,   /* Diffusion_model, neuralsde.py:123-307                        */
       SNSDE_FAMILY_TUTORIAL_LSDE = 1, /* NeuralLSDEFunc, tutorial "Neural LSDE" notebook cell 7     */
       SNSDE_FAMILY_LATENT_SDE = 2   /* LatentSDE augmented system f_aug/g_aug (posterior drift + KL path
                                        accumulator), torch-ists/torch_ists/diff_module/NSDE/latent_sde.py:29-93,
                                        solved by sdeint_adjoint(..., names={'drift':'f_aug','diffusion':'g_aug'})
                                        at :134-141.  hidden = hidden_channels = latent width + 1 (the last state
                                        channel accumulates 0.5*|u|^2, u = (f - theta(mu - y)) / sigma)            */ };
enum { SNSDE_METHOD_EULER = 0,       /* torchsde Euler.step (Ito)                                     */
       SNSDE_METHOD_MILSTEIN = 1,    /* torchsde Milstein.step (Ito, diagonal, derivative-based)      */
       SNSDE_METHOD_SRK = 2          /* torchsde SRK.diagonal_or_scalar_step (SRID2 tableau, strong
                                        order 1.5): the torch-ists wrapper's default, nsde_model.py:67  */ };
enum { SNSDE_PRECISION_FP32 = 0,     /* fp32 FMA kernel, every model/shape                            */
       SNSDE_PRECISION_TC = 1,       /* tcgen05 tensor-core kernel, split-fp16 operands (fp32-class)   */
       SNSDE_PRECISION_AUTO = 2      /* TC when the model/shape is supported, else FP32               */ };

/* Model descriptor.  Replaces the constructor arguments of Diffusion_model
 * (neuralsde.py:124) plus the `method=` kwarg of sdeint (neuralsde.py:35-36). */
typedef struct {
  int32_t family;           /* SNSDE_FAMILY_*                                        */
  int32_t input_option;     /* 0..6   (neuralsde.py:148-156,200-225); 0 for tutorial / latent */
  int32_t noise_option;     /* 0..19  (neuralsde.py:233-288);        0 for tutorial / latent  */
  int32_t input_channels;   /* C                                                     */
  int32_t hidden;           /* H  = hidden_channels                                  */
  int32_t hidden_hidden;    /* HH = hidden_hidden_channels                           */
  int32_t num_hidden_layers;/* L                                                     */
  int32_t method;           /* SNSDE_METHOD_*                                        */
  int32_t precision;        /* SNSDE_PRECISION_*                                     */
} snsde_model_desc;

/* One solver step, built on the host by replaying torchsde's fixed-step loop in fp32
 * (BaseSDESolver.integrate: next_t = min(curr_t + dt, ts[-1])).  40 bytes. */
typedef struct {
  float t0;            /* step start (the time f and g are evaluated at)              */
  float h;             /* fp32(t1 - t0)                                               */
  float sqrt_h;        /* sqrtf(h): in-kernel Brownian increments are N(0,1)*sqrt_h   */
  float sin_t0;        /* time features of neuralsde.py:191-193                       */
  float cos_t0;
  int32_t interval;    /* clamp(bucketize(t0, knots) - 1, 0, K-2)  (CubicSpline)      */
  float frac;          /* fp32(t0 - knots[interval])                                  */
  int32_t emit_begin;  /* outputs produced after this step: emits[emit_begin:emit_end]*/
  int32_t emit_end;
  int32_t reserved;
} snsde_step;

/* One output row: out[slot] = w_prev * y_before_step + w_curr * y_after_step
 * (torchsde linear_interp; w_prev = 0, w_curr = 1 when the step lands on ts[slot]). */
typedef struct {
  int32_t slot;
  float w_prev;
  float w_curr;
} snsde_emit;

/* One evaluation time inside a step (method SRK only): the stages of the SRID2 tableau evaluate f at
 * t0 + {0, 1, 1/2} h and g at t0 + {0, 1/4, 1} h; the host tabulates per step the four points
 * t0, t0 + h/4, t0 + h/2, t0 + h (fp32 arithmetic of torchsde: t0 + c*h).  20 bytes. */
typedef struct {
  float t;             /* evaluation time                                             */
  float sin_t;         /* time features at t                                          */
  float cos_t;
  float frac;          /* fp32(t - knots[interval])                                   */
  int32_t interval;    /* clamp(bucketize(t, knots) - 1, 0, K-2)                      */
} snsde_point;

typedef struct snsde_plan snsde_plan;

int snsde_abi_version(void);
const char* snsde_last_error(void);

/* Number of floats in the weight blob for `desc`, and the blob layout: the tensors of the
 * reference state_dict in declaration order, each row-major as stored by nn.Linear
 * ([out,in] weight then [out] bias):
 *   benchmark: initial_network, linear_in, [emb], linears.0..L-2, linear_out, theta(1),
 *              [sigma(1) | sigma_diag(H)], [noise_t | noise_t.0, noise_t.2 | noise_y | noise_y.0, noise_y.2]
 *   tutorial : linear_X, emb, f_net._model.{0,2,..}, linear_out, noise_in, g_net._model.{0,2,..}
 *   latent   : linear_in [HH, H+1], linears.0..L-2, linear_out [H-1, HH], then the three buffers theta(1), mu(1),
 *              sigma(1) of the prior (latent_sde.py:35-37; not parameters: their gradient slots stay zero)
 * Returns a negative status for an invalid descriptor. */
int64_t snsde_weight_count(const snsde_model_desc* desc);

int snsde_plan_create(const snsde_model_desc* desc, int device, snsde_plan** out_plan);
int snsde_plan_destroy(snsde_plan* plan);

/* Copies and re-lays the blob into the plan's device images (sync only when on_device). */
int snsde_plan_set_weights(snsde_plan* plan, const float* blob, int64_t n_floats, int on_device,
                           void* stream);

/* Which kernel the plan will run: 0 = fp32 FMA, 1 = tcgen05 with resident weights, 2 = general tcgen05
 * (streamed weights / two M tiles / noise networks).  Negative on error. */
int snsde_plan_kernel_kind(const snsde_plan* plan);

/* For plans whose kernel kind is 0, after snsde_plan_set_weights: 0 = the shared-memory interpreter kernel (any shape /
 * method), 1 = the warp-owned kernel (hidden, hidden_hidden and control width <= 32, euler / milstein: one warp owns
 * its rows end to end with a helper warp preparing the state-independent inputs; used for launches of up to 4096 rows, larger
 * batches run on the interpreter kernel, whose 8-row groups amortise the weight reads).  Negative on error. */
int snsde_plan_fma_variant(const snsde_plan* plan);

/* The solve.  Replaces torchsde.sdeint as called at neuralsde.py:78-82.
 *   coeffs_dev       [B, K-1, 4C] fp32, packed cat(a,b,two_c,three_d); row b starts at
 *                    coeffs_dev + b*coeff_row_stride (floats).  May be NULL iff the model
 *                    never reads the control (input options 1,3,5).
 *   y0_dev           [B, H] fp32
 *   steps_host       [S] (S >= 0), emits_host [E]; the first `n_init_emits` emits apply to y0
 *                    itself (slot of ts[0]); n_out = number of output slots
 *   row_slot_dev     NULL: out_dev is [n_out, B, H] (the torchsde layout);
 *                    else int32[B]: out_dev is [B, H] and row b keeps only slot row_slot[b]
 *                    (fused `z_t.gather(final_index)` of neuralsde.py:115-116)
 *   points_host      method SRK: [S][4] evaluation points (see snsde_point); NULL otherwise
 *   dW_dev           NULL: increments drawn in-kernel from Philox4x32-10(seed; feature,
 *                    (row_offset+b)>>2, step); else explicit increments [S, B, H] (parity mode)
 *   dU_dev           method SRK with dW_dev != NULL: the space-time Levy integrals U[S, B, H] paired with dW
 *                    (torchsde `bm(t0, t1, return_U=True)`: U = h (W/2 + Hst), Hst ~ N(0, h/12)); with
 *                    dW_dev == NULL they come from a second Philox stream.  NULL otherwise
 *   row_offset       global index of local row 0 (batch sharding keeps the stream invariant)
 */
int snsde_forward(snsde_plan* plan,
                  const float* coeffs_dev, int64_t coeff_row_stride, int32_t n_knots,
                  const float* y0_dev, int32_t B,
                  const snsde_step* steps_host, int32_t S,
                  const snsde_emit* emits_host, int32_t E, int32_t n_init_emits, int32_t n_out,
                  const snsde_point* points_host,
                  const int32_t* row_slot_dev,
                  const float* dW_dev, const float* dU_dev, uint64_t seed, uint64_t row_offset,
                  float* out_dev, void* stream);

/* Materialises the increments the kernels would draw: dW_dev[s,b,j] = n1 * steps[s].sqrt_h for s<S, b<B, j<H.
 * Used to feed the oracle the identical Brownian path.
 * dU_dev (may be NULL): the SRK kernels' space-time Levy integrals for the same path,
 * U = h (W/2 + sqrt_h sqrt(1/12) n2) with n2 from the second Philox stream (tag 'SNSU'). */
int snsde_philox_fill(uint64_t seed, uint64_t row_offset, int32_t S, int32_t B, int32_t H,
                      const snsde_step* steps_host, float* dW_dev, float* dU_dev, int device, void* stream);

/* Sticky device-side status of the plan's solves since the last call (synchronises `stream`, then clears):
 *   bit 0: a tensor-core kernel met an operand beyond the fp16 range (|v| > 65504) of its split-precision
 *          format and saturated it - the affected rows are NOT within the parity tolerance; rerun the solve
 *          with precision = SNSDE_PRECISION_FP32.  (States of the reference models stay far below this:
 *          |z| <= |z0| + sum(h + 6.7 sqrt(h)) because drift and diffusion are tanh-clipped.)
 * Returns the flags (>= 0) or a negative snsde_status. */
int snsde_plan_status(snsde_plan* plan, void* stream);

/* Non-blocking variant: the flags raised by the solves that have COMPLETED so far (cleared when non-zero); touches
 * no stream.  The Python layer polls it on every call so that a saturated solve is reported on the next call
 * at the latest instead of staying silent. */
int snsde_plan_status_nowait(snsde_plan* plan);

/* ---- backward pass through the solve -------------------------------------------------------------
 * Replaces the autograd graph the reference builds through torchsde.sdeint when it trains
 * (benchmark_classification/common_sde.py:156-162: `pred_y = model(...); loss.backward()`;
 * benchmark_forecasting/common_sde.py:145-150).  Methods euler, milstein (for the elementwise diffusions: every noise
 * option but 14, 15, 18, 19; its diagonal term 0.5 (g v) dg/dy is differentiated through both factors, as torchsde's
 * create_graph vjp is) and srk (points_host as in snsde_forward; dU_dev beside dW_dev in table mode) - the default
 * method of the torch-ists wrappers (nsde_model.py:63-74, latent_sde.py:107-109).
 *
 * Protocol: run snsde_forward with a step plan that emits EVERY solver state (slot s+1 = state after step s;
 * out_dev = states [S+1, B, H]); form the requested outputs from those states (linear interpolation / per-row
 * final_index capture) in the caller; hand the cotangent of every state back here.
 *   states_dev        [S+1, B, H] saved solver states
 *   grad_states_dev   [S+1, B, H] dL/d state from the output selection (zero where a state is not an output)
 *   dW_dev / seed     the same Brownian path as the forward call (table, or Philox replay)
 *   grad_y0_dev       [B, H]   dL/d y0                                    (written)
 *   grad_blob_dev     [snsde_weight_count] dL/d weights in the blob layout of snsde_plan_set_weights (overwritten;
 *                     parameters that do not influence the solve - e.g. initial_network under input options
 *                     1,3,5 - get zeros, where autograd would report None)
 *   workspace_dev     snsde_backward_workspace_bytes(plan, B, S) bytes, 16-byte aligned, owned by the caller: the
 *                     per-op cotangents and activations behind the weight-gradient GEMMs
 * Coefficients are data: no gradient flows to coeffs_dev. */
int64_t snsde_backward_workspace_bytes(const snsde_plan* plan, int32_t B, int32_t S);
int snsde_backward(snsde_plan* plan,
                   const float* coeffs_dev, int64_t coeff_row_stride, int32_t n_knots, int32_t B,
                   const snsde_step* steps_host, int32_t S, const snsde_point* points_host,
                   const float* states_dev, const float* grad_states_dev,
                   const float* dW_dev, const float* dU_dev, uint64_t seed, uint64_t row_offset,
                   float* grad_y0_dev, float* grad_blob_dev,
                   void* workspace_dev, int64_t workspace_bytes, void* stream);

/* Control-path coefficients on device.  Replaces, for NaN-free inputs,
 * torchcde.hermite_cubic_coefficients_with_backward_differences(x, t) as the reference calls it at
 * benchmark_classification/datasets/common.py:82-84: x_dev [B,K,C], knots_dev [K] ->
 * coeffs_dev [B,K-1,4C] = cat(a,b,two_c,three_d), the packing snsde_forward reads. */
int snsde_hermite_coeffs(const float* x_dev, const float* knots_dev, int32_t B, int32_t K, int32_t C,
                         float* coeffs_dev, int device, void* stream);

/* Natural cubic spline coefficients on device: the in-tree builder of the forecasting benchmark,
 * benchmark_forecasting/controldiffeq/interpolate.py:7-53 (called at benchmark_forecasting/datasets/common.py:79-81,
 * packed by benchmark_forecasting/models_sde/neuralsde.py:161).  x_dev [B,K,C] NaN-free, knots_dev [K] ->
 * coeffs_dev [B,K-1,4C].  scratch_dev: 3*K floats owned by the caller (knot-only sweep factors). */
int snsde_natural_coeffs(const float* x_dev, const float* knots_dev, int32_t B, int32_t K, int32_t C,
                         float* coeffs_dev, float* scratch_dev, int device, void* stream);

/* The same builder for paths WITH missing values (NaN): benchmark_classification/controldiffeq/interpolate.py:56-153.
 * Per scalar series: an all-NaN series gives zero coefficients, a NaN at either end is imputed with the nearest
 * observation, the natural spline is built on the observed knots only and every original interval receives the piece
 * it lies in, re-centred at its own knot.  x_dev [B,K,C] (NaN = missing) -> coeffs_dev [B,K-1,4C]; no scratch. */
int snsde_natural_coeffs_missing(const float* x_dev, const float* knots_dev, int32_t B, int32_t K, int32_t C,
                                 float* coeffs_dev, int device, void* stream);

/* Missing-value (NaN) fill that torchcde applies before the Hermite builder (linear_interpolation_coeffs):
 * linear in t between observed neighbours, first observed value at the head, forward fill at the tail.
 * x_dev [B,K,C] -> out_dev [B,K,C] (must not alias x_dev). */
int snsde_fill_missing(const float* x_dev, const float* knots_dev, int32_t B, int32_t K, int32_t C,
                       float* out_dev, int device, void* stream);

/* ---- the seam's neighbours, eval mode ----------------------------------------------------------------
 * z0 = initial_network(X(times[0])): `_prepare_initial_state`, benchmark_classification/models_sde/neuralsde.py:63-69
 * (= benchmark_forecasting/models_sde/neuralsde.py:137-143).  coeffs_dev as in snsde_forward; (interval, frac) locate
 * times[0] in the knots (0, 0.0f when times[0] is the first knot); W_dev [H, C], b_dev [H] = initial_network; z0_dev [B, H]. */
int snsde_initial_state(const float* coeffs_dev, int64_t coeff_row_stride, int32_t B, int32_t C, int32_t n_knots,
                        int32_t interval, float frac, const float* W_dev, const float* b_dev, int32_t H,
                        float* z0_dev, int device, void* stream);

/* pred = Linear2(relu(bn(Linear1(pre(z))))) for R rows: the read-out heads of the three wrappers in EVAL mode
 * (neuralsde.py:59-61,119: Linear, BatchNorm1d, ReLU, Dropout, Linear; forecasting :133-136,185: Linear, ReLU, Linear;
 * torch-ists nsde_model.py:52-55,83: Tanh, Linear, ReLU, Linear).  pre_tanh: apply tanh to z first.  bn_scale_dev /
 * bn_shift_dev [H1] (both or neither): BatchNorm1d in eval mode as  x * scale + shift  with
 * scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale.  W1 [H1, H], W2 [O, H1] (nn.Linear layout). */
int snsde_readout_head(const float* z_dev, int64_t R, int32_t H, int32_t pre_tanh,
                       const float* W1_dev, const float* b1_dev, const float* bn_scale_dev, const float* bn_shift_dev,
                       int32_t H1, const float* W2_dev, const float* b2_dev, int32_t O, float* out_dev,
                       int device, void* stream);

/* Number of engine kernels launched by this plan so far (for bench.py's gpu_launches). */
int64_t snsde_plan_launch_count(const snsde_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* SNSDE_H_ */
