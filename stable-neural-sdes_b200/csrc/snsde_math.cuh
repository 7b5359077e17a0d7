// Device math shared by the FMA and tcgen05 kernels: activations, nan_to_num, the elementwise
// diffusion g(t,y) of neuralsde.py:233-307 and its y-derivative for the Milstein term.
#pragma once
#include <math.h>
#include "snsde_common.cuh"

namespace snsde {

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_RELU) return v < 0.f ? 0.f : v;                       // NaN passes like torch.relu
  if (act == ACT_LIPSWISH) return 0.909f * (v / (1.f + expf(-v)));     // 0.909 * silu(v)
  return v;
}

__device__ __forceinline__ float nan_to_num_f(float v) {               // torch.nan_to_num defaults, branch-free
  v = (v == v) ? v : 0.f;                                              // NaN -> 0
  return fminf(fmaxf(v, -3.4028234663852886e38f), 3.4028234663852886e38f);   // +-inf -> +-FLT_MAX
}
__device__ __forceinline__ bool is_finite_f(float v) { return fabsf(v) <= 3.4028234663852886e38f; }   // false for NaN, +-inf

// tanh with ~4e-7 RELATIVE accuracy on the fast path: odd Taylor polynomial below 0.25 (the multiplicative
// models - tanh(sigma a y), z*tanh(y) - live on the relative accuracy near 0: an absolute 1e-7 error there is
// an O(1) relative error once |y| ~ 1e-7 and the trajectories separate), one ex2 + one rcp above
// (1 - 2/(1+e^{2|x|})); exact limits at +-inf, NaN passes.
__device__ __forceinline__ float tanh_fast(float x) {
  const float ax = fabsf(x);
  const float e = __expf(2.f * ax);
  const float big = 1.f - __fdividef(2.f, e + 1.f);
  const float x2 = x * x;
  const float small = ax * fmaf(x2, fmaf(x2, fmaf(x2, -0.053968254f, 0.13333334f), -0.33333334f), 1.f);
  return copysignf(ax < 0.25f ? small : big, x);
}

// Diffusion value g and (for Milstein) d g / d y at one element.  FAST selects tanh_fast.
template <bool FAST>
__device__ __forceinline__ void diffusion_eval(const TailOp& t, float coef, float y, float tt, float& g, float& dgdy) {
  float raw, draw;
  switch (t.special) {
    case SP_ZERO: raw = 0.f; draw = 0.f; break;
    case SP_SQRT: raw = sqrtf(y); draw = 0.5f / raw; break;
    case SP_CUBE: raw = y * y * y; draw = 3.f * y * y; break;
    case SP_SIGMOID: raw = 1.f / (1.f + expf(-y)); draw = raw * (1.f - raw); break;
    case SP_RELU: raw = y < 0.f ? 0.f : y; draw = y > 0.f ? 1.f : 0.f; break;
    default:
      if (t.mult == MU_Y) { raw = coef * y; draw = coef; }
      else if (t.mult == MU_TY) { raw = tt * y; draw = tt; }
      else if (t.mult == MU_T) { raw = coef * tt; draw = 0.f; }
      else { raw = coef; draw = 0.f; }
  }
  const bool state_dep = (t.special >= SP_SQRT) || (t.special == SP_NONE && (t.mult == MU_Y || t.mult == MU_TY));
  if (t.bounded) {
    const bool fin = is_finite_f(raw);
    const float arg = t.s_theta * nan_to_num_f(raw);
    g = FAST ? tanh_fast(arg) : tanhf(arg);
    // autograd chain of tanh(s * nan_to_num(raw)): (1-g^2) * s * isfinite(raw) * raw'.
    // A g that does not depend on y has no gradient path: torchsde's vjp returns zeros.
    dgdy = state_dep ? ((1.f - g * g) * t.s_theta) * (fin ? 1.f : 0.f) * draw : 0.f;
  } else {
    g = raw;
    dgdy = draw;
  }
}

// Cotangents of the elementwise diffusion g(coef, y, t) (what autograd propagates through neuralsde.py:233-307):
// given a_g = dL/dg returns dL/dy (direct path only), dL/dcoef, and dL/d sigmoid(theta).
// nan_to_num passes a gradient only where its input is finite; relu where its output is positive.
__device__ __forceinline__ void diffusion_backward(const TailOp& t, float coef, float y, float tt, float g, float a_g,
                                                   float& a_y, float& a_coef, float& a_sth) {
  float raw, dy, dc;
  switch (t.special) {
    case SP_ZERO: raw = 0.f; dy = 0.f; dc = 0.f; break;
    case SP_SQRT: raw = sqrtf(y); dy = 0.5f / raw; dc = 0.f; break;
    case SP_CUBE: raw = y * y * y; dy = 3.f * y * y; dc = 0.f; break;
    case SP_SIGMOID: raw = 1.f / (1.f + expf(-y)); dy = raw * (1.f - raw); dc = 0.f; break;
    case SP_RELU: raw = y < 0.f ? 0.f : y; dy = y > 0.f ? 1.f : 0.f; dc = 0.f; break;
    default:
      if (t.mult == MU_Y) { raw = coef * y; dy = coef; dc = y; }
      else if (t.mult == MU_TY) { raw = tt * y; dy = tt; dc = 0.f; }
      else if (t.mult == MU_T) { raw = coef * tt; dy = 0.f; dc = tt; }
      else { raw = coef; dy = 0.f; dc = 1.f; }
  }
  float a_raw;
  if (t.bounded) {
    const float a_arg = a_g * (1.f - g * g);
    a_sth = a_arg * nan_to_num_f(raw);
    a_raw = is_finite_f(raw) ? a_arg * t.s_theta : 0.f;
  } else {
    a_sth = 0.f;
    a_raw = a_g;
  }
  a_y = a_raw * dy;
  a_coef = a_raw * dc;
}

// Cotangents of the diagonal Milstein term  T = 0.5 v g dg/dy,  v = dW^2 - h  (torchsde Milstein.step with
// ForwardSDE.gdg_prod_diagonal: vjp(g, y, grad_outputs = g v, create_graph = True) - under autograd the gradient flows
// through BOTH factors, i.e. through the second derivative of g).  For the elementwise diffusions of neuralsde.py:233-307
// g = tanh(s n(raw(coef, y))) [bounded] or g = raw [unbounded]; with r1 = d raw/dy, r2 = d2 raw/dy2, rc = d raw/d coef,
// r1c = d r1/d coef and q = 1 - g^2:
//   g'  = q s r1          g''  = -2 g g' s r1 + q s r2
//   g_c = q s rc          g'_c = -2 g g_c s r1 + q s r1c
//   g_s = q raw           g'_s = -2 g g_s s r1 + q r1            (s = sigmoid(theta))
//   dT/dy = 0.5 v (g'^2 + g g''),  dT/dcoef = 0.5 v (g_c g' + g g'_c),  dT/ds = 0.5 v (g_s g' + g g'_s).
// nan_to_num passes no gradient where raw is not finite; a g that does not depend on y has T = 0.
__device__ __forceinline__ void milstein_backward(const TailOp& t, float coef, float y, float tt, float g, float v, float a_T,
                                                  float& a_y, float& a_coef, float& a_sth) {
  a_y = 0.f; a_coef = 0.f; a_sth = 0.f;
  float raw, r1, r2 = 0.f, rc = 0.f, r1c = 0.f;
  switch (t.special) {
    case SP_SQRT: raw = sqrtf(y); r1 = 0.5f / raw; r2 = -0.25f / (raw * y); break;
    case SP_CUBE: raw = y * y * y; r1 = 3.f * y * y; r2 = 6.f * y; break;
    case SP_SIGMOID: raw = 1.f / (1.f + expf(-y)); r1 = raw * (1.f - raw); r2 = r1 * (1.f - 2.f * raw); break;
    case SP_RELU: raw = y < 0.f ? 0.f : y; r1 = y > 0.f ? 1.f : 0.f; break;
    case SP_NONE:
      if (t.mult == MU_Y) { raw = coef * y; r1 = coef; rc = y; r1c = 1.f; }
      else if (t.mult == MU_TY) { raw = tt * y; r1 = tt; }
      else return;                                           // state-independent g: no Milstein term
      break;
    default: return;                                         // SP_ZERO
  }
  const float hv = 0.5f * v * a_T;
  if (t.bounded) {
    if (!is_finite_f(raw)) return;
    const float s = t.s_theta, q = 1.f - g * g;
    const float g1 = q * s * r1, g2 = -2.f * g * g1 * s * r1 + q * s * r2;
    const float gc = q * s * rc, g1c = -2.f * g * gc * s * r1 + q * s * r1c;
    const float gs = q * raw, g1s = -2.f * g * gs * s * r1 + q * r1;
    a_y = hv * (g1 * g1 + g * g2);
    a_coef = hv * (gc * g1 + g * g1c);
    a_sth = hv * (gs * g1 + g * g1s);
  } else {                                                    // g = raw
    a_y = hv * (r1 * r1 + raw * r2);
    a_coef = hv * (rc * r1 + raw * r1c);
  }
}

// d act(v) / dv given the pre-activation v (LipSwish) or the activated output (ReLU: out > 0).
__device__ __forceinline__ float act_grad(float pre, float post, int act) {
  if (act == ACT_RELU) return post > 0.f ? 1.f : 0.f;
  if (act == ACT_LIPSWISH) {
    const float sg = 1.f / (1.f + expf(-pre));
    return 0.909f * (sg + pre * sg * (1.f - sg));
  }
  return 1.f;
}

}  // namespace snsde
