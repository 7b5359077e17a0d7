// Control-path coefficient construction on device: Hermite cubic with backward differences,
// natural cubic spline, and the NaN fill that precedes the Hermite builder.
//
// Replaces the data-prep call torchcde.hermite_cubic_coefficients_with_backward_differences(x, t)
// (/root/reference/benchmark_classification/datasets/common.py:82-84, tests/test_neuralsde_core_alignment.py:64)
// for NaN-free inputs: x [B, K, C] -> coeffs [B, K-1, 4C] = cat(a, b, two_c, three_d), the layout the solve reads.
// One pass, HBM-bound: 4 bytes read (+ neighbours from L1/L2) and 16 bytes written per (row, interval, channel).
#include <cuda_runtime.h>
#include "../../include/snsde.h"
#include "snsde_host.cuh"

namespace snsde {

// blockIdx.y = batch row, thread = one (k, c) of that row with c fastest (32-bit index math): reads are coalesced
// along C (the x[k-1], x[k+1] neighbours come from L1/L2), and the 4C floats of one (row, interval) are written by
// C consecutive threads, i.e. into one contiguous 16C-byte run.
__global__ void __launch_bounds__(256) hermite_coeffs_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                             float* __restrict__ out, int B, int K, int C) {
  const int per_row = (K - 1) * C;
  const float* xb = x + (size_t)blockIdx.y * K * C;
  float* ob = out + (size_t)blockIdx.y * (K - 1) * 4 * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_row; i += gridDim.x * blockDim.x) {
    const int k = i / C, c = i - k * C;
    const float* xr = xb + i;                                            // (k, c)
    const float x0 = xr[0], x1 = xr[C];
    const float h = __fsub_rn(t[k + 1], t[k]);
    const float dx = __fsub_rn(x1, x0);
    const float m1 = __fdiv_rn(dx, h);                                   // this interval's secant slope
    float m0 = m1;                                                       // previous interval's (first: its own)
    if (k > 0) m0 = __fdiv_rn(__fsub_rn(x0, xr[-C]), __fsub_rn(t[k], t[k - 1]));
    // two_c = 2 (3 (dx/h - m0) - m1 + m0) / h ;  three_d = (1/h^2) (m1 - m0) - two_c / h     (op order of torchcde)
    const float u = __fsub_rn(__fmul_rn(3.f, __fsub_rn(m1, m0)), m1);
    const float two_c = __fdiv_rn(__fmul_rn(2.f, __fadd_rn(u, m0)), h);
    const float three_d = __fsub_rn(__fmul_rn(__fdiv_rn(1.f, __fmul_rn(h, h)), __fsub_rn(m1, m0)), __fdiv_rn(two_c, h));
    float* o = ob + (size_t)k * 4 * C + c;
    o[0] = x0;
    o[C] = m0;
    o[2 * C] = two_c;
    o[3 * C] = three_d;
  }
}


// ------------------------------------------------------------------------------------------------------------
// Natural cubic spline (second derivative zero at both ends): the in-tree builder of the forecasting benchmark,
// /root/reference/benchmark_forecasting/controldiffeq/interpolate.py:7-53 (tridiagonal solve misc.py:13-66),
// called at benchmark_forecasting/datasets/common.py:79-81 and fed to the solve as cat(a,b,two_c,three_d)
// (benchmark_forecasting/models_sde/neuralsde.py:161).
//
// The knot slopes k_i solve a tridiagonal system whose matrix depends on the knots only; Thomas sweep:
//   w_i = r_{i-1} / cp_{i-1},  cp_i = diag_i - w_i r_{i-1},  d_i = rhs_i - w_i d_{i-1};   k_i = (d_i - r_i k_{i+1}) / cp_i
// Stage 1 (one thread, K steps): r, cp, w from the knots.  Stage 2: one thread per (row, channel) series runs the
// two sweeps; the forward intermediates d_i are parked in the `b` slot of the output (overwritten by k_i on the
// way back), so the pass costs 4 B read + 16 B written (+ 4 B written and read again for d) per element and
// needs no scratch beyond 3K floats.  Operation order = the torch-op chain in data.natural_cubic_coeffs
// (no FMA contraction): results are bit-identical to it.
__global__ void natural_knot_sweep_kernel(const float* __restrict__ t, int K, float* __restrict__ r, float* __restrict__ cp,
                                          float* __restrict__ w) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int i = 0; i < K - 1; ++i) r[i] = __fdiv_rn(1.f, __fsub_rn(t[i + 1], t[i]));
  r[K - 1] = 0.f;
  float prev = 0.f;
  for (int i = 0; i < K; ++i) {
    // diag = zeros; diag[:-1] += 2r; diag[1:] += 2r
    float diag = 0.f;
    if (i < K - 1) diag = __fadd_rn(diag, __fmul_rn(2.f, r[i]));
    if (i > 0) diag = __fadd_rn(diag, __fmul_rn(2.f, r[i - 1]));
    if (i == 0) { w[0] = 0.f; cp[0] = diag; prev = diag; continue; }
    const float wi = __fdiv_rn(r[i - 1], prev);
    const float c = __fsub_rn(diag, __fmul_rn(wi, r[i - 1]));
    w[i] = wi; cp[i] = c; prev = c;
  }
}

__global__ void __launch_bounds__(128) natural_coeffs_kernel(const float* __restrict__ x, const float* __restrict__ r,
                                                             const float* __restrict__ cp, const float* __restrict__ w,
                                                             float* __restrict__ out, int B, int K, int C) {
  const long long series = (long long)blockIdx.x * blockDim.x + threadIdx.x;        // (row, channel), channel fastest
  if (series >= (long long)B * C) return;
  const int b = (int)(series / C), c = (int)(series - (long long)b * C);
  const float* xs = x + (size_t)b * K * C + c;
  float* os = out + (size_t)b * (K - 1) * 4 * C + c;
  if (K == 2) {
    const float x0 = xs[0];
    os[0] = x0; os[C] = __fmul_rn(__fsub_rn(xs[C], x0), r[0]); os[2 * C] = 0.f; os[3 * C] = 0.f;
    return;
  }
  // forward sweep: rhs_i = [3 dx_i r_i^2] + [3 dx_{i-1} r_{i-1}^2];  d_i = rhs_i - w_i d_{i-1}
  // The loads do not depend on the recurrence: they are issued U at a time ahead of the dependent chain (each
  // thread walks its series alone, so without this the sweep is one exposed memory latency per knot).
  constexpr int U = 8;
  auto s_of = [&](float xa, float xb, float ri) { return __fmul_rn(__fmul_rn(3.f, __fsub_rn(xb, xa)), __fmul_rn(ri, ri)); };
  float xprev = xs[0];
  float sprev, dprev;
  {                                                        // i = 0: rhs_0 = s_0
    const float xn = xs[C];
    sprev = s_of(xprev, xn, r[0]);
    dprev = __fadd_rn(0.f, sprev);
    os[C] = dprev;
    xprev = xn;
  }
  int i = 1;
  for (; i + U <= K - 1; i += U) {                         // interior knots 1 .. K-2
    float xv[U], rv[U], wv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { xv[u] = xs[(size_t)(i + u + 1) * C]; rv[u] = r[i + u]; wv[u] = w[i + u]; }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float sc = s_of(xprev, xv[u], rv[u]);
      const float d = __fsub_rn(__fadd_rn(__fadd_rn(0.f, sc), sprev), __fmul_rn(wv[u], dprev));
      os[(size_t)(i + u) * 4 * C + C] = d;                 // parked in the `b` slot
      dprev = d; sprev = sc; xprev = xv[u];
    }
  }
  for (; i < K - 1; ++i) {
    const float xn = xs[(size_t)(i + 1) * C];
    const float sc = s_of(xprev, xn, r[i]);
    const float d = __fsub_rn(__fadd_rn(__fadd_rn(0.f, sc), sprev), __fmul_rn(w[i], dprev));
    os[(size_t)i * 4 * C + C] = d;
    dprev = d; sprev = sc; xprev = xn;
  }
  dprev = __fsub_rn(__fadd_rn(0.f, sprev), __fmul_rn(w[K - 1], dprev));          // i = K-1: rhs = s_{K-2}
  // backward sweep + coefficients of interval i from (k_i, k_{i+1})
  float knext = __fdiv_rn(dprev, cp[K - 1]);
  float xn = xs[(size_t)(K - 1) * C];
  auto finish = [&](int ii, float dv, float x0, float ri, float cpi) {
    float* o = os + (size_t)ii * 4 * C;
    const float k0 = __fdiv_rn(__fsub_rn(dv, __fmul_rn(ri, knext)), cpi);
    const float dx = __fsub_rn(xn, x0);
    // two_c = (6 dx r - 4 k0 - 2 k1) r ;  three_d = (-6 dx r + 3 (k0 + k1)) r^2
    const float two_c = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(__fmul_rn(6.f, dx), ri), __fmul_rn(4.f, k0)), __fmul_rn(2.f, knext)), ri);
    const float three_d = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(-6.f, dx), ri), __fmul_rn(3.f, __fadd_rn(k0, knext))), __fmul_rn(ri, ri));
    o[0] = x0; o[C] = k0; o[2 * C] = two_c; o[3 * C] = three_d;
    knext = k0; xn = x0;
  };
  i = K - 2;
  for (; i - U + 1 >= 0; i -= U) {
    float dv[U], xv[U], rv[U], cv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      dv[u] = os[(size_t)(i - u) * 4 * C + C]; xv[u] = xs[(size_t)(i - u) * C]; rv[u] = r[i - u]; cv[u] = cp[i - u];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) finish(i - u, dv[u], xv[u], rv[u], cv[u]);
  }
  for (; i >= 0; --i) finish(i, os[(size_t)i * 4 * C + C], xs[(size_t)i * C], r[i], cp[i]);
}

// ------------------------------------------------------------------------------------------------------------
// Natural cubic spline WITH missing values: benchmark_classification/controldiffeq/interpolate.py:56-153
// (`_natural_cubic_spline_coeffs_with_missing_values`).  The reference treats every scalar series on its own: an
// all-NaN series gives zero coefficients; a NaN at either end is imputed with the nearest observation; the spline is
// built on the OBSERVED knots only (their own irregular grid) and every original interval [t_i, t_{i+1}) receives the
// piece it lies in, re-centred at t_i.  One thread per series: forward Thomas sweep over the observed knots (modified
// diagonal and right-hand side parked in the b / two_c slots of the knot's own interval), backward sweep that turns each
// observed interval into the coefficients of the original intervals it covers.
__global__ void __launch_bounds__(128) natural_coeffs_missing_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                                     float* __restrict__ out, int B, int K, int C) {
  const long long series = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (series >= (long long)B * C) return;
  const int b = (int)(series / C), c = (int)(series - (long long)b * C);
  const float* xs = x + (size_t)b * K * C + c;
  float* os = out + (size_t)b * (K - 1) * 4 * C + c;
  auto xat = [&](int i) { return xs[(size_t)i * C]; };
  int first = -1, last = -1;
  for (int i = 0; i < K; ++i) { const float v = xat(i); if (v == v) { if (first < 0) first = i; last = i; } }
  if (first < 0) {                                         // nothing observed: constant zero path
    for (int i = 0; i < K - 1; ++i) { float* o = os + (size_t)i * 4 * C; o[0] = 0.f; o[C] = 0.f; o[2 * C] = 0.f; o[3 * C] = 0.f; }
    return;
  }
  const float x_first = xat(first), x_last = xat(last);
  // value at knot i after imputing the two ends; "observed" = not NaN, or an end point
  auto obs = [&](int i) { const float v = xat(i); return i == 0 || i == K - 1 || v == v; };
  auto val = [&](int i) { const float v = xat(i); return (v == v) ? v : (i == 0 ? x_first : x_last); };
  auto next_obs = [&](int i) { int j = i + 1; while (!obs(j)) ++j; return j; };      // i < K-1
  auto prev_obs = [&](int i) { int j = i - 1; while (!obs(j)) --j; return j; };      // i > 0
  // expands the piece (a, bb, c2, d3) anchored at observed knot j0 onto the original intervals [j0, j1)
  auto expand = [&](int j0, int j1, float a, float bb, float c2, float d3) {
    const float t0 = t[j0];
    for (int i = j0; i < j1; ++i) {
      const float off = t0 - t[i];
      const float a_inner = (0.5f * c2 - d3 * off / 3.f) * off;
      float* o = os + (size_t)i * 4 * C;
      o[0] = a + (a_inner - bb) * off;
      o[C] = bb + (d3 * off - c2) * off;
      o[2 * C] = c2 - 2.f * d3 * off;
      o[3 * C] = d3;
    }
  };
  const int second = next_obs(0);
  if (second == K - 1) {                                   // two observed knots: the straight line (interpolate.py:15-19)
    const float x0 = val(0);
    expand(0, K - 1, x0, (val(K - 1) - x0) / (t[K - 1] - t[0]), 0.f, 0.f);
    return;
  }
  // ---- forward sweep over the observed knots j (previous observed p, next observed n) ----
  //   diag_j = 2 (r_j + r_p) [ends: one term], rhs_j = 3 dx_j r_j^2 + 3 dx_p r_p^2, w = r_p / nd_p,
  //   nd_j = diag_j - w r_p, nb_j = rhs_j - w nb_p            (oracle/spline.py, misc.py:52-64)
  float nd_prev, nb_prev, r_prev, s_prev;
  {
    const float r0 = 1.f / (t[second] - t[0]);
    const float s0 = 3.f * (val(second) - val(0)) * (r0 * r0);
    nd_prev = 2.f * r0; nb_prev = s0; r_prev = r0; s_prev = s0;
    os[C] = nb_prev; os[2 * C] = nd_prev;                  // parked at the knot's own interval (knot 0)
  }
  int j = second;
  while (j < K - 1) {
    const int n = next_obs(j);
    const float rj = 1.f / (t[n] - t[j]);
    const float sj = 3.f * (val(n) - val(j)) * (rj * rj);
    const float w = r_prev / nd_prev;
    const float nd = 2.f * (rj + r_prev) - w * r_prev;
    const float nb = (sj + s_prev) - w * nb_prev;
    os[(size_t)j * 4 * C + C] = nb; os[(size_t)j * 4 * C + 2 * C] = nd;
    nd_prev = nd; nb_prev = nb; r_prev = rj; s_prev = sj;
    j = n;
  }
  // last observed knot (index K-1): diag = 2 r_p, rhs = s_p
  float k_next;
  {
    const float w = r_prev / nd_prev;
    const float nd = 2.f * r_prev - w * r_prev;
    const float nb = s_prev - w * nb_prev;
    k_next = nb / nd;
  }
  // ---- backward sweep: k_j, then the piece on [j, n) ----
  int n = K - 1;
  j = prev_obs(K - 1);
  for (;;) {
    const float nb = os[(size_t)j * 4 * C + C], nd = os[(size_t)j * 4 * C + 2 * C];
    const float rj = 1.f / (t[n] - t[j]);
    const float kj = (nb - rj * k_next) / nd;
    const float xj = val(j);
    const float six_dx = 2.f * (3.f * (val(n) - xj));
    const float two_c = (six_dx * rj - 4.f * kj - 2.f * k_next) * rj;
    const float three_d = (-six_dx * rj + 3.f * (kj + k_next)) * (rj * rj);
    expand(j, n, xj, kj, two_c, three_d);
    k_next = kj;
    if (j == 0) break;
    n = j;
    j = prev_obs(j);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Missing values (NaN) before the Hermite builder: torchcde fills them through linear_interpolation_coeffs -
// linear in t between the observed neighbours, first observed value before the first, last observed value after
// the last (forward fill); an all-NaN series stays NaN.  One thread per (row, channel) series, one forward scan.
__global__ void __launch_bounds__(128) fill_missing_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                           float* __restrict__ out, int B, int K, int C) {
  const long long series = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (series >= (long long)B * C) return;
  const int b = (int)(series / C), c = (int)(series - (long long)b * C);
  const float* xs = x + (size_t)b * K * C + c;
  float* os = out + (size_t)b * K * C + c;
  int last = -1;                 // index of the last observed value
  float vlast = 0.f;
  for (int i = 0; i < K; ++i) {
    const float v = xs[(size_t)i * C];
    if (v != v) continue;
    if (last < 0) {
      for (int j = 0; j < i; ++j) os[(size_t)j * C] = v;                 // head: first observed value
    } else if (i - last > 1) {
      const float t0 = t[last], span = __fsub_rn(t[i], t0), dv = __fsub_rn(v, vlast);
      for (int j = last + 1; j < i; ++j)                                 // v[lo] + w * (v[hi] - v[lo])
        os[(size_t)j * C] = __fadd_rn(vlast, __fmul_rn(__fdiv_rn(__fsub_rn(t[j], t0), span), dv));
    }
    os[(size_t)i * C] = v;
    last = i; vlast = v;
  }
  if (last < 0) {
    for (int j = 0; j < K; ++j) os[(size_t)j * C] = xs[(size_t)j * C];   // nothing observed: stays NaN
  } else {
    for (int j = last + 1; j < K; ++j) os[(size_t)j * C] = vlast;        // tail: forward fill
  }
}

}  // namespace snsde

using snsde::fail;

extern "C" int snsde_hermite_coeffs(const float* x_dev, const float* knots_dev, int32_t B, int32_t K, int32_t C,
                                    float* coeffs_dev, int device, void* stream_v) {
  SNSDE_API_BEGIN
  if (!x_dev || !knots_dev || !coeffs_dev) return fail(SNSDE_ERR_BAD_ARG, "hermite_coeffs: x/knots/coeffs is NULL");
  if (B < 1 || K < 2 || C < 1) return fail(SNSDE_ERR_BAD_ARG, "hermite_coeffs: need B >= 1, K >= 2, C >= 1 (got B=%d K=%d C=%d)", B, K, C);
  if ((long long)(K - 1) * C > 0x7fffffffLL || B > 65535 * 64) return fail(SNSDE_ERR_BAD_ARG, "hermite_coeffs: tensor too large (B=%d K=%d C=%d)", B, K, C);
  snsde::DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.err));
  const int per_row = (K - 1) * C;
  const int gx = (per_row + 255) / 256 < 64 ? (per_row + 255) / 256 : 64;
  for (int b0 = 0; b0 < B; b0 += 65535) {                                  // gridDim.y limit
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    snsde::hermite_coeffs_kernel<<<dim3(gx, nb), 256, 0, (cudaStream_t)stream_v>>>(
        x_dev + (size_t)b0 * K * C, knots_dev, coeffs_dev + (size_t)b0 * (K - 1) * 4 * C, nb, K, C);
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "hermite_coeffs launch: %s", cudaGetErrorString(e));
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

extern "C" int snsde_natural_coeffs(const float* x_dev, const float* knots_dev, int32_t B, int32_t K, int32_t C,
                                    float* coeffs_dev, float* scratch_dev, int device, void* stream_v) {
  SNSDE_API_BEGIN
  if (!x_dev || !knots_dev || !coeffs_dev || !scratch_dev) return fail(SNSDE_ERR_BAD_ARG, "natural_coeffs: x/knots/coeffs/scratch is NULL");
  if (B < 1 || K < 2 || C < 1) return fail(SNSDE_ERR_BAD_ARG, "natural_coeffs: need B >= 1, K >= 2, C >= 1 (got B=%d K=%d C=%d)", B, K, C);
  if ((long long)B * C > 0x7fffffffLL * 128LL) return fail(SNSDE_ERR_BAD_ARG, "natural_coeffs: tensor too large");
  snsde::DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.err));
  cudaStream_t st = (cudaStream_t)stream_v;
  float *r = scratch_dev, *cp = scratch_dev + K, *w = scratch_dev + 2 * (size_t)K;
  snsde::natural_knot_sweep_kernel<<<1, 32, 0, st>>>(knots_dev, K, r, cp, w);
  const long long n = (long long)B * C;
  snsde::natural_coeffs_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(x_dev, r, cp, w, coeffs_dev, B, K, C);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "natural_coeffs launch: %s", cudaGetErrorString(e));
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

extern "C" int snsde_natural_coeffs_missing(const float* x_dev, const float* knots_dev, int32_t B, int32_t K, int32_t C,
                                            float* coeffs_dev, int device, void* stream_v) {
  SNSDE_API_BEGIN
  if (!x_dev || !knots_dev || !coeffs_dev) return fail(SNSDE_ERR_BAD_ARG, "natural_coeffs_missing: x/knots/coeffs is NULL");
  if (B < 1 || K < 2 || C < 1) return fail(SNSDE_ERR_BAD_ARG, "natural_coeffs_missing: need B >= 1, K >= 2, C >= 1 (got B=%d K=%d C=%d)", B, K, C);
  const long long n = (long long)B * C;
  if (n > 0x7fffffffLL * 128LL) return fail(SNSDE_ERR_BAD_ARG, "natural_coeffs_missing: tensor too large");
  snsde::DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.err));
  snsde::natural_coeffs_missing_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream_v>>>(x_dev, knots_dev, coeffs_dev, B, K, C);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "natural_coeffs_missing launch: %s", cudaGetErrorString(e));
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

extern "C" int snsde_fill_missing(const float* x_dev, const float* knots_dev, int32_t B, int32_t K, int32_t C,
                                  float* out_dev, int device, void* stream_v) {
  SNSDE_API_BEGIN
  if (!x_dev || !knots_dev || !out_dev) return fail(SNSDE_ERR_BAD_ARG, "fill_missing: x/knots/out is NULL");
  if (B < 1 || K < 1 || C < 1) return fail(SNSDE_ERR_BAD_ARG, "fill_missing: need B, K, C >= 1 (got B=%d K=%d C=%d)", B, K, C);
  if (x_dev == out_dev) return fail(SNSDE_ERR_BAD_ARG, "fill_missing: out must not alias x");
  const long long n = (long long)B * C;
  if (n > 0x7fffffffLL * 128LL) return fail(SNSDE_ERR_BAD_ARG, "fill_missing: tensor too large");
  snsde::DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.err));
  snsde::fill_missing_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream_v>>>(x_dev, knots_dev, out_dev, B, K, C);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "fill_missing launch: %s", cudaGetErrorString(e));
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}
