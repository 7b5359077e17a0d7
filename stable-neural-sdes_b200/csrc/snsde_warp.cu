// Warp-resident fp32 kernel for hidden sizes <= 32 (the north star's "warp-shuffle FMA path for hidden < 64"):
// ONE warp owns R batch rows end to end - weights, activations and the SDE state all live in registers, a layer's
// inputs travel between lanes by warp shuffles, and there is no shared memory and no barrier of any kind in the
// time loop.
//
// Replaces, like snsde_fma.cu, the Python step loop of torchsde.sdeint (Euler.step / Milstein.step) with the
// per-step Diffusion_model.f/g evaluation (/root/reference/benchmark_classification/models_sde/neuralsde.py:295-307),
// torchcde.CubicSpline.evaluate (:296) and the tutorial NeuralLSDEFunc (notebook cell 7) - for the shapes where the
// interpreter kernel is pure latency: a row group there is one warp that walks ~1700 instructions per step through
// shared-memory round trips (profiles/r2_c1_fma_kernel.txt: 6 us per solver step, issue slots 15 % busy,
// 82 KB of SASS against a 32 KB instruction cache).
//
// Decomposition
//   * the compiled dense-op program (snsde_api.cu) is flattened on the host into <= NMV mat-vecs of at most
//     32 x 32 (an op with two sources, emb(cat(yy, Xt)), is two mat-vecs accumulating into one output);
//   * lane j holds column j of every mat-vec's transposed weight image in registers (32 per mat-vec, loaded once),
//     the bias and the two time-feature weights; activations are one register per (buffer, row): lane j = feature j;
//   * out[j] = b[j] + sum_k shfl(a, k) * w[k][j]: 32 shuffles + 32 dependent FMAs, k ascending from a single
//     accumulator - the summation order of the interpreter kernel, so both kernels return bit-identical trajectories
//     (tests/test_engine_gpu.py::test_warp_kernel_is_bit_identical_to_the_interpreter);
//   * lane c < C reads its channel's four spline coefficients of the NEXT step straight from global memory into
//     registers while the current step computes (no staging buffer); step records, the first emit of a step and the
//     row-independent diffusion coefficient are requested a step ahead as well;
//   * Brownian increments: the same Philox4x32-10 stream as every other kernel (keyed by feature, global row >> 2, step).
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>

#include "snsde_fma.cuh"
#include "snsde_warp.cuh"

namespace snsde {

template <int R>
__device__ __forceinline__ float pick_buf(const float (&v)[kNumRowBufs][R], int id, int r) {
  float a = v[0][r];
#pragma unroll
  for (int b = 1; b < kNumRowBufs; ++b) a = (id == b) ? v[b][r] : a;
  return a;
}

template <int NMV, int R>
__global__ void __launch_bounds__(128, 1) snsde_warp_kernel(const FmaParams p, const WarpProg wp) {
  const Program& pg = p.prog;
  const TailOp& t = pg.tail;
  const int H = pg.H, C = pg.C;
  const int lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
  if (row0 >= p.B) return;                                  // whole warp idle
  auto grow = [&](int r) { return min(row0 + r, p.B - 1); };
  const bool jact = lane < H;

  // ---- weights: registers, for the whole trajectory ----
  float w[NMV][32], bias[NMV], tws[NMV], twc[NMV];
#pragma unroll
  for (int i = 0; i < NMV; ++i) {
    bias[i] = 0.f; tws[i] = 0.f; twc[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) w[i][k] = 0.f;
    if (i < wp.n_mv) {
      const WarpMv m = wp.mv[i];
      if (lane < m.N) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < m.K) w[i][k] = p.wimg[m.w_off + k * m.N + lane];
        if (m.first && m.b_off >= 0) bias[i] = p.wimg[m.b_off + lane];
        if (m.first && m.tmode == TM_SINCOS) { tws[i] = p.wimg[m.tw_off + lane]; twc[i] = p.wimg[m.tw_off + m.N + lane]; }
      }
    }
  }

  // ---- state ----
  float y[R], yprev[R];
  int myslot[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    y[r] = jact ? p.y0[(size_t)grow(r) * H + lane] : 0.f;
    yprev[r] = y[r];
    myslot[r] = p.row_slot ? p.row_slot[grow(r)] : -1;
  }
  auto emit = [&](const snsde_emit em) {
    if (!jact) return;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r >= p.B) continue;
      const float v = em.w_prev * yprev[r] + em.w_curr * y[r];
      if (p.row_slot) {
        if (myslot[r] == em.slot) p.out[(size_t)grow(r) * H + lane] = v;
      } else {
        p.out[((size_t)em.slot * p.B + grow(r)) * H + lane] = v;
      }
    }
  };
  for (int e = 0; e < p.n_init_emits; ++e) {
    snsde_emit em = p.emits[e];
    em.w_prev = 0.f; em.w_curr = 1.f;
    emit(em);
  }

  // spline row of a step: lane c holds (a, b, two_c, three_d) of channel c
  float ca[R], cb[R], cc[R], cd[R];
  auto load_control = [&](int interval) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      ca[r] = cb[r] = cc[r] = cd[r] = 0.f;
      if (pg.uses_control && lane < C) {
        const float* src = p.coeffs + (size_t)grow(r) * p.coeff_row_stride + (size_t)interval * 4 * C + lane;
        ca[r] = __ldg(src); cb[r] = __ldg(src + C); cc[r] = __ldg(src + 2 * C); cd[r] = __ldg(src + 3 * C);
      }
    }
  };
  auto vcoef_at = [&](int s) {
    float v = t.coef_scalar;
    if (jact) {
      if (t.coef_src == CO_IMG) v = p.wimg[t.coef_ref + lane];
      else if (t.coef_src == CO_VBUF) v = p.vtab[(size_t)s * H + lane];
    }
    return v;
  };

  snsde_step st_next = p.S > 0 ? p.steps[0] : snsde_step{};
  float vc_next = 0.f;
  if (p.S > 0) { load_control(st_next.interval); vc_next = vcoef_at(0); }
  float nrm[4];

  for (int s = 0; s < p.S; ++s) {
    const snsde_step st = st_next;
    if (s + 1 < p.S) st_next = p.steps[s + 1];
    snsde_emit em0;
    em0.slot = 0; em0.w_prev = 0.f; em0.w_curr = 0.f;
    if (st.emit_end > st.emit_begin) em0 = p.emits[st.emit_begin];
    const float vcoef = vc_next;

    // X(t) = a + (b + (two_c/2 + three_d*frac/3)*frac)*frac   (torchcde op order), then the next step's row
    float v[kNumRowBufs][R];
#pragma unroll
    for (int b = 0; b < kNumRowBufs; ++b)
#pragma unroll
      for (int r = 0; r < R; ++r) v[b][r] = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float inner = 0.5f * cc[r] + __fdiv_rn(cd[r] * st.frac, 3.0f);
      inner = cb[r] + inner * st.frac;
      v[BUF_X][r] = ca[r] + inner * st.frac;
      v[BUF_Y][r] = y[r];
    }
    if (s + 1 < p.S) { load_control(st_next.interval); vc_next = vcoef_at(s + 1); }

    // ---- the dense program: register mat-vecs ----
    float drift[R], acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { drift[r] = 0.f; acc[r] = 0.f; }
#pragma unroll
    for (int i = 0; i < NMV; ++i) {
      if (i < wp.n_mv) {                                         // warp-uniform
        const WarpMv m = wp.mv[i];
        float a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = pick_buf<R>(v, m.src, r);
        if (m.first) {
          const float init = fmaf(st.cos_t0, twc[i], fmaf(st.sin_t0, tws[i], bias[i]));
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = init;
        }
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 4) {
          if (k0 < m.K) {
#pragma unroll
            for (int k = k0; k < k0 + 4; ++k)
#pragma unroll
              for (int r = 0; r < R; ++r) acc[r] = fmaf(__shfl_sync(0xffffffffu, a[r], k), w[i][k], acc[r]);
          }
        }
        if (m.last) {
          if (m.dst == kWarpDstDrift) {
#pragma unroll
            for (int r = 0; r < R; ++r) drift[r] = acc[r];
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float o = act_apply(acc[r], m.act);
#pragma unroll
              for (int b = 0; b < kNumRowBufs; ++b)
                if (m.dst == b) v[b][r] = o;
            }
          }
        }
      }
    }

    // ---- the update (same arithmetic as snsde_fma.cu, Euler / Milstein with the diagonal closed form) ----
    if (jact) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float d = drift[r];
        if (t.geometric) d = d * tanhf(y[r]);
        if (t.clip_drift) d = tanhf(d);
        const float coef = t.coef_src == CO_RBUF ? pick_buf<R>(v, t.coef_ref, r) : vcoef;
        float g, dgdy;
        diffusion_eval<false>(t, coef, y[r], st.t0, g, dgdy);
        float dw;
        if (p.dW != nullptr) {
          dw = p.dW[((size_t)s * p.B + grow(r)) * H + lane];
        } else {
          const unsigned long long gb = p.row_offset + (unsigned long long)(row0 + r);
          if (r == 0 || (gb & 3ull) == 0ull) philox_normals4(p.seed, (uint32_t)lane, (uint32_t)(gb >> 2), (uint32_t)s, nrm);
          dw = __fmul_rn(pick4(nrm, (int)(gb & 3ull)), st.sqrt_h);
        }
        float yn = __fadd_rn(__fadd_rn(y[r], __fmul_rn(d, st.h)), __fmul_rn(g, dw));
        if (t.milstein) {
          const float q = __fmul_rn(dw, dw) - st.h;
          yn = __fadd_rn(yn, 0.5f * ((g * q) * dgdy));
        }
        yprev[r] = y[r];
        y[r] = yn;
      }
    }
    if (st.emit_end > st.emit_begin) emit(em0);
    for (int e = st.emit_begin + 1; e < st.emit_end; ++e) emit(p.emits[e]);
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
bool warp_plan(const Program& pg, int method, WarpProg& wp) {
  wp.n_mv = 0;
  if (method != SNSDE_METHOD_EULER && method != SNSDE_METHOD_MILSTEIN) return false;      // SRK stages: interpreter kernel
  const TailOp& t = pg.tail;
  if (t.latent) return false;
  if (t.milstein && t.vjp_kind != 0) return false;                                         // full vjp through noise_y
  if (std::max(std::max(pg.H, pg.HH), pg.uses_control ? pg.C : 0) > 32) return false;
  if (t.coef_src == CO_RBUF && (t.coef_ref < 0 || t.coef_ref >= kNumRowBufs)) return false;
  for (int o = 0; o < pg.n_ops; ++o) {
    const DenseOp& op = pg.ops[o];
    if (op.vec) continue;                                   // tabulated per step by vec_tables_kernel
    if (op.src < 0 || op.src >= kNumRowBufs || op.K > 32 || op.N > 32 || op.tmode == TM_RAW) return false;
    if (op.src2 >= kNumRowBufs || (op.src2 >= 0 && op.K2 > 32)) return false;
    if (!op.final_drift && (op.dst < 0 || op.dst >= kNumRowBufs)) return false;
    const int n = op.src2 >= 0 ? 2 : 1;
    if (wp.n_mv + n > kWarpMaxMv) return false;
    for (int part = 0; part < n; ++part) {
      WarpMv& m = wp.mv[wp.n_mv++];
      m.src = part == 0 ? op.src : op.src2;
      m.K = part == 0 ? op.K : op.K2;
      m.N = op.N;
      m.w_off = part == 0 ? op.w_off : op.w2_off;
      m.first = part == 0;
      m.last = part == n - 1;
      m.dst = op.final_drift ? kWarpDstDrift : op.dst;
      m.act = op.act;
      m.tmode = op.tmode;
      m.b_off = op.b_off;
      m.tw_off = op.tw_off;
    }
  }
  return wp.n_mv > 0;
}

template <int NMV, int R>
static cudaError_t warp_launch_one(const FmaParams& p, const WarpProg& wp, int num_sms, cudaStream_t stream) {
  const int n_groups = (p.B + R - 1) / R;
  // small batches: one warp per CTA so that every row group gets an SM of its own; otherwise 4 warps per CTA
  const int wpb = std::max(1, std::min(4, (n_groups + num_sms - 1) / num_sms));
  const int grid = (n_groups + wpb - 1) / wpb;
  snsde_warp_kernel<NMV, R><<<grid, wpb * 32, 0, stream>>>(p, wp);
  return cudaGetLastError();
}

cudaError_t warp_launch(const FmaParams& p, const WarpProg& wp, int num_sms, cudaStream_t stream) {
  // two rows per warp (two independent FMA chains) once there are enough rows to fill the machine with such warps
  const bool two = p.B >= 2 * 4 * num_sms;
  if (wp.n_mv <= 4) return two ? warp_launch_one<4, 2>(p, wp, num_sms, stream) : warp_launch_one<4, 1>(p, wp, num_sms, stream);
  return warp_launch_one<kWarpMaxMv, 1>(p, wp, num_sms, stream);
}

}  // namespace snsde
