// fp32 FMA kernels: the whole fixed-step SDE solve in ONE launch, for every model / shape / method.
//
// Replaces the Python step loop of torchsde.sdeint (BaseSDESolver.integrate + Euler / Milstein / SRK .step)
// together with the per-step Diffusion_model.f/g evaluation
// (/root/reference/benchmark_classification/models_sde/neuralsde.py:295-307) and
// torchcde.CubicSpline.evaluate (:296).
//
// Work decomposition (the north star's "warp-shuffle/FMA path" for hidden < 64, and the general fallback):
//   * a ROW GROUP = R batch rows integrated by `nw` warps (thread j of the group = feature j).  A group never
//     talks to another group: its hand-offs are __syncwarp (nw == 1, hidden <= 32) or one named barrier per
//     group (bar.sync id, 32*nw) - there is no __syncthreads in the time loop.  A CTA packs several groups that
//     share one shared-memory copy of the weights;
//   * R is 4 or 8 and groups start at multiples of 4 rows, so ONE Philox4x32 call per (feature, step) yields the
//     normals of 4 rows (the stream is keyed by global_row >> 2);
//   * the row-independent noise networks (noise_t(sin t, cos t), the tutorial's g_net(noise_in(t))) are tabulated
//     per evaluation time by a pre-kernel (vec_tables_kernel) instead of being evaluated on B identical rows;
//   * per step: cp.async-prefetched spline rows -> X(t); the dense-op program -> drift pre-activation in
//     registers; Philox or table increments; the state update; emits (lerp / fused final_index capture).
//   * METHOD 1 = SRK (torchsde SRK.diagonal_or_scalar_step, SRID2 tableau): 3 drift and 4 diffusion evaluations
//     per step at the stage states H0_s / H1_s.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>

#include "snsde_fma.cuh"

namespace snsde {

// ---- row-independent coefficient table ------------------------------------------------------------
// One CTA per (step, evaluation time): runs the program's vec ops on that single "row" and stores the final
// coefficient vector.  The reference evaluates these networks on B identical rows (neuralsde.py:281-282).
__global__ void __launch_bounds__(1024) vec_tables_kernel(const Program pg, const float* __restrict__ wimg,
                                                          const snsde_step* __restrict__ steps,
                                                          const snsde_point* __restrict__ points, int npg,
                                                          float* __restrict__ vtab) {
  extern __shared__ __align__(16) float vsm[];          // [kNumVecBufs][ld]
  const int s = blockIdx.x / npg, q = blockIdx.x - s * npg, j = threadIdx.x, ld = pg.ld;
  TimePoint tp;
  if (points != nullptr) {
    const snsde_point pt = points[s * kSrkPoints + (q == 0 ? 0 : (q == 1 ? 1 : 3))];
    tp.t = pt.t; tp.sin_t = pt.sin_t; tp.cos_t = pt.cos_t;
  } else {
    const snsde_step st = steps[s];
    tp.t = st.t0; tp.sin_t = st.sin_t0; tp.cos_t = st.cos_t0;
  }
  for (int o = 0; o < pg.n_ops; ++o) {
    const DenseOp& op = pg.ops[o];
    if (!op.vec) continue;
    if (j < op.N) {
      float v = op.b_off >= 0 ? wimg[op.b_off + j] : 0.f;
      if (op.tmode == TM_SINCOS) v = fmaf(tp.cos_t, wimg[op.tw_off + op.N + j], fmaf(tp.sin_t, wimg[op.tw_off + j], v));
      else if (op.tmode == TM_RAW) v = fmaf(tp.t, wimg[op.tw_off + j], v);
      if (op.src >= 0) {
        const float* src = vsm + (op.src - BUF_V0) * ld;
        const float* w = wimg + op.w_off + j;
        for (int k = 0; k < op.K; ++k) v = fmaf(src[k], w[(size_t)k * op.N], v);
      }
      vsm[(op.dst - BUF_V0) * ld + j] = act_apply(v, op.act);
    }
    __syncthreads();
  }
  if (j < pg.H) vtab[(size_t)blockIdx.x * pg.H + j] = vsm[(pg.tail.coef_ref - BUF_V0) * ld + j];
}

// ---- the solve ---------------------------------------------------------------------------------------
// WS: the whole weight image is staged in shared memory (true) or read through L1/L2 (false: it does not fit).
template <int R, int NTMAX, int METHOD, bool WS>
__global__ void __launch_bounds__(NTMAX) snsde_fma_kernel(const FmaParams p) {
  extern __shared__ __align__(16) float smem[];
  const Program& pg = p.prog;
  const TailOp& t = pg.tail;
  const int H = pg.H, C = pg.C, ld = pg.ld;
  const int nw = p.nw, GT = nw * 32;                        // threads per group
  const int gid = threadIdx.x / GT, tid = threadIdx.x - gid * GT;
  constexpr int NP = METHOD == 1 ? 3 : 1;                   // drift evaluation points per step
  constexpr int NPG = METHOD == 1 ? kSrkGPoints : 1;

  // ---- shared memory: [weights][group 0: row bufs, spline stages][group 1 ...] ----
  const int stage_floats = pg.uses_control ? R * 4 * C : 0;
  const int group_floats = kNumRowBufs * R * ld + 2 * NP * stage_floats;
  if (WS) {
    stage_weights(smem, p.wimg, p.smem_w_floats);
    __syncthreads();                                        // the only CTA-wide barrier
  }
  const float* __restrict__ const W = WS ? smem : p.wimg;
  GroupSmem sm;
  sm.base = smem + (WS ? p.smem_w_floats : 0) + gid * group_floats;
  sm.row_buf_floats = R * ld;
  sm.stage0 = sm.base + kNumRowBufs * R * ld;
  sm.stage_floats = stage_floats;

  const int row0 = (blockIdx.x * p.groups + gid) * R;
  if (row0 >= p.B) return;                                  // whole group idle (groups never sync with each other)
  float* const sY = sm.buf(BUF_Y);
  float* const sX = sm.buf(BUF_X);
  auto grow = [&](int r) { return min(row0 + r, p.B - 1); };   // clamped row (local to this shard)
  auto gsync = [&]() { group_sync(gid, nw); };

  // which interval a drift evaluation point reads: Euler/Milstein -> the step's own; SRK -> points 0, 3, 2
  auto f_point = [&](int s, int i, TimePoint& tp, int& interval, float& frac) {
    if (METHOD == 1) {
      const snsde_point pt = p.points[s * kSrkPoints + (i == 0 ? 0 : (i == 1 ? 3 : 2))];
      tp.t = pt.t; tp.sin_t = pt.sin_t; tp.cos_t = pt.cos_t; interval = pt.interval; frac = pt.frac;
    } else {
      const snsde_step st = p.steps[s];
      tp.t = st.t0; tp.sin_t = st.sin_t0; tp.cos_t = st.cos_t0; interval = st.interval; frac = st.frac;
    }
  };
  auto stage_buf = [&](int s, int i) { return sm.stage0 + (((s & 1) * NP) + i) * stage_floats; };
  auto prefetch_spline = [&](int s) {
    if (!pg.uses_control || s >= p.S) return;
    for (int i = 0; i < NP; ++i) {
      TimePoint tp; int interval; float frac;
      f_point(s, i, tp, interval, frac);
      float* dst = stage_buf(s, i);
      for (int q = tid; q < R * C; q += GT) {
        const int r = q / C, c = q - r * C;
        const float* src = p.coeffs + (size_t)grow(r) * p.coeff_row_stride + (size_t)interval * 4 * C + 4 * c;
        cp_async16(dst + r * 4 * C + 4 * c, src);
      }
    }
    cp_async_commit();
  };
  prefetch_spline(0);

  // X(t) of drift point i of step s: a + (b + (two_c/2 + three_d*frac/3)*frac)*frac  (torchcde op order)
  auto eval_control = [&](int s, int i, float frac) {
    const float* stg = stage_buf(s, i);
    for (int q = tid; q < R * C; q += GT) {
      const int r = q / C, c = q - r * C;
      const float* row = stg + r * 4 * C;
      float inner = 0.5f * row[2 * C + c] + __fdiv_rn(row[3 * C + c] * frac, 3.0f);
      inner = row[C + c] + inner * frac;
      sX[r * ld + c] = row[c] + inner * frac;
    }
  };

  // ---- state in registers: thread j owns y[.][j] ----
  const bool jact = tid < H;
  float y[R], yprev[R];
  int myslot[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    y[r] = jact ? p.y0[(size_t)grow(r) * H + tid] : 0.f;
    yprev[r] = y[r];
    myslot[r] = p.row_slot ? p.row_slot[grow(r)] : -1;
    if (jact) sY[r * ld + tid] = y[r];
  }

  auto emit = [&](const snsde_emit em) {
    if (!jact) return;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r >= p.B) continue;
      const float v = em.w_prev * yprev[r] + em.w_curr * y[r];
      if (p.row_slot) {
        if (myslot[r] == em.slot) p.out[(size_t)grow(r) * H + tid] = v;
      } else {
        p.out[((size_t)em.slot * p.B + grow(r)) * H + tid] = v;
      }
    }
  };
  for (int e = 0; e < p.n_init_emits; ++e) {
    snsde_emit em = p.emits[e];
    em.w_prev = 0.f; em.w_curr = 1.f;
    emit(em);
  }

  // Runs the per-row ops of one part (0 drift, 1 diffusion, 2 both) on the state currently in sY (and sX);
  // the drift pre-activation stays in `acc`.  Every op is preceded by a group hand-off.
  float acc[R];
  auto run_ops = [&](int part, const TimePoint& tp) {
    for (int o = 0; o < pg.n_ops; ++o) {
      const DenseOp& op = pg.ops[o];
      if (op.vec || (part != 2 && op.part != part)) continue;
      gsync();                              // op o reads what an earlier op (or the state writer) wrote
      if (tid < op.N) {
        if (op.final_drift) {
          dense_eval<R>(acc, op, W, sm, ld, tp, tid);
        } else {
          float a[R];
          dense_eval<R>(a, op, W, sm, ld, tp, tid);
          float* dst = sm.buf(op.dst);
#pragma unroll
          for (int r = 0; r < R; ++r) dst[r * ld + tid] = a[r];
        }
      }
    }
    if (part != 0 && t.coef_src == CO_RBUF) gsync();       // the tail reads the coefficient buffer
  };
  auto drift_value = [&](float d, float ycur) {
    if (t.geometric) d = d * tanhf(ycur);
    if (t.clip_drift) d = tanhf(d);
    return d;
  };
  // LatentSDE.f_aug (latent_sde.py:77-82): the drift of the last state channel is 0.5 * sum_j u_j^2 over the latent
  // features, u = (f - theta (mu - y)) / stable(sigma), evaluated at the state the dense program just read (sY).
  // Whole-group call (shuffles + at most two hand-offs); the partial sums of the warps meet in the BUF_U rows.
  auto latent_drift = [&](float (&f)[R]) {
    if (!t.latent) return;
    float part[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float u = 0.f;
      if (tid < H - 1) u = __fdiv_rn(f[r] - t.lat_theta * (t.lat_mu - sY[r * ld + tid]), t.lat_div);
      part[r] = u * u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part[r] += __shfl_xor_sync(0xffffffffu, part[r], o);
    }
    if (nw > 1) {
      float* const red = sm.buf(BUF_U);                        // (its previous readers are behind the dense program's hand-offs)
      if ((tid & 31) == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) red[r * ld + (tid >> 5)] = part[r];
      }
      gsync();
      if (tid == H - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float v = 0.f;
          for (int q = 0; q < nw; ++q) v += red[r * ld + q];
          part[r] = v;
        }
      }
    }
    if (tid == H - 1) {
#pragma unroll
      for (int r = 0; r < R; ++r) f[r] = 0.5f * part[r];
    }
  };
  // coefficient entering the elementwise diffusion for row r (table / image / scalar / per-row network output)
  auto coef_of = [&](int r, float vcoef) {
    return t.coef_src == CO_RBUF ? sm.buf(t.coef_ref)[r * ld + tid] : vcoef;
  };
  auto vcoef_at = [&](int s, int q) {
    float v = t.coef_scalar;
    if (t.coef_src == CO_IMG) v = W[t.coef_ref + tid];
    else if (t.coef_src == CO_VBUF) v = p.vtab[((size_t)s * NPG + q) * H + tid];
    if (t.latent && tid == H - 1) v = 0.f;                    // g_aug: no noise on the KL accumulator (latent_sde.py:84-90)
    return v;
  };
  // Brownian increment (and, for SRK, the space-time Levy integral U) of row r at step s
  float nrm[4], nrmu[4];
  auto draw = [&](int s, int r, float h, float sqrt_h, float& w, float& u) {
    if (p.dW != nullptr) {
      w = p.dW[((size_t)s * p.B + grow(r)) * H + tid];
      u = (METHOD == 1) ? p.dU[((size_t)s * p.B + grow(r)) * H + tid] : 0.f;
      return;
    }
    const unsigned long long gb = p.row_offset + (unsigned long long)(row0 + r);
    if (r == 0 || (gb & 3ull) == 0ull) {
      philox_normals4(p.seed, (uint32_t)tid, (uint32_t)(gb >> 2), (uint32_t)s, nrm);
      if (METHOD == 1) philox_normals4_u(p.seed, (uint32_t)tid, (uint32_t)(gb >> 2), (uint32_t)s, nrmu);
    }
    w = __fmul_rn(pick4(nrm, (int)(gb & 3ull)), sqrt_h);
    u = (METHOD == 1) ? levy_U(w, pick4(nrmu, (int)(gb & 3ull)), h, sqrt_h) : 0.f;
  };

  snsde_step st_next = p.S > 0 ? p.steps[0] : snsde_step{};
  for (int s = 0; s < p.S; ++s) {
    // step record of the NEXT step and this step's first emit are requested now: their global-memory latency is
    // hidden behind the step instead of opening and closing it
    const snsde_step st = st_next;
    if (s + 1 < p.S) st_next = p.steps[s + 1];
    snsde_emit em0;
    em0.slot = 0; em0.w_prev = 0.f; em0.w_curr = 0.f;
    if (st.emit_end > st.emit_begin) em0 = p.emits[st.emit_begin];

    if (METHOD == 0) {
      // =============================== Euler / Milstein ===============================================
      TimePoint tp{st.t0, st.sin_t0, st.cos_t0};
      float vcoef = 0.f;
      if (jact) vcoef = vcoef_at(s, 0);                       // issued early: consumed after the dense program
      if (pg.uses_control) {
        cp_async_wait_all();
        gsync();
        eval_control(s, 0, st.frac);
        prefetch_spline(s + 1);
      }
      run_ops(2, tp);
      latent_drift(acc);

      const bool net_vjp = t.milstein && t.vjp_kind != 0;       // block-uniform
      float* const sA = sm.buf(BUF_X);                            // scratch of the vjp (dead since the first ops)
      float* const sB = sm.buf(BUF_U);
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float d = drift_value(acc[r], y[r]);
          const float coef = coef_of(r, vcoef);
          float g, dgdy;
          diffusion_eval<false>(t, coef, y[r], st.t0, g, dgdy);
          float w, u;
          draw(s, r, st.h, st.sqrt_h, w, u);
          float yn = __fadd_rn(__fadd_rn(y[r], __fmul_rn(d, st.h)), __fmul_rn(g, w));
          if (t.milstein) {
            const float v = __fmul_rn(w, w) - st.h;
            yn = __fadd_rn(yn, 0.5f * ((g * v) * dgdy));          // direct (diagonal) part of d g_j / d y_j
            if (net_vjp) {
              // cotangent reaching q_j:  (g v) * tanh' * s_theta * [nan_to_num passes] * d raw / d q  (* relu' for kind 2)
              const float raw = (t.mult == MU_Y) ? coef * y[r] : coef;
              float a = (g * v) * ((1.f - g * g) * t.s_theta) * (is_finite_f(raw) ? 1.f : 0.f);
              if (t.mult == MU_Y) a *= y[r];
              if (t.vjp_kind == 2 && !(coef > 0.f)) a = 0.f;
              sA[r * ld + tid] = a;
            }
          }
          yprev[r] = y[r];
          y[r] = yn;
          if (!net_vjp) sY[r * ld + tid] = yn;
        }
      }
      if (net_vjp) {
        // torchsde: + 0.5 * vjp_y(g; g * (dW^2 - h)); the path through noise_y is W1y^T [relu'] W2^T [relu'] a
        gsync();
        const float* src = sA;
        if (t.vjp_kind == 2) {
          if (jact) {
            float b[R];
#pragma unroll
            for (int r = 0; r < R; ++r) b[r] = 0.f;
            dot_rows_T<R>(b, sA, ld, p.wimg + t.vjp_w2 + (size_t)tid * H, H);
            const float* h1 = sm.buf(t.vjp_h1);
#pragma unroll
            for (int r = 0; r < R; ++r) sB[r * ld + tid] = (h1[r * ld + tid] > 0.f) ? b[r] : 0.f;
          }
          gsync();
          src = sB;
        }
        if (jact) {
          float c[R];
#pragma unroll
          for (int r = 0; r < R; ++r) c[r] = 0.f;
          dot_rows_T<R>(c, src, ld, p.wimg + t.vjp_w1 + (size_t)tid * H, H);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            y[r] = __fadd_rn(y[r], 0.5f * c[r]);
            sY[r * ld + tid] = y[r];
          }
        }
      }
    } else {
      // =============================== SRK (SRID2, diagonal noise) ====================================
      // torchsde SRK.diagonal_or_scalar_step; tableau (tableaus/srid2.py, Roessler 2010 SRI2):
      //   C0 = (0, 1, 1/2, 0)   A0 = [[], [1], [1/4, 1/4], [0, 0, 0]]      B0 = [[], [0], [1, 1/2], [0, 0, 0]]
      //   C1 = (0, 1/4, 1, 1/4) A1 = [[], [1/4], [1, 0], [0, 0, 1/4]]      B1 = [[], [-1/2], [1, 0], [2, -1, 1/2]]
      //   alpha = (1/6, 1/6, 2/3, 0)   beta1 = (-1, 4/3, 2/3, 0)   beta2 = (1, -4/3, 1/3, 0)
      //   beta3 = (2, -4/3, -2/3, 0)   beta4 = (-2, 5/3, -2/3, 1)
      // Stage 3 of the drift (H0_3 = y0, alpha_3 = 0) contributes nothing and is skipped.
      const float h = st.h, sqrt_h = st.sqrt_h, rdt = __fdiv_rn(1.0f, h);
      const snsde_point* pts = p.points + s * kSrkPoints;
      TimePoint tp0{pts[0].t, pts[0].sin_t, pts[0].cos_t}, tpq{pts[1].t, pts[1].sin_t, pts[1].cos_t};
      TimePoint tph{pts[2].t, pts[2].sin_t, pts[2].cos_t}, tp1{pts[3].t, pts[3].sin_t, pts[3].cos_t};
      float vc0 = 0.f, vcq = 0.f, vc1 = 0.f;
      if (jact) { vc0 = vcoef_at(s, 0); vcq = vcoef_at(s, 1); vc1 = vcoef_at(s, 2); }
      float y0r[R], w[R], ik0[R], f0[R], g0[R], f1[R], g1[R], f2[R], g2[R], tmp[R];
      const bool per_row_g = t.coef_src == CO_RBUF;
      if (pg.uses_control) { cp_async_wait_all(); gsync(); eval_control(s, 0, pts[0].frac); }
      // stage 0: f0 = f(t0, y0), g0 = g(t0, y0)            (sY holds y0)
      run_ops(2, tp0);
      latent_drift(acc);
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          y0r[r] = y[r];
          float u, dg;
          draw(s, r, h, sqrt_h, w[r], u);
          ik0[r] = u;
          f0[r] = drift_value(acc[r], y[r]);
          diffusion_eval<false>(t, coef_of(r, vc0), y[r], tp0.t, g0[r], dg);
        }
      }
      // stage 1: H0_1 = y0 + f0 h (+ 0 g0 I_k0/h);  H1_1 = y0 + f0 h/4 - g0 sqrt(h)/2
      gsync();
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          sY[r * ld + tid] = y0r[r] + (1.0f * f0[r]) * h + ((0.0f * g0[r]) * ik0[r]) * rdt;
          tmp[r] = y0r[r] + (0.25f * f0[r]) * h + (-0.5f * g0[r]) * sqrt_h;
        }
      }
      if (pg.uses_control) { gsync(); eval_control(s, 1, pts[3].frac); }
      run_ops(0, tp1);                                       // f1 = f(t0 + h, H0_1)
      latent_drift(acc);
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) f1[r] = drift_value(acc[r], sY[r * ld + tid]);
      }
      gsync();
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) sY[r * ld + tid] = tmp[r];
      }
      if (per_row_g) run_ops(1, tpq);                        // g1 = g(t0 + h/4, H1_1)
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) { float dg; diffusion_eval<false>(t, coef_of(r, vcq), tmp[r], tpq.t, g1[r], dg); }
      }
      // stage 2: H0_2 = y0 + f0 h/4 + g0 I_k0/h + f1 h/4 + g1 I_k0/(2h);  H1_2 = y0 + f0 h + g0 sqrt(h) (+ 0 f1 h + 0 g1 sqrt(h))
      gsync();
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float a = y0r[r] + (0.25f * f0[r]) * h + ((1.0f * g0[r]) * ik0[r]) * rdt;
          a = a + (0.25f * f1[r]) * h + ((0.5f * g1[r]) * ik0[r]) * rdt;
          sY[r * ld + tid] = a;
          float b = y0r[r] + (1.0f * f0[r]) * h + (1.0f * g0[r]) * sqrt_h;
          b = b + (0.0f * f1[r]) * h + (0.0f * g1[r]) * sqrt_h;
          tmp[r] = b;
        }
      }
      if (pg.uses_control) { gsync(); eval_control(s, 2, pts[2].frac); prefetch_spline(s + 1); }
      run_ops(0, tph);                                       // f2 = f(t0 + h/2, H0_2)
      latent_drift(acc);
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) f2[r] = drift_value(acc[r], sY[r * ld + tid]);
      }
      gsync();
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) sY[r * ld + tid] = tmp[r];
      }
      if (per_row_g) run_ops(1, tp1);                        // g2 = g(t0 + h, H1_2)
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) { float dg; diffusion_eval<false>(t, coef_of(r, vc1), tmp[r], tp1.t, g2[r], dg); }
      }
      // stage 3: H1_3 = y0 (+0 f0 h) + 2 g0 sqrt(h) (+0 f1 h) - g1 sqrt(h) + f2 h/4 + g2 sqrt(h)/2
      gsync();
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float b = y0r[r] + (0.0f * f0[r]) * h + (2.0f * g0[r]) * sqrt_h;
          b = b + (0.0f * f1[r]) * h + (-1.0f * g1[r]) * sqrt_h;
          b = b + (0.25f * f2[r]) * h + (0.5f * g2[r]) * sqrt_h;
          tmp[r] = b;
          sY[r * ld + tid] = b;
        }
      }
      if (per_row_g) run_ops(1, tpq);                        // g3 = g(t0 + h/4, H1_3)
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float g3, dg;
          diffusion_eval<false>(t, coef_of(r, vcq), tmp[r], tpq.t, g3, dg);
          const float I_k = w[r];
          const float I_kk = (I_k * I_k - h) * 0.5f;
          const float I_kkk = (I_k * I_k * I_k - 3.0f * h * I_k) * (1.0f / 6.0f);
          const float a0 = I_kk / sqrt_h, a1 = ik0[r] * rdt, a2 = I_kkk * rdt;
          const float gw0 = -1.0f * I_k + 1.0f * a0 + 2.0f * a1 + -2.0f * a2;
          const float gw1 = (4.0f / 3.0f) * I_k + (-4.0f / 3.0f) * a0 + (-4.0f / 3.0f) * a1 + (5.0f / 3.0f) * a2;
          const float gw2 = (2.0f / 3.0f) * I_k + (1.0f / 3.0f) * a0 + (-2.0f / 3.0f) * a1 + (-2.0f / 3.0f) * a2;
          const float gw3 = a2;
          float yn = y0r[r] + ((1.0f / 6.0f) * f0[r]) * h + gw0 * g0[r];
          yn = yn + ((1.0f / 6.0f) * f1[r]) * h + gw1 * g1[r];
          yn = yn + ((2.0f / 3.0f) * f2[r]) * h + gw2 * g2[r];
          yn = yn + gw3 * g3;
          yprev[r] = y0r[r];
          y[r] = yn;
        }
      }
      if (jact) {                                            // (every reader of sY is behind a hand-off by now)
#pragma unroll
        for (int r = 0; r < R; ++r) sY[r * ld + tid] = y[r];
      }
    }
    if (st.emit_end > st.emit_begin) emit(em0);
    for (int e = st.emit_begin + 1; e < st.emit_end; ++e) emit(p.emits[e]);
  }
}

// ---- host-side launchers -----------------------------------------------------------------------

size_t fma_group_smem_floats(const Program& pg, int R, int method) {
  const int np = method == SNSDE_METHOD_SRK ? 3 : 1;
  size_t f = (size_t)kNumRowBufs * R * pg.ld;
  if (pg.uses_control) f += (size_t)2 * np * R * 4 * pg.C;
  return f;
}

template <int R, int NTMAX, int METHOD, bool WS>
static cudaError_t launch_one(const FmaParams& p, int grid, int nt, size_t smem, cudaStream_t stream) {
  auto kern = snsde_fma_kernel<R, NTMAX, METHOD, WS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, nt, smem, stream>>>(p);
  return cudaGetLastError();
}

template <int R, int NTMAX>
static cudaError_t launch_rn(const FmaParams& p, int method, int grid, int nt, size_t smem, cudaStream_t stream) {
  const bool ws = p.smem_w_floats > 0;
  if (method == SNSDE_METHOD_SRK)
    return ws ? launch_one<R, NTMAX, 1, true>(p, grid, nt, smem, stream) : launch_one<R, NTMAX, 1, false>(p, grid, nt, smem, stream);
  return ws ? launch_one<R, NTMAX, 0, true>(p, grid, nt, smem, stream) : launch_one<R, NTMAX, 0, false>(p, grid, nt, smem, stream);
}

cudaError_t fma_launch(const FmaParams& p, int R, int method, size_t smem, cudaStream_t stream) {
  const int nt = p.groups * p.nw * 32;
  const int n_groups = (p.B + R - 1) / R;
  const int grid = (n_groups + p.groups - 1) / p.groups;
  if (nt <= 512) {
    if (R == 8 && method != SNSDE_METHOD_SRK) return launch_rn<8, 512>(p, method, grid, nt, smem, stream);
    if (R == 4) return launch_rn<4, 512>(p, method, grid, nt, smem, stream);
    if (R == 1) return launch_rn<1, 512>(p, method, grid, nt, smem, stream);
  } else if (R == 4) {
    return launch_rn<4, 1024>(p, method, grid, nt, smem, stream);
  }
  return cudaErrorInvalidValue;
}

cudaError_t vec_tables_launch(const Program& pg, const float* wimg, const snsde_step* steps, const snsde_point* points,
                              int S, int npg, float* vtab, cudaStream_t stream) {
  const int nt = std::max(32, (std::max(pg.H, pg.HH) + 31) & ~31);
  vec_tables_kernel<<<S * npg, nt, sizeof(float) * kNumVecBufs * pg.ld, stream>>>(pg, wimg, steps, points, npg, vtab);
  return cudaGetLastError();
}

}  // namespace snsde
