// Reverse sweep of the SRK solve (torchsde SRK.diagonal_or_scalar_step, SRID2 tableau): dL/dy0 and the per-op cotangents
// behind the weight gradients, for method 'srk' - the default of the torch-ists NeuralSDE wrapper
// (/root/reference/torch-ists/torch_ists/diff_module/NSDE/nsde_model.py:63-74) and of LatentSDE
// (torch-ists/torch_ists/diff_module/NSDE/latent_sde.py:107-109), which the reference trains by autograd through
// torchsde's step loop (LatentSDE: sdeint_adjoint; here the exact reverse of the discrete solve).
//
// One SRK step evaluates the drift at three states and the diffusion at four (snsde_fma.cu spells the tableau out):
//   f0 = f(t0, y0)          g0 = g(t0, y0)
//   H01 = y0 + f0 h                                   f1 = f(t0+h,   H01)
//   H11 = y0 + f0 h/4 - g0 sqrt(h)/2                  g1 = g(t0+h/4, H11)
//   H02 = y0 + f0 h/4 + g0 U/h + f1 h/4 + g1 U/(2h)   f2 = f(t0+h/2, H02)
//   H12 = y0 + f0 h + g0 sqrt(h)                      g2 = g(t0+h,   H12)
//   H13 = y0 + 2 g0 sqrt(h) - g1 sqrt(h) + f2 h/4 + g2 sqrt(h)/2        g3 = g(t0+h/4, H13)
//   y1  = y0 + h (f0/6 + f1/6 + 2 f2/3) + gw0 g0 + gw1 g1 + gw2 g2 + gw3 g3
// The reverse of a step first recomputes the stage states from y_s (saved by the forward solve), then visits the six
// evaluation SITES in the order g3, g2, f2, g1, f1, (f0, g0) - every site after all of its consumers.  A site re-evaluates
// its part of the dense program at its stage state (activations in shared memory, and in HBM where a weight gradient
// needs them), differentiates the drift / diffusion tail, runs the part's ops in reverse and returns J^T cot, which is
// scattered to the cotangents of the quantities the stage state was built from.
//
// Same decomposition as snsde_bwd.cu (row groups of R rows x nw warps, thread j = feature j, group-local hand-offs).
// Per op the D / P buffers carry one block of S*B rows per site of the op's part (drift: f0, f1, f2; diffusion: g0..g3),
// so the weight gradients stay single GEMMs  dW = D^T P  over all sites (snsde_api.cu).
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>

#include "snsde_bwd.cuh"
#include "snsde_fma.cuh"

namespace snsde {

template <int R, int NTMAX, bool WS>
__global__ void __launch_bounds__(NTMAX) snsde_bwd_srk_kernel(const BwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const Program& pg = p.prog;
  const TailOp& t = pg.tail;
  const int H = pg.H, C = pg.C, ld = pg.ld;
  const int nw = p.nw, GT = nw * 32;
  const int gid = threadIdx.x / GT, tid = threadIdx.x - gid * GT;
  const int n_rops = p.n_rops;
  const int slot = R * ld;

  // ---- shared memory: [weights][group: Y X D | post[n_rops] | cot[n_rops] | pre[n_rops]?] ----
  const int group_floats = (3 + n_rops * (p.has_lipswish ? 3 : 2)) * slot;
  if (WS) {
    stage_weights(smem, p.wimg, p.smem_w_floats);
    __syncthreads();
  }
  const float* __restrict__ const W = WS ? smem : p.wimg;
  float* const gbase = smem + (WS ? p.smem_w_floats : 0) + gid * group_floats;
  float* const sY = gbase;
  float* const sX = gbase + slot;
  float* const sD = gbase + 2 * slot;
  float* const sPost = gbase + 3 * slot;
  float* const sCot = sPost + n_rops * slot;
  float* const sPre = sCot + n_rops * slot;                  // only when has_lipswish

  const int row0 = (blockIdx.x * p.groups + gid) * R;
  if (row0 >= p.B) return;
  auto grow = [&](int r) { return min(row0 + r, p.B - 1); };
  auto gsync = [&]() { group_sync(gid, nw); };
  auto buf_of = [&](int src_op) -> const float* {
    if (src_op == SRC_STATE) return sY;
    if (src_op == SRC_CONTROL) return sX;
    for (int i = 0; i < n_rops; ++i) if (p.rop[i] == src_op) return sPost + i * slot;
    return sY;
  };
  auto rop_index = [&](int op_index) { for (int i = 0; i < n_rops; ++i) if (p.rop[i] == op_index) return i; return -1; };

  const bool jact = tid < H;
  const size_t BH = (size_t)p.B * H, SB = (size_t)p.S * p.B;
  float lam[R];
  bool valid[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    valid[r] = row0 + r < p.B;
    lam[r] = (jact && valid[r]) ? p.grad_states[(size_t)p.S * BH + (size_t)grow(r) * H + tid] : 0.f;
  }
  float acc_theta = 0.f, acc_coef = 0.f;
  const int coef_i = t.coef_src == CO_RBUF ? rop_index(t.coef_op) : -1;
  const bool per_row_g = coef_i >= 0;
  int final_i = -1;
  for (int i = 0; i < n_rops; ++i) if (pg.ops[p.rop[i]].final_drift) final_i = i;

  // ---- building blocks ------------------------------------------------------------------------------------------
  // the state a site evaluates at -> sY (all rows of the group), X(t) of the site -> sX
  auto put_state = [&](const float (&z)[R]) {
    gsync();                                                  // earlier readers of sY are done
    if (jact) {
#pragma unroll
      for (int r = 0; r < R; ++r) sY[r * ld + tid] = z[r];
    }
  };
  auto put_control = [&](const snsde_point& pt, size_t site_rows) {      // site_rows: first row of the site in xbuf, or ~0
    if (!pg.uses_control) return;
    for (int q = tid; q < R * C; q += GT) {
      const int r = q / C, c = q - r * C;
      const float* row = p.coeffs + (size_t)grow(r) * p.coeff_row_stride + (size_t)pt.interval * 4 * C;
      float inner = 0.5f * row[2 * C + c] + __fdiv_rn(row[3 * C + c] * pt.frac, 3.0f);
      inner = row[C + c] + inner * pt.frac;
      const float x = row[c] + inner * pt.frac;
      sX[r * ld + c] = x;
      if (site_rows != ~(size_t)0 && p.xbuf != nullptr && row0 + r < p.B) p.xbuf[(site_rows + row0 + r) * C + c] = x;
    }
  };
  // forward evaluation of the ops of `part` (0 drift, 1 diffusion) on (sY, sX); outputs stay in sPost, the final drift
  // pre-activation in acc; with store_rows != ~0 the activations a weight gradient needs go to HBM at that site's rows
  auto eval_part = [&](int part, const TimePoint& tp, float (&acc)[R], size_t store_rows) {
    for (int i = 0; i < n_rops; ++i) {
      const DenseOp& op = pg.ops[p.rop[i]];
      if (op.part != part) continue;
      gsync();
      if (tid < op.N) {
        float a[R];
        float init = op.b_off >= 0 ? W[op.b_off + tid] : 0.f;
        if (op.tmode == TM_SINCOS) {
          const float* tw = W + op.tw_off;
          init = fmaf(tp.cos_t, tw[op.N + tid], fmaf(tp.sin_t, tw[tid], init));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = init;
        if (op.src >= 0) dot_accumulate<R>(a, buf_of(op.src_op), ld, W + op.w_off + tid, op.K, op.N);
        if (op.src2 >= 0) dot_accumulate<R>(a, buf_of(op.src2_op), ld, W + op.w2_off + tid, op.K2, op.N);
        if (op.final_drift) {
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = a[r];
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float post = act_apply(a[r], op.act);
            sPost[i * slot + r * ld + tid] = post;
            if (p.has_lipswish) sPre[i * slot + r * ld + tid] = a[r];
            if (store_rows != ~(size_t)0 && p.pbuf[p.rop[i]] != nullptr && valid[r])
              p.pbuf[p.rop[i]][(store_rows + row0 + r) * op.N + tid] = post;
          }
        }
      }
    }
    gsync();                                                  // the tail may read another thread's output (coefficient row)
  };
  // the ops of `part` in reverse; cotangents of their outputs are in sCot; D goes to HBM at the site's rows; the
  // contributions to the state cotangent are added to a_y
  auto reverse_part = [&](int part, size_t site_rows, float (&a_y)[R]) {
    for (int i = n_rops - 1; i >= 0; --i) {
      const DenseOp& op = pg.ops[p.rop[i]];
      if (op.part != part) continue;
      gsync();
      if (tid < op.N) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float cot = sCot[i * slot + r * ld + tid];
          const float dlt = op.final_drift ? cot
                                           : cot * act_grad(p.has_lipswish ? sPre[i * slot + r * ld + tid] : 0.f,
                                                            sPost[i * slot + r * ld + tid], op.act);
          sD[r * ld + tid] = dlt;
          if (valid[r]) p.dbuf[p.rop[i]][(site_rows + row0 + r) * op.N + tid] = dlt;
        }
      }
      gsync();
      for (int half = 0; half < 2; ++half) {
        const int src_op = half == 0 ? op.src_op : op.src2_op;
        const int K = half == 0 ? op.K : op.K2;
        const int col = half == 0 ? op.g_col : op.g_col2;
        if ((half == 0 ? op.src : op.src2) < 0 || src_op == SRC_CONTROL || src_op == SRC_ABSENT) continue;
        if (tid < K) {
          float c[R];
#pragma unroll
          for (int r = 0; r < R; ++r) c[r] = 0.f;
          dot_accumulate<R>(c, sD, ld, p.blob + op.g_w + col + tid, op.N, op.g_ldw);
          if (src_op == SRC_STATE) {
#pragma unroll
            for (int r = 0; r < R; ++r) a_y[r] += c[r];
          } else {
            const int si = rop_index(src_op);
#pragma unroll
            for (int r = 0; r < R; ++r) sCot[si * slot + r * ld + tid] = c[r];
          }
        }
      }
    }
    gsync();
  };
  auto time_of = [&](const snsde_point& pt) { return TimePoint{pt.t, pt.sin_t, pt.cos_t}; };
  // drift value from the pre-activation (geometric term, tanh clip); for LatentSDE also the KL channel
  // (feature H-1 = 0.5 sum_j u_j^2, u = (f - theta (mu - z)) / div).  Whole-group call.
  auto drift_of = [&](const float (&acc)[R], const float (&z)[R], float (&f)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float d = acc[r];
      if (t.geometric) d = d * tanhf(z[r]);
      if (t.clip_drift) d = tanhf(d);
      f[r] = d;
    }
    if (t.latent) {
      float part[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float u = 0.f;
        if (tid < H - 1) u = __fdiv_rn(f[r] - t.lat_theta * (t.lat_mu - z[r]), t.lat_div);
        part[r] = u * u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part[r] += __shfl_xor_sync(0xffffffffu, part[r], o);
      }
      gsync();
      if ((tid & 31) == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) sD[r * ld + (tid >> 5)] = part[r];
      }
      gsync();
      if (tid == H - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float v = 0.f;
          for (int q = 0; q < nw; ++q) v += sD[r * ld + q];
          f[r] = 0.5f * v;
        }
      }
      gsync();
    }
  };
  auto coef_at = [&](int s, int q, int r) {
    if (per_row_g) return sPost[coef_i * slot + r * ld + tid];
    float v = t.coef_scalar;
    if (t.coef_src == CO_IMG) v = W[t.coef_ref + tid];
    else if (t.coef_src == CO_VBUF) v = p.vtab[((size_t)s * kSrkGPoints + q) * H + tid];
    if (t.latent && tid == H - 1) v = 0.f;
    return v;
  };

  // ---- sites ----------------------------------------------------------------------------------------------------
  // J_f(z)^T fb for the drift evaluated at (z, point pt): returns the contribution to the state cotangent
  auto site_f = [&](int s, int site, const float (&z)[R], const snsde_point& pt, const float (&fb)[R], float (&zb)[R]) {
    const size_t rows = ((size_t)site * p.S + s) * p.B;
    put_state(z);
    put_control(pt, rows);
    if (jact && p.pstate_f != nullptr) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (valid[r]) p.pstate_f[(rows + row0 + r) * H + tid] = z[r];
    }
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r] = 0.f; zb[r] = 0.f; }
    eval_part(0, time_of(pt), acc, rows);
    // tail: f = clip(acc * tanh(z)) [+ KL channel]
    float fbj[R];
#pragma unroll
    for (int r = 0; r < R; ++r) fbj[r] = fb[r];
    if (t.latent) {
      // f[H-1] = 0.5 sum_j u_j^2: its cotangent reaches every latent feature's drift value and state
      gsync();
      if (tid == H - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) sD[r * ld] = fb[r];
      }
      gsync();
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (tid < H - 1) {
          const float lk = sD[r * ld];
          const float u = __fdiv_rn(acc[r] - t.lat_theta * (t.lat_mu - z[r]), t.lat_div);       // LatentSDE: no clip / geometric
          fbj[r] = fb[r] + lk * __fdiv_rn(u, t.lat_div);
          zb[r] += lk * u * __fdiv_rn(t.lat_theta, t.lat_div);
        } else {
          fbj[r] = 0.f;
        }
      }
      gsync();
    }
    if (jact) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float th = t.geometric ? tanhf(z[r]) : 1.f;
        const float pre = acc[r] * th;
        const float f = t.clip_drift ? tanhf(pre) : pre;
        const float a_pre = t.clip_drift ? fbj[r] * (1.f - f * f) : fbj[r];
        if (t.geometric) zb[r] += a_pre * acc[r] * (1.f - th * th);
        if (final_i >= 0) sCot[final_i * slot + r * ld + tid] = a_pre * th;
      }
    }
    reverse_part(0, rows, zb);
  };
  // J_g(z)^T gb for the diffusion evaluated at (z, point pt, coefficient row q)
  auto site_g = [&](int s, int site, int q, const float (&z)[R], const snsde_point& pt, const float (&gb)[R], float (&zb)[R]) {
    const size_t rows = ((size_t)site * p.S + s) * p.B;
#pragma unroll
    for (int r = 0; r < R; ++r) zb[r] = 0.f;
    if (per_row_g) {
      put_state(z);
      if (jact && p.pstate_g != nullptr) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (valid[r]) p.pstate_g[(rows + row0 + r) * H + tid] = z[r];
      }
      float dummy[R];
      eval_part(1, time_of(pt), dummy, rows);
    }
    float gv_sum = 0.f;
    if (jact) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float coef = coef_at(s, q, r);
        float g, dg;
        diffusion_eval<false>(t, coef, z[r], pt.t, g, dg);
        float ay_g, a_coef, a_sth;
        diffusion_backward(t, coef, z[r], pt.t, g, gb[r], ay_g, a_coef, a_sth);
        zb[r] += ay_g;
        acc_theta += a_sth;
        if (per_row_g) sCot[coef_i * slot + r * ld + tid] = a_coef;
        else { acc_coef += a_coef; gv_sum += a_coef; }
      }
      if (t.coef_src == CO_VBUF && p.gvtab != nullptr) atomicAdd(p.gvtab + ((size_t)s * kSrkGPoints + q) * H + tid, gv_sum);
    }
    if (per_row_g) reverse_part(1, rows, zb);
  };

  // ---- the sweep ------------------------------------------------------------------------------------------------
  for (int s = p.S - 1; s >= 0; --s) {
    const snsde_step st = p.steps[s];
    const snsde_point p0 = p.points[s * kSrkPoints + 0], pq = p.points[s * kSrkPoints + 1];
    const snsde_point ph = p.points[s * kSrkPoints + 2], p1 = p.points[s * kSrkPoints + 3];
    const float h = st.h, sqrt_h = st.sqrt_h, rdt = __fdiv_rn(1.0f, h);

    float y0[R], w[R], u[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { y0[r] = 0.f; w[r] = 0.f; u[r] = 0.f; }
    if (jact) {
      float nrm[4], nrmu[4];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        y0[r] = p.states[(size_t)s * BH + (size_t)grow(r) * H + tid];
        if (p.dW != nullptr) {
          w[r] = p.dW[(size_t)s * BH + (size_t)grow(r) * H + tid];
          u[r] = p.dU[(size_t)s * BH + (size_t)grow(r) * H + tid];
        } else {
          const unsigned long long gb = p.row_offset + (unsigned long long)(row0 + r);
          if (r == 0 || (gb & 3ull) == 0ull) {
            philox_normals4(p.seed, (uint32_t)tid, (uint32_t)(gb >> 2), (uint32_t)s, nrm);
            philox_normals4_u(p.seed, (uint32_t)tid, (uint32_t)(gb >> 2), (uint32_t)s, nrmu);
          }
          w[r] = __fmul_rn(pick4(nrm, (int)(gb & 3ull)), sqrt_h);
          u[r] = levy_U(w[r], pick4(nrmu, (int)(gb & 3ull)), h, sqrt_h);
        }
      }
    }

    // ---- stage states, recomputed from y_s (same arithmetic as the forward kernels) ----
    const size_t none = ~(size_t)0;
    float acc[R], f0[R], g0[R], f1[R], g1[R], f2[R], g2[R], H01[R], H11[R], H02[R], H12[R], H13[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    put_state(y0); put_control(p0, none);
    eval_part(0, time_of(p0), acc, none);
    drift_of(acc, y0, f0);
    if (per_row_g) { float d_[R]; eval_part(1, time_of(p0), d_, none); }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float dg;
      g0[r] = 0.f;
      if (jact) diffusion_eval<false>(t, coef_at(s, 0, r), y0[r], p0.t, g0[r], dg);
      H01[r] = y0[r] + (1.0f * f0[r]) * h + ((0.0f * g0[r]) * u[r]) * rdt;
      H11[r] = y0[r] + (0.25f * f0[r]) * h + (-0.5f * g0[r]) * sqrt_h;
    }
    put_state(H01); put_control(p1, none);
    eval_part(0, time_of(p1), acc, none);
    drift_of(acc, H01, f1);
    if (per_row_g) { put_state(H11); float d_[R]; eval_part(1, time_of(pq), d_, none); }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float dg;
      g1[r] = 0.f;
      if (jact) diffusion_eval<false>(t, coef_at(s, 1, r), H11[r], pq.t, g1[r], dg);
      float a = y0[r] + (0.25f * f0[r]) * h + ((1.0f * g0[r]) * u[r]) * rdt;
      a = a + (0.25f * f1[r]) * h + ((0.5f * g1[r]) * u[r]) * rdt;
      H02[r] = a;
      float b = y0[r] + (1.0f * f0[r]) * h + (1.0f * g0[r]) * sqrt_h;
      b = b + (0.0f * f1[r]) * h + (0.0f * g1[r]) * sqrt_h;
      H12[r] = b;
    }
    put_state(H02); put_control(ph, none);
    eval_part(0, time_of(ph), acc, none);
    drift_of(acc, H02, f2);
    if (per_row_g) { put_state(H12); float d_[R]; eval_part(1, time_of(p1), d_, none); }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float dg;
      g2[r] = 0.f;
      if (jact) diffusion_eval<false>(t, coef_at(s, 2, r), H12[r], p1.t, g2[r], dg);
      float b = y0[r] + (0.0f * f0[r]) * h + (2.0f * g0[r]) * sqrt_h;
      b = b + (0.0f * f1[r]) * h + (-1.0f * g1[r]) * sqrt_h;
      b = b + (0.25f * f2[r]) * h + (0.5f * g2[r]) * sqrt_h;
      H13[r] = b;
    }

    // ---- reverse of the step ----
    float yb[R], fb0[R], fb1[R], fb2[R], gb0[R], gb1[R], gb2[R], gb3[R], zb[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float I_k = w[r];
      const float I_kk = (I_k * I_k - h) * 0.5f;
      const float I_kkk = (I_k * I_k * I_k - 3.0f * h * I_k) * (1.0f / 6.0f);
      const float c0 = I_kk / sqrt_h, c1 = u[r] * rdt, c2 = I_kkk * rdt;
      const float gw0 = -1.0f * I_k + 1.0f * c0 + 2.0f * c1 + -2.0f * c2;
      const float gw1 = (4.0f / 3.0f) * I_k + (-4.0f / 3.0f) * c0 + (-4.0f / 3.0f) * c1 + (5.0f / 3.0f) * c2;
      const float gw2 = (2.0f / 3.0f) * I_k + (1.0f / 3.0f) * c0 + (-2.0f / 3.0f) * c1 + (-2.0f / 3.0f) * c2;
      const float gw3 = c2;
      const float l = lam[r];
      yb[r] = l;
      fb0[r] = l * ((1.0f / 6.0f) * h); fb1[r] = l * ((1.0f / 6.0f) * h); fb2[r] = l * ((2.0f / 3.0f) * h);
      gb0[r] = l * gw0; gb1[r] = l * gw1; gb2[r] = l * gw2; gb3[r] = l * gw3;
    }
    site_g(s, 3, 1, H13, pq, gb3, zb);                        // g3 = g(t0+h/4, H13)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      yb[r] += zb[r]; gb0[r] += 2.0f * sqrt_h * zb[r]; gb1[r] -= sqrt_h * zb[r]; fb2[r] += 0.25f * h * zb[r]; gb2[r] += 0.5f * sqrt_h * zb[r];
    }
    site_g(s, 2, 2, H12, p1, gb2, zb);                        // g2 = g(t0+h, H12)
#pragma unroll
    for (int r = 0; r < R; ++r) { yb[r] += zb[r]; fb0[r] += h * zb[r]; gb0[r] += sqrt_h * zb[r]; }
    site_f(s, 2, H02, ph, fb2, zb);                           // f2 = f(t0+h/2, H02)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      yb[r] += zb[r]; fb0[r] += 0.25f * h * zb[r]; gb0[r] += (u[r] * rdt) * zb[r]; fb1[r] += 0.25f * h * zb[r]; gb1[r] += (0.5f * u[r] * rdt) * zb[r];
    }
    site_g(s, 1, 1, H11, pq, gb1, zb);                        // g1 = g(t0+h/4, H11)
#pragma unroll
    for (int r = 0; r < R; ++r) { yb[r] += zb[r]; fb0[r] += 0.25f * h * zb[r]; gb0[r] -= 0.5f * sqrt_h * zb[r]; }
    site_f(s, 1, H01, p1, fb1, zb);                           // f1 = f(t0+h, H01)
#pragma unroll
    for (int r = 0; r < R; ++r) { yb[r] += zb[r]; fb0[r] += h * zb[r]; }
    site_f(s, 0, y0, p0, fb0, zb);                            // f0 = f(t0, y0)
#pragma unroll
    for (int r = 0; r < R; ++r) yb[r] += zb[r];
    site_g(s, 0, 0, y0, p0, gb0, zb);                         // g0 = g(t0, y0)
#pragma unroll
    for (int r = 0; r < R; ++r) yb[r] += zb[r];

    if (jact) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        lam[r] = valid[r] ? p.grad_states[(size_t)s * BH + (size_t)grow(r) * H + tid] + yb[r] : 0.f;
    }
  }
  (void)SB;

  if (jact) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (valid[r]) p.grad_y0[(size_t)grow(r) * H + tid] = lam[r];
    if (t.bounded && t.g_theta >= 0) {
      float v = acc_theta * t.s_theta * (1.f - t.s_theta);
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) atomicAdd(p.grad_blob + t.g_theta, v);
    }
    if (t.coef_src == CO_IMG && t.g_sigma >= 0) {
      atomicAdd(p.grad_blob + t.g_sigma + tid, acc_coef * p.wimg[t.coef_ref + tid]);      // coef = exp(sigma_diag)
    } else if (t.coef_src == CO_SCALAR && t.g_sigma >= 0) {
      float v = acc_coef * t.coef_scalar;                                                 // coef = exp(sigma)
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) atomicAdd(p.grad_blob + t.g_sigma, v);
    }
  }
}

// aux rows (sin t, cos t, 1) of every evaluation site: drift sites at points (0, 3, 2) = t0, t0+h, t0+h/2 and diffusion
// sites at points (0, 1, 3, 1) = t0, t0+h/4, t0+h, t0+h/4
__global__ void bwd_srk_aux_kernel(const snsde_point* __restrict__ points, int S, int B, float* __restrict__ aux_f,
                                   float* __restrict__ aux_g) {
  const size_t n = (size_t)S * B;
  const int fpt[3] = {0, 3, 2}, gpt[4] = {0, 1, 3, 1};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t s = i / B;
    for (int k = 0; k < 3; ++k) {
      const snsde_point pt = points[s * kSrkPoints + fpt[k]];
      float* a = aux_f + 3 * ((size_t)k * n + i);
      a[0] = pt.sin_t; a[1] = pt.cos_t; a[2] = 1.0f;
    }
    for (int k = 0; k < 4; ++k) {
      const snsde_point pt = points[s * kSrkPoints + gpt[k]];
      float* a = aux_g + 3 * ((size_t)k * n + i);
      a[0] = pt.sin_t; a[1] = pt.cos_t; a[2] = 1.0f;
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
size_t bwd_srk_group_smem_floats(const Program& pg, int n_rops, int R, int has_lipswish) {
  return (size_t)(3 + n_rops * (has_lipswish ? 3 : 2)) * R * pg.ld;
}

template <int R, int NTMAX, bool WS>
static cudaError_t bwd_srk_launch_one(const BwdParams& p, int grid, int nt, size_t smem, cudaStream_t stream) {
  auto kern = snsde_bwd_srk_kernel<R, NTMAX, WS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, nt, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t bwd_srk_launch(const BwdParams& p, int R, size_t smem, cudaStream_t stream) {
  const int nt = p.groups * p.nw * 32;
  const int n_groups = (p.B + R - 1) / R;
  const int grid = (n_groups + p.groups - 1) / p.groups;
  const bool ws = p.smem_w_floats > 0;
  if (R == 4 && nt <= 512) return ws ? bwd_srk_launch_one<4, 512, true>(p, grid, nt, smem, stream) : bwd_srk_launch_one<4, 512, false>(p, grid, nt, smem, stream);
  if (R == 1 && nt <= 512) return ws ? bwd_srk_launch_one<1, 512, true>(p, grid, nt, smem, stream) : bwd_srk_launch_one<1, 512, false>(p, grid, nt, smem, stream);
  if (R == 4) return ws ? bwd_srk_launch_one<4, 1024, true>(p, grid, nt, smem, stream) : bwd_srk_launch_one<4, 1024, false>(p, grid, nt, smem, stream);
  return cudaErrorInvalidValue;
}

cudaError_t bwd_srk_fill_aux(const snsde_point* points, int S, int B, float* aux_f, float* aux_g, cudaStream_t stream) {
  const size_t n = (size_t)S * B;
  if (n == 0) return cudaSuccess;
  bwd_srk_aux_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, stream>>>(points, S, B, aux_f, aux_g);
  return cudaGetLastError();
}

}  // namespace snsde
