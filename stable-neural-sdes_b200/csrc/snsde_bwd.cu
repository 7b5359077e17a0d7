// Reverse sweep of the Euler-Maruyama solve (and of Milstein with an elementwise diffusion): dL/dy0 and the per-op
// cotangents behind the weight gradients.
//
// What it replaces: the autograd graph the reference builds through torchsde.sdeint's Python step loop
// (/root/reference/benchmark_classification/common_sde.py:156-162 `loss.backward()` through
// benchmark_classification/models_sde/neuralsde.py:78-82) - S x ~50 autograd nodes, each a kernel launch.
//
// One launch walks the trajectory backwards.  Same decomposition as the forward FMA kernel (snsde_fma.cu): a row
// group = R batch rows x nw warps, thread j = feature j, group-local hand-offs only.  Per step s (descending):
//   1. y_s (saved by the forward solve) -> registers + shared memory; X(t_s) from the prefetched spline row;
//   2. the dense program is re-evaluated, every op's activated output kept in its own shared-memory slot and
//      written to HBM where a later op consumes it (P buffers: the GEMM operands of the weight gradients);
//   3. the SDE update is differentiated:  y_{s+1} = y + f h + g dW  with lambda = dL/dy_{s+1}
//        a_f = lambda h,  a_g = lambda dW,  through tanh clips / geometric term / nan_to_num / sigmoid(theta);
//   4. ops in reverse: D = cot * act'(.) -> HBM (D buffers) and shared memory; cot_src[k] = sum_j D[j] W[j][k]
//      (W read in its nn.Linear layout, coalesced over k); state inputs accumulate into lambda;
//   5. lambda_s = dL/dy_s(outputs) + a_y.
// theta / sigma / sigma_diag gradients are register partial sums reduced at the end; the cotangent of the
// row-independent coefficient table goes to gvtab[S][H] and is pulled through the noise_t / g_net networks by
// vec_bwd_kernel.  The host (snsde_api.cu) then forms every weight gradient as D^T P with cuBLAS.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>

#include "snsde_bwd.cuh"
#include "snsde_fma.cuh"

namespace snsde {

template <int R, int NTMAX, bool WS>
__global__ void __launch_bounds__(NTMAX) snsde_bwd_kernel(const BwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const Program& pg = p.prog;
  const TailOp& t = pg.tail;
  const int H = pg.H, C = pg.C, ld = pg.ld;
  const int nw = p.nw, GT = nw * 32;
  const int gid = threadIdx.x / GT, tid = threadIdx.x - gid * GT;
  const int n_rops = p.n_rops;
  const int slot = R * ld;

  // ---- shared memory: [weights][group: Y X D | post[n_rops] | cot[n_rops] | pre[n_rops]? | 2 spline stages] ----
  const int stage_floats = pg.uses_control ? R * 4 * C : 0;
  const int group_floats = (3 + n_rops * (p.has_lipswish ? 3 : 2)) * slot + 2 * stage_floats;
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
  float* const sStage = gbase + (3 + n_rops * (p.has_lipswish ? 3 : 2)) * slot;

  const int row0 = (blockIdx.x * p.groups + gid) * R;
  if (row0 >= p.B) return;
  auto grow = [&](int r) { return min(row0 + r, p.B - 1); };
  auto gsync = [&]() { group_sync(gid, nw); };
  auto stage_buf = [&](int s) { return sStage + (s & 1) * stage_floats; };
  auto prefetch_spline = [&](int s) {
    if (!pg.uses_control || s < 0) return;
    const int interval = p.steps[s].interval;
    float* dst = stage_buf(s);
    for (int q = tid; q < R * C; q += GT) {
      const int r = q / C, c = q - r * C;
      const float* src = p.coeffs + (size_t)grow(r) * p.coeff_row_stride + (size_t)interval * 4 * C + 4 * c;
      cp_async16(dst + r * 4 * C + 4 * c, src);
    }
    cp_async_commit();
  };
  prefetch_spline(p.S - 1);

  // the op list reads buffers by id: map each buffer id to the slot of the op that wrote it last (program order)
  auto buf_of = [&](int src_op, int buf_id) -> const float* {
    if (src_op == SRC_STATE) return sY;
    if (src_op == SRC_CONTROL) return sX;
    (void)buf_id;
    for (int i = 0; i < n_rops; ++i) if (p.rop[i] == src_op) return sPost + i * slot;
    return sY;
  };
  auto rop_index = [&](int op_index) { for (int i = 0; i < n_rops; ++i) if (p.rop[i] == op_index) return i; return -1; };

  const bool jact = tid < H;
  const size_t BH = (size_t)p.B * H;
  float lam[R];
  bool valid[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    valid[r] = row0 + r < p.B;
    lam[r] = (jact && valid[r]) ? p.grad_states[(size_t)p.S * BH + (size_t)grow(r) * H + tid] : 0.f;
  }
  float acc_theta = 0.f, acc_coef = 0.f;
  const int final_i = n_rops - 1;                                   // the final drift op is the last per-row op
  const int coef_i = t.coef_src == CO_RBUF ? rop_index(t.coef_op) : -1;

  for (int s = p.S - 1; s >= 0; --s) {
    const snsde_step st = p.steps[s];
    const TimePoint tp{st.t0, st.sin_t0, st.cos_t0};
    const size_t srow = (size_t)s * p.B;

    // ---- 1. state and control of step s ----
    float y[R], w[R];
    float vcoef = t.coef_scalar;
    if (jact) {
      if (t.coef_src == CO_IMG) vcoef = W[t.coef_ref + tid];
      else if (t.coef_src == CO_VBUF) vcoef = p.vtab[(size_t)s * H + tid];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        y[r] = p.states[(size_t)s * BH + (size_t)grow(r) * H + tid];
        sY[r * ld + tid] = y[r];
      }
      float nrm[4];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (p.dW != nullptr) {
          w[r] = p.dW[(size_t)s * BH + (size_t)grow(r) * H + tid];
        } else {
          const unsigned long long gb = p.row_offset + (unsigned long long)(row0 + r);
          if (r == 0 || (gb & 3ull) == 0ull) philox_normals4(p.seed, (uint32_t)tid, (uint32_t)(gb >> 2), (uint32_t)s, nrm);
          w[r] = __fmul_rn(pick4(nrm, (int)(gb & 3ull)), st.sqrt_h);
        }
      }
    }
    if (pg.uses_control) {
      cp_async_wait_all();
      gsync();
      const float* stg = stage_buf(s);
      for (int q = tid; q < R * C; q += GT) {
        const int r = q / C, c = q - r * C;
        const float* row = stg + r * 4 * C;
        float inner = 0.5f * row[2 * C + c] + __fdiv_rn(row[3 * C + c] * st.frac, 3.0f);
        inner = row[C + c] + inner * st.frac;
        const float x = row[c] + inner * st.frac;
        sX[r * ld + c] = x;
        if (p.xbuf != nullptr && row0 + r < p.B) p.xbuf[(srow + row0 + r) * C + c] = x;
      }
      prefetch_spline(s - 1);
    }

    // ---- 2. forward re-evaluation, every op's output in its own slot ----
    float acc[R];
    for (int i = 0; i < n_rops; ++i) {
      const DenseOp& op = pg.ops[p.rop[i]];
      gsync();
      if (tid < op.N) {
        // dense_eval reads sm.buf(id): resolve the producing slots by hand
        float a[R];
        {
          float init = op.b_off >= 0 ? W[op.b_off + tid] : 0.f;
          if (op.tmode == TM_SINCOS) {
            const float* tw = W + op.tw_off;
            init = fmaf(tp.cos_t, tw[op.N + tid], fmaf(tp.sin_t, tw[tid], init));
          }
#pragma unroll
          for (int r = 0; r < R; ++r) a[r] = init;
          if (op.src >= 0) dot_accumulate<R>(a, buf_of(op.src_op, op.src), ld, W + op.w_off + tid, op.K, op.N);
          if (op.src2 >= 0) dot_accumulate<R>(a, buf_of(op.src2_op, op.src2), ld, W + op.w2_off + tid, op.K2, op.N);
        }
        if (op.final_drift) {
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = a[r];
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float post = act_apply(a[r], op.act);
            sPost[i * slot + r * ld + tid] = post;
            if (p.has_lipswish) sPre[i * slot + r * ld + tid] = a[r];
            if (p.pbuf[p.rop[i]] != nullptr && valid[r]) p.pbuf[p.rop[i]][(srow + row0 + r) * op.N + tid] = post;
          }
        }
      }
    }
    if (coef_i >= 0) gsync();

    // ---- 3. the SDE update, differentiated ----
    float a_y[R];
    float gv_sum = 0.f;
    if (t.latent) {
      // LatentSDE.f_aug (latent_sde.py:77-82): y_kl' = y_kl + h * 0.5 sum_j u_j^2, u_j = (f_j - theta (mu - y_j)) / div.
      // lambda of the KL channel reaches every latent feature: broadcast it through sD (idle since the previous
      // step's reverse ops, which are behind the hand-offs of the re-evaluation above).
      if (tid == H - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) sD[r * ld] = lam[r];
      }
      gsync();
      if (jact) {
#pragma unroll
        for (int r = 0; r < R; ++r) {                               // g_aug is constant: the noise term has no cotangent
          float a_f = 0.f, ay = lam[r];
          if (tid < H - 1) {
            const float lk = sD[r * ld] * st.h;
            const float u = __fdiv_rn(acc[r] - t.lat_theta * (t.lat_mu - y[r]), t.lat_div);
            a_f = lam[r] * st.h + lk * __fdiv_rn(u, t.lat_div);
            ay += lk * u * __fdiv_rn(t.lat_theta, t.lat_div);
          }
          a_y[r] = ay;
          sCot[final_i * slot + r * ld + tid] = a_f;
        }
      }
    } else if (jact) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float d = acc[r];
        const float th = t.geometric ? tanhf(y[r]) : 1.f;
        const float pre = d * th;
        const float f = t.clip_drift ? tanhf(pre) : pre;
        const float coef = coef_i >= 0 ? sPost[coef_i * slot + r * ld + tid] : vcoef;
        float g, dgdy;
        diffusion_eval<false>(t, coef, y[r], st.t0, g, dgdy);
        const float a_f = lam[r] * st.h, a_g = lam[r] * w[r];
        const float a_pre = t.clip_drift ? a_f * (1.f - f * f) : a_f;
        float ay = lam[r];
        if (t.geometric) ay += a_pre * d * (1.f - th * th);
        float ay_g, a_coef, a_sth;
        diffusion_backward(t, coef, y[r], st.t0, g, a_g, ay_g, a_coef, a_sth);
        if (t.milstein) {                                    // y_{s+1} += 0.5 (g v) dg/dy, v = dW^2 - h  (diagonal closed form)
          float my, mc, ms;
          milstein_backward(t, coef, y[r], st.t0, g, __fmul_rn(w[r], w[r]) - st.h, lam[r], my, mc, ms);
          ay_g += my; a_coef += mc; a_sth += ms;
        }
        a_y[r] = ay + ay_g;
        acc_theta += a_sth;
        sCot[final_i * slot + r * ld + tid] = a_pre * th;           // cotangent of the drift pre-activation
        if (coef_i >= 0) sCot[coef_i * slot + r * ld + tid] = a_coef;
        else { acc_coef += a_coef; gv_sum += a_coef; }
      }
      if (t.coef_src == CO_VBUF && p.gvtab != nullptr) atomicAdd(p.gvtab + (size_t)s * H + tid, gv_sum);
    }

    // ---- 4. ops in reverse ----
    for (int i = n_rops - 1; i >= 0; --i) {
      const DenseOp& op = pg.ops[p.rop[i]];
      gsync();                                   // cot slot i complete (written by the tail or by later ops)
      if (tid < op.N) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float cot = sCot[i * slot + r * ld + tid];
          const float dlt = op.final_drift ? cot
                                           : cot * act_grad(p.has_lipswish ? sPre[i * slot + r * ld + tid] : 0.f,
                                                            sPost[i * slot + r * ld + tid], op.act);
          sD[r * ld + tid] = dlt;
          if (valid[r]) p.dbuf[p.rop[i]][(srow + row0 + r) * op.N + tid] = dlt;
        }
      }
      gsync();
      for (int part = 0; part < 2; ++part) {
        const int src_op = part == 0 ? op.src_op : op.src2_op;
        const int K = part == 0 ? op.K : op.K2;
        const int col = part == 0 ? op.g_col : op.g_col2;
        if ((part == 0 ? op.src : op.src2) < 0 || src_op == SRC_CONTROL || src_op == SRC_ABSENT) continue;
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

    // ---- 5. lambda_s ----
    if (jact) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        lam[r] = valid[r] ? p.grad_states[(size_t)s * BH + (size_t)grow(r) * H + tid] + a_y[r] : 0.f;
    }
  }

  if (jact) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (valid[r]) p.grad_y0[(size_t)grow(r) * H + tid] = lam[r];
    if (t.bounded && t.g_theta >= 0) {
      // d sigmoid(theta) / d theta = s (1 - s); one atomic per warp
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

// ---- backward of the row-independent coefficient networks ------------------------------------------
// One CTA per step: re-evaluates the vec ops at t_s, takes gvtab[s] as the cotangent of the final vector and
// accumulates weight / bias / time-column gradients into the blob with atomics (S x N x K of them: tiny).
__global__ void __launch_bounds__(1024) vec_bwd_kernel(const Program pg, const float* __restrict__ wimg,
                                                       const float* __restrict__ blob,
                                                       const snsde_step* __restrict__ steps,
                                                       const snsde_point* __restrict__ points, int npg,
                                                       const float* __restrict__ gvtab, float* __restrict__ grad_blob) {
  extern __shared__ __align__(16) float vsm[];          // post[kMaxOps][ld] pre[kMaxOps][ld] cot[kMaxOps][ld] dlt[ld]
  const int s = blockIdx.x / npg, q = blockIdx.x - s * npg, j = threadIdx.x, ld = pg.ld;
  float* const post = vsm;
  float* const pre = vsm + kMaxOps * ld;
  float* const cot = vsm + 2 * kMaxOps * ld;
  float* const dlt = vsm + 3 * kMaxOps * ld;
  snsde_step st = steps[s];
  if (points != nullptr) {                                 // SRK: coefficient rows at t0, t0+h/4, t0+h
    const snsde_point pt = points[s * kSrkPoints + (q == 0 ? 0 : (q == 1 ? 1 : 3))];
    st.t0 = pt.t; st.sin_t0 = pt.sin_t; st.cos_t0 = pt.cos_t;
  }
  int last = -1;
  for (int o = 0; o < pg.n_ops; ++o) {
    const DenseOp& op = pg.ops[o];
    if (!op.vec) continue;
    if (j < op.N) {
      float v = op.b_off >= 0 ? wimg[op.b_off + j] : 0.f;
      if (op.tmode == TM_SINCOS) v = fmaf(st.cos_t0, wimg[op.tw_off + op.N + j], fmaf(st.sin_t0, wimg[op.tw_off + j], v));
      else if (op.tmode == TM_RAW) v = fmaf(st.t0, wimg[op.tw_off + j], v);
      if (op.src >= 0) {
        const float* src = post + op.src_op * ld;
        const float* w = wimg + op.w_off + j;
        for (int k = 0; k < op.K; ++k) v = fmaf(src[k], w[(size_t)k * op.N], v);
      }
      pre[o * ld + j] = v;
      post[o * ld + j] = act_apply(v, op.act);
    }
    last = o;
    __syncthreads();
  }
  if (last < 0) return;
  if (j < pg.ops[last].N) cot[last * ld + j] = gvtab[(size_t)blockIdx.x * pg.H + j];
  for (int o = pg.n_ops - 1; o >= 0; --o) {
    const DenseOp& op = pg.ops[o];
    if (!op.vec) continue;
    __syncthreads();
    if (j < op.N) {
      const float d = cot[o * ld + j] * act_grad(pre[o * ld + j], post[o * ld + j], op.act);
      dlt[j] = d;
      if (op.g_b >= 0) atomicAdd(grad_blob + op.g_b + j, d);
      if (op.tmode == TM_SINCOS) {
        atomicAdd(grad_blob + op.g_w + (size_t)j * op.g_ldw + 0, d * st.sin_t0);
        atomicAdd(grad_blob + op.g_w + (size_t)j * op.g_ldw + 1, d * st.cos_t0);
      } else if (op.tmode == TM_RAW) {
        atomicAdd(grad_blob + op.g_w + (size_t)j * op.g_ldw, d * st.t0);
      }
      if (op.src >= 0) {
        const float* src = post + op.src_op * ld;
        float* gw = grad_blob + op.g_w + (size_t)j * op.g_ldw + op.g_col;
        for (int k = 0; k < op.K; ++k) atomicAdd(gw + k, d * src[k]);
      }
    }
    __syncthreads();
    if (op.src >= 0 && j < op.K) {
      float c = 0.f;
      const float* w = blob + op.g_w + op.g_col + j;
      for (int n = 0; n < op.N; ++n) c = fmaf(dlt[n], w[(size_t)n * op.g_ldw], c);
      cot[op.src_op * ld + j] = c;
    }
  }
}

__global__ void bwd_aux_kernel(const snsde_step* __restrict__ steps, int S, int B, float* __restrict__ aux) {
  const size_t n = (size_t)S * B;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const snsde_step st = steps[i / B];
    aux[3 * i + 0] = st.sin_t0;
    aux[3 * i + 1] = st.cos_t0;
    aux[3 * i + 2] = 1.0f;
  }
}

// ---- host side -------------------------------------------------------------------------------------
size_t bwd_group_smem_floats(const Program& pg, int n_rops, int R, int has_lipswish) {
  size_t f = (size_t)(3 + n_rops * (has_lipswish ? 3 : 2)) * R * pg.ld;
  if (pg.uses_control) f += (size_t)2 * R * 4 * pg.C;
  return f;
}

template <int R, int NTMAX, bool WS>
static cudaError_t bwd_launch_one(const BwdParams& p, int grid, int nt, size_t smem, cudaStream_t stream) {
  auto kern = snsde_bwd_kernel<R, NTMAX, WS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, nt, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t bwd_launch(const BwdParams& p, int R, size_t smem, cudaStream_t stream) {
  const int nt = p.groups * p.nw * 32;
  const int n_groups = (p.B + R - 1) / R;
  const int grid = (n_groups + p.groups - 1) / p.groups;
  const bool ws = p.smem_w_floats > 0;
  if (R == 8 && nt <= 512) return ws ? bwd_launch_one<8, 512, true>(p, grid, nt, smem, stream) : bwd_launch_one<8, 512, false>(p, grid, nt, smem, stream);
  if (R == 4 && nt <= 512) return ws ? bwd_launch_one<4, 512, true>(p, grid, nt, smem, stream) : bwd_launch_one<4, 512, false>(p, grid, nt, smem, stream);
  if (R == 1 && nt <= 512) return ws ? bwd_launch_one<1, 512, true>(p, grid, nt, smem, stream) : bwd_launch_one<1, 512, false>(p, grid, nt, smem, stream);
  if (R == 4) return ws ? bwd_launch_one<4, 1024, true>(p, grid, nt, smem, stream) : bwd_launch_one<4, 1024, false>(p, grid, nt, smem, stream);
  return cudaErrorInvalidValue;
}

cudaError_t bwd_fill_aux(const snsde_step* steps, int S, int B, float* aux, cudaStream_t stream) {
  const size_t n = (size_t)S * B;
  if (n == 0) return cudaSuccess;
  bwd_aux_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, stream>>>(steps, S, B, aux);
  return cudaGetLastError();
}

cudaError_t vec_bwd_launch(const Program& pg, const float* wimg, const float* blob, const snsde_step* steps,
                           const snsde_point* points, int S, int npg, const float* gvtab, float* grad_blob, cudaStream_t stream) {
  if (S == 0) return cudaSuccess;
  const int nt = std::max(32, (std::max(pg.H, pg.HH) + 31) & ~31);
  const size_t smem = sizeof(float) * (size_t)(3 * kMaxOps + 1) * pg.ld;
  cudaError_t e = cudaFuncSetAttribute(vec_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  vec_bwd_kernel<<<S * npg, nt, smem, stream>>>(pg, wimg, blob, steps, points, npg, gvtab, grad_blob);
  return cudaGetLastError();
}

}  // namespace snsde
