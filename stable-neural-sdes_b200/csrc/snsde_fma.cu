// fp32 FMA kernel: the whole fixed-step SDE solve for R batch rows per CTA in ONE launch.
//
// Replaces the Python step loop of torchsde.sdeint (BaseSDESolver.integrate + Euler/Milstein
// .step) together with the per-step Diffusion_model.f/g evaluation
// (/root/reference/benchmark_classification/models_sde/neuralsde.py:295-307) and
// torchcde.CubicSpline.evaluate (:296).  Per step, per CTA:
//
//   cp.async-prefetched spline row  ->  X(t)        (coalesced 16C-byte rows, double buffered)
//   program of dense ops            ->  drift pre-activation in registers (thread j = feature j)
//   row-independent noise nets      ->  computed once per CTA, not once per row
//   Philox/Box-Muller or table dW   ->  y += f*h + g*dW (+ Milstein term), all in registers
//   emits                           ->  out[slot] (lerp) or fused final_index capture
//
// This is the generic path: any (input_option, noise_option), any H/HH/C/L, and the
// north star's "warp-shuffle/FMA" path for hidden < 64.  Weights are staged once into shared
// memory (as much of the image as fits); the rest is read through L2.
#include <cuda_runtime.h>
#include <math.h>

#include "snsde_common.cuh"
#include "snsde_math.cuh"
#include "snsde_rng.cuh"

namespace snsde {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

struct SmemMap {
  float* base;           // [row bufs: kNumRowBufs x R x ld][vec bufs: kNumVecBufs x ld]
  int row_buf_floats;    // R * ld
  int ld;
  float* stage0;         // 2 spline stages of stage_floats each
  int stage_floats;
  const float* w;        // staged weights (first smem_w_floats of the image)
  __device__ __forceinline__ float* buf(int id) const {
    return id < kNumRowBufs ? base + id * row_buf_floats
                            : base + kNumRowBufs * row_buf_floats + (id - kNumRowBufs) * ld;
  }
  __device__ __forceinline__ float* stage(int i) const { return stage0 + (i & 1) * stage_floats; }
};

// acc[r] += sum_k src[r][k] * Wt[k][j]
template <int ROWS>
__device__ __forceinline__ void dot_accumulate(float (&acc)[ROWS], const float* __restrict__ src, int ld,
                                               const float* __restrict__ wt, int K, int N, int j) {
  const float* w = wt + j;
  int k = 0;
  const int K4 = K & ~3;
  for (; k < K4; k += 4) {
    const float w0 = w[(size_t)k * N], w1 = w[(size_t)(k + 1) * N];
    const float w2 = w[(size_t)(k + 2) * N], w3 = w[(size_t)(k + 3) * N];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(src + r * ld + k);
      acc[r] = fmaf(a.x, w0, acc[r]);
      acc[r] = fmaf(a.y, w1, acc[r]);
      acc[r] = fmaf(a.z, w2, acc[r]);
      acc[r] = fmaf(a.w, w3, acc[r]);
    }
  }
  for (; k < K; ++k) {
    const float wk = w[(size_t)k * N];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(src[r * ld + k], wk, acc[r]);
  }
}

// acc[r] += sum_k src[r][k] * wrow[k]   - the TRANSPOSED product: thread j walks row j of an [in][out] image
// (contiguous), used by the Milstein vjp through the noise network; read through L1/L2 (a shared-memory copy
// would be 32-way bank conflicted at this access pattern).
template <int ROWS>
__device__ __forceinline__ void dot_rows_T(float (&acc)[ROWS], const float* __restrict__ src, int ld,
                                           const float* __restrict__ wrow, int K) {
  for (int k = 0; k < K; ++k) {
    const float wk = __ldg(wrow + k);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(src[r * ld + k], wk, acc[r]);
  }
}

template <int ROWS>
__device__ __forceinline__ void dense_eval(float (&acc)[ROWS], const DenseOp& op, const FmaParams& p,
                                           const SmemMap& sm, const snsde_step& st, int j) {
  const int ld = p.prog.ld;
  auto wptr = [&](int off, int count) -> const float* {
    return (off + count <= p.smem_w_floats) ? sm.w + off : p.wimg + off;
  };
  float init = op.b_off >= 0 ? wptr(op.b_off, op.N)[j] : 0.f;
  if (op.tmode == TM_SINCOS) {
    const float* tw = wptr(op.tw_off, 2 * op.N);
    init = fmaf(st.cos_t0, tw[op.N + j], fmaf(st.sin_t0, tw[j], init));
  } else if (op.tmode == TM_RAW) {
    init = fmaf(st.t0, wptr(op.tw_off, op.N)[j], init);
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) acc[r] = init;
  if (op.src >= 0) dot_accumulate<ROWS>(acc, sm.buf(op.src), ld, wptr(op.w_off, op.K * op.N), op.K, op.N, j);
  if (op.src2 >= 0) dot_accumulate<ROWS>(acc, sm.buf(op.src2), ld, wptr(op.w2_off, op.K2 * op.N), op.K2, op.N, j);
#pragma unroll
  for (int r = 0; r < ROWS; ++r) acc[r] = act_apply(acc[r], op.act);
}

template <int R, int NTMAX>
__global__ void __launch_bounds__(NTMAX) snsde_fma_kernel(const FmaParams p) {
  extern __shared__ __align__(16) float smem[];
  const Program& pg = p.prog;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int H = pg.H, C = pg.C, ld = pg.ld;
  const int row0 = blockIdx.x * R;

  // ---- shared memory carve-up: [row bufs][vec bufs][2 spline stages][weights] ----
  SmemMap sm;
  {
    sm.base = smem;
    sm.row_buf_floats = R * ld;
    sm.ld = ld;
    sm.stage0 = smem + kNumRowBufs * R * ld + kNumVecBufs * ld;
    sm.stage_floats = pg.uses_control ? R * 4 * C : 0;
    float* w = sm.stage0 + 2 * sm.stage_floats;
    for (int i = tid; i < p.smem_w_floats; i += NT) w[i] = p.wimg[i];
    sm.w = w;
  }
  float* const sY = sm.buf(BUF_Y);
  float* const sX = sm.buf(BUF_X);
  auto grow = [&](int r) { return min(row0 + r, p.B - 1); };   // clamped row (local to this shard)

  auto prefetch_spline = [&](int s) {
    if (!pg.uses_control || s >= p.S) return;
    const int interval = p.steps[s].interval;
    float* dst = sm.stage(s);
    for (int i = tid; i < R * C; i += NT) {
      const int r = i / C, q = i - r * C;
      const float* src = p.coeffs + (size_t)grow(r) * p.coeff_row_stride + (size_t)interval * 4 * C + 4 * q;
      cp_async16(dst + r * 4 * C + 4 * q, src);
    }
    cp_async_commit();
  };
  prefetch_spline(0);

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

  for (int s = 0; s < p.S; ++s) {
    const snsde_step st = p.steps[s];

    // explicit increments (parity mode): issue the loads now, consume at the end of the step
    float dw[R];
    if (p.dW != nullptr && jact) {
#pragma unroll
      for (int r = 0; r < R; ++r) dw[r] = p.dW[((size_t)s * p.B + grow(r)) * H + tid];
    }

    // ---- control path X(t0): a + (b + (two_c/2 + three_d*frac/3)*frac)*frac ----
    if (pg.uses_control) {
      cp_async_wait_all();
      __syncthreads();
      const float* stg = sm.stage(s);
      for (int i = tid; i < R * C; i += NT) {
        const int r = i / C, c = i - r * C;
        const float* row = stg + r * 4 * C;
        float inner = 0.5f * row[2 * C + c] + __fdiv_rn(row[3 * C + c] * st.frac, 3.0f);
        inner = row[C + c] + inner * st.frac;
        sX[r * ld + c] = row[c] + inner * st.frac;
      }
      prefetch_spline(s + 1);
    }
    __syncthreads();

    // ---- dense program ----
    float acc[R];
    for (int o = 0; o < pg.n_ops; ++o) {
      const DenseOp& op = pg.ops[o];
      if (o > 0) __syncthreads();      // op o reads what op o-1 wrote (dst != src by construction)
      if (tid < op.N) {
        if (op.vec) {
          float a1[1];
          dense_eval<1>(a1, op, p, sm, st, tid);
          sm.buf(op.dst)[tid] = a1[0];
        } else {
          dense_eval<R>(acc, op, p, sm, st, tid);
          if (!op.final_drift) {
            float* dst = sm.buf(op.dst);
#pragma unroll
            for (int r = 0; r < R; ++r) dst[r * ld + tid] = acc[r];
          }
        }
      }
    }

    // ---- SDE update (torchsde Euler.step / Milstein.step) ----
    const TailOp& t = pg.tail;
    const bool net_vjp = t.milstein && t.vjp_kind != 0;       // block-uniform
    float* const sA = sm.buf(BUF_X);                            // scratch of the vjp (dead since the first ops)
    float* const sB = sm.buf(BUF_U);
    if (jact) {
      float vcoef = t.coef_scalar;
      if (t.coef_src == CO_IMG) vcoef = p.wimg[t.coef_ref + tid];
      else if (t.coef_src == CO_VBUF) vcoef = sm.buf(t.coef_ref)[tid];
      const float* rcoef = (t.coef_src == CO_RBUF) ? sm.buf(t.coef_ref) : nullptr;
      float nrm[4];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float d = acc[r];
        if (t.geometric) d = d * tanhf(y[r]);
        if (t.clip_drift) d = tanhf(d);
        const float coef = rcoef ? rcoef[r * ld + tid] : vcoef;
        float g, dgdy;
        diffusion_eval<false>(t, coef, y[r], st.t0, g, dgdy);
        float w;
        if (p.dW != nullptr) {
          w = dw[r];
        } else {
          const unsigned long long gb = p.row_offset + (unsigned long long)(row0 + r);
          if (r == 0 || (gb & 3ull) == 0ull) philox_normals4(p.seed, (uint32_t)tid, (uint32_t)(gb >> 2), (uint32_t)s, nrm);
          w = __fmul_rn(pick4(nrm, (int)(gb & 3ull)), st.sqrt_h);
        }
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
      __syncthreads();
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
        __syncthreads();
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
    for (int e = st.emit_begin; e < st.emit_end; ++e) emit(p.emits[e]);
  }
}

// ---- host-side launcher --------------------------------------------------------------------

size_t fma_smem_bytes(const Program& pg, int R, int smem_w_floats) {
  size_t f = (size_t)kNumRowBufs * R * pg.ld + (size_t)kNumVecBufs * pg.ld;
  if (pg.uses_control) f += (size_t)2 * R * 4 * pg.C;
  f += smem_w_floats;
  return f * sizeof(float);
}

template <int R, int NTMAX>
static cudaError_t launch_one(const FmaParams& p, int grid, int nt, size_t smem, cudaStream_t stream) {
  auto kern = snsde_fma_kernel<R, NTMAX>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, nt, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t fma_launch(const FmaParams& p, int R, int nt, size_t smem, cudaStream_t stream) {
  const int grid = (p.B + R - 1) / R;
  if (nt <= 256) {
    switch (R) {
      case 1: return launch_one<1, 256>(p, grid, nt, smem, stream);
      case 2: return launch_one<2, 256>(p, grid, nt, smem, stream);
      case 4: return launch_one<4, 256>(p, grid, nt, smem, stream);
      case 8: return launch_one<8, 256>(p, grid, nt, smem, stream);
      case 16: return launch_one<16, 256>(p, grid, nt, smem, stream);
    }
  } else {
    switch (R) {
      case 1: return launch_one<1, 1024>(p, grid, nt, smem, stream);
      case 2: return launch_one<2, 1024>(p, grid, nt, smem, stream);
      case 4: return launch_one<4, 1024>(p, grid, nt, smem, stream);
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace snsde
