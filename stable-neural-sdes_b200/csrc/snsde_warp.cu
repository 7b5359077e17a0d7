// Warp-owned fp32 kernel for hidden sizes <= 32 (the north star's "warp-shuffle / FMA path for hidden < 64").
// A ROW GROUP = R batch rows owned end to end by a PAIR of warps: a main warp that runs the dependent chain of a solver
// step (the layers, the SDE update, the outputs) and a helper warp that prepares, one batch of steps ahead, everything
// that does not depend on the state: X(t) of every row, the Brownian increments, the row-independent diffusion
// coefficient.  Lane j = feature j; the SDE state lives in the main warp's registers.  The pair meets at one named
// barrier per batch of 4 steps; inside a step the main warp synchronises with nobody (__syncwarp only).
//
// Replaces, like snsde_fma.cu, the Python step loop of torchsde.sdeint (Euler.step / Milstein.step) with the
// per-step Diffusion_model.f/g evaluation (/root/reference/benchmark_classification/models_sde/neuralsde.py:295-307),
// torchcde.CubicSpline.evaluate (:296) and the tutorial NeuralLSDEFunc (notebook cell 7) - for the shapes where the
// interpreter kernel is pure latency: with one warp per SM sub-partition every instruction issues behind the previous
// one (measured: ~5 cycles per instruction whatever the mix), so the time per step is the number of instructions on the
// main warp's path.  The interpreter walks ~1700 per step through a 55 KB loop body (profiles/r2_c1_fma_kernel.txt:
// 6 us per step).
//
// What the measurements said (profiles/r2_c1_warp_kernel_v*.txt), in order:
//   v1 weights in registers, inputs by warp shuffles, fully unrolled: 255 registers leave room for ~3 shuffles in flight,
//      the shuffle latency shows on every third FMA; the compiler re-selected the source register per chunk (3.5-4.8 us/step);
//   v3 weights in shared memory + shuffles in chunks of 8: 3 instructions per multiply-add plus address arithmetic,
//      1900 instructions per step (3.8 us/step);
//   v4 both operands by 16-byte shared-memory loads (lane j reads four weights of ITS output: rows padded to a stride
//      of 4 mod 8 floats are conflict-free; four activations arrive by one broadcast load): 1.5 instructions per
//      multiply-add, 1080 per step - still 3.0 us/step: the layers were no longer the bulk, Philox + Box-Muller, the
//      spline read, descriptor loads from the constant bank and re-derived shared-window bases were;
//   v5 moves the state-independent work to the helper warp and keeps shared addresses in registers (32-bit window
//      addresses, explicit ld.shared / st.shared): 2.35 us/step;
//   v6 (this file) unrolls the loop over mat-vecs so that descriptor fields are constant-bank operands, and issues all
//      16 operand loads of a 32-wide mat-vec before its FMAs (ld.volatile + FMA chains that start from the last pair
//      loaded: ptxas otherwise pairs every load with its first use): 2.17 us/step, 720 instructions per step on the
//      main warp.  Packed FFMA2 (half the FMA instructions) measured no faster: what is left is the dependent
//      latency chain load -> FMA chain -> reduction -> activation -> store of six layers, not issue slots.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "snsde_fma.cuh"
#include "snsde_warp.cuh"

namespace snsde {

constexpr int kActRow = 32;          // floats per activation row (lane j = feature j; unused lanes hold exact zeros)
constexpr int kBatch = 4;            // solver steps per hand-over between the helper and the main warp (the first
                                     // batch is pure start-up latency for the main warp: 4.6 us at 8 steps per batch)

// Shared memory through 32-bit window addresses kept in registers (plain pointers made the compiler re-derive the
// window base from SR_CgaCtaId inside the loops: ~100 cycles each time on a single warp).
__device__ __forceinline__ float lds_f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
// Operand loads of a mat-vec: ld.volatile keeps their program order in ptxas, which otherwise pairs every load with its
// first use (it schedules for an occupancy that hides the latency; here ONE warp runs per scheduler).
__device__ __forceinline__ float4 lds_f4_ordered(uint32_t a) {
  float4 v;
  asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ int2 lds_i2(uint32_t a) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// One solver step's records out of the staged tables (or, TS = false, out of global memory).
struct StepRec { float t0, h, sqrt_h, sin_t0, cos_t0, frac; int interval, emit_begin, emit_end; };
template <bool TS>
__device__ __forceinline__ StepRec read_step(uint32_t a_steps, const snsde_step* g_steps, int s) {
  StepRec r;
  if (TS) {
    const uint32_t a = a_steps + s * (uint32_t)sizeof(snsde_step);          // 40-byte records, 8-byte aligned
    const int2 v0 = lds_i2(a), v1 = lds_i2(a + 8), v2 = lds_i2(a + 16), v3 = lds_i2(a + 24), v4 = lds_i2(a + 32);
    r.t0 = __int_as_float(v0.x); r.h = __int_as_float(v0.y); r.sqrt_h = __int_as_float(v1.x); r.sin_t0 = __int_as_float(v1.y);
    r.cos_t0 = __int_as_float(v2.x); r.interval = v2.y; r.frac = __int_as_float(v3.x); r.emit_begin = v3.y; r.emit_end = v4.x;
  } else {
    const snsde_step st = g_steps[s];
    r.t0 = st.t0; r.h = st.h; r.sqrt_h = st.sqrt_h; r.sin_t0 = st.sin_t0; r.cos_t0 = st.cos_t0; r.interval = st.interval;
    r.frac = st.frac; r.emit_begin = st.emit_begin; r.emit_end = st.emit_end;
  }
  return r;
}

// NMV: mat-vec slots compiled in (the loop over mat-vecs is unrolled so that every descriptor field is a constant-bank
// operand of the instruction that uses it; slots beyond wp.n_mv are skipped by a uniform branch).
// TS: step and emit tables staged in shared memory.
template <int NMV, int R, bool TS>
__global__ void __launch_bounds__(256) snsde_warp_kernel(const FmaParams p, const WarpProg wp) {
  extern __shared__ __align__(16) float smem[];
  const Program& pg = p.prog;
  const TailOp& t = pg.tail;
  const int H = pg.H, C = pg.C, S = p.S;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, npairs = blockDim.x >> 6;
  const int pair = wid >> 1;
  const bool helper = (wid & 1) != 0;

  // ---- shared memory (floats): [weight image][pairs x activation rows][pairs x ring][tables] ----
  constexpr int kPlanes = 2 * R + 1;                        // per step: R rows of X(t), R rows of dW, the coefficient row
  const int wf = p.smem_w_floats, descf = 0;
  const int actf = kNumRowBufs * R * kActRow, ringf = 2 * kBatch * kPlanes * kActRow;
  for (int i = threadIdx.x; i < (wf >> 2); i += blockDim.x) cp_async16(smem + 4 * i, p.wimg + 4 * i);   // one round trip:
  cp_async_commit();                                        // every 16-byte copy of the image is in flight at once
  {
    float* act_all = smem + wf + descf;
    for (int i = threadIdx.x; i < npairs * actf; i += blockDim.x) act_all[i] = 0.f;
    if (TS) {
      int* tbl = reinterpret_cast<int*>(smem + wf + descf + npairs * (actf + ringf));
      const int* gs = reinterpret_cast<const int*>(p.steps);
      const int nsi = S * (int)(sizeof(snsde_step) / 4);
      for (int i = threadIdx.x; i < nsi; i += blockDim.x) tbl[i] = gs[i];
      const int* ge = reinterpret_cast<const int*>(p.emits);
      const int nei = wp.n_emits * (int)(sizeof(snsde_emit) / 4);
      for (int i = threadIdx.x; i < nei; i += blockDim.x) tbl[nsi + i] = ge[i];
    }
  }
  cp_async_wait_all();
  __syncthreads();                                          // the only CTA-wide barrier
  const uint32_t aW = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t aDesc = aW + 4u * wf;
  const uint32_t aAct = aDesc + 4u * (descf + pair * actf);
  const uint32_t aRing = aDesc + 4u * (descf + npairs * actf + pair * ringf);
  const uint32_t aSteps = aDesc + 4u * (descf + npairs * (actf + ringf));
  const uint32_t aEmits = aSteps + (uint32_t)S * (uint32_t)sizeof(snsde_step);
  auto read_emit = [&](int e) {
    snsde_emit em;
    if (TS) {
      const uint32_t a = aEmits + (uint32_t)e * (uint32_t)sizeof(snsde_emit);
      em.slot = __float_as_int(lds_f(a)); em.w_prev = lds_f(a + 4); em.w_curr = lds_f(a + 8);
    } else {
      em = p.emits[e];
    }
    return em;
  };

  const int row0 = (blockIdx.x * npairs + pair) * R;
  if (row0 >= p.B) return;                                  // whole pair idle
  auto grow = [&](int r) { return min(row0 + r, p.B - 1); };
  const bool jact = lane < H;
  const int n_batches = (S + kBatch - 1) / kBatch;
  const int bar_id = pair + 1;
  auto slot_of = [&](int b, int q) { return aRing + 4u * (((b & 1) * kBatch + q) * kPlanes * kActRow); };

  if (helper) {
    // =========================== helper warp: everything that does not depend on the state ===========================
    const bool ctl = pg.uses_control && lane < C;
    const bool vtab = t.coef_src == CO_VBUF && jact;
    const float* crow[R];
#pragma unroll
    for (int r = 0; r < R; ++r) crow[r] = p.coeffs + (size_t)grow(r) * p.coeff_row_stride + lane;
    for (int b = 0; b < n_batches; ++b) {
      const int s0 = b * kBatch;
      // phase A: every global load of the batch in flight at once
      float ca[kBatch][R], cb[kBatch][R], cc[kBatch][R], cd[kBatch][R], tw[kBatch][R], vc[kBatch], fr[kBatch], sq[kBatch];
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const int s = min(s0 + q, S - 1);
        const StepRec st = read_step<TS>(aSteps, p.steps, s);
        fr[q] = st.frac; sq[q] = st.sqrt_h;
        vc[q] = vtab ? __ldg(p.vtab + (size_t)s * H + lane) : 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          ca[q][r] = cb[q][r] = cc[q][r] = cd[q][r] = 0.f;
          if (ctl) {
            const float* src = crow[r] + (size_t)st.interval * 4 * C;
            ca[q][r] = __ldg(src); cb[q][r] = __ldg(src + C); cc[q][r] = __ldg(src + 2 * C); cd[q][r] = __ldg(src + 3 * C);
          }
          tw[q][r] = (p.dW != nullptr && jact) ? __ldg(p.dW + ((size_t)s * p.B + grow(r)) * H + lane) : 0.f;
        }
      }
      // phase B: X(t) = a + (b + (two_c/2 + three_d*frac/3)*frac)*frac (torchcde op order), increments, coefficient
#pragma unroll 2
      for (int q = 0; q < kBatch; ++q) {
        const int s = s0 + q;
        if (s < S) {
          const uint32_t slot = slot_of(b, q);
          float nrm[4];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float inner = 0.5f * cc[q][r] + __fdiv_rn(cd[q][r] * fr[q], 3.0f);
            inner = cb[q][r] + inner * fr[q];
            sts_f(slot + 4u * (r * kActRow + lane), ca[q][r] + inner * fr[q]);          // lanes >= C: exact zero
            float dw = tw[q][r];
            if (p.dW == nullptr) {
              const unsigned long long gb = p.row_offset + (unsigned long long)(row0 + r);
              if (R == 1) {
                dw = __fmul_rn(philox_normal1(p.seed, (uint32_t)lane, gb, (uint32_t)s), sq[q]);
              } else {
                if (r == 0 || (gb & 3ull) == 0ull) philox_normals4(p.seed, (uint32_t)lane, (uint32_t)(gb >> 2), (uint32_t)s, nrm);
                dw = __fmul_rn(pick4(nrm, (int)(gb & 3ull)), sq[q]);
              }
            }
            sts_f(slot + 4u * ((R + r) * kActRow + lane), dw);
          }
          sts_f(slot + 4u * (2 * R * kActRow + lane), vc[q]);
        }
      }
      pair_sync(bar_id);                                    // batch b handed over (and batch b-1 consumed)
    }
    return;
  }

  // =============================== main warp: the dependent chain of every step ====================================
  float y[R], yprev[R];
  int myslot[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    y[r] = jact ? p.y0[(size_t)grow(r) * H + lane] : 0.f;
    yprev[r] = y[r];
    myslot[r] = p.row_slot ? p.row_slot[grow(r)] : -1;
    if (jact) sts_f(aAct + 4u * ((BUF_Y * R + r) * kActRow + lane), y[r]);
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
    snsde_emit em = read_emit(e);
    em.w_prev = 0.f; em.w_curr = 1.f;
    emit(em);
  }
  const float coef_fixed = (t.coef_src == CO_IMG) ? lds_f(aW + 4u * (wp.coef_off + lane)) : t.coef_scalar;
  __syncwarp();

  for (int b = 0; b < n_batches; ++b) {
    pair_sync(bar_id);                                      // batch b is in the ring
#pragma unroll 1
    for (int q = 0; q < kBatch; ++q) {
      const int s = b * kBatch + q;
      if (s >= S) break;
      const StepRec st = read_step<TS>(aSteps, p.steps, s);
      const uint32_t slot = slot_of(b, q);

      // ---- the dense program ----
      float drift[R], a0[R], a1[R], a2[R], a3[R];
#pragma unroll
      for (int r = 0; r < R; ++r) { drift[r] = 0.f; a0[r] = a1[r] = a2[r] = a3[r] = 0.f; }
#pragma unroll
      for (int i = 0; i < NMV; ++i) {
        if (i < wp.n_mv) {                                   // uniform
          const WarpMv& m = wp.mv[i];
          const int jc = min(lane, m.N - 1);                // lanes beyond N recompute row N-1 (zeroed below)
          const uint32_t wa = aW + 4u * (m.w_off + jc * m.stride);
          const uint32_t xa = m.src == BUF_X ? slot : aAct + 4u * (m.src * R * kActRow);
          if (m.flags & kMvFirst) {
            float init = lds_f(aW + 4u * (m.b_off + lane));
            if (m.flags & kMvSinCos)
              init = fmaf(st.cos_t0, lds_f(aW + 4u * (m.tw_off + 32 + lane)), fmaf(st.sin_t0, lds_f(aW + 4u * (m.tw_off + lane)), init));
#pragma unroll
            for (int r = 0; r < R; ++r) { a0[r] = init; a1[r] = a2[r] = a3[r] = 0.f; }
          }
          if (m.n8 == 4) {
            // 32 inputs: all 8 + 8R operand loads are issued first; the FMA chains start from the LAST pair loaded, so
            // nothing can be scheduled between the loads and one shared-memory latency is paid per mat-vec, not per pair
            float4 w[8], x[R][8];
#pragma unroll
            for (int c = 0; c < 8; ++c) w[c] = lds_f4_ordered(wa + 16u * c);
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
              for (int c = 0; c < 8; ++c) x[r][c] = lds_f4_ordered(xa + 4u * (r * kActRow) + 16u * c);
#pragma unroll
            for (int c = 7; c >= 0; --c)
#pragma unroll
              for (int r = R - 1; r >= 0; --r) {
                a0[r] = fmaf(x[r][c].x, w[c].x, a0[r]); a1[r] = fmaf(x[r][c].y, w[c].y, a1[r]);
                a2[r] = fmaf(x[r][c].z, w[c].z, a2[r]); a3[r] = fmaf(x[r][c].w, w[c].w, a3[r]);
              }
          } else {
#pragma unroll 1
            for (int c = m.n8 - 1; c >= 0; --c) {
              float4 w0 = lds_f4_ordered(wa + 32u * c), w1 = lds_f4_ordered(wa + 32u * c + 16u), x0[R], x1[R];
#pragma unroll
              for (int r = 0; r < R; ++r) {
                x0[r] = lds_f4_ordered(xa + 4u * (r * kActRow) + 32u * c);
                x1[r] = lds_f4_ordered(xa + 4u * (r * kActRow) + 32u * c + 16u);
              }
#pragma unroll
              for (int r = R - 1; r >= 0; --r) {
                a0[r] = fmaf(x1[r].x, w1.x, a0[r]); a1[r] = fmaf(x1[r].y, w1.y, a1[r]);
                a2[r] = fmaf(x1[r].z, w1.z, a2[r]); a3[r] = fmaf(x1[r].w, w1.w, a3[r]);
                a0[r] = fmaf(x0[r].x, w0.x, a0[r]); a1[r] = fmaf(x0[r].y, w0.y, a1[r]);
                a2[r] = fmaf(x0[r].z, w0.z, a2[r]); a3[r] = fmaf(x0[r].w, w0.w, a3[r]);
              }
            }
          }
          if (m.flags & kMvLast) {
            if (m.dst == kWarpDstDrift) {
#pragma unroll
              for (int r = 0; r < R; ++r) drift[r] = (a0[r] + a1[r]) + (a2[r] + a3[r]);
            } else {
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const float o = act_apply((a0[r] + a1[r]) + (a2[r] + a3[r]), m.act);
                sts_f(aAct + 4u * ((m.dst * R + r) * kActRow + lane), lane < m.N ? o : 0.f);  // unused lanes of every row stay zero
              }
              __syncwarp();
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
          float coef = coef_fixed;
          if (t.coef_src == CO_RBUF) coef = lds_f(aAct + 4u * ((t.coef_ref * R + r) * kActRow + lane));
          else if (t.coef_src == CO_VBUF) coef = lds_f(slot + 4u * (2 * R * kActRow + lane));
          float g, dgdy;
          diffusion_eval<false>(t, coef, y[r], st.t0, g, dgdy);
          const float dw = lds_f(slot + 4u * ((R + r) * kActRow + lane));
          float yn = __fadd_rn(__fadd_rn(y[r], __fmul_rn(d, st.h)), __fmul_rn(g, dw));
          if (t.milstein) {
            const float qv = __fmul_rn(dw, dw) - st.h;
            yn = __fadd_rn(yn, 0.5f * ((g * qv) * dgdy));
          }
          yprev[r] = y[r];
          y[r] = yn;
          sts_f(aAct + 4u * ((BUF_Y * R + r) * kActRow + lane), yn);
        }
      }
      for (int e = st.emit_begin; e < st.emit_end; ++e) emit(read_emit(e));
      __syncwarp();                                         // this step's reads precede the next step's writes
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
// Flattens the per-row ops into mat-vecs and lays their weights out for the kernel: per mat-vec N rows of
// `stride` = round8(K) + 4 floats (row j = nn.Linear row j, columns [col, col+K), zero-padded), then for the first
// mat-vec of an output a 32-float bias row and, with time features, the sin / cos weight rows.
bool warp_build(const Program& pg, int method, const float* blob, const float* fma_img, WarpProg& wp, std::vector<float>& img) {
  memset(&wp, 0, sizeof(wp));
  img.clear();
  if (method != SNSDE_METHOD_EULER && method != SNSDE_METHOD_MILSTEIN) return false;      // SRK stages: interpreter kernel
  const TailOp& t = pg.tail;
  if (t.latent) return false;
  if (t.milstein && t.vjp_kind != 0) return false;                                         // full vjp through noise_y
  if (std::max(std::max(pg.H, pg.HH), pg.uses_control ? pg.C : 0) > 32) return false;
  if (t.coef_src == CO_RBUF && (t.coef_ref < 0 || t.coef_ref >= kNumRowBufs)) return false;
  auto row32 = [&](auto value_of, int n) {
    const int off = (int)img.size();
    for (int j = 0; j < 32; ++j) img.push_back(j < n ? value_of(j) : 0.f);
    return off;
  };
  for (int o = 0; o < pg.n_ops; ++o) {
    const DenseOp& op = pg.ops[o];
    if (op.vec) continue;                                   // tabulated per step by vec_tables_kernel
    if (op.src < 0 || op.src >= kNumRowBufs || op.K > 32 || op.K < 1 || op.N > 32 || op.N < 1 || op.tmode == TM_RAW || op.g_w < 0)
      return false;
    if (op.src2 >= kNumRowBufs || (op.src2 >= 0 && (op.K2 > 32 || op.K2 < 1))) return false;
    if (!op.final_drift && (op.dst < 0 || op.dst >= kNumRowBufs || op.dst == BUF_X || op.dst == BUF_Y)) return false;
    const int n = op.src2 >= 0 ? 2 : 1;
    if (wp.n_mv + n > kWarpMaxMv) return false;
    const float* Wl = blob + op.g_w;                         // nn.Linear weight [N][g_ldw]
    for (int part = 0; part < n; ++part) {
      WarpMv& m = wp.mv[wp.n_mv++];
      const int K = part == 0 ? op.K : op.K2, col = part == 0 ? op.g_col : op.g_col2;
      m.src = part == 0 ? op.src : op.src2;
      m.n8 = (K + 7) / 8;
      m.N = op.N;
      m.stride = m.n8 * 8 + 4;
      m.flags = (part == 0 ? kMvFirst : 0) | (part == n - 1 ? kMvLast : 0) | (part == 0 && op.tmode == TM_SINCOS ? kMvSinCos : 0);
      m.dst = op.final_drift ? kWarpDstDrift : op.dst;
      m.act = op.act;
      m.w_off = (int)img.size();
      for (int j = 0; j < op.N; ++j)
        for (int k = 0; k < m.stride; ++k) img.push_back(k < K ? Wl[(size_t)j * op.g_ldw + col + k] : 0.f);
      m.b_off = m.tw_off = 0;
      if (part == 0) {
        m.b_off = row32([&](int j) { return op.g_b >= 0 ? blob[op.g_b + j] : 0.f; }, op.N);
        if (op.tmode == TM_SINCOS) {                         // time features are columns 0 (sin) and 1 (cos) of the weight
          m.tw_off = row32([&](int j) { return Wl[(size_t)j * op.g_ldw + 0]; }, op.N);
          row32([&](int j) { return Wl[(size_t)j * op.g_ldw + 1]; }, op.N);
        }
      }
    }
  }
  wp.coef_off = 0;
  if (t.coef_src == CO_IMG) wp.coef_off = row32([&](int j) { return fma_img[t.coef_ref + j]; }, pg.H);
  while (img.size() & 3) img.push_back(0.f);
  return wp.n_mv > 0;
}

size_t warp_smem_bytes(int img_floats, int n_mv, int pairs, int R, int S, int n_emits, bool tables) {
  (void)n_mv;
  size_t f = (size_t)img_floats + (size_t)pairs * ((size_t)kNumRowBufs * R * kActRow + (size_t)2 * kBatch * (2 * R + 1) * kActRow);
  size_t b = f * 4;
  if (tables) b += (size_t)S * sizeof(snsde_step) + (size_t)n_emits * sizeof(snsde_emit);
  return b;
}

template <int R>
static cudaError_t warp_launch_r(const FmaParams& p, WarpProg wp, int n_emits, int num_sms, int smem_optin, cudaStream_t stream) {
  const int n_groups = (p.B + R - 1) / R;
  // small batches: one pair of warps per CTA so that every row group gets an SM of its own; otherwise up to 4 pairs
  const int pairs = std::max(1, std::min(4, (n_groups + num_sms - 1) / num_sms));
  const int grid = (n_groups + pairs - 1) / pairs;
  const size_t t_bytes = (size_t)p.S * sizeof(snsde_step) + (size_t)n_emits * sizeof(snsde_emit);
  const bool ts = t_bytes <= 64 * 1024 && warp_smem_bytes(p.wimg_floats, wp.n_mv, pairs, R, p.S, n_emits, true) <= (size_t)smem_optin;
  const size_t smem = warp_smem_bytes(p.wimg_floats, wp.n_mv, pairs, R, p.S, n_emits, ts);
  if (smem > (size_t)smem_optin) return cudaErrorInvalidValue;
  FmaParams q = p;
  q.smem_w_floats = p.wimg_floats;
  wp.n_emits = n_emits;
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, pairs * 64, smem, stream>>>(q, wp);
    return cudaGetLastError();
  };
  auto pick = [&](auto nmv) -> cudaError_t {
    constexpr int N = decltype(nmv)::value;
    return ts ? go(snsde_warp_kernel<N, R, true>) : go(snsde_warp_kernel<N, R, false>);
  };
  if (wp.n_mv <= 4) return pick(std::integral_constant<int, 4>());
  if (wp.n_mv <= 6) return pick(std::integral_constant<int, 6>());
  if (wp.n_mv <= 8) return pick(std::integral_constant<int, 8>());
  return pick(std::integral_constant<int, kWarpMaxMv>());
}

// `p.wimg` / `p.wimg_floats` must describe the WARP image (warp_build), not the interpreter's.
// One row per pair: measured on the tutorial function (H = 32, 1xB200, SDE-steps/s, this kernel vs the interpreter
// kernel) 64 rows 3.2e7 / 1.0e7, 512 rows 2.0e8 / 6.0e7, 2048 rows 4.0e8 / 2.4e8, 8192 rows 3.4e8 / 5.8e8, 32768 rows
// 3.5e8 / 1.0e9 - a latency kernel: it wins while rows are scarce, and the interpreter's 8-row groups (each weight
// read feeds 8 FMAs, 15 groups share one staged image) win once the machine is full.  The caller switches at
// kWarpMaxRows; a two-rows-per-pair variant was measured slower than both (165 registers) and is not instantiated.
cudaError_t warp_launch(const FmaParams& p, const WarpProg& wp, int n_emits, int num_sms, int smem_optin, cudaStream_t stream) {
  return warp_launch_r<1>(p, wp, n_emits, num_sms, smem_optin, stream);
}

}  // namespace snsde
