// Warp-owned fp32 kernel for hidden sizes <= 32 (the north star's "warp-shuffle / FMA path for hidden < 64").
// A ROW GROUP = R batch rows (R = 1 is what ships) owned end to end by a PAIR of warps: a main warp that runs the dependent chain of a solver
// step (the layers, the SDE update, the outputs) and a helper warp that prepares, one batch of steps ahead, everything
// that does not depend on the state: X(t) of every row, the Brownian increments, the row-independent diffusion
// coefficient.  Lane j = feature j; the SDE state lives in the main warp's registers.  The pair meets at one named
// barrier per batch of 4 steps; inside a step the main warp synchronises with nobody (__syncwarp only).
//
// Replaces, like snsde_fma.cu, the Python step loop of torchsde.sdeint (Euler.step / Milstein.step /
// SRK.diagonal_or_scalar_step) with the per-step Diffusion_model.f/g evaluation
// (/root/reference/benchmark_classification/models_sde/neuralsde.py:295-307), torchcde.CubicSpline.evaluate (:296), the
// tutorial NeuralLSDEFunc (notebook cell 7) and LatentSDE.f_aug/g_aug (torch-ists/.../latent_sde.py:74-90) - for the shapes where the
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
//      Then: hand-over batches of 4 steps (the first batch is start-up latency), cp.async staging of the image, the
//      coefficient table cached across solves: 101-107 us per c1 solve (315 us on the interpreter kernel).
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

// One evaluation point of an SRK step out of the staged point table (or global memory).
struct PointRec { float t, sin_t, cos_t, frac; int interval; };
template <bool TS>
__device__ __forceinline__ PointRec read_point(uint32_t a_points, const snsde_point* g_points, int idx) {
  PointRec r;
  if (TS) {
    const uint32_t a = a_points + idx * (uint32_t)sizeof(snsde_point);        // 20-byte records
    r.t = lds_f(a); r.sin_t = lds_f(a + 4); r.cos_t = lds_f(a + 8); r.frac = lds_f(a + 12); r.interval = __float_as_int(lds_f(a + 16));
  } else {
    const snsde_point pt = g_points[idx];
    r.t = pt.t; r.sin_t = pt.sin_t; r.cos_t = pt.cos_t; r.frac = pt.frac; r.interval = pt.interval;
  }
  return r;
}

// NMV: mat-vec slots compiled in (the loop over mat-vecs is unrolled so that every descriptor field is a constant-bank
// operand of the instruction that uses it; slots beyond wp.n_mv are skipped by a uniform branch).
// TS: step / emit / point tables staged in shared memory.
// METHOD 0: Euler / Milstein (one pass over the program per step).  METHOD 1: SRK (torchsde SRK.diagonal_or_scalar_step,
// SRID2; same stage algebra as snsde_fma.cu): six passes per step - drift + diffusion at the step's state, then drift at
// the stage states H0_1, H0_2 and the per-row diffusion network at H1_1, H1_2, H1_3 - through ONE copy of the mat-vec code.
template <int NMV, int R, bool TS, int METHOD>
__global__ void __launch_bounds__(256) snsde_warp_kernel(const FmaParams p, const WarpProg wp) {
  extern __shared__ __align__(16) float smem[];
  const Program& pg = p.prog;
  const TailOp& t = pg.tail;
  const int H = pg.H, C = pg.C, S = p.S;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, npairs = blockDim.x >> 6;
  const int pair = wid >> 1;
  const bool helper = (wid & 1) != 0;
  constexpr bool SRK = METHOD == 1;
  constexpr int NX = SRK ? 3 : 1;                           // X(t) rows per batch row and step: t0 | t0, t0+h, t0+h/2
  constexpr int NW = SRK ? 2 : 1;                           // increment rows: dW | dW, U
  constexpr int NV = SRK ? kSrkGPoints : 1;                 // coefficient rows: t0 | t0, t0+h/4, t0+h
  constexpr int kPlanes = (NX + NW) * R + NV;

  // ---- shared memory (floats): [weight image][pairs x activation rows][pairs x ring][step, emit, point tables] ----
  const int wf = p.smem_w_floats;
  const int actf = kNumRowBufs * R * kActRow, ringf = 2 * kBatch * kPlanes * kActRow;
  for (int i = threadIdx.x; i < (wf >> 2); i += blockDim.x) cp_async16(smem + 4 * i, p.wimg + 4 * i);   // one round trip:
  cp_async_commit();                                        // every 16-byte copy of the image is in flight at once
  const int nsi = S * (int)(sizeof(snsde_step) / 4), nei = wp.n_emits * (int)(sizeof(snsde_emit) / 4);
  {
    float* act_all = smem + wf;
    for (int i = threadIdx.x; i < npairs * actf; i += blockDim.x) act_all[i] = 0.f;
    if (TS) {
      int* tbl = reinterpret_cast<int*>(smem + wf + npairs * (actf + ringf));
      const int* gs = reinterpret_cast<const int*>(p.steps);
      for (int i = threadIdx.x; i < nsi; i += blockDim.x) tbl[i] = gs[i];
      const int* ge = reinterpret_cast<const int*>(p.emits);
      for (int i = threadIdx.x; i < nei; i += blockDim.x) tbl[nsi + i] = ge[i];
      if (SRK) {
        const int* gp = reinterpret_cast<const int*>(p.points);
        const int npi = S * kSrkPoints * (int)(sizeof(snsde_point) / 4);
        for (int i = threadIdx.x; i < npi; i += blockDim.x) tbl[nsi + nei + i] = gp[i];
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();                                          // the only CTA-wide barrier
  const uint32_t aW = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t aAct = aW + 4u * (wf + pair * actf);
  const uint32_t aRing = aW + 4u * (wf + npairs * actf + pair * ringf);
  const uint32_t aSteps = aW + 4u * (wf + npairs * (actf + ringf));
  const uint32_t aEmits = aSteps + 4u * nsi;
  const uint32_t aPoints = aEmits + 4u * nei;
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
  // planes of a ring slot: X rows [i][r], increment rows [w][r], coefficient rows [v]
  auto x_plane = [&](int i, int r) { return 4u * ((i * R + r) * kActRow); };
  auto w_plane = [&](int w, int r) { return 4u * (((NX + w) * R + r) * kActRow); };
  auto v_plane = [&](int v) { return 4u * (((NX + NW) * R + v) * kActRow); };

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
      float ca[kBatch][NX][R], cb[kBatch][NX][R], cc[kBatch][NX][R], cd[kBatch][NX][R], tw[kBatch][NW][R], vc[kBatch][NV];
      float fr[kBatch][NX], sq[kBatch], hh[kBatch];
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const int s = min(s0 + q, S - 1);
        const StepRec st = read_step<TS>(aSteps, p.steps, s);
        sq[q] = st.sqrt_h; hh[q] = st.h;
        int interval[NX];
        fr[q][0] = st.frac; interval[0] = st.interval;
        if constexpr (SRK) {                                // drift evaluation points: t0, t0 + h, t0 + h/2  (point table 0, 3, 2)
          const PointRec p1 = read_point<TS>(aPoints, p.points, s * kSrkPoints + 3), p2 = read_point<TS>(aPoints, p.points, s * kSrkPoints + 2);
          const PointRec p0 = read_point<TS>(aPoints, p.points, s * kSrkPoints + 0);
          fr[q][0] = p0.frac; interval[0] = p0.interval;
          fr[q][NX - 2] = p1.frac; interval[NX - 2] = p1.interval;
          fr[q][NX - 1] = p2.frac; interval[NX - 1] = p2.interval;
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) vc[q][v] = vtab ? __ldg(p.vtab + ((size_t)s * NV + v) * H + lane) : 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            ca[q][i][r] = cb[q][i][r] = cc[q][i][r] = cd[q][i][r] = 0.f;
            if (ctl) {
              const float* src = crow[r] + (size_t)interval[i] * 4 * C;
              ca[q][i][r] = __ldg(src); cb[q][i][r] = __ldg(src + C); cc[q][i][r] = __ldg(src + 2 * C); cd[q][i][r] = __ldg(src + 3 * C);
            }
          }
          const size_t wi = ((size_t)s * p.B + grow(r)) * H + lane;
          tw[q][0][r] = (p.dW != nullptr && jact) ? __ldg(p.dW + wi) : 0.f;
          if (SRK) tw[q][NW - 1][r] = (p.dW != nullptr && jact) ? __ldg(p.dU + wi) : 0.f;
        }
      }
      // phase B: X(t) = a + (b + (two_c/2 + three_d*frac/3)*frac)*frac (torchcde op order), increments, coefficient
#pragma unroll 2
      for (int q = 0; q < kBatch; ++q) {
        const int s = s0 + q;
        if (s < S) {
          const uint32_t slot = slot_of(b, q);
          float nrm[4], nrmu[4];
#pragma unroll
          for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int i = 0; i < NX; ++i) {
              float inner = 0.5f * cc[q][i][r] + __fdiv_rn(cd[q][i][r] * fr[q][i], 3.0f);
              inner = cb[q][i][r] + inner * fr[q][i];
              sts_f(slot + x_plane(i, r) + 4u * lane, ca[q][i][r] + inner * fr[q][i]);     // lanes >= C: exact zero
            }
            float dw = tw[q][0][r], du = SRK ? tw[q][NW - 1][r] : 0.f;
            if (p.dW == nullptr) {
              const unsigned long long gb = p.row_offset + (unsigned long long)(row0 + r);
              if (R == 1 && !SRK) {
                dw = __fmul_rn(philox_normal1(p.seed, (uint32_t)lane, gb, (uint32_t)s), sq[q]);
              } else {
                if (r == 0 || (gb & 3ull) == 0ull) {
                  philox_normals4(p.seed, (uint32_t)lane, (uint32_t)(gb >> 2), (uint32_t)s, nrm);
                  if (SRK) philox_normals4_u(p.seed, (uint32_t)lane, (uint32_t)(gb >> 2), (uint32_t)s, nrmu);
                }
                dw = __fmul_rn(pick4(nrm, (int)(gb & 3ull)), sq[q]);
                if (SRK) du = levy_U(dw, pick4(nrmu, (int)(gb & 3ull)), hh[q], sq[q]);
              }
            }
            sts_f(slot + w_plane(0, r) + 4u * lane, dw);
            if (SRK) sts_f(slot + w_plane(NW - 1, r) + 4u * lane, du);
          }
#pragma unroll
          for (int v = 0; v < NV; ++v) sts_f(slot + v_plane(v) + 4u * lane, vc[q][v]);
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
  float coef_fixed = (t.coef_src == CO_IMG) ? lds_f(aW + 4u * (wp.coef_off + lane)) : t.coef_scalar;
  if (t.latent && lane == H - 1) coef_fixed = 0.f;          // g_aug: no noise on the KL accumulator (latent_sde.py:84-90)
  const bool per_row_g = t.coef_src == CO_RBUF;
  auto set_state = [&](int r, float v) { if (jact) sts_f(aAct + 4u * ((BUF_Y * R + r) * kActRow + lane), v); };
  __syncwarp();

  for (int b = 0; b < n_batches; ++b) {
    pair_sync(bar_id);                                      // batch b is in the ring
#pragma unroll 1
    for (int q = 0; q < kBatch; ++q) {
      const int s = b * kBatch + q;
      if (s >= S) break;
      const StepRec st = read_step<TS>(aSteps, p.steps, s);
      const uint32_t slot = slot_of(b, q);
      const float h = st.h, sqrt_h = st.sqrt_h, rdt = SRK ? __fdiv_rn(1.0f, h) : 0.f;
      // SRK stage registers (see snsde_fma.cu for the tableau): only live when METHOD == 1
      float y0r[R], dwv[R], ik0[R], f0[R], g0[R], f1[R], g1[R], f2[R], g2[R], tmp[R], hs[R];
#pragma unroll
      for (int r = 0; r < R; ++r) { y0r[r] = y[r]; hs[r] = y[r]; tmp[r] = y[r]; dwv[r] = ik0[r] = f0[r] = g0[r] = f1[r] = g1[r] = f2[r] = g2[r] = 0.f; }

      constexpr int NPASS = SRK ? 6 : 1;
#pragma unroll 1
      for (int pass = 0; pass < NPASS; ++pass) {
        // what this pass evaluates, where, and on which X(t) row:
        //   0: f, g at t0 on the step's state | 1: f at t0+h on H0_1 | 2: g at t0+h/4 on H1_1 | 3: f at t0+h/2 on H0_2
        //   4: g at t0+h on H1_2              | 5: g at t0+h/4 on H1_3
        const bool want_f = !SRK || pass == 0 || pass == 1 || pass == 3;
        const bool want_g = !SRK || pass == 0 || pass == 2 || pass >= 4;
        float tt = st.t0, tf_sin = st.sin_t0, tf_cos = st.cos_t0;
        int xi = 0, vq = 0;
        if (SRK) {
          const int pi = pass == 0 ? 0 : (pass == 1 || pass == 4 ? 3 : (pass == 3 ? 2 : 1));       // point table index
          const PointRec pt = read_point<TS>(aPoints, p.points, s * kSrkPoints + pi);
          tt = pt.t; tf_sin = pt.sin_t; tf_cos = pt.cos_t;
          xi = pass == 1 ? 1 : (pass == 3 ? 2 : 0);
          vq = pi == 0 ? 0 : (pi == 1 ? 1 : 2);                                                     // t0, t0+h/4, t0+h
        }
        const bool run_nets = want_f || per_row_g;            // a pass that only needs an elementwise g runs no mat-vec

        // ---- the dense program ----
        float drift[R], a0[R], a1[R], a2[R], a3[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { drift[r] = 0.f; a0[r] = a1[r] = a2[r] = a3[r] = 0.f; }
        if (run_nets) {
#pragma unroll
          for (int i = 0; i < NMV; ++i) {
            if (i < wp.n_mv) {                                 // uniform
              const WarpMv& m = wp.mv[i];
              const bool is_g = (m.flags & kMvDiff) != 0;
              if (SRK && (is_g ? !want_g : !want_f)) continue;
              const int jc = min(lane, m.N - 1);              // lanes beyond N recompute row N-1 (zeroed below)
              const uint32_t wa = aW + 4u * (m.w_off + jc * m.stride);
              const uint32_t xa = m.src == BUF_X ? slot + x_plane(xi, 0) : aAct + 4u * (m.src * R * kActRow);
              if (m.flags & kMvFirst) {
                float init = lds_f(aW + 4u * (m.b_off + lane));
                if (m.flags & kMvSinCos)
                  init = fmaf(tf_cos, lds_f(aW + 4u * (m.tw_off + 32 + lane)), fmaf(tf_sin, lds_f(aW + 4u * (m.tw_off + lane)), init));
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
                  for (int r = 0; r < R; ++r) drift[r] = lane < m.N ? (a0[r] + a1[r]) + (a2[r] + a3[r]) : 0.f;
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
        }

        // drift value at the state this pass read: geometric term, tanh clip, LatentSDE's KL channel (latent_sde.py:77-82:
        // 0.5 * sum_j ((f_j - theta (mu - y_j)) / stable(sigma))^2 over the latent features, into feature H-1)
        float fv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float d = drift[r];
          if (want_f) {
            if (t.geometric) d = d * tanhf(hs[r]);
            if (t.clip_drift) d = tanhf(d);
            if (t.latent) {
              float u = 0.f;
              if (lane < H - 1) u = __fdiv_rn(d - t.lat_theta * (t.lat_mu - hs[r]), t.lat_div);
              float part = u * u;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
              if (lane == H - 1) d = 0.5f * part;
            }
          }
          fv[r] = d;
        }
        // diffusion value at the state this pass read
        float gv[R], dgv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          gv[r] = 0.f; dgv[r] = 0.f;
          if (want_g && jact) {
            float coef = coef_fixed;
            if (per_row_g) coef = lds_f(aAct + 4u * ((t.coef_ref * R + r) * kActRow + lane));
            else if (t.coef_src == CO_VBUF) coef = lds_f(slot + v_plane(vq) + 4u * lane);
            const float ys = SRK ? (pass == 0 ? y0r[r] : tmp[r]) : y[r];
            diffusion_eval<false>(t, coef, ys, tt, gv[r], dgv[r]);
          }
        }

        if (!SRK) {
          // ---- Euler / Milstein update (same arithmetic as snsde_fma.cu) ----
          if (jact) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float dw = lds_f(slot + w_plane(0, r) + 4u * lane);
              float yn = __fadd_rn(__fadd_rn(y[r], __fmul_rn(fv[r], h)), __fmul_rn(gv[r], dw));
              if (t.milstein) {
                const float qv = __fmul_rn(dw, dw) - h;
                yn = __fadd_rn(yn, 0.5f * ((gv[r] * qv) * dgv[r]));
              }
              yprev[r] = y[r];
              y[r] = yn;
              hs[r] = yn;
              set_state(r, yn);
            }
          }
        } else if (jact) {
          // ---- SRK stage algebra (SRID2; coefficients written out as in snsde_fma.cu) ----
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (pass == 0) {
              dwv[r] = lds_f(slot + w_plane(0, r) + 4u * lane);
              ik0[r] = lds_f(slot + w_plane(NW - 1, r) + 4u * lane);
              f0[r] = fv[r]; g0[r] = gv[r];
              hs[r] = y0r[r] + (1.0f * f0[r]) * h + ((0.0f * g0[r]) * ik0[r]) * rdt;               // H0_1
              tmp[r] = y0r[r] + (0.25f * f0[r]) * h + (-0.5f * g0[r]) * sqrt_h;                      // H1_1
              set_state(r, hs[r]);
            } else if (pass == 1) {
              f1[r] = fv[r];
              set_state(r, tmp[r]);
            } else if (pass == 2) {
              g1[r] = gv[r];
              float a = y0r[r] + (0.25f * f0[r]) * h + ((1.0f * g0[r]) * ik0[r]) * rdt;
              a = a + (0.25f * f1[r]) * h + ((0.5f * g1[r]) * ik0[r]) * rdt;                         // H0_2
              float bq = y0r[r] + (1.0f * f0[r]) * h + (1.0f * g0[r]) * sqrt_h;
              bq = bq + (0.0f * f1[r]) * h + (0.0f * g1[r]) * sqrt_h;                                // H1_2
              hs[r] = a; tmp[r] = bq;
              set_state(r, a);
            } else if (pass == 3) {
              f2[r] = fv[r];
              set_state(r, tmp[r]);
            } else if (pass == 4) {
              g2[r] = gv[r];
              float bq = y0r[r] + (0.0f * f0[r]) * h + (2.0f * g0[r]) * sqrt_h;
              bq = bq + (0.0f * f1[r]) * h + (-1.0f * g1[r]) * sqrt_h;
              bq = bq + (0.25f * f2[r]) * h + (0.5f * g2[r]) * sqrt_h;                               // H1_3
              tmp[r] = bq;
              set_state(r, bq);
            } else {
              const float g3 = gv[r];
              const float I_k = dwv[r];
              const float I_kk = (I_k * I_k - h) * 0.5f;
              const float I_kkk = (I_k * I_k * I_k - 3.0f * h * I_k) * (1.0f / 6.0f);
              const float c0 = I_kk / sqrt_h, c1 = ik0[r] * rdt, c2 = I_kkk * rdt;
              const float gw0 = -1.0f * I_k + 1.0f * c0 + 2.0f * c1 + -2.0f * c2;
              const float gw1 = (4.0f / 3.0f) * I_k + (-4.0f / 3.0f) * c0 + (-4.0f / 3.0f) * c1 + (5.0f / 3.0f) * c2;
              const float gw2 = (2.0f / 3.0f) * I_k + (1.0f / 3.0f) * c0 + (-2.0f / 3.0f) * c1 + (-2.0f / 3.0f) * c2;
              const float gw3 = c2;
              float yn = y0r[r] + ((1.0f / 6.0f) * f0[r]) * h + gw0 * g0[r];
              yn = yn + ((1.0f / 6.0f) * f1[r]) * h + gw1 * g1[r];
              yn = yn + ((2.0f / 3.0f) * f2[r]) * h + gw2 * g2[r];
              yn = yn + gw3 * g3;
              yprev[r] = y0r[r];
              y[r] = yn;
              set_state(r, yn);
            }
          }
        }
        __syncwarp();                                         // this pass's reads precede the next pass's (step's) writes
      }
      for (int e = st.emit_begin; e < st.emit_end; ++e) emit(read_emit(e));
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
  if (method != SNSDE_METHOD_EULER && method != SNSDE_METHOD_MILSTEIN && method != SNSDE_METHOD_SRK) return false;
  const TailOp& t = pg.tail;
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
      m.flags = (part == 0 ? kMvFirst : 0) | (part == n - 1 ? kMvLast : 0) | (part == 0 && op.tmode == TM_SINCOS ? kMvSinCos : 0) |
                (op.part == 1 ? kMvDiff : 0);
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

static size_t warp_table_bytes(int S, int n_emits, bool srk) {
  return (size_t)S * sizeof(snsde_step) + (size_t)n_emits * sizeof(snsde_emit) + (srk ? (size_t)S * kSrkPoints * sizeof(snsde_point) : 0);
}
size_t warp_smem_bytes(int img_floats, int pairs, int R, int S, int n_emits, bool tables, bool srk) {
  const int planes = srk ? (3 + 2) * R + kSrkGPoints : 2 * R + 1;
  size_t f = (size_t)img_floats + (size_t)pairs * ((size_t)kNumRowBufs * R * kActRow + (size_t)2 * kBatch * planes * kActRow);
  return f * 4 + (tables ? warp_table_bytes(S, n_emits, srk) : 0);
}

template <int R, int METHOD>
static cudaError_t warp_launch_r(const FmaParams& p, WarpProg wp, int n_emits, int num_sms, int smem_optin, cudaStream_t stream) {
  constexpr bool srk = METHOD == 1;
  const int n_groups = (p.B + R - 1) / R;
  // small batches: one pair of warps per CTA so that every row group gets an SM of its own; otherwise up to 4 pairs
  const int pairs = std::max(1, std::min(4, (n_groups + num_sms - 1) / num_sms));
  const int grid = (n_groups + pairs - 1) / pairs;
  const bool ts = warp_table_bytes(p.S, n_emits, srk) <= 96 * 1024 &&
                  warp_smem_bytes(p.wimg_floats, pairs, R, p.S, n_emits, true, srk) <= (size_t)smem_optin;
  const size_t smem = warp_smem_bytes(p.wimg_floats, pairs, R, p.S, n_emits, ts, srk);
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
    return ts ? go(snsde_warp_kernel<N, R, true, METHOD>) : go(snsde_warp_kernel<N, R, false, METHOD>);
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
cudaError_t warp_launch(const FmaParams& p, const WarpProg& wp, int method, int n_emits, int num_sms, int smem_optin, cudaStream_t stream) {
  if (method == SNSDE_METHOD_SRK) return warp_launch_r<1, 1>(p, wp, n_emits, num_sms, smem_optin, stream);
  return warp_launch_r<1, 0>(p, wp, n_emits, num_sms, smem_optin, stream);
}

}  // namespace snsde
