// General tcgen05 kernel (see snsde_tcg.cuh).  Included by the instantiation units only.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "snsde_common.cuh"
#include "snsde_math.cuh"
#include "snsde_rng.cuh"
#include "snsde_tc_common.cuh"
#include "snsde_tc_ptx.cuh"
#include "snsde_tcg.cuh"

namespace snsde {

constexpr int kGEpiPerQuad = 4;
constexpr int kGEpiWarps = 4 * kGEpiPerQuad;          // warps 0..7
constexpr int kGMmaWarp = kGEpiWarps;                 // warp 8
constexpr int kGProdWarp0 = kGMmaWarp + 1;            // warps 9..10 : X(t) producer
constexpr int kGProdWarps = 2;
constexpr int kGStreamWarp = kGProdWarp0 + kGProdWarps; // warp 11 : weight streamer
constexpr int kGPrepWarp0 = kGStreamWarp + 1;         // warps 12..15 : step prefetch
constexpr int kGPrepWarps = 4;
constexpr int kTcgThreads = 32 * (kGPrepWarp0 + kGPrepWarps);
constexpr int kGProdThreads = 32 * kGProdWarps;
constexpr uint32_t kGALbo = 2048, kGASbo = 128;       // inside one 4 KB tile image: [k/8][row/8][row%8][k%8]
// named hardware barriers of the CTA-internal hand-offs (see snsde_tc.cu); 1 is the producers' own
constexpr int kGBarIn = 2, kGBarPFull = 3, kGBarPEmpty = 5;
constexpr int kGCntIn = 32 * (kGEpiWarps + 1), kGCntPrep = 32 * (kGEpiWarps + kGPrepWarps);

struct TcgSmem {
  int w, ring, b0, b1, x, stg, prep, bias, bars, total;
  int lbo_b, b_bytes, x_slot_bytes, stg_bytes, prep_bytes, n_bars;
};

// msplit: the launch is a cluster of CTA pairs, each CTA owning 128 of the 256 output features (HP = 128 local features)
// while the B operands hold all 256 input features (two buffers, ping-pong per phase) and two more mbarriers count the
// peer's half arriving.
__host__ __device__ inline TcgSmem tcg_smem_layout(int wres_bytes, int nslot, int HP, int nets, int C, int Cpad, int N, int NR,
                                                   int nx, int nstg, int NP, int uses_control, int msplit = 0) {
  TcgSmem s;
  s.lbo_b = (2 * N / 8) * 128 + 16;
  s.w = 0;
  s.ring = (wres_bytes + 127) & ~127;
  s.b0 = s.ring + nslot * kTcgSlotBytes;
  s.b_bytes = ((msplit ? 2 * HP : HP) / 8) * s.lbo_b;
  s.b1 = s.b0 + s.b_bytes;
  s.x = s.b1 + ((nets > 1 || msplit) ? s.b_bytes : 0);
  s.x_slot_bytes = uses_control ? (Cpad / 8) * s.lbo_b : 0;
  s.stg = s.x + nx * s.x_slot_bytes;
  s.stg_bytes = uses_control ? NR * 16 * C : 0;
  s.prep = (s.stg + nstg * s.stg_bytes + 15) & ~15;
  s.prep_bytes = (NR + 2) * HP * 4 + 32;
  s.bias = s.prep + 2 * s.prep_bytes;
  s.bars = s.bias + NP * nets * HP * 4;
  s.n_bars = 2 + 2 * nx + nstg + 4 + 2 * nslot + 2;     // the last two: peer-half arrival (msplit)
  s.total = s.bars + 8 * s.n_bars + 16;
  return s;
}

// MS ("M-split", hidden 129..256 with every weight tile resident): the launch is a cluster of CTA PAIRS.  Both CTAs of
// a pair integrate the same NR rows; CTA c owns output features [128c, 128c+128) of every layer - half the weights,
// which then fit in tensor memory + shared memory with nothing streamed - and after each layer pushes its half of the
// next B operand into the peer's shared memory with one DSMEM bulk copy (cp.async.bulk.shared::cluster.shared::cta)
// that completes on the peer's mbarrier.  The MMA warp issues the K chunks of its OWN half first and waits for the
// peer's half only then, so the exchange overlaps the first half of the layer's MMAs.  B operands ping-pong between two
// buffers per phase, which makes the exchange race-free without credits (a CTA can be at most one phase ahead).
template <int NR, int CH, int MT, int DIFF, bool MS>
__global__ void __launch_bounds__(kTcgThreads, 1) snsde_tcg_kernel(const TcgParams p) {
  static_assert(!MS || MT == 1, "M-split CTAs own one 128-feature tile");
  constexpr int N = NR < 16 ? 16 : NR;
  using Acc = AccRegion<N, CH>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, HP = p.HP, C = p.C, Cpad = p.Cpad, NP = p.NP, nets = p.nets;
  const TcgSmem L = tcg_smem_layout(p.wres_bytes, p.nslot, HP, nets, C, Cpad, N, NR, p.nx, p.nstg, NP, p.uses_control, MS ? 1 : 0);
  const uint32_t region_cols = (uint32_t)(nets * MT) * Acc::kCols;       // 2 regions: phase 0 | later phases
  // Thread-block cluster (1 = none): every CTA of a cluster streams the SAME weight tiles in the same order, so each
  // tile is fetched from L2 once and multicast into all their rings (c5: the chip-wide L2 stream, not the ring, bounded
  // the step).  CTA r issues the copies of the slots with index % CL == r; a slot is free once all CL CTAs have consumed it.
  const uint32_t CL = cluster_nctarank(), crank = cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
  const int fbase = MS ? 128 * (int)crank : 0;                           // global index of this CTA's local feature 0
  const int row0 = (MS ? (int)(blockIdx.x >> 1) : (int)blockIdx.x) * NR; // M-split: both CTAs of a pair share the rows

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  const uint32_t bar_acc = smem_u32(&bars[1]);
  const uint32_t bar_xfull = smem_u32(&bars[2]), bar_xempty = bar_xfull + 8 * p.nx;
  const uint32_t bar_cfull = bar_xempty + 8 * p.nx;
  const uint32_t bar_pfull = bar_cfull + 8 * p.nstg, bar_pempty = bar_pfull + 16;
  const uint32_t bar_rfull = bar_pempty + 16, bar_rempty = bar_rfull + 8 * p.nslot;
  const uint32_t bar_peer = bar_rempty + 8 * p.nslot;                    // [2]: the peer's half of operand buffer 0 / 1 landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[L.n_bars]);

  // ---- one-time setup ----
  for (int j = 0; j < p.n_jobs; ++j) {                 // resident weight segments -> smem
    const TcgJob& jb = p.jobs[j];
    if (jb.stream || jb.tmem_col >= 0) continue;
    const uint4* src = reinterpret_cast<const uint4*>(p.wblob + ((MS && crank) ? jb.g_off1 : jb.g_off));
    uint4* dst = reinterpret_cast<uint4*>(smem + L.w + jb.a_off);
    for (int i = tid; i < jb.nk * (kTcgSlotBytes / 16); i += kTcgThreads) dst[i] = src[i];
  }
  {
    uint4* z = reinterpret_cast<uint4*>(smem + L.b0);
    const int zn = (L.stg - L.b0) / 16;
    for (int i = tid; i < zn; i += kTcgThreads) z[i] = make_uint4(0, 0, 0, 0);
    float* sb = reinterpret_cast<float*>(smem + L.bias);
    for (int i = tid; i < NP * nets * HP; i += kTcgThreads) {
      const int f = i % HP, pn = i / HP, net = pn % nets, ph = pn / nets;
      const int off = p.bias[ph][net];
      sb[i] = (off >= 0 && fbase + f < H) ? p.vec[off + fbase + f] : 0.f;
    }
  }
  if (tid == 0) {
    mbar_init(bar_acc, 1);
    for (int i = 0; i < p.nx; ++i) { mbar_init(bar_xfull + 8 * i, kGProdWarps); mbar_init(bar_xempty + 8 * i, 1); }
    for (int i = 0; i < p.nstg; ++i) mbar_init(bar_cfull + 8 * i, 1);
    for (int i = 0; i < p.nslot; ++i) { mbar_init(bar_rfull + 8 * i, 1); mbar_init(bar_rempty + 8 * i, CL); }
    mbar_init(bar_peer, 1); mbar_init(bar_peer + 8, 1);
    mbar_fence_init();
  }
  if (warp == kGMmaWarp) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();            // every CTA's barriers exist before a peer arrives on them / copies into the rings
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM-resident jobs: lane m = weight row m of the tile, each 32-bit column packs two consecutive K elements
  // (one 16-byte core-matrix row of the tile image = 4 columns).  Warp w may only touch lane quadrant w % 4.
  if (warp < kGEpiWarps) {
    const int m = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int j = 0; j < p.n_jobs; ++j) {
      const int col = p.jobs[j].tmem_col, nk = p.jobs[j].nk;
      if (col < 0) continue;
      const uint8_t* src = p.wblob + ((MS && crank) ? p.jobs[j].g_off1 : p.jobs[j].g_off) + (m >> 3) * kGASbo + (m & 7) * 16;
      for (int i = (warp >> 2); i < 2 * nk; i += kGEpiPerQuad) {          // i < nk: hi image of chunk i; else lo image
        const int kb = i < nk ? i : i - nk;
        const uint8_t* q = src + (size_t)kb * kTcgSlotBytes + (i < nk ? 0 : 4096);
        const uint4 lo = *reinterpret_cast<const uint4*>(q);
        const uint4 hi = *reinterpret_cast<const uint4*>(q + kGALbo);
        const uint32_t r[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        tmem_st8(tmem + lane_base + (uint32_t)(col + 8 * i), r);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < kGEpiWarps) {
    // =========================== EPILOGUE / SDE STATE ===========================
    constexpr int RT = NR / kGEpiPerQuad;
    constexpr int LW = RT < 8 ? RT : 8;
    const int h = (warp & 3) * 32 + lane;
    const int rbase = (warp >> 2) * RT;
    const TailOp t = p.tail;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const bool net2 = nets > 1;
    float vmax = 0.f;
    auto write_operand = [&](int buf_off, int f, int r, float v) {
      __half hi, lo;
      split_f16(v, hi, lo);
      vmax = fmaxf(vmax, fabsf(v));                 // range check of the split-fp16 operands: one flag write at the end
      uint8_t* q = smem + buf_off + (f >> 3) * L.lbo_b + (f & 7) * 2 + (r >> 3) * 128 + (r & 7) * 16;
      *reinterpret_cast<__half*>(q) = hi;
      *reinterpret_cast<__half*>(q + (N / 8) * 128) = lo;
    };
    auto hand_over = [&]() {
      tc_fence_before();
      fence_proxy_async_smem();
      named_arrive(kGBarIn, kGCntIn);
    };
    const float* sbias = reinterpret_cast<const float*>(smem + L.bias);
    auto bias_of = [&](int ph, int net, int f) { return sbias[(ph * nets + net) * HP + f]; };
    // accumulators of set `acc` in the region of phase `ph` -> vm (main) / vc (scaled correction), this thread's rows
    auto load_acc = [&](int ph, int acc, float (&vm)[RT], float (&vc)[RT]) {
      const uint32_t dreg = tmem + lane_base + (ph == 0 ? 0u : region_cols) + (uint32_t)acc * Acc::kCols + rbase;
#pragma unroll
      for (int c = 0; c < RT; c += LW) {
        float m8[CH][LW], a8[CH][LW];
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          tmem_ldw<LW>(dreg + Acc::a(ch) + c, m8[ch]);
          tmem_ldw<LW>(dreg + Acc::a(ch) + N + c, a8[ch]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < LW; ++i) {
          float m = m8[0][i], cc = a8[0][i];
#pragma unroll
          for (int ch = 1; ch < CH; ++ch) { m += m8[ch][i]; cc += a8[ch][i]; }
          vm[c + i] = m; vc[c + i] = cc;
        }
      }
    };

    constexpr bool PRE_OK = RT <= 8;
    const bool pre = PRE_OK && !net2;                 // diffusion precomputed in the MMA shadow (needs no net output)
    constexpr int NPRE = PRE_OK ? RT : 1;
    float y[MT][RT], yprev[MT][RT], qn[MT][RT], gv[MT][NPRE], dg[MT][NPRE], thy[MT][NPRE];
    int myslot[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      const int b = min(row0 + rbase + i, p.B - 1);
      myslot[i] = p.row_slot ? p.row_slot[b] : -1;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int f = fbase + h + 128 * mt;                 // global feature
        y[mt][i] = f < H ? p.y0[(size_t)b * H + f] : 0.f;
        yprev[mt][i] = y[mt][i];
        qn[mt][i] = 0.f;
        if (f < H) write_operand(L.b0, f, rbase + i, y[mt][i]);
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < NPRE; ++i) gv[mt][i] = dg[mt][i] = thy[mt][i] = 0.f;

    auto emit = [&](snsde_emit em) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int f = fbase + h + 128 * mt;
        if (f >= H) continue;
#pragma unroll
        for (int i = 0; i < RT; ++i) {
          const int gr = row0 + rbase + i;
          if (gr >= p.B) continue;
          const float v = em.w_prev * yprev[mt][i] + em.w_curr * y[mt][i];
          if (p.row_slot) {
            if (myslot[i] == em.slot) p.out[(size_t)gr * H + f] = v;
          } else {
            p.out[((size_t)em.slot * p.B + gr) * H + f] = v;
          }
        }
      }
    };
    hand_over();
    // Outputs are written at ONE place in the code (top of the loop: the emits the previous step left pending,
    // initially the outputs at ts[0] with weights (0, 1)): the hot code has to fit the 32 KB instruction cache.
    int pend_n = p.n_init_emits, pend_begin = 0;
    snsde_emit pend_first;
    pend_first.slot = 0; pend_first.w_prev = 0.f; pend_first.w_curr = 1.f;
    if (pend_n > 0) { pend_first = p.emits[0]; pend_first.w_prev = 0.f; pend_first.w_curr = 1.f; }

    auto state_terms = [&](float yr, float cf, float t0, float& g, float& dgy, float& th) {
      if (DIFF == 1) {
        const float raw = cf * yr;
        const bool fin = (raw == raw) && (fabsf(raw) != INFINITY);
        g = tanh_fast(t.s_theta * nan_to_num_f(raw));
        dgy = t.milstein ? ((1.f - g * g) * t.s_theta) * (fin ? 1.f : 0.f) * cf : 0.f;
      } else if (DIFF == 2) {
        // state-network noise (options 14,15,18,19; BASELINE c4 is (3,18)): cf = q, the network's output for this row.
        // Euler only (Milstein needs the vjp through the network: fp32 kernel), so no derivative.  Kept apart from the
        // generic form below, whose inlined switch (sqrt / sigmoid / division slow paths) made this kernel 128 KB of
        // SASS against a 32 KB instruction cache.
        const float raw = (t.mult == MU_Y) ? cf * yr : cf;
        g = tanh_fast(t.s_theta * nan_to_num_f(raw));
        dgy = 0.f;
      } else {
        diffusion_eval<true>(t, cf, yr, t0, g, dgy);
      }
      th = t.geometric ? tanh_fast(yr) : 1.f;
    };

    uint32_t pacc = 0;
    for (int s = 0; s <= p.S; ++s) {
#pragma unroll 1
      for (int e = 0; e < pend_n; ++e) {
        snsde_emit em = pend_first;
        if (e > 0) {
          em = p.emits[pend_begin + e];
          if (s == 0) { em.w_prev = 0.f; em.w_curr = 1.f; }
        }
        emit(em);
      }
      if (s == p.S) break;
      const uint8_t* slot = smem + L.prep + (s & 1) * L.prep_bytes;
      const float* sdw = reinterpret_cast<const float*>(slot);
      named_sync(kGBarPFull + (s & 1), kGCntPrep);
      const StepInfo si = *reinterpret_cast<const StepInfo*>(slot + (NR + 2) * HP * 4);
      float add0[MT], vec1[MT];                       // folded layer-0 bias; diffusion coefficient or noise-net layer-0 bias
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        add0[mt] = sdw[NR * HP + h + 128 * mt];
        vec1[mt] = sdw[(NR + 1) * HP + h + 128 * mt];
      }
      if (pre) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int i = 0; i < NPRE; ++i)
            if (fbase + h + 128 * mt < H) state_terms(y[mt][i], vec1[mt], si.t0, gv[mt][i], dg[mt][i], thy[mt][i]);
      }
      for (int ph = 0; ph < NP; ++ph) {
        mbar_wait(bar_acc, pacc);
        pacc ^= 1;
        tc_fence_after();
        TC_TRACE(tid == 0 && ph < 2, s, ph == 0 ? EV_EPI_ACC0 : EV_EPI_ACC1);
        // M-split: the operand of global phase q lives in buffer q & 1; this phase (q = s*NP + ph) writes the next one's
        const int ob = MS ? ((((s * NP + ph + 1) & 1) != 0) ? L.b1 : L.b0) : L.b0;
        // ---- noise network (state-dependent noise options), layers 0..NN-1 ride on phases 0..NN-1 ----
        if (net2 && ph < p.NN) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int f = h + 128 * mt;
            float vm[RT], vc[RT];
            load_acc(ph, MT + mt, vm, vc);
            if (f < H) {
              const float add = (ph == 0) ? vec1[mt] : bias_of(ph, 1, f);
#pragma unroll
              for (int i = 0; i < RT; ++i) {
                const float v = act_apply(fmaf(vc[i], kLoInv, vm[i]) + add, p.noise_act[ph]);
                if (ph < p.NN - 1) write_operand(L.b1, f, rbase + i, v);
                else qn[mt][i] = v;
              }
            }
          }
        }
        // ---- drift network ----
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int fl = h + 128 * mt, f = fbase + fl;        // local (tables in shared memory) / global feature
          float vm[RT], vc[RT];
          load_acc(ph, mt, vm, vc);
          TC_TRACE(tid == 0 && ph < 2 && mt == 0, s, ph == 0 ? EV_EPI_LD0 : EV_EPI_LD1);
          if (f >= H) continue;
          if (ph < NP - 1) {
            const float add = (ph == 0) ? add0[mt] : bias_of(ph, 0, fl);
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              float v = fmaf(vc[i], kLoInv, vm[i]) + add;
              v = v < 0.f ? 0.f : v;
              write_operand(ob, f, rbase + i, v);
            }
          } else {
            const float bl = bias_of(ph, 0, fl);
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              float d = fmaf(vc[i], kLoInv, vm[i]) + bl;
              float g, dgy, th;
              if (PRE_OK && pre) { g = gv[mt][i < NPRE ? i : 0]; dgy = dg[mt][i < NPRE ? i : 0]; th = thy[mt][i < NPRE ? i : 0]; }
              else state_terms(y[mt][i], net2 ? qn[mt][i] : vec1[mt], si.t0, g, dgy, th);
              if (t.geometric) d *= th;
              if (t.clip_drift) d = tanh_fast(d);
              const float dw = sdw[(rbase + i) * HP + fl];
              float yn = __fadd_rn(__fadd_rn(y[mt][i], __fmul_rn(d, si.h)), __fmul_rn(g, dw));
              if (t.milstein) {
                const float v2 = __fmul_rn(dw, dw) - si.h;
                yn = __fadd_rn(yn, 0.5f * ((g * v2) * dgy));
              }
              yprev[mt][i] = y[mt][i];
              y[mt][i] = yn;
              write_operand(ob, f, rbase + i, yn);
            }
          }
        }
        hand_over();
        TC_TRACE(tid == 0 && ph < 2, s, ph == 0 ? EV_EPI_DONE0 : EV_EPI_DONE1);
      }
      named_arrive(kGBarPEmpty + (s & 1), kGCntPrep);
      pend_n = si.n_emits; pend_begin = si.emit_begin; pend_first = si.first;
      TC_TRACE(tid == 0, s, EV_EPI_SHADOW_END);
    }
    if (vmax > 65504.f) *p.status = 1;              // an operand beyond the fp16 range was saturated (sticky flag)
  } else if (warp == kGMmaWarp) {
    // =========================== MMA ISSUER (warp-uniform, one elected lane) ===========================
    const bool leader = elect_one();
    const uint32_t w_base = smem_u32(smem + L.w), ring_base = smem_u32(smem + L.ring);
    const uint32_t b_base0 = smem_u32(smem + L.b0), b_base1 = smem_u32(smem + L.b1);
    const uint32_t x_base = smem_u32(smem + L.x);
    constexpr uint32_t idesc2 = umma_idesc_f16(128, 2 * N), idesc1 = umma_idesc_f16(128, N);
    const uint64_t b_step = (uint64_t)((2 * L.lbo_b) >> 4);
    uint32_t rslot = 0, rphase = 0, xphase = 0;
    int xslot = 0;
    // Descriptor arithmetic is incremental (one 64-bit add per operand and chunk) and the resident / streamed
    // cases are separate loops: the issue rate of this warp bounds the step time.
    const uint64_t ring_desc0 = umma_smem_desc(ring_base, kGALbo, kGASbo);
    static_assert(CH == 1, "the resident issue loops assume one accumulator chain");
    // chunks [kb0, kb1) of job j; `acc0`: accumulate flag of the first MMA issued (later ones always accumulate)
    auto issue_job = [&](int j, uint32_t bbase, int kb0, int kb1, uint32_t acc0) {
      const int nk = p.jobs[j].nk;
      const uint32_t d = tmem + (p.jobs[j].phase == 0 ? 0u : region_cols) + (uint32_t)p.jobs[j].acc * Acc::kCols;
      uint64_t db = umma_smem_desc(bbase + (p.jobs[j].b_chunk0 + kb0) * 2 * L.lbo_b, L.lbo_b, 128);
      uint32_t acc = acc0;
      // Resident tiles: the elected lane alone runs the issue loop (the others wait at the warp sync), unrolled by
      // 4 chunks: with the election test and the loop bookkeeping inside every iteration an MMA took ~50 cycles to issue
      // (clock trace, c4: 1584 cycles for the 32 MMAs of a phase) against ~20 in the resident kernel.
      if (p.jobs[j].tmem_col >= 0) {                     // tiles resident in TMEM: TS-form MMAs (11-17 vs >= 39 cycles)
        if (leader) {
          uint32_t ah = tmem + (uint32_t)p.jobs[j].tmem_col + 8u * (uint32_t)kb0, al = ah + 8u * (uint32_t)nk;
#pragma unroll 4
          for (int kb = kb0; kb < kb1; ++kb) {
            umma_f16_ts(d + Acc::a(0), ah, db, idesc2, acc);
            umma_f16_ts(d + Acc::b(0), al, db, idesc1, 1u);
            ah += 8; al += 8;
            db += b_step;
            acc = 1u;
          }
        }
        __syncwarp();
      } else if (!p.jobs[j].stream) {
        if (leader) {
          uint64_t da = umma_smem_desc(w_base + p.jobs[j].a_off + kb0 * kTcgSlotBytes, kGALbo, kGASbo);
#pragma unroll 4
          for (int kb = kb0; kb < kb1; ++kb) {
            umma_f16(d + Acc::a(0), da, db, idesc2, acc);
            umma_f16(d + Acc::b(0), da + (4096 >> 4), db, idesc1, 1u);     // corr columns: the hi product just initialised them
            da += (uint64_t)(kTcgSlotBytes >> 4);
            db += b_step;
            acc = 1u;
          }
        }
        __syncwarp();
      } else {                                           // streamed tiles (never under M-split): whole job, ring order
        for (int kb = kb0; kb < kb1; kb += CH) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            mbar_wait(bar_rfull + 8 * rslot, rphase);
            tc_fence_after();
            const uint64_t da = ring_desc0 + (uint64_t)(rslot * (kTcgSlotBytes >> 4));
            if (leader) {
              umma_f16(d + Acc::a(c), da, db, idesc2, acc);
              umma_f16(d + Acc::b(c), da + (4096 >> 4), db, idesc1, 1u);     // corr columns: the hi product just initialised them
              if (CL > 1) umma_commit_multicast(bar_rempty + 8 * rslot, cmask);   // "consumed" to every CTA of the cluster
              else umma_commit(bar_rempty + 8 * rslot);
            }
            __syncwarp();
            if (++rslot == (uint32_t)p.nslot) { rslot = 0; rphase ^= 1; }
            db += b_step;
          }
          acc = 1u;
        }
      }
    };
    // One step = NP phases of drift/noise jobs + the X(t) jobs of the NEXT step's phase 0 (issued a phase early);
    // every job goes through ONE inlined copy of issue_job (instruction cache).
    const int n_main = p.n_jobs - p.n_xjobs;
    for (int s = -1; s < p.S; ++s) {
      int j = 0;
      for (int ph = (s < 0 ? NP : 0); ph <= NP; ++ph) {
        const bool isx = ph == NP;
        if (isx && !(p.n_xjobs > 0 && s + 1 < p.S)) continue;
        int j_end;
        uint32_t commit_bar, xb = 0;
        if (isx) {
          mbar_wait(bar_xfull + 8 * xslot, xphase);
          j = n_main; j_end = p.n_jobs;
          xb = x_base + xslot * L.x_slot_bytes;
          commit_bar = bar_xempty + 8 * xslot;
          if (++xslot == p.nx) { xslot = 0; xphase ^= 1; }
        } else {
          named_sync(kGBarIn, kGCntIn);
          j_end = j;
          while (j_end < n_main && p.jobs[j_end].phase == ph) ++j_end;
          commit_bar = bar_acc;
        }
        tc_fence_after();
        TC_TRACE(lane == 0 && ph < 2, s, ph == 0 ? EV_MMA_WAKE0 : EV_MMA_WAKE1);
        if (MS && !isx) {
          // operand of global phase q = s*NP + ph: buffer q & 1.  Send this CTA's half to the peer, run the MMAs over the
          // own half, wait for the peer's half, run the rest.
          const uint32_t q = (uint32_t)(s * NP + ph), buf = q & 1u;
          const uint32_t bbase = buf ? b_base1 : b_base0;
          const uint32_t half_bytes = 16u * (uint32_t)L.lbo_b, own_off = crank * half_bytes;
          if (lane == 0) {
            mbar_arrive_expect_tx(bar_peer + 8 * buf, half_bytes);                 // the bytes the PEER will deliver here
            dsmem_bulk_copy(mapa_u32(bbase + own_off, crank ^ 1u), bbase + own_off, half_bytes, mapa_u32(bar_peer + 8 * buf, crank ^ 1u));
          }
          __syncwarp();
          const int k_own = 8 * (int)crank, k_peer = 8 * (int)(crank ^ 1u);
          const int j0 = j;
#pragma unroll 1
          for (; j < j_end; ++j) issue_job(j, bbase, k_own, k_own + 8, p.jobs[j].fresh ? 0u : 1u);
          mbar_wait(bar_peer + 8 * buf, (q >> 1) & 1u);
          tc_fence_after();
#pragma unroll 1
          for (j = j0; j < j_end; ++j) issue_job(j, bbase, k_peer, k_peer + 8, 1u);
        } else {
#pragma unroll 1
          for (; j < j_end; ++j)
            issue_job(j, isx ? xb : (p.jobs[j].b_src ? b_base1 : b_base0), 0, p.jobs[j].nk, p.jobs[j].fresh ? 0u : 1u);
        }
        if (leader) umma_commit(commit_bar);
        __syncwarp();
        TC_TRACE(lane == 0 && (ph < 2 || isx), s, isx ? EV_MMA_X_DONE : (ph == 0 ? EV_MMA_COMMIT0 : EV_MMA_COMMIT1));
      }
    }
  } else if (warp < kGStreamWarp) {
    // =========================== CONTROL PRODUCER ===========================
    // Same scheme as the resident kernel (snsde_tc.cu): copies spread over the producer warps, the items of a thread
    // tabulated once and evaluated interleaved, x/3 by the FMA-corrected reciprocal product.
    if (p.uses_control) {
      const int ptid = tid - 32 * kGProdWarp0;
      const int pwarp = warp - kGProdWarp0;
      const uint32_t row_bytes = 16u * C;
      const int rows_per_warp = (NR + kGProdWarps - 1) / kGProdWarps;
      auto fetch = [&](int s, int interval) {           // spline rows of step s -> staging slot
        if (s >= p.S) return;
        const int stg = s % p.nstg;
        const uint32_t bar = bar_cfull + 8 * stg;
        const int r = pwarp * rows_per_warp + lane;
        if (lane < rows_per_warp && r < NR) {
          const int b = min(row0 + r, p.B - 1);
          const float* src = p.coeffs + (size_t)b * p.coeff_row_stride + (size_t)interval * 4 * C;
          bulk_g2s(smem_u32(smem + L.stg + stg * L.stg_bytes + r * row_bytes), src, row_bytes, bar);
        }
      };
      auto expect = [&](int s) {                        // one thread announces the bytes of step s
        if (ptid == 0 && s < p.S) mbar_arrive_expect_tx(bar_cfull + 8 * (s % p.nstg), row_bytes * NR);
      };
      constexpr int kItems = 4;
      int item_src[kItems], item_dst[kItems];
#pragma unroll
      for (int k = 0; k < kItems; ++k) {
        const int i = ptid + k * kGProdThreads;
        const int r = i / C, c = i - r * C;
        item_src[k] = (i < NR * C) ? r * 4 * C + c : -1;
        item_dst[k] = (c >> 3) * L.lbo_b + (c & 7) * 2 + (r >> 3) * 128 + (r & 7) * 16;
      }
      float xmax = 0.f;
      auto eval_item = [&](const float* rows, uint8_t* xs, int src, int dst, float frac) {
        const float* q0 = rows + src;
        const float v = q0[3 * C] * frac;
        float q = v * 0.333333343f;                     // v / 3, correctly rounded: Newton step on the residual
        q = fmaf(fmaf(-3.0f, q, v), 0.333333343f, q);
        float inner = 0.5f * q0[2 * C] + q;
        inner = q0[C] + inner * frac;
        const float x = q0[0] + inner * frac;
        __half hi, lo;
        split_f16(x, hi, lo);
        xmax = fmaxf(xmax, fabsf(x));
        *reinterpret_cast<__half*>(xs + dst) = hi;
        *reinterpret_cast<__half*>(xs + dst + (N / 8) * 128) = lo;
      };
#pragma unroll 1
      for (int s = 0; s < p.nstg - 1; ++s) {
        expect(s);
        asm volatile("bar.sync 1, %0;" ::"n"(kGProdThreads));
        fetch(s, s < p.S ? p.steps[s].interval : 0);
      }
      int interval_ahead = (p.nstg - 1 < p.S) ? p.steps[p.nstg - 1].interval : 0;
      float frac_cur = p.S > 0 ? p.steps[0].frac : 0.f;
      for (int s = 0; s < p.S; ++s) {
        expect(s + p.nstg - 1);
        asm volatile("bar.sync 1, %0;" ::"n"(kGProdThreads));      // all producer warps are done with step s-1
        fetch(s + p.nstg - 1, interval_ahead);
        const int sa = s + p.nstg;
        const int interval_next = sa < p.S ? p.steps[sa].interval : 0;       // consumed next iteration
        const float frac_next = s + 1 < p.S ? p.steps[s + 1].frac : 0.f;
        const int stg = s % p.nstg, slot = s % p.nx;
        const float frac = frac_cur;
        mbar_wait(bar_cfull + 8 * stg, (uint32_t)((s / p.nstg) & 1));
        if (s >= p.nx) mbar_wait(bar_xempty + 8 * slot, (uint32_t)(((s / p.nx) - 1) & 1));
        const float* rows = reinterpret_cast<const float*>(smem + L.stg + stg * L.stg_bytes);
        uint8_t* xs = smem + L.x + slot * L.x_slot_bytes;
#pragma unroll
        for (int k = 0; k < kItems; ++k)
          if (item_src[k] >= 0) eval_item(rows, xs, item_src[k], item_dst[k], frac);
#pragma unroll 1
        for (int i = ptid + kItems * kGProdThreads; i < NR * C; i += kGProdThreads) {
          const int r = i / C, c = i - r * C;
          eval_item(rows, xs, r * 4 * C + c, (c >> 3) * L.lbo_b + (c & 7) * 2 + (r >> 3) * 128 + (r & 7) * 16, frac);
        }
        interval_ahead = interval_next;
        frac_cur = frac_next;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_xfull + 8 * slot);
      }
      if (xmax > 65504.f) *p.status = 1;
    }
  } else if (warp >= kGPrepWarp0) {
    // =========================== STEP PREFETCH (time-only work) ===========================
    const int h = tid - 32 * kGPrepWarp0;
    const TailOp t = p.tail;
    float c0[MT], csin[MT], ccos[MT], n0[MT], nsin[MT], ncos[MT], coef[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int f = fbase + h + 128 * mt;
      const bool a = f < H;
      c0[mt] = (a && p.bias[0][0] >= 0) ? p.vec[p.bias[0][0] + f] : 0.f;
      csin[mt] = (a && p.c_sin[0] >= 0) ? p.vec[p.c_sin[0] + f] : 0.f;
      ccos[mt] = (a && p.c_cos[0] >= 0) ? p.vec[p.c_cos[0] + f] : 0.f;
      n0[mt] = (a && nets > 1 && p.bias[0][1] >= 0) ? p.vec[p.bias[0][1] + f] : 0.f;
      nsin[mt] = (a && p.c_sin[1] >= 0) ? p.vec[p.c_sin[1] + f] : 0.f;
      ncos[mt] = (a && p.c_cos[1] >= 0) ? p.vec[p.c_cos[1] + f] : 0.f;
      coef[mt] = t.coef_scalar;
      if (t.coef_src == CO_IMG && a) coef[mt] = p.vec[p.coef_vec + f];
    }
    // Global rows of this CTA start at gb0; Philox yields 4 normals for the aligned row quad gb >> 2, so the NR
    // rows span nq quads.  One quad per iteration of a NOT unrolled loop (I-cache).
    const unsigned long long gb0 = p.row_offset + (unsigned long long)row0;
    const int lane0 = (int)(gb0 & 3ull);
    const int nq = (lane0 + NR + 3) >> 2;
    for (int s = 0; s < p.S; ++s) {
      const snsde_step st = p.steps[s];
      uint8_t* slot = smem + L.prep + (s & 1) * L.prep_bytes;
      float* sdw = reinterpret_cast<float*>(slot);
      float v1[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int f = fbase + h + 128 * mt;
        v1[mt] = coef[mt];
        if (f >= H) continue;
        if (nets > 1) v1[mt] = fmaf(st.cos_t0, ncos[mt], fmaf(st.sin_t0, nsin[mt], n0[mt]));
        else if (t.coef_src == CO_VBUF) v1[mt] = p.a_tab[(size_t)s * H + f];
      }
      StepInfo si;
      si.h = st.h; si.t0 = st.t0; si.n_emits = st.emit_end - st.emit_begin; si.emit_begin = st.emit_begin;
      si.first.slot = 0; si.first.w_prev = 0.f; si.first.w_curr = 0.f;
      if (h == 0 && si.n_emits > 0) si.first = p.emits[st.emit_begin];
      if (s >= 2) named_sync(kGBarPEmpty + (s & 1), kGCntPrep);
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int fl = h + 128 * mt, f = fbase + fl;
        if (f >= H) continue;
        if (p.dW != nullptr) {
#pragma unroll 4
          for (int r = 0; r < NR; ++r) sdw[r * HP + fl] = p.dW[((size_t)s * p.B + min(row0 + r, p.B - 1)) * H + f];
        } else {
#pragma unroll 1
          for (int q = 0; q < nq; ++q) {
            float nrm[4];
            philox_normals4(p.seed, (uint32_t)f, (uint32_t)((gb0 >> 2) + (unsigned long long)q), (uint32_t)s, nrm);
            const int rq = q * 4 - lane0;               // CTA-local row of the quad's first normal
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if ((unsigned)(rq + k) < (unsigned)NR) sdw[(rq + k) * HP + fl] = __fmul_rn(nrm[k], st.sqrt_h);
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int fl = h + 128 * mt;
        if (fbase + fl >= H) continue;
        sdw[NR * HP + fl] = fmaf(st.cos_t0, ccos[mt], fmaf(st.sin_t0, csin[mt], c0[mt]));
        sdw[(NR + 1) * HP + fl] = v1[mt];
      }
      if (h == 0) *reinterpret_cast<StepInfo*>(slot + (NR + 2) * HP * 4) = si;
      named_arrive(kGBarPFull + (s & 1), kGCntPrep);
    }
  } else {
    // =========================== WEIGHT STREAMER ===========================
    // The streamed segments are consumed in the same order every step, so one lane keeps the ring full with
    // 8 KB bulk copies, running ahead of the MMA warp across layers and steps.
    if (lane == 0 && p.n_stream_chunks > 0) {
      uint32_t slot = 0, phase = 0, turn = 0;               // turn: whose copy this is (round robin over the cluster)
      for (int s = 0; s < p.S; ++s) {
        for (int j = 0; j < p.n_jobs; ++j) {
          const TcgJob& jb = p.jobs[j];
          if (!jb.stream) continue;
          for (int kb = 0; kb < jb.nk; ++kb) {
            mbar_wait(bar_rempty + 8 * slot, phase ^ 1);      // tight poll: a sleeping streamer caps the ring at ~30 B/cycle
            mbar_arrive_expect_tx(bar_rfull + 8 * slot, kTcgSlotBytes);
            const uint8_t* src = p.wblob + jb.g_off + (size_t)kb * kTcgSlotBytes;
            const uint32_t dst = smem_u32(smem + L.ring + slot * kTcgSlotBytes);
            if (CL == 1) bulk_g2s(dst, src, kTcgSlotBytes, bar_rfull + 8 * slot);
            else if (turn == crank) bulk_g2s_multicast(dst, src, kTcgSlotBytes, bar_rfull + 8 * slot, cmask);
            if (++turn == CL) turn = 0;
            if (++slot == (uint32_t)p.nslot) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kGMmaWarp) tmem_dealloc(tmem, 512);
  if (CL > 1) cluster_sync_all();            // no CTA leaves while a peer may still multicast into its ring / barriers
}

template <int NR, int CH, int MT, int DIFF, bool MS>
cudaError_t tcg_launch(const TcgParams& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = snsde_tcg_kernel<NR, CH, MT, DIFF, MS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // Streamed weights: launch thread-block clusters so that one L2 read of a tile feeds several CTAs (multicast).
  // Largest cluster (4, 2) the device can keep resident for the whole grid in one wave; SNSDE_TCG_CLUSTER overrides.
  int cl = 1;
  if (MS) cl = 2;                              // `grid` already counts both CTAs of every pair
  else if (p.n_stream_chunks > 0) {
    const char* env = getenv("SNSDE_TCG_CLUSTER");
    for (int want : {4, 2}) {
      if (env != nullptr && atoi(env) != want) continue;
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3((unsigned)((grid + want - 1) / want * want)); q.blockDim = dim3(kTcgThreads); q.dynamicSmemBytes = smem;
      cudaLaunchAttribute a[1];
      a[0].id = cudaLaunchAttributeClusterDimension;
      a[0].val.clusterDim.x = want; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
      q.attrs = a; q.numAttrs = 1;
      int n_clusters = 0;
      const cudaError_t eo = cudaOccupancyMaxActiveClusters(&n_clusters, kern, &q);
      if (getenv("SNSDE_TCG_DEBUG") != nullptr)
        fprintf(stderr, "[snsde] cluster %d: occupancy query %s, %d active clusters, grid %u\n", want, cudaGetErrorString(eo), n_clusters, q.gridDim.x);
      if (eo == cudaSuccess && n_clusters * want >= (int)q.gridDim.x) { cl = want; break; }
      (void)cudaGetLastError();
    }
    if (env != nullptr && atoi(env) == 1) cl = 1;
  }
  if (cl == 1) {
    kern<<<grid, kTcgThreads, smem, stream>>>(p);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)((grid + cl - 1) / cl * cl)); cfg.blockDim = dim3(kTcgThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

}  // namespace snsde
