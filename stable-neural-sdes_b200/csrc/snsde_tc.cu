// tcgen05 tensor-core kernel: the whole fixed-step SDE solve for NR batch rows per CTA in ONE
// persistent launch, with the drift MLP contractions on the 5th-gen tensor cores.
//
// Replaces, for the named models, the same reference code as snsde_fma.cu:
// torchsde.sdeint's step loop + Diffusion_model.f/g
// (/root/reference/benchmark_classification/models_sde/neuralsde.py:295-307) + CubicSpline.evaluate.
//
// Orientation.  The batch is tiny per SM (1024 rows / 148 SMs ~ 8 rows), the hidden width is
// 64..128.  So the WEIGHTS are the M=128 operand A - resident in TENSOR MEMORY for the whole kernel
// (TS-form tcgen05.mma; what does not fit stays in shared memory as SS-form images) - and the batch
// rows are the small N: D^T[feature, row] = W[feature, :] . act[row, :].
// TMEM lane i then holds feature i for every row, i.e. epilogue thread i owns feature i - the
// same thread/feature mapping as the FMA kernel, so the SDE state never leaves registers.
//
// Precision.  fp32 parity (1e-4 over hundreds of recurrent steps) on fp16 tensor cores:
// every operand v is split v ~ hi + 2^-11 * lo', hi = fp16(v), lo' = fp16((v - hi) * 2^11)
// (22 significant bits), and   W.a ~ Whi.ahi + 2^-11 (Wlo'.ahi + Whi.alo')   with the main and the
// correction sums in separate fp32 TMEM accumulator columns, combined in the epilogue.  Two MMAs per
// 16-wide K chunk:  Whi x [ahi ; alo'] (N' = 2N columns: [main | corr]) and Wlo' x ahi -> the same corr columns.
//
// Algebra.  Input options 2,4,6 have no nonlinearity between linear_in and emb
// (neuralsde.py:202,210), so layer 0 is collapsed on the host in double precision:
//   z0 = (We1 Win_y) y + (We2 Wi) X(t) + [be + We1 bin + We2 bi] + (We1 Win_tau) [sin t, cos t].
//
// Warp roles (768 threads, 1 CTA/SM):
//   warps 0-15  epilogue (four warps per TMEM lane quadrant, each thread owns one feature of a quarter of
//               the rows): tcgen05.ld accumulators -> bias/activation or SDE update -> split -> write the
//               next B operand; outputs and the diffusion of the new state run in the shadow of the MMAs
//   warp  16    issues every tcgen05.mma (warp-uniform code, one elected lane) and commits to mbarriers
//   warps 17-19 control producer: 1-D bulk async copies (TMA) of the spline rows several steps
//               ahead, cubic evaluation, split, write of the X(t) operand ring
//   warps 20-23 step prefetch: everything that depends on time only - Philox/Box-Muller (or table)
//               increments, folded layer-0 bias, diffusion coefficient, step/emit descriptors -
//               one step ahead, through a 2-deep shared-memory ring
// Hand-offs between warps are hardware named barriers; mbarriers only where the async proxy (tcgen05.commit,
// TMA) is the producer.  The hot code is kept inside the 32 KB instruction cache (compact loops, one emit site).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "snsde_common.cuh"
#include "snsde_math.cuh"
#include "snsde_rng.cuh"
#include "snsde_tc.cuh"
#include "snsde_tc_ptx.cuh"
#include "snsde_tc_common.cuh"

namespace snsde {

using namespace ptx;

constexpr int kEpiPerQuad = 4;                     // epilogue warps per TMEM lane quadrant, each owning NR/kEpiPerQuad rows
constexpr int kEpiWarps = 4 * kEpiPerQuad;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kProdWarp0 = kMmaWarp + 1, kPrepWarp0 = kProdWarp0 + 3;
constexpr int kTcThreads = 32 * (kPrepWarp0 + 4);
constexpr int kProdWarps = 3;
constexpr int kProdThreads = 32 * kProdWarps;
constexpr int kPrepWarps = 4;
constexpr int kCntIn = 32 * (kEpiWarps + 1), kCntPrep = 32 * (kEpiWarps + kPrepWarps);
constexpr uint32_t kASbo = 128, kALbo = 2048;      // A images: 16 row groups contiguous, then K chunks
// Warp-to-warp hand-offs inside the CTA use HARDWARE named barriers (bar.arrive / bar.sync), not mbarriers: with 16
// epilogue warps every mbarrier operation (a shared-memory atomic through the SYNCS unit) was measured at 230-460
// cycles on the critical path (trace: arrive 467, try_wait 229 / 231); only the hand-offs whose producer is the
// async proxy (tcgen05.commit, TMA) need an mbarrier.  Barrier 0 is __syncthreads, 1 the producers' own.
constexpr int kBarIn = 2;                           // epilogue warps arrive, MMA warp syncs
constexpr int kBarPFull = 3;                        // +slot: step-prefetch warps arrive, epilogue warps sync
constexpr int kBarPEmpty = 5;                       // +slot: epilogue warps arrive, step-prefetch warps sync


struct TcSmem {
  int w, b, x, stg, prep, bias, outst, bars, total;
  int lbo_b;            // bytes between K chunks (8 columns) of a B operand
  int x_slot_bytes, stg_bytes, prep_bytes;
};

// prep slot: [NR][128] dW floats | [128] add0 | [128] coef | StepInfo
__host__ __device__ inline TcSmem tc_smem_layout(int wimg_bytes, int H, int C, int Cpad, int N, int NR, int nx, int nstg,
                                                 int uses_control) {
  TcSmem s;
  s.lbo_b = (2 * N / 8) * 128 + 16;                 // +16: de-conflicts the epilogue's column-strided stores
  s.w = 0;
  s.b = (wimg_bytes + 127) & ~127;
  s.x = s.b + (H / 8) * s.lbo_b;
  s.x_slot_bytes = uses_control ? (Cpad / 8) * s.lbo_b : 0;
  s.stg = (s.x + nx * s.x_slot_bytes + 127) & ~127;  // TMA tensor copies land here: 128-byte aligned
  s.stg_bytes = uses_control ? NR * 16 * C : 0;
  s.prep = (s.stg + nstg * s.stg_bytes + 15) & ~15;
  s.prep_bytes = (NR + 2) * 128 * 4 + 32;
  s.bias = s.prep + 2 * s.prep_bytes;                // [kTcMaxLayers][128] per-layer bias vectors
  s.outst = s.bias + kTcMaxLayers * 512;             // output staging ring: 2 x ([NR][H] floats + 16-byte header)
  s.bars = s.outst + 2 * (NR * H * 4 + 16);
  s.total = s.bars + 8 * (2 + 2 * nx + nstg + 4) + 16;
  return s;
}

// All MMAs of one operand segment (executed warp-uniformly; `leader` is the one issuing lane).
// nk (16-wide K chunks) is a multiple of CH; chunk j feeds chain j % CH, so the chain index is a compile-time
// constant inside the unrolled body.  `fresh`: the region holds no partial sums yet (first segment of a layer).
// TSH / TSL: the hi / lo weight image is resident in TMEM (a_hi / a_lo is then a TMEM address and the MMA takes
// the TS form).  Measured on B200 (tests/cuda/umma_probe.cu, M=128 K=16): an SS-form MMA costs >= 39 cycles
// whatever N <= 32 is (it re-reads the 4 KB A tile from shared memory), a TS-form one 11 (N=16) / 17 (N=32).
// The operands of one segment, computed BEFORE the issuer waits for the segment's inputs: the descriptor and
// address arithmetic (~25 dependent uniform-datapath instructions plus indexed constant loads) used to sit between
// the wake-up and the first MMA, i.e. on the critical path of every layer.
struct SegOps {
  int ts, nk;
  uint32_t a_hi, a_lo, d;
  uint64_t db, da_hi, da_lo;
  uint32_t acc0;
};
__device__ __forceinline__ SegOps seg_ops(int ts, uint32_t a_hi, uint32_t a_lo, uint32_t b_base, int nk, uint32_t lbo_b,
                                          uint32_t d_tmem, bool fresh) {
  SegOps o;
  o.ts = ts; o.nk = nk; o.a_hi = a_hi; o.a_lo = a_lo; o.d = d_tmem;
  o.db = umma_smem_desc(b_base, lbo_b, 128);
  o.da_hi = umma_smem_desc(a_hi, kALbo, kASbo);        // only meaningful for an SS-form image
  o.da_lo = umma_smem_desc(a_lo, kALbo, kASbo);
  o.acc0 = fresh ? 0u : 1u;
  return o;
}

// All MMAs of one operand segment (executed warp-uniformly; `leader` is the one issuing lane).
// nk (16-wide K chunks) is a multiple of CH; chunk j feeds chain j % CH, so the chain index is a compile-time
// constant inside the unrolled body.  acc0 = 0: the region holds no partial sums yet (first segment of a layer).
// ts bit 0 / 1: the hi / lo weight image is resident in TMEM (a_hi / a_lo is then a TMEM address and the MMA takes
// the TS form).  Measured on B200 (tests/cuda/umma_probe.cu, M=128 K=16): an SS-form MMA costs >= 39 cycles
// whatever N <= 32 is (it re-reads the 4 KB A tile from shared memory), a TS-form one 11 (N=16) / 17 (N=32).
template <int N, int CH, int NK>
__device__ __forceinline__ void issue_ts_unrolled(bool leader, const SegOps& o, uint32_t lbo_b) {
  constexpr uint32_t idesc2 = umma_idesc_f16(128, 2 * N), idesc1 = umma_idesc_f16(128, N);
  const uint64_t b_step = (uint64_t)((2 * lbo_b) >> 4);
  if (leader) {
#pragma unroll
    for (int kb = 0; kb < NK; ++kb) {
      const uint32_t acc = kb < CH ? o.acc0 : 1u;
      umma_f16_ts(o.d + AccRegion<N, CH>::a(kb % CH), o.a_hi + 8 * kb, o.db + b_step * kb, idesc2, acc);
      umma_f16_ts(o.d + AccRegion<N, CH>::b(kb % CH), o.a_lo + 8 * kb, o.db + b_step * kb, idesc1, 1u);   // corr columns: the hi product just initialised them
    }
  }
}
template <int N, int CH>
__device__ __forceinline__ void issue_segment(bool leader, const SegOps& o, uint32_t lbo_b) {
  constexpr uint32_t idesc2 = umma_idesc_f16(128, 2 * N), idesc1 = umma_idesc_f16(128, N);
  const uint64_t a_step = (uint64_t)((2 * kALbo) >> 4), b_step = (uint64_t)((2 * lbo_b) >> 4);
  if (o.ts == 3 && o.nk == 8) { issue_ts_unrolled<N, CH, 8>(leader, o, lbo_b); return; }     // H = 128
  if (o.ts == 3 && o.nk == 4) { issue_ts_unrolled<N, CH, 4>(leader, o, lbo_b); return; }     // H = 64, C <= 64
  uint32_t a_hi = o.a_hi, a_lo = o.a_lo, acc = o.acc0;
  uint64_t db = o.db, da_hi = o.da_hi, da_lo = o.da_lo;
  const bool tsh = (o.ts & 1) != 0, tsl = (o.ts & 2) != 0;  // warp-uniform
#pragma unroll 1
  for (int kb = 0; kb < o.nk; kb += CH) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (leader) {
        if (tsh) umma_f16_ts(o.d + AccRegion<N, CH>::a(c), a_hi, db, idesc2, acc);
        else umma_f16(o.d + AccRegion<N, CH>::a(c), da_hi, db, idesc2, acc);
        if (tsl) umma_f16_ts(o.d + AccRegion<N, CH>::b(c), a_lo, db, idesc1, 1u);     // corr columns: always accumulate
        else umma_f16(o.d + AccRegion<N, CH>::b(c), da_lo, db, idesc1, 1u);
      }
      a_hi += 8; a_lo += 8;                              // 8 TMEM columns = 16 fp16 of K
      da_hi += a_step; da_lo += a_step;
      db += b_step;
    }
    acc = 1u;
  }
}

// DIFF = 1: the diffusion is tanh(sigmoid(theta) * nan_to_num(coef * y)) (noise options 3,6,13,17) - the
// form of every proposed model with multiplicative noise; DIFF = 0: generic (runtime-selected) form.
template <int NR, int DIFF, int CH>
__global__ void __launch_bounds__(kTcThreads, 1) snsde_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr int N = NR < 16 ? 16 : NR;              // MMA N (rows padded to >= 16)
  using Acc = AccRegion<N, CH>;                     // CH accumulator chains per product; 2 regions of CH*2N columns
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, C = p.C, Cpad = p.Cpad, NL = p.NL;
  const TcSmem L = tc_smem_layout(p.w_smem_bytes, H, C, Cpad, N, NR, p.nx, p.nstg, p.uses_control);
  const int row0 = blockIdx.x * NR;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  const uint32_t bar_acc = smem_u32(&bars[1]);
  const uint32_t bar_xfull = smem_u32(&bars[2]), bar_xempty = smem_u32(&bars[2 + p.nx]);
  const uint32_t bar_cfull = smem_u32(&bars[2 + 2 * p.nx]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[2 + 2 * p.nx + p.nstg + 4]);
  // Two accumulator regions suffice: layer 0 owns region 0 (the X(t) segment of the NEXT step is issued into it
  // while the last layer's epilogue still reads), every later layer reuses region 1 (its MMAs are only issued
  // after the previous layer's epilogue has drained that region).
  // The weight images that fit are kept in the TMEM columns behind the accumulators (p.img[].tmem_col).
  static_assert(2 * Acc::kCols <= 512, "TMEM budget");
  const uint32_t kTmemCols = (uint32_t)p.tmem_cols;

  // ---- one-time setup: weights -> smem, zero the operand buffers, barriers, TMEM ----
  {
    for (int k = 0; k < p.n_img; ++k) {                 // SS-form images -> shared memory
      if (p.img[k].tmem_col >= 0) continue;
      const uint4* src = reinterpret_cast<const uint4*>(p.wimg + p.img[k].g_off);
      uint4* dst = reinterpret_cast<uint4*>(smem + L.w + p.img[k].s_off);
      for (int i = tid; i < p.img[k].bytes / 16; i += kTcThreads) dst[i] = src[i];
    }
    uint4* z = reinterpret_cast<uint4*>(smem + L.b);
    const int zn = (L.stg - L.b) / 16;
    for (int i = tid; i < zn; i += kTcThreads) z[i] = make_uint4(0, 0, 0, 0);
    float* sb = reinterpret_cast<float*>(smem + L.bias);
    for (int i = tid; i < kTcMaxLayers * 128; i += kTcThreads) {
      const int l = i >> 7, j = i & 127;
      sb[i] = (l < NL && j < H) ? p.vec[p.layer[l].bias + j] : 0.f;
    }
  }
  if (tid == 0) {
    mbar_init(bar_acc, 1);
    for (int i = 0; i < p.nx; ++i) { mbar_init(bar_xfull + 8 * i, kProdWarps); mbar_init(bar_xempty + 8 * i, 1); }
    for (int i = 0; i < p.nstg; ++i) mbar_init(bar_cfull + 8 * i, 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  auto dcol = [&](int l) -> uint32_t { return (uint32_t)((l == 0 ? 0 : 1) * Acc::kCols); };
  // TS-form images -> TMEM: lane m = weight row m, each 32-bit column packs two consecutive K elements, i.e. one
  // 16-byte core-matrix row of the canonical image is 4 columns.  Warp w may only touch lane quadrant w % 4.
  if (warp < kEpiWarps) {
    const int m = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int k = 0; k < p.n_img; ++k) {
      const TcImg im = p.img[k];
      if (im.tmem_col < 0) continue;
      const uint8_t* src = p.wimg + im.g_off + (m >> 3) * kASbo + (m & 7) * 16;
      const int nkc = im.bytes / (2 * (int)kALbo);
      // two chunks per trip: all four 16-byte loads are in flight before the first tcgen05.st (the prologue is one
      // L2 round trip per trip; it is ~8 % of the launch at S = 200)
      for (int kc = (warp >> 2); kc < nkc; kc += 2 * kEpiPerQuad) {
        const int kc2 = kc + kEpiPerQuad;
        const bool two = kc2 < nkc;
        const uint4 lo = *reinterpret_cast<const uint4*>(src + (size_t)(2 * kc) * kALbo);
        const uint4 hi = *reinterpret_cast<const uint4*>(src + (size_t)(2 * kc + 1) * kALbo);
        uint4 lo2 = lo, hi2 = hi;
        if (two) {
          lo2 = *reinterpret_cast<const uint4*>(src + (size_t)(2 * kc2) * kALbo);
          hi2 = *reinterpret_cast<const uint4*>(src + (size_t)(2 * kc2 + 1) * kALbo);
        }
        const uint32_t r[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        tmem_st8(tmem + lane_base + (uint32_t)(im.tmem_col + kc * 8), r);
        if (two) {
          const uint32_t r2[8] = {lo2.x, lo2.y, lo2.z, lo2.w, hi2.x, hi2.y, hi2.z, hi2.w};
          tmem_st8(tmem + lane_base + (uint32_t)(im.tmem_col + kc2 * 8), r2);
        }
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < kEpiWarps) {
    // =========================== EPILOGUE / SDE STATE ===========================
    // thread = (feature h, row group): TMEM lane quadrant = warp % 4, rows [(warp/4)*RT, (warp/4)*RT + RT)
    constexpr int RT = NR / kEpiPerQuad;               // rows per thread
    constexpr int LW = RT < 8 ? RT : 8;                // TMEM load width
    const int h = (warp & 3) * 32 + lane;
    const int rbase = (warp >> 2) * RT;
    const bool act = h < H;
    const TailOp t = p.tail;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint8_t* bact = smem + L.b + (h >> 3) * L.lbo_b + (h & 7) * 2;       // + (r/8)*128 + (r%8)*16 ; lo: + (N/8)*128
    float vmax = 0.f;
    // (A 4x4 lane transpose that turns the four 2-byte operand stores of a thread into one 8-byte store was measured:
    //  the two dependent shuffles cost more than the stores they save - layer epilogues +170..190 cycles.)
    auto write_operand = [&](int r, float v) {
      __half hi, lo;
      split_f16(v, hi, lo);
      vmax = fmaxf(vmax, fabsf(v));                 // range check of the split-fp16 operands: one flag write at the end
      uint8_t* q = bact + (r >> 3) * 128 + (r & 7) * 16;
      *reinterpret_cast<__half*>(q) = hi;
      *reinterpret_cast<__half*>(q + (N / 8) * 128) = lo;
    };
    auto hand_over = [&]() {                 // operands written -> MMA issuer (one arrival per warp)
      tc_fence_before();
      fence_proxy_async_smem();
      named_arrive(kBarIn, kCntIn);
    };
    const float* sbias = reinterpret_cast<const float*>(smem + L.bias) + h;
    const float bias_last = sbias[(NL - 1) * 128];
    // small row counts: diffusion / tanh(y) are evaluated in the shadow of the layer-0 MMAs and kept in
    // registers; RT = 16 has no registers to spare and evaluates them inside the update loop
    constexpr bool PRE = RT <= 8;
    constexpr int NP = PRE ? RT : 1;

    float y[RT], yprev[RT], gv[NP], dg[NP], thy[NP];
    int myslot[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      const int b = min(row0 + rbase + i, p.B - 1);
      y[i] = act ? p.y0[(size_t)b * H + h] : 0.f;
      yprev[i] = y[i];
      myslot[i] = p.row_slot ? p.row_slot[b] : -1;
      if (PRE) gv[i] = dg[i] = thy[i] = 0.f;
      if (act) write_operand(rbase + i, y[i]);
    }
    auto emit = [&](snsde_emit em) {
      if (!act) return;
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const int gr = row0 + rbase + i;
        if (gr >= p.B) continue;
        const float v = em.w_prev * yprev[i] + em.w_curr * y[i];
        if (p.row_slot) {
          if (myslot[i] == em.slot) p.out[(size_t)gr * H + h] = v;
        } else {
          p.out[((size_t)em.slot * p.B + gr) * H + h] = v;
        }
      }
    };
    for (int e = 0; e < p.n_init_emits; ++e) {
      snsde_emit em = p.emits[e];
      em.w_prev = 0.f; em.w_curr = 1.f;
      emit(em);
    }
    hand_over();

    // diffusion of the CURRENT state for the coming step (needs only y and the step's coefficient)
    // Row loops are kept branch-free (model flags are tested OUTSIDE the loops) so that the compiler interleaves
    // the rows' dependent MUFU/FMA chains; with per-row branches each row's ~150-cycle chain ran back to back.
    auto diffusion_rows = [&](const float (&yv)[RT], float cf, float t0, float (&g)[RT], float (&dgy)[RT]) {
      if (DIFF == 1) {
        if (t.milstein) {
#pragma unroll
          for (int i = 0; i < RT; ++i) {
            const float raw = cf * yv[i];
            g[i] = tanh_fast(t.s_theta * nan_to_num_f(raw));
            dgy[i] = ((1.f - g[i] * g[i]) * t.s_theta) * (is_finite_f(raw) ? cf : 0.f);
          }
        } else {                                     // Euler: the derivative is never used
#pragma unroll
          for (int i = 0; i < RT; ++i) {
            g[i] = tanh_fast(t.s_theta * nan_to_num_f(cf * yv[i]));
            dgy[i] = 0.f;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < RT; ++i) diffusion_eval<true>(t, cf, yv[i], t0, g[i], dgy[i]);
      }
    };
    auto prepare_state = [&](float cf, float t0) {
      if (!act || !PRE) return;
      float g[RT], d[RT];
      diffusion_rows(y, cf, t0, g, d);
#pragma unroll
      for (int i = 0; i < NP; ++i) { gv[i] = g[i]; dg[i] = d[i]; }
      if (t.geometric) {
#pragma unroll
        for (int i = 0; i < NP; ++i) thy[i] = tanh_fast(y[i]);
      }
    };

    uint32_t pacc = 0;
    for (int s = 0; s < p.S; ++s) {
      // ---- step data from the prefetch warps (normally already there) ----
      const uint8_t* slot = smem + L.prep + (s & 1) * L.prep_bytes;
      const float* sdw = reinterpret_cast<const float*>(slot);
      named_sync(kBarPFull + (s & 1), kCntPrep);
      TC_TRACE(tid == 0, s, EV_EPI_PFULL);
      const StepInfo si = *reinterpret_cast<const StepInfo*>(slot + (NR + 2) * 512);
      const float add0 = sdw[NR * 128 + h];
      const float cf = sdw[(NR + 1) * 128 + h];
      TC_TRACE(tid == 0, s, EV_EPI_PREPARED);
      for (int l = 0; l < NL; ++l) {
        mbar_wait(bar_acc, pacc);
        pacc ^= 1;
        tc_fence_after();
        TC_TRACE(tid == 0 && l < 2, s, l == 0 ? EV_EPI_ACC0 : EV_EPI_ACC1);
        float vm[RT], vc[RT];
        const uint32_t dreg = tmem + lane_base + dcol(l) + rbase;
#pragma unroll
        for (int c = 0; c < RT; c += LW) {
          float m8[CH][LW], a8[CH][LW];
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            tmem_ldw<LW>(dreg + Acc::a(ch) + c, m8[ch]);
            tmem_ldw<LW>(dreg + Acc::a(ch) + N + c, a8[ch]);
          }
          tmem_ld_wait();
          TC_TRACE(tid == 0 && l < 2 && c == 0, s, l == 0 ? EV_EPI_LD0 : EV_EPI_LD1);
#pragma unroll
          for (int i = 0; i < LW; ++i) {
            float m = m8[0][i], cc = a8[0][i];
#pragma unroll
            for (int ch = 1; ch < CH; ++ch) { m += m8[ch][i]; cc += a8[ch][i]; }
            vm[c + i] = m; vc[c + i] = cc;
          }
        }
        if (l < NL - 1) {
          if (act) {
            const float add = (l == 0) ? add0 : sbias[l * 128];
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              float v = fmaf(vc[i], kLoInv, vm[i]) + add;
              v = v < 0.f ? 0.f : v;                                   // relu feeding the next Linear
              write_operand(rbase + i, v);
            }
          }
        } else if (act) {
          float d[RT], g[RT], dgy[RT], dw[RT];
#pragma unroll
          for (int i = 0; i < RT; ++i) {
            d[i] = fmaf(vc[i], kLoInv, vm[i]) + bias_last;
            dw[i] = sdw[(rbase + i) * 128 + h];
          }
          if (PRE) {
#pragma unroll
            for (int i = 0; i < NP; ++i) { g[i] = gv[i]; dgy[i] = dg[i]; }
            if (t.geometric) {
#pragma unroll
              for (int i = 0; i < NP; ++i) d[i] *= thy[i];
            }
          } else {
            diffusion_rows(y, cf, si.t0, g, dgy);
            if (t.geometric) {
#pragma unroll
              for (int i = 0; i < RT; ++i) d[i] *= tanh_fast(y[i]);
            }
          }
          if (t.clip_drift) {
#pragma unroll
            for (int i = 0; i < RT; ++i) d[i] = tanh_fast(d[i]);
          }
          float yn[RT];
#pragma unroll
          for (int i = 0; i < RT; ++i) yn[i] = __fadd_rn(__fadd_rn(y[i], __fmul_rn(d[i], si.h)), __fmul_rn(g[i], dw[i]));
          if (t.milstein) {
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              const float v2 = __fmul_rn(dw[i], dw[i]) - si.h;
              yn[i] = __fadd_rn(yn[i], 0.5f * ((g[i] * v2) * dgy[i]));
            }
          }
          // Output of this step (torchsde linear_interp between the two states) -> staging ring; the MMA warp sends it
          // to HBM with TMA bulk stores once this hand-over has made it visible to the async proxy.  The epilogue issues
          // no global store in the time loop: its hand-over fence (fence.proxy.async = MEMBAR.ALL.CTA) would wait for the
          // store's round trip to L2 (clock trace: ~430 cycles per step when the emits were plain stores here).
          {
            float* ost = reinterpret_cast<float*>(smem + L.outst + (s & 1) * (NR * H * 4 + 16));
            if (si.n_emits > 0) {
              // per-row final_index capture: only the rows whose slot this is are staged (a handful per step; every
              // shared-memory store in flight lengthens the hand-over fence by ~50 cycles)
#pragma unroll
              for (int i = 0; i < RT; ++i)
                if (p.row_slot == nullptr || myslot[i] == si.first.slot)
                  ost[4 + (rbase + i) * H + h] = si.first.w_prev * y[i] + si.first.w_curr * yn[i];
            }
          }
#pragma unroll
          for (int i = 0; i < RT; ++i) {
            yprev[i] = y[i];
            y[i] = yn[i];
            write_operand(rbase + i, yn[i]);
          }
        }
        hand_over();
        TC_TRACE(tid == 0 && l < 2, s, l == 0 ? EV_EPI_DONE0 : EV_EPI_DONE1);
        // diffusion / tanh of the current state for this step's update: in the shadow of the layer-1 MMAs (the
        // layer-0 window is already filled by the outputs of the previous step)
        if (l == 0) prepare_state(cf, si.t0);
      }
      // ---- in the shadow of the next step's layer-0 MMAs ----
      named_arrive(kBarPEmpty + (s & 1), kCntPrep);                     // slot consumed
      for (int e = 1; e < si.n_emits; ++e) emit(p.emits[si.emit_begin + e]);      // several output times inside one step: rare
      TC_TRACE(tid == 0, s, EV_EPI_SHADOW_END);
    }
    if (vmax > 65504.f) *p.status = 1;              // an operand beyond the fp16 range was saturated (sticky flag)
  } else if (warp == kMmaWarp) {
    // =========================== MMA ISSUER ===========================
    // The whole warp runs this code (descriptor arithmetic stays on the uniform datapath); one elected
    // lane issues the tcgen05 instructions.  A step is NL layer segments plus the X(t) segment of the NEXT
    // step's layer 0 (issued a layer early, into accumulator region 0); all of them go through ONE copy of
    // the issue loop - the kernel's hot code has to stay inside the 32 KB instruction cache.
    const bool leader = elect_one();
    const uint32_t w_base = smem_u32(smem + L.w), b_base = smem_u32(smem + L.b), x_base = smem_u32(smem + L.x);
    const bool has_x = p.uses_control != 0;
    uint32_t xphase = 0;
    int xslot = 0;
    auto layer_ops = [&](int l) {
      const int ts = p.layer[l].ts;
      return seg_ops(ts, ((ts & 1) ? tmem : w_base) + p.layer[l].h_hi, ((ts & 2) ? tmem : w_base) + p.layer[l].h_lo,
                     b_base, p.layer[l].K / 16, L.lbo_b, tmem + dcol(l), !(l == 0 && has_x));
    };
    auto x_ops = [&](int slot) {
      const int ts = p.x_ts;
      return seg_ops(ts, ((ts & 1) ? tmem : w_base) + p.hx_hi, ((ts & 2) ? tmem : w_base) + p.hx_lo,
                     x_base + slot * L.x_slot_bytes, Cpad / 16, L.lbo_b, tmem + dcol(0), true);
    };
    // Outputs of step s: staged by the epilogue before the hand-over that woke this warp; sent with bulk async copies
    // (lane r = row r for the per-row final_index capture, one copy of the whole [rows][H] block for streamed outputs).
    const int my_row_slot = (p.row_slot != nullptr && lane < NR) ? p.row_slot[min(row0 + lane, p.B - 1)] : -1;
    const int valid_rows = min(NR, p.B - row0);
    // first output slot of a step (-1: the step emits nothing), read from the global step / emit tables one step ahead
    auto load_eslot = [&](int s) -> int {
      if (s >= p.S) return -1;
      const int eb = p.steps[s].emit_begin, ee = p.steps[s].emit_end;
      return ee > eb ? p.emits[eb].slot : -1;
    };
    int eslot_next = load_eslot(0);
    auto send_outputs = [&](int s) {
      const uint8_t* ost = smem + L.outst + (s & 1) * (NR * H * 4 + 16);
      const int eslot = eslot_next;
      eslot_next = load_eslot(s + 1);                    // consumed one step later: its latency is never waited for
      if (eslot >= 0) {
        if (p.row_slot != nullptr) {
          if (lane < valid_rows && my_row_slot == eslot)
            bulk_s2g(p.out + (size_t)(row0 + lane) * H, smem_u32(ost + 16 + lane * H * 4), (uint32_t)(H * 4));
        } else if (lane == 0) {
          bulk_s2g(p.out + ((size_t)eslot * p.B + row0) * H, smem_u32(ost + 16), (uint32_t)(valid_rows * H * 4));
        }
      }
      bulk_commit_group();
      bulk_wait_group_read<1>();          // the copies of step s-1 have read their staging slot: the epilogue may reuse it
    };
    auto run_segment = [&](const SegOps& o, uint32_t wait_bar, uint32_t wait_par, uint32_t commit_bar, int s = -1, int ev = -1) {
      if (wait_bar == 0) named_sync(kBarIn, kCntIn);        // operands from the epilogue warps
      else mbar_wait(wait_bar, wait_par);                   // X(t) operand ring
      tc_fence_after();
      TC_TRACE(lane == 0 && ev >= 0, s, ev);
      issue_segment<N, CH>(leader, o, L.lbo_b);
      if (leader) umma_commit(commit_bar);
      __syncwarp();
    };
    if (NL == 2) {
      // One hidden layer (every named model at its headline configuration): the operands of the three segments of a
      // step live in uniform registers for the whole kernel, so nothing but the MMAs follows a wake-up.  (Computing
      // them per segment - indexed constant loads plus ~25 dependent uniform-datapath instructions - cost ~350
      // cycles per segment on the critical path.)
      const SegOps o0 = layer_ops(0), o1 = layer_ops(1);
      const SegOps ox = x_ops(0);
      const uint64_t x_slot_desc = (uint64_t)(L.x_slot_bytes >> 4);
      for (int s = -1; s < p.S; ++s) {
        if (s >= 0) {
          run_segment(o0, 0, 0, bar_acc, s, EV_MMA_WAKE0);
          TC_TRACE(lane == 0, s, EV_MMA_COMMIT0);
          if (s > 0) send_outputs(s - 1);                // in the shadow of the layer-0 MMAs
          run_segment(o1, 0, 0, bar_acc, s, EV_MMA_WAKE1);
          TC_TRACE(lane == 0, s, EV_MMA_COMMIT1);
        }
        if (has_x && s + 1 < p.S) {
          SegOps o = ox;
          o.db += x_slot_desc * (uint64_t)xslot;
          run_segment(o, bar_xfull + 8 * xslot, xphase, bar_xempty + 8 * xslot);
          if (++xslot == p.nx) { xslot = 0; xphase ^= 1; }
          TC_TRACE(lane == 0 && s >= 0, s, EV_MMA_X_DONE);
        }
      }
    } else {
      for (int s = -1; s < p.S; ++s) {
        for (int seg = (s < 0 ? NL : 0); seg <= NL; ++seg) {
          const bool isx = seg == NL;
          if (isx && !(has_x && s + 1 < p.S)) continue;
          // operands first (independent of the data we are about to wait for), then the wait
          if (isx) {
            const SegOps o = x_ops(xslot);
            asm volatile("" ::"r"(o.ts), "r"(o.nk), "r"(o.a_hi), "r"(o.a_lo), "r"(o.d), "l"(o.db), "r"(o.acc0));
            run_segment(o, bar_xfull + 8 * xslot, xphase, bar_xempty + 8 * xslot);
            if (++xslot == p.nx) { xslot = 0; xphase ^= 1; }
          } else {
            const SegOps o = layer_ops(seg);
            asm volatile("" ::"r"(o.ts), "r"(o.nk), "r"(o.a_hi), "r"(o.a_lo), "r"(o.d), "l"(o.db), "r"(o.acc0));
            run_segment(o, 0, 0, bar_acc);
            if (seg == 0 && s > 0) send_outputs(s - 1);
          }
        }
      }
    }
    if (p.S > 0) {                                       // outputs of the last step: its hand-over has no MMA segment behind it
      named_sync(kBarIn, kCntIn);
      send_outputs(p.S - 1);
    }
    bulk_wait_all();
  } else if (warp < kPrepWarp0) {
    // =========================== CONTROL PRODUCER ===========================
    // This role has to deliver one X(t) operand per solver step; measured (clock64 trace) at 3500 cycles per
    // iteration it was THE bottleneck of the kernel once the MMA phases got short: 840 cycles for 8 serialized
    // bulk-copy issues from one warp, 2 x 320 for the barrier tests, 1760 for three dependent evaluate chains.
    // Now: the copies are spread over the three warps, the items of a thread are tabulated once and evaluated
    // interleaved (ILP across up to 4 items, one shared body), and x/3 uses the FMA-corrected reciprocal product
    // (correctly rounded like the division it replaces, no subroutine call).
    if (p.uses_control) {
      const int ptid = tid - 32 * kProdWarp0;            // 0..95
      const int pwarp = warp - kProdWarp0;
      const uint32_t row_bytes = 16u * C;
      const int rows_per_warp = (NR + kProdWarps - 1) / kProdWarps;
      auto fetch = [&](int s, int interval) {           // spline rows of step s -> staging slot
        if (s >= p.S) return;
        const int stg = s % p.nstg;
        const uint32_t bar = bar_cfull + 8 * stg;
        if (p.use_tmap) {                               // ONE tensor copy: box [NR rows][4C floats] at (interval*4C, row0)
          if (ptid == 0) tma_load_2d(smem_u32(smem + L.stg + stg * L.stg_bytes), &p.tmap, interval * 4 * C, row0, bar);
          return;
        }
        const int r = pwarp * rows_per_warp + lane;
        if (lane < rows_per_warp && r < NR) {
          const int b = min(row0 + r, p.B - 1);
          const float* src = p.coeffs + (size_t)b * p.coeff_row_stride + (size_t)interval * 4 * C;
          bulk_g2s(smem_u32(smem + L.stg + stg * L.stg_bytes + r * row_bytes), src, row_bytes, bar);
        }
      };
      auto expect = [&](int s) {                        // one thread announces the bytes of step s (before any copy of it)
        if (ptid == 0 && s < p.S) mbar_arrive_expect_tx(bar_cfull + 8 * (s % p.nstg), row_bytes * NR);
      };
      constexpr int kItems = 4;                         // items evaluated together; more (wide inputs) loop around
      int item_src[kItems], item_dst[kItems];
#pragma unroll
      for (int k = 0; k < kItems; ++k) {
        const int i = ptid + k * kProdThreads;
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
        asm volatile("bar.sync 1, %0;" ::"n"(kProdThreads));
        fetch(s, s < p.S ? p.steps[s].interval : 0);
      }
      int interval_ahead = (p.nstg - 1 < p.S) ? p.steps[p.nstg - 1].interval : 0;
      float frac_cur = p.S > 0 ? p.steps[0].frac : 0.f;
      for (int s = 0; s < p.S; ++s) {
        expect(s + p.nstg - 1);
        asm volatile("bar.sync 1, %0;" ::"n"(kProdThreads));      // all producer warps are done with step s-1; expect_tx is posted
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
        for (int i = ptid + kItems * kProdThreads; i < NR * C; i += kProdThreads) {      // wide inputs / many rows
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
  } else {
    // =========================== STEP PREFETCH (time-only work) ===========================
    // Brownian increments (Philox or table), the folded layer-0 bias, the diffusion coefficient and the
    // step/emit descriptors for step s, written to a 2-deep ring one step ahead of the epilogue.
    const int h = tid - 32 * kPrepWarp0;
    const bool act = h < H;
    const TailOp t = p.tail;
    const float c0 = act ? p.vec[p.layer[0].bias + h] : 0.f;
    const float csin = (p.c_sin >= 0 && act) ? p.vec[p.c_sin + h] : 0.f;
    const float ccos = (p.c_cos >= 0 && act) ? p.vec[p.c_cos + h] : 0.f;
    float coef = t.coef_scalar;
    if (t.coef_src == CO_IMG && act) coef = p.vec[p.coef_vec + h];
    // Global rows of this CTA start at gb0; Philox yields 4 normals for the aligned row quad gb >> 2, so the NR
    // rows span nq = (gb0 % 4 + NR + 3) / 4 quads.  One quad per iteration of a NOT unrolled loop (I-cache).
    const unsigned long long gb0 = p.row_offset + (unsigned long long)row0;
    const int lane0 = (int)(gb0 & 3ull);
    const int nq = (lane0 + NR + 3) >> 2;
    if (t.coef_src == CO_VBUF) asm volatile("griddepcontrol.wait;" ::: "memory");   // a_tab comes from the tables kernel (PDL)
    for (int s = 0; s < p.S; ++s) {
      const snsde_step st = p.steps[s];
      uint8_t* slot = smem + L.prep + (s & 1) * L.prep_bytes;
      float* sdw = reinterpret_cast<float*>(slot);
      float cf = coef;
      if (t.coef_src == CO_VBUF && act) cf = p.a_tab[(size_t)s * H + h];
      StepInfo si;
      si.h = st.h; si.t0 = st.t0; si.n_emits = st.emit_end - st.emit_begin; si.emit_begin = st.emit_begin;
      si.first.slot = 0; si.first.w_prev = 0.f; si.first.w_curr = 0.f;
      if (h == 0 && si.n_emits > 0) si.first = p.emits[st.emit_begin];
      if (s >= 2) named_sync(kBarPEmpty + (s & 1), kCntPrep);
      if (act) {
        if (p.dW != nullptr) {
#pragma unroll 4
          for (int r = 0; r < NR; ++r) sdw[r * 128 + h] = p.dW[((size_t)s * p.B + min(row0 + r, p.B - 1)) * H + h];
        } else {
#pragma unroll 1
          for (int q = 0; q < nq; ++q) {
            float nrm[4];
            philox_normals4(p.seed, (uint32_t)h, (uint32_t)((gb0 >> 2) + (unsigned long long)q), (uint32_t)s, nrm);
            const int rq = q * 4 - lane0;               // CTA-local row of the quad's first normal
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if ((unsigned)(rq + k) < (unsigned)NR) sdw[(rq + k) * 128 + h] = __fmul_rn(nrm[k], st.sqrt_h);
          }
        }
        sdw[NR * 128 + h] = fmaf(st.cos_t0, ccos, fmaf(st.sin_t0, csin, c0));
        sdw[(NR + 1) * 128 + h] = cf;
      }
      if (h == 0) *reinterpret_cast<StepInfo*>(slot + (NR + 2) * 512) = si;
      named_arrive(kBarPFull + (s & 1), kCntPrep);

    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

// Per-step coefficient of the row-independent noise networks: a_tab[s][h]
// (noise_t(time_features), neuralsde.py:271-282; evaluated ONCE per step, not once per batch row).
__global__ void __launch_bounds__(256) snsde_tc_tables_kernel(const float* __restrict__ vec, TcNoiseNet nn, int H,
                                                              const snsde_step* __restrict__ steps, float* __restrict__ a_tab) {
  __shared__ float h1[256];
  // programmatic dependent launch: the solve kernel may start its prologue (weights -> TMEM/smem, barriers) now;
  // it executes griddepcontrol.wait before the first read of a_tab
  asm volatile("griddepcontrol.launch_dependents;");
  const int s = blockIdx.x, h = threadIdx.x;
  const snsde_step st = steps[s];
  float v = 0.f;
  if (h < H) {
    v = fmaf(st.cos_t0, vec[nn.w1t + H + h], fmaf(st.sin_t0, vec[nn.w1t + h], vec[nn.b1 + h]));
    if (nn.kind == 2) v = v < 0.f ? 0.f : v;
  }
  h1[h] = v;
  __syncthreads();
  if (h >= H) return;
  if (nn.kind == 2) {
    float acc = vec[nn.b2 + h];
    for (int k = 0; k < H; ++k) acc = fmaf(h1[k], vec[nn.w2t + k * H + h], acc);
    v = acc < 0.f ? 0.f : acc;
  }
  a_tab[(size_t)s * H + h] = v;
}

// =================================== host side ===================================================
static thread_local std::string g_reason = "";
const char* tc_unsupported_reason() { return g_reason.c_str(); }

static bool is_time_opt(int io) { return io >= 3 && io <= 6; }
static bool is_emb_opt(int io) { return io == 2 || io == 4 || io == 6; }

// Where each operand image lives for one launch shape: TMEM columns behind the two accumulator regions while
// they last (layers first - they are on the critical path - then the control segment), shared memory otherwise.
// Fills p.img / the layers' handles / p.tmem_cols and returns the bytes of the shared-memory weight area.
static int tc_place(TcParams& p, int N, int CH, bool use_tmem) {
  int col = 2 * CH * 2 * N, soff = 0;
  p.n_img = 0;
  auto place = [&](int g_off, int K, int& handle, int& ts, int bit) {
    TcImg im;
    im.g_off = g_off; im.bytes = (K / 8) * (int)kALbo;
    const int need = K / 2;
    if (use_tmem && col + need <= 512) { im.tmem_col = col; im.s_off = -1; handle = col; ts |= bit; col += need; }
    else { im.tmem_col = -1; im.s_off = soff; handle = soff; soff += im.bytes; }
    p.img[p.n_img++] = im;
  };
  for (int l = 0; l < p.NL; ++l) {
    p.layer[l].ts = 0;
    place(p.layer[l].a_hi, p.layer[l].K, p.layer[l].h_hi, p.layer[l].ts, 1);
    place(p.layer[l].a_lo, p.layer[l].K, p.layer[l].h_lo, p.layer[l].ts, 2);
  }
  p.x_ts = 0; p.hx_hi = p.hx_lo = 0;
  if (p.uses_control) {
    place(p.ax_hi, p.Cpad, p.hx_hi, p.x_ts, 1);
    place(p.ax_lo, p.Cpad, p.hx_lo, p.x_ts, 2);
  }
  p.tmem_cols = col <= 128 ? 128 : (col <= 256 ? 256 : 512);
  p.w_smem_bytes = soff;
  return soff;
}
static bool tc_env_no_tmem() { return getenv("SNSDE_TC_NO_TMEM") != nullptr; }     // testing aid: every image SS-form
static int tc_env_chains() { const char* e = getenv("SNSDE_TC_CH"); return (e && atoi(e) == 2) ? 2 : 1; }

bool tc_supported(const snsde_model_desc& d, int cc_major, int smem_optin) {
  const int no = d.noise_option, io = d.input_option;
  if (cc_major != 10) { g_reason = "needs an sm_100 device"; return false; }
  if (d.family != SNSDE_FAMILY_BENCHMARK) { g_reason = "the tutorial and LatentSDE families run on the FMA kernel"; return false; }
  if (io == 0) { g_reason = "input_option 0 (control only) runs on the FMA kernel"; return false; }
  if (no == 14 || no == 15 || no == 18 || no == 19) { g_reason = "state-network noise options run on the FMA kernel"; return false; }
  if (d.hidden != d.hidden_hidden) { g_reason = "needs hidden_hidden == hidden"; return false; }
  // K chunks are issued in pairs, one per accumulator chain
  if (d.hidden % 32 || d.hidden < 32 || d.hidden > 128) { g_reason = "needs hidden in {32,64,96,128}"; return false; }
  if (d.num_hidden_layers + 1 > kTcMaxLayers) { g_reason = "too many hidden layers"; return false; }
  const int Cpad = (d.input_channels + 31) & ~31;
  if (Cpad > 256) { g_reason = "too many input channels for the resident kernel"; return false; }
  TcParams q;                                         // smallest launch shape: 8 rows per CTA, one accumulator chain
  memset(&q, 0, sizeof(q));
  q.NL = d.num_hidden_layers + 1; q.uses_control = is_emb_opt(io); q.Cpad = Cpad;
  for (int l = 0; l < q.NL; ++l) q.layer[l].K = d.hidden;
  const int wsm = tc_place(q, 16, 1, !tc_env_no_tmem());
  const TcSmem L = tc_smem_layout(wsm, d.hidden, d.input_channels, Cpad, 16, 8, 2, 2, is_emb_opt(io));
  if (L.total > smem_optin) { g_reason = "weights + operand buffers exceed shared memory"; return false; }
  return true;
}

namespace {
struct TcImage {
  std::vector<uint8_t> bytes;
  std::vector<float> vec;
  float max_abs = 0.f;
  // W: [rows][K] row-major doubles (rows <= 128); appends canonical hi and lo' images padded to Kpad
  void add_matrix(const std::vector<double>& W, int rows, int K, int Kpad, int& off_hi, int& off_lo) {
    const size_t sz = (size_t)(Kpad / 8) * kALbo;
    off_hi = (int)bytes.size(); bytes.resize(bytes.size() + sz, 0);
    off_lo = (int)bytes.size(); bytes.resize(bytes.size() + sz, 0);
    for (int m = 0; m < rows; ++m)
      for (int k = 0; k < K; ++k) {
        const float w = (float)W[(size_t)m * K + k];
        max_abs = std::max(max_abs, fabsf(w));
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn((w - __half2float(hi)) * kLoScale);
        const size_t o = (size_t)(m / 8) * kASbo + (size_t)(k / 8) * kALbo + (m % 8) * 16 + (k % 8) * 2;
        memcpy(&bytes[off_hi + o], &hi, 2);
        memcpy(&bytes[off_lo + o], &lo, 2);
      }
  }
  int add_vec(const std::vector<double>& v) {
    const int off = (int)vec.size();
    for (double x : v) vec.push_back((float)x);
    while (vec.size() % 128) vec.push_back(0.f);
    return off;
  }
};
}  // namespace

int tc_set_weights(TcPlan& tc, const snsde_model_desc& d, const Program& pg, const float* blob, int num_sms, int smem_optin,
                   cudaStream_t stream) {
  const int C = d.input_channels, H = d.hidden, L = d.num_hidden_layers, io = d.input_option, no = d.noise_option;
  const int tau = is_time_opt(io) ? 2 : 0;
  const int Cpad = (C + 31) & ~31;          // whole pairs of 16-wide K chunks (two accumulator chains)
  const float* q = blob;
  auto take = [&](size_t n) { const float* r = q; q += n; return r; };
  const float* Wi = take((size_t)H * C); const float* bi = take(H);
  const float* Win = take((size_t)H * (H + tau)); const float* bin = take(H);
  const float *We = nullptr, *be = nullptr;
  if (is_emb_opt(io)) { We = take((size_t)H * 2 * H); be = take(H); }
  std::vector<const float*> Wl(L - 1), bl(L - 1);
  for (int l = 0; l < L - 1; ++l) { Wl[l] = take((size_t)H * H); bl[l] = take(H); }
  const float* Wo = take((size_t)H * H); const float* bo = take(H);
  take(1);                                                      // theta (already folded into pg.tail.s_theta)

  TcImage img;
  TcParams& P = tc.proto;
  memset(&P, 0, sizeof(P));
  P.H = H; P.C = C; P.Cpad = Cpad; P.NL = L + 1; P.uses_control = is_emb_opt(io);
  P.c_sin = P.c_cos = P.coef_vec = -1;
  P.tail = pg.tail;

  // ---- layer 0 (collapsed in double precision) ----
  std::vector<double> W0((size_t)H * H), W0x, c0(H), cs(H, 0.0), cc(H, 0.0);
  if (is_emb_opt(io)) {
    W0x.assign((size_t)H * C, 0.0);
    for (int i = 0; i < H; ++i) {
      double acc0 = be[i];
      for (int j = 0; j < H; ++j) {
        const double e1 = We[(size_t)i * 2 * H + j], e2 = We[(size_t)i * 2 * H + H + j];
        acc0 += e1 * bin[j] + e2 * bi[j];
        if (tau) { cs[i] += e1 * Win[(size_t)j * (H + tau)]; cc[i] += e1 * Win[(size_t)j * (H + tau) + 1]; }
        for (int k = 0; k < H; ++k) W0[(size_t)i * H + k] += e1 * Win[(size_t)j * (H + tau) + tau + k];
        for (int c = 0; c < C; ++c) W0x[(size_t)i * C + c] += e2 * Wi[(size_t)j * C + c];
      }
      c0[i] = acc0;
    }
  } else {
    for (int i = 0; i < H; ++i) {
      c0[i] = bin[i];
      if (tau) { cs[i] = Win[(size_t)i * (H + tau)]; cc[i] = Win[(size_t)i * (H + tau) + 1]; }
      for (int k = 0; k < H; ++k) W0[(size_t)i * H + k] = Win[(size_t)i * (H + tau) + tau + k];
    }
  }
  img.add_matrix(W0, H, H, H, P.layer[0].a_hi, P.layer[0].a_lo);
  P.layer[0].K = H;
  P.layer[0].bias = img.add_vec(c0);
  if (tau) { P.c_sin = img.add_vec(cs); P.c_cos = img.add_vec(cc); }
  if (P.uses_control) img.add_matrix(W0x, H, C, Cpad, P.ax_hi, P.ax_lo);
  auto as_double = [&](const float* W, size_t n) { return std::vector<double>(W, W + n); };
  for (int l = 0; l < L - 1; ++l) {
    img.add_matrix(as_double(Wl[l], (size_t)H * H), H, H, H, P.layer[1 + l].a_hi, P.layer[1 + l].a_lo);
    P.layer[1 + l].K = H;
    P.layer[1 + l].bias = img.add_vec(as_double(bl[l], H));
  }
  img.add_matrix(as_double(Wo, (size_t)H * H), H, H, H, P.layer[L].a_hi, P.layer[L].a_lo);
  P.layer[L].K = H;
  P.layer[L].bias = img.add_vec(as_double(bo, H));

  // ---- diffusion coefficients ----
  tc.noise.kind = 0;
  if (no >= 1 && no <= 3) take(1);                              // exp(sigma) already in pg.tail.coef_scalar
  if (no >= 4 && no <= 6) {
    const float* sd = take(H);
    std::vector<double> e(H);
    for (int j = 0; j < H; ++j) e[j] = expf(sd[j]);
    P.coef_vec = img.add_vec(e);
  }
  if (no == 12 || no == 13 || no == 16 || no == 17) {
    const float* W1 = take((size_t)H * 2); const float* b1 = take(H);
    std::vector<double> w1t(2 * (size_t)H);
    for (int j = 0; j < H; ++j) { w1t[j] = W1[2 * j]; w1t[H + j] = W1[2 * j + 1]; }
    tc.noise.kind = 1;
    tc.noise.w1t = img.add_vec(w1t);
    tc.noise.b1 = img.add_vec(as_double(b1, H));
    if (no >= 16) {
      const float* W2 = take((size_t)H * H); const float* b2 = take(H);
      std::vector<double> w2t((size_t)H * H);
      for (int j = 0; j < H; ++j)
        for (int k = 0; k < H; ++k) w2t[(size_t)k * H + j] = W2[(size_t)j * H + k];
      tc.noise.kind = 2;
      tc.noise.w2t = img.add_vec(w2t);
      tc.noise.b2 = img.add_vec(as_double(b2, H));
    }
  }
  if (!(img.max_abs < 6.0e4f)) { g_reason = "a weight exceeds the fp16 range of the split-precision operands"; return SNSDE_ERR_UNSUPPORTED; }

  if ((int)img.bytes.size() > tc.wimg_bytes) {
    cudaFree(tc.d_wimg); tc.d_wimg = nullptr;
    if (cudaMalloc(&tc.d_wimg, img.bytes.size()) != cudaSuccess) { g_reason = "cudaMalloc failed"; return SNSDE_ERR_CUDA; }
  }
  if ((int)img.vec.size() > tc.vec_floats) {
    cudaFree(tc.d_vec); tc.d_vec = nullptr;
    if (cudaMalloc(&tc.d_vec, img.vec.size() * sizeof(float)) != cudaSuccess) { g_reason = "cudaMalloc failed"; return SNSDE_ERR_CUDA; }
  }
  tc.wimg_bytes = (int)img.bytes.size();
  tc.vec_floats = (int)img.vec.size();
  cudaMemcpyAsync(tc.d_wimg, img.bytes.data(), img.bytes.size(), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(tc.d_vec, img.vec.data(), img.vec.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
  P.wimg = tc.d_wimg; P.wimg_bytes = tc.wimg_bytes; P.vec = tc.d_vec;
  tc.num_sms = num_sms; tc.smem_optin = smem_optin;
  tc.atab_valid = false;                    // the noise tables depend on the weights
  tc.ready = true;
  return SNSDE_OK;
}

template <int NR, int DIFF, int CH>
static cudaError_t tc_launch_one(const TcParams& p, int grid, size_t smem, cudaStream_t stream, bool pdl) {
  // pdl: the launch follows the tables kernel - overlap the prologue with it (programmatic dependent launch)
  auto kern = snsde_tc_kernel<NR, DIFF, CH>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}
// coefficients [B][(K-1)*4C] fp32 (row stride in floats), box [NR][4C]
static bool make_coeff_tmap(CUtensorMap* tm, const float* coeffs, long long row_stride, int B, int n_intervals, int C, int NR) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || getenv("SNSDE_TC_NO_TMAP") != nullptr) return false;
  if (4 * C > 256 || NR > 256 || ((row_stride * 4) & 15) || ((uintptr_t)coeffs & 15)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)n_intervals * 4 * C, (cuuint64_t)B};
  const cuuint64_t strides[1] = {(cuuint64_t)row_stride * 4};
  const cuuint32_t box[2] = {(cuuint32_t)(4 * C), (cuuint32_t)NR};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)coeffs, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t tc_forward(TcPlan& tc, const TcForwardArgs& a, cudaStream_t stream, int* n_launches) {
  TcParams p = tc.proto;
  p.use_tmap = 0;
  p.coeffs = a.coeffs; p.coeff_row_stride = a.coeff_row_stride; p.y0 = a.y0; p.B = a.B;
  p.steps = a.steps; p.S = a.S; p.emits = a.emits; p.n_init_emits = a.n_init_emits; p.n_out = a.n_out;
  p.row_slot = a.row_slot; p.dW = a.dW; p.seed = a.seed; p.row_offset = a.row_offset; p.out = a.out;
  p.status = a.status;
  *n_launches = 0;
  bool pdl = false;
  if (tc.noise.kind != 0 && a.S > 0) {
    unsigned long long key = 1469598103934665603ull ^ (unsigned long long)a.S;
    for (int s2 = 0; s2 < a.S; ++s2) {
      unsigned int w2[2];
      memcpy(&w2[0], &a.steps_host[s2].sin_t0, 4); memcpy(&w2[1], &a.steps_host[s2].cos_t0, 4);
      key = (key ^ w2[0]) * 1099511628211ull;
      key = (key ^ w2[1]) * 1099511628211ull;
    }
    if (!(tc.atab_valid && tc.atab_key == key) || getenv("SNSDE_NO_ATAB_CACHE") != nullptr) {
      if (a.S * p.H > tc.atab_cap) {
        cudaFree(tc.d_atab); tc.d_atab = nullptr;
        cudaError_t e = cudaMalloc(&tc.d_atab, sizeof(float) * (size_t)a.S * p.H);
        if (e != cudaSuccess) return e;
        tc.atab_cap = a.S * p.H;
      }
      snsde_tc_tables_kernel<<<a.S, 128, 0, stream>>>(tc.d_vec, tc.noise, p.H, a.steps, tc.d_atab);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
      *n_launches += 1;
      pdl = getenv("SNSDE_NO_PDL") == nullptr;
      tc.atab_key = key; tc.atab_valid = true;
    }
  }
  p.a_tab = tc.d_atab;
  p.dbg = nullptr;
  static long long* s_dbg = nullptr;
  if (getenv("SNSDE_TC_TRACE") != nullptr && a.S > 0) {
    if (s_dbg == nullptr) cudaMalloc(&s_dbg, sizeof(long long) * 16 * 4096);
    cudaMemsetAsync(s_dbg, 0, sizeof(long long) * 16 * 4096, stream);
    if (a.S <= 4096) p.dbg = s_dbg;
  }
  // rows per CTA: the fewest that still covers the batch in one wave; then whatever shared memory allows
  int NR = 8;
  while (NR < 32 && (a.B + NR - 1) / NR > tc.num_sms) NR *= 2;
  const int CH = tc_env_chains();
  const bool use_tmem = !tc_env_no_tmem();
  TcSmem L;
  for (;;) {
    const int N = NR < 16 ? 16 : NR;
    const int wsm = tc_place(p, N, CH, use_tmem);
    bool ok = false;
    for (int cfg = 0; cfg < 3 && !ok; ++cfg) {
      p.nx = cfg == 0 ? 4 : 2;
      p.nstg = cfg == 2 ? 2 : 4;
      L = tc_smem_layout(wsm, p.H, p.C, p.Cpad, N, NR, p.nx, p.nstg, p.uses_control);
      ok = L.total <= tc.smem_optin;
    }
    if (ok) break;
    if (NR == 8) return cudaErrorInvalidConfiguration;
    NR /= 2;
  }
  const int grid = (a.B + NR - 1) / NR;
  if (p.uses_control && a.S > 0)
    p.use_tmap = make_coeff_tmap(&p.tmap, a.coeffs, a.coeff_row_stride, a.B, a.n_knots - 1, p.C, NR) ? 1 : 0;
  cudaError_t e;
  const bool fast_diff = p.tail.bounded && p.tail.special == SP_NONE && p.tail.mult == MU_Y;
  // Accumulator chains per product: with the weights in TMEM the MMA phase is short and one chain is best (fewer
  // TMEM loads/adds in the epilogue); two chains stay selectable (SNSDE_TC_CH=2) for experiments.
  switch ((CH == 2 ? 128 : 0) + NR * 2 + (fast_diff ? 1 : 0)) {
    case 16: e = tc_launch_one<8, 0, 1>(p, grid, L.total, stream, pdl); break;
    case 17: e = tc_launch_one<8, 1, 1>(p, grid, L.total, stream, pdl); break;
    case 32: e = tc_launch_one<16, 0, 1>(p, grid, L.total, stream, pdl); break;
    case 33: e = tc_launch_one<16, 1, 1>(p, grid, L.total, stream, pdl); break;
    case 64: e = tc_launch_one<32, 0, 1>(p, grid, L.total, stream, pdl); break;
    case 65: e = tc_launch_one<32, 1, 1>(p, grid, L.total, stream, pdl); break;
    case 128 + 16: e = tc_launch_one<8, 0, 2>(p, grid, L.total, stream, pdl); break;
    case 128 + 17: e = tc_launch_one<8, 1, 2>(p, grid, L.total, stream, pdl); break;
    case 128 + 32: e = tc_launch_one<16, 0, 2>(p, grid, L.total, stream, pdl); break;
    case 128 + 33: e = tc_launch_one<16, 1, 2>(p, grid, L.total, stream, pdl); break;
    case 128 + 64: e = tc_launch_one<32, 0, 2>(p, grid, L.total, stream, pdl); break;
    default: e = tc_launch_one<32, 1, 2>(p, grid, L.total, stream, pdl); break;
  }
  if (e == cudaSuccess) *n_launches += 1;
  if (p.dbg != nullptr && e == cudaSuccess) {             // debug only: synchronous dump of the trace
    std::vector<long long> hbuf((size_t)16 * a.S);
    cudaStreamSynchronize(stream);
    cudaMemcpy(hbuf.data(), p.dbg, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    FILE* f = fopen(getenv("SNSDE_TC_TRACE"), "w");
    if (f) {
      for (int s2 = 0; s2 < a.S; ++s2) {
        for (int k = 0; k < 16; ++k) fprintf(f, "%lld%c", hbuf[(size_t)s2 * 16 + k], k == 15 ? '\n' : ' ');
      }
      fclose(f);
    }
  }
  return e;
}

void tc_release(TcPlan& tc) {
  cudaFree(tc.d_wimg); cudaFree(tc.d_vec); cudaFree(tc.d_atab);
  tc.d_wimg = nullptr; tc.d_vec = nullptr; tc.d_atab = nullptr; tc.ready = false;
}

}  // namespace snsde
