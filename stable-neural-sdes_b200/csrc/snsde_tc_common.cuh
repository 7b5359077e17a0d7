// Device helpers shared by the two tcgen05 kernels (snsde_tc.cu: weights resident; snsde_tcg_kernel.cuh: general).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>
#include "snsde_common.cuh"
#include "snsde_tc_ptx.cuh"

namespace snsde {

using namespace ptx;

constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

// clock64 trace (debug aid, enabled by the SNSDE_TC_TRACE environment variable): 16 events per step, CTA 0.
enum TraceEv { EV_EPI_ACC0 = 0, EV_EPI_LD0, EV_EPI_DONE0, EV_EPI_ACC1, EV_EPI_LD1, EV_EPI_DONE1, EV_EPI_SHADOW_END,
               EV_MMA_WAKE0, EV_MMA_COMMIT0, EV_MMA_WAKE1, EV_MMA_COMMIT1, EV_MMA_X_DONE, EV_PREP_DONE, EV_PROD_DONE,
               EV_EPI_PFULL, EV_EPI_PREPARED };
// The trace code is compiled in only with -DSNSDE_TC_TRACE_BUILD (SNSDE_TRACE_BUILD=1 python .../build.py): the hot code of
// the persistent kernels has to fit the 32 KB instruction cache and every event costs ~10 instructions.
#ifdef SNSDE_TC_TRACE_BUILD
#define TC_TRACE(cond, step, ev) do { if (p.dbg != nullptr && blockIdx.x == 0 && (step) >= 0 && (cond)) p.dbg[(size_t)(step) * 16 + (ev)] = clock64(); } while (0)
#else
#define TC_TRACE(cond, step, ev) do { } while (0)
#endif

// Per-step broadcast block written by the step-prefetch warps (ring of 2).
struct StepInfo {
  float h, t0;
  int n_emits, emit_begin;
  snsde_emit first;          // the first emit of the step (almost every step has at most one)
};

// fp16 split of an fp32 value: hi (saturating) and the 2^11-scaled residual
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  hi = __ushort_as_half(h);
  lo = __float2half_rn((v - __half2float(hi)) * kLoScale);
}

// TMEM accumulator region of one layer: per chain c the 2N columns [main | corr] at c*2N.
//   hi product  Whi x [ahi ; alo']  (N' = 2N)  -> [main | corr]
//   lo product  Wlo' x ahi          (N)        -> accumulated into the SAME corr columns
// so the epilogue reads two values per element, not three: the TMEM read port (~64 B/cycle) bounds the
// accumulator load of a phase (3 x NR columns x 128 lanes x 4 B = 12 KB was ~200 cycles at NR = 8).
// Back-to-back MMAs into the same columns cost nothing extra on the in-order tensor pipe (measured with
// tests/cuda/umma_probe.cu: 10.8 cycles per TS-form MMA whether or not consecutive ones share the accumulator).
// CH > 1 spreads the K chunks over independent chains summed in the epilogue (kept for experiments).
template <int N, int CH>
struct AccRegion {
  static constexpr int kCols = CH * 2 * N;
  __device__ static constexpr uint32_t a(int c) { return (uint32_t)(c * 2 * N); }
  __device__ static constexpr uint32_t b(int c) { return (uint32_t)(c * 2 * N + N); }
};

}  // namespace snsde
