// Device helpers shared by the fp32 FMA forward kernels (snsde_fma.cu) and the reverse sweep (snsde_bwd.cu).
#pragma once
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

// Hand-off inside one row group.
__device__ __forceinline__ void group_sync(int gid, int nw) {
  if (nw == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(nw * 32) : "memory");
}

struct GroupSmem {
  float* base;           // [kNumRowBufs][R][ld]
  int row_buf_floats;    // R * ld
  float* stage0;         // 2 x NP spline stages of stage_floats each
  int stage_floats;
  __device__ __forceinline__ float* buf(int id) const { return base + id * row_buf_floats; }
};

// Copies the weight image into shared memory (whole CTA, 16-byte accesses); the caller synchronises.
__device__ __forceinline__ void stage_weights(float* __restrict__ dst, const float* __restrict__ src, int n_floats) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  const int n4 = n_floats >> 2;
#pragma unroll 4
  for (int i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = s4[i];
}

// acc[r] += sum_k src[r][k] * w[k * stride]      (w already offset by the thread's output feature)
// The weight pointer walks by `stride`; its address space (shared when the image is staged, global otherwise) is a
// compile-time property of the calling kernel, so these are plain LDS / LDG with immediate offsets.
template <int ROWS>
__device__ __forceinline__ void dot_accumulate(float (&acc)[ROWS], const float* __restrict__ src, int ld,
                                               const float* __restrict__ w, int K, int stride) {
  const float* wk = w;
  int k = 0;
  if (ROWS == 1) {
    // A single-row group is pure latency: batch 16 weight reads and 4 activation reads, THEN the 16 dependent FMAs
    // (same k order as the general path below, so a row's result does not depend on how the batch was grouped).
    const int K16 = K & ~15;
    for (; k < K16; k += 16) {
      float wv[16];
      float4 av[4];
#pragma unroll
      for (int i = 0; i < 16; ++i) wv[i] = wk[i * stride];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(src + k + 4 * i);
      wk += 16 * stride;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[0] = fmaf(av[i].x, wv[4 * i + 0], acc[0]);
        acc[0] = fmaf(av[i].y, wv[4 * i + 1], acc[0]);
        acc[0] = fmaf(av[i].z, wv[4 * i + 2], acc[0]);
        acc[0] = fmaf(av[i].w, wv[4 * i + 3], acc[0]);
      }
    }
  }
  const int K4 = K & ~3;
  const int s2 = 2 * stride, s3 = 3 * stride, s4 = 4 * stride;
#pragma unroll 2
  for (; k < K4; k += 4, wk += s4) {
    const float w0 = wk[0], w1 = wk[stride], w2 = wk[s2], w3 = wk[s3];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(src + r * ld + k);
      acc[r] = fmaf(a.x, w0, acc[r]);
      acc[r] = fmaf(a.y, w1, acc[r]);
      acc[r] = fmaf(a.z, w2, acc[r]);
      acc[r] = fmaf(a.w, w3, acc[r]);
    }
  }
  for (; k < K; ++k, wk += stride) {
    const float wv = wk[0];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(src[r * ld + k], wv, acc[r]);
  }
}

// acc[r] += sum_k src[r][k] * wrow[k]   - the TRANSPOSED product: thread j walks row j of an [in][out] image
// (contiguous), used by the Milstein vjp through the noise network.
template <int ROWS>
__device__ __forceinline__ void dot_rows_T(float (&acc)[ROWS], const float* __restrict__ src, int ld,
                                           const float* __restrict__ wrow, int K) {
  for (int k = 0; k < K; ++k) {
    const float wk = __ldg(wrow + k);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(src[r * ld + k], wk, acc[r]);
  }
}

struct TimePoint { float t, sin_t, cos_t; };

// Pre-activation (ACT = false) or activated output of one dense op for the R rows of a group, feature j.
// W = base of the weight image in the address space the kernel was compiled for.
template <int ROWS, bool ACT = true>
__device__ __forceinline__ void dense_eval(float (&acc)[ROWS], const DenseOp& op, const float* __restrict__ W,
                                           const GroupSmem& sm, int ld, const TimePoint& tp, int j) {
  float init = op.b_off >= 0 ? W[op.b_off + j] : 0.f;
  if (op.tmode == TM_SINCOS) {
    const float* tw = W + op.tw_off;
    init = fmaf(tp.cos_t, tw[op.N + j], fmaf(tp.sin_t, tw[j], init));
  } else if (op.tmode == TM_RAW) {
    init = fmaf(tp.t, W[op.tw_off + j], init);
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) acc[r] = init;
  if (op.src >= 0) dot_accumulate<ROWS>(acc, sm.buf(op.src), ld, W + op.w_off + j, op.K, op.N);
  if (op.src2 >= 0) dot_accumulate<ROWS>(acc, sm.buf(op.src2), ld, W + op.w2_off + j, op.K2, op.N);
  if (ACT) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = act_apply(acc[r], op.act);
  }
}

}  // namespace snsde
