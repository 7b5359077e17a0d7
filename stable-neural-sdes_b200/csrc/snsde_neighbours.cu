// The seam's neighbours: the producer of the initial state and the read-out head, as engine kernels so
// that an inference forward of the reference wrappers runs without a PyTorch launch between the host buffers and the
// prediction.
//
//   snsde_initial_state   z0 = initial_network(X(times[0]))
//        /root/reference/benchmark_classification/models_sde/neuralsde.py:63-69 (`_prepare_initial_state`), same code in
//        benchmark_forecasting/models_sde/neuralsde.py:137-143 and torch-ists nsde_model.py:57-61.
//   snsde_readout_head    pred = Linear2(relu(bn(Linear1(pre(z)))))  in EVAL mode:
//        classification  Linear -> BatchNorm1d -> ReLU -> Dropout -> Linear      (neuralsde.py:59-61, applied at :119)
//        forecasting     Linear -> ReLU -> Linear                                 (benchmark_forecasting/...:133-136, :185)
//        torch-ists      Tanh -> Linear -> ReLU -> Linear                         (nsde_model.py:52-55, :83)
//      BatchNorm1d in eval mode is the per-feature affine map (x - running_mean) / sqrt(running_var + eps) * gamma + beta,
//      passed in as scale / shift; Dropout in eval mode is the identity.  Training mode (batch statistics, dropout
//      masks) stays in PyTorch.
//
// Both are tiny next to the solve (2 C H resp. 2 H (H + O) FLOP per row): thread j = output feature j, ROWS rows per
// CTA staged in shared memory, weights read through L1/L2 in their nn.Linear layout via a transposed walk.
#include <cuda_runtime.h>
#include <math.h>

#include "snsde_host.cuh"

namespace snsde {

constexpr int kNbRows = 8;

// X(t) of one spline row at (interval, frac): a + (b + (two_c/2 + three_d*frac/3)*frac)*frac  (torchcde op order)
__device__ __forceinline__ float spline_value(const float* __restrict__ row, int C, int c, float frac) {
  float inner = 0.5f * row[2 * C + c] + __fdiv_rn(row[3 * C + c] * frac, 3.0f);
  inner = row[C + c] + inner * frac;
  return row[c] + inner * frac;
}

__global__ void __launch_bounds__(1024) initial_state_kernel(const float* __restrict__ coeffs, long long row_stride, int B, int C,
                                                             int interval, float frac, const float* __restrict__ W,
                                                             const float* __restrict__ b, int H, float* __restrict__ z0) {
  extern __shared__ float sx[];                      // [kNbRows][C]
  const int row0 = blockIdx.x * kNbRows;
  for (int i = threadIdx.x; i < kNbRows * C; i += blockDim.x) {
    const int r = i / C, c = i - r * C;
    const int gr = min(row0 + r, B - 1);
    sx[i] = spline_value(coeffs + (size_t)gr * row_stride + (size_t)interval * 4 * C, C, c, frac);
  }
  __syncthreads();
  const int j = threadIdx.x;
  if (j >= H) return;
  float acc[kNbRows];
  const float bj = b[j];
#pragma unroll
  for (int r = 0; r < kNbRows; ++r) acc[r] = bj;
  const float* w = W + (size_t)j * C;
  for (int k = 0; k < C; ++k) {
    const float wk = w[k];
#pragma unroll
    for (int r = 0; r < kNbRows; ++r) acc[r] = fmaf(sx[r * C + k], wk, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kNbRows; ++r)
    if (row0 + r < B) z0[(size_t)(row0 + r) * H + j] = acc[r];
}

__global__ void __launch_bounds__(1024) readout_head_kernel(const float* __restrict__ z, long long R, int H, int pre_tanh,
                                                            const float* __restrict__ W1, const float* __restrict__ b1,
                                                            const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                                            int H1, const float* __restrict__ W2, const float* __restrict__ b2,
                                                            int O, float* __restrict__ out) {
  extern __shared__ float sm[];                      // [kNbRows][H] input, then [kNbRows][H1] hidden
  float* sz = sm;
  float* sh = sm + kNbRows * H;
  const long long row0 = (long long)blockIdx.x * kNbRows;
  for (int i = threadIdx.x; i < kNbRows * H; i += blockDim.x) {
    const int r = i / H, c = i - r * H;
    const long long gr = row0 + r < R ? row0 + r : R - 1;
    const float v = z[(size_t)gr * H + c];
    sz[i] = pre_tanh ? tanhf(v) : v;
  }
  __syncthreads();
  const int j = threadIdx.x;
  if (j < H1) {
    float acc[kNbRows];
    const float bj = b1[j];
#pragma unroll
    for (int r = 0; r < kNbRows; ++r) acc[r] = bj;
    const float* w = W1 + (size_t)j * H;
    for (int k = 0; k < H; ++k) {
      const float wk = w[k];
#pragma unroll
      for (int r = 0; r < kNbRows; ++r) acc[r] = fmaf(sz[r * H + k], wk, acc[r]);
    }
    const float sc = bn_scale ? bn_scale[j] : 1.f, sf = bn_shift ? bn_shift[j] : 0.f;
#pragma unroll
    for (int r = 0; r < kNbRows; ++r) {
      float v = bn_scale ? fmaf(acc[r], sc, sf) : acc[r];
      sh[r * H1 + j] = v < 0.f ? 0.f : v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kNbRows * O; i += blockDim.x) {
    const int r = i / O, o = i - r * O;
    if (row0 + r >= R) continue;
    float acc = b2[o];
    const float* w = W2 + (size_t)o * H1;
    for (int k = 0; k < H1; ++k) acc = fmaf(sh[r * H1 + k], w[k], acc);
    out[(size_t)(row0 + r) * O + o] = acc;
  }
}

}  // namespace snsde

using snsde::fail;

extern "C" int snsde_initial_state(const float* coeffs_dev, int64_t coeff_row_stride, int32_t B, int32_t C, int32_t n_knots,
                                   int32_t interval, float frac, const float* W_dev, const float* b_dev, int32_t H,
                                   float* z0_dev, int device, void* stream_v) {
  SNSDE_API_BEGIN
  if (!coeffs_dev || !W_dev || !b_dev || !z0_dev) return fail(SNSDE_ERR_BAD_ARG, "initial_state: coeffs/W/b/z0 is NULL");
  if (B < 1 || C < 1 || H < 1 || n_knots < 2) return fail(SNSDE_ERR_BAD_ARG, "initial_state: need B, C, H >= 1 and at least 2 knots");
  if (H > 1024) return fail(SNSDE_ERR_UNSUPPORTED, "initial_state: hidden sizes above 1024 are not supported");
  if (interval < 0 || interval > n_knots - 2) return fail(SNSDE_ERR_BAD_ARG, "initial_state: spline interval %d outside [0,%d]", interval, n_knots - 2);
  if (coeff_row_stride < (int64_t)(n_knots - 1) * 4 * C) return fail(SNSDE_ERR_BAD_ARG, "initial_state: coeff_row_stride %lld < (K-1)*4C", (long long)coeff_row_stride);
  snsde::DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.err));
  const int nt = ((H > 64 ? H : 64) + 31) & ~31;
  const int grid = (B + snsde::kNbRows - 1) / snsde::kNbRows;
  snsde::initial_state_kernel<<<grid, nt, sizeof(float) * snsde::kNbRows * C, (cudaStream_t)stream_v>>>(
      coeffs_dev, coeff_row_stride, B, C, interval, frac, W_dev, b_dev, H, z0_dev);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "initial_state launch: %s", cudaGetErrorString(e));
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

extern "C" int snsde_readout_head(const float* z_dev, int64_t R, int32_t H, int32_t pre_tanh,
                                  const float* W1_dev, const float* b1_dev, const float* bn_scale_dev, const float* bn_shift_dev,
                                  int32_t H1, const float* W2_dev, const float* b2_dev, int32_t O, float* out_dev,
                                  int device, void* stream_v) {
  SNSDE_API_BEGIN
  if (!z_dev || !W1_dev || !b1_dev || !W2_dev || !b2_dev || !out_dev) return fail(SNSDE_ERR_BAD_ARG, "readout_head: a required pointer is NULL");
  if ((bn_scale_dev == nullptr) != (bn_shift_dev == nullptr)) return fail(SNSDE_ERR_BAD_ARG, "readout_head: bn scale and shift come together");
  if (R < 1 || H < 1 || H1 < 1 || O < 1) return fail(SNSDE_ERR_BAD_ARG, "readout_head: need R, H, H1, O >= 1");
  if (H1 > 1024 || H > 4096) return fail(SNSDE_ERR_UNSUPPORTED, "readout_head: hidden sizes above 1024 are not supported");
  const size_t smem = sizeof(float) * snsde::kNbRows * ((size_t)H + H1);
  if (smem > 200 * 1024) return fail(SNSDE_ERR_UNSUPPORTED, "readout_head: rows do not fit in shared memory");
  snsde::DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.err));
  const int nt = ((H1 > 64 ? H1 : 64) + 31) & ~31;
  const long long grid = (R + snsde::kNbRows - 1) / snsde::kNbRows;
  if (grid > 0x7fffffffLL) return fail(SNSDE_ERR_BAD_ARG, "readout_head: too many rows");
  if (smem > 48 * 1024) {
    const cudaError_t ea = cudaFuncSetAttribute(snsde::readout_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ea != cudaSuccess) return fail(SNSDE_ERR_CUDA, "readout_head: %s", cudaGetErrorString(ea));
  }
  snsde::readout_head_kernel<<<(unsigned)grid, nt, smem, (cudaStream_t)stream_v>>>(
      z_dev, R, H, pre_tanh, W1_dev, b1_dev, bn_scale_dev, bn_shift_dev, H1, W2_dev, b2_dev, O, out_dev);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "readout_head launch: %s", cudaGetErrorString(e));
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}
