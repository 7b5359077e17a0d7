// Control-path coefficient construction on device: Hermite cubic with backward differences.
//
// Replaces the data-prep call torchcde.hermite_cubic_coefficients_with_backward_differences(x, t)
// (/root/reference/benchmark_classification/datasets/common.py:82-84, tests/test_neuralsde_core_alignment.py:64)
// for NaN-free inputs: x [B, K, C] -> coeffs [B, K-1, 4C] = cat(a, b, two_c, three_d), the layout the solve reads.
// One pass, HBM-bound: 4 bytes read (+ neighbours from L1/L2) and 16 bytes written per (row, interval, channel).
#include <cuda_runtime.h>
#include "../../include/snsde.h"

namespace snsde {

// blockIdx.y = batch row, thread = one (k, c) of that row with c fastest (32-bit index math): reads are coalesced
// along C (the x[k-1], x[k+1] neighbours come from L1/L2), and the 4C floats of one (row, interval) are written by
// C consecutive threads, i.e. into one contiguous 16C-byte run.
__global__ void __launch_bounds__(256) hermite_coeffs_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                             float* __restrict__ out, int B, int K, int C) {
  const int per_row = (K - 1) * C;
  const float* xb = x + (size_t)blockIdx.y * K * C;
  float* ob = out + (size_t)blockIdx.y * (K - 1) * 4 * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_row; i += gridDim.x * blockDim.x) {
    const int k = i / C, c = i - k * C;
    const float* xr = xb + i;                                            // (k, c)
    const float x0 = xr[0], x1 = xr[C];
    const float h = __fsub_rn(t[k + 1], t[k]);
    const float dx = __fsub_rn(x1, x0);
    const float m1 = __fdiv_rn(dx, h);                                   // this interval's secant slope
    float m0 = m1;                                                       // previous interval's (first: its own)
    if (k > 0) m0 = __fdiv_rn(__fsub_rn(x0, xr[-C]), __fsub_rn(t[k], t[k - 1]));
    // two_c = 2 (3 (dx/h - m0) - m1 + m0) / h ;  three_d = (1/h^2) (m1 - m0) - two_c / h     (op order of torchcde)
    const float u = __fsub_rn(__fmul_rn(3.f, __fsub_rn(m1, m0)), m1);
    const float two_c = __fdiv_rn(__fmul_rn(2.f, __fadd_rn(u, m0)), h);
    const float three_d = __fsub_rn(__fmul_rn(__fdiv_rn(1.f, __fmul_rn(h, h)), __fsub_rn(m1, m0)), __fdiv_rn(two_c, h));
    float* o = ob + (size_t)k * 4 * C + c;
    o[0] = x0;
    o[C] = m0;
    o[2 * C] = two_c;
    o[3 * C] = three_d;
  }
}

}  // namespace snsde

extern "C" int snsde_hermite_coeffs(const float* x_dev, const float* knots_dev, int32_t B, int32_t K, int32_t C,
                                    float* coeffs_dev, int device, void* stream_v) {
  if (!x_dev || !knots_dev || !coeffs_dev || B < 1 || K < 2 || C < 1) return SNSDE_ERR_BAD_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return SNSDE_ERR_CUDA;
  if ((long long)(K - 1) * C > 0x7fffffffLL || B > 65535 * 64) return SNSDE_ERR_BAD_ARG;
  const int per_row = (K - 1) * C;
  const int gx = (per_row + 255) / 256 < 64 ? (per_row + 255) / 256 : 64;
  for (int b0 = 0; b0 < B; b0 += 65535) {                                  // gridDim.y limit
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    snsde::hermite_coeffs_kernel<<<dim3(gx, nb), 256, 0, (cudaStream_t)stream_v>>>(
        x_dev + (size_t)b0 * K * C, knots_dev, coeffs_dev + (size_t)b0 * (K - 1) * 4 * C, nb, K, C);
  }
  return cudaGetLastError() == cudaSuccess ? SNSDE_OK : SNSDE_ERR_CUDA;
}
