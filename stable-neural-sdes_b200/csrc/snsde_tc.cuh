// Tensor-core (tcgen05) path: host-visible interface used by snsde_api.cu.
#pragma once
#include <cuda_runtime.h>
#include "snsde_common.cuh"

namespace snsde {

struct TcPlan {
  void* d_image = nullptr;        // packed fp16 hi/lo operand images + fp32 vectors
  size_t image_bytes = 0;
  float* d_tables = nullptr;      // per-step bias / diffusion-coefficient tables [S][...]
  size_t tables_floats = 0;
  int cfg[32] = {0};
};

struct TcForwardArgs {
  const float* coeffs; long long coeff_row_stride;
  const float* y0; int B;
  const snsde_step* steps; const snsde_step* steps_host; int S;
  const snsde_emit* emits; int n_init_emits; int n_out;
  const int* row_slot; const float* dW;
  unsigned long long seed, row_offset;
  float* out;
};

bool tc_supported(const snsde_model_desc& d, int cc_major, int smem_optin);
const char* tc_unsupported_reason();
int tc_set_weights(TcPlan& tc, const snsde_model_desc& d, const float* blob, int num_sms, int smem_optin, cudaStream_t stream);
cudaError_t tc_forward(TcPlan& tc, const TcForwardArgs& a, cudaStream_t stream, int* n_launches);
void tc_release(TcPlan& tc);

}  // namespace snsde
