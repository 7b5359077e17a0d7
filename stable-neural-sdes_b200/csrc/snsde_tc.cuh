// Tensor-core (tcgen05) path: host-visible interface used by snsde_api.cu.
#pragma once
#include <cuda.h>             // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>
#include "snsde_common.cuh"

namespace snsde {

constexpr int kTcMaxLayers = 6;

struct TcLayer {
  int a_hi, a_lo;   // byte offsets of the fp16 hi / scaled-lo operand images inside the (global) weight image
  int K;            // contraction length (multiple of 16)
  int bias;         // float offset into the vector region
  // per-launch placement of the two operand images: a TMEM column offset (TS-form MMA, bit set in `ts`) or a
  // byte offset into the shared-memory weight area (SS-form MMA)
  int h_hi, h_lo, ts;   // ts bit 0: hi image in TMEM, bit 1: lo image in TMEM
};

// One operand image (hi or lo part of one weight matrix) and where this launch keeps it.
struct TcImg {
  int g_off, bytes;     // location inside the global weight image
  int tmem_col;         // >= 0: resident in TMEM from this column on (K/2 columns); -1: shared memory
  int s_off;            // byte offset inside the shared-memory weight area (when tmem_col < 0)
};
constexpr int kTcMaxImgs = 2 * (kTcMaxLayers + 1);

// Everything the tcgen05 kernel needs; built by tc_set_weights (model part) and tc_forward (call part).
struct TcParams {
  int H, C, Cpad, NL, uses_control;
  TcLayer layer[kTcMaxLayers];
  int ax_hi, ax_lo;          // control segment of layer 0 (K = Cpad), byte offsets in the global image
  int hx_hi, hx_lo, x_ts;    // its per-launch placement (see TcLayer)
  TcImg img[kTcMaxImgs]; int n_img;
  int tmem_cols;             // TMEM columns to allocate (power of two)
  int c_sin, c_cos;          // float offsets of the folded time-feature vectors of layer 0 (-1: none)
  int coef_vec;              // float offset of a per-feature diffusion coefficient vector (-1: none)
  TailOp tail;
  const uint8_t* wimg; int wimg_bytes;   // global weight image
  int w_smem_bytes;          // bytes of the shared-memory weight area (SS-form images of this launch)
  const float* vec;
  const float* a_tab;        // [S][H] per-step diffusion coefficient (noise_t networks), or null
  // per call
  const float* coeffs; long long coeff_row_stride;
  const float* y0; int B;
  const snsde_step* steps; int S;
  const snsde_emit* emits; int n_init_emits; int n_out;
  const int* row_slot; const float* dW;
  unsigned long long seed, row_offset;
  float* out;
  int nx, nstg;              // control-operand ring depth, coefficient staging depth
  long long* dbg;            // optional: clock64 trace of CTA 0 (SNSDE_TC_TRACE env), [step][event]
  int* status;               // sticky flags (bit 0: operand beyond the fp16 range was saturated)
  // 2-D tensor map of the coefficients [B][(K-1)*4C] with box [NR][4C]: one TMA tensor copy per step fetches the
  // spline rows of all NR rows of the CTA (use_tmap = 0: per-row 1-D bulk copies, e.g. 4C > 256)
  int use_tmap;
  CUtensorMap tmap;
};

// Per-step table of the row-independent noise networks (options 12,13,16,17).
struct TcNoiseNet {
  int kind;                  // 0 none, 1 = Linear(2,H), 2 = relu(Linear(H,H)(relu(Linear(2,H))))
  int w1t, b1, w2t, b2;      // float offsets into the vector region (w1t [2][H], w2t [H][H] transposed)
};

struct TcPlan {
  uint8_t* d_wimg = nullptr; int wimg_bytes = 0;
  float* d_vec = nullptr; int vec_floats = 0;
  float* d_atab = nullptr; int atab_cap = 0;
  // a_tab depends on the step times and the noise-net weights only: it is reused while both are unchanged
  // (every batch of a dataset shares its knots).  Key = FNV-1a of the steps' (sin t0, cos t0) + S.
  unsigned long long atab_key = 0; bool atab_valid = false;
  TcParams proto;            // model part of the params
  TcNoiseNet noise;
  int num_sms = 0, smem_optin = 0;
  bool ready = false;
};

struct TcForwardArgs {
  const float* coeffs; long long coeff_row_stride; int n_knots;
  const float* y0; int B;
  const snsde_step* steps; const snsde_step* steps_host; int S;
  const snsde_emit* emits; int n_init_emits; int n_out;
  const int* row_slot; const float* dW;
  unsigned long long seed, row_offset;
  float* out;
  int* status;
};

bool tc_supported(const snsde_model_desc& d, int cc_major, int smem_optin);
const char* tc_unsupported_reason();
int tc_set_weights(TcPlan& tc, const snsde_model_desc& d, const Program& pg, const float* blob, int num_sms,
                   int smem_optin, cudaStream_t stream);
cudaError_t tc_forward(TcPlan& tc, const TcForwardArgs& a, cudaStream_t stream, int* n_launches);
void tc_release(TcPlan& tc);

}  // namespace snsde
