// General tcgen05 path: host-visible interface + kernel parameter block.
//
// Same algorithm, orientation and precision scheme as snsde_tc.cu, without its capacity limits:
//   * weight operand segments are RESIDENT in tensor memory (TS-form MMAs, while the columns behind the
//     accumulators last), RESIDENT in shared memory, or STREAMED from L2 every step through a
//     ring of 8 KB slots filled by 1-D bulk async copies (the weight sequence of a step is static, so the
//     producer runs ahead across layers and steps);
//   * the output features span up to two 128-row M tiles (hidden <= 256: BASELINE config c5);
//   * a phase may carry a second network beside the drift: the state-dependent noise networks of
//     noise options 14,15,18,19 (neuralsde.py:172-179, 275-286; BASELINE config c4 is (3,18)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "snsde_common.cuh"
#include "snsde_tc.cuh"

namespace snsde {

constexpr int kTcgMaxJobs = 32;
constexpr int kTcgMaxPhases = 6;
constexpr int kTcgSlotBytes = 8192;        // one 16-wide K chunk of one 128-row tile: hi image (4 KB) + lo image (4 KB)

// One MMA group: accumulator set `acc` of phase `phase` (+)= A(job) x B(b_src).
struct TcgJob {
  int phase;        // 0..NP-1
  int acc;          // accumulator set inside the phase's TMEM region: net * MT + mtile
  int nk;           // 16-wide K chunks (multiple of the chain count)
  int b_src;        // 0: B0 (state y, then drift activations)  1: B1 (noise-net activations)  2: X(t) ring
  int b_chunk0;     // first K chunk inside the B operand
  int fresh;        // first group into this accumulator set within the phase
  int stream;       // A tiles come through the ring
  int a_off;        // resident: byte offset in the shared-memory weight area (packed per launch)
  int g_off;        // byte offset of the job's tiles in the global weight blob ([chunk][hi 4 KB | lo 4 KB])
  int g_off1;       // M-split launches: the tiles of the cluster's second CTA (output features 128..255)
  int tmem_col;     // >= 0: the job's tiles are resident in TMEM (TS-form MMAs): hi images at [col, col + 8 nk),
                    // lo images at [col + 8 nk, col + 16 nk); -1: shared memory / ring
};

struct TcgParams {
  int H, HP, MT, C, Cpad, NP, NN, uses_control, nets;   // HP = 128*MT; NN = noise-net layers (0,1,2)
  int n_jobs, n_xjobs;                       // jobs[0:n_jobs) per step in issue order; the last n_xjobs are the X(t) groups
  TcgJob jobs[kTcgMaxJobs];
  int bias[kTcgMaxPhases][2];                // float offsets into vec: [phase][net] (-1: none), HP floats each
  int c_sin[2], c_cos[2];                    // time-feature vectors of layer 0 per net (-1: none)
  int coef_vec;
  int noise_act[2];                          // activation of the noise-net layers (ACT_*)
  TailOp tail;
  const uint8_t* wblob;                      // every job's tiles (global): copied once (resident) or read every step (streamed)
  int wres_bytes;                            // bytes of the packed resident area in shared memory
  int n_stream_chunks;                       // slots consumed per step
  const float* vec;
  const float* a_tab;
  // per call
  const float* coeffs; long long coeff_row_stride;
  const float* y0; int B;
  const snsde_step* steps; int S;
  const snsde_emit* emits; int n_init_emits; int n_out;
  const int* row_slot; const float* dW;
  unsigned long long seed, row_offset;
  float* out;
  int nx, nstg, nslot;
  long long* dbg;
  int* status;               // sticky flags (bit 0: operand beyond the fp16 range was saturated)
};

struct TcgPlan {
  uint8_t* d_wblob = nullptr; int wblob_cap = 0;
  float* d_vec = nullptr; int vec_cap = 0;
  float* d_atab = nullptr; int atab_cap = 0;
  TcgParams proto;
  TcNoiseNet noise;
  int num_sms = 0, smem_optin = 0;
  bool ready = false;
};

bool tcg_supported(const snsde_model_desc& d, int cc_major, int smem_optin);
const char* tcg_unsupported_reason();
int tcg_set_weights(TcgPlan& tc, const snsde_model_desc& d, const Program& pg, const float* blob, int num_sms,
                    int smem_optin, cudaStream_t stream);
cudaError_t tcg_forward(TcgPlan& tc, const TcForwardArgs& a, cudaStream_t stream, int* n_launches);
void tcg_release(TcgPlan& tc);

// launchers instantiated in snsde_tcg_inst*.cu
template <int NR, int CH, int MT, int DIFF, bool MS = false>
cudaError_t tcg_launch(const TcgParams& p, int grid, size_t smem, cudaStream_t stream);

}  // namespace snsde
