// C-ABI entry points (include/snsde.h): plan management, model -> kernel program compilation,
// weight re-layout, launch.  Host side only; kernels live in snsde_fma.cu / snsde_tc.cu.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "snsde_common.cuh"
#include "snsde_rng.cuh"
#include "snsde_tc.cuh"
#include "snsde_tcg.cuh"

namespace snsde {
size_t fma_smem_bytes(const Program& pg, int R, int smem_w_floats);
cudaError_t fma_launch(const FmaParams& p, int R, int nt, size_t smem, cudaStream_t stream);
}  // namespace snsde

using namespace snsde;

// ---- error plumbing ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) return fail(SNSDE_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

// ---- plan -------------------------------------------------------------------------------------
struct snsde_plan {
  snsde_model_desc desc;
  int device = 0, num_sms = 0, smem_optin = 0;
  int kind = 0;                       // 0 FMA, 1 tcgen05 (weights resident), 2 tcgen05 (general: streamed weights, 2 M tiles, noise nets)
  bool has_weights = false;
  Program prog;
  float* d_wimg = nullptr;
  int wimg_floats = 0;
  TcPlan tc;                          // tensor-core path state (kind == 1)
  TcgPlan tcg;                        // general tensor-core path state (kind == 2)
  snsde_step* d_steps = nullptr; int steps_cap = 0; std::vector<snsde_step> h_steps;
  snsde_emit* d_emits = nullptr; int emits_cap = 0; std::vector<snsde_emit> h_emits;
  int64_t launches = 0;
  int* d_status = nullptr;            // sticky device flags (snsde_plan_status)
  // The plan owns device tables (steps, emits, noise tables) that every solve reads: consecutive solves of one
  // plan are serialised across streams by this event (the caller's copies on other streams still overlap).
  cudaEvent_t done_ev = nullptr; void* last_stream = nullptr; bool ev_pending = false;
};

static bool is_time_opt(int io) { return io >= 3 && io <= 6; }
static bool is_emb_opt(int io) { return io == 2 || io == 4 || io == 6; }
static bool uses_control(int io) { return io == 0 || is_emb_opt(io); }

static int validate(const snsde_model_desc* d) {
  if (!d) return fail(SNSDE_ERR_BAD_ARG, "desc is NULL");
  if (d->family != SNSDE_FAMILY_BENCHMARK && d->family != SNSDE_FAMILY_TUTORIAL_LSDE)
    return fail(SNSDE_ERR_BAD_ARG, "unknown family %d", d->family);
  if (d->input_channels < 1 || d->hidden < 1 || d->hidden_hidden < 1 || d->num_hidden_layers < 1)
    return fail(SNSDE_ERR_BAD_ARG, "C/H/HH/L must be >= 1");
  if (std::max(d->hidden, d->hidden_hidden) > 1024)
    return fail(SNSDE_ERR_UNSUPPORTED, "hidden sizes above 1024 are not supported");
  if (d->method != SNSDE_METHOD_EULER && d->method != SNSDE_METHOD_MILSTEIN)
    return fail(SNSDE_ERR_UNSUPPORTED, "method %d: only euler(0) and milstein(1) are implemented ('srk' is SURVEY 8f2)", d->method);
  if (d->precision < 0 || d->precision > SNSDE_PRECISION_AUTO) return fail(SNSDE_ERR_BAD_ARG, "bad precision %d", d->precision);
  if (d->family == SNSDE_FAMILY_BENCHMARK) {
    if (d->input_option < 0 || d->input_option > 6) return fail(SNSDE_ERR_BAD_ARG, "input_option %d not in 0..6", d->input_option);
    if (d->noise_option < 0 || d->noise_option > 19)
      return fail(SNSDE_ERR_BAD_ARG, "Unknown noise_option %d.", d->noise_option);   // neuralsde.py:288
    if ((d->input_option == 0 || is_emb_opt(d->input_option)) && d->hidden != d->hidden_hidden)
      return fail(SNSDE_ERR_BAD_ARG, "input_option %d requires hidden_hidden == hidden (emb is Linear(2H,H), neuralsde.py:154,210)", d->input_option);
  }
  const int n_ops = d->family == SNSDE_FAMILY_BENCHMARK ? 6 + d->num_hidden_layers : 6 + 2 * d->num_hidden_layers;
  if (n_ops > kMaxOps) return fail(SNSDE_ERR_UNSUPPORTED, "num_hidden_layers too large");
  return SNSDE_OK;
}

static int64_t weight_count(const snsde_model_desc* d) {
  const int64_t C = d->input_channels, H = d->hidden, HH = d->hidden_hidden, L = d->num_hidden_layers;
  int64_t n = 0;
  if (d->family == SNSDE_FAMILY_BENCHMARK) {
    const int io = d->input_option, no = d->noise_option;
    n += H * C + H;
    n += HH * (H + (is_time_opt(io) ? 2 : 0)) + HH;
    if (is_emb_opt(io)) n += H * 2 * H + H;
    n += (L - 1) * (HH * HH + HH);
    n += H * HH + H;
    n += 1;
    if (no >= 1 && no <= 3) n += 1;
    if (no >= 4 && no <= 6) n += H;
    if (no == 12 || no == 13) n += H * 2 + H;
    if (no == 14 || no == 15) n += H * (H + 2) + H;
    if (no == 16 || no == 17) n += H * 2 + H + H * H + H;
    if (no == 18 || no == 19) n += H * (H + 2) + H + H * H + H;
  } else {
    const int64_t mlp = (HH * H + HH) + (L - 1) * (HH * HH + HH) + (H * HH + H);
    n += H * C + H;            // linear_X
    n += H * 2 * H + H;        // emb
    n += mlp;                  // f_net
    n += H * H + H;            // linear_out
    n += H * 1 + H;            // noise_in
    n += mlp;                  // g_net
  }
  return n;
}

// ---- weight image for the FMA kernel ------------------------------------------------------------
struct ImageBuilder {
  std::vector<float> img;
  void pad4() { while (img.size() & 3) img.push_back(0.f); }
  // W is [N][ldw] row-major (nn.Linear); appends Wt[k][j] = W[j][col0 + k], k < ncols.
  int add_T(const float* W, int N, int ldw, int col0, int ncols) {
    pad4();
    const int off = (int)img.size();
    for (int k = 0; k < ncols; ++k)
      for (int j = 0; j < N; ++j) img.push_back(W[(size_t)j * ldw + col0 + k]);
    return off;
  }
  int add_vec(const float* v, int n) {
    pad4();
    const int off = (int)img.size();
    img.insert(img.end(), v, v + n);
    return off;
  }
};

static DenseOp make_op(int dst, int src, int K, int N, int w, int b, int act) {
  DenseOp o;
  memset(&o, 0, sizeof(o));
  o.dst = dst; o.src = src; o.src2 = BUF_NONE; o.K = K; o.N = N;
  o.w_off = w; o.w2_off = -1; o.b_off = b; o.tw_off = -1; o.tmode = TM_NONE; o.act = act;
  return o;
}

struct BlobCursor {
  const float* q;
  const float* take(size_t n) { const float* r = q; q += n; return r; }
};

static void compile_benchmark(const snsde_model_desc& d, const float* blob, Program& pg, ImageBuilder& ib) {
  const int C = d.input_channels, H = d.hidden, HH = d.hidden_hidden, L = d.num_hidden_layers;
  const int io = d.input_option, no = d.noise_option;
  const int tau = is_time_opt(io) ? 2 : 0;
  BlobCursor bc{blob};
  const float* Wi = bc.take((size_t)H * C); const float* bi = bc.take(H);
  const float* Win = bc.take((size_t)HH * (H + tau)); const float* bin = bc.take(HH);
  const float *We = nullptr, *be = nullptr;
  if (is_emb_opt(io)) { We = bc.take((size_t)H * 2 * H); be = bc.take(H); }
  std::vector<const float*> Wl(L - 1), bl(L - 1);
  for (int l = 0; l < L - 1; ++l) { Wl[l] = bc.take((size_t)HH * HH); bl[l] = bc.take(HH); }
  const float* Wo = bc.take((size_t)H * HH); const float* bo = bc.take(H);
  const float theta = *bc.take(1);

  memset(&pg, 0, sizeof(pg));
  pg.C = C; pg.H = H; pg.HH = HH;
  pg.ld = (std::max(std::max(H, HH), C) + 3) & ~3;
  pg.uses_control = uses_control(io);
  int n = 0;
  int cur;
  if (pg.uses_control) {                      // Xt = initial_network(X(t))      neuralsde.py:296-297
    pg.ops[n++] = make_op(BUF_U, BUF_X, C, H, ib.add_T(Wi, H, C, 0, C), ib.add_vec(bi, H),
                          io == 0 ? ACT_RELU : ACT_NONE);
  }
  if (io == 0) {
    cur = BUF_U;                              // z = Xt                         :206-207
  } else {                                    // yy = linear_in([tf,] y)        :200-204
    DenseOp o = make_op(BUF_A, BUF_Y, H, HH, ib.add_T(Win, HH, H + tau, tau, H), ib.add_vec(bin, HH),
                        is_emb_opt(io) ? ACT_NONE : ACT_RELU);
    if (tau) { o.tmode = TM_SINCOS; o.tw_off = ib.add_T(Win, HH, H + tau, 0, 2); }
    pg.ops[n++] = o;
    cur = BUF_A;
    if (is_emb_opt(io)) {                     // z = emb(cat(yy, Xt))           :210
      DenseOp e = make_op(BUF_B, BUF_A, H, H, ib.add_T(We, H, 2 * H, 0, H), ib.add_vec(be, H), ACT_RELU);
      e.src2 = BUF_U; e.K2 = H; e.w2_off = ib.add_T(We, H, 2 * H, H, H);
      pg.ops[n++] = e;
      cur = BUF_B;
    }
  }
  // noise networks (row-independent ones run on ONE row per CTA)              :170-179, 271-286
  TailOp& t = pg.tail;
  t.geometric = (io == 5 || io == 6);
  t.clip_drift = 1;
  t.bounded = 1;
  t.s_theta = 1.f / (1.f + expf(-theta));
  t.milstein = d.method == SNSDE_METHOD_MILSTEIN;
  t.special = SP_NONE; t.mult = MU_ONE; t.coef_src = CO_NONE; t.coef_scalar = 0.f;
  if (no == 0) t.special = SP_ZERO;
  if (no >= 1 && no <= 3) {
    t.coef_src = CO_SCALAR; t.coef_scalar = expf(*bc.take(1));
    t.mult = no == 1 ? MU_ONE : (no == 2 ? MU_T : MU_Y);
  }
  if (no >= 4 && no <= 6) {
    const float* sd = bc.take(H);
    std::vector<float> e(H);
    for (int j = 0; j < H; ++j) e[j] = expf(sd[j]);
    t.coef_src = CO_IMG; t.coef_ref = ib.add_vec(e.data(), H);
    t.mult = no == 4 ? MU_ONE : (no == 5 ? MU_T : MU_Y);
  }
  if (no == 7) t.special = SP_SQRT;
  if (no == 8) t.special = SP_CUBE;
  if (no == 9) t.special = SP_SIGMOID;
  if (no == 10) t.special = SP_RELU;
  if (no == 11) t.mult = MU_TY;
  if (no == 12 || no == 13 || no == 16 || no == 17) {
    const float* W1 = bc.take((size_t)H * 2); const float* b1 = bc.take(H);
    DenseOp o = make_op(BUF_V0, BUF_NONE, 0, H, -1, ib.add_vec(b1, H), no >= 16 ? ACT_RELU : ACT_NONE);
    o.vec = 1; o.tmode = TM_SINCOS; o.tw_off = ib.add_T(W1, H, 2, 0, 2);
    pg.ops[n++] = o;
    t.coef_src = CO_VBUF; t.coef_ref = BUF_V0;
    if (no >= 16) {
      const float* W2 = bc.take((size_t)H * H); const float* b2 = bc.take(H);
      DenseOp o2 = make_op(BUF_V1, BUF_V0, H, H, ib.add_T(W2, H, H, 0, H), ib.add_vec(b2, H), ACT_RELU);
      o2.vec = 1;
      pg.ops[n++] = o2;
      t.coef_ref = BUF_V1;
    }
    t.mult = (no == 13 || no == 17) ? MU_Y : MU_ONE;
  }
  if (no == 14 || no == 15 || no == 18 || no == 19) {
    const float* W1 = bc.take((size_t)H * (H + 2)); const float* b1 = bc.take(H);
    const bool deep = no >= 18;
    DenseOp o = make_op(deep ? BUF_P : BUF_Q, BUF_Y, H, H, ib.add_T(W1, H, H + 2, 2, H), ib.add_vec(b1, H),
                        deep ? ACT_RELU : ACT_NONE);
    o.tmode = TM_SINCOS; o.tw_off = ib.add_T(W1, H, H + 2, 0, 2);
    pg.ops[n++] = o;
    t.vjp_kind = deep ? 2 : 1; t.vjp_w1 = o.w_off; t.vjp_w2 = -1; t.vjp_h1 = BUF_P;
    if (deep) {
      const float* W2 = bc.take((size_t)H * H); const float* b2 = bc.take(H);
      pg.ops[n++] = make_op(BUF_Q, BUF_P, H, H, ib.add_T(W2, H, H, 0, H), ib.add_vec(b2, H), ACT_RELU);
      t.vjp_w2 = pg.ops[n - 1].w_off;
    }
    t.coef_src = CO_RBUF; t.coef_ref = BUF_Q;
    t.mult = (no == 15 || no == 19) ? MU_Y : MU_ONE;
  }
  // shared MLP tail: relu -> (Linear, relu)* -> linear_out                     :212-217
  for (int l = 0; l < L - 1; ++l) {
    const int dst = (cur == BUF_A) ? BUF_B : BUF_A;
    pg.ops[n++] = make_op(dst, cur, HH, HH, ib.add_T(Wl[l], HH, HH, 0, HH), ib.add_vec(bl[l], HH), ACT_RELU);
    cur = dst;
  }
  DenseOp fo = make_op(BUF_NONE, cur, HH, H, ib.add_T(Wo, H, HH, 0, HH), ib.add_vec(bo, H), ACT_NONE);
  fo.final_drift = 1;
  pg.ops[n++] = fo;
  pg.n_ops = n;
}

static void compile_tutorial(const snsde_model_desc& d, const float* blob, Program& pg, ImageBuilder& ib) {
  const int C = d.input_channels, H = d.hidden, HH = d.hidden_hidden, L = d.num_hidden_layers;
  BlobCursor bc{blob};
  memset(&pg, 0, sizeof(pg));
  pg.C = C; pg.H = H; pg.HH = HH;
  pg.ld = (std::max(std::max(H, HH), C) + 3) & ~3;
  pg.uses_control = 1;
  int n = 0;
  const float* WX = bc.take((size_t)H * C); const float* bX = bc.take(H);
  const float* We = bc.take((size_t)H * 2 * H); const float* be = bc.take(H);
  pg.ops[n++] = make_op(BUF_U, BUF_X, C, H, ib.add_T(WX, H, C, 0, C), ib.add_vec(bX, H), ACT_NONE);
  DenseOp e = make_op(BUF_A, BUF_Y, H, H, ib.add_T(We, H, 2 * H, 0, H), ib.add_vec(be, H), ACT_NONE);
  e.src2 = BUF_U; e.K2 = H; e.w2_off = ib.add_T(We, H, 2 * H, H, H);     // emb(cat(y, Xt))
  pg.ops[n++] = e;
  auto mlp = [&](int cur, int b0, int b1, bool vec) {
    int in = H;
    for (int l = 0; l < L; ++l) {            // Linear(in->HH) + LipSwish, L times
      const float* W = bc.take((size_t)HH * in); const float* b = bc.take(HH);
      const int dst = (cur == b0) ? b1 : b0;
      DenseOp o = make_op(dst, cur, in, HH, ib.add_T(W, HH, in, 0, in), ib.add_vec(b, HH), ACT_LIPSWISH);
      o.vec = vec;
      pg.ops[n++] = o;
      cur = dst; in = HH;
    }
    const float* W = bc.take((size_t)H * HH); const float* b = bc.take(H);
    const int dst = (cur == b0) ? b1 : b0;
    DenseOp o = make_op(dst, cur, HH, H, ib.add_T(W, H, HH, 0, HH), ib.add_vec(b, H), ACT_NONE);
    o.vec = vec;
    pg.ops[n++] = o;
    return dst;
  };
  const int fcur = mlp(BUF_A, BUF_A, BUF_B, false);
  const float* Wlo = bc.take((size_t)H * H); const float* blo = bc.take(H);
  const float* Wni = bc.take((size_t)H); const float* bni = bc.take(H);
  DenseOp ni = make_op(BUF_V0, BUF_NONE, 0, H, -1, ib.add_vec(bni, H), ACT_NONE);
  ni.vec = 1; ni.tmode = TM_RAW; ni.tw_off = ib.add_T(Wni, H, 1, 0, 1);
  pg.ops[n++] = ni;
  const int gcur = mlp(BUF_V0, BUF_V0, BUF_V1, true);
  DenseOp fo = make_op(BUF_NONE, fcur, H, H, ib.add_T(Wlo, H, H, 0, H), ib.add_vec(blo, H), ACT_NONE);
  fo.final_drift = 1;
  pg.ops[n++] = fo;
  pg.n_ops = n;
  TailOp& t = pg.tail;
  memset(&t, 0, sizeof(t));
  t.coef_src = CO_VBUF; t.coef_ref = gcur; t.mult = MU_ONE; t.special = SP_NONE;
  t.bounded = 0; t.clip_drift = 0; t.geometric = 0; t.s_theta = 1.f;
  t.milstein = d.method == SNSDE_METHOD_MILSTEIN;
}

// ---- Philox materialisation kernel -------------------------------------------------------------
__global__ void philox_fill_kernel(unsigned long long seed, unsigned long long row_offset, int S, int B, int H,
                                   const float* __restrict__ sqrt_h, float* __restrict__ dW) {
  const size_t n = (size_t)S * B * H;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % H);
    const size_t sb = i / H;
    const int b = (int)(sb % B), s = (int)(sb / B);
    const unsigned long long gb = row_offset + (unsigned long long)b;
    float nrm[4];
    philox_normals4(seed, (uint32_t)j, (uint32_t)(gb >> 2), (uint32_t)s, nrm);
    dW[i] = __fmul_rn(pick4(nrm, (int)(gb & 3ull)), sqrt_h[s]);
  }
}

// ---- ABI ----------------------------------------------------------------------------------------
extern "C" {

int snsde_abi_version(void) { return SNSDE_ABI_VERSION; }
const char* snsde_last_error(void) { return g_err; }

int64_t snsde_weight_count(const snsde_model_desc* desc) {
  const int rc = validate(desc);
  if (rc != SNSDE_OK) return rc;
  return weight_count(desc);
}

int snsde_plan_create(const snsde_model_desc* desc, int device, snsde_plan** out_plan) {
  if (!out_plan) return fail(SNSDE_ERR_BAD_ARG, "out_plan is NULL");
  *out_plan = nullptr;
  const int rc = validate(desc);
  if (rc != SNSDE_OK) return rc;
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(SNSDE_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
  snsde_plan* p = new snsde_plan();
  p->desc = *desc;
  p->device = device;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { delete p; return fail(SNSDE_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
  p->num_sms = prop.multiProcessorCount;
  p->smem_optin = (int)prop.sharedMemPerBlockOptin;
  const bool force_general = getenv("SNSDE_FORCE_TCG") != nullptr;           // testing aid: general kernel even where the resident one applies
  const bool tc_ok = tc_supported(*desc, prop.major, p->smem_optin) && !force_general;
  const int no_ = desc->noise_option;
  // Milstein through a state-dependent noise network needs the full vjp: implemented in the FMA kernel only
  const bool net_vjp = desc->family == SNSDE_FAMILY_BENCHMARK && desc->method == SNSDE_METHOD_MILSTEIN &&
                       (no_ == 14 || no_ == 15 || no_ == 18 || no_ == 19);
  const bool tcg_ok = tcg_supported(*desc, prop.major, p->smem_optin) && !net_vjp;
  if (desc->precision == SNSDE_PRECISION_TC && !tc_ok && !tcg_ok) {
    delete p;
    return fail(SNSDE_ERR_UNSUPPORTED, "tensor-core path does not support this model/shape/device: %s", tcg_unsupported_reason());
  }
  p->kind = desc->precision == SNSDE_PRECISION_FP32 ? 0 : (tc_ok ? 1 : (tcg_ok ? 2 : 0));
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p->d_status, sizeof(int)) != cudaSuccess ||
      cudaMemset(p->d_status, 0, sizeof(int)) != cudaSuccess) {
    delete p;
    return fail(SNSDE_ERR_CUDA, "cannot allocate the plan status word");
  }
  if (cudaEventCreateWithFlags(&p->done_ev, cudaEventDisableTiming) != cudaSuccess) p->done_ev = nullptr;
  *out_plan = p;
  return SNSDE_OK;
}

int snsde_plan_destroy(snsde_plan* p) {
  if (!p) return SNSDE_OK;
  cudaSetDevice(p->device);
  cudaFree(p->d_wimg);
  cudaFree(p->d_steps);
  cudaFree(p->d_emits);
  cudaFree(p->d_status);
  if (p->done_ev) cudaEventDestroy(p->done_ev);
  tc_release(p->tc);
  tcg_release(p->tcg);
  delete p;
  return SNSDE_OK;
}

int snsde_plan_kernel_kind(const snsde_plan* p) {
  if (!p) return fail(SNSDE_ERR_BAD_ARG, "plan is NULL");
  return p->kind;
}

int64_t snsde_plan_launch_count(const snsde_plan* p) { return p ? p->launches : 0; }

int snsde_plan_status(snsde_plan* p, void* stream_v) {
  if (!p) return fail(SNSDE_ERR_BAD_ARG, "plan is NULL");
  cudaStream_t stream = (cudaStream_t)stream_v;
  int flags = 0;
  CUDA_TRY(cudaSetDevice(p->device));
  CUDA_TRY(cudaMemcpyAsync(&flags, p->d_status, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaMemsetAsync(p->d_status, 0, sizeof(int), stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  return flags;
}

int snsde_plan_set_weights(snsde_plan* p, const float* blob, int64_t n_floats, int on_device, void* stream_v) {
  if (!p || !blob) return fail(SNSDE_ERR_BAD_ARG, "plan/blob is NULL");
  const int64_t want = weight_count(&p->desc);
  if (n_floats != want) return fail(SNSDE_ERR_BAD_ARG, "weight blob has %lld floats, model needs %lld", (long long)n_floats, (long long)want);
  cudaStream_t stream = (cudaStream_t)stream_v;
  CUDA_TRY(cudaSetDevice(p->device));
  std::vector<float> host;
  if (on_device) {
    host.resize(n_floats);
    CUDA_TRY(cudaMemcpyAsync(host.data(), blob, n_floats * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    blob = host.data();
  }
  ImageBuilder ib;
  if (p->desc.family == SNSDE_FAMILY_BENCHMARK) compile_benchmark(p->desc, blob, p->prog, ib);
  else compile_tutorial(p->desc, blob, p->prog, ib);
  ib.pad4();
  if ((int)ib.img.size() > p->wimg_floats) {
    cudaFree(p->d_wimg);
    p->d_wimg = nullptr;
    CUDA_TRY(cudaMalloc(&p->d_wimg, ib.img.size() * sizeof(float)));
  }
  p->wimg_floats = (int)ib.img.size();
  // pageable source: the runtime stages it before returning, so `ib` may die at scope exit
  CUDA_TRY(cudaMemcpyAsync(p->d_wimg, ib.img.data(), ib.img.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
  if (p->kind == 1) {
    const int rc = tc_set_weights(p->tc, p->desc, p->prog, blob, p->num_sms, p->smem_optin, stream);
    if (rc == SNSDE_ERR_UNSUPPORTED && p->desc.precision == SNSDE_PRECISION_AUTO) p->kind = 0;   // e.g. weights beyond fp16 range
    else if (rc != SNSDE_OK) return fail(rc, "tensor-core weight packing failed: %s", tc_unsupported_reason());
  }
  if (p->kind == 2) {
    const int rc = tcg_set_weights(p->tcg, p->desc, p->prog, blob, p->num_sms, p->smem_optin, stream);
    if (rc == SNSDE_ERR_UNSUPPORTED && p->desc.precision == SNSDE_PRECISION_AUTO) p->kind = 0;
    else if (rc != SNSDE_OK) return fail(rc, "tensor-core weight packing failed: %s", tcg_unsupported_reason());
  }
  p->has_weights = true;
  return SNSDE_OK;
}

static int upload_tables(snsde_plan* p, const snsde_step* steps, int S, const snsde_emit* emits, int E,
                         cudaStream_t stream) {
  if (S > p->steps_cap) {
    cudaFree(p->d_steps); p->d_steps = nullptr; p->h_steps.clear();
    CUDA_TRY(cudaMalloc(&p->d_steps, sizeof(snsde_step) * (size_t)std::max(S, 64)));
    p->steps_cap = std::max(S, 64);
  }
  if (E > p->emits_cap) {
    cudaFree(p->d_emits); p->d_emits = nullptr; p->h_emits.clear();
    CUDA_TRY(cudaMalloc(&p->d_emits, sizeof(snsde_emit) * (size_t)std::max(E, 64)));
    p->emits_cap = std::max(E, 64);
  }
  if ((int)p->h_steps.size() != S || (S && memcmp(p->h_steps.data(), steps, sizeof(snsde_step) * S) != 0)) {
    p->h_steps.assign(steps, steps + S);
    if (S) CUDA_TRY(cudaMemcpyAsync(p->d_steps, steps, sizeof(snsde_step) * S, cudaMemcpyHostToDevice, stream));
  }
  if ((int)p->h_emits.size() != E || (E && memcmp(p->h_emits.data(), emits, sizeof(snsde_emit) * E) != 0)) {
    p->h_emits.assign(emits, emits + E);
    if (E) CUDA_TRY(cudaMemcpyAsync(p->d_emits, emits, sizeof(snsde_emit) * E, cudaMemcpyHostToDevice, stream));
  }
  return SNSDE_OK;
}

int snsde_forward(snsde_plan* p, const float* coeffs_dev, int64_t coeff_row_stride, int32_t n_knots,
                  const float* y0_dev, int32_t B, const snsde_step* steps_host, int32_t S,
                  const snsde_emit* emits_host, int32_t E, int32_t n_init_emits, int32_t n_out,
                  const int32_t* row_slot_dev, const float* dW_dev, uint64_t seed, uint64_t row_offset,
                  float* out_dev, void* stream_v) {
  if (!p) return fail(SNSDE_ERR_BAD_ARG, "plan is NULL");
  if (!p->has_weights) return fail(SNSDE_ERR_NO_WEIGHTS, "snsde_forward before snsde_plan_set_weights");
  if (!y0_dev || !out_dev) return fail(SNSDE_ERR_BAD_ARG, "y0/out is NULL");
  if (B < 1 || S < 0 || E < 0 || n_out < 1) return fail(SNSDE_ERR_BAD_ARG, "bad sizes B=%d S=%d E=%d n_out=%d", B, S, E, n_out);
  if ((S && !steps_host) || (E && !emits_host)) return fail(SNSDE_ERR_BAD_ARG, "step/emit table is NULL");
  if (n_init_emits < 0 || n_init_emits > E) return fail(SNSDE_ERR_BAD_ARG, "n_init_emits out of range");
  const Program& pg = p->prog;
  if (pg.uses_control) {
    if (!coeffs_dev) return fail(SNSDE_ERR_BAD_ARG, "model reads the control path but coeffs is NULL");
    if (n_knots < 2) return fail(SNSDE_ERR_BAD_ARG, "need at least 2 knots");
    if (coeff_row_stride < (int64_t)(n_knots - 1) * 4 * pg.C)
      return fail(SNSDE_ERR_BAD_ARG, "coeff_row_stride %lld < (K-1)*4C", (long long)coeff_row_stride);
    if (((uintptr_t)coeffs_dev & 15) || (coeff_row_stride & 3))
      return fail(SNSDE_ERR_BAD_ARG, "coeffs must be 16-byte aligned with a row stride multiple of 4 floats");
  }
  int prev_end = n_init_emits;
  for (int s = 0; s < S; ++s) {
    const snsde_step& st = steps_host[s];
    if (pg.uses_control && (st.interval < 0 || st.interval > n_knots - 2))
      return fail(SNSDE_ERR_BAD_ARG, "step %d: spline interval %d outside [0,%d]", s, st.interval, n_knots - 2);
    if (st.emit_begin != prev_end || st.emit_end < st.emit_begin || st.emit_end > E)
      return fail(SNSDE_ERR_BAD_ARG, "step %d: emit range [%d,%d) is not contiguous with the previous step", s, st.emit_begin, st.emit_end);
    prev_end = st.emit_end;
  }
  if (prev_end != E) return fail(SNSDE_ERR_BAD_ARG, "emit table has %d entries but steps consume %d", E, prev_end);
  for (int e = 0; e < E; ++e)
    if (emits_host[e].slot < 0 || emits_host[e].slot >= n_out)
      return fail(SNSDE_ERR_BAD_ARG, "emit %d: slot %d outside [0,%d)", e, emits_host[e].slot, n_out);

  cudaStream_t stream = (cudaStream_t)stream_v;
  CUDA_TRY(cudaSetDevice(p->device));
  if (p->ev_pending && p->last_stream != stream_v) CUDA_TRY(cudaStreamWaitEvent(stream, p->done_ev, 0));
  int rc = upload_tables(p, steps_host, S, emits_host, E, stream);
  if (rc != SNSDE_OK) return rc;
  struct Done {                      // record the completion event on every exit path after this point
    snsde_plan* p; cudaStream_t st; void* sv;
    ~Done() { if (p->done_ev && cudaEventRecord(p->done_ev, st) == cudaSuccess) { p->ev_pending = true; p->last_stream = sv; } }
  } done_guard{p, stream, stream_v};

  if (p->kind >= 1) {
    TcForwardArgs a;
    a.coeffs = coeffs_dev; a.coeff_row_stride = coeff_row_stride; a.y0 = y0_dev; a.B = B;
    a.steps = p->d_steps; a.steps_host = steps_host; a.S = S; a.emits = p->d_emits; a.n_init_emits = n_init_emits;
    a.n_out = n_out; a.row_slot = row_slot_dev; a.dW = dW_dev; a.seed = seed; a.row_offset = row_offset; a.out = out_dev;
    a.status = p->d_status;
    int nl = 0;
    cudaError_t e = p->kind == 1 ? tc_forward(p->tc, a, stream, &nl) : tcg_forward(p->tcg, a, stream, &nl);
    if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "tcgen05 kernel launch: %s", cudaGetErrorString(e));
    p->launches += nl;
    return SNSDE_OK;
  }

  FmaParams fp;
  fp.prog = pg;
  fp.wimg = p->d_wimg; fp.wimg_floats = p->wimg_floats;
  fp.coeffs = coeffs_dev; fp.coeff_row_stride = coeff_row_stride;
  fp.y0 = y0_dev; fp.B = B;
  fp.steps = p->d_steps; fp.S = S; fp.emits = p->d_emits; fp.n_init_emits = n_init_emits; fp.n_out = n_out;
  fp.row_slot = row_slot_dev; fp.dW = dW_dev; fp.seed = seed; fp.row_offset = row_offset; fp.out = out_dev;

  const int width = std::max(pg.H, pg.HH);
  const int nt = std::max(32, (width + 31) & ~31);
  const int r_max = nt <= 256 ? 16 : 4;
  int R = 1;
  while (R < r_max && (B + R - 1) / R > p->num_sms) R *= 2;
  // stage as much of the weight image as fits beside the activation buffers
  size_t fixed = fma_smem_bytes(pg, R, 0);
  while (fixed > (size_t)p->smem_optin && R > 1) { R /= 2; fixed = fma_smem_bytes(pg, R, 0); }
  if (fixed > (size_t)p->smem_optin) return fail(SNSDE_ERR_UNSUPPORTED, "activation buffers do not fit in shared memory");
  const int room = (int)((p->smem_optin - fixed) / sizeof(float)) & ~3;
  fp.smem_w_floats = std::min(p->wimg_floats, room);
  const size_t smem = fma_smem_bytes(pg, R, fp.smem_w_floats);
  cudaError_t e = fma_launch(fp, R, nt, smem, stream);
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "fma kernel launch (R=%d nt=%d smem=%zu): %s", R, nt, smem, cudaGetErrorString(e));
  p->launches += 1;
  return SNSDE_OK;
}

int snsde_philox_fill(uint64_t seed, uint64_t row_offset, int32_t S, int32_t B, int32_t H,
                      const float* sqrt_h_host, float* dW_dev, int device, void* stream_v) {
  if (S < 0 || B < 1 || H < 1 || !dW_dev || (S && !sqrt_h_host)) return fail(SNSDE_ERR_BAD_ARG, "bad philox_fill arguments");
  if (S == 0) return SNSDE_OK;
  cudaStream_t stream = (cudaStream_t)stream_v;
  CUDA_TRY(cudaSetDevice(device));
  float* d_sq = nullptr;
  CUDA_TRY(cudaMallocAsync(&d_sq, sizeof(float) * S, stream));
  CUDA_TRY(cudaMemcpyAsync(d_sq, sqrt_h_host, sizeof(float) * S, cudaMemcpyHostToDevice, stream));
  const size_t n = (size_t)S * B * H;
  const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
  philox_fill_kernel<<<grid, 256, 0, stream>>>(seed, row_offset, S, B, H, d_sq, dW_dev);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(d_sq, stream);
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "philox_fill launch: %s", cudaGetErrorString(e));
  return SNSDE_OK;
}

}  // extern "C"
