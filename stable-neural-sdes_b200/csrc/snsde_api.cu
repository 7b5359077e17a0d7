// C-ABI entry points (include/snsde.h): plan management, model -> kernel program compilation,
// weight re-layout, launch.  Host side only; kernels live in snsde_fma.cu / snsde_tc.cu.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include <cublas_v2.h>

#include "snsde_bwd.cuh"
#include "snsde_common.cuh"
#include "snsde_host.cuh"
#include "snsde_rng.cuh"
#include "snsde_tc.cuh"
#include "snsde_tcg.cuh"
#include "snsde_warp.cuh"

namespace snsde {
size_t fma_group_smem_floats(const Program& pg, int R, int method);
cudaError_t fma_launch(const FmaParams& p, int R, int method, size_t smem, cudaStream_t stream);
cudaError_t vec_tables_launch(const Program& pg, const float* wimg, const snsde_step* steps, const snsde_point* points,
                              int S, int npg, float* vtab, cudaStream_t stream);

// ---- error plumbing ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace snsde

using namespace snsde;
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
  bool warp_ok = false; WarpProg wprog;   // kind == 0: the program fits the warp-owned kernel (hidden <= 32, snsde_warp.cu)
  float* d_wimg_warp = nullptr; int wimg_warp_floats = 0, wimg_warp_cap = 0;
  TcPlan tc;                          // tensor-core path state (kind == 1)
  TcgPlan tcg;                        // general tensor-core path state (kind == 2)
  float* d_blob = nullptr; int blob_floats = 0;     // raw nn.Linear blob (the backward pass reads W in its [out][in] layout)
  snsde_step* d_steps = nullptr; int steps_cap = 0; std::vector<snsde_step> h_steps;
  snsde_emit* d_emits = nullptr; int emits_cap = 0; std::vector<snsde_emit> h_emits;
  snsde_point* d_points = nullptr; int points_cap = 0; std::vector<snsde_point> h_points;
  float* d_vtab = nullptr; size_t vtab_cap = 0;     // [S][npg][H] row-independent diffusion coefficient (FMA kernels)
  bool vtab_valid = false; int vtab_S = 0, vtab_npg = 0;   // cached while the weights and the evaluation times are unchanged
  cublasHandle_t cublas = nullptr;                  // weight-gradient GEMMs of the backward pass (created lazily)
  int64_t launches = 0;
  // sticky flags raised by the kernels: one word of mapped page-locked host memory (the rare device store travels over
  // PCIe; the host reads it without touching any stream)
  int* h_status = nullptr; int* d_status = nullptr;
  // The plan owns device tables (steps, emits, noise tables) that every solve reads: consecutive solves of one
  // plan are serialised across streams by this event (the caller's copies on other streams still overlap).
  cudaEvent_t done_ev = nullptr; void* last_stream = nullptr; bool ev_pending = false;
};

static bool is_time_opt(int io) { return io >= 3 && io <= 6; }
static bool is_emb_opt(int io) { return io == 2 || io == 4 || io == 6; }
static bool uses_control(int io) { return io == 0 || is_emb_opt(io); }

static int validate(const snsde_model_desc* d) {
  if (!d) return fail(SNSDE_ERR_BAD_ARG, "desc is NULL");
  if (d->family != SNSDE_FAMILY_BENCHMARK && d->family != SNSDE_FAMILY_TUTORIAL_LSDE && d->family != SNSDE_FAMILY_LATENT_SDE)
    return fail(SNSDE_ERR_BAD_ARG, "unknown family %d", d->family);
  if (d->family == SNSDE_FAMILY_LATENT_SDE && d->hidden < 2)
    return fail(SNSDE_ERR_BAD_ARG, "LatentSDE needs hidden >= 2 (latent width hidden-1 plus the KL accumulator channel)");
  if (d->input_channels < 1 || d->hidden < 1 || d->hidden_hidden < 1 || d->num_hidden_layers < 1)
    return fail(SNSDE_ERR_BAD_ARG, "C/H/HH/L must be >= 1");
  if (std::max(d->hidden, d->hidden_hidden) > 1024)
    return fail(SNSDE_ERR_UNSUPPORTED, "hidden sizes above 1024 are not supported");
  if (d->method != SNSDE_METHOD_EULER && d->method != SNSDE_METHOD_MILSTEIN && d->method != SNSDE_METHOD_SRK)
    return fail(SNSDE_ERR_UNSUPPORTED, "method %d: implemented methods are euler(0), milstein(1), srk(2)", d->method);
  if (d->precision < 0 || d->precision > SNSDE_PRECISION_AUTO) return fail(SNSDE_ERR_BAD_ARG, "bad precision %d", d->precision);
  if (d->family == SNSDE_FAMILY_BENCHMARK) {
    if (d->input_option < 0 || d->input_option > 6) return fail(SNSDE_ERR_BAD_ARG, "input_option %d not in 0..6", d->input_option);
    if (d->noise_option < 0 || d->noise_option > 19)
      return fail(SNSDE_ERR_BAD_ARG, "Unknown noise_option %d.", d->noise_option);   // neuralsde.py:288
    if ((d->input_option == 0 || is_emb_opt(d->input_option)) && d->hidden != d->hidden_hidden)
      return fail(SNSDE_ERR_BAD_ARG, "input_option %d requires hidden_hidden == hidden (emb is Linear(2H,H), neuralsde.py:154,210)", d->input_option);
  }
  const int n_ops = d->family == SNSDE_FAMILY_TUTORIAL_LSDE ? 6 + 2 * d->num_hidden_layers : 6 + d->num_hidden_layers;
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
  } else if (d->family == SNSDE_FAMILY_LATENT_SDE) {
    n += HH * (H + 1) + HH;                    // linear_in: Linear(hidden+2-1, HH)     latent_sde.py:48
    n += (L - 1) * (HH * HH + HH);             // linears                               :49-50
    n += (H - 1) * HH + (H - 1);               // linear_out: Linear(HH, hidden-1)      :51
    n += 3;                                    // theta, mu, sigma buffers              :35-37
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
  o.g_w = -1; o.g_ldw = 0; o.g_col = 0; o.g_col2 = 0; o.g_b = -1;
  o.src_op = SRC_ABSENT; o.src2_op = SRC_ABSENT;
  return o;
}

struct BlobCursor {
  const float* base;
  const float* q;
  explicit BlobCursor(const float* b) : base(b), q(b) {}
  const float* take(size_t n) { const float* r = q; q += n; return r; }
  int off(const float* ptr) const { return (int)(ptr - base); }
};

// Appends ops to a program and records, per op, which earlier op produced each of its inputs (the backward
// pass keeps every op's output in its own slot) and where its parameters sit in the blob.
struct ProgramBuilder {
  Program& pg;
  const BlobCursor& bc;
  int n = 0;
  int writer[BUF_COUNT];
  ProgramBuilder(Program& p, const BlobCursor& b) : pg(p), bc(b) { for (int& w : writer) w = SRC_ABSENT; }
  int producer(int buf) const { return buf == BUF_Y ? SRC_STATE : (buf == BUF_X ? SRC_CONTROL : (buf < 0 ? SRC_ABSENT : writer[buf])); }
  // W: the nn.Linear weight [N][ldw] the op reads (src at columns col.., src2 at col2..), bias b
  DenseOp& push(DenseOp o, int part, const float* W, int ldw, int col, int col2, const float* b) {
    o.part = part;
    o.g_w = W ? bc.off(W) : -1; o.g_ldw = ldw; o.g_col = col; o.g_col2 = col2; o.g_b = b ? bc.off(b) : -1;
    o.src_op = producer(o.src); o.src2_op = producer(o.src2);
    pg.ops[n] = o;
    if (o.dst >= 0) writer[o.dst] = n;
    return pg.ops[n++];
  }
};

static void compile_benchmark(const snsde_model_desc& d, const float* blob, Program& pg, ImageBuilder& ib) {
  const int C = d.input_channels, H = d.hidden, HH = d.hidden_hidden, L = d.num_hidden_layers;
  const int io = d.input_option, no = d.noise_option;
  const int tau = is_time_opt(io) ? 2 : 0;
  BlobCursor bc(blob);
  const float* Wi = bc.take((size_t)H * C); const float* bi = bc.take(H);
  const float* Win = bc.take((size_t)HH * (H + tau)); const float* bin = bc.take(HH);
  const float *We = nullptr, *be = nullptr;
  if (is_emb_opt(io)) { We = bc.take((size_t)H * 2 * H); be = bc.take(H); }
  std::vector<const float*> Wl(L - 1), bl(L - 1);
  for (int l = 0; l < L - 1; ++l) { Wl[l] = bc.take((size_t)HH * HH); bl[l] = bc.take(HH); }
  const float* Wo = bc.take((size_t)H * HH); const float* bo = bc.take(H);
  const float* theta_p = bc.take(1);
  const float theta = *theta_p;

  memset(&pg, 0, sizeof(pg));
  pg.C = C; pg.H = H; pg.HH = HH;
  pg.ld = (std::max(std::max(H, HH), C) + 3) & ~3;
  pg.uses_control = uses_control(io);
  ProgramBuilder pb(pg, bc);
  int cur;
  if (pg.uses_control) {                      // Xt = initial_network(X(t))      neuralsde.py:296-297
    pb.push(make_op(BUF_U, BUF_X, C, H, ib.add_T(Wi, H, C, 0, C), ib.add_vec(bi, H), io == 0 ? ACT_RELU : ACT_NONE),
            0, Wi, C, 0, 0, bi);
  }
  if (io == 0) {
    cur = BUF_U;                              // z = Xt                         :206-207
  } else {                                    // yy = linear_in([tf,] y)        :200-204
    DenseOp o = make_op(BUF_A, BUF_Y, H, HH, ib.add_T(Win, HH, H + tau, tau, H), ib.add_vec(bin, HH),
                        is_emb_opt(io) ? ACT_NONE : ACT_RELU);
    if (tau) { o.tmode = TM_SINCOS; o.tw_off = ib.add_T(Win, HH, H + tau, 0, 2); }
    pb.push(o, 0, Win, H + tau, tau, 0, bin);
    cur = BUF_A;
    if (is_emb_opt(io)) {                     // z = emb(cat(yy, Xt))           :210
      DenseOp e = make_op(BUF_B, BUF_A, H, H, ib.add_T(We, H, 2 * H, 0, H), ib.add_vec(be, H), ACT_RELU);
      e.src2 = BUF_U; e.K2 = H; e.w2_off = ib.add_T(We, H, 2 * H, H, H);
      pb.push(e, 0, We, 2 * H, 0, H, be);
      cur = BUF_B;
    }
  }
  // noise networks (row-independent ones are tabulated per step, not evaluated per row)  :170-179, 271-286
  TailOp& t = pg.tail;
  t.geometric = (io == 5 || io == 6);
  t.clip_drift = 1;
  t.bounded = 1;
  t.s_theta = 1.f / (1.f + expf(-theta));
  t.milstein = d.method == SNSDE_METHOD_MILSTEIN;
  t.special = SP_NONE; t.mult = MU_ONE; t.coef_src = CO_NONE; t.coef_scalar = 0.f;
  t.g_theta = bc.off(theta_p); t.g_sigma = -1; t.coef_op = -1;
  if (no == 0) t.special = SP_ZERO;
  if (no >= 1 && no <= 3) {
    const float* sg = bc.take(1);
    t.coef_src = CO_SCALAR; t.coef_scalar = expf(*sg); t.g_sigma = bc.off(sg);
    t.mult = no == 1 ? MU_ONE : (no == 2 ? MU_T : MU_Y);
  }
  if (no >= 4 && no <= 6) {
    const float* sd = bc.take(H);
    std::vector<float> e(H);
    for (int j = 0; j < H; ++j) e[j] = expf(sd[j]);
    t.coef_src = CO_IMG; t.coef_ref = ib.add_vec(e.data(), H); t.g_sigma = bc.off(sd);
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
    pb.push(o, 1, W1, 2, 2, 0, b1);
    t.coef_src = CO_VBUF; t.coef_ref = BUF_V0;
    if (no >= 16) {
      const float* W2 = bc.take((size_t)H * H); const float* b2 = bc.take(H);
      DenseOp o2 = make_op(BUF_V1, BUF_V0, H, H, ib.add_T(W2, H, H, 0, H), ib.add_vec(b2, H), ACT_RELU);
      o2.vec = 1;
      pb.push(o2, 1, W2, H, 0, 0, b2);
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
    const DenseOp& o1 = pb.push(o, 1, W1, H + 2, 2, 0, b1);
    t.vjp_kind = deep ? 2 : 1; t.vjp_w1 = o1.w_off; t.vjp_w2 = -1; t.vjp_h1 = BUF_P;
    if (deep) {
      const float* W2 = bc.take((size_t)H * H); const float* b2 = bc.take(H);
      const DenseOp& o2 = pb.push(make_op(BUF_Q, BUF_P, H, H, ib.add_T(W2, H, H, 0, H), ib.add_vec(b2, H), ACT_RELU),
                                  1, W2, H, 0, 0, b2);
      t.vjp_w2 = o2.w_off;
    }
    t.coef_src = CO_RBUF; t.coef_ref = BUF_Q; t.coef_op = pb.n - 1;
    t.mult = (no == 15 || no == 19) ? MU_Y : MU_ONE;
  }
  // shared MLP tail: relu -> (Linear, relu)* -> linear_out                     :212-217
  for (int l = 0; l < L - 1; ++l) {
    const int dst = (cur == BUF_A) ? BUF_B : BUF_A;
    pb.push(make_op(dst, cur, HH, HH, ib.add_T(Wl[l], HH, HH, 0, HH), ib.add_vec(bl[l], HH), ACT_RELU),
            0, Wl[l], HH, 0, 0, bl[l]);
    cur = dst;
  }
  DenseOp fo = make_op(BUF_NONE, cur, HH, H, ib.add_T(Wo, H, HH, 0, HH), ib.add_vec(bo, H), ACT_NONE);
  fo.final_drift = 1;
  pb.push(fo, 0, Wo, HH, 0, 0, bo);
  pg.n_ops = pb.n;
}

static void compile_tutorial(const snsde_model_desc& d, const float* blob, Program& pg, ImageBuilder& ib) {
  const int C = d.input_channels, H = d.hidden, HH = d.hidden_hidden, L = d.num_hidden_layers;
  BlobCursor bc(blob);
  memset(&pg, 0, sizeof(pg));
  pg.C = C; pg.H = H; pg.HH = HH;
  pg.ld = (std::max(std::max(H, HH), C) + 3) & ~3;
  pg.uses_control = 1;
  ProgramBuilder pb(pg, bc);
  const float* WX = bc.take((size_t)H * C); const float* bX = bc.take(H);
  const float* We = bc.take((size_t)H * 2 * H); const float* be = bc.take(H);
  pb.push(make_op(BUF_U, BUF_X, C, H, ib.add_T(WX, H, C, 0, C), ib.add_vec(bX, H), ACT_NONE), 0, WX, C, 0, 0, bX);
  DenseOp e = make_op(BUF_A, BUF_Y, H, H, ib.add_T(We, H, 2 * H, 0, H), ib.add_vec(be, H), ACT_NONE);
  e.src2 = BUF_U; e.K2 = H; e.w2_off = ib.add_T(We, H, 2 * H, H, H);     // emb(cat(y, Xt))
  pb.push(e, 0, We, 2 * H, 0, H, be);
  auto mlp = [&](int cur, int b0, int b1, bool vec) {
    int in = H;
    for (int l = 0; l < L; ++l) {            // Linear(in->HH) + LipSwish, L times
      const float* W = bc.take((size_t)HH * in); const float* b = bc.take(HH);
      const int dst = (cur == b0) ? b1 : b0;
      DenseOp o = make_op(dst, cur, in, HH, ib.add_T(W, HH, in, 0, in), ib.add_vec(b, HH), ACT_LIPSWISH);
      o.vec = vec;
      pb.push(o, vec ? 1 : 0, W, in, 0, 0, b);
      cur = dst; in = HH;
    }
    const float* W = bc.take((size_t)H * HH); const float* b = bc.take(H);
    const int dst = (cur == b0) ? b1 : b0;
    DenseOp o = make_op(dst, cur, HH, H, ib.add_T(W, H, HH, 0, HH), ib.add_vec(b, H), ACT_NONE);
    o.vec = vec;
    pb.push(o, vec ? 1 : 0, W, HH, 0, 0, b);
    return dst;
  };
  const int fcur = mlp(BUF_A, BUF_A, BUF_B, false);
  const float* Wlo = bc.take((size_t)H * H); const float* blo = bc.take(H);
  const float* Wni = bc.take((size_t)H); const float* bni = bc.take(H);
  DenseOp ni = make_op(BUF_V0, BUF_NONE, 0, H, -1, ib.add_vec(bni, H), ACT_NONE);
  ni.vec = 1; ni.tmode = TM_RAW; ni.tw_off = ib.add_T(Wni, H, 1, 0, 1);
  pb.push(ni, 1, Wni, 1, 1, 0, bni);
  const int gcur = mlp(BUF_V0, BUF_V0, BUF_V1, true);
  DenseOp fo = make_op(BUF_NONE, fcur, H, H, ib.add_T(Wlo, H, H, 0, H), ib.add_vec(blo, H), ACT_NONE);
  fo.final_drift = 1;
  pb.push(fo, 0, Wlo, H, 0, 0, blo);
  pg.n_ops = pb.n;
  TailOp& t = pg.tail;
  memset(&t, 0, sizeof(t));
  t.coef_src = CO_VBUF; t.coef_ref = gcur; t.mult = MU_ONE; t.special = SP_NONE;
  t.bounded = 0; t.clip_drift = 0; t.geometric = 0; t.s_theta = 1.f;
  t.milstein = d.method == SNSDE_METHOD_MILSTEIN;
  t.g_theta = -1; t.g_sigma = -1; t.coef_op = -1;
}

// LatentSDE.f_aug / g_aug (torch-ists/torch_ists/diff_module/NSDE/latent_sde.py:57-90): the posterior drift is an MLP
// on (sin t, cos t, y[:, :-1]) with relu between layers and NO clip; the KL channel and the masked constant diffusion are
// tail work (TailOp::latent).
static void compile_latent(const snsde_model_desc& d, const float* blob, Program& pg, ImageBuilder& ib) {
  const int H = d.hidden, HH = d.hidden_hidden, L = d.num_hidden_layers, Hl = H - 1;
  BlobCursor bc(blob);
  memset(&pg, 0, sizeof(pg));
  pg.C = d.input_channels; pg.H = H; pg.HH = HH;
  pg.ld = (std::max(H, HH) + 3) & ~3;
  pg.uses_control = 0;
  ProgramBuilder pb(pg, bc);
  const float* Win = bc.take((size_t)HH * (Hl + 2)); const float* bin = bc.take(HH);
  DenseOp o = make_op(BUF_A, BUF_Y, Hl, HH, ib.add_T(Win, HH, Hl + 2, 2, Hl), ib.add_vec(bin, HH), ACT_RELU);
  o.tmode = TM_SINCOS; o.tw_off = ib.add_T(Win, HH, Hl + 2, 0, 2);
  pb.push(o, 0, Win, Hl + 2, 2, 0, bin);
  int cur = BUF_A;
  for (int l = 0; l < L - 1; ++l) {
    const float* W = bc.take((size_t)HH * HH); const float* b = bc.take(HH);
    const int dst = (cur == BUF_A) ? BUF_B : BUF_A;
    pb.push(make_op(dst, cur, HH, HH, ib.add_T(W, HH, HH, 0, HH), ib.add_vec(b, HH), ACT_RELU), 0, W, HH, 0, 0, b);
    cur = dst;
  }
  const float* Wo = bc.take((size_t)Hl * HH); const float* bo = bc.take(Hl);
  DenseOp fo = make_op(BUF_NONE, cur, HH, Hl, ib.add_T(Wo, Hl, HH, 0, HH), ib.add_vec(bo, Hl), ACT_NONE);
  fo.final_drift = 1;
  pb.push(fo, 0, Wo, HH, 0, 0, bo);
  pg.n_ops = pb.n;
  const float theta = *bc.take(1), mu = *bc.take(1), sigma = *bc.take(1);
  TailOp& t = pg.tail;
  memset(&t, 0, sizeof(t));
  t.coef_src = CO_SCALAR; t.coef_scalar = sigma; t.mult = MU_ONE; t.special = SP_NONE;
  t.bounded = 0; t.clip_drift = 0; t.geometric = 0; t.s_theta = 1.f;
  t.milstein = d.method == SNSDE_METHOD_MILSTEIN;
  t.g_theta = -1; t.g_sigma = -1; t.coef_op = -1;
  t.latent = 1; t.lat_theta = theta; t.lat_mu = mu;
  // _stable_division(a, b, eps=1e-7): b = where(|b| > eps, b, eps * sign(b))        latent_sde.py:24-26
  const float eps = 1e-7f;
  t.lat_div = fabsf(sigma) > eps ? sigma : eps * (sigma > 0.f ? 1.f : (sigma < 0.f ? -1.f : 0.f));
}

// ---- Philox materialisation kernel -------------------------------------------------------------
__global__ void philox_fill_kernel(unsigned long long seed, unsigned long long row_offset, int S, int B, int H,
                                   const snsde_step* __restrict__ steps, float* __restrict__ dW, float* __restrict__ dU) {
  const size_t n = (size_t)S * B * H;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % H);
    const size_t sb = i / H;
    const int b = (int)(sb % B), s = (int)(sb / B);
    const unsigned long long gb = row_offset + (unsigned long long)b;
    const float h = steps[s].h, sqrt_h = steps[s].sqrt_h;
    float nrm[4];
    philox_normals4(seed, (uint32_t)j, (uint32_t)(gb >> 2), (uint32_t)s, nrm);
    const float w = __fmul_rn(pick4(nrm, (int)(gb & 3ull)), sqrt_h);
    if (dW != nullptr) dW[i] = w;
    if (dU != nullptr) {
      float nu[4];
      philox_normals4_u(seed, (uint32_t)j, (uint32_t)(gb >> 2), (uint32_t)s, nu);
      dU[i] = levy_U(w, pick4(nu, (int)(gb & 3ull)), h, sqrt_h);
    }
  }
}

// ---- table upload (steps / emits / points are host data; uploaded only when they change) ------------
template <typename T>
static int upload_table(T*& d_ptr, int& cap, std::vector<T>& host_copy, const T* src, int n, cudaStream_t stream,
                        bool* changed = nullptr) {
  if (n > cap) {
    cudaFree(d_ptr); d_ptr = nullptr; host_copy.clear(); cap = 0;
    CUDA_TRY(cudaMalloc(&d_ptr, sizeof(T) * (size_t)std::max(n, 64)));
    cap = std::max(n, 64);
  }
  if ((int)host_copy.size() != n || (n && memcmp(host_copy.data(), src, sizeof(T) * n) != 0)) {
    host_copy.assign(src, src + n);
    if (changed) *changed = true;
    // pageable source: the runtime stages it before returning
    if (n) CUDA_TRY(cudaMemcpyAsync(d_ptr, src, sizeof(T) * n, cudaMemcpyHostToDevice, stream));
  }
  return SNSDE_OK;
}

static int upload_tables(snsde_plan* p, const snsde_step* steps, int S, const snsde_emit* emits, int E,
                         const snsde_point* points, cudaStream_t stream) {
  bool times_changed = false;
  int rc = upload_table(p->d_steps, p->steps_cap, p->h_steps, steps, S, stream, &times_changed);
  if (rc != SNSDE_OK) return rc;
  if (emits || E == 0) rc = upload_table(p->d_emits, p->emits_cap, p->h_emits, emits, E, stream);
  if (rc != SNSDE_OK) return rc;
  if (points) rc = upload_table(p->d_points, p->points_cap, p->h_points, points, S * kSrkPoints, stream, &times_changed);
  if (times_changed) p->vtab_valid = false;              // the row-independent coefficient table is per evaluation time
  return rc;
}

static int validate_steps(const Program& pg, int n_knots, const snsde_step* steps_host, int S) {
  for (int s = 0; s < S; ++s)
    if (pg.uses_control && (steps_host[s].interval < 0 || steps_host[s].interval > n_knots - 2))
      return fail(SNSDE_ERR_BAD_ARG, "step %d: spline interval %d outside [0,%d]", s, steps_host[s].interval, n_knots - 2);
  return SNSDE_OK;
}

static int validate_control(const Program& pg, const float* coeffs_dev, int64_t coeff_row_stride, int n_knots) {
  if (!pg.uses_control) return SNSDE_OK;
  if (!coeffs_dev) return fail(SNSDE_ERR_BAD_ARG, "model reads the control path but coeffs is NULL");
  if (n_knots < 2) return fail(SNSDE_ERR_BAD_ARG, "need at least 2 knots");
  if (coeff_row_stride < (int64_t)(n_knots - 1) * 4 * pg.C)
    return fail(SNSDE_ERR_BAD_ARG, "coeff_row_stride %lld < (K-1)*4C", (long long)coeff_row_stride);
  if (((uintptr_t)coeffs_dev & 15) || (coeff_row_stride & 3))
    return fail(SNSDE_ERR_BAD_ARG, "coeffs must be 16-byte aligned with a row stride multiple of 4 floats");
  return SNSDE_OK;
}

// Serialises consecutive uses of one plan across streams (the plan owns device tables every launch reads) and
// records the completion event on every exit path.
struct PlanStreamGuard {
  snsde_plan* p; cudaStream_t st; void* sv;
  PlanStreamGuard(snsde_plan* p_, void* sv_) : p(p_), st((cudaStream_t)sv_), sv(sv_) {
    if (p->ev_pending && p->last_stream != sv) cudaStreamWaitEvent(st, p->done_ev, 0);
  }
  ~PlanStreamGuard() {
    if (p->done_ev && cudaEventRecord(p->done_ev, st) == cudaSuccess) { p->ev_pending = true; p->last_stream = sv; }
  }
};

// Rows per group, row groups per CTA, warps per group and whether the weight image is staged in shared memory, for
// the FMA-style kernels.  group_floats(R) = shared-memory floats one group needs.
struct GroupConfig { int R, nw, groups, smem_w_floats; size_t smem; };
template <typename F>
static bool pick_group_config(const snsde_plan* p, int B, F group_floats, bool allow_r8, GroupConfig& g) {
  const Program& pg = p->prog;
  g.nw = (std::max(pg.H, pg.HH) + 31) / 32;
  const int max_threads = g.nw > 16 ? 1024 : 512;
  const size_t optin = (size_t)p->smem_optin, img = (size_t)p->wimg_floats * sizeof(float);
  const bool img_fits = img + group_floats(4) * sizeof(float) <= optin;
  // Parallelism first: the chain of dependent ops makes every group latency-bound, so small batches take fewer rows per
  // group (down to one) until there is about a warp per scheduler; 8 rows per group amortise each weight read over
  // twice the FMAs and pay when the image is read through L2 or the batch is large.
  const int want_warps = 4 * p->num_sms;
  int R = 1;
  if (g.nw > 16) R = 4;
  else if (allow_r8 && (B / 8) * g.nw >= (img_fits ? want_warps : p->num_sms / 2)) R = 8;
  else if ((B / 4) * g.nw >= p->num_sms / 2) R = 4;
  for (;; R = (R == 8 ? 4 : 1)) {
    if (group_floats(R) * sizeof(float) <= optin) break;
    if (R == 1 || (R == 4 && g.nw > 16)) return false;
  }
  g.R = R;
  const size_t gf = group_floats(R) * sizeof(float);
  const int n_groups = (B + R - 1) / R;
  int G = std::max(1, n_groups / std::max(1, p->num_sms));             // one wave of CTAs first
  G = std::min(G, std::min(15, max_threads / (g.nw * 32)));             // 15 named barriers, thread limit
  G = std::max(G, 1);
  while (G > 1 && (size_t)G * gf > optin) --G;
  while (G > 1 && (size_t)G * gf + img > optin && gf + img <= optin) --G;   // trade groups for the staged image
  g.groups = G;
  g.smem_w_floats = ((size_t)G * gf + img <= optin) ? p->wimg_floats : 0;   // all or nothing
  g.smem = (size_t)g.smem_w_floats * sizeof(float) + (size_t)G * gf;
  return true;
}

static int ensure_vtab(snsde_plan* p, size_t floats) {
  if (floats > p->vtab_cap) {
    cudaFree(p->d_vtab); p->d_vtab = nullptr; p->vtab_cap = 0;
    CUDA_TRY(cudaMalloc(&p->d_vtab, floats * sizeof(float)));
    p->vtab_cap = floats;
  }
  return SNSDE_OK;
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
  SNSDE_API_BEGIN
  if (!out_plan) return fail(SNSDE_ERR_BAD_ARG, "out_plan is NULL");
  *out_plan = nullptr;
  const int rc = validate(desc);
  if (rc != SNSDE_OK) return rc;
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(SNSDE_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.err));
  std::unique_ptr<snsde_plan> p(new snsde_plan());
  p->desc = *desc;
  p->device = device;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  p->num_sms = prop.multiProcessorCount;
  p->smem_optin = (int)prop.sharedMemPerBlockOptin;
  const bool force_general = getenv("SNSDE_FORCE_TCG") != nullptr;           // testing aid: general kernel even where the resident one applies
  const bool srk = desc->method == SNSDE_METHOD_SRK;                          // SRK stages run on the FMA kernels
  const bool tc_ok = !srk && tc_supported(*desc, prop.major, p->smem_optin) && !force_general;
  const int no_ = desc->noise_option;
  // Milstein through a state-dependent noise network needs the full vjp: implemented in the FMA kernel only
  const bool net_vjp = desc->family == SNSDE_FAMILY_BENCHMARK && desc->method == SNSDE_METHOD_MILSTEIN &&
                       (no_ == 14 || no_ == 15 || no_ == 18 || no_ == 19);
  const bool tcg_ok = !srk && tcg_supported(*desc, prop.major, p->smem_optin) && !net_vjp;
  if (desc->precision == SNSDE_PRECISION_TC && !tc_ok && !tcg_ok)
    return fail(SNSDE_ERR_UNSUPPORTED, "tensor-core path does not support this model/shape/method/device: %s",
                srk ? "method 'srk' runs on the fp32 kernels" : tcg_unsupported_reason());
  p->kind = desc->precision == SNSDE_PRECISION_FP32 ? 0 : (tc_ok ? 1 : (tcg_ok ? 2 : 0));
  CUDA_TRY(cudaHostAlloc(&p->h_status, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
  *p->h_status = 0;
  if (cudaHostGetDevicePointer(&p->d_status, p->h_status, 0) != cudaSuccess) {
    cudaFreeHost(p->h_status);
    return fail(SNSDE_ERR_CUDA, "cannot map the plan status word");
  }
  if (cudaEventCreateWithFlags(&p->done_ev, cudaEventDisableTiming) != cudaSuccess) p->done_ev = nullptr;
  *out_plan = p.release();
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

int snsde_plan_destroy(snsde_plan* p) {
  SNSDE_API_BEGIN
  if (!p) return SNSDE_OK;
  DeviceGuard guard(p->device);
  cudaFree(p->d_wimg);
  cudaFree(p->d_wimg_warp);
  cudaFree(p->d_blob);
  cudaFree(p->d_steps);
  cudaFree(p->d_emits);
  cudaFree(p->d_points);
  cudaFree(p->d_vtab);
  cudaFreeHost(p->h_status);
  if (p->cublas) cublasDestroy(p->cublas);
  if (p->done_ev) cudaEventDestroy(p->done_ev);
  tc_release(p->tc);
  tcg_release(p->tcg);
  delete p;
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

int snsde_plan_kernel_kind(const snsde_plan* p) {
  if (!p) return fail(SNSDE_ERR_BAD_ARG, "plan is NULL");
  return p->kind;
}

int64_t snsde_plan_launch_count(const snsde_plan* p) { return p ? p->launches : 0; }

int snsde_plan_fma_variant(const snsde_plan* p) {
  if (!p) return fail(SNSDE_ERR_BAD_ARG, "plan is NULL");
  if (!p->has_weights) return fail(SNSDE_ERR_NO_WEIGHTS, "the variant is chosen when the weights are set");
  return (p->kind == 0 && p->warp_ok) ? 1 : 0;
}

int snsde_plan_status(snsde_plan* p, void* stream_v) {
  SNSDE_API_BEGIN
  if (!p) return fail(SNSDE_ERR_BAD_ARG, "plan is NULL");
  DeviceGuard guard(p->device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream_v));
  volatile int* w = p->h_status;
  const int flags = *w;
  *w = 0;
  return flags;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

int snsde_plan_status_nowait(snsde_plan* p) {
  if (!p) return fail(SNSDE_ERR_BAD_ARG, "plan is NULL");
  volatile int* w = p->h_status;
  const int flags = *w;
  if (flags) *w = 0;
  return flags;
}

int snsde_plan_set_weights(snsde_plan* p, const float* blob, int64_t n_floats, int on_device, void* stream_v) {
  SNSDE_API_BEGIN
  if (!p || !blob) return fail(SNSDE_ERR_BAD_ARG, "plan/blob is NULL");
  const int64_t want = weight_count(&p->desc);
  if (n_floats != want) return fail(SNSDE_ERR_BAD_ARG, "weight blob has %lld floats, model needs %lld", (long long)n_floats, (long long)want);
  cudaStream_t stream = (cudaStream_t)stream_v;
  DeviceGuard guard(p->device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
  PlanStreamGuard order(p, stream_v);                 // earlier solves of this plan still read the old images
  std::vector<float> host;
  if (on_device) {
    host.resize(n_floats);
    CUDA_TRY(cudaMemcpyAsync(host.data(), blob, n_floats * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    blob = host.data();
  }
  ImageBuilder ib;
  if (p->desc.family == SNSDE_FAMILY_BENCHMARK) compile_benchmark(p->desc, blob, p->prog, ib);
  else if (p->desc.family == SNSDE_FAMILY_LATENT_SDE) compile_latent(p->desc, blob, p->prog, ib);
  else compile_tutorial(p->desc, blob, p->prog, ib);
  ib.pad4();
  {
    std::vector<float> wimg_warp;
    p->warp_ok = getenv("SNSDE_NO_WARP") == nullptr &&                                               // env: testing aid
                 warp_build(p->prog, p->desc.method, blob, ib.img.data(), p->wprog, wimg_warp) &&
                 warp_smem_bytes((int)wimg_warp.size(), 1, 1, 0, 0, false, p->desc.method == SNSDE_METHOD_SRK) <= (size_t)p->smem_optin;
    if (p->warp_ok) {
      if ((int)wimg_warp.size() > p->wimg_warp_cap) {
        cudaFree(p->d_wimg_warp);
        p->d_wimg_warp = nullptr; p->wimg_warp_cap = 0;
        CUDA_TRY(cudaMalloc(&p->d_wimg_warp, wimg_warp.size() * sizeof(float)));
        p->wimg_warp_cap = (int)wimg_warp.size();
      }
      p->wimg_warp_floats = (int)wimg_warp.size();
      CUDA_TRY(cudaMemcpyAsync(p->d_wimg_warp, wimg_warp.data(), wimg_warp.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    }
  }
  if ((int)ib.img.size() > p->wimg_floats) {
    cudaFree(p->d_wimg);
    p->d_wimg = nullptr; p->wimg_floats = 0;
    CUDA_TRY(cudaMalloc(&p->d_wimg, ib.img.size() * sizeof(float)));
  }
  p->wimg_floats = (int)ib.img.size();
  // pageable source: the runtime stages it before returning, so `ib` may die at scope exit
  CUDA_TRY(cudaMemcpyAsync(p->d_wimg, ib.img.data(), ib.img.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
  if ((int)n_floats > p->blob_floats) {
    cudaFree(p->d_blob);
    p->d_blob = nullptr; p->blob_floats = 0;
    CUDA_TRY(cudaMalloc(&p->d_blob, (size_t)n_floats * sizeof(float)));
    p->blob_floats = (int)n_floats;
  }
  CUDA_TRY(cudaMemcpyAsync(p->d_blob, blob, (size_t)n_floats * sizeof(float), cudaMemcpyHostToDevice, stream));
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
  p->vtab_valid = false;                              // the coefficient table is a function of the weights
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

int snsde_forward(snsde_plan* p, const float* coeffs_dev, int64_t coeff_row_stride, int32_t n_knots,
                  const float* y0_dev, int32_t B, const snsde_step* steps_host, int32_t S,
                  const snsde_emit* emits_host, int32_t E, int32_t n_init_emits, int32_t n_out,
                  const snsde_point* points_host, const int32_t* row_slot_dev,
                  const float* dW_dev, const float* dU_dev, uint64_t seed, uint64_t row_offset,
                  float* out_dev, void* stream_v) {
  SNSDE_API_BEGIN
  if (!p) return fail(SNSDE_ERR_BAD_ARG, "plan is NULL");
  if (!p->has_weights) return fail(SNSDE_ERR_NO_WEIGHTS, "snsde_forward before snsde_plan_set_weights");
  if (!y0_dev || !out_dev) return fail(SNSDE_ERR_BAD_ARG, "y0/out is NULL");
  if (B < 1 || S < 0 || E < 0 || n_out < 1) return fail(SNSDE_ERR_BAD_ARG, "bad sizes B=%d S=%d E=%d n_out=%d", B, S, E, n_out);
  if ((S && !steps_host) || (E && !emits_host)) return fail(SNSDE_ERR_BAD_ARG, "step/emit table is NULL");
  if (n_init_emits < 0 || n_init_emits > E) return fail(SNSDE_ERR_BAD_ARG, "n_init_emits out of range");
  const bool srk = p->desc.method == SNSDE_METHOD_SRK;
  if (srk && S && !points_host) return fail(SNSDE_ERR_BAD_ARG, "method srk needs the per-step evaluation points");
  if (srk && dW_dev && !dU_dev) return fail(SNSDE_ERR_BAD_ARG, "method srk with explicit increments needs dU (space-time Levy integrals) beside dW");
  const Program& pg = p->prog;
  int rc = validate_control(pg, coeffs_dev, coeff_row_stride, n_knots);
  if (rc != SNSDE_OK) return rc;
  rc = validate_steps(pg, n_knots, steps_host, S);
  if (rc != SNSDE_OK) return rc;
  if (srk && pg.uses_control)
    for (int i = 0; i < S * kSrkPoints; ++i)
      if (points_host[i].interval < 0 || points_host[i].interval > n_knots - 2)
        return fail(SNSDE_ERR_BAD_ARG, "point %d: spline interval %d outside [0,%d]", i, points_host[i].interval, n_knots - 2);
  int prev_end = n_init_emits;
  for (int s = 0; s < S; ++s) {
    const snsde_step& st = steps_host[s];
    if (st.emit_begin != prev_end || st.emit_end < st.emit_begin || st.emit_end > E)
      return fail(SNSDE_ERR_BAD_ARG, "step %d: emit range [%d,%d) is not contiguous with the previous step", s, st.emit_begin, st.emit_end);
    prev_end = st.emit_end;
  }
  if (prev_end != E) return fail(SNSDE_ERR_BAD_ARG, "emit table has %d entries but steps consume %d", E, prev_end);
  for (int e = 0; e < E; ++e)
    if (emits_host[e].slot < 0 || emits_host[e].slot >= n_out)
      return fail(SNSDE_ERR_BAD_ARG, "emit %d: slot %d outside [0,%d)", e, emits_host[e].slot, n_out);

  cudaStream_t stream = (cudaStream_t)stream_v;
  DeviceGuard guard(p->device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
  PlanStreamGuard order(p, stream_v);
  rc = upload_tables(p, steps_host, S, emits_host, E, srk ? points_host : nullptr, stream);
  if (rc != SNSDE_OK) return rc;

  if (p->kind >= 1) {
    TcForwardArgs a;
    a.coeffs = coeffs_dev; a.coeff_row_stride = coeff_row_stride; a.n_knots = n_knots; a.y0 = y0_dev; a.B = B;
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
  fp.steps = p->d_steps; fp.S = S; fp.points = srk ? p->d_points : nullptr;
  fp.emits = p->d_emits; fp.n_init_emits = n_init_emits; fp.n_out = n_out;
  fp.row_slot = row_slot_dev; fp.dW = dW_dev; fp.dU = dU_dev; fp.seed = seed; fp.row_offset = row_offset; fp.out = out_dev;
  fp.vtab = nullptr;

  fp.groups = 1; fp.nw = 1; fp.smem_w_floats = 0;
  if (pg.tail.coef_src == CO_VBUF && S > 0) {
    const int npg = srk ? kSrkGPoints : 1;
    if (!p->vtab_valid || p->vtab_S != S || p->vtab_npg != npg) {
      p->vtab_valid = false;
      rc = ensure_vtab(p, (size_t)S * npg * pg.H);
      if (rc != SNSDE_OK) return rc;
      cudaError_t e = vec_tables_launch(pg, p->d_wimg, p->d_steps, fp.points, S, npg, p->d_vtab, stream);
      if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "coefficient table kernel launch: %s", cudaGetErrorString(e));
      p->launches += 1;
      p->vtab_valid = true; p->vtab_S = S; p->vtab_npg = npg;
    }
    fp.vtab = p->d_vtab;
  }
  if (p->warp_ok && B <= kWarpMaxRows) {   // hidden <= 32 and rows scarce: a pair of warps owns each row end to end
    fp.wimg = p->d_wimg_warp; fp.wimg_floats = p->wimg_warp_floats;
    cudaError_t e = warp_launch(fp, p->wprog, p->desc.method, E, p->num_sms, p->smem_optin, stream);
    if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "warp kernel launch: %s", cudaGetErrorString(e));
    p->launches += 1;
    return SNSDE_OK;
  }
  GroupConfig gc;
  const int method_ = p->desc.method;
  if (!pick_group_config(p, B, [&](int R) { return fma_group_smem_floats(pg, R, method_); }, !srk, gc))
    return fail(SNSDE_ERR_UNSUPPORTED, "activation buffers do not fit in shared memory");
  fp.groups = gc.groups; fp.nw = gc.nw; fp.smem_w_floats = gc.smem_w_floats;
  cudaError_t e = fma_launch(fp, gc.R, p->desc.method, gc.smem, stream);
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "fma kernel launch (R=%d nw=%d groups=%d smem=%zu): %s", gc.R, gc.nw, gc.groups, gc.smem, cudaGetErrorString(e));
  p->launches += 1;
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

// ---- backward -----------------------------------------------------------------------------------
struct BwdLayout {
  size_t aux = 0, aux_g = 0, xbuf = 0, gvtab = 0, vtab = 0, pstate_f = 0, pstate_g = 0, total = 0;     // float offsets
  size_t dbuf[kMaxOps], pbuf[kMaxOps];
  int n_rops = 0; int rop[kMaxOps];
  bool has_lipswish = false, has_g_ops = false;
};
// srk: every per-row buffer carries one block of S*B rows per evaluation site of the op's part (drift: 3, diffusion: 4)
static BwdLayout bwd_layout(const Program& pg, int B, int S, bool srk) {
  BwdLayout L;
  const size_t SB = (size_t)S * B;
  const size_t nf = srk ? 3 : 1, ng = srk ? 4 : 1, nv = srk ? kSrkGPoints : 1;
  auto take = [&](size_t n) { const size_t o = L.total; L.total += (n + 3) & ~(size_t)3; return o; };
  L.aux = take(nf * SB * 3);
  L.xbuf = pg.uses_control ? take(nf * SB * pg.C) : 0;
  L.vtab = take((size_t)S * nv * pg.H);
  L.gvtab = take((size_t)S * nv * pg.H);
  bool consumed[kMaxOps] = {false};
  for (int o = 0; o < pg.n_ops; ++o) {
    const DenseOp& op = pg.ops[o];
    if (op.vec) continue;
    if (op.part == 1) L.has_g_ops = true;
    if (op.src_op >= 0) consumed[op.src_op] = true;
    if (op.src2 >= 0 && op.src2_op >= 0) consumed[op.src2_op] = true;
    if (op.act == ACT_LIPSWISH) L.has_lipswish = true;
  }
  if (srk) {
    L.aux_g = take(ng * SB * 3);
    L.pstate_f = take(nf * SB * pg.H);
    if (L.has_g_ops) L.pstate_g = take(ng * SB * pg.H);
  }
  for (int o = 0; o < pg.n_ops; ++o) {
    const DenseOp& op = pg.ops[o];
    L.dbuf[o] = L.pbuf[o] = (size_t)-1;
    if (op.vec) continue;
    const size_t ns = op.part == 1 ? ng : nf;
    L.rop[L.n_rops++] = o;
    L.dbuf[o] = take(ns * SB * op.N);
    if (consumed[o]) L.pbuf[o] = take(ns * SB * op.N);
  }
  return L;
}

int64_t snsde_backward_workspace_bytes(const snsde_plan* p, int32_t B, int32_t S) {
  SNSDE_API_BEGIN
  if (!p || !p->has_weights) return fail(SNSDE_ERR_NO_WEIGHTS, "plan has no weights");
  if (B < 1 || S < 0) return fail(SNSDE_ERR_BAD_ARG, "bad sizes B=%d S=%d", B, S);
  return (int64_t)(bwd_layout(p->prog, B, S, p->desc.method == SNSDE_METHOD_SRK).total * sizeof(float));
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

#define CUBLAS_TRY(expr)                                                                        \
  do {                                                                                          \
    cublasStatus_t s__ = (expr);                                                                \
    if (s__ != CUBLAS_STATUS_SUCCESS) return fail(SNSDE_ERR_CUDA, "%s: cuBLAS status %d", #expr, (int)s__); \
  } while (0)

int snsde_backward(snsde_plan* p, const float* coeffs_dev, int64_t coeff_row_stride, int32_t n_knots, int32_t B,
                   const snsde_step* steps_host, int32_t S, const snsde_point* points_host,
                   const float* states_dev, const float* grad_states_dev,
                   const float* dW_dev, const float* dU_dev, uint64_t seed, uint64_t row_offset,
                   float* grad_y0_dev, float* grad_blob_dev, void* workspace_dev, int64_t workspace_bytes,
                   void* stream_v) {
  SNSDE_API_BEGIN
  if (!p) return fail(SNSDE_ERR_BAD_ARG, "plan is NULL");
  if (!p->has_weights) return fail(SNSDE_ERR_NO_WEIGHTS, "snsde_backward before snsde_plan_set_weights");
  const bool srk = p->desc.method == SNSDE_METHOD_SRK;
  if (srk && S && !points_host) return fail(SNSDE_ERR_BAD_ARG, "method srk needs the per-step evaluation points");
  if (srk && dW_dev && !dU_dev) return fail(SNSDE_ERR_BAD_ARG, "method srk with explicit increments needs dU beside dW");
  if (p->desc.method == SNSDE_METHOD_MILSTEIN && p->prog.tail.vjp_kind != 0)
    return fail(SNSDE_ERR_UNSUPPORTED, "backward of 'milstein' through a state-dependent noise network (noise options 14, 15, 18, 19) needs second derivatives of the network: not implemented");
  if (!states_dev || !grad_states_dev || !grad_y0_dev || !grad_blob_dev) return fail(SNSDE_ERR_BAD_ARG, "states/grad_states/grad_y0/grad_blob is NULL");
  if (B < 1 || S < 0 || (S && !steps_host)) return fail(SNSDE_ERR_BAD_ARG, "bad sizes B=%d S=%d", B, S);
  const Program& pg = p->prog;
  int rc = validate_control(pg, coeffs_dev, coeff_row_stride, n_knots);
  if (rc != SNSDE_OK) return rc;
  rc = validate_steps(pg, n_knots, steps_host, S);
  if (rc != SNSDE_OK) return rc;
  if (srk && pg.uses_control)
    for (int i = 0; i < S * kSrkPoints; ++i)
      if (points_host[i].interval < 0 || points_host[i].interval > n_knots - 2)
        return fail(SNSDE_ERR_BAD_ARG, "point %d: spline interval %d outside [0,%d]", i, points_host[i].interval, n_knots - 2);
  const BwdLayout L = bwd_layout(pg, B, S, srk);
  if (workspace_bytes < (int64_t)(L.total * sizeof(float)) || (L.total && !workspace_dev))
    return fail(SNSDE_ERR_BAD_ARG, "workspace has %lld bytes, snsde_backward_workspace_bytes says %lld", (long long)workspace_bytes, (long long)(L.total * sizeof(float)));
  if ((uintptr_t)workspace_dev & 15) return fail(SNSDE_ERR_BAD_ARG, "workspace must be 16-byte aligned");

  cudaStream_t stream = (cudaStream_t)stream_v;
  DeviceGuard guard(p->device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
  PlanStreamGuard order(p, stream_v);
  const int64_t n_w = weight_count(&p->desc);
  CUDA_TRY(cudaMemsetAsync(grad_blob_dev, 0, sizeof(float) * (size_t)n_w, stream));
  if (S == 0) {                                       // no step: y_0 is the only state
    CUDA_TRY(cudaMemcpyAsync(grad_y0_dev, grad_states_dev, sizeof(float) * (size_t)B * pg.H, cudaMemcpyDeviceToDevice, stream));
    return SNSDE_OK;
  }
  rc = upload_tables(p, steps_host, S, nullptr, 0, srk ? points_host : nullptr, stream);
  if (rc != SNSDE_OK) return rc;
  float* ws = (float*)workspace_dev;
  const int nf = srk ? 3 : 1, ng = srk ? 4 : 1, nv = srk ? kSrkGPoints : 1;

  BwdParams bp;
  memset(&bp, 0, sizeof(bp));
  bp.prog = pg;
  bp.wimg = p->d_wimg; bp.wimg_floats = p->wimg_floats; bp.blob = p->d_blob;
  bp.coeffs = coeffs_dev; bp.coeff_row_stride = coeff_row_stride;
  bp.B = B; bp.S = S; bp.steps = p->d_steps;
  bp.states = states_dev; bp.grad_states = grad_states_dev; bp.dW = dW_dev; bp.dU = dU_dev; bp.seed = seed; bp.row_offset = row_offset;
  bp.points = srk ? p->d_points : nullptr;
  bp.pstate_f = srk ? ws + L.pstate_f : nullptr;
  bp.pstate_g = (srk && L.has_g_ops) ? ws + L.pstate_g : nullptr;
  bp.grad_y0 = grad_y0_dev; bp.grad_blob = grad_blob_dev;
  bp.xbuf = pg.uses_control ? ws + L.xbuf : nullptr;
  bp.n_rops = L.n_rops; bp.has_lipswish = L.has_lipswish;
  for (int i = 0; i < L.n_rops; ++i) bp.rop[i] = L.rop[i];
  for (int o = 0; o < pg.n_ops; ++o) {
    bp.dbuf[o] = L.dbuf[o] == (size_t)-1 ? nullptr : ws + L.dbuf[o];
    bp.pbuf[o] = L.pbuf[o] == (size_t)-1 ? nullptr : ws + L.pbuf[o];
  }
  const bool vbuf = pg.tail.coef_src == CO_VBUF;
  if (vbuf) {
    cudaError_t e = vec_tables_launch(pg, p->d_wimg, p->d_steps, bp.points, S, nv, ws + L.vtab, stream);
    if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "coefficient table kernel launch: %s", cudaGetErrorString(e));
    CUDA_TRY(cudaMemsetAsync(ws + L.gvtab, 0, sizeof(float) * (size_t)S * nv * pg.H, stream));
    bp.vtab = ws + L.vtab; bp.gvtab = ws + L.gvtab;
    p->launches += 1;
  }
  GroupConfig gc;
  auto group_floats = [&](int R) {
    return srk ? bwd_srk_group_smem_floats(pg, L.n_rops, R, L.has_lipswish) : bwd_group_smem_floats(pg, L.n_rops, R, L.has_lipswish);
  };
  if (!pick_group_config(p, B, group_floats, false, gc))   // 8 rows per group measured slower (13.5 vs 10.8 ms at c2)
    return fail(SNSDE_ERR_UNSUPPORTED, "the backward pass keeps every activation of a step in shared memory: hidden size too large (%zu bytes per row group)",
                group_floats(1) * sizeof(float));
  bp.groups = gc.groups; bp.nw = gc.nw; bp.smem_w_floats = gc.smem_w_floats;
  cudaError_t e = srk ? bwd_srk_fill_aux(p->d_points, S, B, ws + L.aux, ws + L.aux_g, stream) : bwd_fill_aux(p->d_steps, S, B, ws + L.aux, stream);
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "aux kernel launch: %s", cudaGetErrorString(e));
  e = srk ? bwd_srk_launch(bp, gc.R, gc.smem, stream) : bwd_launch(bp, gc.R, gc.smem, stream);
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "reverse-sweep kernel launch (nw=%d groups=%d smem=%zu): %s", gc.nw, gc.groups, gc.smem, cudaGetErrorString(e));
  p->launches += 2;
  if (vbuf) {
    e = vec_bwd_launch(pg, p->d_wimg, p->d_blob, p->d_steps, bp.points, S, nv, ws + L.gvtab, grad_blob_dev, stream);
    if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "coefficient-network backward launch: %s", cudaGetErrorString(e));
    p->launches += 1;
  }

  // ---- weight gradients: dW[N][K] = D^T P as library GEMMs (row-major operands seen column-major) ----
  if (!p->cublas) CUBLAS_TRY(cublasCreate(&p->cublas));
  CUBLAS_TRY(cublasSetStream(p->cublas, stream));
  CUBLAS_TRY(cublasSetMathMode(p->cublas, CUBLAS_PEDANTIC_MATH));                  // fp32 accumulate, no TF32
  const int SB = S * B;
  const float one = 1.f;
  // C'[K x N] (ldc = ldw: the [N][ldw] blob rows) += P'[K x M] (lda) * D'[N x M]^T (ldb = N); M = rows of all the op's sites
  auto gemm = [&](int M, const float* P, int lda, int K, const float* D, int N, float* C, int ldc) -> cublasStatus_t {
    return cublasSgemm(p->cublas, CUBLAS_OP_N, CUBLAS_OP_T, K, N, M, &one, P, lda, D, N, &one, C, ldc);
  };
  for (int i = 0; i < L.n_rops; ++i) {
    const int o = L.rop[i];
    const DenseOp& op = pg.ops[o];
    const float* D = bp.dbuf[o];
    float* gW = grad_blob_dev + op.g_w;
    const bool gpart = op.part == 1;
    const int M = (gpart ? ng : nf) * SB;
    const float* aux = ws + ((srk && gpart) ? L.aux_g : L.aux);
    for (int part = 0; part < 2; ++part) {
      const int src = part == 0 ? op.src : op.src2, src_op = part == 0 ? op.src_op : op.src2_op;
      const int K = part == 0 ? op.K : op.K2, col = part == 0 ? op.g_col : op.g_col2;
      if (src < 0) continue;
      const float* P; int lda;
      if (src_op == SRC_STATE) { P = srk ? (gpart ? bp.pstate_g : bp.pstate_f) : states_dev; lda = pg.H; }
      else if (src_op == SRC_CONTROL) { P = bp.xbuf; lda = pg.C; }
      else { P = bp.pbuf[src_op]; lda = pg.ops[src_op].N; }
      CUBLAS_TRY(gemm(M, P, lda, K, D, op.N, gW + col, op.g_ldw));
    }
    if (op.tmode == TM_SINCOS) CUBLAS_TRY(gemm(M, aux, 3, 2, D, op.N, gW, op.g_ldw));
    if (op.g_b >= 0) CUBLAS_TRY(gemm(M, aux + 2, 3, 1, D, op.N, grad_blob_dev + op.g_b, 1));
  }
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

int snsde_philox_fill(uint64_t seed, uint64_t row_offset, int32_t S, int32_t B, int32_t H,
                      const snsde_step* steps_host, float* dW_dev, float* dU_dev, int device, void* stream_v) {
  SNSDE_API_BEGIN
  if (S < 0 || B < 1 || H < 1 || (!dW_dev && !dU_dev) || (S && !steps_host)) return fail(SNSDE_ERR_BAD_ARG, "bad philox_fill arguments");
  if (S == 0) return SNSDE_OK;
  cudaStream_t stream = (cudaStream_t)stream_v;
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return fail(SNSDE_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.err));
  snsde_step* d_st = nullptr;
  CUDA_TRY(cudaMallocAsync(&d_st, sizeof(snsde_step) * S, stream));
  cudaError_t e = cudaMemcpyAsync(d_st, steps_host, sizeof(snsde_step) * S, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) {
    const size_t n = (size_t)S * B * H;
    const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    philox_fill_kernel<<<grid, 256, 0, stream>>>(seed, row_offset, S, B, H, d_st, dW_dev, dU_dev);
    e = cudaGetLastError();
  }
  cudaFreeAsync(d_st, stream);
  if (e != cudaSuccess) return fail(SNSDE_ERR_CUDA, "philox_fill: %s", cudaGetErrorString(e));
  return SNSDE_OK;
  SNSDE_API_END(SNSDE_ERR_INTERNAL)
}

}  // extern "C"
