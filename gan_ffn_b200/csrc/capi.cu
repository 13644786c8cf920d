// C ABI (include/ganffn.h) over the kernel files, plus the library state and the GEMM engine
// dispatcher.
#include <stdarg.h>
#include <algorithm>
#include <utility>
#include <vector>

#include "kernels.h"

namespace ganffn {

unsigned long long g_launches = 0;
int g_gemm_engine = GANFFN_GEMM_AUTO;
int g_side_streams = 1;
int g_deterministic = 0;
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- GEMM engine dispatch -------------------------------------------------------------------------------
// AUTO: tcgen05 3xTF32 tiles when the problem has tensor-sized extents, else the FFMA engine.
static bool use_tc(bool transA, bool b_is_nk, int lda, int ldb, int ldc, int M, int N, int K, const void* A,
                   const void* B) {
  if (g_gemm_engine == GANFFN_GEMM_SIMT) return false;
  return gemm_tc_supported(transA, b_is_nk, lda, ldb, ldc, M, N, K, A, B);
}

// ---- per-GEMM CUDA-event profiling (bench.py roofline leg) ------------------------------------------------
// When enabled, every GEMM (including its split-K fold) is bracketed by two events on the launching
// stream; ganffn_gemm_profile_collect() synchronises and sums elapsed time and algorithmic FLOPs per engine.
struct ProfEntry { cudaEvent_t a, b; double flops; int engine; int M, N, K, transA, b_is_nk; };
static bool g_prof = false;
static std::vector<ProfEntry> g_prof_entries;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;

// Inside a stream capture the brackets become *external* event-record nodes: the replayed graph then timestamps every
// GEMM without any host launch gap between the two records (bench.py times its roofline leg this way).
static void prof_record(cudaEvent_t ev, cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (cs == cudaStreamCaptureStatusActive) cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal);
  else cudaEventRecord(ev, st);
}

static ProfEntry prof_begin(double flops, int engine, cudaStream_t st) {
  ProfEntry e;
  if (!g_prof_pool.empty()) {
    e.a = g_prof_pool.back().first; e.b = g_prof_pool.back().second;
    g_prof_pool.pop_back();
  } else {
    cudaEventCreate(&e.a); cudaEventCreate(&e.b);
  }
  e.flops = flops; e.engine = engine;
  prof_record(e.a, st);
  return e;
}

int gemm(const float* A, int lda, bool transA, const float* B, int ldb, bool b_is_nk, float* C, int ldc, int M, int N,
         int K, const Epilogue& ep, float* scratch, int64_t scratch_floats, cudaStream_t st) {
  const bool tc = use_tc(transA, b_is_nk, lda, ldb, ldc, M, N, K, A, B);
  ProfEntry pe{};
  if (g_prof) {
    pe = prof_begin(2.0 * M * N * K, tc ? GANFFN_GEMM_TC : GANFFN_GEMM_SIMT, st);
    pe.M = M; pe.N = N; pe.K = K; pe.transA = transA; pe.b_is_nk = b_is_nk;
  }
  int rc;
  if (tc) {
    rc = gemm_tc(A, lda, transA, B, ldb, b_is_nk, C, ldc, M, N, K, ep, scratch, scratch_floats, st);
  } else {
    Epilogue e2 = ep;
    e2.rowsum = nullptr;
    e2.ln_out = nullptr;
    e2.lnb_dz = nullptr;
    rc = gemm_simt(A, lda, transA, B, ldb, b_is_nk, C, ldc, M, N, K, e2, scratch, scratch_floats, st);
    if (rc == GANFFN_OK && ep.ln_out) rc = layernorm_fwd(C, ep.ln_gamma, ep.ln_beta, ep.ln_out, M, N, st);   // not fused on the FFMA engine
    if (rc == GANFFN_OK && ep.lnb_dz) rc = layernorm_bwd_after(ep, C, M, N, st);
  }
  if (g_prof) {
    prof_record(pe.b, st);
    g_prof_entries.push_back(pe);
  }
  return rc;
}

int64_t gemm_scratch_floats(int M, int N, int K) {
  return std::max(gemm_simt_scratch_floats(M, N, K), gemm_tc_scratch_floats(M, N, K));
}

int linear_wgrad(const float* dy, const float* x, float* dw, float* db, int M, int N, int K, int accumulate,
                 float* gemm_scratch, int64_t gemm_scratch_n, cudaStream_t st) {
  if (!accumulate) {
    cudaMemsetAsync(dw, 0, (size_t)N * K * sizeof(float), st);
    if (db) cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st);
  }
  if (g_deterministic) {
    // Fixed-order accumulation: dw = 1 * dw + dy^T x through the non-atomic path (split-K partials folded in slice
    // order by the reduce kernel), bias gradient by the single-writer column sum.  One backward pass at a time
    // may add to an arena in this mode (the host side runs the networks serially).
    Epilogue ep;
    ep.beta = 1.0f;
    GANFFN_TRY(gemm(dy, N, true, x, K, false, dw, K, N, K, M, ep, gemm_scratch, gemm_scratch_n, st));
    if (db) GANFFN_TRY(colsum(dy, M, N, db, 1, st));
    return GANFFN_OK;
  }
  Epilogue ep;
  ep.atomic_acc = 1;
  const bool tc = gemm_uses_tc(dy, N, true, x, K, false, K, N, K, M);
  if (tc) ep.rowsum = db;   // bias gradient from the A producers' registers
  GANFFN_TRY(gemm(dy, N, true, x, K, false, dw, K, N, K, M, ep, gemm_scratch, gemm_scratch_n, st));
  if (db && !tc) GANFFN_TRY(colsum(dy, M, N, db, 1, st));
  return GANFFN_OK;
}

bool gemm_uses_tc(const float* A, int lda, bool transA, const float* B, int ldb, bool b_is_nk, int ldc, int M, int N,
                  int K) {
  return use_tc(transA, b_is_nk, lda, ldb, ldc, M, N, K, A, B);
}

}  // namespace ganffn

using namespace ganffn;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int ganffn_version(void) { return 100; }
const char* ganffn_last_error(void) { return g_err; }
unsigned long long ganffn_launch_count(void) { return g_launches; }
void ganffn_reset_launch_count(void) { g_launches = 0; }
void ganffn_gemm_profile_enable(int on) { g_prof = on != 0; }
int ganffn_set_deterministic(int on) { const int prev = g_deterministic; g_deterministic = on != 0; return prev; }
int ganffn_set_side_streams(int on) { const int prev = g_side_streams; g_side_streams = on != 0; return prev; }
int ganffn_gemm_profile_collect(int engine, double* total_ms, double* total_flops, int64_t* launches) {
  GANFFN_CHECK_ARG(total_ms && total_flops && launches, "gemm_profile_collect: null pointer");
  double ms = 0.0, fl = 0.0;
  int64_t n = 0;
  for (auto& e : g_prof_entries) {
    if (cudaEventSynchronize(e.b) != cudaSuccess) { set_error("gemm_profile_collect: event sync failed"); return GANFFN_ERR_CUDA; }
    float t = 0.f;
    cudaEventElapsedTime(&t, e.a, e.b);
    if (engine == GANFFN_GEMM_AUTO || engine == e.engine) { ms += t; fl += e.flops; ++n; }
    g_prof_pool.emplace_back(e.a, e.b);
  }
  g_prof_entries.clear();
  *total_ms = ms; *total_flops = fl; *launches = n;
  return GANFFN_OK;
}
int64_t ganffn_gemm_profile_table(int64_t* shapes, double* ms, int64_t max_rows) {
  if (!shapes || !ms || max_rows <= 0) { set_error("gemm_profile_table: null pointer"); return -1; }
  int64_t rows = 0;
  for (auto& e : g_prof_entries) {
    if (cudaEventSynchronize(e.b) != cudaSuccess) { set_error("gemm_profile_table: event sync failed"); return -1; }
    float t = 0.f;
    if (cudaEventElapsedTime(&t, e.a, e.b) != cudaSuccess) { set_error("gemm_profile_table: elapsed time failed"); return -1; }
    int64_t r = 0;
    for (; r < rows; ++r)
      if (shapes[6 * r] == e.M && shapes[6 * r + 1] == e.N && shapes[6 * r + 2] == e.K && shapes[6 * r + 3] == e.transA &&
          shapes[6 * r + 4] == e.b_is_nk && shapes[6 * r + 5] == e.engine)
        break;
    if (r == rows) {
      if (rows == max_rows) continue;
      shapes[6 * r] = e.M; shapes[6 * r + 1] = e.N; shapes[6 * r + 2] = e.K; shapes[6 * r + 3] = e.transA;
      shapes[6 * r + 4] = e.b_is_nk; shapes[6 * r + 5] = e.engine;
      ms[2 * r] = 0.0; ms[2 * r + 1] = 0.0;
      ++rows;
    }
    ms[2 * r] += t; ms[2 * r + 1] += 1.0;
  }
  return rows;
}
int ganffn_set_gemm_engine(int engine) {
  int prev = g_gemm_engine;
  if (engine >= GANFFN_GEMM_AUTO && engine <= GANFFN_GEMM_TF32X1) g_gemm_engine = engine;
  return prev;
}

int ganffn_linear_fwd(const float* x, const float* w, const float* bias, const float* residual, float* y, float* pre,
                      int M, int N, int K, int act, int drop_before_act, float p_drop, uint64_t seed, int site,
                      float* scratch, int64_t scratch_floats, void* stream) {
  GANFFN_CHECK_ARG(x && w && y, "linear_fwd: null pointer");
  GANFFN_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "linear_fwd: dropout p=%f", p_drop);
  Epilogue ep;
  ep.bias = bias; ep.residual = residual; ep.ldr = N; ep.pre = pre; ep.act = act; ep.drop_before_act = drop_before_act;
  ep.p_drop = p_drop; ep.seed = seed; ep.site = (uint32_t)site;
  return gemm(x, K, false, w, K, true, y, N, M, N, K, ep, scratch, scratch_floats, S(stream));
}

int ganffn_linear_ln_fwd(const float* x, const float* w, const float* bias, const float* residual, const float* gamma,
                         const float* beta, float* z, float* y, int M, int N, int K, float p_drop, uint64_t seed, int site,
                         float* scratch, int64_t scratch_floats, void* stream) {
  GANFFN_CHECK_ARG(x && w && bias && residual && gamma && beta && z && y, "linear_ln_fwd: null pointer");
  GANFFN_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "linear_ln_fwd: dropout p=%f", p_drop);
  GANFFN_CHECK_ARG(N % 4 == 0 && N <= 512, "linear_ln_fwd: N=%d must be a multiple of 4 and <= 512 (LayerNorm width)", N);
  Epilogue ep;
  ep.bias = bias; ep.residual = residual; ep.ldr = N; ep.p_drop = p_drop; ep.seed = seed; ep.site = (uint32_t)site;
  ep.ln_gamma = gamma; ep.ln_beta = beta; ep.ln_out = y;
  return gemm(x, K, false, w, K, true, z, N, M, N, K, ep, scratch, scratch_floats, S(stream));
}

int ganffn_disc_head_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                         const float* b3, float* g0, float* f1, float* a1, float* f2, float* a2, float* prob, int T, int d,
                         float p_drop, uint64_t seed, int site0, void* stream) {
  GANFFN_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f && T >= 1, "disc_head_fwd: T=%d p=%f", T, p_drop);
  return disc_head_fwd(x, w1, b1, w2, b2, w3, b3, g0, f1, a1, f2, a2, prob, T, d, p_drop, Seed(seed), site0, S(stream));
}

int ganffn_disc_head_bwd(const float* d_prob, const float* prob, const float* x, const float* g0, const float* f1, const float* a1,
                         const float* f2, const float* a2, const float* w1, const float* w2, const float* w3, float* dx,
                         float* dw1, float* db1, float* dw2, float* db2, float* dw3, float* db3, int T, int d, float p_drop,
                         uint64_t seed, int site0, void* stream) {
  GANFFN_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f && T >= 1, "disc_head_bwd: T=%d p=%f", T, p_drop);
  return disc_head_bwd(d_prob, prob, x, g0, f1, a1, f2, a2, w1, w2, w3, dx, dw1, db1, dw2, db2, dw3, db3, T, d, p_drop, Seed(seed),
                       site0, S(stream));
}

int ganffn_linear_dgrad(const float* dy, const float* w, const float* residual, float* dx, int M, int N, int K,
                        float* scratch, int64_t scratch_floats, void* stream) {
  GANFFN_CHECK_ARG(dy && w && dx, "linear_dgrad: null pointer");
  Epilogue ep;
  ep.residual = residual; ep.ldr = K;
  return gemm(dy, N, false, w, K, false, dx, K, M, K, N, ep, scratch, scratch_floats, S(stream));
}

int64_t ganffn_wgrad_scratch_floats(int M, int N, int K) {
  (void)M;
  return round_up(gemm_scratch_floats(N, K, M), 32);
}

int64_t ganffn_gemm_scratch_floats(int M, int N, int K) { return gemm_scratch_floats(M, N, K); }

int ganffn_linear_wgrad(const float* dy, const float* x, float* dw, float* db, int M, int N, int K, int accumulate,
                        float* scratch, void* stream) {
  GANFFN_CHECK_ARG(dy && x && dw && scratch, "linear_wgrad: null pointer");
  return linear_wgrad(dy, x, dw, db, M, N, K, accumulate, scratch, round_up(gemm_scratch_floats(N, K, M), 32), S(stream));
}

int ganffn_attention_fwd(const float* qkv, float* o, float* lse, int S_, int B, int d, int nhead, float p_drop,
                         uint64_t seed, int site, void* stream) {
  GANFFN_CHECK_ARG(qkv && o && lse, "attention_fwd: null pointer");
  return attention_fwd(qkv, o, lse, S_, B, d, nhead, p_drop, seed, site, S(stream));
}

int ganffn_attention_bwd(const float* qkv, const float* o, const float* lse, const float* d_o, float* dqkv, int S_,
                         int B, int d, int nhead, float p_drop, uint64_t seed, int site, void* stream) {
  GANFFN_CHECK_ARG(qkv && o && lse && d_o && dqkv, "attention_bwd: null pointer");
  return attention_bwd(qkv, o, lse, d_o, dqkv, S_, B, d, nhead, p_drop, seed, site, S(stream));
}

int ganffn_layernorm_fwd(const float* z, const float* gamma, const float* beta, float* y, int T, int d, void* stream) {
  GANFFN_CHECK_ARG(z && gamma && beta && y, "layernorm_fwd: null pointer");
  return layernorm_fwd(z, gamma, beta, y, T, d, S(stream));
}

int ganffn_layernorm_bwd(const float* dy, const float* z, const float* gamma, float* dz, float* dz_drop, float* dgamma,
                         float* dbeta, int T, int d, int accumulate, float p_drop, uint64_t seed, int site,
                         float* scratch, void* stream) {
  GANFFN_CHECK_ARG(dy && z && gamma && dz && dgamma && dbeta, "layernorm_bwd: null pointer");
  (void)scratch;
  return layernorm_bwd(dy, z, gamma, dz, dz_drop, dgamma, dbeta, nullptr, T, d, accumulate, p_drop, seed, site, S(stream));
}

int64_t ganffn_layernorm_scratch_floats(int T, int d) { (void)T; (void)d; return 0; }

int ganffn_posenc_fwd(const float* x, const float* pe, float* y, int S_, int B, int d, float p_drop, uint64_t seed,
                      void* stream) {
  GANFFN_CHECK_ARG(x && pe && y, "posenc_fwd: null pointer");
  return posenc_fwd(x, pe, y, S_, B, d, p_drop, seed, S(stream));
}

int ganffn_dropout_mask(float* out, int64_t rows, int64_t cols, int64_t row_stride, float p_drop, uint64_t seed,
                        int site, void* stream) {
  GANFFN_CHECK_ARG(out, "dropout_mask: null pointer");
  return dropout_mask(out, rows, cols, row_stride, p_drop, seed, site, S(stream));
}

int ganffn_fuse_cls_fwd(const float* a, const float* v, const float* t, const float* w, const float* b, float* fusion,
                        float* log_prob, int T, int dh, int C, void* stream) {
  GANFFN_CHECK_ARG(a && v && t && w && b && log_prob, "fuse_cls_fwd: null pointer");
  return fuse_cls_fwd(a, v, t, w, b, fusion, log_prob, T, dh, C, S(stream));
}

int ganffn_fuse_cls_bwd(const float* d_log_prob, const float* log_prob, const float* fusion, const float* w,
                        float* d_fusion, float* dw, float* db, int T, int dh, int C, int accumulate, float* scratch,
                        void* stream) {
  GANFFN_CHECK_ARG(d_log_prob && log_prob && fusion && w && d_fusion && dw && db, "fuse_cls_bwd: null pointer");
  return fuse_cls_bwd(d_log_prob, log_prob, fusion, w, d_fusion, dw, db, T, dh, C, accumulate, scratch, S(stream));
}

int64_t ganffn_fuse_cls_scratch_floats(int T, int dh, int C) { return fuse_cls_scratch_floats(T, dh, C); }

int ganffn_masked_nll_fwd(const float* pred, const int64_t* target, const float* mask, const float* weight,
                          float* loss_and_den, int64_t n, int C, float den_override, void* stream) {
  GANFFN_CHECK_ARG(pred && target && mask && loss_and_den, "masked_nll_fwd: null pointer");
  return masked_nll_fwd(pred, target, mask, weight, loss_and_den, n, C, den_override, S(stream));
}

int ganffn_masked_nll_bwd(const float* d_loss, const float* loss_and_den, const int64_t* target, const float* mask,
                          const float* weight, float* d_pred, int64_t n, int C, void* stream) {
  GANFFN_CHECK_ARG(d_loss && loss_and_den && target && mask && d_pred, "masked_nll_bwd: null pointer");
  return masked_nll_bwd(d_loss, loss_and_den, target, mask, weight, d_pred, n, C, S(stream));
}

int ganffn_bce_fwd(const float* prob, const float* target, float* loss, int64_t n, float scale, void* stream) {
  GANFFN_CHECK_ARG(prob && target && loss, "bce_fwd: null pointer");
  return bce_fwd(prob, target, loss, n, scale, S(stream));
}

int ganffn_bce_bwd(const float* d_loss, const float* prob, const float* target, float* d_prob, int64_t n, float scale,
                   void* stream) {
  GANFFN_CHECK_ARG(d_loss && prob && target && d_prob, "bce_bwd: null pointer");
  return bce_bwd(d_loss, prob, target, d_prob, n, scale, S(stream));
}

int ganffn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int step, float lr, float beta1,
                     float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
  GANFFN_CHECK_ARG(p && g && m && v, "adam_step: null pointer");
  return adam_step(p, g, m, v, n, step, lr, beta1, beta2, eps, weight_decay, grad_scale, S(stream));
}

int ganffn_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const int* step_dev, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
  GANFFN_CHECK_ARG(p && g && m && v && step_dev, "adam_step_dev: null pointer");
  return adam_step_dev(p, g, m, v, n, step_dev, lr, beta1, beta2, eps, weight_decay, grad_scale, S(stream));
}

static NetDims dims(int kind, int S_, int B, int d_in, int d, int nhead, int dff, int nlayers, int h1, int h2) {
  NetDims nd;
  nd.kind = kind; nd.S = S_; nd.B = B; nd.d_in = d_in; nd.d = d; nd.nhead = nhead; nd.dff = dff; nd.L = nlayers;
  nd.h1 = h1; nd.h2 = h2;
  return nd;
}

int ganffn_net_fwd(int kind, const float* params, const int64_t* off, const float* pe, const float* x, float* out,
                   float* stash, float* scratch, int S_, int B, int d_in, int d, int nhead, int dff, int nlayers, int h1,
                   int h2, int train, float p_head, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  return net_fwd(dims(kind, S_, B, d_in, d, nhead, dff, nlayers, h1, h2), params, off, pe, x, out, stash, scratch, train,
                 p_head, Seed(seed, seed_dev), S(stream));
}

int ganffn_net_bwd(int kind, const float* params, const int64_t* off, const float* x, const float* out,
                   const float* d_out_grad, const float* stash, float* grads, float* dx, float* scratch, int S_, int B,
                   int d_in, int d, int nhead, int dff, int nlayers, int h1, int h2, int train, float p_head,
                   uint64_t seed, const uint64_t* seed_dev, int accumulate, void* stream) {
  return net_bwd(dims(kind, S_, B, d_in, d, nhead, dff, nlayers, h1, h2), params, off, x, out, d_out_grad, stash, grads,
                 dx, scratch, train, p_head, Seed(seed, seed_dev), accumulate, S(stream));
}

int ganffn_net_bwd_layer_wait(void* bwd_stream, int layer, void* waiting_stream) {
  return net_bwd_layer_wait(S(bwd_stream), layer, S(waiting_stream));
}

int64_t ganffn_net_stash_floats(int kind, int S_, int B, int d_in, int d, int nhead, int dff, int nlayers, int h1,
                                int h2) {
  NetDims nd = dims(kind, S_, B, d_in, d, nhead, dff, nlayers, h1, h2);
  if (net_check(nd) != GANFFN_OK) return -1;
  return net_stash_floats(nd);
}

int64_t ganffn_net_scratch_floats(int kind, int S_, int B, int d_in, int d, int nhead, int dff, int nlayers, int h1,
                                  int h2) {
  NetDims nd = dims(kind, S_, B, d_in, d, nhead, dff, nlayers, h1, h2);
  if (net_check(nd) != GANFFN_OK) return -1;
  return net_scratch_floats(nd);
}

}  // extern "C"
