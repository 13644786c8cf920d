// Whole-network forward / backward: [object] -> PE -> L encoder layers -> head, as one host-side
// sequence of kernel launches on the caller's stream (no allocation, no synchronisation, so the
// whole sequence can be captured into a CUDA graph by the caller).
//
// Reference: generators model.py:1221-1231, 1255-1263, 1286-1294; discriminators model.py:1320-1327,
// 1354-1364, 1390-1397; encoder layer torch/nn/modules/transformer.py:944-982 (post-norm, ReLU).
#include <map>
#include <mutex>
#include <stdlib.h>

#include "kernels.h"

namespace ganffn {

namespace {

constexpr float P_PE = 0.2f;   // model.py:1179
constexpr float P_ENC = 0.1f;  // torch TransformerEncoderLayer default

inline int64_t al(int64_t n) { return round_up(n, 32); }  // 128-byte aligned regions

// parameter table indices (see include/ganffn.h)
enum { IN_W = 0, IN_B, OUT_W, OUT_B, L1_W, L1_B, L2_W, L2_B, N1_W, N1_B, N2_W, N2_B, PER_LAYER };
enum { FC1_W = 0, FC1_B, FC2_W, FC2_B, FC3_W, FC3_B, OBJ_W, OBJ_B };

struct Stash {
  int64_t obj, x0, layer0, per_layer, qkv, lse, o, z1, x1, h, z2, x2, g0, f1, a1, f2, a2, total;
};

Stash stash_layout(const NetDims& nd) {
  Stash s;
  const int64_t T = nd.T(), d = nd.d;
  int64_t p = 0;
  s.obj = p; if (nd.has_object()) p += al(T * d);
  s.x0 = p; p += al(T * d);
  s.layer0 = p;
  int64_t q = 0;
  s.qkv = q; q += al(T * 3 * d);
  s.lse = q; q += al((int64_t)nd.B * nd.nhead * nd.S);
  s.o = q; q += al(T * d);
  s.z1 = q; q += al(T * d);
  s.x1 = q; q += al(T * d);
  s.h = q; q += al(T * nd.dff);
  s.z2 = q; q += al(T * d);
  s.x2 = q; q += al(T * d);
  s.per_layer = q;
  p += q * nd.L;
  s.g0 = p; p += al(T * d);
  s.f1 = p; p += al(T * nd.h1);
  s.a1 = p; p += al(T * nd.h1);
  s.f2 = p; p += al(T * nd.h2);
  s.a2 = p; if (nd.kind == GANFFN_NET_DISCRIMINATOR) p += al(T * nd.h2);
  s.total = p;
  return s;
}

struct Scratch {
  int64_t da, db, dz, dzd, dz1, dzd1, d_o, dqkv, dh, hb1, hb2, hb3, gemm, gemm_floats, gemm2, red, total;
};

Scratch scratch_layout(const NetDims& nd) {
  Scratch s;
  const int64_t T = nd.T(), d = nd.d;
  int64_t p = 0;
  s.da = p; p += al(T * d);
  s.db = p; p += al(T * d);
  s.dz = p; p += al(T * d);
  s.dzd = p; p += al(T * d);
  s.dz1 = p; p += al(T * d);     // LayerNorm-1 backward writes its own pair, so the weight-gradient stream can still
  s.dzd1 = p; p += al(T * d);    // be reading the LayerNorm-2 pair (and vice versa)
  s.d_o = p; p += al(T * d);
  s.dqkv = p; p += al(T * 3 * d);
  s.dh = p; p += al(T * nd.dff);
  s.hb1 = p; p += al(T * nd.h1);
  s.hb2 = p; p += al(T * nd.h2);
  s.hb3 = p; p += al(T);
  int64_t g = 0;
  auto mx = [&](int M, int N, int K) {
    g = std::max(g, gemm_scratch_floats(M, N, K));
  };
  const int Ti = (int)T;
  // forward + dgrad shapes
  mx(Ti, 3 * nd.d, nd.d); mx(Ti, nd.d, nd.d); mx(Ti, nd.dff, nd.d); mx(Ti, nd.d, nd.dff); mx(Ti, nd.d, 3 * nd.d);
  mx(Ti, nd.h1, nd.d); mx(Ti, nd.h2, nd.h1); mx(Ti, nd.h1, nd.h2); mx(Ti, nd.d, nd.h1); mx(Ti, 1, nd.h2); mx(Ti, nd.h2, 1);
  if (nd.has_object()) { mx(Ti, nd.d, nd.d_in); mx(Ti, nd.d_in, nd.d); mx(nd.d, nd.d_in, Ti); }
  // wgrad shapes
  mx(3 * nd.d, nd.d, Ti); mx(nd.d, nd.d, Ti); mx(nd.dff, nd.d, Ti); mx(nd.d, nd.dff, Ti);
  mx(nd.h1, nd.d, Ti); mx(nd.h2, nd.h1, Ti); mx(1, nd.h2, Ti);
  s.gemm = p; s.gemm_floats = al(g); p += s.gemm_floats;
  s.gemm2 = p; p += s.gemm_floats;   // workspace of the weight-gradient stream
  s.red = p; p += al(32);
  s.total = p;
  return s;
}

struct Ctx {
  const NetDims& nd;
  float* gemm_scratch;
  int64_t gemm_floats;
  float* red;
  cudaStream_t st;

  int linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int N, int K, Epilogue ep) const {
    ep.bias = b;
    return gemm(x, K, false, w, K, true, y, N, M, N, K, ep, gemm_scratch, gemm_floats, st);
  }
  int dgrad(const float* dy, const float* w, float* dx, int M, int N, int K, Epilogue ep) const {
    return gemm(dy, N, false, w, K, false, dx, K, M, K, N, ep, gemm_scratch, gemm_floats, st);
  }
  // gradients always accumulate into the arena (net_bwd zeroes it first when asked not to accumulate)
  int wgrad(const float* dy, const float* x, float* dw, float* dbias, int M, int N, int K, int /*accumulate*/) const {
    return linear_wgrad(dy, x, dw, dbias, M, N, K, 1, gemm_scratch, gemm_floats, st);
  }
};

// ---- weight-gradient side stream -----------------------------------------------------------------------------------
// In the backward pass only the data gradients are on the critical path (LN -> dgrad -> dgrad -> LN -> dgrad ->
// attention -> dgrad per layer); the four weight-gradient products of a layer just have to be finished before the
// optimizer runs.  They are issued on a second stream that forks from / joins the caller's stream with events (legal
// inside CUDA-graph capture), so they fill SMs the 24..144-CTA data-gradient kernels leave idle.  One side stream
// per caller stream (networks on different lanes each get their own); created on first use, never destroyed.
constexpr int MAX_LAYER_EVENTS = 32;
struct SideCtx {
  cudaStream_t side = nullptr;
  cudaEvent_t ready[4], done[4], join;   // slots: L2, L1, OUT, IN
  bool pending[4] = {false, false, false, false};
  // layer_done[l]: every gradient of encoder layer l (weight gradients on the side stream AND the LayerNorm / bias
  // gradients the data-gradient stream wrote) has landed -- what a per-layer gradient all-reduce waits for
  // (ganffn_net_bwd_layer_wait; parallel.py overlaps the all-reduce of layer l with the backward of layers l-1 .. 0)
  cudaEvent_t layer_done[MAX_LAYER_EVENTS], layer_tmp;
  int layers_recorded = 0;
};
enum { W_L2 = 0, W_L1 = 1, W_OUT = 2, W_IN = 3 };

SideCtx* side_ctx(cudaStream_t st) {
  static std::mutex mu;
  static std::map<cudaStream_t, SideCtx*> table;
  static const bool off = getenv("GANFFN_NO_WGRAD_STREAM") != nullptr;
  if (off || !g_side_streams) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  auto it = table.find(st);
  if (it != table.end()) return it->second;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (cs != cudaStreamCaptureStatusNone) return nullptr;   // never create streams/events while capturing
  SideCtx* c = new SideCtx();
  if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess) { delete c; return nullptr; }
  for (int i = 0; i < 4; ++i) {
    cudaEventCreateWithFlags(&c->ready[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&c->join, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->layer_tmp, cudaEventDisableTiming);
  for (int i = 0; i < MAX_LAYER_EVENTS; ++i) cudaEventCreateWithFlags(&c->layer_done[i], cudaEventDisableTiming);
  table[st] = c;
  return c;
}

}  // namespace

// `waiting` waits (device-side) until every gradient of encoder layer `layer` of the LAST ganffn_net_bwd issued on
// `bwd_stream` has landed.  GANFFN_ERR_ARG when that call recorded no per-layer events (side streams off, or created
// during a stream capture): the caller then waits for the whole backward pass instead.
int net_bwd_layer_wait(cudaStream_t bwd_stream, int layer, cudaStream_t waiting) {
  if (!g_side_streams) { set_error("net_bwd_layer_wait: side streams are off"); return GANFFN_ERR_ARG; }
  SideCtx* sx = side_ctx(bwd_stream);
  if (sx == nullptr || layer < 0 || layer >= sx->layers_recorded) {
    set_error("net_bwd_layer_wait: no per-layer event for layer %d on this stream", layer);
    return GANFFN_ERR_ARG;
  }
  if (cudaStreamWaitEvent(waiting, sx->layer_done[layer], 0) != cudaSuccess) {
    set_error("net_bwd_layer_wait: cudaStreamWaitEvent failed");
    return GANFFN_ERR_CUDA;
  }
  return GANFFN_OK;
}

int net_check(const NetDims& nd) {
  GANFFN_CHECK_ARG(nd.kind == GANFFN_NET_GENERATOR || nd.kind == GANFFN_NET_DISCRIMINATOR, "net: kind %d", nd.kind);
  GANFFN_CHECK_ARG(nd.S >= 1 && nd.S <= GANFFN_MAX_SEQ, "net: seq_len %d outside [1,%d] (PositionalEncoding max_len, model.py:1179)",
                   nd.S, GANFFN_MAX_SEQ);
  GANFFN_CHECK_ARG(nd.B >= 1, "net: batch %d", nd.B);
  GANFFN_CHECK_ARG(nd.d > 0 && nd.d % 4 == 0 && nd.d <= 512, "net: d_model %d must be a multiple of 4 and <= 512", nd.d);
  GANFFN_CHECK_ARG(nd.nhead > 0 && nd.d % nd.nhead == 0, "net: d_model %d not divisible by nhead %d", nd.d, nd.nhead);
  GANFFN_CHECK_ARG(nd.dff > 0 && nd.dff % 4 == 0 && nd.L >= 1, "net: dff %d / layers %d", nd.dff, nd.L);
  GANFFN_CHECK_ARG(nd.h1 > 0 && nd.h2 > 0 && nd.h1 % 4 == 0 && nd.h2 % 4 == 0, "net: head widths %d,%d must be multiples of 4", nd.h1, nd.h2);
  GANFFN_CHECK_ARG(nd.d_in == nd.d || (nd.kind == GANFFN_NET_DISCRIMINATOR && nd.d_in % 4 == 0),
                   "net: input width %d does not match d_model %d (only the visual discriminator projects, model.py:1355)",
                   nd.d_in, nd.d);
  return GANFFN_OK;
}

int64_t net_stash_floats(const NetDims& nd) { return stash_layout(nd).total; }
int64_t net_scratch_floats(const NetDims& nd) { return scratch_layout(nd).total; }

int net_fwd(const NetDims& nd, const float* params, const int64_t* off, const float* pe, const float* x, float* out,
            float* stash, float* scratch, int train, float p_head, Seed seed, cudaStream_t st) {
  GANFFN_TRY(net_check(nd));
  GANFFN_CHECK_ARG(params && off && pe && x && out && stash && scratch, "net_fwd: null pointer");
  const Stash sl = stash_layout(nd);
  const Scratch sc = scratch_layout(nd);
  const Ctx cx{nd, scratch + sc.gemm, sc.gemm_floats, scratch + sc.red, st};
  const int T = nd.T(), d = nd.d;
  const float p_pe = train ? P_PE : 0.f, p_enc = train ? P_ENC : 0.f, p_hd = train ? p_head : 0.f;
  const int64_t* hoff = off + (int64_t)nd.L * PER_LAYER;
  auto P = [&](int64_t o) { return params + o; };

  const float* xin = x;
  if (nd.has_object()) {
    GANFFN_CHECK_ARG(hoff[OBJ_W] >= 0, "net_fwd: input is %d wide but the network has no `object` projection", nd.d_in);
    GANFFN_TRY(cx.linear_fwd(x, P(hoff[OBJ_W]), P(hoff[OBJ_B]), stash + sl.obj, T, d, nd.d_in, Epilogue{}));
    xin = stash + sl.obj;
  }
  GANFFN_TRY(posenc_fwd(xin, pe, stash + sl.x0, nd.S, nd.B, d, p_pe, seed, st));

  const float* cur = stash + sl.x0;
  for (int l = 0; l < nd.L; ++l) {
    float* base = stash + sl.layer0 + (int64_t)l * sl.per_layer;
    const int64_t* lo = off + (int64_t)l * PER_LAYER;
    GANFFN_TRY(cx.linear_fwd(cur, P(lo[IN_W]), P(lo[IN_B]), base + sl.qkv, T, 3 * d, d, Epilogue{}));
    GANFFN_TRY(attention_fwd(base + sl.qkv, base + sl.o, base + sl.lse, nd.S, nd.B, d, nd.nhead, p_enc, seed,
                             GANFFN_SITE_LAYER(l, 0), st));
    {
      Epilogue ep; ep.p_drop = p_enc; ep.seed = seed; ep.site = GANFFN_SITE_LAYER(l, 1);
      ep.residual = cur; ep.ldr = d;
      ep.ln_gamma = P(lo[N1_W]); ep.ln_beta = P(lo[N1_B]); ep.ln_out = base + sl.x1;   // x1 = LN1(z1), fused where the engine can
      GANFFN_TRY(cx.linear_fwd(base + sl.o, P(lo[OUT_W]), P(lo[OUT_B]), base + sl.z1, T, d, d, ep));
    }
    {
      Epilogue ep; ep.act = GANFFN_ACT_RELU; ep.p_drop = p_enc; ep.seed = seed; ep.site = GANFFN_SITE_LAYER(l, 2);
      GANFFN_TRY(cx.linear_fwd(base + sl.x1, P(lo[L1_W]), P(lo[L1_B]), base + sl.h, T, nd.dff, d, ep));
    }
    {
      Epilogue ep; ep.p_drop = p_enc; ep.seed = seed; ep.site = GANFFN_SITE_LAYER(l, 3);
      ep.residual = base + sl.x1; ep.ldr = d;
      ep.ln_gamma = P(lo[N2_W]); ep.ln_beta = P(lo[N2_B]); ep.ln_out = base + sl.x2;   // x2 = LN2(z2)
      GANFFN_TRY(cx.linear_fwd(base + sl.h, P(lo[L2_W]), P(lo[L2_B]), base + sl.z2, T, d, nd.dff, ep));
    }
    cur = base + sl.x2;
  }

  // head
  const bool gen = nd.kind == GANFFN_NET_GENERATOR;
  if (!gen && disc_head_fusable(d, nd.h1, nd.h2))   // the whole 100 -> 64 -> 16 -> 1 discriminator head in one kernel (head.cu)
    return disc_head_fwd(cur, P(hoff[FC1_W]), P(hoff[FC1_B]), P(hoff[FC2_W]), P(hoff[FC2_B]), P(hoff[FC3_W]), P(hoff[FC3_B]),
                         stash + sl.g0, stash + sl.f1, stash + sl.a1, stash + sl.f2, stash + sl.a2, out, T, d, p_hd, seed,
                         GANFFN_SITE_HEAD, st);
  GANFFN_TRY(elementwise(cur, nullptr, stash + sl.g0, (int64_t)T * d, EW_GELU_DROP, gen ? p_hd : 0.f, seed,
                         GANFFN_SITE_HEAD + 0, st));
  {
    Epilogue ep; ep.act = GANFFN_ACT_GELU; ep.drop_before_act = 1; ep.p_drop = p_hd; ep.seed = seed;
    ep.site = GANFFN_SITE_HEAD + 1; ep.pre = stash + sl.f1;
    GANFFN_TRY(cx.linear_fwd(stash + sl.g0, P(hoff[FC1_W]), P(hoff[FC1_B]), stash + sl.a1, T, nd.h1, d, ep));
  }
  {
    Epilogue ep; ep.act = GANFFN_ACT_GELU; ep.drop_before_act = 1; ep.p_drop = p_hd; ep.seed = seed;
    ep.site = GANFFN_SITE_HEAD + 2; ep.pre = stash + sl.f2;
    float* dst = gen ? out : stash + sl.a2;
    GANFFN_TRY(cx.linear_fwd(stash + sl.a1, P(hoff[FC2_W]), P(hoff[FC2_B]), dst, T, nd.h2, nd.h1, ep));
  }
  if (!gen) {
    Epilogue ep; ep.act = GANFFN_ACT_SIGMOID; ep.drop_before_act = 1; ep.p_drop = p_hd; ep.seed = seed;
    ep.site = GANFFN_SITE_HEAD + 3;
    GANFFN_TRY(cx.linear_fwd(stash + sl.a2, P(hoff[FC3_W]), P(hoff[FC3_B]), out, T, 1, nd.h2, ep));
  }
  return GANFFN_OK;
}

int net_bwd(const NetDims& nd, const float* params, const int64_t* off, const float* x, const float* out,
            const float* d_out, const float* stash, float* grads, float* dx, float* scratch, int train, float p_head,
            Seed seed, int accumulate, cudaStream_t st) {
  GANFFN_TRY(net_check(nd));
  GANFFN_CHECK_ARG(params && off && x && out && d_out && stash && scratch, "net_bwd: null pointer");
  // grads == NULL: data gradient only (a frozen network, e.g. the discriminator inside train_gen whose parameter
  // gradients the next train_disc zeroes before use, train_IEMOCAP.py:221): every weight / bias / LayerNorm gradient is skipped
  const bool pg = grads != nullptr;
  GANFFN_CHECK_ARG(pg || dx, "net_bwd: neither parameter gradients nor dx requested");
  const Stash sl = stash_layout(nd);
  const Scratch sc = scratch_layout(nd);
  const Ctx cx{nd, scratch + sc.gemm, sc.gemm_floats, scratch + sc.red, st};
  const int T = nd.T(), d = nd.d;
  const float p_pe = train ? P_PE : 0.f, p_enc = train ? P_ENC : 0.f, p_hd = train ? p_head : 0.f;
  const int64_t* hoff = off + (int64_t)nd.L * PER_LAYER;
  auto P = [&](int64_t o) { return params + o; };
  auto G = [&](int64_t o) -> float* { return pg ? grads + o : nullptr; };
  const bool gen = nd.kind == GANFFN_NET_GENERATOR;

  if (pg && !accumulate) {
    // every gradient below is accumulated (red.global.add from the GEMM / LayerNorm kernels): start from zero
    auto Z = [&](int64_t o, int64_t n) { if (o >= 0) cudaMemsetAsync(grads + o, 0, (size_t)n * sizeof(float), st); };
    for (int l = 0; l < nd.L; ++l) {
      const int64_t* lo = off + (int64_t)l * PER_LAYER;
      Z(lo[IN_W], (int64_t)3 * d * d); Z(lo[IN_B], 3 * d); Z(lo[OUT_W], (int64_t)d * d); Z(lo[OUT_B], d);
      Z(lo[L1_W], (int64_t)nd.dff * d); Z(lo[L1_B], nd.dff); Z(lo[L2_W], (int64_t)d * nd.dff); Z(lo[L2_B], d);
      Z(lo[N1_W], d); Z(lo[N1_B], d); Z(lo[N2_W], d); Z(lo[N2_B], d);
    }
    Z(hoff[FC1_W], (int64_t)nd.h1 * d); Z(hoff[FC1_B], nd.h1); Z(hoff[FC2_W], (int64_t)nd.h2 * nd.h1); Z(hoff[FC2_B], nd.h2);
    if (!gen) { Z(hoff[FC3_W], nd.h2); Z(hoff[FC3_B], 1); }
    if (nd.has_object()) { Z(hoff[OBJ_W], (int64_t)d * nd.d_in); Z(hoff[OBJ_B], d); }
  }

  float* da = scratch + sc.da;    // gradient w.r.t. the current layer output
  float* db = scratch + sc.db;
  float* dz = scratch + sc.dz;
  float* dzd_buf = scratch + sc.dzd;
  float* hb1 = scratch + sc.hb1;
  float* hb2 = scratch + sc.hb2;
  float* hb3 = scratch + sc.hb3;
  const float* last_x2 = stash + sl.layer0 + (int64_t)(nd.L - 1) * sl.per_layer + sl.x2;

  // ---- head ----
  if (!gen && !g_deterministic && disc_head_fusable(d, nd.h1, nd.h2)) {
    // the whole discriminator head backward in one kernel (head.cu); its gradients accumulate with red.global.add, so
    // the deterministic mode keeps the sequence of launches below
    GANFFN_TRY(disc_head_bwd(d_out, out, last_x2, stash + sl.g0, stash + sl.f1, stash + sl.a1, stash + sl.f2, stash + sl.a2,
                             P(hoff[FC1_W]), P(hoff[FC2_W]), P(hoff[FC3_W]), da, G(hoff[FC1_W]), G(hoff[FC1_B]), G(hoff[FC2_W]),
                             G(hoff[FC2_B]), G(hoff[FC3_W]), G(hoff[FC3_B]), T, d, p_hd, seed, GANFFN_SITE_HEAD, st));
  } else {
  if (gen) {
    // out = gelu(f2), f2 = drop(fc2(a1))
    GANFFN_TRY(elementwise(d_out, stash + sl.f2, hb2, (int64_t)T * nd.h2, EW_DGELU_MASK, p_hd, seed, GANFFN_SITE_HEAD + 2, st));
  } else {
    // prob = sigmoid(drop(fc3(a2)))
    GANFFN_TRY(elementwise(d_out, out, hb3, (int64_t)T, EW_DSIGMOID_MASK, p_hd, seed, GANFFN_SITE_HEAD + 3, st));
    if (pg) GANFFN_TRY(cx.wgrad(hb3, stash + sl.a2, G(hoff[FC3_W]), G(hoff[FC3_B]), T, 1, nd.h2, accumulate));
    Epilogue ep; ep.dact = DACT_GELU; ep.dact_src = stash + sl.f2; ep.p_drop = p_hd; ep.seed = seed;
    ep.site = GANFFN_SITE_HEAD + 2;
    GANFFN_TRY(cx.dgrad(hb3, P(hoff[FC3_W]), hb2, T, 1, nd.h2, ep));
  }
  if (pg) GANFFN_TRY(cx.wgrad(hb2, stash + sl.a1, G(hoff[FC2_W]), G(hoff[FC2_B]), T, nd.h2, nd.h1, accumulate));
  {
    Epilogue ep; ep.dact = DACT_GELU; ep.dact_src = stash + sl.f1; ep.p_drop = p_hd; ep.seed = seed;
    ep.site = GANFFN_SITE_HEAD + 1;
    GANFFN_TRY(cx.dgrad(hb2, P(hoff[FC2_W]), hb1, T, nd.h2, nd.h1, ep));
  }
  if (pg) GANFFN_TRY(cx.wgrad(hb1, stash + sl.g0, G(hoff[FC1_W]), G(hoff[FC1_B]), T, nd.h1, d, accumulate));
  {
    Epilogue ep; ep.dact = DACT_GELU; ep.dact_src = last_x2; ep.p_drop = gen ? p_hd : 0.f; ep.seed = seed;
    ep.site = GANFFN_SITE_HEAD + 0;
    GANFFN_TRY(cx.dgrad(hb1, P(hoff[FC1_W]), da, T, nd.h1, d, ep));
  }
  }

  // ---- encoder layers, last to first ----
  // Data gradients on the caller's stream, weight gradients on the side stream (see SideCtx).  wait_slot(k) orders a
  // kernel that overwrites the input of a still-pending weight-gradient product of slot k behind that product.
  SideCtx* sx = pg ? side_ctx(st) : nullptr;
  float* gemm2 = scratch + sc.gemm2;
  float* dz1 = scratch + sc.dz1;
  float* dzd1_buf = scratch + sc.dzd1;
  if (sx) for (int i = 0; i < 4; ++i) sx->pending[i] = false;
  auto wgrad_side = [&](int slot, const float* dy, const float* xa, float* dw, float* dbias, int M, int N, int K) -> int {
    if (!pg) return GANFFN_OK;
    if (!sx) return cx.wgrad(dy, xa, dw, dbias, M, N, K, accumulate);
    cudaEventRecord(sx->ready[slot], st);
    cudaStreamWaitEvent(sx->side, sx->ready[slot], 0);
    const int rc = linear_wgrad(dy, xa, dw, dbias, M, N, K, 1, gemm2, sc.gemm_floats, sx->side);
    cudaEventRecord(sx->done[slot], sx->side);
    sx->pending[slot] = true;
    return rc;
  };
  auto wait_slot = [&](int slot) {
    if (sx && sx->pending[slot]) {
      cudaStreamWaitEvent(st, sx->done[slot], 0);
      sx->pending[slot] = false;
    }
  };
  for (int l = nd.L - 1; l >= 0; --l) {
    const float* base = stash + sl.layer0 + (int64_t)l * sl.per_layer;
    const int64_t* lo = off + (int64_t)l * PER_LAYER;
    const float* xin = l == 0 ? stash + sl.x0 : stash + sl.layer0 + (int64_t)(l - 1) * sl.per_layer + sl.x2;
    float* dzd = p_enc > 0.f ? dzd_buf : dz;
    float* dzd1 = p_enc > 0.f ? dzd1_buf : dz1;

    // x2 = LN2(z2), z2 = x1 + drop(linear2(h))
    wait_slot(W_L2);
    GANFFN_TRY(layernorm_bwd(da, base + sl.z2, P(lo[N2_W]), dz, p_enc > 0.f ? dzd_buf : nullptr, G(lo[N2_W]), G(lo[N2_B]),
                             G(lo[L2_B]), T, d, 1, p_enc, seed, GANFFN_SITE_LAYER(l, 3), st));
    GANFFN_TRY(wgrad_side(W_L2, dzd, base + sl.h, G(lo[L2_W]), nullptr, T, d, nd.dff));
    wait_slot(W_L1);
    {
      Epilogue ep; ep.dact = DACT_NONZERO; ep.dact_src = base + sl.h; ep.dact_scale = p_enc > 0.f ? 1.f / (1.f - p_enc) : 1.f;
      GANFFN_TRY(cx.dgrad(dzd, P(lo[L2_W]), scratch + sc.dh, T, d, nd.dff, ep));
    }
    GANFFN_TRY(wgrad_side(W_L1, scratch + sc.dh, base + sl.x1, G(lo[L1_W]), G(lo[L1_B]), T, nd.dff, d));
    // x1 = LN1(z1), z1 = xin + drop(out_proj(o)): the LayerNorm-1 backward rides behind the product that makes its dy
    // (Epilogue::lnb_*: inside the split-K fold on the tcgen05 engine, stand-alone otherwise)
    wait_slot(W_OUT);     // the weight-gradient stream may still be reading dz1 / dzd1 of the layer above
    {
      Epilogue ep; ep.residual = dz; ep.ldr = d;
      ep.lnb_z = base + sl.z1; ep.lnb_gamma = P(lo[N1_W]); ep.lnb_dz = dz1; ep.lnb_dzd = p_enc > 0.f ? dzd1_buf : nullptr;
      ep.lnb_dgamma = G(lo[N1_W]); ep.lnb_dbeta = G(lo[N1_B]); ep.lnb_dbias = G(lo[OUT_B]);
      ep.lnb_p = p_enc; ep.lnb_seed = seed; ep.lnb_site = GANFFN_SITE_LAYER(l, 1);
      GANFFN_TRY(cx.dgrad(scratch + sc.dh, P(lo[L1_W]), db, T, nd.dff, d, ep));
    }
    GANFFN_TRY(wgrad_side(W_OUT, dzd1, base + sl.o, G(lo[OUT_W]), nullptr, T, d, d));
    GANFFN_TRY(cx.dgrad(dzd1, P(lo[OUT_W]), scratch + sc.d_o, T, d, d, Epilogue{}));
    wait_slot(W_IN);
    GANFFN_TRY(attention_bwd(base + sl.qkv, base + sl.o, base + sl.lse, scratch + sc.d_o, scratch + sc.dqkv, nd.S, nd.B,
                             d, nd.nhead, p_enc, seed, GANFFN_SITE_LAYER(l, 0), st));
    GANFFN_TRY(wgrad_side(W_IN, scratch + sc.dqkv, xin, G(lo[IN_W]), G(lo[IN_B]), T, 3 * d, d));
    {
      Epilogue ep; ep.residual = dz1; ep.ldr = d;
      GANFFN_TRY(cx.dgrad(scratch + sc.dqkv, P(lo[IN_W]), da, T, 3 * d, d, ep));
    }
    if (sx && l < MAX_LAYER_EVENTS) {   // layer l's gradients are complete once both streams get here
      cudaEventRecord(sx->layer_tmp, st);
      cudaStreamWaitEvent(sx->side, sx->layer_tmp, 0);
      cudaEventRecord(sx->layer_done[l], sx->side);
    }
  }
  if (sx) sx->layers_recorded = nd.L < MAX_LAYER_EVENTS ? nd.L : MAX_LAYER_EVENTS;
  if (sx) {   // join: every weight gradient has landed before the caller's stream continues
    cudaEventRecord(sx->join, sx->side);
    cudaStreamWaitEvent(st, sx->join, 0);
    for (int i = 0; i < 4; ++i) sx->pending[i] = false;
  }

  // ---- positional encoding (dropout only) and the optional `object` projection ----
  if (nd.has_object()) {
    const float* dpe = da;
    if (p_pe > 0.f) {
      GANFFN_TRY(elementwise(da, nullptr, db, (int64_t)T * d, EW_MASK, p_pe, seed, GANFFN_SITE_PE, st));
      dpe = db;
    }
    if (pg) GANFFN_TRY(cx.wgrad(dpe, x, G(hoff[OBJ_W]), G(hoff[OBJ_B]), T, d, nd.d_in, accumulate));
    if (dx) GANFFN_TRY(cx.dgrad(dpe, P(hoff[OBJ_W]), dx, T, d, nd.d_in, Epilogue{}));
  } else if (dx) {
    GANFFN_TRY(elementwise(da, nullptr, dx, (int64_t)T * d, EW_MASK, p_pe, seed, GANFFN_SITE_PE, st));
  }
  return GANFFN_OK;
}

}  // namespace ganffn
