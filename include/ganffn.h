/*
 * ganffn.h -- C ABI of the B200 (sm_100a) GAN-FFN fusion hot path.
 *
 * The reference (Jing-yilin/GAN-FFN) has no FFI of its own: its seam is the Python
 * nn.Module surface of model.py.  Every entry point below therefore
 * cites the reference *module code* it replaces; the Python host in gan_ffn_b200/
 * keeps the reference's constructors and forward signatures and calls these through
 * ctypes with raw device pointers (tensor.data_ptr()).
 *
 * Conventions
 *  - All tensors are fp32, contiguous, device memory owned by the caller.
 *  - Activations are seq-major (S,B,d) exactly as the reference feeds them
 *    (model.py:1194), i.e. a row-major [T=S*B, d] matrix with row t = s*B + b.
 *  - `stream` is a cudaStream_t passed as void*.  Nothing here allocates,
 *    synchronises or throws.  Return value: GANFFN_OK or an error code;
 *    ganffn_last_error() gives the text.
 *  - Dropout: counter-based SplitMix64 hash keyed by (seed, site, element group of 4), 16-bit uniforms.  The same
 *    (seed, site) regenerates the same mask in the backward pass; p == 0 is eval mode.
 *    ganffn_dropout_mask() exports the mask a site draws so tests can inject it into
 *    the CPU oracle.
 */
#ifndef GANFFN_H
#define GANFFN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GANFFN_OK 0
#define GANFFN_ERR_ARG 1   /* shape / pointer precondition violated */
#define GANFFN_ERR_CUDA 2  /* a kernel launch failed */

#define GANFFN_MAX_SEQ 110 /* PositionalEncoding(max_len=110), model.py:1179 */

/* dropout sites (shared with oracle/ganffn_oracle.py) */
#define GANFFN_SITE_PE 0
#define GANFFN_SITE_LAYER(l, k) (16 * ((l) + 1) + (k)) /* k: 0 attn P, 1 out-proj, 2 FFN hidden, 3 linear2 */
#define GANFFN_SITE_HEAD 200                            /* +0 gelu(enc) [gen], +1 fc1, +2 fc2, +3 fc3 */

/* activation codes for ganffn_linear_fwd */
#define GANFFN_ACT_NONE 0
#define GANFFN_ACT_RELU 1
#define GANFFN_ACT_GELU 2
#define GANFFN_ACT_SIGMOID 3

/* network kinds for the whole-network entry points */
#define GANFFN_NET_GENERATOR 0     /* model.py:1200-1294 */
#define GANFFN_NET_DISCRIMINATOR 1 /* model.py:1297-1397 */

/* GEMM engines */
#define GANFFN_GEMM_AUTO 0
#define GANFFN_GEMM_SIMT 1   /* fp32 FFMA tiles */
#define GANFFN_GEMM_TC 2     /* tcgen05 kind::tf32, 3xTF32 error-compensated */
#define GANFFN_GEMM_TF32X1 3 /* REDUCED PRECISION variant: one tcgen05 kind::tf32 MMA per product (10-bit mantissa operands,
                              * fp32 accumulate) -- the reduced-precision variant north_star allows at rtol 2e-2 (it names
                              * bf16; single-pass TF32 keeps 3 more mantissa bits at the same one-MMA-per-product cost model).
                              * Never the fp32-parity headline: bench.py reports it as a separate line. */

/* ---- library state -------------------------------------------------------------------- */
int ganffn_version(void);
const char* ganffn_last_error(void);
/* Number of kernels this library has launched since load / last reset (bench.py: gpu_launches). */
unsigned long long ganffn_launch_count(void);
void ganffn_reset_launch_count(void);
/* Select the GEMM engine (GANFFN_GEMM_*); returns the previous value. */
int ganffn_set_gemm_engine(int engine);

/* Deterministic gradient accumulation (the reference pins determinism: train_IEMOCAP.py:46-53).  Off (default): weight
 * gradients of split-K slices, LayerNorm parameter gradients and bias gradients are accumulated with red.global.add --
 * fast, order-dependent in the last bits.  On: split-K slices go through partial buffers folded in slice order,
 * LayerNorm partials are folded in block order by the block that finishes last, bias gradients have a single writer;
 * two runs from the same state are then bit-identical, provided only one backward pass at a time adds to a gradient
 * arena (the host side runs the networks serially in this mode).  Returns the previous setting. */
int ganffn_set_deterministic(int on);

/* Weight-gradient products of ganffn_net_bwd on a library-owned side stream (forked from / joined to `stream` by
 * events; default on).  Returns the previous setting.  bench.py switches it off for its per-kernel roofline leg. */
int ganffn_set_side_streams(int on);

/* GEMM profiling for the roofline leg of bench.py: when enabled every GEMM (with its split-K
 * fold) is bracketed by CUDA events on the launching stream.  collect() synchronises, sums the
 * elapsed milliseconds and algorithmic FLOPs (2*M*N*K) of the GEMMs run by `engine`
 * (GANFFN_GEMM_AUTO = all) since the last collect, and clears the record. */
void ganffn_gemm_profile_enable(int on);
int ganffn_gemm_profile_collect(int engine, double* total_ms, double* total_flops, int64_t* launches);
/* Per-shape view of the same record (call before collect(), which clears it): row r is
 * shapes[6r..6r+5] = {M, N, K, transA, b_is_nk, engine}, ms[2r] = summed milliseconds, ms[2r+1] = launches.
 * Returns the number of rows (<= max_rows), -1 on error.  When the bracketed calls were recorded inside a stream
 * capture the brackets are external event-record nodes: replay the graph, then read the table. */
int64_t ganffn_gemm_profile_table(int64_t* shapes, double* ms, int64_t max_rows);

/* ---- primitives (each is also used by the whole-network calls) ------------------------ */

/* y[M,N] = epilogue(x[M,K] @ w[N,K]^T + bias).  Replaces nn.Linear inside
 * TransformerEncoderLayer (torch transformer.py:961-982) and fc1/fc2/fc3/object/fc
 * (model.py:1214-1215, 1311-1313, 1344, 1432).
 *   drop_before_act = 1: y = act(drop(v))  (heads, model.py:1227-1228, 1324-1326)
 *   drop_before_act = 0: y = drop(act(v))  (FFN hidden, transformer.py:981)
 *   residual (optional) is added last; pre (optional) receives the value fed to act. */
int ganffn_linear_fwd(const float* x, const float* w, const float* bias, const float* residual,
                      float* y, float* pre, int M, int N, int K, int act, int drop_before_act,
                      float p_drop, uint64_t seed, int site, float* scratch,
                      int64_t scratch_floats, void* stream);
/* Post-norm sublayer tail in one call: z[M,N] = residual + drop(x[M,K] @ w[N,K]^T + bias), y[M,N] = LayerNorm(z) * gamma
 * + beta (eps 1e-5).  Replaces `x = norm(x + dropout(linear(..)))` of TransformerEncoderLayer (torch transformer.py
 * :944-982, post-norm branch; out-proj + norm1 and linear2 + norm2).  z is kept because the backward pass reads it.
 * The tcgen05 engine normalises inside the GEMM epilogue when N <= 128 (or inside its split-K fold); other shapes and
 * the FFMA engine run the stand-alone LayerNorm kernel behind the product -- same results either way. */
int ganffn_linear_ln_fwd(const float* x, const float* w, const float* bias, const float* residual,
                         const float* gamma, const float* beta, float* z, float* y, int M, int N, int K,
                         float p_drop, uint64_t seed, int site, float* scratch, int64_t scratch_floats,
                         void* stream);
/* The discriminator head (model.py:1320-1327, 1354-1364, 1390-1397) as one kernel each way:
 *   g0 = gelu(x);  f1 = drop(fc1 g0) [64];  a1 = gelu(f1);  f2 = drop(fc2 a1) [16];  a2 = gelu(f2);  prob = sigmoid(drop(fc3 a2))
 * x [T,d] (d <= 128, multiple of 4) is the last encoder output; w1 [64,d], w2 [16,64], w3 [16]; g0, f1, a1, f2, a2 are the
 * intermediates the backward pass reads.  Dropout sites `site0 + 1..3` (fc1, fc2, fc3), masks = ganffn_dropout_mask.
 * Backward: dx [T,d]; dw1 .. db3 are ACCUMULATED into (red.global.add); pass all six as NULL for a frozen network. */
int ganffn_disc_head_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                         const float* w3, const float* b3, float* g0, float* f1, float* a1, float* f2, float* a2,
                         float* prob, int T, int d, float p_drop, uint64_t seed, int site0, void* stream);
int ganffn_disc_head_bwd(const float* d_prob, const float* prob, const float* x, const float* g0, const float* f1,
                         const float* a1, const float* f2, const float* a2, const float* w1, const float* w2,
                         const float* w3, float* dx, float* dw1, float* db1, float* dw2, float* db2, float* dw3,
                         float* db3, int T, int d, float p_drop, uint64_t seed, int site0, void* stream);
/* Split-K workspace (floats) the GEMM engines want for an [M,N,K] product; may be 0. */
int64_t ganffn_gemm_scratch_floats(int M, int N, int K);

/* dx[M,K] = dy[M,N] @ w[N,K] (+ residual[M,K]).  Backward-data of nn.Linear. */
int ganffn_linear_dgrad(const float* dy, const float* w, const float* residual, float* dx,
                        int M, int N, int K, float* scratch, int64_t scratch_floats,
                        void* stream);

/* dw[N,K] (+)= dy[M,N]^T @ x[M,K]; db[N] (+)= column sums of dy.  Backward-weight of
 * nn.Linear.  accumulate != 0 adds into dw/db.  scratch: >= ganffn_wgrad_scratch_floats(). */
int ganffn_linear_wgrad(const float* dy, const float* x, float* dw, float* db, int M, int N, int K,
                        int accumulate, float* scratch, void* stream);
int64_t ganffn_wgrad_scratch_floats(int M, int N, int K);

/* Self-attention core for all (dialogue, head) pairs: o = softmax(q k^T / sqrt(hd)) v with
 * dropout on the probabilities.  qkv is the packed in-proj output [T, 3d]; o is [T, d];
 * lse [B*nhead*S] receives the row log-sum-exp for the backward pass.
 * Replaces F.scaled_dot_product_attention inside nn.MultiheadAttention
 * (torch functional.py multi_head_attention_forward). */
int ganffn_attention_fwd(const float* qkv, float* o, float* lse, int S, int B, int d, int nhead,
                         float p_drop, uint64_t seed, int site, void* stream);
int ganffn_attention_bwd(const float* qkv, const float* o, const float* lse, const float* d_o,
                         float* dqkv, int S, int B, int d, int nhead, float p_drop, uint64_t seed,
                         int site, void* stream);

/* y = LayerNorm(z) (eps 1e-5, torch default).  z already holds x + sublayer(x). */
int ganffn_layernorm_fwd(const float* z, const float* gamma, const float* beta, float* y, int T,
                         int d, void* stream);
/* dz = LN backward; dz_drop (optional) = dz * dropout_mask(site) for the sublayer branch;
 * dgamma/dbeta (+)= reductions.  scratch: >= ganffn_layernorm_scratch_floats(). */
int ganffn_layernorm_bwd(const float* dy, const float* z, const float* gamma, float* dz,
                         float* dz_drop, float* dgamma, float* dbeta, int T, int d, int accumulate,
                         float p_drop, uint64_t seed, int site, float* scratch, void* stream);
int64_t ganffn_layernorm_scratch_floats(int T, int d);

/* y[s,b,:] = drop(x[s,b,:] + pe[s,:])  (PositionalEncoding.forward, model.py:1191-1197).
 * pe is the module's registered buffer viewed as [max_len, d] (model.py:1186-1189). */
int ganffn_posenc_fwd(const float* x, const float* pe, float* y, int S, int B, int d, float p_drop,
                      uint64_t seed, void* stream);

/* Writes out[rows,cols] = the scaled keep mask (0 or 1/(1-p)) that `site` draws; element (r,c)
 * is dropout element r*row_stride + c.  row_stride == cols for every [T,N] site; the attention
 * site uses rows = B*nhead*S, cols = S, row_stride = round_up(S,4). */
int ganffn_dropout_mask(float* out, int64_t rows, int64_t cols, int64_t row_stride, float p_drop,
                        uint64_t seed, int site, void* stream);

/* ---- losses ---------------------------------------------------------------------------- */
/* log_prob[T,C] = log_softmax((a+v+t) @ w[C,100]^T + b)  (GAN_FFN.forward, model.py:1444-1449).
 * fusion[T,dh] (optional) receives a+v+t (GAN_FFN_DialogueRNN.forward, model.py:1524). */
int ganffn_fuse_cls_fwd(const float* a, const float* v, const float* t, const float* w,
                        const float* b, float* fusion, float* log_prob, int T, int dh, int C,
                        void* stream);
/* Given d_log_prob: d_fusion[T,dh] (same for a, v and t), dw[C,dh] (+)=, db[C] (+)=. */
int ganffn_fuse_cls_bwd(const float* d_log_prob, const float* log_prob, const float* fusion,
                        const float* w, float* d_fusion, float* dw, float* db, int T, int dh, int C,
                        int accumulate, float* scratch, void* stream);
int64_t ganffn_fuse_cls_scratch_floats(int T, int dh, int C);

/* MaskedNLLLoss.forward (model.py:68-81): pred [n,C] log-probs, target int64 [n], mask [n],
 * weight [C] or NULL.  loss_and_den[0] = loss, [1] = denominator sum(w[t]*m).
 * `den_override` > 0 replaces the denominator (global denominator under dialogue sharding). */
int ganffn_masked_nll_fwd(const float* pred, const int64_t* target, const float* mask,
                          const float* weight, float* loss_and_den, int64_t n, int C,
                          float den_override, void* stream);
int ganffn_masked_nll_bwd(const float* d_loss, const float* loss_and_den, const int64_t* target,
                          const float* mask, const float* weight, float* d_pred, int64_t n, int C,
                          void* stream);

/* torch.nn.BCELoss() (train_IEMOCAP.py:300): mean over n of -[y log p + (1-y) log(1-p)], logs
 * clamped at -100.  `scale` multiplies the mean (1/world_size under dialogue sharding). */
int ganffn_bce_fwd(const float* prob, const float* target, float* loss, int64_t n, float scale,
                   void* stream);
int ganffn_bce_bwd(const float* d_loss, const float* prob, const float* target, float* d_prob,
                   int64_t n, float scale, void* stream);

/* ---- optimizer ------------------------------------------------------------------------- */
/* torch.optim.Adam over one flat parameter arena (train_IEMOCAP.py:292-297, :661): L2 term
 * folded into the gradient, bias-corrected, eps outside the sqrt.  `step` is 1-based.
 * grad_scale multiplies g first (1/world_size after a sum-allreduce). */
int ganffn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int step, float lr,
                     float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                     void* stream);
/* Same, with the step count read from device memory (`*step_dev`, >= 1) so that the call can be captured in a CUDA
 * graph; the caller increments the counter on the same stream before the call. */
int ganffn_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const int* step_dev, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream);

/* ---- whole networks -------------------------------------------------------------------- */
/* One generator / discriminator forward: [object] -> PE -> nlayers x encoder layer -> head.
 *   params      flat fp32 arena; off[] gives each tensor's float offset in canonical order:
 *               per layer l (12 entries): in_proj_weight, in_proj_bias, out_proj.weight,
 *               out_proj.bias, linear1.weight, linear1.bias, linear2.weight, linear2.bias,
 *               norm1.weight, norm1.bias, norm2.weight, norm2.bias; then head entries:
 *               generator: fc1.w, fc1.b, fc2.w, fc2.b;
 *               discriminator: fc1.w, fc1.b, fc2.w, fc2.b, fc3.w, fc3.b, object.w, object.b
 *               (object.* = -1 when absent).
 *   x           (S,B,d_in); d_in == d, or 512 with d == 100 for the visual discriminator's
 *               real input (model.py:1355-1356).
 *   pe          the PositionalEncoding buffer viewed as [max_len, d].
 *   h1, h2      head widths: generator fc1 d->h1, fc2 h1->h2 (= D_h); discriminator fc1 d->h1 (64),
 *               fc2 h1->h2 (16), fc3 h2->1.
 *   out         generator: (S,B,h2) fused feature; discriminator: (S,B,1) probability.
 *   stash       activation stash for the backward pass, ganffn_net_stash_floats() floats.
 *   scratch     ganffn_net_scratch_floats() floats of workspace (split-K partials etc.).
 *   p_scale     0 = eval mode (all dropout off); 1 = train mode (reference probabilities:
 *               PE 0.2, encoder 0.1, head `p_head`).
 *   seed        dropout seed of this forward call; when `seed_dev` is not NULL the kernels read the seed from
 *               that device word instead, so a captured CUDA graph draws fresh masks on every replay.
 * Replaces model.py:1221-1231, 1255-1263, 1286-1294, 1320-1327, 1354-1364, 1390-1397. */
int ganffn_net_fwd(int kind, const float* params, const int64_t* off, const float* pe,
                   const float* x, float* out, float* stash, float* scratch, int S, int B, int d_in,
                   int d, int nhead, int dff, int nlayers, int h1, int h2, int train, float p_head,
                   uint64_t seed, const uint64_t* seed_dev, void* stream);
/* Backward of the above.  grads has the arena's layout; accumulate != 0 adds into it.
 * dx may be NULL when the input needs no gradient.  grads may be NULL for a frozen network (data gradient only:
 * no weight, bias or LayerNorm gradient is computed -- the discriminator inside train_gen, whose parameter gradients
 * the next train_disc zeroes before use, train_IEMOCAP.py:221, 245-251).  scratch: ganffn_net_scratch_floats(). */
int ganffn_net_bwd(int kind, const float* params, const int64_t* off, const float* x,
                   const float* out, const float* d_out_grad, const float* stash, float* grads,
                   float* dx, float* scratch, int S, int B, int d_in, int d, int nhead, int dff,
                   int nlayers, int h1, int h2, int train, float p_head, uint64_t seed,
                   const uint64_t* seed_dev, int accumulate, void* stream);
/* Per-layer gradient completion, for overlapping the data-parallel all-reduce with the rest of the backward pass
 * ("bucket per encoder layer and launch as each layer's weight-grad completes"): makes
 * `waiting_stream` wait, on the device, until every gradient of encoder layer `layer` written by the LAST
 * ganffn_net_bwd issued on `bwd_stream` has landed.  Layers complete from nlayers-1 down to 0; the head gradients and
 * layer 0 are complete when ganffn_net_bwd's own work on `bwd_stream` is.  Returns GANFFN_ERR_ARG when no per-layer
 * event exists (side streams off): wait for the whole pass instead. */
int ganffn_net_bwd_layer_wait(void* bwd_stream, int layer, void* waiting_stream);
/* Both return -1 when the shape violates a precondition (ganffn_last_error() says which). */
int64_t ganffn_net_stash_floats(int kind, int S, int B, int d_in, int d, int nhead, int dff,
                                int nlayers, int h1, int h2);
int64_t ganffn_net_scratch_floats(int kind, int S, int B, int d_in, int d, int nhead, int dff,
                                  int nlayers, int h1, int h2);

/* ---- dialogue graph (north_star parts 2-3) --------------------------------------------------------------------
 * ABSENT FROM THE REFERENCE: /root/reference has no edge construction and no graph
 * convolution, so there is nothing to be bit-exact against ("parity unpinned -- no reference implementation").
 * The semantics below are this library's own (DialogueGCN-style window graph), pinned by oracle/graph_oracle.py:
 *   - nodes: the real utterances, dialogue-major: node(b, t) = node_off[b] + t, t < lengths[b];
 *   - edges: for every target i of a dialogue, one edge from every source j with i - wp <= j <= i + wf (self
 *     loop included), stored as CSR over targets with sources ascending -- this is the canonical edge order, and
 *     edge_index[0][e] = source, edge_index[1][e] = target in that order;
 *   - edge_type[e] = ((speaker[j] * n_speakers + speaker[i]) << 1) | (j < i ? 0 : 1)   (2 * n_speakers^2 relations);
 *   - the transposed structure (CSR over sources: targets i with j - wf <= i <= j + wp, ascending) makes the
 *     backward scatter-add a segmented gather as well: no atomics anywhere.
 * lengths [B] and speakers [N] are int32 device arrays; offsets are int64; col / etype are int32. */
int64_t ganffn_graph_num_edges_host(const int* lengths_host, int n_dialogues, int wp, int wf, int64_t* n_nodes);
/* node_off[B+1], edge_off[B+1] (exclusive scans; one thread block). */
int ganffn_graph_offsets(const int* lengths, int n_dialogues, int wp, int wf, int64_t* node_off,
                         int64_t* edge_off, void* stream);
/* Warp per dialogue, straight into CSR.  transposed = 0: rows are targets (rowptr/col/etype as above; node_b /
 * node_t [N] and inv_cnt [N, n_rel] = 1 / #edges of that relation into the node (0 if none) are written when not
 * null; edge_index [2, E] int64 optional).  transposed = 1: rows are sources, col = targets, same etype. */
int ganffn_graph_build(const int* lengths, const int* speakers, const int64_t* node_off,
                       const int64_t* edge_off, int n_dialogues, int wp, int wf, int n_speakers,
                       int transposed, int64_t* rowptr, int* col, int* etype, int64_t* edge_index,
                       int64_t n_edges, int* node_b, int* node_t, float* inv_cnt, void* stream);
/* (S,B,d) zero-padded batch <-> packed [N,d] node features (and the same pair for gradients). */
int ganffn_graph_pack(const float* x_sbd, const int* node_b, const int* node_t, float* x_nodes,
                      int64_t n_nodes, int B, int d, void* stream);
int ganffn_graph_unpack(const float* x_nodes, const int* lengths, const int64_t* node_off, float* x_sbd,
                        int S, int B, int d, void* stream);
/* Device-side collate (reference dataloader.py:55-58, `pad_sequence` inside collate_fn).  The loader's batch is copied
 * to the device PACKED -- features [N, d] of the real utterances only, speakers [N], labels [N], lengths [B],
 * node_off [B+1] = exclusive prefix sums of lengths -- and padded here: features with ganffn_graph_unpack (packed
 * [N, d] -> zero-padded (S,B,d)), the per-utterance metadata with this call:
 *   qmask (S,B,n_speakers) one-hot speakers, umask (B,S) 1/0, label (B,S) int64, all zero at padded slots. */
int ganffn_collate_meta(const int* speakers, const int64_t* labels, const int* lengths, const int64_t* node_off,
                        float* qmask, float* umask, int64_t* label, int S, int B, int n_speakers, void* stream);
/* Relation-typed mean aggregation: out[n, r, :] = inv_cnt[n, r] * sum_{e in row n, etype[e] = r} x[col[e], :].
 * out is [N, n_rel, d] (empty relations are written as zeros): the dense contraction with the relation weights is
 * then one GEMM of [N, n_rel*d] x [n_rel*d, h] (ganffn_linear_fwd).
 * node_off [B+1] / n_dialogues / max_len (optional: NULL, 0, 0): when given and a dialogue's rows fit in shared
 * memory, one CTA stages one dialogue and HBM sees every input row once (otherwise rows are re-read from L2). */
int ganffn_graph_gather_typed(const float* x, const int64_t* rowptr, const int* col, const int* etype,
                              const float* inv_cnt, float* out, int64_t n_nodes, int n_rel, int d,
                              const int64_t* node_off, int n_dialogues, int max_len, void* stream);
/* Plain segmented gather-sum: out[n, :] = sum_{e in row n} w_e * in[col[e], slot_e, :], in is [N, in_slots, d];
 * slot_e = etype[e] when in_slots > 1 else 0; w_e = inv_cnt[col[e], etype[e]] when inv_cnt != NULL else 1.
 * Serves GraphConv forward (CSR), GraphConv backward (transposed CSR) and the backward of
 * ganffn_graph_gather_typed (transposed CSR, in = d_out [N, n_rel, d], weights = inv_cnt). */
int ganffn_graph_gather_sum(const float* in, const int64_t* rowptr, const int* col, const int* etype,
                            const float* inv_cnt, float* out, int64_t n_nodes, int in_slots, int n_rel,
                            int d, const int64_t* node_off, int n_dialogues, int max_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GANFFN_H */
