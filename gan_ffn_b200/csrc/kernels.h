// Internal (C++) entry points of the kernel files; capi.cu wraps them as the C ABI.
#pragma once
#include "common.cuh"

namespace ganffn {

// gemm_simt.cu
int gemm_simt(const float* A, int lda, bool transA, const float* B, int ldb, bool b_is_nk, float* C, int ldc, int M,
              int N, int K, const Epilogue& ep, float* scratch, int64_t scratch_floats, cudaStream_t st);
int64_t gemm_simt_scratch_floats(int M, int N, int K);

// gemm_tc.cu (tcgen05 3xTF32)
bool gemm_tc_supported(bool transA, bool b_is_nk, int lda, int ldb, int ldc, int M, int N, int K, const void* A,
                       const void* B);
int gemm_tc(const float* A, int lda, bool transA, const float* B, int ldb, bool b_is_nk, float* C, int ldc, int M, int N,
            int K, const Epilogue& ep, float* scratch, int64_t scratch_floats, cudaStream_t st);
int64_t gemm_tc_scratch_floats(int M, int N, int K);

// attention.cu
int attention_fwd(const float* qkv, float* o, float* lse, int S, int B, int d, int nhead, float p, Seed seed,
                  int site, cudaStream_t st);
int attention_bwd(const float* qkv, const float* o, const float* lse, const float* d_o, float* dqkv, int S, int B, int d,
                  int nhead, float p, Seed seed, int site, cudaStream_t st);

// attention_mma.cu (tensor-core path, 3xTF32 on mma.sync fragments); return -1 for an unsupported head_dim
int attention_fwd_mma(const float* qkv, float* o, float* lse, int S, int B, int d, int nhead, float p, Seed seed,
                      int site, cudaStream_t st);
int attention_bwd_mma(const float* qkv, const float* o, const float* lse, const float* d_o, float* dqkv, int S, int B, int d,
                      int nhead, float p, Seed seed, int site, cudaStream_t st);

// rowwise.cu
int layernorm_fwd(const float* z, const float* gamma, const float* beta, float* y, int T, int d, cudaStream_t st);
int layernorm_bwd(const float* dy, const float* z, const float* gamma, float* dz, float* dz_drop, float* dgamma,
                  float* dbeta, float* dbias_sub, int T, int d, int accumulate, float p, Seed seed, int site,
                  cudaStream_t st);
// the stand-alone LayerNorm backward behind a product whose Epilogue asked for it (lnb_* fields) and could not fuse it
int layernorm_bwd_after(const Epilogue& ep, const float* dy, int M, int N, cudaStream_t st);
int posenc_fwd(const float* x, const float* pe, float* y, int S, int B, int d, float p, Seed seed, cudaStream_t st);
enum { EW_GELU_DROP = 0, EW_DGELU_MASK = 1, EW_DSIGMOID_MASK = 2, EW_MASK = 3 };
int elementwise(const float* a, const float* src, float* y, int64_t n, int mode, float p, Seed seed, int site,
                cudaStream_t st);
int colsum(const float* a, int M, int N, float* out, int accumulate, cudaStream_t st);
// dw[N,K] (+)= dy[M,N]^T x[M,K];  db[N] (+)= column sums of dy (db may be null).  accumulate = 0 zeroes first.
int linear_wgrad(const float* dy, const float* x, float* dw, float* db, int M, int N, int K, int accumulate,
                 float* gemm_scratch, int64_t gemm_scratch_floats, cudaStream_t st);
int dropout_mask(float* out, int64_t rows, int64_t cols, int64_t row_stride, float p, uint64_t seed, int site,
                 cudaStream_t st);

// losses.cu
int fuse_cls_fwd(const float* a, const float* v, const float* t, const float* w, const float* b, float* fusion,
                 float* logp, int T, int dh, int C, cudaStream_t st);
int fuse_cls_bwd(const float* dlp, const float* logp, const float* fusion, const float* w, float* d_fusion, float* dw,
                 float* db, int T, int dh, int C, int accumulate, float* scratch, cudaStream_t st);
int64_t fuse_cls_scratch_floats(int T, int dh, int C);
int masked_nll_fwd(const float* pred, const int64_t* target, const float* mask, const float* weight, float* out,
                   int64_t n, int C, float den_override, cudaStream_t st);
int masked_nll_bwd(const float* d_loss, const float* loss_and_den, const int64_t* target, const float* mask,
                   const float* weight, float* d_pred, int64_t n, int C, cudaStream_t st);
int bce_fwd(const float* prob, const float* target, float* loss, int64_t n, float scale, cudaStream_t st);
int bce_bwd(const float* d_loss, const float* prob, const float* target, float* d_prob, int64_t n, float scale,
            cudaStream_t st);
int adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const int* step_dev, float lr, float b1,
                  float b2, float eps, float wd, float gscale, cudaStream_t st);
int adam_step(float* p, const float* g, float* m, float* v, int64_t n, int step, float lr, float b1, float b2, float eps,
              float wd, float gscale, cudaStream_t st);

// head.cu: the discriminator head (gelu -> fc1 -> fc2 -> fc3 -> sigmoid, model.py:1320-1327) as one kernel
bool disc_head_fusable(int d, int h1, int h2);
int disc_head_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                  const float* b3, float* g0, float* f1, float* a1, float* f2, float* a2, float* out, int T, int d, float p_drop,
                  Seed seed, int site0, cudaStream_t st);

int disc_head_bwd(const float* d_out, const float* out, const float* x, const float* g0, const float* f1, const float* a1,
                  const float* f2, const float* a2, const float* w1, const float* w2, const float* w3, float* dx, float* dw1,
                  float* db1, float* dw2, float* db2, float* dw3, float* db3, int T, int d, float p_drop, Seed seed, int site0,
                  cudaStream_t st);

// net.cu
struct NetDims {
  int kind, S, B, d_in, d, nhead, dff, L, h1, h2;
  int T() const { return S * B; }
  bool has_object() const { return d_in != d; }
};
int net_fwd(const NetDims& nd, const float* params, const int64_t* off, const float* pe, const float* x, float* out,
            float* stash, float* scratch, int train, float p_head, Seed seed, cudaStream_t st);
int net_bwd(const NetDims& nd, const float* params, const int64_t* off, const float* x, const float* out,
            const float* d_out, const float* stash, float* grads, float* dx, float* scratch, int train, float p_head,
            Seed seed, int accumulate, cudaStream_t st);
int net_bwd_layer_wait(cudaStream_t bwd_stream, int layer, cudaStream_t waiting);
int64_t net_stash_floats(const NetDims& nd);
int64_t net_scratch_floats(const NetDims& nd);
int net_check(const NetDims& nd);

}  // namespace ganffn
