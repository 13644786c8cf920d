// Shared device/host helpers for the GAN-FFN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <algorithm>
#include <atomic>

#include "../../include/ganffn.h"

namespace ganffn {

// ---- library state (defined in capi.cu) ------------------------------------------------
extern unsigned long long g_launches;
extern int g_gemm_engine;
extern int g_deterministic;  // fixed-order gradient accumulation (ganffn_set_deterministic)
extern int g_side_streams;   // weight gradients on a side stream inside net_bwd (ganffn_set_side_streams)
void set_error(const char* fmt, ...);

#define GANFFN_CHECK_ARG(cond, ...)                \
  do {                                             \
    if (!(cond)) {                                 \
      ::ganffn::set_error(__VA_ARGS__);            \
      return GANFFN_ERR_ARG;                       \
    }                                              \
  } while (0)

// Call after every kernel launch: counts it and turns a launch error into a status code.
#define GANFFN_LAUNCHED(name)                                                     \
  do {                                                                            \
    ++::ganffn::g_launches;                                                       \
    cudaError_t e__ = cudaPeekAtLastError();                                      \
    if (e__ != cudaSuccess) {                                                     \
      ::ganffn::set_error("%s: %s", name, cudaGetErrorString(e__));               \
      return GANFFN_ERR_CUDA;                                                     \
    }                                                                             \
  } while (0)

#define GANFFN_TRY(expr)             \
  do {                               \
    int rc__ = (expr);               \
    if (rc__ != GANFFN_OK) return rc__; \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: opt in once per (kernel, device).  One
// static bit mask per call site (= per kernel instantiation); the atomic makes the check safe between the main
// thread and autograd's worker thread (setting the attribute twice is harmless).
#define GANFFN_SMEM_OPTIN(kernel, bytes)                                                                  \
  do {                                                                                                    \
    static std::atomic<unsigned long long> done__{0ull};                                                  \
    int dev__ = 0;                                                                                        \
    cudaGetDevice(&dev__);                                                                                \
    const unsigned long long bit__ = 1ull << (dev__ & 63);                                                \
    if (!(done__.load(std::memory_order_acquire) & bit__)) {                                              \
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));            \
      done__.fetch_or(bit__, std::memory_order_release);                                                  \
    }                                                                                                     \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---- dropout seed: an immediate, or (CUDA-graph safe) a device word that a replayed graph re-reads ------------
struct Seed {
  uint64_t v = 0;
  const uint64_t* p = nullptr;
  Seed() = default;
  Seed(uint64_t value) : v(value), p(nullptr) {}   // NOLINT: implicit on purpose (C ABI passes plain integers)
  Seed(uint64_t value, const uint64_t* ptr) : v(value), p(ptr) {}
};
__device__ __forceinline__ uint64_t seed_value(const Seed& s) { return s.p ? __ldg(reinterpret_cast<const unsigned long long*>(s.p)) : s.v; }

// ---- dropout bits: counter-based SplitMix64 ------------------------------------------------------
// One 64-bit word per group of four consecutive elements of a dropout site; element j of the group owns bits
// 16j..16j+15 as a 16-bit uniform.  word = mix64(key + (group+1) * gamma) is exactly SplitMix64's output `group+1`
// steps after state `key` (Steele, Lea, Flood 2014; mix64 = Stafford's variant 13, passes BigCrush on sequential
// counters); key = mix64(seed + site * gamma) gives every (seed, site) its own random offset into the 2^64 cycle.
// ~35 integer instructions per four elements.  (The first version used Philox4x32-10, ~120 instructions per four
// elements: in the 2048-wide FFN epilogue and the attention kernels the mask cost more than the arithmetic.)
// keep <=> u16 >= round(p * 65536): the drop probability is p quantised to 2^-16 (0.1 -> 0.100006).
constexpr uint64_t kGamma = 0x9E3779B97F4A7C15ull;
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t drop_key(uint64_t seed, uint32_t site) { return mix64(seed + (uint64_t)site * kGamma); }
__device__ __forceinline__ uint64_t drop_word(uint64_t key, uint64_t group) { return mix64(key + (group + 1) * kGamma); }
__device__ __forceinline__ uint32_t drop_threshold(float p) { return (uint32_t)__float2int_rn(p * 65536.0f); }

// Scaled keep mask of element `e` (0 or 1/(1-p)).
__device__ __forceinline__ float dropout_scale1(uint64_t seed, uint32_t site, uint64_t e, float p, float scale) {
  const uint64_t w = drop_word(drop_key(seed, site), e >> 2);
  const uint32_t u = (uint32_t)(w >> (16 * (uint32_t)(e & 3))) & 0xFFFFu;
  return u >= drop_threshold(p) ? scale : 0.0f;
}

// Four consecutive elements e..e+3.  Fast path when e is 4-aligned.  (seed, site, p) are loop-invariant at every
// call site, so the key and the threshold are hoisted out of the callers' loops by the compiler.
__device__ __forceinline__ void dropout_scale4(uint64_t seed, uint32_t site, uint64_t e, float p, float scale,
                                               float out[4]) {
  if ((e & 3) == 0) {
    const uint64_t w = drop_word(drop_key(seed, site), e >> 2);
    const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32), thr = drop_threshold(p);
    out[0] = (lo & 0xFFFFu) >= thr ? scale : 0.0f;
    out[1] = (lo >> 16) >= thr ? scale : 0.0f;
    out[2] = (hi & 0xFFFFu) >= thr ? scale : 0.0f;
    out[3] = (hi >> 16) >= thr ? scale : 0.0f;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = dropout_scale1(seed, site, e + j, p, scale);
  }
}

// Loop-hoisted form for 4-aligned element indices: the key and the threshold are computed once per thread.
struct DropCtx { uint64_t key; uint32_t thr; float scale; };
__device__ __forceinline__ DropCtx make_drop_ctx(uint64_t seed, uint32_t site, float p) {
  DropCtx c;
  c.key = drop_key(seed, site); c.thr = drop_threshold(p); c.scale = 1.0f / (1.0f - p);
  return c;
}
__device__ __forceinline__ void dropout_scale4_aligned(const DropCtx& c, uint64_t e, float out[4]) {
  const uint64_t w = drop_word(c.key, e >> 2);
  const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
  out[0] = (lo & 0xFFFFu) >= c.thr ? c.scale : 0.0f;
  out[1] = (lo >> 16) >= c.thr ? c.scale : 0.0f;
  out[2] = (hi & 0xFFFFu) >= c.thr ? c.scale : 0.0f;
  out[3] = (hi >> 16) >= c.thr ? c.scale : 0.0f;
}

// ---- activations ---------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float inv_sqrt_2pi = 0.39894228040143267794f;
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * inv_sqrt_2pi * expf(-0.5f * x * x);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
// Round-to-nearest (ties away from zero) of an fp32 value to TF32, as two integer instructions: add half a TF32 ulp to
// the magnitude bits and clear the 13 low mantissa bits -- the carry propagates into the exponent exactly as rounding
// requires.  Bit-identical to `cvt.rna.tf32.f32` for finite inputs (Inf stays Inf), which on sm_100a compiles to a
// five-to-six instruction sequence (VIADD, LOP3, FSETP, SEL, ...: r2 SASS of attention_bwd_mma_kernel<64> -- 45 % of its
// 26.6 M warp instructions were the conversions of the 3xTF32 operand splits).
__device__ __forceinline__ uint32_t tf32_rna_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

// 2^x in one MUFU instruction (rel. error ~2^-22).  The attention kernels keep their scores in log2 units (the 1/sqrt(hd)
// scale of Q carries a log2(e) factor) so that every probability is a single ex2: expf() is ~12 instructions, and the
// small-head kernels are bound by instruction issue (one or two exponentials per (query, key) pair of 45 - 110 instructions).
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float ex2_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case GANFFN_ACT_RELU: return v > 0.0f ? v : 0.0f;
    case GANFFN_ACT_GELU: return gelu_f(v);
    case GANFFN_ACT_SIGMOID: return sigmoid_f(v);
    default: return v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- asynchronous global -> shared copies (LDGSTS) ---------------------------------------------------------------
// The attention kernels stage a dialogue's Q/K/V/dO head slices (rows of 8..64 floats at a stride of B*3*d floats) in
// shared memory.  Through registers that is a loop of dependent LDG -> STS pairs, i.e. one exposed global-memory
// latency per iteration (r2 ncu of attention_fwd_small_kernel: 38 % of the stall cycles were long-scoreboard waits of
// the eight tile-load iterations); with cp.async every chunk of the tile is in flight at once.
__device__ __forceinline__ void cp_async_8(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- GEMM epilogue shared by the SIMT and tcgen05 engines --------------------------------------
// Backward-activation codes (dact)
enum { DACT_NONE = 0, DACT_NONZERO = 1, DACT_GELU = 2 };

struct Epilogue {
  const float* bias = nullptr;      // [N]
  const float* residual = nullptr;  // [M, ldr], added last
  int ldr = 0;
  float* pre = nullptr;             // [M, ldc] optional: value fed to act
  int act = GANFFN_ACT_NONE;
  int drop_before_act = 0;
  float p_drop = 0.0f;              // forward dropout (or backward mask regeneration for dact)
  Seed seed;
  uint32_t site = 0;
  // backward-through-activation: v *= f(dact_src[m,n])
  int dact = DACT_NONE;
  const float* dact_src = nullptr;  // [M, ldc]
  float dact_scale = 1.0f;          // DACT_NONZERO: multiply kept lanes by this (1/(1-p))
  float beta = 0.0f;                // C = beta*C + v
  // Gradient accumulation (wgrad): C += v with red.global.add, so split-K slices need no partial buffers and no
  // fold kernel; the tcgen05 engine can also emit rowsum[m] += sum_k A[m,k] (the bias gradient, since A = dY^T)
  // from the registers of its A producers.  The FFMA engine adds with red.global.add too (its split-K fold kernel
  // or its direct epilogue) and ignores `rowsum`.
  int atomic_acc = 0;
  float* rowsum = nullptr;
  // LayerNorm behind the epilogue (post-norm encoder layer: x = LN(residual + drop(linear(..)))).  When ln_out is set,
  // gemm() also writes ln_out[m, :] = LN(C[m, :]) * ln_gamma + ln_beta (row stride ldc, needs ldc == N).  The tcgen05
  // engine does it inside the GEMM kernel when one tile spans the row (N <= 128) or inside the split-K fold; every
  // other case runs the stand-alone LayerNorm kernel after the product.
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  float* ln_out = nullptr;
  // LayerNorm BACKWARD behind a data-gradient product: the product's result (residual included) is dy of a post-norm
  // LayerNorm x = LN(z).  When lnb_dz is set, gemm() also computes dz = dLN(dy; z, gamma) -> lnb_dz, the dropout-masked
  // branch gradient dz * mask(lnb_site) -> lnb_dzd (optional), and accumulates dgamma / dbeta / the sublayer's bias
  // gradient (column sums of dy * xhat, dy, dz * mask; optional).  The tcgen05 engine does it inside its split-K fold
  // (N <= 128); every other case runs layernorm_bwd() behind the product -- same results.
  const float* lnb_z = nullptr;
  const float* lnb_gamma = nullptr;
  float* lnb_dz = nullptr;
  float* lnb_dzd = nullptr;
  float* lnb_dgamma = nullptr;
  float* lnb_dbeta = nullptr;
  float* lnb_dbias = nullptr;
  float lnb_p = 0.0f;
  uint32_t lnb_site = 0;
  Seed lnb_seed;
};

// Applies the epilogue to the four accumulators of row m, columns n..n+3 (n % 4 == 0) and
// stores the valid ones.  N is the logical row length (also the dropout index stride).
__device__ __forceinline__ void epilogue_store4(const Epilogue& ep, float* C, int ldc, int M, int N, int m, int n,
                                                float v[4]) {
  if (m >= M || n >= N) return;
  const int nv = min(4, N - n);
  if (ep.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nv) v[j] += __ldg(ep.bias + n + j);
  }
  const bool drop = ep.p_drop > 0.0f;
  float msk[4] = {1.f, 1.f, 1.f, 1.f};
  if (drop) dropout_scale4(seed_value(ep.seed), ep.site, (uint64_t)m * (uint64_t)N + (uint64_t)n, ep.p_drop, 1.0f / (1.0f - ep.p_drop), msk);
  if (ep.dact == DACT_NONE) {
    if (drop && ep.drop_before_act) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= msk[j];
    }
    if (ep.pre) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nv) ep.pre[(size_t)m * ldc + n + j] = v[j];
    }
    if (ep.act != GANFFN_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], ep.act);
    }
    if (drop && !ep.drop_before_act) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= msk[j];
    }
  } else if (ep.dact == DACT_NONZERO) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nv) v[j] = (ep.dact_src[(size_t)m * ldc + n + j] != 0.0f) ? v[j] * ep.dact_scale : 0.0f;
  } else {  // DACT_GELU: v *= gelu'(pre) * dropmask
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nv) v[j] *= gelu_grad_f(ep.dact_src[(size_t)m * ldc + n + j]) * msk[j];
  }
  if (ep.residual) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nv) v[j] += ep.residual[(size_t)m * ep.ldr + n + j];
  }
  float* c = C + (size_t)m * ldc + n;
  if (ep.atomic_acc) {   // gradient accumulation: several backward passes may add to C concurrently (network lanes)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nv) atomicAdd(c + j, v[j]);
    return;
  }
  if (ep.beta != 0.0f) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nv) v[j] += ep.beta * c[j];
  }
  if (nv == 4 && ((((uintptr_t)c) & 15) == 0)) {
    *reinterpret_cast<float4*>(c) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nv) c[j] = v[j];
  }
}

// ---- GEMM front door (gemm.cu): C[M,N] = op(A)[M,K] * op(B)[K,N] with epilogue -----------------
//   transA = false: A stored [M, lda] (k contiguous);  true: A stored [K, lda] (m contiguous)
//   b_is_nk = true: B stored [N, ldb] (k contiguous, i.e. nn.Linear weight); false: stored [K, ldb]
// scratch is used for split-K partials (may be null => no split).
int gemm(const float* A, int lda, bool transA, const float* B, int ldb, bool b_is_nk, float* C, int ldc, int M, int N,
         int K, const Epilogue& ep, float* scratch, int64_t scratch_floats, cudaStream_t st);
int64_t gemm_scratch_floats(int M, int N, int K);
// true when gemm() would run this product on the tcgen05 engine (which honours Epilogue::rowsum)
bool gemm_uses_tc(const float* A, int lda, bool transA, const float* B, int ldb, bool b_is_nk, int ldc, int M, int N, int K);

}  // namespace ganffn
