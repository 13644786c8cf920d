// Self-attention core for whole dialogues: one CTA per (dialogue, head), S <= 110 so a full
// dialogue's K and V live in shared memory and every thread owns one query (forward) or one
// query then one key (backward).  Padded slots are real tokens here, exactly as in the
// reference, which never passes a key-padding mask.
#include "common.cuh"

namespace ganffn {
namespace {

constexpr int ATT_THREADS = 128;  // >= GANFFN_MAX_SEQ

template <int HD>
struct Vec {
  static constexpr int W = (HD % 4 == 0) ? 4 : (HD % 2 == 0) ? 2 : 1;
};

template <int HD>
__device__ __forceinline__ void load_row(const float* __restrict__ g, float* r) {
  constexpr int W = Vec<HD>::W;
  if (W == 4) {
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      float4 v = __ldg(reinterpret_cast<const float4*>(g) + c);
      r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
  } else if (W == 2) {
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) {
      float2 v = __ldg(reinterpret_cast<const float2*>(g) + c);
      r[2 * c] = v.x; r[2 * c + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < HD; ++c) r[c] = __ldg(g + c);
  }
}

template <int HD>
__device__ __forceinline__ void store_row(float* g, const float* r) {
  constexpr int W = Vec<HD>::W;
  if (W == 4) {
#pragma unroll
    for (int c = 0; c < HD / 4; ++c)
      reinterpret_cast<float4*>(g)[c] = make_float4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
  } else if (W == 2) {
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) reinterpret_cast<float2*>(g)[c] = make_float2(r[2 * c], r[2 * c + 1]);
  } else {
#pragma unroll
    for (int c = 0; c < HD; ++c) g[c] = r[c];
  }
}

// Cooperative copy of one head's [S, HD] slice (row stride `ld` floats in global) to smem [S][HD].
template <int HD>
__device__ __forceinline__ void load_tile(const float* __restrict__ g, int ld, float* s, int S) {
  for (int idx = threadIdx.x; idx < S * HD; idx += blockDim.x) {
    int r = idx / HD, c = idx % HD;
    s[idx] = __ldg(g + (size_t)r * ld + c);
  }
}

template <int HD>
__device__ __forceinline__ float dot_smem(const float* q, const float* __restrict__ krow) {
  float s = 0.f;
  constexpr int W = Vec<HD>::W;
  if (W == 4) {
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      float4 k = *reinterpret_cast<const float4*>(krow + 4 * c);
      s = fmaf(q[4 * c], k.x, s); s = fmaf(q[4 * c + 1], k.y, s);
      s = fmaf(q[4 * c + 2], k.z, s); s = fmaf(q[4 * c + 3], k.w, s);
    }
  } else if (W == 2) {
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) {
      float2 k = *reinterpret_cast<const float2*>(krow + 2 * c);
      s = fmaf(q[2 * c], k.x, s); s = fmaf(q[2 * c + 1], k.y, s);
    }
  } else {
#pragma unroll
    for (int c = 0; c < HD; ++c) s = fmaf(q[c], krow[c], s);
  }
  return s;
}

template <int HD>
__device__ __forceinline__ void axpy_smem(float a, const float* __restrict__ row, float* acc) {
  constexpr int W = Vec<HD>::W;
  if (W == 4) {
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      float4 v = *reinterpret_cast<const float4*>(row + 4 * c);
      acc[4 * c] = fmaf(a, v.x, acc[4 * c]); acc[4 * c + 1] = fmaf(a, v.y, acc[4 * c + 1]);
      acc[4 * c + 2] = fmaf(a, v.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(a, v.w, acc[4 * c + 3]);
    }
  } else if (W == 2) {
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) {
      float2 v = *reinterpret_cast<const float2*>(row + 2 * c);
      acc[2 * c] = fmaf(a, v.x, acc[2 * c]); acc[2 * c + 1] = fmaf(a, v.y, acc[2 * c + 1]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = fmaf(a, row[c], acc[c]);
  }
}

// ---- forward ----------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attention_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ o,
                                                                    float* __restrict__ lse, int S, int B, int d,
                                                                    int nhead, float p_drop, uint64_t seed,
                                                                    uint32_t site) {
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;            // [S][HD]
  float* Vs = smem + S * HD;   // [S][HD]
  const int b = blockIdx.x / nhead, h = blockIdx.x % nhead;
  const int ld = B * 3 * d;    // row stride between consecutive s for fixed b
  const float* base = qkv + (size_t)b * 3 * d + (size_t)h * HD;
  load_tile<HD>(base + d, ld, Ks, S);
  load_tile<HD>(base + 2 * d, ld, Vs, S);
  __syncthreads();

  const int i = threadIdx.x;
  if (i >= S) return;
  const float scale = rsqrtf((float)HD);
  float q[HD], acc[HD];
  load_row<HD>(base + (size_t)i * ld, q);
#pragma unroll
  for (int c = 0; c < HD; ++c) { q[c] *= scale; acc[c] = 0.f; }

  const bool drop = p_drop > 0.f;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  const int S4 = (S + 3) & ~3;
  const uint64_t ebase = ((uint64_t)blockIdx.x * S + i) * S4;

  float mrun = -INFINITY, lrun = 0.f;
  for (int j0 = 0; j0 < S; j0 += 8) {
    float s[8];
    float cmax = -INFINITY;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u;
      s[u] = (j < S) ? dot_smem<HD>(q, Ks + j * HD) : -INFINITY;
      cmax = fmaxf(cmax, s[u]);
    }
    const float mnew = fmaxf(mrun, cmax);
    const float corr = __expf(mrun - mnew);  // 0 on the first chunk
    lrun *= corr;
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] *= corr;
    float msk[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if (drop) {
      dropout_scale4(seed, site, ebase + j0, p_drop, dscale, msk);
      dropout_scale4(seed, site, ebase + j0 + 4, p_drop, dscale, msk + 4);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u;
      if (j < S) {
        const float pj = expf(s[u] - mnew);
        lrun += pj;
        axpy_smem<HD>(pj * msk[u], Vs + j * HD, acc);
      }
    }
    mrun = mnew;
  }
  const float inv = 1.f / lrun;
#pragma unroll
  for (int c = 0; c < HD; ++c) acc[c] *= inv;
  store_row<HD>(o + ((size_t)i * B + b) * d + (size_t)h * HD, acc);
  lse[(size_t)blockIdx.x * S + i] = mrun + logf(lrun);
}

// ---- backward ---------------------------------------------------------------------------------
// Phase A (thread = query i): recompute P row, dS row; dQ_i; park Pd and dS in smem.
// Phase B (thread = key j):   dV_j = sum_i Pd[i][j] dO_i ; dK_j = sum_i dS[i][j] Q_i.
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attention_bwd_kernel(
    const float* __restrict__ qkv, const float* __restrict__ o, const float* __restrict__ lse,
    const float* __restrict__ d_o, float* __restrict__ dqkv, int S, int B, int d, int nhead, float p_drop,
    uint64_t seed, uint32_t site) {
  extern __shared__ __align__(16) float smem[];
  const int SP = S | 1;  // odd row stride for the S x S matrices
  float* Qs = smem;                 // [S][HD]
  float* Ks = Qs + S * HD;
  float* Vs = Ks + S * HD;
  float* dOs = Vs + S * HD;
  float* Ps = dOs + S * HD;         // [S][SP]  dropped, scaled probabilities
  float* dSs = Ps + S * SP;         // [S][SP]
  float* Dv = dSs + S * SP;         // [S] rowsum(dO * O)

  const int b = blockIdx.x / nhead, h = blockIdx.x % nhead;
  const int ld = B * 3 * d;
  const int ldo = B * d;
  const float* base = qkv + (size_t)b * 3 * d + (size_t)h * HD;
  const float* obase = o + (size_t)b * d + (size_t)h * HD;
  const float* dobase = d_o + (size_t)b * d + (size_t)h * HD;
  load_tile<HD>(base, ld, Qs, S);
  load_tile<HD>(base + d, ld, Ks, S);
  load_tile<HD>(base + 2 * d, ld, Vs, S);
  load_tile<HD>(dobase, ldo, dOs, S);
  __syncthreads();
  {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = w; r < S; r += ATT_THREADS / 32) {
      float acc = 0.f;
      for (int c = lane; c < HD; c += 32) acc += dOs[r * HD + c] * __ldg(obase + (size_t)r * ldo + c);
      acc = warp_sum(acc);
      if (lane == 0) Dv[r] = acc;
    }
  }
  __syncthreads();

  const float scale = rsqrtf((float)HD);
  const bool drop = p_drop > 0.f;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  const int S4 = (S + 3) & ~3;
  const int i = threadIdx.x;

  if (i < S) {
    float q[HD], g[HD], dq[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) { q[c] = Qs[i * HD + c] * scale; g[c] = dOs[i * HD + c]; dq[c] = 0.f; }
    const float li = lse[(size_t)blockIdx.x * S + i];
    const float Di = Dv[i];
    const uint64_t ebase = ((uint64_t)blockIdx.x * S + i) * S4;
    for (int j0 = 0; j0 < S; j0 += 4) {
      float msk[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop) dropout_scale4(seed, site, ebase + j0, p_drop, dscale, msk);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        if (j < S) {
          const float s = dot_smem<HD>(q, Ks + j * HD);
          const float pj = expf(s - li);
          const float dpd = dot_smem<HD>(g, Vs + j * HD);
          const float ds = pj * (dpd * msk[u] - Di);
          Ps[i * SP + j] = pj * msk[u];
          dSs[i * SP + j] = ds * scale;
          axpy_smem<HD>(ds, Ks + j * HD, dq);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) dq[c] *= scale;
    store_row<HD>(dqkv + ((size_t)i * B + b) * 3 * d + (size_t)h * HD, dq);
  }
  __syncthreads();
  if (i < S) {
    const int j = i;
    float dk[HD], dv[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
    for (int r = 0; r < S; ++r) {
      const float pd = Ps[r * SP + j];
      const float ds = dSs[r * SP + j];
      axpy_smem<HD>(pd, dOs + r * HD, dv);
      axpy_smem<HD>(ds, Qs + r * HD, dk);
    }
    float* out = dqkv + ((size_t)j * B + b) * 3 * d + (size_t)h * HD;
    store_row<HD>(out + d, dk);
    store_row<HD>(out + 2 * d, dv);
  }
}

template <int HD>
int launch_fwd(const float* qkv, float* o, float* lse, int S, int B, int d, int nhead, float p, uint64_t seed, int site,
               cudaStream_t st) {
  const size_t smem = (size_t)2 * S * HD * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attention_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(2 * GANFFN_MAX_SEQ * HD * sizeof(float)));
    attr_done = true;
  }
  attention_fwd_kernel<HD><<<B * nhead, ATT_THREADS, smem, st>>>(qkv, o, lse, S, B, d, nhead, p, seed, (uint32_t)site);
  GANFFN_LAUNCHED("attention_fwd_kernel");
  return GANFFN_OK;
}

template <int HD>
int launch_bwd(const float* qkv, const float* o, const float* lse, const float* d_o, float* dqkv, int S, int B, int d,
               int nhead, float p, uint64_t seed, int site, cudaStream_t st) {
  auto bytes = [](int s) { return ((size_t)4 * s * HD + (size_t)2 * s * (s | 1) + s) * sizeof(float); };
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attention_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)bytes(GANFFN_MAX_SEQ));
    attr_done = true;
  }
  attention_bwd_kernel<HD><<<B * nhead, ATT_THREADS, bytes(S), st>>>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed,
                                                                    (uint32_t)site);
  GANFFN_LAUNCHED("attention_bwd_kernel");
  return GANFFN_OK;
}

}  // namespace

int attention_fwd(const float* qkv, float* o, float* lse, int S, int B, int d, int nhead, float p, uint64_t seed,
                  int site, cudaStream_t st) {
  GANFFN_CHECK_ARG(S >= 1 && S <= GANFFN_MAX_SEQ, "attention: seq_len %d outside [1,%d] (model.py:1179)", S,
                   GANFFN_MAX_SEQ);
  GANFFN_CHECK_ARG(B >= 1 && nhead >= 1 && d % nhead == 0, "attention: d=%d not divisible by nhead=%d", d, nhead);
  GANFFN_CHECK_ARG(p >= 0.f && p < 1.f, "attention: dropout p=%f", p);
  switch (d / nhead) {
    case 8: return launch_fwd<8>(qkv, o, lse, S, B, d, nhead, p, seed, site, st);
    case 10: return launch_fwd<10>(qkv, o, lse, S, B, d, nhead, p, seed, site, st);
    case 16: return launch_fwd<16>(qkv, o, lse, S, B, d, nhead, p, seed, site, st);
    case 32: return launch_fwd<32>(qkv, o, lse, S, B, d, nhead, p, seed, site, st);
    case 64: return launch_fwd<64>(qkv, o, lse, S, B, d, nhead, p, seed, site, st);
  }
  set_error("attention: unsupported head_dim %d (supported 8,10,16,32,64)", d / nhead);
  return GANFFN_ERR_ARG;
}

int attention_bwd(const float* qkv, const float* o, const float* lse, const float* d_o, float* dqkv, int S, int B, int d,
                  int nhead, float p, uint64_t seed, int site, cudaStream_t st) {
  GANFFN_CHECK_ARG(S >= 1 && S <= GANFFN_MAX_SEQ, "attention: seq_len %d outside [1,%d] (model.py:1179)", S,
                   GANFFN_MAX_SEQ);
  GANFFN_CHECK_ARG(B >= 1 && nhead >= 1 && d % nhead == 0, "attention: d=%d not divisible by nhead=%d", d, nhead);
  switch (d / nhead) {
    case 8: return launch_bwd<8>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed, site, st);
    case 10: return launch_bwd<10>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed, site, st);
    case 16: return launch_bwd<16>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed, site, st);
    case 32: return launch_bwd<32>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed, site, st);
    case 64: return launch_bwd<64>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed, site, st);
  }
  set_error("attention: unsupported head_dim %d (supported 8,10,16,32,64)", d / nhead);
  return GANFFN_ERR_ARG;
}

}  // namespace ganffn
