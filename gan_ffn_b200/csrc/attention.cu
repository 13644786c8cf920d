// Self-attention core for whole dialogues: one CTA per (dialogue, head).  S <= 110, so a dialogue's Q/K/V (and in
// the backward pass the S x S probability / score-gradient matrices) live in shared memory.  Padded slots are real
// tokens here, exactly as in the reference, which never passes a key-padding mask.
//
// Work decomposition: lane = one row (query, or key in the dK/dV sweeps), warps come in four groups.  In the sweeps
// over keys (scores, dP) group g owns a quarter of the keys; in the sweeps that produce HD-wide rows (P V, dQ, dK,
// dV) group g owns a quarter of the output columns, so nothing is reduced across threads.  Every shared-memory
// operand read is either warp-uniform (broadcast of a K/V/Q/dO row) or walks consecutive floats of an odd-stride
// S x S row: one wavefront per instruction.  (A first version with one thread per query and no groups had 3 warps
// per CTA and ran latency-bound: 26/100 us fwd/bwd per d=100 layer at S=94, B=32.)
#include <stdlib.h>
#include "kernels.h"

namespace ganffn {
namespace {

constexpr int NG = 4;                                // warp groups
// resident CTAs per SM the small-head kernels are compiled for (register cap 65536 / (128 * n)); build-time tuning knobs
#ifndef GANFFN_ATT_FWD_MINB
#define GANFFN_ATT_FWD_MINB 5
#endif
#ifndef GANFFN_ATT_BWD_MINB
#define GANFFN_ATT_BWD_MINB 5
#endif
constexpr int MAX_RW = (GANFFN_MAX_SEQ + 31) / 32;   // row warps per group
constexpr int KPG = ((GANFFN_MAX_SEQ + 4 * NG - 1) / (4 * NG)) * 4;   // max keys per group (multiple of 4)

template <int HD>
struct Cfg {
  static constexpr int W = (HD % 4 == 0) ? 4 : (HD % 2 == 0) ? 2 : 1;   // vector width of a row
  static constexpr int CW = (HD + NG - 1) / NG;                          // output columns per group
  static constexpr int CWV = (CW % 4 == 0) ? 4 : 1;
};

// Cooperative asynchronous copy (cp.async, see common.cuh) of one head's [S, HD] slice (row stride `ld` floats in
// global) to smem [S][HD]; the caller runs cp_async_wait_all() before the __syncthreads() that publishes the tile.
template <int HD>
__device__ __forceinline__ void load_tile(const float* __restrict__ g, int ld, float* s, int S) {
  constexpr int W = Cfg<HD>::W, CPR = HD / W;
  for (int idx = threadIdx.x; idx < S * CPR; idx += blockDim.x) {
    const int r = idx / CPR, c = (idx % CPR) * W;
    const float* src = g + (size_t)r * ld + c;
    float* dst = s + r * HD + c;
    if (W == 4) cp_async_16(dst, src);
    else if (W == 2) cp_async_8(dst, src);
    else cp_async_4(dst, src);
  }
}

template <int HD>
__device__ __forceinline__ void row_to_regs(const float* __restrict__ s, float* r, float mul) {
  constexpr int W = Cfg<HD>::W;
  if (W == 4) {
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(s + 4 * c);
      r[4 * c] = v.x * mul; r[4 * c + 1] = v.y * mul; r[4 * c + 2] = v.z * mul; r[4 * c + 3] = v.w * mul;
    }
  } else if (W == 2) {
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) {
      const float2 v = *reinterpret_cast<const float2*>(s + 2 * c);
      r[2 * c] = v.x * mul; r[2 * c + 1] = v.y * mul;
    }
  } else {
#pragma unroll
    for (int c = 0; c < HD; ++c) r[c] = s[c] * mul;
  }
}

// Packed fp32 FMA (sm_100 FFMA2: two fused multiply-adds per instruction).  These kernels are bound by instruction
// issue, not by the FMA pipe (r2 ncu: fma pipe 15 % busy, issue slots 37 % with every warp waiting on its own
// dependent chain), so halving the FMA instruction count is worth more than any reordering.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// q (registers) . row (warp-uniform smem address)
template <int HD>
__device__ __forceinline__ float dot_smem(const float* q, const float* __restrict__ krow) {
  float s = 0.f;
  constexpr int W = Cfg<HD>::W;
  if (W == 4) {
    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      float4 k = *reinterpret_cast<const float4*>(krow + 4 * c);
      s2 = ffma2(make_float2(q[4 * c], q[4 * c + 1]), make_float2(k.x, k.y), s2);
      s2 = ffma2(make_float2(q[4 * c + 2], q[4 * c + 3]), make_float2(k.z, k.w), s2);
    }
    s = s2.x + s2.y;
  } else if (W == 2) {
    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) {
      float2 k = *reinterpret_cast<const float2*>(krow + 2 * c);
      s2 = ffma2(make_float2(q[2 * c], q[2 * c + 1]), k, s2);
    }
    s = s2.x + s2.y;
  } else {
#pragma unroll
    for (int c = 0; c < HD; ++c) s = fmaf(q[c], krow[c], s);
  }
  return s;
}

// acc[0..CW) += a * row[c0 .. c0+CW)  (warp-uniform row; columns past HD are skipped)
template <int HD>
__device__ __forceinline__ void axpy_cols(float a, const float* __restrict__ row, int c0, float* acc) {
  constexpr int CW = Cfg<HD>::CW;
  if (Cfg<HD>::CWV == 4) {
#pragma unroll
    for (int c = 0; c < CW / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(row + c0 + 4 * c);
      acc[4 * c] = fmaf(a, v.x, acc[4 * c]); acc[4 * c + 1] = fmaf(a, v.y, acc[4 * c + 1]);
      acc[4 * c + 2] = fmaf(a, v.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(a, v.w, acc[4 * c + 3]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < CW; ++c)
      if (c0 + c < HD) acc[c] = fmaf(a, row[c0 + c], acc[c]);
  }
}

template <int HD>
__device__ __forceinline__ void store_cols(float* g, int c0, const float* acc, float mul) {
  constexpr int CW = Cfg<HD>::CW;
  if (Cfg<HD>::CWV == 4) {
#pragma unroll
    for (int c = 0; c < CW / 4; ++c)
      *reinterpret_cast<float4*>(g + c0 + 4 * c) =
          make_float4(acc[4 * c] * mul, acc[4 * c + 1] * mul, acc[4 * c + 2] * mul, acc[4 * c + 3] * mul);
  } else {
#pragma unroll
    for (int c = 0; c < CW; ++c)
      if (c0 + c < HD) g[c0 + c] = acc[c] * mul;
  }
}

// Sweep producing an HD-wide row per lane: acc = sum_r M[r][row] or M[row][r] (see callers) * X[r][c0..c0+CW).
// `m` points at the lane's first coefficient, `mstride` is the step between consecutive r.
template <int HD>
__device__ __forceinline__ void sweep_cols(const float* __restrict__ m, int mstride, const float* __restrict__ X, int S,
                                           int c0, float* acc) {
#pragma unroll
  for (int c = 0; c < Cfg<HD>::CW; ++c) acc[c] = 0.f;
  int r = 0;
  for (; r + 4 <= S; r += 4) {
    const float a0 = m[(r + 0) * mstride], a1 = m[(r + 1) * mstride], a2 = m[(r + 2) * mstride], a3 = m[(r + 3) * mstride];
    axpy_cols<HD>(a0, X + (r + 0) * HD, c0, acc);
    axpy_cols<HD>(a1, X + (r + 1) * HD, c0, acc);
    axpy_cols<HD>(a2, X + (r + 2) * HD, c0, acc);
    axpy_cols<HD>(a3, X + (r + 3) * HD, c0, acc);
  }
  for (; r < S; ++r) axpy_cols<HD>(m[r * mstride], X + r * HD, c0, acc);
}

// ---- forward ----------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(NG * MAX_RW * 32) attention_fwd_kernel(const float* __restrict__ qkv,
                                                                         float* __restrict__ o, float* __restrict__ lse,
                                                                         int S, int B, int d, int nhead, float p_drop,
                                                                         const Seed seed_ref, uint32_t site) {
  extern __shared__ __align__(16) float smem[];
  const int SP = S | 1;
  float* Qs = smem;              // [S][HD]
  float* Ks = Qs + S * HD;
  float* Vs = Ks + S * HD;
  float* Ps = Vs + S * HD;       // [S][SP] dropped probabilities, not yet normalised
  float* red = Ps + S * SP;      // [NG][S] partial row max, then partial row sums
  const int b = blockIdx.x / nhead, h = blockIdx.x % nhead;
  const int ld = B * 3 * d;      // row stride between consecutive s for fixed b
  const float* base = qkv + (size_t)b * 3 * d + (size_t)h * HD;
  load_tile<HD>(base, ld, Qs, S);
  load_tile<HD>(base + d, ld, Ks, S);
  load_tile<HD>(base + 2 * d, ld, Vs, S);
  cp_async_wait_all();
  __syncthreads();

  const int nrw = (S + 31) >> 5;                       // row warps per group
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = warp / nrw, i = (warp % nrw) * 32 + lane;
  const bool active = i < S;
  const int ir = active ? i : S - 1;
  const int kpg = (((S + 3) >> 2) + NG - 1) / NG * 4;  // keys per group, multiple of 4
  const int j_beg = g * kpg, j_end = min(S, j_beg + kpg);
  const float scale = rsqrtf((float)HD);

  float s[KPG];
  float mx = -INFINITY;
  {
    float q[HD];
    row_to_regs<HD>(Qs + ir * HD, q, scale);
#pragma unroll
    for (int k = 0; k < KPG; ++k) {
      const int j = j_beg + k;
      s[k] = (j < j_end) ? dot_smem<HD>(q, Ks + j * HD) : -INFINITY;
      mx = fmaxf(mx, s[k]);
    }
  }
  red[g * S + ir] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[ir], red[S + ir]), fmaxf(red[2 * S + ir], red[3 * S + ir]));
  __syncthreads();
  const bool drop = p_drop > 0.f;
  const uint64_t seed = drop ? seed_value(seed_ref) : 0ull;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  const int S4 = (S + 3) & ~3;
  const uint64_t ebase = ((uint64_t)blockIdx.x * S + ir) * S4;
  float lsum = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < KPG; k4 += 4) {
    const int j0 = j_beg + k4;
    if (j0 < j_end) {
      float msk[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop) dropout_scale4(seed, site, ebase + j0, p_drop, dscale, msk);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + u < j_end) {
          const float pj = expf(s[k4 + u] - mx);
          lsum += pj;
          if (active) Ps[i * SP + j0 + u] = pj * msk[u];
        }
      }
    }
  }
  red[g * S + ir] = lsum;
  __syncthreads();
  lsum = (red[ir] + red[S + ir]) + (red[2 * S + ir] + red[3 * S + ir]);

  // O_i[c0 .. c0+CW) = sum_j P_ij V_j[c0 ..): group g now owns a quarter of the columns
  constexpr int CW = Cfg<HD>::CW;
  const int c0 = g * CW;
  float acc[CW];
  sweep_cols<HD>(Ps + ir * SP, 1, Vs, S, c0, acc);
  if (active) {
    store_cols<HD>(o + ((size_t)i * B + b) * d + (size_t)h * HD, c0, acc, 1.f / lsum);
    if (g == 0) lse[(size_t)blockIdx.x * S + i] = mx + logf(lsum);
  }
}

// ---- backward ---------------------------------------------------------------------------------
//   lane = query i, group = key range:   P_ij = exp(q_i.k_j - lse_i);  dP_ij = dO_i.v_j;
//                                        Ps = P m,  dSs = P (dP m - D_i) scale
//   lane = query i, group = columns:     dQ_i = sum_j dSs_ij k_j
//   lane = key j,   group = columns:     dV_j = sum_i Ps_ij dO_i ;  dK_j = sum_i dSs_ij q_i
template <int HD>
__global__ void __launch_bounds__(NG * MAX_RW * 32) attention_bwd_kernel(
    const float* __restrict__ qkv, const float* __restrict__ o, const float* __restrict__ lse,
    const float* __restrict__ d_o, float* __restrict__ dqkv, int S, int B, int d, int nhead, float p_drop,
    const Seed seed_ref, uint32_t site) {
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
  cp_async_wait_all();
  __syncthreads();
  {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = w; r < S; r += nw) {
      float acc = 0.f;
      for (int c = lane; c < HD; c += 32) acc += dOs[r * HD + c] * __ldg(obase + (size_t)r * ldo + c);
      acc = warp_sum(acc);
      if (lane == 0) Dv[r] = acc;
    }
  }
  __syncthreads();

  const int nrw = (S + 31) >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = warp / nrw, i = (warp % nrw) * 32 + lane;
  const bool active = i < S;
  const int ir = active ? i : S - 1;
  const int kpg = (((S + 3) >> 2) + NG - 1) / NG * 4;
  const int j_beg = g * kpg, j_end = min(S, j_beg + kpg);
  const float scale = rsqrtf((float)HD);
  const bool drop = p_drop > 0.f;
  const uint64_t seed = drop ? seed_value(seed_ref) : 0ull;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  const int S4 = (S + 3) & ~3;
  const uint64_t ebase = ((uint64_t)blockIdx.x * S + ir) * S4;

  if (active) {
    float r[HD];
    const float li = lse[(size_t)blockIdx.x * S + i];
    row_to_regs<HD>(Qs + i * HD, r, scale);
    for (int j = j_beg; j < j_end; ++j) Ps[i * SP + j] = expf(dot_smem<HD>(r, Ks + j * HD) - li);
    row_to_regs<HD>(dOs + i * HD, r, 1.f);
    const float Di = Dv[i];
    for (int j0 = j_beg; j0 < j_end; j0 += 4) {
      float msk[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop) dropout_scale4(seed, site, ebase + j0, p_drop, dscale, msk);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        if (j < j_end) {
          const float pj = Ps[i * SP + j];
          const float dpd = dot_smem<HD>(r, Vs + j * HD);
          Ps[i * SP + j] = pj * msk[u];
          dSs[i * SP + j] = pj * (dpd * msk[u] - Di) * scale;
        }
      }
    }
  }
  __syncthreads();

  constexpr int CW = Cfg<HD>::CW;
  const int c0 = g * CW;
  float acc[CW];
  float* out = dqkv + ((size_t)ir * B + b) * 3 * d + (size_t)h * HD;
  sweep_cols<HD>(dSs + ir * SP, 1, Ks, S, c0, acc);        // dQ_i (row i of dSs)
  if (active) store_cols<HD>(out, c0, acc, 1.f);
  sweep_cols<HD>(Ps + ir, SP, dOs, S, c0, acc);            // dV_j (column j = ir of Ps)
  if (active) store_cols<HD>(out + 2 * d, c0, acc, 1.f);
  sweep_cols<HD>(dSs + ir, SP, Qs, S, c0, acc);            // dK_j (column j of dSs)
  if (active) store_cols<HD>(out + d, c0, acc, 1.f);
}


// ================================================================================================================
// Small head_dim (<= 16: the five d=100 networks, head_dim 10).  Same decomposition -- lane = query row, four warp
// groups split the keys -- but the probabilities never leave the registers of the thread that computed them: each
// group accumulates P V (and dQ in the backward pass) over its own quarter of the keys and the four partial rows
// are folded through shared memory.  ncu on the first version (S x S matrix in smem, column-split P V sweeps): 11.7 M
// warp instructions forward / 23.8 M backward per d=100 layer at S=94 B=32, issue-bound at ~50 % issue utilisation
// with two CTAs per SM; this version executes about half of that and needs no S x S matrix in the forward pass.
// ================================================================================================================
template <int HD>
__device__ __forceinline__ void axpy_row(float a, const float* __restrict__ row, float* acc) {
  constexpr int W = Cfg<HD>::W;
  const float2 a2 = make_float2(a, a);
  if (W == 4) {
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(row + 4 * c);
      const float2 r0 = ffma2(a2, make_float2(v.x, v.y), make_float2(acc[4 * c], acc[4 * c + 1]));
      const float2 r1 = ffma2(a2, make_float2(v.z, v.w), make_float2(acc[4 * c + 2], acc[4 * c + 3]));
      acc[4 * c] = r0.x; acc[4 * c + 1] = r0.y; acc[4 * c + 2] = r1.x; acc[4 * c + 3] = r1.y;
    }
  } else if (W == 2) {
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) {
      const float2 v = *reinterpret_cast<const float2*>(row + 2 * c);
      const float2 r = ffma2(a2, v, make_float2(acc[2 * c], acc[2 * c + 1]));
      acc[2 * c] = r.x; acc[2 * c + 1] = r.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = fmaf(a, row[c], acc[c]);
  }
}

// row of HD floats in global memory (8-byte aligned when HD is even, 16-byte when HD % 4 == 0)
template <int HD>
__device__ __forceinline__ void store_row(float* g, const float* acc, float mul) {
  constexpr int W = Cfg<HD>::W;
  if (W == 4) {
#pragma unroll
    for (int c = 0; c < HD / 4; ++c)
      *reinterpret_cast<float4*>(g + 4 * c) = make_float4(acc[4 * c] * mul, acc[4 * c + 1] * mul, acc[4 * c + 2] * mul, acc[4 * c + 3] * mul);
  } else if (W == 2) {
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) *reinterpret_cast<float2*>(g + 2 * c) = make_float2(acc[2 * c] * mul, acc[2 * c + 1] * mul);
  } else {
#pragma unroll
    for (int c = 0; c < HD; ++c) g[c] = acc[c] * mul;
  }
}

// q (registers) . row (registers)
template <int HD>
__device__ __forceinline__ float dot_regs(const float* a, const float* b) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < HD; ++c) s = fmaf(a[c], b[c], s);
  return s;
}

// One CTA = (dialogue, head, block of 32 rows), four warps = four groups.  r2: the (dialogue, head) CTAs of the first
// version held 384 threads / 110 KB and ran two per SM, so the 320 CTAs of an S=94, B=32 layer took two waves on 296
// slots (the second with 24 CTAs): 24.7 / 37.4 us for 5.9 / 12 us of issue time.  With 32-row CTAs of 128 threads
// (7.5 - 16 KB of shared memory) 960 CTAs are all resident at once and nothing is quantised.
template <int HD>
__global__ void __launch_bounds__(NG * 32, GANFFN_ATT_FWD_MINB) attention_fwd_small_kernel(const float* __restrict__ qkv,
                                                                        float* __restrict__ o, float* __restrict__ lse,
                                                                        int S, int B, int d, int nhead, float p_drop,
                                                                        const Seed seed_ref, uint32_t site) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PW = HD | 1;         // odd row stride of the partial rows (lane = row: conflict-free)
  float* Ks = smem;                  // [S][HD]
  float* Vs = Ks + S * HD;
  float* redm = Vs + S * HD;         // [NG][32] partial row max
  float* reds = redm + NG * 32;      // [NG][32] partial row sums
  float* part = reds + NG * 32;      // [NG-1][32][PW] partial P V rows of groups 1..NG-1
  const int nrw = (S + 31) >> 5;
  const int bh = blockIdx.x / nrw, rw = blockIdx.x % nrw;
  const int b = bh / nhead, h = bh % nhead;
  const int ld = B * 3 * d;
  const float* base = qkv + (size_t)b * 3 * d + (size_t)h * HD;
  load_tile<HD>(base + d, ld, Ks, S);
  load_tile<HD>(base + 2 * d, ld, Vs, S);

  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = rw * 32 + lane;
  const bool active = i < S;
  const int ir = active ? i : S - 1;
  const int kpg = (((S + 3) >> 2) + NG - 1) / NG * 4;
  const int j_beg = g * kpg, j_end = min(S, j_beg + kpg);
  const float scale = rsqrtf((float)HD) * kLog2e;      // scores in log2 units: p = 2^(s - max)
  float q[HD];
  row_to_regs<HD>(base + (size_t)ir * ld, q, scale);   // the thread's own query row straight from global memory
  cp_async_wait_all();
  __syncthreads();

  float s[KPG];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < KPG; ++k) {
    const int j = j_beg + k;
    s[k] = (j < j_end) ? dot_smem<HD>(q, Ks + j * HD) : -INFINITY;
    mx = fmaxf(mx, s[k]);
  }
  redm[g * 32 + lane] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(redm[lane], redm[32 + lane]), fmaxf(redm[64 + lane], redm[96 + lane]));

  const bool drop = p_drop > 0.f;
  const uint64_t seed = drop ? seed_value(seed_ref) : 0ull;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  const int S4 = (S + 3) & ~3;
  const uint64_t ebase = ((uint64_t)bh * S + ir) * S4;
  float lsum = 0.f;
  float acc[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) acc[c] = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < KPG; k4 += 4) {
    const int j0 = j_beg + k4;
    if (j0 < j_end) {
      float msk[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop) dropout_scale4(seed, site, ebase + j0, p_drop, dscale, msk);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + u < j_end) {
          const float pj = ex2_f(s[k4 + u] - mx);
          lsum += pj;
          axpy_row<HD>(pj * msk[u], Vs + (j0 + u) * HD, acc);
        }
      }
    }
  }
  reds[g * 32 + lane] = lsum;
  if (g > 0) {
    float* pr = part + ((size_t)(g - 1) * 32 + lane) * PW;
#pragma unroll
    for (int c = 0; c < HD; ++c) pr[c] = acc[c];
  }
  __syncthreads();
  if (g == 0 && active) {
    lsum = (reds[lane] + reds[32 + lane]) + (reds[64 + lane] + reds[96 + lane]);
#pragma unroll
    for (int gg = 0; gg < NG - 1; ++gg) {
      const float* pr = part + ((size_t)gg * 32 + lane) * PW;
#pragma unroll
      for (int c = 0; c < HD; ++c) acc[c] += pr[c];
    }
    store_row<HD>(o + ((size_t)i * B + b) * d + (size_t)h * HD, acc, 1.f / lsum);
    lse[(size_t)bh * S + i] = mx * kLn2 + logf(lsum);   // natural-log lse (the C ABI's contract)
  }
}

// Backward, same CTA shape, no S x S matrix anywhere: the probabilities are recomputed in both sweeps.
//   sweep A  lane = query i of the CTA's row block, group = key range:
//              P_ij = exp(q_i.k_j - lse_i), dP_ij = dO_i.v_j, dS_ij = P_ij (dP_ij m_ij - D_i) scale;  dQ_i += dS_ij k_j
//   sweep B  lane = key j of the CTA's row block, group = query range: the same P_ij / dS_ij from the key's side;
//              dV_j += P_ij m_ij dO_i,  dK_j += dS_ij q_i
// Recomputing costs ~45 % more instructions than reading an S x S matrix back, but the first version's 110 KB of shared
// memory per (dialogue, head) CTA is what capped it at two CTAs per SM (see the forward kernel).  Dropout bits in
// sweep B: one 64-bit word covers four consecutive keys of one query, i.e. four neighbouring lanes; for a chunk of four
// queries each lane generates the word of query (lane & 3) of its key quad and the quad exchanges them by shuffle.
template <int HD>
__global__ void __launch_bounds__(NG * 32, GANFFN_ATT_BWD_MINB) attention_bwd_small_kernel(
    const float* __restrict__ qkv, const float* __restrict__ o, const float* __restrict__ lse,
    const float* __restrict__ d_o, float* __restrict__ dqkv, int S, int B, int d, int nhead, float p_drop,
    const Seed seed_ref, uint32_t site) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PW = (2 * HD) | 1;
  float* Qs = smem;                 // [S][HD]
  float* Ks = Qs + S * HD;
  float* Vs = Ks + S * HD;
  float* dOs = Vs + S * HD;
  float* Ls = dOs + S * HD;         // [S] row log-sum-exp
  float* Ds = Ls + S;               // [S] rowsum(dO * O)
  float* part = Ds + S;             // [NG-1][32][PW] partial rows: dQ (HD wide), later dV | dK (2 HD wide)

  const int nrw = (S + 31) >> 5;
  const int bh = blockIdx.x / nrw, rw = blockIdx.x % nrw;
  const int b = bh / nhead, h = bh % nhead;
  const int ld = B * 3 * d;
  const int ldo = B * d;
  const float* base = qkv + (size_t)b * 3 * d + (size_t)h * HD;
  const float* obase = o + (size_t)b * d + (size_t)h * HD;
  const float* dobase = d_o + (size_t)b * d + (size_t)h * HD;
  load_tile<HD>(base, ld, Qs, S);
  load_tile<HD>(base + d, ld, Ks, S);
  load_tile<HD>(base + 2 * d, ld, Vs, S);
  load_tile<HD>(dobase, ldo, dOs, S);
  for (int r = threadIdx.x; r < S; r += blockDim.x) {
    float orow[HD], drow[HD];
    row_to_regs<HD>(obase + (size_t)r * ldo, orow, 1.f);
    row_to_regs<HD>(dobase + (size_t)r * ldo, drow, 1.f);
    Ds[r] = dot_regs<HD>(drow, orow);
    Ls[r] = lse[(size_t)bh * S + r] * kLog2e;   // log2 units, like the recomputed scores below
  }
  cp_async_wait_all();
  __syncthreads();

  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = rw * 32 + lane;       // the thread's row: a query in sweep A, a key in sweep B
  const bool active = i < S;
  const int ir = active ? i : S - 1;
  const int kpg = (((S + 3) >> 2) + NG - 1) / NG * 4;
  const int j_beg = g * kpg, j_end = min(S, j_beg + kpg);
  const float scale = rsqrtf((float)HD);
  const bool drop = p_drop > 0.f;
  const uint64_t seed = drop ? seed_value(seed_ref) : 0ull;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  const int S4 = (S + 3) & ~3;
  float* out = dqkv + ((size_t)ir * B + b) * 3 * d + (size_t)h * HD;

  // ---- sweep A: dQ ----
  {
    float acc[HD], q[HD], r[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = 0.f;
    row_to_regs<HD>(Qs + ir * HD, q, scale * kLog2e);   // only feeds the exponent: P_ij = 2^(q'_i.k_j - lse'_i)
    row_to_regs<HD>(dOs + ir * HD, r, 1.f);
    const float Di = Ds[ir], li = Ls[ir];
    const uint64_t ebase = ((uint64_t)bh * S + ir) * S4;
    for (int j0 = j_beg; j0 < j_end; j0 += 4) {
      float msk[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop) dropout_scale4(seed, site, ebase + j0, p_drop, dscale, msk);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        if (j < j_end) {
          const float pj = ex2_f(dot_smem<HD>(q, Ks + j * HD) - li);
          const float dpd = dot_smem<HD>(r, Vs + j * HD);
          axpy_row<HD>(pj * (dpd * msk[u] - Di) * scale, Ks + j * HD, acc);
        }
      }
    }
    if (g > 0) {
      float* pr = part + ((size_t)(g - 1) * 32 + lane) * PW;
#pragma unroll
      for (int c = 0; c < HD; ++c) pr[c] = acc[c];
    }
    __syncthreads();
    if (g == 0 && active) {
#pragma unroll
      for (int gg = 0; gg < NG - 1; ++gg) {
        const float* pr = part + ((size_t)gg * 32 + lane) * PW;
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] += pr[c];
      }
      store_row<HD>(out, acc, 1.f);
    }
    __syncthreads();   // the partial rows are rewritten by sweep B
  }

  // ---- sweep B: dV, dK ----
  float kj[HD], vj[HD], accv[HD], acck[HD];
  row_to_regs<HD>(Ks + ir * HD, kj, scale * kLog2e);   // scaled once here: P_ij = 2^((k_j scale log2e).q_i - lse'_i)
  row_to_regs<HD>(Vs + ir * HD, vj, 1.f);
#pragma unroll
  for (int c = 0; c < HD; ++c) { accv[c] = 0.f; acck[c] = 0.f; }
  const uint64_t key = drop ? drop_key(seed, site) : 0ull;
  const uint32_t thr = drop ? drop_threshold(p_drop) : 0u;
  const uint32_t wq = (uint32_t)(ir >> 2);           // the key quad's word within a query's mask row
  const uint32_t sh = 16u * (uint32_t)(ir & 3);
  const uint32_t words_per_row = (uint32_t)(S4 >> 2);
  for (int i0 = j_beg; i0 < j_end; i0 += 4) {
    uint32_t wlo = 0u, whi = 0u;
    if (drop) {
      const int iq = min(i0 + (lane & 3), S - 1);
      const uint64_t w = drop_word(key, ((uint64_t)bh * S + iq) * words_per_row + wq);
      wlo = (uint32_t)w; whi = (uint32_t)(w >> 32);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int qi = i0 + u;          // warp-uniform
      float m = 1.f;
      if (drop) {
        const uint32_t lo = __shfl_sync(0xffffffffu, wlo, (lane & ~3) | u);
        const uint32_t hi = __shfl_sync(0xffffffffu, whi, (lane & ~3) | u);
        const uint64_t w = ((uint64_t)hi << 32) | lo;
        m = ((uint32_t)(w >> sh) & 0xFFFFu) >= thr ? dscale : 0.f;
      }
      if (qi < j_end) {
        const float* qrow = Qs + qi * HD;
        const float* drow = dOs + qi * HD;
        const float p = ex2_f(dot_smem<HD>(kj, qrow) - Ls[qi]);
        const float dpd = dot_smem<HD>(vj, drow);
        axpy_row<HD>(p * m, drow, accv);
        axpy_row<HD>(p * (dpd * m - Ds[qi]) * scale, qrow, acck);
      }
    }
  }
  if (g > 0) {
    float* pr = part + ((size_t)(g - 1) * 32 + lane) * PW;
#pragma unroll
    for (int c = 0; c < HD; ++c) { pr[c] = accv[c]; pr[HD + c] = acck[c]; }
  }
  __syncthreads();
  if (g == 0 && active) {
#pragma unroll
    for (int gg = 0; gg < NG - 1; ++gg) {
      const float* pr = part + ((size_t)gg * 32 + lane) * PW;
#pragma unroll
      for (int c = 0; c < HD; ++c) { accv[c] += pr[c]; acck[c] += pr[HD + c]; }
    }
    store_row<HD>(out + 2 * d, accv, 1.f);
    store_row<HD>(out + d, acck, 1.f);
  }
}

inline int att_threads(int S) { return NG * ((S + 31) / 32) * 32; }

template <int HD>
int launch_fwd(const float* qkv, float* o, float* lse, int S, int B, int d, int nhead, float p, Seed seed, int site,
               cudaStream_t st) {
  if constexpr (HD <= 16) {
    auto bytes = [](int s) { return ((size_t)2 * s * HD + (size_t)2 * NG * 32 + (size_t)(NG - 1) * 32 * (HD | 1)) * sizeof(float); };
    attention_fwd_small_kernel<HD><<<B * nhead * ((S + 31) / 32), NG * 32, bytes(S), st>>>(qkv, o, lse, S, B, d, nhead, p, seed, (uint32_t)site);
    GANFFN_LAUNCHED("attention_fwd_small_kernel");
    return GANFFN_OK;
  } else {
  auto bytes = [](int s) { return ((size_t)3 * s * HD + (size_t)s * (s | 1) + (size_t)NG * s) * sizeof(float); };
  GANFFN_SMEM_OPTIN(attention_fwd_kernel<HD>, bytes(GANFFN_MAX_SEQ));
  attention_fwd_kernel<HD><<<B * nhead, att_threads(S), bytes(S), st>>>(qkv, o, lse, S, B, d, nhead, p, seed, (uint32_t)site);
  GANFFN_LAUNCHED("attention_fwd_kernel");
  return GANFFN_OK;
  }
}

template <int HD>
int launch_bwd(const float* qkv, const float* o, const float* lse, const float* d_o, float* dqkv, int S, int B, int d,
               int nhead, float p, Seed seed, int site, cudaStream_t st) {
  if constexpr (HD <= 16) {
    auto bytes = [](int s) { return ((size_t)4 * s * HD + (size_t)2 * s + (size_t)(NG - 1) * 32 * ((2 * HD) | 1)) * sizeof(float); };
    attention_bwd_small_kernel<HD><<<B * nhead * ((S + 31) / 32), NG * 32, bytes(S), st>>>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p,
                                                                                          seed, (uint32_t)site);
    GANFFN_LAUNCHED("attention_bwd_small_kernel");
    return GANFFN_OK;
  } else {
  auto bytes = [](int s) { return ((size_t)4 * s * HD + (size_t)2 * s * (s | 1) + s) * sizeof(float); };
  GANFFN_SMEM_OPTIN(attention_bwd_kernel<HD>, bytes(GANFFN_MAX_SEQ));
  attention_bwd_kernel<HD><<<B * nhead, att_threads(S), bytes(S), st>>>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed,
                                                                       (uint32_t)site);
  GANFFN_LAUNCHED("attention_bwd_kernel");
  return GANFFN_OK;
  }
}

}  // namespace

// bit 0: forward on the tensor-core kernels (attention_mma.cu), bit 1: backward.  GANFFN_ATTN=0 selects the FFMA
// kernels of this file (kept as the exact-fp32 cross-check and for A/B runs), 1 / 2 one direction only.
// Default (measured, warm L2, S=94 B=32, tools/attn_time.py): head_dim 64 -- tensor-core kernels 28-30 us forward /
// 84-95 us backward against 40-42 / 97-98 us for the FFMA kernels; head_dim 10 -- both are bound by the latency of one
// warp's dependent chain (16-20 / 39-50 us on mma.sync, 16-17 / 29-33 us on FFMA with four warps per 32 rows), so the
// narrow heads stay on the FFMA kernels.
static int attn_engine(int head_dim) {
  static const int mode = getenv("GANFFN_ATTN") ? atoi(getenv("GANFFN_ATTN")) : -1;
  if (mode >= 0) return mode;
  return head_dim >= 32 ? 3 : 0;
}

int attention_fwd(const float* qkv, float* o, float* lse, int S, int B, int d, int nhead, float p, Seed seed,
                  int site, cudaStream_t st) {
  GANFFN_CHECK_ARG(S >= 1 && S <= GANFFN_MAX_SEQ, "attention: seq_len %d outside [1,%d] (model.py:1179)", S,
                   GANFFN_MAX_SEQ);
  GANFFN_CHECK_ARG(B >= 1 && nhead >= 1 && d % nhead == 0, "attention: d=%d not divisible by nhead=%d", d, nhead);
  GANFFN_CHECK_ARG(p >= 0.f && p < 1.f, "attention: dropout p=%f", p);
  if (attn_engine(d / nhead) & 1) {
    const int rc = attention_fwd_mma(qkv, o, lse, S, B, d, nhead, p, seed, site, st);
    if (rc >= 0) return rc;
  }
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
                  int nhead, float p, Seed seed, int site, cudaStream_t st) {
  GANFFN_CHECK_ARG(S >= 1 && S <= GANFFN_MAX_SEQ, "attention: seq_len %d outside [1,%d] (model.py:1179)", S,
                   GANFFN_MAX_SEQ);
  GANFFN_CHECK_ARG(B >= 1 && nhead >= 1 && d % nhead == 0, "attention: d=%d not divisible by nhead=%d", d, nhead);
  if (attn_engine(d / nhead) & 2) {
    const int rc = attention_bwd_mma(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed, site, st);
    if (rc >= 0) return rc;
  }
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
