// Whole-dialogue self-attention on the tensor cores with fp32 parity (3xTF32), forward and backward.
//
// One CTA = one (dialogue, head); one warp = 16 query rows (forward, backward sweep A) or 16 key rows (backward
// sweep B); every product is a chain of mma.sync.m16n8k8 TF32 instructions on register fragments, error-compensated
// like the GEMM engine: x = hi + lo, hi = rna_tf32(x); a.b ~ lo_a.hi_b + hi_a.lo_b + hi_a.hi_b (small terms first).
// S <= 110 keys and head_dim <= 64 mean the operands of a whole (dialogue, head) fit in shared memory (K/V, and in the
// backward pass Q/dO too) and a warp's 16 x S score strip is recomputed rather than stored, so there is no S x S
// matrix anywhere.  (tcgen05 is the wrong tool here: its smallest A tile is 128 rows x 8 per instruction with the
// operands staged through TMEM / swizzled shared memory, while a dialogue has 94 rows and a head 10 columns; the
// r2 profile of the FFMA kernels this file replaces showed 8.1 M warp instructions per d=100 layer forward at 37 %
// issue utilisation -- this formulation issues about a quarter of that.)
//
// The score tile of a warp (16 queries x 8 keys) comes out of the MMA as a C fragment: lane (g = lane/4, t = lane%4)
// holds rows g, g+8 and key columns 2t, 2t+1.  As the A operand of the next product (P V, dS K, ...) the same values are
// fed without any shuffle by renaming the contraction index: A's k-slot t stands for key 2t and k-slot t+4 for key 2t+1,
// and the B fragment (V, K, dO or Q rows) is loaded with the same renaming.
//
// Reference: F.scaled_dot_product_attention inside nn.MultiheadAttention (torch functional.py
// multi_head_attention_forward), dropout on the probabilities; padded slots are real tokens.
#include <stdlib.h>
#include "common.cuh"

namespace ganffn {
namespace {

__device__ __forceinline__ uint32_t tf32_hi(float x) { return tf32_rna_bits(x); }
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_hi(x);
  lo = __float_as_uint(x - __uint_as_float(hi));   // the tensor core reads the top 19 bits: lo is truncated, |err| <= 2^-21 |x|
}
// d += a (16x8, row) * b (8x8, col)
__device__ __forceinline__ void mma8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 3xTF32: d += (ahi + alo) * (bhi + blo) without the lo*lo term, small terms first
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], float b0, float b1) {
  uint32_t h0, l0, h1, l1;
  split(b0, h0, l0);
  split(b1, h1, l1);
  mma8(d, alo, h0, h1);
  mma8(d, ahi, l0, l1);
  mma8(d, ahi, h0, h1);
}

template <int HD>
struct Dims {
  static constexpr int HDP = (HD + 7) / 8 * 8;   // head_dim padded to the MMA k / n granularity (zeros)
  static constexpr int KS = HDP / 8;             // 8-wide steps along the head dimension
  static constexpr int LD = HDP + 4;             // row stride of the shared-memory tiles: conflict-free fragment loads
  static constexpr int W = (HD % 4 == 0) ? 4 : (HD % 2 == 0) ? 2 : 1;
};

// [S, HD] slice of one head (row stride ld floats in global) -> smem [rows][LD], zero padded to `rows` x HDP.
// Asynchronous (cp.async, common.cuh): every chunk of the tile is in flight at once instead of one exposed global
// latency per loop iteration; the caller runs cp_async_wait_all() + __syncthreads() (and scale_head for Q) afterwards.
template <int HD>
__device__ __forceinline__ void load_head(const float* __restrict__ g, int ld, float* s, int S, int rows) {
  constexpr int LD = Dims<HD>::LD, HDP = Dims<HD>::HDP, W = Dims<HD>::W, CPR = HDP / W;
  static_assert(LD % W == 0, "cp.async destination alignment");
  for (int idx = threadIdx.x; idx < rows * CPR; idx += blockDim.x) {
    const int r = idx / CPR, c = (idx % CPR) * W;
    float* dst = s + r * LD + c;
    if (r < S && c < HD) {
      const float* src = g + (size_t)r * ld + c;
      if (W == 4) cp_async_16(dst, src);
      else if (W == 2) cp_async_8(dst, src);
      else cp_async_4(dst, src);
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) dst[j] = 0.f;
    }
  }
}
// in-place scaling of a staged tile (Q is kept pre-multiplied by 1/sqrt(head_dim)); between two __syncthreads()
template <int HD>
__device__ __forceinline__ void scale_head(float* s, int rows, float mul) {
  constexpr int LD = Dims<HD>::LD, HDP = Dims<HD>::HDP;
  for (int idx = threadIdx.x; idx < rows * HDP; idx += blockDim.x) s[(idx / HDP) * LD + idx % HDP] *= mul;
}

// A fragments (hi, lo) of 16 rows starting at row0 of a [rows][LD] smem tile, all KS k-steps
template <int HD>
__device__ __forceinline__ void load_a_frags(const float* __restrict__ s, int row0, int g, int t,
                                             uint32_t (&hi)[Dims<HD>::KS][4], uint32_t (&lo)[Dims<HD>::KS][4]) {
  constexpr int LD = Dims<HD>::LD;
#pragma unroll
  for (int kk = 0; kk < Dims<HD>::KS; ++kk) {
    split(s[(row0 + g) * LD + 8 * kk + t], hi[kk][0], lo[kk][0]);
    split(s[(row0 + g + 8) * LD + 8 * kk + t], hi[kk][1], lo[kk][1]);
    split(s[(row0 + g) * LD + 8 * kk + t + 4], hi[kk][2], lo[kk][2]);
    split(s[(row0 + g + 8) * LD + 8 * kk + t + 4], hi[kk][3], lo[kk][3]);
  }
}

// c (16 x 8) = A (16 x HDP, fragments) . X[j0 .. j0+8)^T : contraction over the head dimension
template <int HD>
__device__ __forceinline__ void dot_tile(float (&c)[4], const uint32_t (&ahi)[Dims<HD>::KS][4], const uint32_t (&alo)[Dims<HD>::KS][4],
                                         const float* __restrict__ X, int j0, int g, int t) {
  constexpr int LD = Dims<HD>::LD;
  c[0] = c[1] = c[2] = c[3] = 0.f;
  const float* row = X + (j0 + g) * LD + t;
#pragma unroll
  for (int kk = 0; kk < Dims<HD>::KS; ++kk) mma3(c, ahi[kk], alo[kk], row[8 * kk], row[8 * kk + 4]);
}

// acc[nn] (16 x 8 each, HDP/8 of them) += P (16 x 8 keys, C-fragment values as the A operand) . X[j0 .. j0+8)[:, 8 nn ..]
// k-slot t <-> row j0 + 2t, k-slot t+4 <-> row j0 + 2t + 1 (see the file header)
template <int HD>
__device__ __forceinline__ void acc_tile(float (&acc)[Dims<HD>::KS][4], const float (&p)[4], const float* __restrict__ X, int j0,
                                         int g, int t) {
  constexpr int LD = Dims<HD>::LD;
  uint32_t phi[4], plo[4];
  split(p[0], phi[0], plo[0]);   // row g,   key 2t    -> a0 (row g,   k t)
  split(p[2], phi[1], plo[1]);   // row g+8, key 2t    -> a1 (row g+8, k t)
  split(p[1], phi[2], plo[2]);   // row g,   key 2t+1  -> a2 (row g,   k t+4)
  split(p[3], phi[3], plo[3]);   // row g+8, key 2t+1  -> a3 (row g+8, k t+4)
  const float* r0 = X + (j0 + 2 * t) * LD + g;
#pragma unroll
  for (int nn = 0; nn < Dims<HD>::KS; ++nn) mma3(acc[nn], phi, plo, r0[8 * nn], r0[LD + 8 * nn]);
}

// Scaled keep masks of a C fragment: rows r0 + g, r0 + g + 8, keys j0 + 2t, j0 + 2t + 1 (one 64-bit word per row:
// both keys lie in the same group of four because j0 % 8 == 0).
__device__ __forceinline__ void frag_masks(float (&m)[4], uint64_t key, uint32_t thr, float dscale, uint64_t row_word0,
                                           uint32_t words_per_row, int j0, int t) {
  const uint32_t wq = (uint32_t)((j0 + 2 * t) >> 2), sh = 16u * (uint32_t)((2 * t) & 3);
  const uint64_t w0 = drop_word(key, row_word0 + wq);
  const uint64_t w1 = drop_word(key, row_word0 + (uint64_t)8 * words_per_row + wq);
  m[0] = ((uint32_t)(w0 >> sh) & 0xFFFFu) >= thr ? dscale : 0.f;
  m[1] = ((uint32_t)(w0 >> (sh + 16)) & 0xFFFFu) >= thr ? dscale : 0.f;
  m[2] = ((uint32_t)(w1 >> sh) & 0xFFFFu) >= thr ? dscale : 0.f;
  m[3] = ((uint32_t)(w1 >> (sh + 16)) & 0xFFFFu) >= thr ? dscale : 0.f;
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ---- forward ----------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(224) attention_fwd_mma_kernel(const float* __restrict__ qkv, float* __restrict__ o,
                                                                float* __restrict__ lse, int S, int B, int d, int nhead,
                                                                float p_drop, const Seed seed_ref, uint32_t site) {
  extern __shared__ __align__(16) float smem[];
  constexpr int LD = Dims<HD>::LD, KS = Dims<HD>::KS;
  const int S8 = (S + 7) & ~7, S16 = (S + 15) & ~15;
  float* Qs = smem;                 // [S16][LD], pre-scaled by 1/sqrt(HD)
  float* Ks = Qs + S16 * LD;        // [S8][LD]
  float* Vs = Ks + S8 * LD;         // [S8][LD]
  const int b = blockIdx.x / nhead, h = blockIdx.x % nhead;
  const int ld = B * 3 * d;
  const float* base = qkv + (size_t)b * 3 * d + (size_t)h * HD;
  load_head<HD>(base, ld, Qs, S, S16);
  load_head<HD>(base + d, ld, Ks, S, S8);
  load_head<HD>(base + 2 * d, ld, Vs, S, S8);
  cp_async_wait_all();
  __syncthreads();
  scale_head<HD>(Qs, S16, rsqrtf((float)HD) * kLog2e);   // scores in log2 units: every probability is one ex2
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16;
  if (r0 >= S) return;
  uint32_t qhi[KS][4], qlo[KS][4];
  load_a_frags<HD>(Qs, r0, g, t, qhi, qlo);
  const int nt = S8 >> 3;

  // pass 1: row maxima (rows r0 + g and r0 + g + 8)
  float mx0 = -INFINITY, mx1 = -INFINITY;
  for (int n = 0; n < nt; ++n) {
    float c[4];
    dot_tile<HD>(c, qhi, qlo, Ks, 8 * n, g, t);
    const int j = 8 * n + 2 * t;
    if (j < S) { mx0 = fmaxf(mx0, c[0]); mx1 = fmaxf(mx1, c[2]); }
    if (j + 1 < S) { mx0 = fmaxf(mx0, c[1]); mx1 = fmaxf(mx1, c[3]); }
  }
  mx0 = quad_max(mx0);
  mx1 = quad_max(mx1);

  // pass 2: probabilities, row sums, P V
  const bool drop = p_drop > 0.f;
  const uint64_t key = drop ? drop_key(seed_value(seed_ref), site) : 0ull;
  const uint32_t thr = drop ? drop_threshold(p_drop) : 0u;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  const uint32_t words_per_row = (uint32_t)(((S + 3) & ~3) >> 2);
  const uint64_t row_word0 = ((uint64_t)blockIdx.x * S + (uint64_t)(r0 + g)) * words_per_row;
  float acc[KS][4];
#pragma unroll
  for (int nn = 0; nn < KS; ++nn) acc[nn][0] = acc[nn][1] = acc[nn][2] = acc[nn][3] = 0.f;
  float l0 = 0.f, l1 = 0.f;
  for (int n = 0; n < nt; ++n) {
    float c[4];
    dot_tile<HD>(c, qhi, qlo, Ks, 8 * n, g, t);
    const int j = 8 * n + 2 * t;
    float p[4];
    p[0] = j < S ? ex2_f(c[0] - mx0) : 0.f;
    p[1] = j + 1 < S ? ex2_f(c[1] - mx0) : 0.f;
    p[2] = j < S ? ex2_f(c[2] - mx1) : 0.f;
    p[3] = j + 1 < S ? ex2_f(c[3] - mx1) : 0.f;
    l0 += p[0] + p[1];
    l1 += p[2] + p[3];
    if (drop) {
      float m[4];
      frag_masks(m, key, thr, dscale, row_word0, words_per_row, 8 * n, t);
      p[0] *= m[0]; p[1] *= m[1]; p[2] *= m[2]; p[3] *= m[3];
    }
    acc_tile<HD>(acc, p, Vs, 8 * n, g, t);
  }
  l0 = quad_sum(l0);
  l1 = quad_sum(l1);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int i0 = r0 + g, i1 = r0 + g + 8;
  float* o0 = o + ((size_t)i0 * B + b) * d + (size_t)h * HD;
  float* o1 = o + ((size_t)i1 * B + b) * d + (size_t)h * HD;
#pragma unroll
  for (int nn = 0; nn < KS; ++nn) {
    const int c = 8 * nn + 2 * t;     // C fragment columns 2t, 2t+1 of n-tile nn (HD is even for every supported head_dim)
    if (c < HD) {
      if (i0 < S) *reinterpret_cast<float2*>(o0 + c) = make_float2(acc[nn][0] * inv0, acc[nn][1] * inv0);
      if (i1 < S) *reinterpret_cast<float2*>(o1 + c) = make_float2(acc[nn][2] * inv1, acc[nn][3] * inv1);
    }
  }
  if (t == 0) {
    if (i0 < S) lse[(size_t)blockIdx.x * S + i0] = mx0 * kLn2 + logf(l0);   // natural-log lse
    if (i1 < S) lse[(size_t)blockIdx.x * S + i1] = mx1 * kLn2 + logf(l1);
  }
}

// ---- backward ---------------------------------------------------------------------------------------------------------
//   sweep A  warp = 16 queries:  S = Q K^T, dP = dO V^T (both contract over the head dimension),
//                                P = exp(S - lse), dS = P (dP m - D) scale;          dQ += dS K   (contracts over keys)
//   sweep B  warp = 16 keys:     S^T = K Q^T, dP^T = V dO^T, the same P / dS seen from the key's side;
//                                dV += (P m)^T dO,  dK += dS^T Q                     (contract over queries)
// Q is kept pre-scaled (q / sqrt(hd)) in shared memory, so dK = dS_unscaled^T Q_scaled needs no extra factor and
// dQ = (dS_unscaled K) / sqrt(hd) is scaled once at the store.
template <int HD>
__global__ void __launch_bounds__(224) attention_bwd_mma_kernel(
    const float* __restrict__ qkv, const float* __restrict__ o, const float* __restrict__ lse,
    const float* __restrict__ d_o, float* __restrict__ dqkv, int S, int B, int d, int nhead, float p_drop,
    const Seed seed_ref, uint32_t site, int parts) {
  extern __shared__ __align__(16) float smem[];
  constexpr int LD = Dims<HD>::LD, KS = Dims<HD>::KS;
  const int S16 = (S + 15) & ~15;
  float* Qs = smem;                 // [S16][LD] scaled by 1/sqrt(HD)
  float* Ks = Qs + S16 * LD;
  float* Vs = Ks + S16 * LD;
  float* dOs = Vs + S16 * LD;
  float* Ls = dOs + S16 * LD;       // [S16] row log-sum-exp
  float* Ds = Ls + S16;             // [S16] rowsum(dO * O)
  // `parts` CTAs may share one (dialogue, head), each staging all four tiles and owning a contiguous range of 16-row
  // strips (tuning knob GANFFN_ATTN_BWD_PARTS; head_dim 64 needs ~250 registers per thread, i.e. one 192-thread CTA per
  // SM and two waves for the 256 CTAs of an S=94, B=32 layer -- but splitting did not pay: see the launcher).
  const int bh = blockIdx.x / parts, part = blockIdx.x % parts;
  const int b = bh / nhead, h = bh % nhead;
  const int ld = B * 3 * d, ldo = B * d;
  const float scale = rsqrtf((float)HD);
  const float* base = qkv + (size_t)b * 3 * d + (size_t)h * HD;
  const float* obase = o + (size_t)b * d + (size_t)h * HD;
  const float* dobase = d_o + (size_t)b * d + (size_t)h * HD;
  load_head<HD>(base, ld, Qs, S, S16);
  load_head<HD>(base + d, ld, Ks, S, S16);
  load_head<HD>(base + 2 * d, ld, Vs, S, S16);
  load_head<HD>(dobase, ldo, dOs, S, S16);
  for (int r = threadIdx.x; r < S16; r += blockDim.x) {
    float acc = 0.f, l = 0.f;
    if (r < S) {
      constexpr int W = Dims<HD>::W;
      const float* orow = obase + (size_t)r * ldo;
      const float* drow = dobase + (size_t)r * ldo;
#pragma unroll
      for (int c = 0; c < HD; c += W) {
        if (W == 4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(orow + c)), e = __ldg(reinterpret_cast<const float4*>(drow + c));
          acc = fmaf(a.x, e.x, acc); acc = fmaf(a.y, e.y, acc); acc = fmaf(a.z, e.z, acc); acc = fmaf(a.w, e.w, acc);
        } else if (W == 2) {
          const float2 a = __ldg(reinterpret_cast<const float2*>(orow + c)), e = __ldg(reinterpret_cast<const float2*>(drow + c));
          acc = fmaf(a.x, e.x, acc); acc = fmaf(a.y, e.y, acc);
        } else {
          acc = fmaf(__ldg(orow + c), __ldg(drow + c), acc);
        }
      }
      l = lse[(size_t)bh * S + r] * kLog2e;   // log2 units, like the scores (Q carries log2e below)
    }
    Ds[r] = acc;
    Ls[r] = l;
  }
  cp_async_wait_all();
  __syncthreads();
  scale_head<HD>(Qs, S16, scale * kLog2e);   // Q' = Q log2(e) / sqrt(hd): dK = dS^T Q' is multiplied by ln 2 at the store
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = (part * (int)(blockDim.x >> 5) + warp) * 16;
  if (r0 >= S) return;
  const int nt = S16 >> 3;
  const bool drop = p_drop > 0.f;
  const uint64_t key = drop ? drop_key(seed_value(seed_ref), site) : 0ull;
  const uint32_t thr = drop ? drop_threshold(p_drop) : 0u;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  const uint32_t words_per_row = (uint32_t)(((S + 3) & ~3) >> 2);
  const int i0 = r0 + g, i1 = r0 + g + 8;
  float* out0 = dqkv + ((size_t)i0 * B + b) * 3 * d + (size_t)h * HD;
  float* out1 = dqkv + ((size_t)i1 * B + b) * 3 * d + (size_t)h * HD;

  // ---- sweep A: rows = queries r0 .. r0+15 ----
  {
    uint32_t qhi[KS][4], qlo[KS][4], ghi[KS][4], glo[KS][4];
    load_a_frags<HD>(Qs, r0, g, t, qhi, qlo);
    load_a_frags<HD>(dOs, r0, g, t, ghi, glo);
    const float L0 = Ls[i0], L1 = Ls[i1], D0 = Ds[i0], D1 = Ds[i1];
    const uint64_t row_word0 = ((uint64_t)bh * S + (uint64_t)i0) * words_per_row;
    float acc[KS][4];
#pragma unroll
    for (int nn = 0; nn < KS; ++nn) acc[nn][0] = acc[nn][1] = acc[nn][2] = acc[nn][3] = 0.f;
    for (int n = 0; n < nt; ++n) {
      float c[4], e[4];
      dot_tile<HD>(c, qhi, qlo, Ks, 8 * n, g, t);     // scores (already scaled through Q)
      dot_tile<HD>(e, ghi, glo, Vs, 8 * n, g, t);     // dP
      const int j = 8 * n + 2 * t;
      float m[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop) frag_masks(m, key, thr, dscale, row_word0, words_per_row, 8 * n, t);
      float ds[4];
      ds[0] = j < S ? ex2_f(c[0] - L0) * (e[0] * m[0] - D0) : 0.f;
      ds[1] = j + 1 < S ? ex2_f(c[1] - L0) * (e[1] * m[1] - D0) : 0.f;
      ds[2] = j < S ? ex2_f(c[2] - L1) * (e[2] * m[2] - D1) : 0.f;
      ds[3] = j + 1 < S ? ex2_f(c[3] - L1) * (e[3] * m[3] - D1) : 0.f;
      acc_tile<HD>(acc, ds, Ks, 8 * n, g, t);
    }
#pragma unroll
    for (int nn = 0; nn < KS; ++nn) {
      const int c = 8 * nn + 2 * t;
      if (c < HD) {
        if (i0 < S) *reinterpret_cast<float2*>(out0 + c) = make_float2(acc[nn][0] * scale, acc[nn][1] * scale);
        if (i1 < S) *reinterpret_cast<float2*>(out1 + c) = make_float2(acc[nn][2] * scale, acc[nn][3] * scale);
      }
    }
  }

  // ---- sweep B: rows = keys r0 .. r0+15, columns = queries ----
  {
    uint32_t khi[KS][4], klo[KS][4], vhi[KS][4], vlo[KS][4];
    load_a_frags<HD>(Ks, r0, g, t, khi, klo);
    load_a_frags<HD>(Vs, r0, g, t, vhi, vlo);
    float accv[KS][4], acck[KS][4];
#pragma unroll
    for (int nn = 0; nn < KS; ++nn) {
      accv[nn][0] = accv[nn][1] = accv[nn][2] = accv[nn][3] = 0.f;
      acck[nn][0] = acck[nn][1] = acck[nn][2] = acck[nn][3] = 0.f;
    }
    // dropout bits of element (query q, key k): word (q, k / 4), field k % 4; here the fragment holds keys i0, i1 (rows)
    // and queries 8n + 2t, 8n + 2t + 1 (columns): four words per tile
    const uint32_t wk0 = (uint32_t)(i0 >> 2), wk1 = (uint32_t)(i1 >> 2);
    const uint32_t sh0 = 16u * (uint32_t)(i0 & 3), sh1 = 16u * (uint32_t)(i1 & 3);
    for (int n = 0; n < nt; ++n) {
      float c[4], e[4];
      dot_tile<HD>(c, khi, klo, Qs, 8 * n, g, t);     // S^T
      dot_tile<HD>(e, vhi, vlo, dOs, 8 * n, g, t);    // dP^T
      const int q = 8 * n + 2 * t;                    // queries q, q + 1 (columns)
      const float La = Ls[q], Lb = Ls[q + 1], Da = Ds[q], Db = Ds[q + 1];
      float m[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop) {
        const uint64_t rowa = ((uint64_t)bh * S + (uint64_t)min(q, S - 1)) * words_per_row;
        const uint64_t rowb = ((uint64_t)bh * S + (uint64_t)min(q + 1, S - 1)) * words_per_row;
        const uint64_t wa0 = drop_word(key, rowa + wk0), wb0 = drop_word(key, rowb + wk0);
        const uint64_t wa1 = drop_word(key, rowa + wk1), wb1 = drop_word(key, rowb + wk1);
        m[0] = ((uint32_t)(wa0 >> sh0) & 0xFFFFu) >= thr ? dscale : 0.f;   // key i0, query q
        m[1] = ((uint32_t)(wb0 >> sh0) & 0xFFFFu) >= thr ? dscale : 0.f;   // key i0, query q+1
        m[2] = ((uint32_t)(wa1 >> sh1) & 0xFFFFu) >= thr ? dscale : 0.f;   // key i1, query q
        m[3] = ((uint32_t)(wb1 >> sh1) & 0xFFFFu) >= thr ? dscale : 0.f;   // key i1, query q+1
      }
      const bool va = q < S, vb = q + 1 < S;
      float p[4], pm[4], ds[4];
      p[0] = va ? ex2_f(c[0] - La) : 0.f;
      p[1] = vb ? ex2_f(c[1] - Lb) : 0.f;
      p[2] = va ? ex2_f(c[2] - La) : 0.f;
      p[3] = vb ? ex2_f(c[3] - Lb) : 0.f;
      pm[0] = p[0] * m[0]; pm[1] = p[1] * m[1]; pm[2] = p[2] * m[2]; pm[3] = p[3] * m[3];
      ds[0] = p[0] * (e[0] * m[0] - Da);
      ds[1] = p[1] * (e[1] * m[1] - Db);
      ds[2] = p[2] * (e[2] * m[2] - Da);
      ds[3] = p[3] * (e[3] * m[3] - Db);
      acc_tile<HD>(accv, pm, dOs, 8 * n, g, t);       // dV += (P m)^T dO
      acc_tile<HD>(acck, ds, Qs, 8 * n, g, t);        // dK += dS^T (Q / sqrt(hd))
    }
#pragma unroll
    for (int nn = 0; nn < KS; ++nn) {
      const int c = 8 * nn + 2 * t;
      if (c < HD) {
        if (i0 < S) {
          *reinterpret_cast<float2*>(out0 + 2 * d + c) = make_float2(accv[nn][0], accv[nn][1]);
          *reinterpret_cast<float2*>(out0 + d + c) = make_float2(acck[nn][0] * kLn2, acck[nn][1] * kLn2);
        }
        if (i1 < S) {
          *reinterpret_cast<float2*>(out1 + 2 * d + c) = make_float2(accv[nn][2], accv[nn][3]);
          *reinterpret_cast<float2*>(out1 + d + c) = make_float2(acck[nn][2] * kLn2, acck[nn][3] * kLn2);
        }
      }
    }
  }
}

inline int mma_threads(int S) { return ((S + 15) / 16) * 32; }

}  // namespace

template <int HD>
static int launch_fwd_mma(const float* qkv, float* o, float* lse, int S, int B, int d, int nhead, float p, Seed seed, int site,
                          cudaStream_t st) {
  auto bytes = [](int s) { return (size_t)(((s + 15) & ~15) + 2 * ((s + 7) & ~7)) * Dims<HD>::LD * sizeof(float); };
  GANFFN_SMEM_OPTIN(attention_fwd_mma_kernel<HD>, bytes(GANFFN_MAX_SEQ));
  attention_fwd_mma_kernel<HD><<<B * nhead, mma_threads(S), bytes(S), st>>>(qkv, o, lse, S, B, d, nhead, p, seed, (uint32_t)site);
  GANFFN_LAUNCHED("attention_fwd_mma_kernel");
  return GANFFN_OK;
}

template <int HD>
static int launch_bwd_mma(const float* qkv, const float* o, const float* lse, const float* d_o, float* dqkv, int S, int B, int d,
                          int nhead, float p, Seed seed, int site, cudaStream_t st) {
  auto bytes = [](int s) { const int s16 = (s + 15) & ~15; return ((size_t)4 * s16 * Dims<HD>::LD + 2 * s16) * sizeof(float); };
  GANFFN_SMEM_OPTIN(attention_bwd_mma_kernel<HD>, bytes(GANFFN_MAX_SEQ));
  static const int parts_env = getenv("GANFFN_ATTN_BWD_PARTS") ? atoi(getenv("GANFFN_ATTN_BWD_PARTS")) : 0;   // A/B switch
  const int strips = (S + 15) / 16;
  int parts = parts_env > 0 ? parts_env : 1;   // measured (r2, S=94 B=32 head_dim 64): 1 / 2 / 3 parts = 80 / 87 / 103 us -- every CTA stages all four tiles
  parts = std::min(parts, strips);
  const int warps = (strips + parts - 1) / parts;
  attention_bwd_mma_kernel<HD><<<B * nhead * parts, warps * 32, bytes(S), st>>>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed,
                                                                                (uint32_t)site, parts);
  GANFFN_LAUNCHED("attention_bwd_mma_kernel");
  return GANFFN_OK;
}

// head_dim must be even (float2 stores of C-fragment column pairs) -- 8, 10, 16, 32, 64 all are
int attention_fwd_mma(const float* qkv, float* o, float* lse, int S, int B, int d, int nhead, float p, Seed seed, int site,
                      cudaStream_t st) {
  switch (d / nhead) {
    case 8: return launch_fwd_mma<8>(qkv, o, lse, S, B, d, nhead, p, seed, site, st);
    case 10: return launch_fwd_mma<10>(qkv, o, lse, S, B, d, nhead, p, seed, site, st);
    case 16: return launch_fwd_mma<16>(qkv, o, lse, S, B, d, nhead, p, seed, site, st);
    case 32: return launch_fwd_mma<32>(qkv, o, lse, S, B, d, nhead, p, seed, site, st);
    case 64: return launch_fwd_mma<64>(qkv, o, lse, S, B, d, nhead, p, seed, site, st);
  }
  return -1;
}

int attention_bwd_mma(const float* qkv, const float* o, const float* lse, const float* d_o, float* dqkv, int S, int B, int d,
                      int nhead, float p, Seed seed, int site, cudaStream_t st) {
  switch (d / nhead) {
    case 8: return launch_bwd_mma<8>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed, site, st);
    case 10: return launch_bwd_mma<10>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed, site, st);
    case 16: return launch_bwd_mma<16>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed, site, st);
    case 32: return launch_bwd_mma<32>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed, site, st);
    case 64: return launch_bwd_mma<64>(qkv, o, lse, d_o, dqkv, S, B, d, nhead, p, seed, site, st);
  }
  return -1;
}

}  // namespace ganffn
