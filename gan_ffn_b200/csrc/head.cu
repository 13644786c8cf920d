// The discriminator head as ONE kernel (north_star kernel (1): "each modality's small ... discriminator MLP fused into
// one kernel with weights staged in shared memory and coalesced, vectorised 128-bit loads of utterance features"):
//   g0 = gelu(x)            x = last encoder output [T, d], d <= 128
//   f1 = drop(fc1(g0))      d  -> 64     a1 = gelu(f1)
//   f2 = drop(fc2(a1))      64 -> 16     a2 = gelu(f2)
//   out = sigmoid(drop(fc3(a2)))   16 -> 1
// Reference: AcousticDiscriminator / VisualDiscriminator / TextDiscriminator.forward, model.py:1320-1327, 1354-1364,
// 1390-1397 (dropout BEFORE the activation, as the reference orders them).  These layers are too narrow to be dense
// contractions worth a tensor-core tile (7 440 MACs per utterance): exact-fp32 FFMA from shared memory, one CTA per 32
// utterances, every intermediate the backward pass reads (g0, f1, a1, f2, a2) written once with 128-bit stores.
// Replaces one element-wise launch + three GEMM launches (+ a split-K fold) per discriminator pass.
#include <stdlib.h>
#include "kernels.h"

namespace ganffn {
namespace {

constexpr int HR = 32;          // utterances per CTA
constexpr int H1 = 64, H2 = 16;
constexpr int HD_MAX = 128;

__global__ void __launch_bounds__(256) disc_head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                            const float* __restrict__ b1, const float* __restrict__ w2,
                                                            const float* __restrict__ b2, const float* __restrict__ w3,
                                                            const float* __restrict__ b3, float* __restrict__ g0,
                                                            float* __restrict__ f1, float* __restrict__ a1, float* __restrict__ f2,
                                                            float* __restrict__ a2, float* __restrict__ out, int T, int d,
                                                            float p_drop, const Seed seed_ref, uint32_t site0) {
  extern __shared__ __align__(16) float sm[];
  float* w1t = sm;                          // [d][H1]   fc1 weight, k-major
  float* w2t = w1t + d * H1;                // [H1][H2]
  float* g0s = w2t + H1 * H2;               // [HR][d + 1]
  float* a1s = g0s + HR * (d + 1);          // [HR][H1 + 1]
  float* a2s = a1s + HR * (H1 + 1);         // [HR][H2 + 1]
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * HR;
  const int d4 = d >> 2;
  // ---- stage the weights (transposed: k-major) and gelu(x) of this CTA's rows ----
  for (int idx = t; idx < H1 * d4; idx += 256) {          // w1 [H1][d]: 128-bit loads along k
    const int n = idx / d4, k4 = idx - n * d4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(w1 + (size_t)n * d) + k4);
    w1t[(4 * k4 + 0) * H1 + n] = v.x; w1t[(4 * k4 + 1) * H1 + n] = v.y;
    w1t[(4 * k4 + 2) * H1 + n] = v.z; w1t[(4 * k4 + 3) * H1 + n] = v.w;
  }
  for (int idx = t; idx < H2 * (H1 / 4); idx += 256) {    // w2 [H2][H1]
    const int n = idx / (H1 / 4), k4 = idx - n * (H1 / 4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(w2 + (size_t)n * H1) + k4);
    w2t[(4 * k4 + 0) * H2 + n] = v.x; w2t[(4 * k4 + 1) * H2 + n] = v.y;
    w2t[(4 * k4 + 2) * H2 + n] = v.z; w2t[(4 * k4 + 3) * H2 + n] = v.w;
  }
  for (int idx = t; idx < HR * d4; idx += 256) {
    const int r = idx / d4, k4 = idx - r * d4, m = m0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < T) {
      v = __ldg(reinterpret_cast<const float4*>(x + (size_t)m * d) + k4);
      v = make_float4(gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w));
      reinterpret_cast<float4*>(g0 + (size_t)m * d)[k4] = v;
    }
    float* gs = g0s + r * (d + 1) + 4 * k4;
    gs[0] = v.x; gs[1] = v.y; gs[2] = v.z; gs[3] = v.w;
  }
  const bool drop = p_drop > 0.f;
  const uint64_t seed = drop ? seed_value(seed_ref) : 0ull;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  __syncthreads();
  // ---- fc1: [HR x d] . [d x 64]; thread = 2 rows x 4 columns ----
  {
    const int ty = t >> 4, tx = t & 15;
    const int r0 = 2 * ty, c0 = 4 * tx;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const float* ga = g0s + r0 * (d + 1);
    const float* gb = ga + (d + 1);
#pragma unroll 4
    for (int k = 0; k < d; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(w1t + k * H1 + c0);
      const float xa = ga[k], xb = gb[k];
      acc[0][0] = fmaf(xa, w.x, acc[0][0]); acc[0][1] = fmaf(xa, w.y, acc[0][1]);
      acc[0][2] = fmaf(xa, w.z, acc[0][2]); acc[0][3] = fmaf(xa, w.w, acc[0][3]);
      acc[1][0] = fmaf(xb, w.x, acc[1][0]); acc[1][1] = fmaf(xb, w.y, acc[1][1]);
      acc[1][2] = fmaf(xb, w.z, acc[1][2]); acc[1][3] = fmaf(xb, w.w, acc[1][3]);
    }
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + c0));
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = r0 + rr, m = m0 + r;
      float v[4] = {acc[rr][0] + bb.x, acc[rr][1] + bb.y, acc[rr][2] + bb.z, acc[rr][3] + bb.w};
      if (drop) {
        float msk[4];
        dropout_scale4(seed, site0 + 1, (uint64_t)m * H1 + (uint64_t)c0, p_drop, dscale, msk);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] *= msk[j];
      }
      const float4 act = make_float4(gelu_f(v[0]), gelu_f(v[1]), gelu_f(v[2]), gelu_f(v[3]));
      if (m < T) {
        *reinterpret_cast<float4*>(f1 + (size_t)m * H1 + c0) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(a1 + (size_t)m * H1 + c0) = act;
      }
      float* as = a1s + r * (H1 + 1) + c0;
      as[0] = act.x; as[1] = act.y; as[2] = act.z; as[3] = act.w;
    }
  }
  __syncthreads();
  // ---- fc2: [HR x 64] . [64 x 16]; thread = 1 row x 2 columns ----
  {
    const int r = t >> 3, c0 = 2 * (t & 7), m = m0 + r;
    float s0 = 0.f, s1 = 0.f;
    const float* ar = a1s + r * (H1 + 1);
#pragma unroll 8
    for (int k = 0; k < H1; ++k) {
      const float2 w = *reinterpret_cast<const float2*>(w2t + k * H2 + c0);
      s0 = fmaf(ar[k], w.x, s0);
      s1 = fmaf(ar[k], w.y, s1);
    }
    float v0 = s0 + __ldg(b2 + c0), v1 = s1 + __ldg(b2 + c0 + 1);
    if (drop) {
      v0 *= dropout_scale1(seed, site0 + 2, (uint64_t)m * H2 + (uint64_t)c0, p_drop, dscale);
      v1 *= dropout_scale1(seed, site0 + 2, (uint64_t)m * H2 + (uint64_t)c0 + 1, p_drop, dscale);
    }
    const float g0v = gelu_f(v0), g1v = gelu_f(v1);
    if (m < T) {
      *reinterpret_cast<float2*>(f2 + (size_t)m * H2 + c0) = make_float2(v0, v1);
      *reinterpret_cast<float2*>(a2 + (size_t)m * H2 + c0) = make_float2(g0v, g1v);
    }
    a2s[r * (H2 + 1) + c0] = g0v;
    a2s[r * (H2 + 1) + c0 + 1] = g1v;
  }
  __syncthreads();
  // ---- fc3 + sigmoid: thread = row ----
  if (t < HR) {
    const int m = m0 + t;
    if (m < T) {
      float s = __ldg(b3);
      const float* ar = a2s + t * (H2 + 1);
#pragma unroll
      for (int k = 0; k < H2; ++k) s = fmaf(ar[k], __ldg(w3 + k), s);
      if (drop) s *= dropout_scale1(seed, site0 + 3, (uint64_t)m, p_drop, dscale);
      out[m] = sigmoid_f(s);
    }
  }
}

// Backward of the head above in one kernel.  One CTA walks four blocks of 32 utterances: per block the chain
//   hb3 = d_out sigmoid'(out) m3;  hb2 = (hb3 w3) gelu'(f2) m2;  hb1 = (hb2 W2) gelu'(f1) m1;  dx = (hb1 W1) gelu'(x)
// runs out of shared memory (W1, W2, w3 staged once per CTA), and -- unless the network is frozen (grads == NULL) -- the
// weight / bias gradients are accumulated in registers across the four blocks (25 + 4 + 1 values per thread) and leave with
// one red.global.add per element and CTA (T/128 CTAs: 24 - 47 adds per weight instead of one per 32 utterances).
constexpr int HB_BLOCKS = 4;
__global__ void __launch_bounds__(256) disc_head_bwd_kernel(
    const float* __restrict__ d_out, const float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ g0,
    const float* __restrict__ f1, const float* __restrict__ a1, const float* __restrict__ f2, const float* __restrict__ a2,
    const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ w3, float* __restrict__ dx, float* dw1,
    float* db1, float* dw2, float* db2, float* dw3, float* db3, int T, int d, float p_drop, const Seed seed_ref, uint32_t site0) {
  extern __shared__ __align__(16) float sm[];
  float* w1s = sm;                           // [H1][d]
  float* w2s = w1s + H1 * d;                 // [H2][H1]
  float* w3s = w2s + H2 * H1;                // [H2]
  float* hb3s = w3s + H2;                    // [HR]
  float* hb2s = hb3s + HR;                   // [HR][H2 + 1]
  float* hb1s = hb2s + HR * (H2 + 1);        // [HR][H1 + 1]
  float* g0s = hb1s + HR * (H1 + 1);         // [HR][d + 1]   (weight gradients only)
  float* a1s = g0s + HR * (d + 1);           // [HR][H1 + 1]
  float* a2s = a1s + HR * (H1 + 1);          // [HR][H2 + 1]
  const int t = threadIdx.x;
  const int d4 = d >> 2;
  const bool pg = dw1 != nullptr;
  for (int idx = t; idx < H1 * d4; idx += 256) reinterpret_cast<float4*>(w1s)[idx] = __ldg(reinterpret_cast<const float4*>(w1) + idx);
  for (int idx = t; idx < H2 * H1 / 4; idx += 256) reinterpret_cast<float4*>(w2s)[idx] = __ldg(reinterpret_cast<const float4*>(w2) + idx);
  if (t < H2) w3s[t] = __ldg(w3 + t);
  const bool drop = p_drop > 0.f;
  const uint64_t seed = drop ? seed_value(seed_ref) : 0ull;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  // register accumulators of the weight gradients (see P4)
  constexpr int W1_PER_T = (H1 * HD_MAX + 255) / 256;   // 32 >= 64 * d / 256
  float gw1[W1_PER_T];
#pragma unroll
  for (int i = 0; i < W1_PER_T; ++i) gw1[i] = 0.f;
  float gw2[4] = {0.f, 0.f, 0.f, 0.f}, gsm = 0.f;       // gsm: this thread's bias / w3 gradient (see P4)
  for (int sb = 0; sb < HB_BLOCKS; ++sb) {
    const int m0 = (blockIdx.x * HB_BLOCKS + sb) * HR;
    if (m0 >= T) break;                                   // uniform
    __syncthreads();                                      // the previous block's P4 reads are done (and the weights are staged)
    // ---- P0: hb3 and, for the weight gradients, the stashed activations of the block ----
    if (t < HR) {
      const int m = m0 + t;
      float v = 0.f;
      if (m < T) {
        const float pv = __ldg(out + m);
        v = __ldg(d_out + m) * pv * (1.f - pv);
        if (drop) v *= dropout_scale1(seed, site0 + 3, (uint64_t)m, p_drop, dscale);
      }
      hb3s[t] = v;
    }
    if (pg) {
      for (int idx = t; idx < HR * d4; idx += 256) {
        const int r = idx / d4, k4 = idx - r * d4, m = m0 + r;
        const float4 v = m < T ? __ldg(reinterpret_cast<const float4*>(g0 + (size_t)m * d) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float* gs = g0s + r * (d + 1) + 4 * k4;
        gs[0] = v.x; gs[1] = v.y; gs[2] = v.z; gs[3] = v.w;
      }
      for (int idx = t; idx < HR * (H1 / 4); idx += 256) {
        const int r = idx / (H1 / 4), k4 = idx - r * (H1 / 4), m = m0 + r;
        const float4 v = m < T ? __ldg(reinterpret_cast<const float4*>(a1 + (size_t)m * H1) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float* as = a1s + r * (H1 + 1) + 4 * k4;
        as[0] = v.x; as[1] = v.y; as[2] = v.z; as[3] = v.w;
      }
      for (int idx = t; idx < HR * H2; idx += 256) {
        const int r = idx / H2, k = idx - r * H2, m = m0 + r;
        a2s[r * (H2 + 1) + k] = m < T ? __ldg(a2 + (size_t)m * H2 + k) : 0.f;
      }
    }
    __syncthreads();
    // ---- P1: hb2 [HR x 16]: thread = 1 row x 2 columns ----
    {
      const int r = t >> 3, c0 = 2 * (t & 7), m = m0 + r;
      float v0 = 0.f, v1 = 0.f;
      if (m < T) {
        const float2 pre = __ldg(reinterpret_cast<const float2*>(f2 + (size_t)m * H2 + c0));
        v0 = hb3s[r] * w3s[c0] * gelu_grad_f(pre.x);
        v1 = hb3s[r] * w3s[c0 + 1] * gelu_grad_f(pre.y);
        if (drop) {
          v0 *= dropout_scale1(seed, site0 + 2, (uint64_t)m * H2 + (uint64_t)c0, p_drop, dscale);
          v1 *= dropout_scale1(seed, site0 + 2, (uint64_t)m * H2 + (uint64_t)c0 + 1, p_drop, dscale);
        }
      }
      hb2s[r * (H2 + 1) + c0] = v0;
      hb2s[r * (H2 + 1) + c0 + 1] = v1;
    }
    __syncthreads();
    // ---- P2: hb1 [HR x 64] = (hb2 W2) gelu'(f1) m1: thread = 2 rows x 4 columns ----
    {
      const int ty = t >> 4, tx = t & 15, c0 = 4 * tx;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = 2 * ty + rr, m = m0 + r;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < T) {
          const float* hr = hb2s + r * (H2 + 1);
#pragma unroll
          for (int j = 0; j < H2; ++j) {
            const float4 w = *reinterpret_cast<const float4*>(w2s + j * H1 + c0);
            v[0] = fmaf(hr[j], w.x, v[0]); v[1] = fmaf(hr[j], w.y, v[1]);
            v[2] = fmaf(hr[j], w.z, v[2]); v[3] = fmaf(hr[j], w.w, v[3]);
          }
          const float4 pre = __ldg(reinterpret_cast<const float4*>(f1 + (size_t)m * H1 + c0));
          float msk[4] = {1.f, 1.f, 1.f, 1.f};
          if (drop) dropout_scale4(seed, site0 + 1, (uint64_t)m * H1 + (uint64_t)c0, p_drop, dscale, msk);
          v[0] *= gelu_grad_f(pre.x) * msk[0]; v[1] *= gelu_grad_f(pre.y) * msk[1];
          v[2] *= gelu_grad_f(pre.z) * msk[2]; v[3] *= gelu_grad_f(pre.w) * msk[3];
        }
        float* hs = hb1s + r * (H1 + 1) + c0;
        hs[0] = v[0]; hs[1] = v[1]; hs[2] = v[2]; hs[3] = v[3];
      }
    }
    __syncthreads();
    // ---- P3: dx [HR x d] = (hb1 W1) gelu'(x): thread = 4 rows x one float4 of columns ----
    {
      const int ty = t >> 5, tx = t & 31;
      if (tx < d4) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        const float* h0 = hb1s + (4 * ty) * (H1 + 1);
#pragma unroll 4
        for (int n = 0; n < H1; ++n) {
          const float4 w = *reinterpret_cast<const float4*>(w1s + n * d + 4 * tx);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float h = h0[i * (H1 + 1) + n];
            acc[i][0] = fmaf(h, w.x, acc[i][0]); acc[i][1] = fmaf(h, w.y, acc[i][1]);
            acc[i][2] = fmaf(h, w.z, acc[i][2]); acc[i][3] = fmaf(h, w.w, acc[i][3]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = m0 + 4 * ty + i;
          if (m < T) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)m * d) + tx);
            reinterpret_cast<float4*>(dx + (size_t)m * d)[tx] =
                make_float4(acc[i][0] * gelu_grad_f(xv.x), acc[i][1] * gelu_grad_f(xv.y), acc[i][2] * gelu_grad_f(xv.z),
                            acc[i][3] * gelu_grad_f(xv.w));
          }
        }
      }
    }
    // ---- P4: weight / bias gradient partials of the block, kept in registers ----
    if (pg) {
#pragma unroll
      for (int i = 0; i < W1_PER_T; ++i) {                // dW1[n][k] = sum_r hb1[r][n] g0[r][k]
        const int idx = t + 256 * i;
        if (idx < H1 * d) {
          const int n = idx / d, k = idx - n * d;
          float s = 0.f;
#pragma unroll 8
          for (int r = 0; r < HR; ++r) s = fmaf(hb1s[r * (H1 + 1) + n], g0s[r * (d + 1) + k], s);
          gw1[i] += s;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {                       // dW2[n][k] = sum_r hb2[r][n] a1[r][k]
        const int idx = t + 256 * i, n = idx / H1, k = idx - n * H1;
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < HR; ++r) s = fmaf(hb2s[r * (H2 + 1) + n], a1s[r * (H1 + 1) + k], s);
        gw2[i] += s;
      }
      // threads 0-63: db1[t]; 64-79: db2; 80-95: dW3; 96: db3
      float s = 0.f;
      if (t < H1) { for (int r = 0; r < HR; ++r) s += hb1s[r * (H1 + 1) + t]; }
      else if (t < H1 + H2) { for (int r = 0; r < HR; ++r) s += hb2s[r * (H2 + 1) + (t - H1)]; }
      else if (t < H1 + 2 * H2) { for (int r = 0; r < HR; ++r) s = fmaf(hb3s[r], a2s[r * (H2 + 1) + (t - H1 - H2)], s); }
      else if (t == H1 + 2 * H2) { for (int r = 0; r < HR; ++r) s += hb3s[r]; }
      gsm += s;
    }
  }
  if (pg) {
#pragma unroll
    for (int i = 0; i < W1_PER_T; ++i) {
      const int idx = t + 256 * i;
      if (idx < H1 * d) atomicAdd(dw1 + idx, gw1[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(dw2 + t + 256 * i, gw2[i]);
    if (t < H1) atomicAdd(db1 + t, gsm);
    else if (t < H1 + H2) atomicAdd(db2 + (t - H1), gsm);
    else if (t < H1 + 2 * H2) atomicAdd(dw3 + (t - H1 - H2), gsm);
    else if (t == H1 + 2 * H2) atomicAdd(db3, gsm);
  }
}

}  // namespace

bool disc_head_fusable(int d, int h1, int h2) {
  static const bool off = getenv("GANFFN_NO_HEAD_FUSE") != nullptr;   // A/B switch
  return !off && h1 == H1 && h2 == H2 && d % 4 == 0 && d >= 4 && d <= HD_MAX;
}

int disc_head_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                  const float* b3, float* g0, float* f1, float* a1, float* f2, float* a2, float* out, int T, int d, float p_drop,
                  Seed seed, int site0, cudaStream_t st) {
  GANFFN_CHECK_ARG(x && w1 && b1 && w2 && b2 && w3 && b3 && g0 && f1 && a1 && f2 && a2 && out, "disc_head_fwd: null pointer");
  GANFFN_CHECK_ARG(disc_head_fusable(d, H1, H2) || (d % 4 == 0 && d <= HD_MAX), "disc_head_fwd: d=%d", d);
  const size_t smem = ((size_t)d * H1 + H1 * H2 + HR * (d + 1) + HR * (H1 + 1) + HR * (H2 + 1)) * sizeof(float);
  GANFFN_SMEM_OPTIN(disc_head_fwd_kernel, 100 * 1024);
  disc_head_fwd_kernel<<<cdiv(T, HR), 256, smem, st>>>(x, w1, b1, w2, b2, w3, b3, g0, f1, a1, f2, a2, out, T, d, p_drop, seed,
                                                      (uint32_t)site0);
  GANFFN_LAUNCHED("disc_head_fwd_kernel");
  return GANFFN_OK;
}

// grads (dw1 .. db3) all NULL: data gradient only (frozen discriminator).  Gradients are ACCUMULATED (red.global.add).
int disc_head_bwd(const float* d_out, const float* out, const float* x, const float* g0, const float* f1, const float* a1,
                  const float* f2, const float* a2, const float* w1, const float* w2, const float* w3, float* dx, float* dw1,
                  float* db1, float* dw2, float* db2, float* dw3, float* db3, int T, int d, float p_drop, Seed seed, int site0,
                  cudaStream_t st) {
  GANFFN_CHECK_ARG(d_out && out && x && f1 && f2 && w1 && w2 && w3 && dx, "disc_head_bwd: null pointer");
  GANFFN_CHECK_ARG(d % 4 == 0 && d >= 4 && d <= HD_MAX, "disc_head_bwd: d=%d", d);
  const bool pg = dw1 != nullptr;
  GANFFN_CHECK_ARG(!pg || (db1 && dw2 && db2 && dw3 && db3 && g0 && a1 && a2), "disc_head_bwd: all six gradients or none");
  const size_t smem = ((size_t)H1 * d + H2 * H1 + H2 + HR + HR * (H2 + 1) + HR * (H1 + 1) + HR * (d + 1) + HR * (H1 + 1) + HR * (H2 + 1)) *
                      sizeof(float);
  GANFFN_SMEM_OPTIN(disc_head_bwd_kernel, 100 * 1024);
  disc_head_bwd_kernel<<<cdiv(T, HR * HB_BLOCKS), 256, smem, st>>>(d_out, out, x, g0, f1, a1, f2, a2, w1, w2, w3, dx, dw1, db1, dw2, db2, dw3,
                                                                  db3, T, d, p_drop, seed, (uint32_t)site0);
  GANFFN_LAUNCHED("disc_head_bwd_kernel");
  return GANFFN_OK;
}

}  // namespace ganffn
