// Row-wise and element-wise kernels: LayerNorm fwd/bwd, positional encoding, GELU/dropout
// element-wise passes, bias-gradient column sums, dropout-mask export.  All HBM-bound: float4
// accesses, one warp per row for the row reductions.
#include "common.cuh"

namespace ganffn {
namespace {

constexpr int LN_MAXV = 4;   // float4 chunks per lane: d <= 4*32*4 = 512
constexpr float LN_EPS = 1e-5f;

// ---- LayerNorm forward: one warp per row ---------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y, int T,
                                                            int d) {
  const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = d >> 2;
  for (int row = blockIdx.x * warps + w; row < T; row += gridDim.x * warps) {
    const float4* zr = reinterpret_cast<const float4*>(z + (size_t)row * d);
    float4 v[LN_MAXV];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
      const int c = lane + 32 * k;
      if (c < nv) { v[k] = zr[c]; sum += v[k].x + v[k].y + v[k].z + v[k].w; }
    }
    const float mean = warp_sum(sum) / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
      const int c = lane + 32 * k;
      if (c < nv) {
        float a = v[k].x - mean, b = v[k].y - mean, e = v[k].z - mean, f = v[k].w - mean;
        sq += a * a + b * b + e * e + f * f;
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)d + LN_EPS);
    float4* yr = reinterpret_cast<float4*>(y + (size_t)row * d);
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
      const int c = lane + 32 * k;
      if (c < nv) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
        float4 r;
        r.x = (v[k].x - mean) * rstd * g.x + b.x;
        r.y = (v[k].y - mean) * rstd * g.y + b.y;
        r.z = (v[k].z - mean) * rstd * g.z + b.z;
        r.w = (v[k].w - mean) * rstd * g.w + b.w;
        yr[c] = r;
      }
    }
  }
}

// ---- LayerNorm backward -----------------------------------------------------------------------------
// dz = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dy*gamma.  Each block folds its warps in shared memory
// and adds its dgamma / dbeta / sublayer-bias-gradient partials to the (pre-zeroed or accumulating) outputs
// with red.global.add: no partial buffer and no second kernel (r1 launch list: 836 fold launches per step).
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                            const float* __restrict__ gamma, float* __restrict__ dz,
                                                            float* __restrict__ dz_drop, float* dgamma, float* dbeta,
                                                            float* dbias_sub, int T, int d, float p_drop, const Seed seed_ref,
                                                            uint32_t site) {
  extern __shared__ __align__(16) float sm[];  // [warps][3][d]: dgamma, dbeta, colsum(dz after dropout)
  const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = d >> 2;
  float4 dg[LN_MAXV], db[LN_MAXV], gm[LN_MAXV], ds[LN_MAXV];
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    dg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    ds[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = lane + 32 * k;
    gm[k] = c < nv ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const bool drop = (dz_drop != nullptr) && p_drop > 0.f;
  const uint64_t seed = drop ? seed_value(seed_ref) : 0ull;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;

  for (int row = blockIdx.x * warps + w; row < T; row += gridDim.x * warps) {
    const float4* zr = reinterpret_cast<const float4*>(z + (size_t)row * d);
    const float4* gr = reinterpret_cast<const float4*>(dy + (size_t)row * d);
    float4 v[LN_MAXV], g[LN_MAXV];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
      const int c = lane + 32 * k;
      if (c < nv) { v[k] = zr[c]; g[k] = gr[c]; sum += v[k].x + v[k].y + v[k].z + v[k].w; }
    }
    const float mean = warp_sum(sum) / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
      const int c = lane + 32 * k;
      if (c < nv) {
        v[k].x -= mean; v[k].y -= mean; v[k].z -= mean; v[k].w -= mean;
        sq += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)d + LN_EPS);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
      const int c = lane + 32 * k;
      if (c < nv) {
        v[k].x *= rstd; v[k].y *= rstd; v[k].z *= rstd; v[k].w *= rstd;  // xhat
        dg[k].x += g[k].x * v[k].x; dg[k].y += g[k].y * v[k].y; dg[k].z += g[k].z * v[k].z; dg[k].w += g[k].w * v[k].w;
        db[k].x += g[k].x; db[k].y += g[k].y; db[k].z += g[k].z; db[k].w += g[k].w;
        g[k].x *= gm[k].x; g[k].y *= gm[k].y; g[k].z *= gm[k].z; g[k].w *= gm[k].w;
        s1 += g[k].x + g[k].y + g[k].z + g[k].w;
        s2 += g[k].x * v[k].x + g[k].y * v[k].y + g[k].z * v[k].z + g[k].w * v[k].w;
      }
    }
    s1 = warp_sum(s1) / (float)d;
    s2 = warp_sum(s2) / (float)d;
    float4* outr = reinterpret_cast<float4*>(dz + (size_t)row * d);
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
      const int c = lane + 32 * k;
      if (c < nv) {
        float4 r;
        r.x = rstd * (g[k].x - s1 - v[k].x * s2);
        r.y = rstd * (g[k].y - s1 - v[k].y * s2);
        r.z = rstd * (g[k].z - s1 - v[k].z * s2);
        r.w = rstd * (g[k].w - s1 - v[k].w * s2);
        outr[c] = r;
        if (drop) {
          float m[4];
          dropout_scale4(seed, site, (uint64_t)row * d + 4 * c, p_drop, dscale, m);
          r = make_float4(r.x * m[0], r.y * m[1], r.z * m[2], r.w * m[3]);
          reinterpret_cast<float4*>(dz_drop + (size_t)row * d)[c] = r;
        }
        ds[k].x += r.x; ds[k].y += r.y; ds[k].z += r.z; ds[k].w += r.w;
      }
    }
  }
  // fold the block's warps
  float* mine = sm + (size_t)w * 3 * d;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    const int c = lane + 32 * k;
    if (c < nv) {
      reinterpret_cast<float4*>(mine)[c] = dg[k];
      reinterpret_cast<float4*>(mine + d)[c] = db[k];
      reinterpret_cast<float4*>(mine + 2 * d)[c] = ds[k];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * d; c += blockDim.x) {
    const int seg = c / d;
    float* out = seg == 0 ? dgamma : seg == 1 ? dbeta : dbias_sub;
    if (out == nullptr) continue;
    float s = 0.f;
    for (int ww = 0; ww < warps; ++ww) s += sm[(size_t)ww * 3 * d + c];
    atomicAdd(out + (c - seg * d), s);
  }
}

// ---- positional encoding ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) posenc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pe,
                                                         float* __restrict__ y, int64_t nvec, int B, int d, float p_drop,
                                                         const Seed seed_ref) {
  const bool drop = p_drop > 0.f;
  const uint64_t seed = drop ? seed_value(seed_ref) : 0ull;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  const int dv = d >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / dv;
    const int c = (int)(i % dv);
    const int s = (int)(row / B);
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pe + (size_t)s * d) + c);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    if (drop) {
      float m[4];
      dropout_scale4(seed, GANFFN_SITE_PE, (uint64_t)i * 4, p_drop, dscale, m);
      v.x *= m[0]; v.y *= m[1]; v.z *= m[2]; v.w *= m[3];
    }
    reinterpret_cast<float4*>(y)[i] = v;
  }
}

// ---- element-wise passes ---------------------------------------------------------------------------------
// mode 0: y = drop(gelu(x))                                  (forward, generators: model.py:1223-1226)
// mode 1: y = dy * gelu'(src) * mask                         (backward through gelu then dropout-before-act)
// mode 2: y = dy * src*(1-src) * mask                        (backward through sigmoid, src = probability)
// mode 3: y = dy * mask                                      (backward through dropout only)
__global__ void __launch_bounds__(256) elementwise_kernel(const float* __restrict__ a, const float* __restrict__ src,
                                                          float* __restrict__ y, int64_t n, int mode, float p_drop,
                                                          const Seed seed_ref, uint32_t site) {
  const bool drop = p_drop > 0.f;
  const uint64_t seed = drop ? seed_value(seed_ref) : 0ull;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;
  const int64_t nvec = (n + 3) >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 4;
    float m[4] = {1.f, 1.f, 1.f, 1.f};
    if (drop) dropout_scale4(seed, site, (uint64_t)e, p_drop, dscale, m);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (e + j < n) {
        const float av = a[e + j];
        float r;
        if (mode == 0) r = gelu_f(av) * m[j];
        else if (mode == 1) r = av * gelu_grad_f(src[e + j]) * m[j];
        else if (mode == 2) { const float pv = src[e + j]; r = av * pv * (1.f - pv) * m[j]; }
        else r = av * m[j];
        y[e + j] = r;
      }
    }
  }
}

// ---- column sums (bias gradients) ---------------------------------------------------------------------------
// grid.x covers 32-column strips, grid.y splits the rows; each block adds its partial with red.global.add.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ a, int M, int N, int rows_per_blk,
                                                     float* __restrict__ out) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_blk, r1 = min(M, r0 + rows_per_blk);
  float s = 0.f;
  if (c < N)
    for (int r = r0 + ty; r < r1; r += 8) s += __ldg(a + (size_t)r * N + c);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][tx];
    atomicAdd(out + c, t);
  }
}

__global__ void __launch_bounds__(256) dropout_mask_kernel(float* __restrict__ out, int64_t rows, int64_t cols,
                                                           int64_t row_stride, float p_drop, uint64_t seed,
                                                           uint32_t site) {
  const float dscale = 1.f / (1.f - p_drop);
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    out[i] = p_drop > 0.f ? dropout_scale1(seed, site, (uint64_t)(r * row_stride + c), p_drop, dscale) : 1.f;
  }
}

}  // namespace

int layernorm_fwd(const float* z, const float* gamma, const float* beta, float* y, int T, int d, cudaStream_t st) {
  GANFFN_CHECK_ARG(T > 0 && d > 0 && d % 4 == 0 && d <= 512, "layernorm: d=%d must be a multiple of 4 and <= 512", d);
  const int grid = min(cdiv(T, 8), 148 * 8);
  layernorm_fwd_kernel<<<grid, 256, 0, st>>>(z, gamma, beta, y, T, d);
  GANFFN_LAUNCHED("layernorm_fwd_kernel");
  return GANFFN_OK;
}

static int ln_bwd_blocks(int T) { return min(cdiv(T, 8), 148 * 2); }

// dbias_sub (optional): column sums of the gradient that flows into the sublayer branch (dz after its
// dropout mask) = the bias gradient of the linear layer that fed this LayerNorm's residual add.
// accumulate = 0 zeroes dgamma / dbeta / dbias_sub first (the network path always accumulates into the arena).
int layernorm_bwd(const float* dy, const float* z, const float* gamma, float* dz, float* dz_drop, float* dgamma,
                  float* dbeta, float* dbias_sub, int T, int d, int accumulate, float p, Seed seed, int site,
                  cudaStream_t st) {
  GANFFN_CHECK_ARG(T > 0 && d > 0 && d % 4 == 0 && d <= 512, "layernorm: d=%d must be a multiple of 4 and <= 512", d);
  if (!accumulate) {
    if (dgamma) cudaMemsetAsync(dgamma, 0, (size_t)d * sizeof(float), st);
    if (dbeta) cudaMemsetAsync(dbeta, 0, (size_t)d * sizeof(float), st);
    if (dbias_sub) cudaMemsetAsync(dbias_sub, 0, (size_t)d * sizeof(float), st);
  }
  const int grid = ln_bwd_blocks(T);
  layernorm_bwd_kernel<<<grid, 256, (size_t)8 * 3 * d * sizeof(float), st>>>(dy, z, gamma, dz, dz_drop, dgamma, dbeta,
                                                                             dbias_sub, T, d, p, seed, (uint32_t)site);
  GANFFN_LAUNCHED("layernorm_bwd_kernel");
  return GANFFN_OK;
}

int posenc_fwd(const float* x, const float* pe, float* y, int S, int B, int d, float p, Seed seed, cudaStream_t st) {
  GANFFN_CHECK_ARG(S >= 1 && S <= GANFFN_MAX_SEQ, "posenc: seq_len %d outside [1,%d] (model.py:1179)", S, GANFFN_MAX_SEQ);
  GANFFN_CHECK_ARG(d % 4 == 0, "posenc: d=%d must be a multiple of 4", d);
  const int64_t nvec = (int64_t)S * B * (d / 4);
  const int grid = (int)std::min<int64_t>(cdiv(nvec, 256), 148 * 8);
  posenc_fwd_kernel<<<grid, 256, 0, st>>>(x, pe, y, nvec, B, d, p, seed);
  GANFFN_LAUNCHED("posenc_fwd_kernel");
  return GANFFN_OK;
}

int elementwise(const float* a, const float* src, float* y, int64_t n, int mode, float p, Seed seed, int site,
                cudaStream_t st) {
  if (n <= 0) return GANFFN_OK;
  const int grid = (int)std::min<int64_t>(cdiv((n + 3) / 4, 256), 148 * 8);
  elementwise_kernel<<<grid, 256, 0, st>>>(a, src, y, n, mode, p, seed, (uint32_t)site);
  GANFFN_LAUNCHED("elementwise_kernel");
  return GANFFN_OK;
}

static int colsum_rowblocks(int M, int N) {
  const int strips = cdiv(N, 32);
  int yb = cdiv(296, strips);
  const int maxy = cdiv(M, 64);
  if (yb > maxy) yb = maxy;
  return yb < 1 ? 1 : yb;
}

// out[N] (+)= column sums of a[M,N]
int colsum(const float* a, int M, int N, float* out, int accumulate, cudaStream_t st) {
  if (!accumulate) cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st);
  const int yb = colsum_rowblocks(M, N);
  const int rpb = cdiv(M, yb);
  dim3 grid(cdiv(N, 32), yb);
  colsum_kernel<<<grid, 256, 0, st>>>(a, M, N, rpb, out);
  GANFFN_LAUNCHED("colsum_kernel");
  return GANFFN_OK;
}

int dropout_mask(float* out, int64_t rows, int64_t cols, int64_t row_stride, float p, uint64_t seed, int site,
                 cudaStream_t st) {
  GANFFN_CHECK_ARG(p >= 0.f && p < 1.f, "dropout_mask: p=%f", p);
  const int64_t n = rows * cols;
  if (n <= 0) return GANFFN_OK;
  dropout_mask_kernel<<<(int)std::min<int64_t>(cdiv(n, 256), 148 * 8), 256, 0, st>>>(out, rows, cols, row_stride, p, seed,
                                                                               (uint32_t)site);
  GANFFN_LAUNCHED("dropout_mask_kernel");
  return GANFFN_OK;
}

}  // namespace ganffn
