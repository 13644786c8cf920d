// Row-wise and element-wise kernels: LayerNorm fwd/bwd, positional encoding, GELU/dropout
// element-wise passes, bias-gradient column sums, dropout-mask export.  All HBM-bound: float4
// accesses, one warp per row for the row reductions.
#include "common.cuh"

namespace ganffn {
namespace {

constexpr int LN_MAXV = 4;   // float4 chunks per lane: d <= 4*32*4 = 512
constexpr float LN_EPS = 1e-5f;

// ---- LayerNorm forward: one warp per RPW rows -----------------------------------------------------------------
// MAXV float4 per lane (d <= 128 * MAXV), RPW rows per warp in flight.  d <= 128 (the five d=100 networks) runs as
// <1, 4>: the four rows' loads are issued together, so a warp pays one memory latency for four rows and the grid
// shrinks from 376 to 94 CTAs -- the first version (<4, 1>, 147 registers, one row per warp) held every SM for 4 us
// (forward) / 10 us (backward) per call, 11 % of the train step's SM time for 3.6 MB of traffic per call.
template <int MAXV, int RPW>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y, int T,
                                                            int d) {
  const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = d >> 2;
  float4 gm[MAXV], bt[MAXV];
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int c = lane + 32 * k;
    gm[k] = c < nv ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[k] = c < nv ? __ldg(reinterpret_cast<const float4*>(beta) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row0 = (blockIdx.x * warps + w) * RPW; row0 < T; row0 += gridDim.x * warps * RPW) {
    float4 v[RPW][MAXV];
    float sum[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      sum[r] = 0.f;
      const float4* zr = reinterpret_cast<const float4*>(z + (size_t)(row0 + r) * d);
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        const int c = lane + 32 * k;
        v[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < nv && row0 + r < T) v[r][k] = zr[c];
        sum[r] += v[r][k].x + v[r][k].y + v[r][k].z + v[r][k].w;
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const float mean = warp_sum(sum[r]) / (float)d;
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        const int c = lane + 32 * k;
        if (c < nv) {
          v[r][k].x -= mean; v[r][k].y -= mean; v[r][k].z -= mean; v[r][k].w -= mean;
          sq += v[r][k].x * v[r][k].x + v[r][k].y * v[r][k].y + v[r][k].z * v[r][k].z + v[r][k].w * v[r][k].w;
        }
      }
      const float rstd = rsqrtf(warp_sum(sq) / (float)d + LN_EPS);
      if (row0 + r < T) {
        float4* yr = reinterpret_cast<float4*>(y + (size_t)(row0 + r) * d);
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
          const int c = lane + 32 * k;
          if (c < nv) {
            float4 o;
            o.x = v[r][k].x * rstd * gm[k].x + bt[k].x;
            o.y = v[r][k].y * rstd * gm[k].y + bt[k].y;
            o.z = v[r][k].z * rstd * gm[k].z + bt[k].z;
            o.w = v[r][k].w * rstd * gm[k].w + bt[k].w;
            yr[c] = o;
          }
        }
      }
    }
  }
}

// ---- LayerNorm backward -----------------------------------------------------------------------------
constexpr int LN_BWD_MAX_BLOCKS = 148 * 2;
__device__ float g_ln_partial[LN_BWD_MAX_BLOCKS * 3 * 512];   // deterministic mode only (see the kernel's tail)
__device__ unsigned g_ln_counter = 0u;
// dz = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dy*gamma.  Each block folds its warps in shared memory
// and adds its dgamma / dbeta / sublayer-bias-gradient partials to the (pre-zeroed or accumulating) outputs
// with red.global.add: no partial buffer and no second kernel (r1 launch list: 836 fold launches per step).
template <int MAXV, int RPW>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                            const float* __restrict__ gamma, float* __restrict__ dz,
                                                            float* __restrict__ dz_drop, float* dgamma, float* dbeta,
                                                            float* dbias_sub, int T, int d, float p_drop, const Seed seed_ref,
                                                            uint32_t site, int det) {
  extern __shared__ __align__(16) float sm[];  // [warps][3][d]: dgamma, dbeta, colsum(dz after dropout)
  const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = d >> 2;
  float4 dg[MAXV], db[MAXV], gm[MAXV], ds[MAXV];
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    dg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    ds[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = lane + 32 * k;
    gm[k] = c < nv ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const bool drop = (dz_drop != nullptr) && p_drop > 0.f;
  const uint64_t seed = drop ? seed_value(seed_ref) : 0ull;
  const float dscale = drop ? 1.f / (1.f - p_drop) : 1.f;

  for (int row0 = (blockIdx.x * warps + w) * RPW; row0 < T; row0 += gridDim.x * warps * RPW) {
    float4 v[RPW][MAXV], g[RPW][MAXV];
    float sum[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      sum[r] = 0.f;
      const float4* zr = reinterpret_cast<const float4*>(z + (size_t)(row0 + r) * d);
      const float4* gr = reinterpret_cast<const float4*>(dy + (size_t)(row0 + r) * d);
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        const int c = lane + 32 * k;
        v[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        g[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < nv && row0 + r < T) { v[r][k] = zr[c]; g[r][k] = gr[c]; }
        sum[r] += v[r][k].x + v[r][k].y + v[r][k].z + v[r][k].w;
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int row = row0 + r;
      const float mean = warp_sum(sum[r]) / (float)d;
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        const int c = lane + 32 * k;
        if (c < nv) {
          v[r][k].x -= mean; v[r][k].y -= mean; v[r][k].z -= mean; v[r][k].w -= mean;
          sq += v[r][k].x * v[r][k].x + v[r][k].y * v[r][k].y + v[r][k].z * v[r][k].z + v[r][k].w * v[r][k].w;
        }
      }
      const float rstd = rsqrtf(warp_sum(sq) / (float)d + LN_EPS);
      float s1 = 0.f, s2 = 0.f;
      if (row < T) {
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
          const int c = lane + 32 * k;
          if (c < nv) {
            float4& x = v[r][k];
            float4& q = g[r][k];
            x.x *= rstd; x.y *= rstd; x.z *= rstd; x.w *= rstd;  // xhat
            dg[k].x += q.x * x.x; dg[k].y += q.y * x.y; dg[k].z += q.z * x.z; dg[k].w += q.w * x.w;
            db[k].x += q.x; db[k].y += q.y; db[k].z += q.z; db[k].w += q.w;
            q.x *= gm[k].x; q.y *= gm[k].y; q.z *= gm[k].z; q.w *= gm[k].w;
            s1 += q.x + q.y + q.z + q.w;
            s2 += q.x * x.x + q.y * x.y + q.z * x.z + q.w * x.w;
          }
        }
      }
      s1 = warp_sum(s1) / (float)d;
      s2 = warp_sum(s2) / (float)d;
      if (row < T) {
        float4* outr = reinterpret_cast<float4*>(dz + (size_t)row * d);
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
          const int c = lane + 32 * k;
          if (c < nv) {
            const float4 x = v[r][k], q = g[r][k];
            float4 o;
            o.x = rstd * (q.x - s1 - x.x * s2);
            o.y = rstd * (q.y - s1 - x.y * s2);
            o.z = rstd * (q.z - s1 - x.z * s2);
            o.w = rstd * (q.w - s1 - x.w * s2);
            outr[c] = o;
            if (drop) {
              float m[4];
              dropout_scale4(seed, site, (uint64_t)row * d + 4 * c, p_drop, dscale, m);
              o = make_float4(o.x * m[0], o.y * m[1], o.z * m[2], o.w * m[3]);
              reinterpret_cast<float4*>(dz_drop + (size_t)row * d)[c] = o;
            }
            ds[k].x += o.x; ds[k].y += o.y; ds[k].z += o.z; ds[k].w += o.w;
          }
        }
      }
    }
  }
  // fold the block's warps
  float* mine = sm + (size_t)w * 3 * d;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int c = lane + 32 * k;
    if (c < nv) {
      reinterpret_cast<float4*>(mine)[c] = dg[k];
      reinterpret_cast<float4*>(mine + d)[c] = db[k];
      reinterpret_cast<float4*>(mine + 2 * d)[c] = ds[k];
    }
  }
  __syncthreads();
  if (!det) {
    for (int c = threadIdx.x; c < 3 * d; c += blockDim.x) {
      const int seg = c / d;
      float* out = seg == 0 ? dgamma : seg == 1 ? dbeta : dbias_sub;
      if (out == nullptr) continue;
      float s = 0.f;
      for (int ww = 0; ww < warps; ++ww) s += sm[(size_t)ww * 3 * d + c];
      atomicAdd(out + (c - seg * d), s);
    }
    return;
  }
  // Deterministic mode (ganffn_set_deterministic): block partials go to a static device buffer, the block that
  // finishes last folds them in block order and does the one read-modify-write per output element.  The buffer is
  // per device, so at most one LayerNorm backward may be in flight (deterministic mode runs networks serially).
  for (int c = threadIdx.x; c < 3 * d; c += blockDim.x) {
    float s = 0.f;
    for (int ww = 0; ww < warps; ++ww) s += sm[(size_t)ww * 3 * d + c];
    g_ln_partial[(size_t)blockIdx.x * 3 * d + c] = s;
  }
  __threadfence();
  __syncthreads();
  // (no static __shared__ here: at d = 512 the dynamic buffer already is the whole 48 KB default limit)
  const unsigned ticket = threadIdx.x == 0 ? atomicAdd(&g_ln_counter, 1u) : 0u;
  if (!__syncthreads_or(threadIdx.x == 0 && ticket == gridDim.x - 1)) return;
  __threadfence();
  for (int c = threadIdx.x; c < 3 * d; c += blockDim.x) {
    const int seg = c / d;
    float* out = seg == 0 ? dgamma : seg == 1 ? dbeta : dbias_sub;
    if (out == nullptr) continue;
    float s = 0.f;
    for (unsigned bb = 0; bb < gridDim.x; ++bb) s += __ldcg(&g_ln_partial[(size_t)bb * 3 * d + c]);
    out[c - seg * d] += s;
  }
  if (threadIdx.x == 0) g_ln_counter = 0u;
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
  if (d <= 128) {
    layernorm_fwd_kernel<1, 4><<<min(cdiv(T, 8 * 4), 148 * 8), 256, 0, st>>>(z, gamma, beta, y, T, d);
  } else {
    layernorm_fwd_kernel<LN_MAXV, 1><<<min(cdiv(T, 8), 148 * 8), 256, 0, st>>>(z, gamma, beta, y, T, d);
  }
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
  const size_t smem = (size_t)8 * 3 * d * sizeof(float);
  if (d <= 128) {
    layernorm_bwd_kernel<1, 4><<<min(cdiv(T, 8 * 4), 148 * 2), 256, smem, st>>>(dy, z, gamma, dz, dz_drop, dgamma, dbeta, dbias_sub,
                                                                             T, d, p, seed, (uint32_t)site, g_deterministic);
  } else {
    layernorm_bwd_kernel<LN_MAXV, 1><<<ln_bwd_blocks(T), 256, smem, st>>>(dy, z, gamma, dz, dz_drop, dgamma, dbeta, dbias_sub, T, d,
                                                                        p, seed, (uint32_t)site, g_deterministic);
  }
  GANFFN_LAUNCHED("layernorm_bwd_kernel");
  return GANFFN_OK;
}

int layernorm_bwd_after(const Epilogue& ep, const float* dy, int M, int N, cudaStream_t st) {
  return layernorm_bwd(dy, ep.lnb_z, ep.lnb_gamma, ep.lnb_dz, ep.lnb_p > 0.f ? ep.lnb_dzd : nullptr, ep.lnb_dgamma, ep.lnb_dbeta,
                       ep.lnb_dbias, M, N, 1, ep.lnb_p, ep.lnb_seed, (int)ep.lnb_site, st);
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
  // deterministic mode: one block per 32-column strip sums all rows in a fixed order (the kernel's red.global.add
  // is then the only writer of its element)
  const int yb = g_deterministic ? 1 : colsum_rowblocks(M, N);
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
