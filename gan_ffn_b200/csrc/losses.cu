// Fusion + classifier + log-softmax, masked NLL, BCE, and the fused Adam step.
#include "common.cuh"

namespace ganffn {
namespace {

constexpr int CLS_MAXC = 8;    // classes (6 IEMOCAP, 7 MELD)
constexpr int CLS_MAXPL = 4;   // columns per lane: dh <= 128

// ---- fusion + fc + log_softmax: one warp per utterance slot ------------------------------------------
__global__ void __launch_bounds__(256) fuse_cls_fwd_kernel(const float* __restrict__ a, const float* __restrict__ v,
                                                           const float* __restrict__ t, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ fusion,
                                                           float* __restrict__ logp, int T, int dh, int C) {
  extern __shared__ float ws[];  // [C][dh]
  for (int i = threadIdx.x; i < C * dh; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps + wid; row < T; row += gridDim.x * warps) {
    float f[CLS_MAXPL];
#pragma unroll
    for (int k = 0; k < CLS_MAXPL; ++k) {
      const int c = lane + 32 * k;
      f[k] = 0.f;
      if (c < dh) {
        const size_t idx = (size_t)row * dh + c;
        f[k] = a[idx] + v[idx] + t[idx];   // same association as model.py:1444 ((a+v)+t)
        if (fusion) fusion[idx] = f[k];
      }
    }
    float logit[CLS_MAXC];
    float mx = -INFINITY;
#pragma unroll
    for (int cc = 0; cc < CLS_MAXC; ++cc) {
      if (cc < C) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < CLS_MAXPL; ++k) {
          const int c = lane + 32 * k;
          if (c < dh) s = fmaf(f[k], ws[cc * dh + c], s);
        }
        logit[cc] = warp_sum(s) + bias[cc];
        mx = fmaxf(mx, logit[cc]);
      }
    }
    float se = 0.f;
#pragma unroll
    for (int cc = 0; cc < CLS_MAXC; ++cc)
      if (cc < C) se += expf(logit[cc] - mx);
    const float lz = mx + logf(se);
    if (lane == 0) {
#pragma unroll
      for (int cc = 0; cc < CLS_MAXC; ++cc)
        if (cc < C) logp[(size_t)row * C + cc] = logit[cc] - lz;
    }
  }
}

// dlogits = dlp - softmax * sum(dlp); d_fusion = dlogits @ W; per-block partial dW, db.
__global__ void __launch_bounds__(256) fuse_cls_bwd_kernel(const float* __restrict__ dlp, const float* __restrict__ logp,
                                                           const float* __restrict__ fusion, const float* __restrict__ w,
                                                           float* __restrict__ d_fusion, float* __restrict__ partial,
                                                           int T, int dh, int C) {
  extern __shared__ float sm[];  // ws [C][dh]  then  red [warps][C*dh + C]
  float* ws = sm;
  float* red = sm + C * dh;
  for (int i = threadIdx.x; i < C * dh; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float dw[CLS_MAXC][CLS_MAXPL];
  float dbias[CLS_MAXC];
#pragma unroll
  for (int cc = 0; cc < CLS_MAXC; ++cc) {
    dbias[cc] = 0.f;
#pragma unroll
    for (int k = 0; k < CLS_MAXPL; ++k) dw[cc][k] = 0.f;
  }
  for (int row = blockIdx.x * warps + wid; row < T; row += gridDim.x * warps) {
    float g[CLS_MAXC];
    float gs = 0.f;
#pragma unroll
    for (int cc = 0; cc < CLS_MAXC; ++cc) {
      g[cc] = cc < C ? dlp[(size_t)row * C + cc] : 0.f;
      gs += g[cc];
    }
#pragma unroll
    for (int cc = 0; cc < CLS_MAXC; ++cc)
      if (cc < C) {
        g[cc] -= expf(logp[(size_t)row * C + cc]) * gs;
        dbias[cc] += g[cc];
      }
#pragma unroll
    for (int k = 0; k < CLS_MAXPL; ++k) {
      const int c = lane + 32 * k;
      if (c < dh) {
        const float f = fusion[(size_t)row * dh + c];
        float dx = 0.f;
#pragma unroll
        for (int cc = 0; cc < CLS_MAXC; ++cc)
          if (cc < C) {
            dx = fmaf(g[cc], ws[cc * dh + c], dx);
            dw[cc][k] = fmaf(g[cc], f, dw[cc][k]);
          }
        d_fusion[(size_t)row * dh + c] = dx;
      }
    }
  }
  const int n = C * dh + C;
  float* mine = red + (size_t)wid * n;
#pragma unroll
  for (int cc = 0; cc < CLS_MAXC; ++cc)
    if (cc < C) {
#pragma unroll
      for (int k = 0; k < CLS_MAXPL; ++k) {
        const int c = lane + 32 * k;
        if (c < dh) mine[cc * dh + c] = dw[cc][k];
      }
      if (lane == 0) mine[C * dh + cc] = dbias[cc];   // every lane holds the same dbias
    }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int ww = 0; ww < warps; ++ww) s += red[(size_t)ww * n + i];
    partial[(size_t)blockIdx.x * n + i] = s;
  }
}

__global__ void __launch_bounds__(256) fold2_kernel(const float* __restrict__ partial, int nblk, int n, float* out0,
                                                    float* out1, int split, int accumulate) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < n)
    for (int b = ty; b < nblk; b += 8) s += partial[(size_t)b * n + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][tx];
    float* dst = c < split ? out0 + c : out1 + (c - split);
    *dst = accumulate ? *dst + t : t;
  }
}

// ---- block reduction helper ------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_1024(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

// ---- MaskedNLLLoss (model.py:68-81): single deterministic block ---------------------------------------------
__global__ void __launch_bounds__(1024) masked_nll_fwd_kernel(const float* __restrict__ pred,
                                                              const int64_t* __restrict__ target,
                                                              const float* __restrict__ mask,
                                                              const float* __restrict__ weight, float* __restrict__ out,
                                                              int64_t n, int C, float den_override) {
  __shared__ float red[32];
  float num = 0.f, den = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t t = target[i];
    const float wm = (weight ? weight[t] : 1.f) * mask[i];
    num -= wm * pred[i * C + t];
    den += wm;
  }
  num = block_sum_1024(num, red);
  den = block_sum_1024(den, red);
  if (threadIdx.x == 0) {
    if (den_override > 0.f) den = den_override;
    out[0] = num / den;
    out[1] = den;
  }
}

__global__ void __launch_bounds__(256) masked_nll_bwd_kernel(const float* __restrict__ d_loss,
                                                             const float* __restrict__ loss_and_den,
                                                             const int64_t* __restrict__ target,
                                                             const float* __restrict__ mask,
                                                             const float* __restrict__ weight, float* __restrict__ d_pred,
                                                             int64_t n, int C) {
  const float g = d_loss[0] / loss_and_den[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * C; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i % C);
    const int64_t t = target[r];
    d_pred[i] = (c == t) ? -(weight ? weight[t] : 1.f) * mask[r] * g : 0.f;
  }
}

// ---- BCELoss -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bce_fwd_kernel(const float* __restrict__ prob, const float* __restrict__ target,
                                                       float* __restrict__ loss, int64_t n, float scale) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float p = prob[i], y = target[i];
    const float lp = fmaxf(logf(p), -100.f);
    const float l1p = fmaxf(log1pf(-p), -100.f);
    s -= y * lp + (1.f - y) * l1p;
  }
  s = block_sum_1024(s, red);
  if (threadIdx.x == 0) loss[0] = s / (float)n * scale;
}

__global__ void __launch_bounds__(256) bce_bwd_kernel(const float* __restrict__ d_loss, const float* __restrict__ prob,
                                                      const float* __restrict__ target, float* __restrict__ d_prob,
                                                      int64_t n, float scale) {
  const float g = d_loss[0] * scale / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = prob[i], y = target[i];
    d_prob[i] = g * (p - y) / fmaxf((1.f - p) * p, 1e-12f);
  }
}

// ---- Adam over a flat arena: 28 B/param of HBM traffic, float4 ---------------------------------------------
// step_dev == nullptr: lr_bc1 / inv_sqrt_bc2 are the host-computed bias corrections.  Otherwise lr_bc1 carries the
// plain learning rate and the corrections are derived here from the device-resident step count (graph replay).
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float lr_bc1,
                                                   float inv_sqrt_bc2, float b1, float b2, float eps, float wd,
                                                   float gscale, const int* __restrict__ step_dev) {
  if (step_dev != nullptr) {
    // the fp64 pow / sqrt of the bias corrections once per block, not once per thread
    __shared__ float s_bc[2];
    if (threadIdx.x == 0) {
      const double step = (double)__ldg(step_dev);
      s_bc[0] = (float)((double)lr_bc1 / (1.0 - pow((double)b1, step)));
      s_bc[1] = (float)(1.0 / sqrt(1.0 - pow((double)b2, step)));
    }
    __syncthreads();
    lr_bc1 = s_bc[0];
    inv_sqrt_bc2 = s_bc[1];
  }
  const int64_t nvec = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto update = [&](float4& pp, const float4& gg, float4& mm, float4& vv) {
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ga[j] * gscale + wd * pa[j];
      ma[j] = b1 * ma[j] + (1.f - b1) * gr;
      va[j] = b2 * va[j] + (1.f - b2) * gr * gr;
      pa[j] -= lr_bc1 * ma[j] / (sqrtf(va[j]) * inv_sqrt_bc2 + eps);
    }
  };
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < nvec; i += 2 * stride) {   // two float4 per array in flight per thread (8 x 16 B requests)
    const int64_t i2 = i + stride;
    float4 p0 = reinterpret_cast<float4*>(p)[i], p1 = reinterpret_cast<float4*>(p)[i2];
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + i), g1 = __ldg(reinterpret_cast<const float4*>(g) + i2);
    float4 m0 = reinterpret_cast<float4*>(m)[i], m1 = reinterpret_cast<float4*>(m)[i2];
    float4 v0 = reinterpret_cast<float4*>(v)[i], v1 = reinterpret_cast<float4*>(v)[i2];
    update(p0, g0, m0, v0);
    update(p1, g1, m1, v1);
    reinterpret_cast<float4*>(p)[i] = p0; reinterpret_cast<float4*>(m)[i] = m0; reinterpret_cast<float4*>(v)[i] = v0;
    reinterpret_cast<float4*>(p)[i2] = p1; reinterpret_cast<float4*>(m)[i2] = m1; reinterpret_cast<float4*>(v)[i2] = v1;
  }
  for (; i < nvec; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    update(pp, gg, mm, vv);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (n % 4), handled by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (nvec << 2) + threadIdx.x;
    const float gr = g[i] * gscale + wd * p[i];
    m[i] = b1 * m[i] + (1.f - b1) * gr;
    v[i] = b2 * v[i] + (1.f - b2) * gr * gr;
    p[i] -= lr_bc1 * m[i] / (sqrtf(v[i]) * inv_sqrt_bc2 + eps);
  }
}

}  // namespace

int fuse_cls_fwd(const float* a, const float* v, const float* t, const float* w, const float* b, float* fusion,
                 float* logp, int T, int dh, int C, cudaStream_t st) {
  GANFFN_CHECK_ARG(T > 0 && dh > 0 && dh <= 32 * CLS_MAXPL && C > 0 && C <= CLS_MAXC,
                   "fuse_cls: dh=%d (<=128) C=%d (<=8) unsupported", dh, C);
  const int grid = min(cdiv(T, 8), 148 * 4);
  fuse_cls_fwd_kernel<<<grid, 256, (size_t)C * dh * sizeof(float), st>>>(a, v, t, w, b, fusion, logp, T, dh, C);
  GANFFN_LAUNCHED("fuse_cls_fwd_kernel");
  return GANFFN_OK;
}

static int cls_bwd_blocks(int T) { return min(cdiv(T, 8), 148); }

int64_t fuse_cls_scratch_floats(int T, int dh, int C) { return (int64_t)cls_bwd_blocks(T) * (C * dh + C); }

int fuse_cls_bwd(const float* dlp, const float* logp, const float* fusion, const float* w, float* d_fusion, float* dw,
                 float* db, int T, int dh, int C, int accumulate, float* scratch, cudaStream_t st) {
  GANFFN_CHECK_ARG(T > 0 && dh > 0 && dh <= 32 * CLS_MAXPL && C > 0 && C <= CLS_MAXC,
                   "fuse_cls: dh=%d (<=128) C=%d (<=8) unsupported", dh, C);
  GANFFN_CHECK_ARG(scratch != nullptr, "fuse_cls_bwd: scratch is null");
  const int grid = cls_bwd_blocks(T);
  const int n = C * dh + C;
  const size_t smem = ((size_t)C * dh + (size_t)8 * n) * sizeof(float);
  fuse_cls_bwd_kernel<<<grid, 256, smem, st>>>(dlp, logp, fusion, w, d_fusion, scratch, T, dh, C);
  GANFFN_LAUNCHED("fuse_cls_bwd_kernel");
  fold2_kernel<<<cdiv(n, 32), 256, 0, st>>>(scratch, grid, n, dw, db, C * dh, accumulate);
  GANFFN_LAUNCHED("fold2_kernel");
  return GANFFN_OK;
}

int masked_nll_fwd(const float* pred, const int64_t* target, const float* mask, const float* weight, float* out,
                   int64_t n, int C, float den_override, cudaStream_t st) {
  GANFFN_CHECK_ARG(n > 0 && C > 0, "masked_nll: empty input");
  masked_nll_fwd_kernel<<<1, 1024, 0, st>>>(pred, target, mask, weight, out, n, C, den_override);
  GANFFN_LAUNCHED("masked_nll_fwd_kernel");
  return GANFFN_OK;
}

int masked_nll_bwd(const float* d_loss, const float* loss_and_den, const int64_t* target, const float* mask,
                   const float* weight, float* d_pred, int64_t n, int C, cudaStream_t st) {
  GANFFN_CHECK_ARG(n > 0 && C > 0, "masked_nll: empty input");
  masked_nll_bwd_kernel<<<(int)std::min<int64_t>(cdiv(n * C, 256), 148 * 4), 256, 0, st>>>(d_loss, loss_and_den, target, mask,
                                                                                     weight, d_pred, n, C);
  GANFFN_LAUNCHED("masked_nll_bwd_kernel");
  return GANFFN_OK;
}

int bce_fwd(const float* prob, const float* target, float* loss, int64_t n, float scale, cudaStream_t st) {
  GANFFN_CHECK_ARG(n > 0, "bce: empty input");
  bce_fwd_kernel<<<1, 1024, 0, st>>>(prob, target, loss, n, scale);
  GANFFN_LAUNCHED("bce_fwd_kernel");
  return GANFFN_OK;
}

int bce_bwd(const float* d_loss, const float* prob, const float* target, float* d_prob, int64_t n, float scale,
            cudaStream_t st) {
  GANFFN_CHECK_ARG(n > 0, "bce: empty input");
  bce_bwd_kernel<<<(int)std::min<int64_t>(cdiv(n, 256), 148 * 4), 256, 0, st>>>(d_loss, prob, target, d_prob, n, scale);
  GANFFN_LAUNCHED("bce_bwd_kernel");
  return GANFFN_OK;
}

int adam_step(float* p, const float* g, float* m, float* v, int64_t n, int step, float lr, float b1, float b2, float eps,
              float wd, float gscale, cudaStream_t st) {
  GANFFN_CHECK_ARG(n > 0 && step >= 1, "adam: n=%lld step=%d", (long long)n, step);
  GANFFN_CHECK_ARG(((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0,
                   "adam: arenas must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  const int grid = (int)std::min<int64_t>(cdiv((n >> 2) + 1, 256), 148 * 8);
  adam_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n, (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), b1, b2, eps, wd,
                                    gscale, nullptr);
  GANFFN_LAUNCHED("adam_kernel");
  return GANFFN_OK;
}

int adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const int* step_dev, float lr, float b1,
                  float b2, float eps, float wd, float gscale, cudaStream_t st) {
  GANFFN_CHECK_ARG(n > 0, "adam: n=%lld", (long long)n);
  GANFFN_CHECK_ARG(((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0,
                   "adam: arenas must be 16-byte aligned");
  const int grid = (int)std::min<int64_t>(cdiv((n >> 2) + 1, 256), 148 * 8);
  adam_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n, lr, 1.0f, b1, b2, eps, wd, gscale, step_dev);
  GANFFN_LAUNCHED("adam_kernel");
  return GANFFN_OK;
}

}  // namespace ganffn
