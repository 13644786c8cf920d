// fp32 FFMA GEMM with fused epilogue and deterministic split-K.
//
// This is the exact-fp32 engine: it is the GPU-side ground truth for the tcgen05 3xTF32
// engine (gemm_tc.cu) and the engine for shapes too small or too ragged for tensor tiles
// (N = 1, 6, 16, 64; see DESIGN.md "GEMM engines").
#include "common.cuh"

namespace ganffn {

namespace {

constexpr int BK = 16;
constexpr int NT = 256;

struct GemmParams {
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  int M, N, K;
  int k_per_split;   // multiple of BK
  float* partial;    // [splits, M, Np] when gridDim.z > 1
  int Np;
  bool vecA, vecB;
  Epilogue ep;
};

// (row, col..col+3) of a row-major matrix with bounds; zero outside.
__device__ __forceinline__ float4 load4(const float* __restrict__ base, int row, int ld, int col, int nrows, int ncols,
                                        bool vec_ok) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= nrows || col >= ncols) return r;
  const float* p = base + (size_t)row * ld + col;
  if (vec_ok && col + 3 < ncols) return __ldg(reinterpret_cast<const float4*>(p));
  r.x = __ldg(p);
  if (col + 1 < ncols) r.y = __ldg(p + 1);
  if (col + 2 < ncols) r.z = __ldg(p + 2);
  if (col + 3 < ncols) r.w = __ldg(p + 3);
  return r;
}

template <int BM, int BN, bool TA, bool TB>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const GemmParams p) {
  constexpr int TM = BM / 16, TN = BN / 16;  // per-thread micro tile
  constexpr int GA = TM / 4, GB = TN / 4;    // groups of 4 rows / cols
  constexpr int LA = BM / 64, LB = BN / 64;  // float4 loads per thread per tile
  constexpr int SA = BM + 4, SB = BN + 4;

  __shared__ __align__(16) float As[2][BK][SA];
  __shared__ __align__(16) float Bs[2][BK][SB];

  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);
  const int ntiles = (kend - kbeg + BK - 1) / BK;

  float acc[GA][4][GB][4];
#pragma unroll
  for (int a = 0; a < GA; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int b = 0; b < GB; ++b)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[a][i][b][j] = 0.f;

  float4 ra[LA], rb[LB];

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int v = t + i * NT;
      if (!TA) {  // A[m, k], vector along k
        int m = v >> 2, kq = v & 3;
        ra[i] = load4(p.A, m0 + m, p.lda, k0 + kq * 4, p.M, kend, p.vecA);
      } else {    // A[k, m], vector along m
        int k = v / (BM / 4), mq = v % (BM / 4);
        ra[i] = load4(p.A, k0 + k, p.lda, m0 + mq * 4, kend, p.M, p.vecA);
      }
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int v = t + i * NT;
      if (TB) {   // B[n, k], vector along k
        int n = v >> 2, kq = v & 3;
        rb[i] = load4(p.B, n0 + n, p.ldb, k0 + kq * 4, p.N, kend, p.vecB);
      } else {    // B[k, n], vector along n
        int k = v / (BN / 4), nq = v % (BN / 4);
        rb[i] = load4(p.B, k0 + k, p.ldb, n0 + nq * 4, kend, p.N, p.vecB);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int v = t + i * NT;
      if (!TA) {
        int m = v >> 2, kq = v & 3;
        As[buf][kq * 4 + 0][m] = ra[i].x; As[buf][kq * 4 + 1][m] = ra[i].y;
        As[buf][kq * 4 + 2][m] = ra[i].z; As[buf][kq * 4 + 3][m] = ra[i].w;
      } else {
        int k = v / (BM / 4), mq = v % (BM / 4);
        *reinterpret_cast<float4*>(&As[buf][k][mq * 4]) = ra[i];
      }
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int v = t + i * NT;
      if (TB) {
        int n = v >> 2, kq = v & 3;
        Bs[buf][kq * 4 + 0][n] = rb[i].x; Bs[buf][kq * 4 + 1][n] = rb[i].y;
        Bs[buf][kq * 4 + 2][n] = rb[i].z; Bs[buf][kq * 4 + 3][n] = rb[i].w;
      } else {
        int k = v / (BN / 4), nq = v % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][k][nq * 4]) = rb[i];
      }
    }
  };

  if (ntiles > 0) {
    gload(kbeg);
    sstore(0);
  }
  __syncthreads();

  for (int it = 0; it < ntiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < ntiles) gload(kbeg + (it + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[GA][4], b[GB][4];
#pragma unroll
      for (int g = 0; g < GA; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][g * 64 + ty * 4]);
        a[g][0] = v.x; a[g][1] = v.y; a[g][2] = v.z; a[g][3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < GB; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * 64 + tx * 4]);
        b[g][0] = v.x; b[g][1] = v.y; b[g][2] = v.z; b[g][3] = v.w;
      }
#pragma unroll
      for (int ga = 0; ga < GA; ++ga)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int gb = 0; gb < GB; ++gb)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[ga][i][gb][j] = fmaf(a[ga][i], b[gb][j], acc[ga][i][gb][j]);
    }
    if (it + 1 < ntiles) sstore(buf ^ 1);
    __syncthreads();
  }

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int ga = 0; ga < GA; ++ga)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ga * 64 + ty * 4 + i;
#pragma unroll
      for (int gb = 0; gb < GB; ++gb) {
        const int n = n0 + gb * 64 + tx * 4;
        float v[4] = {acc[ga][i][gb][0], acc[ga][i][gb][1], acc[ga][i][gb][2], acc[ga][i][gb][3]};
        if (!split) {
          epilogue_store4(p.ep, p.C, p.ldc, p.M, p.N, m, n, v);
        } else if (m < p.M && n < p.Np) {
          float* dst = p.partial + ((size_t)blockIdx.z * p.M + m) * p.Np + n;
          *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int splits, float* C,
                                                            int ldc, int M, int N, int Np, const Epilogue ep) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nq = Np >> 2;
  if (idx >= (int64_t)M * nq) return;
  int m, n;
  if ((int64_t)M * nq < (int64_t)1 << 31) {   // 32-bit index arithmetic (a 64-bit divide costs more than the fold itself)
    const unsigned i32 = (unsigned)idx, mq = i32 / (unsigned)nq;
    m = (int)mq; n = (int)(i32 - mq * (unsigned)nq) * 4;
  } else {
    m = (int)(idx / nq); n = (int)(idx % nq) * 4;
  }
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z = 0; z < splits; ++z) {
    float4 v = *reinterpret_cast<const float4*>(partial + ((size_t)z * M + m) * Np + n);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float v[4] = {s.x, s.y, s.z, s.w};
  epilogue_store4(ep, C, ldc, M, N, m, n, v);
}

struct Plan { int bm; int splits; int k_per_split; };

Plan make_plan(int M, int N, int K) {
  Plan pl;
  const int64_t tiles128 = (int64_t)cdiv(M, 128) * cdiv(N, 128);
  pl.bm = tiles128 >= 120 ? 128 : 64;
  const int64_t tiles = pl.bm == 128 ? tiles128 : (int64_t)cdiv(M, 64) * cdiv(N, 64);
  const int target = pl.bm == 128 ? 148 : 296;
  int splits = 1;
  if (tiles < target) {
    splits = (int)((target + tiles - 1) / tiles);
    const int max_by_k = K / (4 * BK) > 0 ? K / (4 * BK) : 1;  // >= 64 k per split
    if (splits > max_by_k) splits = max_by_k;
    if (splits > 32) splits = 32;
  }
  int kps = (int)round_up(cdiv(K, splits), BK);
  splits = cdiv(K, kps);
  pl.splits = splits < 1 ? 1 : splits;
  pl.k_per_split = kps;
  return pl;
}

template <int BM, int BN>
void launch_simt(const GemmParams& p, bool TA, bool TB, dim3 grid, cudaStream_t st) {
  if (!TA && TB) gemm_simt_kernel<BM, BN, false, true><<<grid, NT, 0, st>>>(p);
  else if (!TA && !TB) gemm_simt_kernel<BM, BN, false, false><<<grid, NT, 0, st>>>(p);
  else if (TA && !TB) gemm_simt_kernel<BM, BN, true, false><<<grid, NT, 0, st>>>(p);
  else gemm_simt_kernel<BM, BN, true, true><<<grid, NT, 0, st>>>(p);
}

}  // namespace

int64_t gemm_simt_scratch_floats(int M, int N, int K) {
  Plan pl = make_plan(M, N, K);
  return pl.splits > 1 ? (int64_t)pl.splits * M * round_up(N, 4) : 0;
}

int gemm_simt(const float* A, int lda, bool transA, const float* B, int ldb, bool b_is_nk, float* C, int ldc, int M,
              int N, int K, const Epilogue& ep, float* scratch, int64_t scratch_floats, cudaStream_t st) {
  GANFFN_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem %dx%dx%d", M, N, K);
  Plan pl = make_plan(M, N, K);
  const int Np = (int)round_up(N, 4);
  if (pl.splits > 1 && (scratch == nullptr || scratch_floats < (int64_t)pl.splits * M * Np)) {
    pl.splits = 1;  // no room for partials: run unsplit (still exact, just fewer CTAs)
    pl.k_per_split = (int)round_up(K, BK);
  }
  GemmParams p;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.M = M; p.N = N; p.K = K;
  p.k_per_split = pl.k_per_split;
  p.partial = scratch; p.Np = Np;
  p.vecA = (lda % 4 == 0) && ((((uintptr_t)A) & 15) == 0);
  p.vecB = (ldb % 4 == 0) && ((((uintptr_t)B) & 15) == 0);
  p.ep = ep;
  if (pl.bm == 128) {
    dim3 grid(cdiv(N, 128), cdiv(M, 128), pl.splits);
    launch_simt<128, 128>(p, transA, b_is_nk, grid, st);
  } else {
    dim3 grid(cdiv(N, 64), cdiv(M, 64), pl.splits);
    launch_simt<64, 64>(p, transA, b_is_nk, grid, st);
  }
  GANFFN_LAUNCHED("gemm_simt_kernel");
  if (pl.splits > 1) {
    const int64_t nvec = (int64_t)M * (Np / 4);
    splitk_reduce_kernel<<<cdiv(nvec, 256), 256, 0, st>>>(scratch, pl.splits, C, ldc, M, N, Np, ep);
    GANFFN_LAUNCHED("splitk_reduce_kernel");
  }
  return GANFFN_OK;
}

}  // namespace ganffn
