// tcgen05 GEMM engine with fp32 parity: 3xTF32 error-compensated products on the 5th-gen tensor
// cores (kind::tf32), fp32 accumulators in TMEM.
//
//   C[M,N] = op(A)[M,K] * op(B)[K,N]   (fp32 in, fp32 out, shared fused epilogue)
//
// Why 3xTF32: north_star asks for rtol 1e-4 against the fp32 reference through 8 post-norm layers and
// their backward; one TF32 pass (10-bit mantissa) cannot hold that.  Each operand x is split as
// x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi); the product uses hi*hi + lo*hi + hi*lo
// (the dropped lo*lo term and the rounding of lo are ~2^-21 relative), three MMAs per K-step.
//
// Structure of one CTA (one 128 x BN output tile, optional split-K slice):
//   warps 0-7  producers: coalesced 128-bit global loads of the fp32 A/B tiles (double-buffered in
//              registers), hi/lo split, conflict-free st.shared into the canonical SWIZZLE_128B UMMA
//              layouts (K-major or MN-major, so nn.Linear forward, dgrad and wgrad all read their
//              operands as they lie in HBM, no transposed copies), fence.proxy.async, mbarrier arrive.
//              Afterwards the same warps run the epilogue: tcgen05.ld (TMEM -> registers), a per-warp
//              smem transpose so global traffic is row-contiguous, fused epilogue, 128-bit stores.
//   warp 8     one elected thread issues tcgen05.mma (12 per 32-wide K block) and tcgen05.commit to
//              release smem stages / publish the accumulator; it also owns the TMEM allocation.
// TMA is deliberately not used for the operands: they must pass through registers anyway for the
// hi/lo split, and the producer warps have the time (3 MMAs per loaded byte).
#include "kernels.h"

namespace ganffn {
namespace {

constexpr int BM = 128;
constexpr int BK = 32;               // fp32 elements per K block = one 128-byte swizzle row
constexpr int NPROD = 256;           // producer / epilogue threads
constexpr int NTHREADS = NPROD + 32; // + MMA warp
constexpr int A_TILE = BM * BK * 4;  // bytes of one A tile (hi or lo)

template <int BN> struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 2 : 3;
  static constexpr int B_TILE = BN * BK * 4;
  static constexpr int STAGE = 2 * (A_TILE + B_TILE);
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct TcParams {
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  int M, N, K;
  int k_per_split;   // multiple of BK
  float* partial; int Np;
  Epilogue ep;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// Shared-memory matrix descriptor (sm_100 UMMA), version 1.
//   K-major, SWIZZLE_128B: rows of 128 B (32 tf32 along K), 8-row atoms of 1024 B, 16-byte chunk index
//             XORed with (row % 8); SBO = stride between 8-row groups.
//   MN-major, SWIZZLE_128B_BASE32B (the only MN-major layout for 32-bit operands): rows of 128 B
//             (32 elements along M/N) per k, 4-k atoms of 512 B, 32-byte chunk index XORed with (k % 4);
//             LBO = stride between 32-element M/N groups, SBO = stride between 4-k groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

// ---- operand tile movers ----------------------------------------------------------------------------------
// One tile = R rows (M or N extent) x 32 k.  CH = 16-byte chunks per producer thread.
template <int R> struct Mover {
  static constexpr int CH = R * 8 / NPROD;

  // global -> registers.  KMAJ: element (r,k) at base[(row0+r)*ld + k]; else at base[k*ld + row0 + r].
  template <bool KMAJ>
  static __device__ __forceinline__ void load(float4 (&v)[CH], const float* __restrict__ base, int ld, int row0,
                                              int rows, int k0, int kend, int t) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int q = t + i * NPROD;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (KMAJ) {
        const int r = q >> 3, c = q & 7;
        const int gr = row0 + r, gk = k0 + c * 4;
        if (gr < rows && gk < kend) v[i] = __ldg(reinterpret_cast<const float4*>(base + (size_t)gr * ld + gk));
      } else {
        const int k = q / (R / 4), mq = q % (R / 4);
        const int gk = k0 + k, gr = row0 + mq * 4;
        if (gk < kend && gr < rows) v[i] = __ldg(reinterpret_cast<const float4*>(base + (size_t)gk * ld + gr));
      }
    }
  }

  // registers -> hi/lo split -> swizzled smem tiles
  template <bool KMAJ>
  static __device__ __forceinline__ void store(const float4 (&v)[CH], uint8_t* hi, uint8_t* lo, int t) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int q = t + i * NPROD;
      uint32_t off;
      if (KMAJ) {
        const int r = q >> 3, c = q & 7;
        off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
      } else {
        const int k = q / (R / 4), mq = q % (R / 4);
        // 512-byte atoms ordered [k-group of 4][m-group of 32]: LBO = 512 B, SBO = (R/32) * 512 B
        off = (uint32_t)(((k >> 2) * (R / 32) + (mq >> 3)) * 512 + (k & 3) * 128 + ((((mq & 7) >> 1) ^ (k & 3)) << 5) +
                         ((mq & 1) << 4));
      }
      const float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      uint32_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = tf32_rna(x[j]);
        l[j] = tf32_rna(x[j] - __uint_as_float(h[j]));
      }
      *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
};

// ---- the kernel ----------------------------------------------------------------------------------------------
template <int BN, bool TA, bool TB>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(const TcParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE);
  // bars[0..S) full, [S..2S) empty, [2S] accum; then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 1);

  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);
  const int nkb = (kend - kbeg + BK - 1) / BK;

  if (t == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(bars + s), NPROD);
      mbar_init(smem_u32(bars + C::STAGES + s), 1);
    }
    mbar_init(smem_u32(bars + 2 * C::STAGES), 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 2 * BN);   // main + correction accumulators
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ================= producers =================
    using MA = Mover<BM>;
    using MB = Mover<BN>;
    float4 a0[MA::CH], b0[MB::CH], a1[MA::CH], b1[MB::CH];
    auto gload = [&](float4 (&ra)[MA::CH], float4 (&rb)[MB::CH], int kb) {
      const int k0 = kbeg + kb * BK;
      MA::template load<!TA>(ra, p.A, p.lda, m0, p.M, k0, kend, t);
      MB::template load<TB>(rb, p.B, p.ldb, n0, p.N, k0, kend, t);
    };
    auto consume = [&](const float4 (&ra)[MA::CH], const float4 (&rb)[MB::CH], int kb) {
      const int s = kb % C::STAGES;
      const uint32_t ph = (uint32_t)(kb / C::STAGES) & 1u;
      mbar_wait(smem_u32(bars + C::STAGES + s), ph ^ 1u);
      uint8_t* st = smem + s * C::STAGE;
      MA::template store<!TA>(ra, st, st + A_TILE, t);
      MB::template store<TB>(rb, st + 2 * A_TILE, st + 2 * A_TILE + C::B_TILE, t);
      fence_proxy_async();
      mbar_arrive(smem_u32(bars + s));
    };
    if (nkb > 0) gload(a0, b0, 0);
    for (int kb = 0; kb < nkb; kb += 2) {
      if (kb + 1 < nkb) gload(a1, b1, kb + 1);
      consume(a0, b0, kb);
      if (kb + 2 < nkb) gload(a0, b0, kb + 2);
      if (kb + 1 < nkb) consume(a1, b1, kb + 1);
    }

    // ================= epilogue =================
    mbar_wait(smem_u32(bars + 2 * C::STAGES), 0);
    tc_fence_after();
    float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 36);   // private 32 x 36 fp32 transpose buffer
    const int quad = warp & 3, half = warp >> 2;
    const bool split = gridDim.z > 1;
#pragma unroll 1
    for (int cc = 0; cc < BN / 2; cc += 32) {
      const int col0 = half * (BN / 2) + cc;
      uint32_t r[32], rl[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0, r);
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(BN + col0), rl);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 o;
        o.x = __uint_as_float(r[4 * q]) + __uint_as_float(rl[4 * q]);
        o.y = __uint_as_float(r[4 * q + 1]) + __uint_as_float(rl[4 * q + 1]);
        o.z = __uint_as_float(r[4 * q + 2]) + __uint_as_float(rl[4 * q + 2]);
        o.w = __uint_as_float(r[4 * q + 3]) + __uint_as_float(rl[4 * q + 3]);
        *reinterpret_cast<float4*>(stage + lane * 36 + q * 4) = o;
      }
      __syncwarp();
      // not unrolled on purpose: the generic epilogue is large and eight copies of it thrashed the
      // instruction cache (ncu r1: 19% of samples stalled on no_inst inside the epilogue)
#pragma unroll 1
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + (lane >> 3);
        const int cq = (lane & 7) * 4;
        const float4 v4 = *reinterpret_cast<const float4*>(stage + rr * 36 + cq);
        float v[4] = {v4.x, v4.y, v4.z, v4.w};
        const int m = m0 + quad * 32 + rr, n = n0 + col0 + cq;
        if (!split) {
          epilogue_store4(p.ep, p.C, p.ldc, p.M, p.N, m, n, v);
        } else if (m < p.M && n < p.Np) {
          *reinterpret_cast<float4*>(p.partial + ((size_t)blockIdx.z * p.M + m) * p.Np + n) = v4;
        }
      }
      __syncwarp();
    }
    tc_fence_before();
  } else {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((TA ? 1u : 0u) << 15) | ((TB ? 0u : 1u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (uint32_t)(kb / C::STAGES) & 1u;
        mbar_wait(smem_u32(bars + s), ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * C::STAGE), a_lo = a_hi + A_TILE;
        const uint32_t b_hi = a_hi + 2 * A_TILE, b_lo = b_hi + C::B_TILE;
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          // A: K-major -> advance 32 B inside the swizzle row; MN-major (TA) -> advance one 8-k atom group
          // (one MMA = 8 k = two 4-k atom groups in the MN-major layout)
          const uint32_t ao = TA ? (uint32_t)j * 2 * (BM / 32) * 512 : (uint32_t)j * 32;
          const uint32_t bo = TB ? (uint32_t)j * 32 : (uint32_t)j * 2 * (BN / 32) * 512;
          const uint32_t a_lbo = TA ? 512 : 16, a_sbo = TA ? (BM / 32) * 512 : 1024, a_lt = TA ? 1u : 2u;
          const uint32_t b_lbo = TB ? 16 : 512, b_sbo = TB ? 1024 : (BN / 32) * 512, b_lt = TB ? 2u : 1u;
          const uint64_t dah = make_desc(a_hi + ao, a_lbo, a_sbo, a_lt), dal = make_desc(a_lo + ao, a_lbo, a_sbo, a_lt);
          const uint64_t dbh = make_desc(b_hi + bo, b_lbo, b_sbo, b_lt), dbl = make_desc(b_lo + bo, b_lbo, b_sbo, b_lt);
          // The tensor core accumulates in fp32 with truncation, so every accumulation step costs up to one
          // ulp of the running sum.  The two correction products (2^-11 of the result) go to their own
          // accumulator: the main one then sees K/8 accumulations instead of 3K/8, and the corrections' own
          // truncation error is scaled down by 2^-11.  The epilogue adds the two in round-to-nearest.
          umma_tf32(tmem_base + BN, dal, dbh, idesc, (kb | j) ? 1u : 0u);
          umma_tf32(tmem_base + BN, dah, dbl, idesc, 1u);
          umma_tf32(tmem_base, dah, dbh, idesc, (kb | j) ? 1u : 0u);
        }
        umma_commit(smem_u32(bars + C::STAGES + s));   // frees the stage when these MMAs retire
      }
      umma_commit(smem_u32(bars + 2 * C::STAGES));      // accumulator complete
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

__global__ void __launch_bounds__(256) tc_splitk_reduce_kernel(const float* __restrict__ partial, int splits, float* C,
                                                               int ldc, int M, int N, int Np, const Epilogue ep) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nq = Np >> 2;
  if (idx >= (int64_t)M * nq) return;
  const int m = (int)(idx / nq), n = (int)(idx % nq) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z = 0; z < splits; ++z) {
    const float4 v = *reinterpret_cast<const float4*>(partial + ((size_t)z * M + m) * Np + n);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float v[4] = {s.x, s.y, s.z, s.w};
  epilogue_store4(ep, C, ldc, M, N, m, n, v);
}

struct TcPlan { int bn; int splits; int kps; };

// Pick the tile width and split-K factor that minimise a simple wave model of the kernel time.
TcPlan tc_plan(int M, int N, int K) {
  const int tm = cdiv(M, BM);
  const int nkb = cdiv(K, BK);
  TcPlan best{128, 1, (int)round_up(K, BK)};
  double best_cost = 1e300;
  const int cand_splits[] = {1, 2, 3, 4, 6, 8, 12, 16, 24};
  for (int bn : {128, 256}) {
    if (bn == 256 && N <= 128) continue;
    const int tiles = tm * cdiv(N, bn);
    for (int sp : cand_splits) {
      if (sp > 1 && nkb / sp < 4) continue;
      const int kps = (int)round_up(cdiv(K, sp), BK);
      if (kps > 1024 && nkb / (sp + 1) >= 4) continue;   // cap the fp32-truncating accumulation chain at 128 steps
      const int splits = cdiv(K, kps);
      const int waves = cdiv((int64_t)tiles * splits, 148);
      const double per_kb = (bn == 256 ? 1536.0 : 768.0 * 1.15);
      double cost = waves * (cdiv(kps, BK) * per_kb + 2500.0 + bn * 12.0);
      if (splits > 1) cost += 3000.0 + (double)M * N * (splits + 1) * 4.0 / 4000.0;   // fold kernel (~4 KB/cycle)
      if (cost < best_cost) { best_cost = cost; best = TcPlan{bn, splits, kps}; }
    }
  }
  return best;
}

template <int BN>
int launch_tc(const TcParams& p, bool TA, bool TB, dim3 grid, cudaStream_t st) {
  constexpr int smem = Cfg<BN>::SMEM;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gemm_tc_kernel<BN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gemm_tc_kernel<BN, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gemm_tc_kernel<BN, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gemm_tc_kernel<BN, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_done = true;
  }
  if (!TA && TB) gemm_tc_kernel<BN, false, true><<<grid, NTHREADS, smem, st>>>(p);
  else if (!TA && !TB) gemm_tc_kernel<BN, false, false><<<grid, NTHREADS, smem, st>>>(p);
  else if (TA && !TB) gemm_tc_kernel<BN, true, false><<<grid, NTHREADS, smem, st>>>(p);
  else gemm_tc_kernel<BN, true, true><<<grid, NTHREADS, smem, st>>>(p);
  GANFFN_LAUNCHED("gemm_tc_kernel");
  return GANFFN_OK;
}

}  // namespace

bool gemm_tc_supported(bool transA, bool b_is_nk, int lda, int ldb, int ldc, int M, int N, int K, const void* A,
                       const void* B) {
  if (M < 96 || N < 64 || K < 32) return false;                       // too small for 128-row tensor tiles
  if ((lda & 3) || (ldb & 3) || (((uintptr_t)A | (uintptr_t)B) & 15)) return false;
  if (transA ? (M & 3) : (K & 3)) return false;                       // 128-bit chunks along the contiguous dim
  if (b_is_nk ? (K & 3) : (N & 3)) return false;
  (void)ldc;
  return true;
}

int64_t gemm_tc_scratch_floats(int M, int N, int K) {
  if (M < 96 || N < 64 || K < 32) return 0;
  const TcPlan pl = tc_plan(M, N, K);
  return pl.splits > 1 ? (int64_t)pl.splits * M * round_up(N, 4) : 0;
}

int gemm_tc(const float* A, int lda, bool transA, const float* B, int ldb, bool b_is_nk, float* C, int ldc, int M, int N,
            int K, const Epilogue& ep, float* scratch, int64_t scratch_floats, cudaStream_t st) {
  TcPlan pl = tc_plan(M, N, K);
  const int Np = (int)round_up(N, 4);
  if (pl.splits > 1 && (scratch == nullptr || scratch_floats < (int64_t)pl.splits * M * Np)) {
    pl.splits = 1;
    pl.kps = (int)round_up(K, BK);
  }
  TcParams p;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.M = M; p.N = N; p.K = K; p.k_per_split = pl.kps; p.partial = scratch; p.Np = Np; p.ep = ep;
  dim3 grid(cdiv(N, pl.bn), cdiv(M, BM), pl.splits);
  if (pl.bn == 256) GANFFN_TRY(launch_tc<256>(p, transA, b_is_nk, grid, st));
  else GANFFN_TRY(launch_tc<128>(p, transA, b_is_nk, grid, st));
  if (pl.splits > 1) {
    const int64_t nvec = (int64_t)M * (Np / 4);
    tc_splitk_reduce_kernel<<<cdiv(nvec, 256), 256, 0, st>>>(scratch, pl.splits, C, ldc, M, N, Np, ep);
    GANFFN_LAUNCHED("tc_splitk_reduce_kernel");
  }
  return GANFFN_OK;
}

}  // namespace ganffn
