// tcgen05 GEMM engine with fp32 parity: 3xTF32 error-compensated products on the 5th-gen tensor
// cores (kind::tf32), fp32 accumulators in TMEM.
//
//   C[M,N] = op(A)[M,K] * op(B)[K,N]   (fp32 in, fp32 out, fused epilogue)
//
// Why 3xTF32: north_star asks for rtol 1e-4 against the fp32 reference through 8 post-norm layers and
// their backward; one TF32 pass (10-bit mantissa) cannot hold that.  Each operand x is split as
// x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi); the product uses hi*hi + lo*hi + hi*lo
// (the dropped lo*lo term and the rounding of lo are ~2^-21 relative), three MMAs per K-step.
//
// Data flow of one CTA (one 128 x 128 output tile, optional split-K slice), 17 warps:
//   warps 0-7   A producers.  Thread = one row of the tile (TMEM lane), half of a 32-wide K block.
//               global -> registers (ring of 3 K blocks in flight) -> hi/lo split -> tcgen05.st into the
//               TMEM A ring.  The A operand is read by the tensor core from TMEM (".ts" form of
//               tcgen05.mma): an SS-form 128x128x8 tf32 MMA needs 8 KB of shared-memory reads per 64
//               cycles, i.e. the whole 128 B/clk shared-memory port, leaving nothing for the producers'
//               writes (r1 in-kernel timeline: 2200 cycles per K block against 768 of MMA).  With A in
//               TMEM the port carries only B: 64 B/clk of MMA reads + 42 B/clk of producer writes.
//   warps 8-15  B producers.  global -> registers (ring of 3) -> hi/lo split -> conflict-free st.shared
//               into the canonical UMMA layouts (K-major SWIZZLE_128B or MN-major SWIZZLE_128B_BASE32B), so
//               forward (X W^T), dgrad (dY W) and wgrad (dY^T X) read W / X as they lie in HBM.
//   warp 16     one elected thread issues tcgen05.mma (3 per 8-wide K step) and tcgen05.commit; owns TMEM.
//   epilogue    all 16 producer warps: one 32x32 sub-tile each (TMEM -> registers -> smem transpose ->
//               row-contiguous 128-bit global traffic) with a mode-specialised fused epilogue.
// One full/empty mbarrier pair per stage covers both operand rings (arrivals are per warp, not per
// thread: 512 arrivals on one mbarrier word serialise for ~1000 cycles per K block).
#include <stdlib.h>
#include "kernels.h"

namespace ganffn {
namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 32;                  // fp32 elements per K block = one 128-byte swizzle row
constexpr int NAW = 8, NBW = 8;         // A / B producer warps
constexpr int NPW = NAW + NBW;
constexpr int NB = NBW * 32;            // B producer threads
constexpr int NTHREADS = (NPW + 1) * 32;
constexpr int STAGES = 4;
#ifndef GANFFN_TC_RING
#define GANFFN_TC_RING 3
#endif
constexpr int RING = GANFFN_TC_RING;    // K blocks each producer thread keeps in flight in registers
constexpr int B_TILE = BN * BK * 4;     // bytes of one B tile (hi or lo)
constexpr int STAGE = 2 * B_TILE;
constexpr int SMEM = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TM_MAIN = 0, TM_CORR = BN, TM_A = 2 * BN;   // TMEM columns; A ring: 64 per stage (32 hi | 32 lo)
constexpr int TM_COLS = 512;
static_assert(TM_A + STAGES * 64 <= TM_COLS, "TMEM budget");
static_assert(NPW * 32 * 36 * 4 <= STAGES * STAGE, "epilogue staging reuses the B ring");

enum { EPI_PLAIN = 0, EPI_DROP = 1, EPI_FULL = 2 };

#ifdef GANFFN_TC_TRACE
__device__ long long g_tc_trace[128];
#define TR(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tc_trace[i] = clock64(); } while (0)
__device__ int g_tc_dbg = 0;
#define DBG(bit) (g_tc_dbg & (bit))
#else
#define DBG(bit) 0
#define TR(i) do {} while (0)
#endif

struct TcParams {
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  int M, N, K;
  int k_per_split;   // multiple of BK
  float* partial; int Np;
  int out_mode;      // 0 fused epilogue, 1 split-K partials, 2 red.global.add (vector), 3 red.global.add (scalar)
  int x1;            // reduced-precision variant (GANFFN_GEMM_TF32X1): hi parts only, one MMA per product
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

__device__ __forceinline__ uint32_t tf32_rna(float x) { return tf32_rna_bits(x); }   // common.cuh: two integer instructions

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


// Explicit shared-space accesses: the dynamic smem base goes through integer alignment arithmetic, after which the
// compiler only knows a generic pointer and would emit generic LD/ST (seen in SASS as ST.E.128 / LD.E.128).
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// hi / lo split of two values at once: hi = rna_tf32(x) (two integer instructions each), lo = x - hi as ONE packed FFMA2
// (fma(hi, -1, x) is exact: x - hi is representable).  The producers share their schedulers with the MMA-issuing thread,
// so every instruction they do not issue shortens the main loop (§3.6 of DESIGN.md).
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& h0, uint32_t& h1, uint32_t& l0, uint32_t& l1) {
  h0 = tf32_rna(x0);
  h1 = tf32_rna(x1);
  const float2 lo = __ffma2_rn(make_float2(__uint_as_float(h0), __uint_as_float(h1)), make_float2(-1.0f, -1.0f), make_float2(x0, x1));
  l0 = __float_as_uint(lo.x);
  l1 = __float_as_uint(lo.y);
}

// A operand from tensor memory (".ts"): lanes = rows of the tile, one tf32 per 32-bit column.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 16 lanes x 16 columns.  Register i of thread t lands in lane (t/4) + 8*((i>>1)&1), column 8*(i>>2) + 2*(t%4) + (i&1)
// (probed on B200 with tools/tmem_probe.cu): a quad owns 8 consecutive bytes of a row, so a K-major operand can be
// loaded with sector-aligned 64-bit global loads and stored without any cross-lane exchange.
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- A operand: registers -> TMEM -----------------------------------------------------------------------------
// One warp = 32 rows (its TMEM lane quadrant) x 16 k (half of a K block); 16 values per thread.
//  K-major A ([M, lda], k contiguous): tcgen05.st.16x256b fragments.  v[8h + i] is row 16h + lane/4 + 8*((i>>1)&1),
//    k = 8*(i>>2) + 2*(lane%4) + (i&1): eight 64-bit loads per thread, each warp instruction reads 8 rows x one
//    full 32-byte sector.  (Thread-per-row 128-bit loads touched 32 half sectors per instruction and throttled the
//    L1 miss path to ~17 B/clk/SM; a quad-cooperative load + shuffle transpose cost 100 extra instructions per K
//    block in an issue-bound producer.)
//  MN-major A ([K, lda], m contiguous): thread = row, tcgen05.st.32x32b; the warp reads 128 contiguous bytes per k.
// `q` points at the thread's first element of this K block (see the kernel for the per-thread base pointers).
template <bool KMAJ>
__device__ __forceinline__ void load_a(float (&v)[16], const float* __restrict__ q, int lda, int rows_left, int k_left) {
  if (KMAJ) {
    const bool interior = rows_left > 24 && k_left > 8;   // every one of the eight loads is in range
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int cg = 0; cg < 2; ++cg)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          float2 x = make_float2(0.f, 0.f);
          if (interior || (16 * h + 8 * rr < rows_left && 8 * cg < k_left))
            x = __ldg(reinterpret_cast<const float2*>(q + (size_t)(16 * h + 8 * rr) * lda + 8 * cg));
          v[8 * h + 4 * cg + 2 * rr] = x.x;
          v[8 * h + 4 * cg + 2 * rr + 1] = x.y;
        }
  } else {
    if (rows_left > 0 && k_left >= 16) {   // interior block: sixteen unpredicated loads at a constant stride
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __ldg(q + (size_t)j * lda);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] = 0.f;
        if (rows_left > 0 && j < k_left) v[j] = __ldg(q + (size_t)j * lda);
      }
    }
  }
}

// hi = rna_tf32(x); lo = x - hi is exact in fp32 and handed over as is: the tensor core reads the top 19 bits of a
// tf32 operand, i.e. truncates lo (|lo| <= 2^-11 |x|, so the truncation is <= 2^-21 |x|).
// taddr: the warp's lane quadrant, first column of its k-half (hi); lo lives 32 columns further.
template <bool KMAJ>
__device__ __forceinline__ void store_a(const float (&v)[16], uint32_t taddr, int x1 = 0) {
  uint32_t h[16], l[16];
#pragma unroll
  for (int j = 0; j < 16; j += 2) split2(v[j], v[j + 1], h[j], h[j + 1], l[j], l[j + 1]);
  if (KMAJ) {
    tmem_st_16x256b_x2(taddr, h);
    tmem_st_16x256b_x2(taddr + (16u << 16), h + 8);
    if (!x1) {
      tmem_st_16x256b_x2(taddr + 32, l);
      tmem_st_16x256b_x2(taddr + (16u << 16) + 32, l + 8);
    }
  } else {
    tmem_st16(taddr, h);
    if (!x1) tmem_st16(taddr + 32, l);
  }
  tmem_st_wait();
}

// ---- B operand: registers -> swizzled shared memory --------------------------------------------------------------
// One tile = BN rows (N extent) x 32 k, 4 16-byte chunks per B-producer thread.
constexpr int BCH = BN * 8 / NB;

// Chunk i of a thread: K-major (element (n,k) at base[(row0+n)*ld + k]): row (tb>>3) + 32 i, k chunk tb&7;
// MN-major (element at base[k*ld + row0 + n]): k = tb/32 + 8 i, n chunk tb%32.  Either way the global pointer of
// chunk i is q + i*istride and its smem offset off0 + i*4096, so nothing is recomputed per K block.
template <bool KMAJ>
__device__ __forceinline__ void load_b(float4 (&v)[BCH], const float* __restrict__ q, int64_t istride, int rows_left,
                                       int k_left) {
  const bool interior = KMAJ ? (rows_left > 32 * (BCH - 1) && k_left > 0) : (rows_left > 0 && k_left > 8 * (BCH - 1));
  if (interior) {
#pragma unroll
    for (int i = 0; i < BCH; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(q + (size_t)i * istride));
    return;
  }
#pragma unroll
  for (int i = 0; i < BCH; ++i) {
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool ok = KMAJ ? (32 * i < rows_left && k_left > 0) : (rows_left > 0 && 8 * i < k_left);
    if (ok) v[i] = __ldg(reinterpret_cast<const float4*>(q + (size_t)i * istride));
  }
}

__device__ __forceinline__ void store_b(const float4 (&v)[BCH], uint32_t hi, uint32_t lo, int x1 = 0) {
#pragma unroll
  for (int i = 0; i < BCH; ++i) {
    const float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
    uint32_t h[4], l[4];
    split2(x[0], x[1], h[0], h[1], l[0], l[1]);   // lo is truncated by the tensor core (see store_a)
    split2(x[2], x[3], h[2], h[3], l[2], l[3]);
    sts128(hi + i * 4096, h[0], h[1], h[2], h[3]);
    if (!x1) sts128(lo + i * 4096, l[0], l[1], l[2], l[3]);
  }
}

// ---- lean vector epilogue for the hot flag combinations ------------------------------------------------------------
// The generic loop in epilogue_subtile() tests its flags at run time and costs ~140 instructions per float4 (SASS,
// r1); with four warps per scheduler that made the epilogue of a 128 x 128 tile 6 000 of the CTA's 16 000 cycles for
// the K = 100 products.  Here the flags are template parameters, the dropout counter advances by addition
// (SplitMix64: state += 4N/4 * gamma per row step) and two row groups are in flight per iteration.
template <bool BIAS, bool RELU, bool DROP, bool DNZ, bool RES>
__device__ __forceinline__ void epi_vec(const TcParams& p, uint32_t stage, int m_base, int n, int lane) {
  const Epilogue& ep = p.ep;
  const int cq = (lane & 7) * 4, r0 = lane >> 3;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (BIAS) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
  const int m_first = m_base + r0;
  int rows_left = p.M - m_first;                    // row group i8 is valid iff 4*i8 < rows_left
  float* cptr = p.C + (size_t)m_first * p.ldc + n;
  const size_t cstep = (size_t)4 * p.ldc;
  const float* aux_base = RES ? ep.residual : (DNZ ? ep.dact_src : nullptr);
  const int aux_ld = RES ? ep.ldr : p.ldc;
  const float* aptr = (RES || DNZ) ? aux_base + (size_t)m_first * aux_ld + n : nullptr;
  const size_t astep = (size_t)4 * aux_ld;
  uint32_t sptr = stage + (uint32_t)(r0 * 36 + cq) * 4;
  // dropout: word(g) = mix64(key + (g + 1) * gamma), g = (m * N + n) / 4; a row step of 4 adds N to g
  uint64_t z = 0, zstep = 0;
  uint32_t thr = 0;
  float dscale = 1.0f;
  if (DROP) {
    const uint64_t key = drop_key(seed_value(ep.seed), ep.site);
    const uint64_t g = ((uint64_t)m_first * (uint64_t)p.N + (uint64_t)n) >> 2;
    z = key + (g + 1) * kGamma;
    zstep = (uint64_t)p.N * kGamma;
    thr = drop_threshold(ep.p_drop);
    dscale = 1.0f / (1.0f - ep.p_drop);
  }
  const float dact_scale = ep.dact_scale;
  const bool ln_after = ep.ln_out != nullptr;
  // all eight row groups of the residual / activation source are requested up front: one memory latency per
  // sub-tile instead of one per row group (the dH = dZ W2 product reads 24.6 MB of h here)
  float4 aux[8];
  if (RES || DNZ) {
#pragma unroll
    for (int i8 = 0; i8 < 8; ++i8) {
      aux[i8] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * i8 < rows_left) aux[i8] = *reinterpret_cast<const float4*>(aptr + (size_t)i8 * astep);
    }
  }
#pragma unroll
  for (int i8 = 0; i8 < 8; ++i8) {
    if (rows_left > 0) {
      float4 aux4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (RES || DNZ) aux4 = aux[i8];
      const float4 a4 = lds128(sptr);
      float v[4] = {a4.x + b4.x, a4.y + b4.y, a4.z + b4.z, a4.w + b4.w};
      if (RELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.0f);
      }
      if (DROP) {
        const uint64_t w = mix64(z);
        const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
        v[0] = (lo & 0xFFFFu) >= thr ? v[0] * dscale : 0.0f;
        v[1] = (lo >> 16) >= thr ? v[1] * dscale : 0.0f;
        v[2] = (hi & 0xFFFFu) >= thr ? v[2] * dscale : 0.0f;
        v[3] = (hi >> 16) >= thr ? v[3] * dscale : 0.0f;
      }
      const float ax[4] = {aux4.x, aux4.y, aux4.z, aux4.w};
      if (DNZ) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = ax[j] != 0.0f ? v[j] * dact_scale : 0.0f;
      }
      if (RES) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += ax[j];
      }
      *reinterpret_cast<float4*>(cptr) = make_float4(v[0], v[1], v[2], v[3]);
      if (RES && BIAS && ln_after)   // the fused LayerNorm reads the finished rows from the staging tiles (epilogue_layernorm)
        sts128(sptr, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
    cptr += cstep;
    sptr += 4 * 36 * 4;
    rows_left -= 4;
    if (DROP) z += zstep;
  }
}

// ---- LayerNorm fused behind the epilogue (one tile spans the row: N <= 128) ----------------------------------------------
// The four warps of a lane quadrant (32 rows x 4 column groups) have written z = residual + drop(acc + bias) to global
// memory and back into their staging tiles; after a named barrier over those 128 threads each warp normalises eight of
// the 32 rows: lane = column within a 32-wide group, the same two-pass arithmetic as layernorm_fwd_kernel (rowwise.cu).
constexpr float TC_LN_EPS = 1e-5f;
__device__ __forceinline__ void epilogue_layernorm(const TcParams& p, uint32_t smem_base, int m_quad, int quad, int cg, int lane) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");
  const Epilogue& ep = p.ep;
  float gm[4], bt[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = lane + 32 * k;
    gm[k] = c < p.N ? __ldg(ep.ln_gamma + c) : 0.f;
    bt[k] = c < p.N ? __ldg(ep.ln_beta + c) : 0.f;
  }
  const float inv_n = 1.0f / (float)p.N;
#pragma unroll 2
  for (int rr = 0; rr < 8; ++rr) {
    const int row = cg * 8 + rr, m = m_quad + row;
    float v[4], sum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = 0.f;
      if (lane + 32 * k < p.N) {
        const uint32_t a = smem_base + (uint32_t)(((quad + 4 * k) * (32 * 36) + row * 36 + lane) * 4);
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[k]) : "r"(a) : "memory");
      }
      sum += v[k];
    }
    const float mean = warp_sum(sum) * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (lane + 32 * k < p.N) { v[k] -= mean; sq += v[k] * v[k]; }
    const float rstd = rsqrtf(warp_sum(sq) * inv_n + TC_LN_EPS);
    if (m < p.M) {
      float* y = ep.ln_out + (size_t)m * p.ldc;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (lane + 32 * k < p.N) y[lane + 32 * k] = v[k] * rstd * gm[k] + bt[k];
    }
  }
}

// ---- epilogue of one 32 x 32 sub-tile (already transposed into `stage`, 32 rows x 36 floats) ------------------------
// Lane -> rows i8*4 + lane/8 (i8 = 0..7), columns (lane%8)*4 .. +3: every global access is a 128-byte row segment.
template <int EPI>
__device__ __forceinline__ void epilogue_subtile(const TcParams& p, uint32_t stage, int m_base, int n_base, int lane,
                                                 int z) {
  const int cq = (lane & 7) * 4, r0 = lane >> 3;
  const int n = n_base + cq;
  // All loops below stay rolled on purpose: this code runs once per CTA, so its cost is instruction fetch, not
  // issue (r1 timeline: 3700 cycles for the fully unrolled version of 8 x {LDS, 4 FADD, STG}; the instruction
  // cache is cold at every launch).
  if (p.out_mode >= 2) {   // gradient accumulation: C += tile
    if (n < p.N) {
#pragma unroll 1
      for (int i8 = 0; i8 < 8; ++i8) {
        const int rr = i8 * 4 + r0, m = m_base + rr;
        if (m < p.M) {
          const float4 v = lds128(stage + (uint32_t)(rr * 36 + cq) * 4);
          float* c = p.C + (size_t)m * p.ldc + n;
          if (p.out_mode == 2) {
            atomicAdd(reinterpret_cast<float4*>(c), v);
          } else {
            atomicAdd(c, v.x);
            if (n + 1 < p.N) atomicAdd(c + 1, v.y);
            if (n + 2 < p.N) atomicAdd(c + 2, v.z);
            if (n + 3 < p.N) atomicAdd(c + 3, v.w);
          }
        }
      }
    }
    return;
  }
  if (p.out_mode == 1) {
    if (n < p.Np) {
#pragma unroll 1
      for (int i8 = 0; i8 < 8; ++i8) {
        const int rr = i8 * 4 + r0, m = m_base + rr;
        if (m < p.M)
          *reinterpret_cast<float4*>(p.partial + ((size_t)z * p.M + m) * p.Np + n) = lds128(stage + (uint32_t)(rr * 36 + cq) * 4);
      }
    }
    return;
  }
  if (EPI == EPI_FULL) {
#pragma unroll 1
    for (int i8 = 0; i8 < 8; ++i8) {
      const int rr = i8 * 4 + r0;
      const float4 v4 = lds128(stage + (uint32_t)(rr * 36 + cq) * 4);
      float v[4] = {v4.x, v4.y, v4.z, v4.w};
      epilogue_store4(p.ep, p.C, p.ldc, p.M, p.N, m_base + rr, n, v);
    }
    return;
  }
  // Vector fast paths (host guarantees N % 4 == 0 and 16-byte alignment of every row segment touched).
  if (n >= p.N) return;
  const Epilogue& ep = p.ep;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ep.bias) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
  const float* aux_base = ep.residual ? ep.residual : (EPI == EPI_DROP && ep.dact == DACT_NONZERO ? ep.dact_src : nullptr);
  const int aux_ld = ep.residual ? ep.ldr : p.ldc;
  const bool has_beta = ep.beta != 0.0f;
  const bool relu = EPI == EPI_DROP && ep.act == GANFFN_ACT_RELU;
  const bool drop = EPI == EPI_DROP && ep.p_drop > 0.0f && ep.dact == DACT_NONE;
  const bool dnz = EPI == EPI_DROP && ep.dact == DACT_NONZERO;
  if (!has_beta) {   // hot combinations: compile-time flags (epi_vec); anything else takes the generic loop below
    const bool hb = ep.bias != nullptr, hr = ep.residual != nullptr;
    if (EPI == EPI_DROP) {
      if (hb && relu && drop && !hr) return epi_vec<true, true, true, false, false>(p, stage, m_base, n, lane);
      if (!hb && !relu && dnz && !hr) return epi_vec<false, false, false, true, false>(p, stage, m_base, n, lane);
      if (hb && !relu && drop && hr) return epi_vec<true, false, true, false, true>(p, stage, m_base, n, lane);
    } else {
      if (hb && !hr) return epi_vec<true, false, false, false, false>(p, stage, m_base, n, lane);
      if (!hb && hr) return epi_vec<false, false, false, false, true>(p, stage, m_base, n, lane);
      if (hb && hr) return epi_vec<true, false, false, false, true>(p, stage, m_base, n, lane);
      if (!hb && !hr) return epi_vec<false, false, false, false, false>(p, stage, m_base, n, lane);
    }
  }
  // N % 4 == 0 on this path, so every element index below is 4-aligned: key / threshold hoisted out of the loop
  const DropCtx dctx = drop ? make_drop_ctx(seed_value(ep.seed), ep.site, ep.p_drop) : DropCtx{0ull, 0u, 1.0f};
  // Software pipeline (the residual / dact source / old C of row group i8+1 is in flight while i8 is computed) with
  // strength-reduced pointers: the loop is issue-bound, 16 warps run it at once.
  const int m_first = m_base + r0;
  float* cptr = p.C + (size_t)m_first * p.ldc + n;
  const float* aptr = aux_base ? aux_base + (size_t)m_first * aux_ld + n : nullptr;
  const size_t cstep = (size_t)4 * p.ldc, astep = (size_t)4 * aux_ld;
  uint32_t sptr = stage + (uint32_t)(r0 * 36 + cq) * 4;
  int rows_left = p.M - m_first;                    // row i8 is valid iff 4*i8 < rows_left
  uint64_t eidx = (uint64_t)m_first * (uint64_t)p.N + (uint64_t)n;   // dropout element index
  float4 aux_n = make_float4(0.f, 0.f, 0.f, 0.f), old_n = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rows_left > 0) {
    if (aptr) aux_n = *reinterpret_cast<const float4*>(aptr);
    if (has_beta) old_n = *reinterpret_cast<const float4*>(cptr);
  }
#pragma unroll 1
  for (int i8 = 0; i8 < 8; ++i8) {
    if (rows_left <= 0) break;
    const float4 aux4 = aux_n, old4 = old_n;
    if (rows_left > 4) {
      if (aptr) aux_n = *reinterpret_cast<const float4*>(aptr + astep);
      if (has_beta) old_n = *reinterpret_cast<const float4*>(cptr + cstep);
    }
    const float4 a4 = lds128(sptr);
    float v[4] = {a4.x + b4.x, a4.y + b4.y, a4.z + b4.z, a4.w + b4.w};
    const float ax[4] = {aux4.x, aux4.y, aux4.z, aux4.w};
    if (relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = v[j] > 0.0f ? v[j] : 0.0f;
    }
    if (drop) {
      float msk[4];
      dropout_scale4_aligned(dctx, eidx, msk);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= msk[j];
    }
    if (dnz) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = ax[j] != 0.0f ? v[j] * ep.dact_scale : 0.0f;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += ax[j];   // residual (zeros when absent)
    }
    if (has_beta) {
      v[0] += ep.beta * old4.x; v[1] += ep.beta * old4.y; v[2] += ep.beta * old4.z; v[3] += ep.beta * old4.w;
    }
    *reinterpret_cast<float4*>(cptr) = make_float4(v[0], v[1], v[2], v[3]);
    cptr += cstep;
    if (aptr) aptr += astep;
    sptr += 4 * 36 * 4;
    rows_left -= 4;
    eidx += (uint64_t)4 * (uint64_t)p.N;
  }
}

// ---- the kernel ----------------------------------------------------------------------------------------------
template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* full = bars;                 // [STAGES] producers (one arrival per warp) -> MMA
  uint64_t* empty = bars + STAGES;       // [STAGES] MMA (tcgen05.commit) -> producers
  uint64_t* accum = bars + 2 * STAGES;   // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  if (t == 0) TR(0);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);
  const int nkb = (kend - kbeg + BK - 1) / BK;

  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full + s), NPW);
      mbar_init(smem_u32(empty + s), 1);
    }
    mbar_init(smem_u32(accum), 1);
    fence_barrier_init();
  }
  if (warp == NPW) tmem_alloc(smem_u32(tmem_slot), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (t == 0) TR(1);

  if (warp < NAW) {
    // ================= A producers: global -> registers -> TMEM =================
    const int quad = warp & 3, half = warp >> 2;
    const int row_base = m0 + quad * 32;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(TM_A + half * 16);
    // per-thread base pointer (K block 0), elements to advance per K block, remaining rows / first k of the thread
    const float* ap = !TA ? p.A + (size_t)(row_base + (lane >> 2)) * p.lda + kbeg + half * 16 + 2 * (lane & 3)
                          : p.A + (size_t)(kbeg + half * 16) * p.lda + row_base + lane;
    const int64_t a_kstep = !TA ? BK : (int64_t)BK * p.lda;
    const int a_rows_left = !TA ? p.M - (row_base + (lane >> 2)) : p.M - (row_base + lane);
    const int a_k0 = kbeg + half * 16 + (!TA ? 2 * (lane & 3) : 0);
    const bool want_rowsum = TA && p.ep.rowsum != nullptr && blockIdx.x == 0;   // bias gradient of a wgrad product
    float buf[RING][16];
    float rowsum = 0.f;   // MN-major A: thread = row, so the row sum over k needs no exchange
    // kb0 starts at -RING: the first pass only issues the loads of K blocks 0..RING-1 (same code as steady state,
    // so nothing here is executed-once straight-line code)
    for (int kb0 = -RING; kb0 < nkb; kb0 += RING) {
#pragma unroll
      for (int r = 0; r < RING; ++r) {
        const int kb = kb0 + r;
        if (kb >= 0 && kb < nkb) {
          const int s = kb % STAGES;
          const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
          mbar_wait(smem_u32(empty + s), ph ^ 1u);
          tc_fence_after();
          if (want_rowsum) {
#pragma unroll
            for (int j = 0; j < 16; ++j) rowsum += buf[r][j];
          }
          store_a<!TA>(buf[r], trow + (uint32_t)(s * 64), p.x1);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(full + s));
          if (t == 0 && kb < 16) TR(8 + kb);
        }
        const int kn = kb + RING;
        if (kn < nkb) load_a<!TA>(buf[r], ap + kn * a_kstep, p.lda, a_rows_left, kend - (a_k0 + kn * BK));
      }
    }
    if (want_rowsum && row_base + lane < p.M) atomicAdd(p.ep.rowsum + row_base + lane, rowsum);
  } else if (warp < NPW) {
    // ================= B producers: global -> registers -> shared memory =================
    const int tb = t - NAW * 32;
    const float* bp;
    int64_t b_kstep, b_istride;
    int b_rows_left, b_k0;
    uint32_t b_off;
    if (TB) {   // K-major
      const int r = tb >> 3, c = tb & 7;
      bp = p.B + (size_t)(n0 + r) * p.ldb + kbeg + 4 * c;
      b_kstep = BK; b_istride = (int64_t)32 * p.ldb;
      b_rows_left = p.N - (n0 + r); b_k0 = kbeg + 4 * c;
      b_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
    } else {    // MN-major: 512-byte atoms ordered [k-group of 4][n-group of 32]: LBO = 512 B, SBO = (BN/32) * 512 B
      const int k = tb >> 5, mq = tb & 31;
      bp = p.B + (size_t)(kbeg + k) * p.ldb + n0 + 4 * mq;
      b_kstep = (int64_t)BK * p.ldb; b_istride = (int64_t)8 * p.ldb;
      b_rows_left = p.N - (n0 + 4 * mq); b_k0 = kbeg + k;
      b_off = (uint32_t)(((k >> 2) * (BN / 32) + (mq >> 3)) * 512 + (k & 3) * 128 + ((((mq & 7) >> 1) ^ (k & 3)) << 5) +
                         ((mq & 1) << 4));
    }
    const uint32_t b_smem = smem_u32(smem) + b_off;
    float4 buf[RING][BCH];
    for (int kb0 = -RING; kb0 < nkb; kb0 += RING) {
#pragma unroll
      for (int r = 0; r < RING; ++r) {
        const int kb = kb0 + r;
        if (kb >= 0 && kb < nkb) {
          const int s = kb % STAGES;
          const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
          mbar_wait(smem_u32(empty + s), ph ^ 1u);
          const uint32_t st = b_smem + (uint32_t)(s * STAGE);
          store_b(buf[r], st, st + B_TILE, p.x1);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(full + s));
          if (tb == 0 && kb < 16) TR(24 + kb);
        }
        const int kn = kb + RING;
        if (kn < nkb) load_b<TB>(buf[r], bp + kn * b_kstep, b_istride, b_rows_left, kend - (b_k0 + kn * BK));
      }
    }
  } else {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      const int n_mma = min(BN, (int)((p.N - n0 + 15) & ~15));   // UMMA N: multiple of 16
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((TB ? 0u : 1u) << 16) | ((uint32_t)(n_mma >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      // Descriptors are built once; per MMA only the 14-bit start-address field (low word) moves.  The issue
      // loop must stay a handful of instructions per MMA: this warp shares its scheduler with four busy
      // producer warps (r1 timeline: 1600 cycles per K block with the descriptors rebuilt inside the loop).
      const uint32_t b_lbo = TB ? 16 : 512, b_sbo = TB ? 1024 : (BN / 32) * 512, b_lt = TB ? 2u : 1u;
      const uint64_t bdesc0 = make_desc(smem_u32(smem), b_lbo, b_sbo, b_lt);
      const uint32_t bdesc_hi32 = (uint32_t)(bdesc0 >> 32), bdesc_lo32 = (uint32_t)bdesc0;
      // B: K-major -> advance 32 B inside the swizzle row; MN-major -> advance two 4-k atom groups
      constexpr uint32_t B_JSTEP = (TB ? 32u : 2u * (BN / 32) * 512u) >> 4;
      const uint32_t d_main = tmem_base + TM_MAIN, d_corr = tmem_base + TM_CORR;
      // The tensor core accumulates in fp32 with truncation, so every accumulation step costs up to one ulp of
      // the running sum.  The two correction products (2^-11 of the result) go to their own accumulator: the
      // main one then sees K/8 accumulations instead of 3K/8, and the corrections' own truncation error is
      // scaled down by 2^-11.  The epilogue adds the two in round-to-nearest.
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(smem_u32(full + s), ph);
        tc_fence_after();
        if (kb < 16) TR(48 + 2 * kb);
        const uint32_t b_lo32 = bdesc_lo32 + (uint32_t)((s * STAGE) >> 4);
        const uint32_t a_hi = tmem_base + (uint32_t)(TM_A + s * 64);
        const int ksteps = min(BK / 8, (kend - (kbeg + kb * BK) + 7) >> 3);   // skip all-zero K steps of a ragged tail
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          if (j < ksteps) {
            const uint64_t dbh = ((uint64_t)bdesc_hi32 << 32) | (uint64_t)(b_lo32 + j * B_JSTEP);
            const uint64_t dbl = ((uint64_t)bdesc_hi32 << 32) | (uint64_t)(b_lo32 + j * B_JSTEP + (B_TILE >> 4));
            const uint32_t acc = (j > 0 || kb > 0) ? 1u : 0u;
            if (!p.x1) {
              umma_tf32_ts(d_corr, a_hi + 32 + 8 * j, dbh, idesc, acc);
              umma_tf32_ts(d_corr, a_hi + 8 * j, dbl, idesc, 1u);
            }
            umma_tf32_ts(d_main, a_hi + 8 * j, dbh, idesc, acc);
          }
        }
        umma_commit(smem_u32(empty + s));   // frees the stage (both rings) when these MMAs retire
        if (kb < 16) TR(49 + 2 * kb);
      }
      umma_commit(smem_u32(accum));         // accumulators complete
    }
    __syncwarp();
  }

  if (warp < NPW) {
    // ================= epilogue: warp -> TMEM lane quadrant warp%4, columns 32*(warp/4) .. +31 =================
    if (t == 0) TR(3);
    mbar_wait(smem_u32(accum), 0);
    tc_fence_after();
    if (t == 0) TR(4);
    const int quad = warp & 3, col0 = (warp >> 2) * 32;
    const bool live = nkb > 0 && n0 + col0 < p.N;
    if (live) {
      const uint32_t stage = smem_u32(smem) + (uint32_t)(warp * (32 * 36) * 4);   // private 32 x 36 fp32 transpose buffer
      uint32_t r[32], rl[32];
      const uint32_t tr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0;
      tmem_ld32(tr + TM_MAIN, r);
      if (!p.x1) {
        tmem_ld32(tr + TM_CORR, rl);
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) rl[q] = 0u;
      }
      tmem_ld_wait();
      if (t == 0) TR(70);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float ox = __uint_as_float(r[4 * q]) + __uint_as_float(rl[4 * q]);
        const float oy = __uint_as_float(r[4 * q + 1]) + __uint_as_float(rl[4 * q + 1]);
        const float oz = __uint_as_float(r[4 * q + 2]) + __uint_as_float(rl[4 * q + 2]);
        const float ow = __uint_as_float(r[4 * q + 3]) + __uint_as_float(rl[4 * q + 3]);
        sts128(stage + (uint32_t)(lane * 36 + q * 4) * 4, __float_as_uint(ox), __float_as_uint(oy), __float_as_uint(oz),
               __float_as_uint(ow));
      }
      __syncwarp();
      if (t == 0) TR(71);
      epilogue_subtile<EPI>(p, stage, m0 + quad * 32, n0 + col0, lane, blockIdx.z);
      if (t == 0) TR(72);
    }
    // host side: only set for out_mode 0, one tile across N, and an epilogue served by epi_vec<BIAS, .., RES>
    if (EPI != EPI_FULL && p.ep.ln_out != nullptr && p.out_mode == 0)
      epilogue_layernorm(p, smem_u32(smem), m0 + quad * 32, quad, warp >> 2, lane);
    tc_fence_before();
    if (t == 0) TR(5);
  }
  __syncthreads();
  if (t == 0) TR(6);
  if (warp == NPW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

// ---- A-stationary variant: K <= 128, wide N ---------------------------------------------------------------------
// The d=100 networks' two widest products -- linear1 forward ([T,100] x [100,2048]) and its mirror in the backward
// pass (dH = dZ W2, same shape) -- have four K blocks per 128 x 128 tile: the generic kernel above spends its life in
// prologue, first-load latency and epilogue (ncu r1: 36-41 us for 1.2 GFLOP, 384 CTAs in 2.6 waves, nothing
// overlapped across CTAs because each owns the SM's whole TMEM and 197 KB of smem).  Here one CTA owns a 128-row
// block of A for its whole life: A (hi | lo, all of K) is written to TMEM once, then the CTA walks over 64-wide N
// tiles with B tiles double-buffered in shared memory and the accumulators double-buffered in TMEM, so the MMAs of
// tile i+1 run under the epilogue of tile i:
//   warps 0-7   load A once (same tcgen05.st path as above), then are the epilogue warps (TMEM -> registers ->
//               smem transpose -> fused epilogue -> row-contiguous global stores)
//   warps 8-15  B producers: all K blocks of one 64-wide tile per stage (64 KB), two stages
//   warp 16     MMA issuer: 3 x ceil(K/8) MMAs per tile into accumulator buffer (tile & 1)
// TMEM: [0,128) accumulators of buffer 0 (main | corr), [128,256) buffer 1, [256,512) A (4 K blocks x (32 hi | 32 lo)).
constexpr int ABN = 64;
constexpr int AKB = 4;
constexpr int AB_TILE = ABN * BK * 4;          // one hi (or lo) tile of one K block: 8 KB
constexpr int AB_KB = 2 * AB_TILE;
constexpr int AB_STAGE = AKB * AB_KB;          // 64 KB
constexpr int ASTAGES = 2;
constexpr int A_EPI = NAW * 32 * 36 * 4;
constexpr int ASMEM = ASTAGES * AB_STAGE + A_EPI + 1024 + 256;
constexpr int ATM_A = 256;
constexpr int ANBW = 4;                        // B producer warps (13 warps per CTA: up to 152 registers per thread)
constexpr int ANPW = NAW + ANBW;
constexpr int A_NTHREADS = (ANPW + 1) * 32;
static_assert(ASMEM <= 227 * 1024, "smem budget");

template <bool TB, int EPI>
__global__ void __launch_bounds__(A_NTHREADS, 1) gemm_tc_astat_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* epi_smem = smem + ASTAGES * AB_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + A_EPI);
  uint64_t* a_full = bars;                    // A producers -> MMA (once)
  uint64_t* b_full = bars + 1;                // [2] B producers -> MMA
  uint64_t* b_empty = bars + 3;               // [2] MMA -> B producers
  uint64_t* acc_full = bars + 5;              // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 7;             // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  if (t == 0) TR(0);
  const int m0 = blockIdx.y * BM;
  const int nkb = (p.K + BK - 1) / BK;
  const int ntiles = (p.N + ABN - 1) / ABN;

  if (t == 0) {
    mbar_init(smem_u32(a_full), NAW);
    for (int s = 0; s < ASTAGES; ++s) {
      mbar_init(smem_u32(b_full + s), ANBW);
      mbar_init(smem_u32(b_empty + s), 1);
      mbar_init(smem_u32(acc_full + s), 1);
      mbar_init(smem_u32(acc_empty + s), NAW);
    }
    fence_barrier_init();
  }
  if (warp == ANPW) tmem_alloc(smem_u32(tmem_slot), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (t == 0) TR(1);

  if (warp < NAW) {
    const int quad = warp & 3, half = warp >> 2;
    const int row_base = m0 + quad * 32;
    {
      // ================= A: global -> registers -> TMEM, all K blocks, once =================
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ATM_A + half * 16);
      const float* ap = p.A + (size_t)(row_base + (lane >> 2)) * p.lda + half * 16 + 2 * (lane & 3);
      const int a_rows_left = p.M - (row_base + (lane >> 2));
      const int a_k0 = half * 16 + 2 * (lane & 3);
      float buf[AKB][16];
#pragma unroll
      for (int kb = 0; kb < AKB; ++kb)
        if (kb < nkb) load_a<true>(buf[kb], ap + kb * BK, p.lda, a_rows_left, p.K - (a_k0 + kb * BK));
#pragma unroll
      for (int kb = 0; kb < AKB; ++kb)
        if (kb < nkb) store_a<true>(buf[kb], trow + (uint32_t)(kb * 64), p.x1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(a_full));
      if (t == 0) TR(2);
    }
    // ================= epilogue: warp -> lane quadrant warp%4, columns 32*(warp/4) .. +31 of the 64-wide tile ==========
    const int col0 = half * 32;
    const uint32_t stage = smem_u32(epi_smem) + (uint32_t)(warp * (32 * 36) * 4);
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      const int n0 = tile * ABN;
      mbar_wait(smem_u32(acc_full + s), ph);
      tc_fence_after();
      if (t == 0 && it < 8) TR(8 + it);
      const bool live = n0 + col0 < p.N;
      uint32_t r[32], rl[32];
      if (live) {
        const uint32_t tr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 2 * ABN + col0);
        tmem_ld32(tr, r);
        if (!p.x1) {
          tmem_ld32(tr + ABN, rl);
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) rl[q] = 0u;
        }
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(acc_empty + s));   // the accumulator buffer is free again
      if (live) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float ox = __uint_as_float(r[4 * q]) + __uint_as_float(rl[4 * q]);
          const float oy = __uint_as_float(r[4 * q + 1]) + __uint_as_float(rl[4 * q + 1]);
          const float oz = __uint_as_float(r[4 * q + 2]) + __uint_as_float(rl[4 * q + 2]);
          const float ow = __uint_as_float(r[4 * q + 3]) + __uint_as_float(rl[4 * q + 3]);
          sts128(stage + (uint32_t)(lane * 36 + q * 4) * 4, __float_as_uint(ox), __float_as_uint(oy), __float_as_uint(oz),
                 __float_as_uint(ow));
        }
        __syncwarp();
        epilogue_subtile<EPI>(p, stage, row_base, n0 + col0, lane, 0);
        __syncwarp();   // the transpose buffer is rewritten by the next tile
      }
      if (t == 0 && it < 8) TR(80 + it);
    }
    tc_fence_before();
  } else if (warp < ANPW) {
    // ================= B producers: one 64-wide tile (all K blocks) per stage =================
    const int tb = t - NAW * 32;
    int64_t g_off, g_istride, g_kstep, g_nstep;
    int rows_off, k_off, i_rows, i_k;   // validity: row/k offsets of chunk 0 and their step per chunk index
    uint32_t b_off;
    if (TB) {   // K-major: element (n, k) at B[n * ldb + k]
      const int r = tb >> 3, c = tb & 7;   // rows r + 16 i
      g_off = (int64_t)r * p.ldb + 4 * c; g_istride = (int64_t)16 * p.ldb; g_kstep = BK; g_nstep = (int64_t)ABN * p.ldb;
      rows_off = r; k_off = 4 * c; i_rows = 16; i_k = 0;
      b_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
    } else {    // MN-major: element (n, k) at B[k * ldb + n]
      const int k = tb >> 4, mq = tb & 15;   // k + 8 i
      g_off = (int64_t)k * p.ldb + 4 * mq; g_istride = (int64_t)8 * p.ldb; g_kstep = (int64_t)BK * p.ldb; g_nstep = ABN;
      rows_off = 4 * mq; k_off = k; i_rows = 0; i_k = 8;
      b_off = (uint32_t)(((k >> 2) * (ABN / 32) + (mq >> 3)) * 512 + (k & 3) * 128 + ((((mq & 7) >> 1) ^ (k & 3)) << 5) +
                         ((mq & 1) << 4));
    }
    const uint32_t b_smem = smem_u32(smem) + b_off;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      const int n0 = tile * ABN;
      const float* bp = p.B + g_off + (int64_t)tile * g_nstep;
      float4 v[AKB][4];
#pragma unroll
      for (int kb = 0; kb < AKB; ++kb)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[kb][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          const bool ok = kb < nkb && (n0 + rows_off + i * i_rows < p.N) && (kb * BK + k_off + i * i_k < p.K);
          if (ok) v[kb][i] = __ldg(reinterpret_cast<const float4*>(bp + kb * g_kstep + i * g_istride));
        }
      mbar_wait(smem_u32(b_empty + s), ph ^ 1u);
#pragma unroll
      for (int kb = 0; kb < AKB; ++kb) {
        if (kb < nkb) {
          const uint32_t hi = b_smem + (uint32_t)(s * AB_STAGE + kb * AB_KB);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x[4] = {v[kb][i].x, v[kb][i].y, v[kb][i].z, v[kb][i].w};
            uint32_t h[4], l[4];
            split2(x[0], x[1], h[0], h[1], l[0], l[1]);
            split2(x[2], x[3], h[2], h[3], l[2], l[3]);
            sts128(hi + i * 2048, h[0], h[1], h[2], h[3]);
            if (!p.x1) sts128(hi + AB_TILE + i * 2048, l[0], l[1], l[2], l[3]);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(b_full + s));
      if (tb == 0 && it < 8) TR(24 + it);
    }
  } else {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t b_lbo = TB ? 16 : 512, b_sbo = TB ? 1024 : (ABN / 32) * 512, b_lt = TB ? 2u : 1u;
      const uint64_t bdesc0 = make_desc(smem_u32(smem), b_lbo, b_sbo, b_lt);
      const uint32_t bdesc_hi32 = (uint32_t)(bdesc0 >> 32), bdesc_lo32 = (uint32_t)bdesc0;
      constexpr uint32_t B_JSTEP = (TB ? 32u : 2u * (ABN / 32) * 512u) >> 4;
      mbar_wait(smem_u32(a_full), 0);
      tc_fence_after();
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (uint32_t)(it >> 1) & 1u;
        const int n0 = tile * ABN;
        const int n_mma = min(ABN, (int)((p.N - n0 + 15) & ~15));
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((TB ? 0u : 1u) << 16) | ((uint32_t)(n_mma >> 3) << 17) |
                               ((uint32_t)(BM >> 4) << 24);
        mbar_wait(smem_u32(acc_empty + s), ph ^ 1u);
        mbar_wait(smem_u32(b_full + s), ph);
        tc_fence_after();
        if (it < 8) TR(48 + 2 * it);
        const uint32_t d_main = tmem_base + (uint32_t)(s * 2 * ABN), d_corr = d_main + ABN;
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t b_lo32 = bdesc_lo32 + (uint32_t)((s * AB_STAGE + kb * AB_KB) >> 4);
          const uint32_t a_hi = tmem_base + (uint32_t)(ATM_A + kb * 64);
          const int ksteps = min(BK / 8, (p.K - kb * BK + 7) >> 3);
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            if (j < ksteps) {
              const uint64_t dbh = ((uint64_t)bdesc_hi32 << 32) | (uint64_t)(b_lo32 + j * B_JSTEP);
              const uint64_t dbl = ((uint64_t)bdesc_hi32 << 32) | (uint64_t)(b_lo32 + j * B_JSTEP + (AB_TILE >> 4));
              const uint32_t acc = (j > 0 || kb > 0) ? 1u : 0u;
              if (!p.x1) {
                umma_tf32_ts(d_corr, a_hi + 32 + 8 * j, dbh, idesc, acc);
                umma_tf32_ts(d_corr, a_hi + 8 * j, dbl, idesc, 1u);
              }
              umma_tf32_ts(d_main, a_hi + 8 * j, dbh, idesc, acc);
            }
          }
        }
        umma_commit(smem_u32(b_empty + s));
        umma_commit(smem_u32(acc_full + s));
        if (it < 8) TR(49 + 2 * it);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (t == 0) TR(6);
  if (warp == ANPW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

__global__ void __launch_bounds__(256) tc_splitk_reduce_kernel(const float* __restrict__ partial, int splits, float* C,
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
    const float4 v = *reinterpret_cast<const float4*>(partial + ((size_t)z * M + m) * Np + n);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float v[4] = {s.x, s.y, s.z, s.w};
  epilogue_store4(ep, C, ldc, M, N, m, n, v);
}

// Split-K fold + bias + dropout + residual + LayerNorm for row-spanning outputs (N <= 128, N % 4 == 0): warp = row,
// lane = one float4 of the row.  Replaces the fold and the stand-alone LayerNorm launch of the d <= 128 FFN's linear2.
__global__ void __launch_bounds__(256) tc_splitk_reduce_ln_kernel(const float* __restrict__ partial, int splits, float* C,
                                                                  int ldc, int M, int N, int Np, const Epilogue ep) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int n = 4 * lane;
  const bool on = n < N;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (on) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
      const float4 q = *reinterpret_cast<const float4*>(partial + ((size_t)z * M + m) * Np + n);
      s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
    }
    const float4 b4 = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[0] = s.x + b4.x; v[1] = s.y + b4.y; v[2] = s.z + b4.z; v[3] = s.w + b4.w;
    if (ep.p_drop > 0.0f) {
      float msk[4];
      dropout_scale4(seed_value(ep.seed), ep.site, (uint64_t)m * (uint64_t)N + (uint64_t)n, ep.p_drop, 1.0f / (1.0f - ep.p_drop), msk);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= msk[j];
    }
    if (ep.residual) {
      const float4 r4 = *reinterpret_cast<const float4*>(ep.residual + (size_t)m * ep.ldr + n);
      v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
    }
    *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
  }
  const float inv_n = 1.0f / (float)N;
  const float mean = warp_sum(v[0] + v[1] + v[2] + v[3]) * inv_n;
  float sq = 0.f;
  if (on) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] -= mean; sq += v[j] * v[j]; }
  }
  const float rstd = rsqrtf(warp_sum(sq) * inv_n + TC_LN_EPS);
  if (on) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(ep.ln_gamma + n)), e4 = __ldg(reinterpret_cast<const float4*>(ep.ln_beta + n));
    *reinterpret_cast<float4*>(ep.ln_out + (size_t)m * ldc + n) =
        make_float4(v[0] * rstd * g4.x + e4.x, v[1] * rstd * g4.y + e4.y, v[2] * rstd * g4.z + e4.z, v[3] * rstd * g4.w + e4.w);
  }
}

// Split-K fold + residual + LayerNorm BACKWARD (N <= 128, N % 4 == 0): the folded product is dy of x = LN(z); warp = row,
// lane = one float4 of the row, two rows per warp; the column sums (dgamma, dbeta, bias gradient of the sublayer) are
// folded across the block's warps in shared memory and added to the arena with one red.global.add per column and block.
// Replaces the fold and the stand-alone LayerNorm-backward launch behind the d <= 128 FFN's dX = dH W1 product.
constexpr int LNB_ROWS = 16;   // rows per 256-thread block
__global__ void __launch_bounds__(256) tc_splitk_reduce_lnb_kernel(const float* __restrict__ partial, int splits, float* C,
                                                                   int ldc, int M, int N, int Np, const Epilogue ep) {
  __shared__ float4 red[8][3][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = 4 * lane;
  const bool on = n < N;
  const float inv_n = 1.0f / (float)N;
  const float4 gm = on ? __ldg(reinterpret_cast<const float4*>(ep.lnb_gamma + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool drop = ep.lnb_dzd != nullptr && ep.lnb_p > 0.0f;
  const uint64_t seed = drop ? seed_value(ep.lnb_seed) : 0ull;
  const float dscale = drop ? 1.0f / (1.0f - ep.lnb_p) : 1.0f;
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg, ds = dg;
#pragma unroll
  for (int rr = 0; rr < LNB_ROWS / 8; ++rr) {
    const int m = blockIdx.x * LNB_ROWS + rr * 8 + warp;
    if (m >= M) continue;                       // warp-uniform
    float4 dy = make_float4(0.f, 0.f, 0.f, 0.f), z = dy;
    if (on) {
      for (int zi = 0; zi < splits; ++zi) {
        const float4 q = *reinterpret_cast<const float4*>(partial + ((size_t)zi * M + m) * Np + n);
        dy.x += q.x; dy.y += q.y; dy.z += q.z; dy.w += q.w;
      }
      if (ep.residual) {
        const float4 r4 = *reinterpret_cast<const float4*>(ep.residual + (size_t)m * ep.ldr + n);
        dy.x += r4.x; dy.y += r4.y; dy.z += r4.z; dy.w += r4.w;
      }
      z = *reinterpret_cast<const float4*>(ep.lnb_z + (size_t)m * ldc + n);
      *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) = dy;
    }
    const float mean = warp_sum(z.x + z.y + z.z + z.w) * inv_n;
    float sq = 0.f;
    if (on) { z.x -= mean; z.y -= mean; z.z -= mean; z.w -= mean; sq = z.x * z.x + z.y * z.y + z.z * z.z + z.w * z.w; }
    const float rstd = rsqrtf(warp_sum(sq) * inv_n + TC_LN_EPS);
    float s1 = 0.f, s2 = 0.f;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) {
      z.x *= rstd; z.y *= rstd; z.z *= rstd; z.w *= rstd;       // xhat
      dg.x += dy.x * z.x; dg.y += dy.y * z.y; dg.z += dy.z * z.z; dg.w += dy.w * z.w;
      db.x += dy.x; db.y += dy.y; db.z += dy.z; db.w += dy.w;
      g = make_float4(dy.x * gm.x, dy.y * gm.y, dy.z * gm.z, dy.w * gm.w);
      s1 = g.x + g.y + g.z + g.w;
      s2 = g.x * z.x + g.y * z.y + g.z * z.z + g.w * z.w;
    }
    s1 = warp_sum(s1) * inv_n;
    s2 = warp_sum(s2) * inv_n;
    if (on) {
      float4 o = make_float4(rstd * (g.x - s1 - z.x * s2), rstd * (g.y - s1 - z.y * s2), rstd * (g.z - s1 - z.z * s2),
                             rstd * (g.w - s1 - z.w * s2));
      *reinterpret_cast<float4*>(ep.lnb_dz + (size_t)m * ldc + n) = o;
      if (drop) {
        float msk[4];
        dropout_scale4(seed, ep.lnb_site, (uint64_t)m * (uint64_t)N + (uint64_t)n, ep.lnb_p, dscale, msk);
        o = make_float4(o.x * msk[0], o.y * msk[1], o.z * msk[2], o.w * msk[3]);
        *reinterpret_cast<float4*>(ep.lnb_dzd + (size_t)m * ldc + n) = o;
      }
      ds.x += o.x; ds.y += o.y; ds.z += o.z; ds.w += o.w;
    }
  }
  red[warp][0][lane] = dg;
  red[warp][1][lane] = db;
  red[warp][2][lane] = ds;
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * N; c += blockDim.x) {
    const int seg = c / N, col = c - seg * N;
    float* out = seg == 0 ? ep.lnb_dgamma : seg == 1 ? ep.lnb_dbeta : ep.lnb_dbias;
    if (out == nullptr) continue;
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sacc += reinterpret_cast<const float*>(&red[w][seg][0])[col];
    atomicAdd(out + col, sacc);
  }
}

struct TcPlan { int splits; int kps; };

// Split-K for an accumulating (red.global.add) product: no fold kernel, so the only costs are the per-CTA
// prologue / atomic epilogue and the waves.  These are the weight-gradient products: they run on the side stream
// of net_bwd, off the critical path, and the train step is bound by aggregate SM time -- so the
// plan is the one with the least SM time (CTAs x cycles per CTA) among those within 2x of the lowest latency, not
// the lowest-latency one (which splits K as finely as the SM count allows and pays the fixed CTA cost 144 times).
TcPlan tc_plan_atomic(int M, int N, int K) {
  const int tiles = cdiv(M, BM) * cdiv(N, BN);
  const int nkb = cdiv(K, BK);
  struct Cand { TcPlan pl; double latency, sm_time; };
  Cand cands[64];
  int n = 0;
  double best_latency = 1e300;
  // candidates start at the smallest split that respects the 128-accumulation chain cap (kps <= 1024): a reduction
  // over a million rows (the graph layers' weight gradients on the 1M-utterance sweep, K = 1 000 030) needs ~1000
  // slices -- r2: with the old "sp <= 64" scan no candidate qualified and the product ran on 7 CTAs for 13 ms
  const int sp_min = std::max(1, cdiv(K, 1024));
  for (int sp = sp_min; sp <= nkb && sp < sp_min + 64 && n < 64; ++sp) {
    const int kps = (int)round_up(cdiv(K, sp), BK);
    if (kps > 1024) continue;
    const int splits = cdiv(K, kps);
    const int ctas = tiles * splits;
    const int waves = cdiv((int64_t)ctas, 148);
    const double per_cta = cdiv(kps, BK) * 800.0 + 6000.0;
    const double latency = waves * per_cta + splits * 150.0;
    cands[n++] = Cand{TcPlan{splits, kps}, latency, ctas * per_cta};
    best_latency = std::min(best_latency, latency);
  }
  TcPlan best{1, (int)round_up(K, BK)};
  double best_sm = 1e300;
  static const double cap = getenv("GANFFN_WGRAD_CAP") ? atof(getenv("GANFFN_WGRAD_CAP")) : 2.0;   // tuning switch
  for (int i = 0; i < n; ++i)
    if (cands[i].latency <= cap * best_latency && cands[i].sm_time < best_sm) { best_sm = cands[i].sm_time; best = cands[i].pl; }
  return best;
}

// Split-K factor from a simple wave model of the kernel time (cycles).  kps <= 1024 caps the truncating fp32
// accumulation chain of the tensor core at 128 steps per partial.
TcPlan tc_plan(int M, int N, int K) {
  const int tiles = cdiv(M, BM) * cdiv(N, BN);
  const int nkb = cdiv(K, BK);
  TcPlan best{1, (int)round_up(K, BK)};
  double best_cost = 1e300;
  const int cand_splits[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, std::max(33, cdiv(K, 1024))};   // last: chain cap for huge K
  for (int sp : cand_splits) {
    if (sp > 1 && nkb / sp < 4) continue;
    const int kps = (int)round_up(cdiv(K, sp), BK);
    if (kps > 1024 && nkb / (sp + 1) >= 4) continue;
    const int splits = cdiv(K, kps);
    const int waves = cdiv((int64_t)tiles * splits, 148);
    const double per_cta = cdiv(kps, BK) * 800.0 + 3500.0;
    double cost = waves * per_cta;
    if (splits > 1) cost += 4000.0 + (double)M * N * (splits + 1) * 4.0 / 4000.0;   // fold kernel (~4 KB/cycle)
    // With concurrent sub-step chains the step is bound by SM-slot time as much as by latency (one CTA of this
    // kernel owns its SM): weigh the slots a plan occupies (tuning knob, r2 A/B runs).
    static const double slot_w = getenv("GANFFN_SLOT_WEIGHT") ? atof(getenv("GANFFN_SLOT_WEIGHT")) : 0.0;
    cost += slot_w * per_cta * (double)tiles * splits / 148.0;
    if (cost < best_cost) { best_cost = cost; best = TcPlan{splits, kps}; }
  }
  return best;
}

template <bool TA, bool TB>
int launch_tc2(const TcParams& p, int epi, dim3 grid, cudaStream_t st) {
  GANFFN_SMEM_OPTIN((gemm_tc_kernel<TA, TB, EPI_PLAIN>), SMEM);
  GANFFN_SMEM_OPTIN((gemm_tc_kernel<TA, TB, EPI_DROP>), SMEM);
  GANFFN_SMEM_OPTIN((gemm_tc_kernel<TA, TB, EPI_FULL>), SMEM);
  if (epi == EPI_PLAIN) gemm_tc_kernel<TA, TB, EPI_PLAIN><<<grid, NTHREADS, SMEM, st>>>(p);
  else if (epi == EPI_DROP) gemm_tc_kernel<TA, TB, EPI_DROP><<<grid, NTHREADS, SMEM, st>>>(p);
  else gemm_tc_kernel<TA, TB, EPI_FULL><<<grid, NTHREADS, SMEM, st>>>(p);
  GANFFN_LAUNCHED("gemm_tc_kernel");
  return GANFFN_OK;
}

int launch_tc(const TcParams& p, bool TA, bool TB, int epi, dim3 grid, cudaStream_t st) {
  if (!TA && TB) return launch_tc2<false, true>(p, epi, grid, st);
  if (!TA && !TB) return launch_tc2<false, false>(p, epi, grid, st);
  if (TA && !TB) return launch_tc2<true, false>(p, epi, grid, st);
  return launch_tc2<true, true>(p, epi, grid, st);
}

template <bool TB>
int launch_astat2(const TcParams& p, int epi, dim3 grid, cudaStream_t st) {
  GANFFN_SMEM_OPTIN((gemm_tc_astat_kernel<TB, EPI_PLAIN>), ASMEM);
  GANFFN_SMEM_OPTIN((gemm_tc_astat_kernel<TB, EPI_DROP>), ASMEM);
  GANFFN_SMEM_OPTIN((gemm_tc_astat_kernel<TB, EPI_FULL>), ASMEM);
  if (epi == EPI_PLAIN) gemm_tc_astat_kernel<TB, EPI_PLAIN><<<grid, A_NTHREADS, ASMEM, st>>>(p);
  else if (epi == EPI_DROP) gemm_tc_astat_kernel<TB, EPI_DROP><<<grid, A_NTHREADS, ASMEM, st>>>(p);
  else gemm_tc_astat_kernel<TB, EPI_FULL><<<grid, A_NTHREADS, ASMEM, st>>>(p);
  GANFFN_LAUNCHED("gemm_tc_astat_kernel");
  return GANFFN_OK;
}

// K <= 128 and at least four 64-wide N tiles: one pass over A per CTA pays.
inline bool astat_applies(bool transA, int M, int N, int K, const Epilogue& ep) {
  static const bool off = getenv("GANFFN_NO_ASTAT") != nullptr;   // A/B switch for profiling
  if (off) return false;
  return !transA && K <= AKB * BK && N >= 4 * ABN && M >= 96 && !ep.atomic_acc && ep.rowsum == nullptr;
}

inline bool al16(const void* q) { return (((uintptr_t)q) & 15) == 0; }

// Which epilogue specialisation can serve this call.
int pick_epilogue(const Epilogue& ep, const float* C, int ldc, int N) {
  const bool vec_ok = (N & 3) == 0 && (ldc & 3) == 0 && al16(C) && (!ep.bias || al16(ep.bias)) &&
                      (!ep.residual || ((ep.ldr & 3) == 0 && al16(ep.residual))) && (!ep.dact_src || al16(ep.dact_src));
  if (!vec_ok || ep.pre || ep.drop_before_act) return EPI_FULL;
  if (ep.act == GANFFN_ACT_NONE && ep.p_drop == 0.0f && ep.dact == DACT_NONE) return EPI_PLAIN;
  if ((ep.act == GANFFN_ACT_NONE || ep.act == GANFFN_ACT_RELU) && (ep.dact == DACT_NONE || ep.dact == DACT_NONZERO) &&
      !(ep.dact == DACT_NONZERO && (ep.residual || ep.act != GANFFN_ACT_NONE)))
    return EPI_DROP;
  return EPI_FULL;
}

}  // namespace

bool gemm_tc_supported(bool transA, bool b_is_nk, int lda, int ldb, int ldc, int M, int N, int K, const void* A,
                       const void* B) {
  if (M < 96 || N < 64 || K < 32) return false;                       // too small for 128-row tensor tiles
  if ((lda & 3) || (ldb & 3) || (((uintptr_t)A | (uintptr_t)B) & 15)) return false;
  if (!transA && (K & 3)) return false;                               // 128-bit chunks along the contiguous dim
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
  GANFFN_CHECK_ARG(ep.rowsum == nullptr || transA, "gemm_tc: rowsum needs an MN-major (transposed) A operand");
  const int Np = (int)round_up(N, 4);
  if (ep.atomic_acc) {
    const TcPlan pa = tc_plan_atomic(M, N, K);
    TcParams p;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.k_per_split = pa.kps; p.partial = nullptr; p.Np = Np; p.ep = ep;
    p.x1 = g_gemm_engine == GANFFN_GEMM_TF32X1;
    p.ep.ln_out = nullptr;   // no LayerNorm behind an accumulating product
    p.out_mode = ((N & 3) == 0 && (ldc & 3) == 0 && al16(C)) ? 2 : 3;
    return launch_tc(p, transA, b_is_nk, EPI_PLAIN, dim3(cdiv(N, BN), cdiv(M, BM), pa.splits), st);
  }
  if (astat_applies(transA, M, N, K, ep)) {
    TcParams p;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.k_per_split = (int)round_up(K, BK); p.partial = nullptr; p.Np = Np; p.ep = ep;
    p.x1 = g_gemm_engine == GANFFN_GEMM_TF32X1;
    p.out_mode = 0;
    const int mtiles = cdiv(M, BM), ntiles = cdiv(N, ABN);
    static const int astat_div = getenv("GANFFN_ASTAT_DIV") ? atoi(getenv("GANFFN_ASTAT_DIV")) : 1;   // tuning knob
    const int groups = std::max(1, std::min(ntiles, 148 / std::max(1, mtiles)) / std::max(1, astat_div));
    const int epi = pick_epilogue(ep, C, ldc, N);
    p.ep.ln_out = nullptr;   // N >= 256 here: the row spans several tiles
    GANFFN_TRY(b_is_nk ? launch_astat2<true>(p, epi, dim3(groups, mtiles, 1), st)
                       : launch_astat2<false>(p, epi, dim3(groups, mtiles, 1), st));
    if (ep.ln_out) GANFFN_TRY(layernorm_fwd(C, ep.ln_gamma, ep.ln_beta, ep.ln_out, M, N, st));
    if (ep.lnb_dz) GANFFN_TRY(layernorm_bwd_after(ep, C, M, N, st));
    return GANFFN_OK;
  }
  TcPlan pl = tc_plan(M, N, K);
  if (pl.splits > 1 && (scratch == nullptr || scratch_floats < (int64_t)pl.splits * M * Np)) {
    pl.splits = 1;
    pl.kps = (int)round_up(K, BK);
  }
  TcParams p;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.M = M; p.N = N; p.K = K; p.k_per_split = pl.kps; p.partial = scratch; p.Np = Np; p.ep = ep;
  p.x1 = g_gemm_engine == GANFFN_GEMM_TF32X1;
  p.out_mode = pl.splits > 1 ? 1 : 0;
  dim3 grid(cdiv(N, BN), cdiv(M, BM), pl.splits);
  const int epi = pl.splits > 1 ? EPI_PLAIN : pick_epilogue(ep, C, ldc, N);
  // LayerNorm behind the product (Epilogue::ln_out): in the kernel's epilogue when one tile spans the row and the
  // epilogue is the bias + [dropout] + residual specialisation, in the split-K fold when K was split, else stand-alone.
  static const bool ln_fuse_off = getenv("GANFFN_NO_LN_FUSE") != nullptr;   // A/B switch
  const bool ln = ep.ln_out != nullptr;
  const bool ln_simple = ln && !ln_fuse_off && N <= BN && (N & 3) == 0 && ldc == N && ep.act == GANFFN_ACT_NONE &&
                         ep.dact == DACT_NONE && !ep.pre && !ep.drop_before_act && ep.beta == 0.0f && ep.bias && ep.residual &&
                         al16(ep.ln_gamma) && al16(ep.ln_beta) && al16(ep.ln_out) && al16(C) && al16(ep.bias) &&
                         (ep.ldr & 3) == 0 && al16(ep.residual);
  const bool ln_in_kernel = ln_simple && pl.splits == 1 && epi != EPI_FULL;
  const bool ln_in_fold = ln_simple && pl.splits > 1;
  if (!ln_in_kernel) p.ep.ln_out = nullptr;
  // LayerNorm backward behind the product (Epilogue::lnb_dz): inside the split-K fold, else stand-alone
  const bool lnb = ep.lnb_dz != nullptr;
  const bool lnb_in_fold = lnb && !ln_fuse_off && !g_deterministic && pl.splits > 1 && N <= BN && (N & 3) == 0 && ldc == N &&
                           ep.act == GANFFN_ACT_NONE && ep.dact == DACT_NONE && !ep.pre && ep.beta == 0.0f && !ep.bias &&
                           ep.p_drop == 0.0f && !ln && al16(C) && al16(ep.lnb_z) && al16(ep.lnb_gamma) && al16(ep.lnb_dz) &&
                           (!ep.lnb_dzd || al16(ep.lnb_dzd)) && (!ep.residual || ((ep.ldr & 3) == 0 && al16(ep.residual)));
  GANFFN_TRY(launch_tc(p, transA, b_is_nk, epi, grid, st));
  if (pl.splits > 1) {
    if (ln_in_fold) {
      tc_splitk_reduce_ln_kernel<<<cdiv(M, 8), 256, 0, st>>>(scratch, pl.splits, C, ldc, M, N, Np, ep);
      GANFFN_LAUNCHED("tc_splitk_reduce_ln_kernel");
    } else if (lnb_in_fold) {
      tc_splitk_reduce_lnb_kernel<<<cdiv(M, LNB_ROWS), 256, 0, st>>>(scratch, pl.splits, C, ldc, M, N, Np, ep);
      GANFFN_LAUNCHED("tc_splitk_reduce_lnb_kernel");
    } else {
      const int64_t nvec = (int64_t)M * (Np / 4);
      tc_splitk_reduce_kernel<<<cdiv(nvec, 256), 256, 0, st>>>(scratch, pl.splits, C, ldc, M, N, Np, ep);
      GANFFN_LAUNCHED("tc_splitk_reduce_kernel");
    }
  }
  if (ln && !ln_in_kernel && !ln_in_fold) GANFFN_TRY(layernorm_fwd(C, ep.ln_gamma, ep.ln_beta, ep.ln_out, M, N, st));
  if (lnb && !lnb_in_fold) GANFFN_TRY(layernorm_bwd_after(ep, C, M, N, st));
  return GANFFN_OK;
}

}  // namespace ganffn

#ifdef GANFFN_TC_TRACE
extern "C" int ganffn_debug_tc_flags(int f) {
  return (int)cudaMemcpyToSymbol(ganffn::g_tc_dbg, &f, sizeof(int));
}
extern "C" int ganffn_debug_tc_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, ganffn::g_tc_trace, sizeof(long long) * 128);
}
#endif
