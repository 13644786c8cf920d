// Dialogue-graph kernels (north_star parts 2-3): window-graph construction straight into CSR and segmented,
// atomic-free gather / scatter-add for the relation-typed and plain graph convolutions.
//
// ABSENT FROM THE REFERENCE: the semantics are this library's own, stated in include/ganffn.h
// and pinned by oracle/graph_oracle.py ("parity unpinned -- no reference implementation").
//
// All of this is HBM-bound integer / gather work (no tensor cores): the design rules are coalescing and grid size.
//   * build: one warp per dialogue.  Lane = target utterance; its degree is a closed form of (t, len, wp, wf), a
//     warp scan gives the row offsets, and each lane writes its (<= wp+wf+1) edges contiguously.  Rows of one
//     dialogue are adjacent, so a warp's stores cover one contiguous span of col / etype / edge_index.
//   * gather: one warp per node.  Lane c owns one float4 of the feature row, so every neighbour row is one coalesced
//     400-byte (d=100) read; neighbour rows are re-read by <= wp+wf+1 nodes of the same dialogue, i.e. from L2.
//     The backward "scatter-add" runs on the transposed CSR as the same gather: no atomics, deterministic.
#include <stdlib.h>
#include "kernels.h"

namespace ganffn {
namespace {

__device__ __forceinline__ int rel_id(int spk_src, int spk_dst, bool past, int n_spk) {
  return ((spk_src * n_spk + spk_dst) << 1) | (past ? 0 : 1);
}

__device__ __forceinline__ int64_t window_edges(int L, int wp, int wf) {
  int64_t e = 0;
  for (int i = 0; i < L; ++i) e += min(L - 1, i + wf) - max(0, i - wp) + 1;
  return e;
}

// One block: exclusive scans of the per-dialogue node and edge counts (B is at most a few 10^4).
__global__ void __launch_bounds__(1024) graph_offsets_kernel(const int* __restrict__ lengths, int B, int wp, int wf,
                                                             int64_t* __restrict__ node_off, int64_t* __restrict__ edge_off) {
  __shared__ int64_t sn[1024], se[1024];
  __shared__ int64_t carry_n, carry_e;
  const int t = threadIdx.x;
  if (t == 0) { carry_n = 0; carry_e = 0; }
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int b = base + t;
    const int L = b < B ? lengths[b] : 0;
    sn[t] = L;
    se[t] = b < B ? window_edges(L, wp, wf) : 0;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {   // Hillis-Steele inclusive scan
      const int64_t a = t >= o ? sn[t - o] : 0, c = t >= o ? se[t - o] : 0;
      __syncthreads();
      sn[t] += a; se[t] += c;
      __syncthreads();
    }
    if (b < B) {
      node_off[b] = carry_n + sn[t] - L;
      edge_off[b] = carry_e + se[t] - window_edges(L, wp, wf);
    }
    __syncthreads();
    if (t == 1023) { carry_n += sn[1023]; carry_e += se[1023]; }
    __syncthreads();
  }
  if (t == 0) { node_off[B] = carry_n; edge_off[B] = carry_e; }
}

// Warp per dialogue.  Rows = targets with window [i-wp, i+wf] (transposed: rows = sources with window [j-wf, j+wp]).
//
// Dialogues of up to BUILD_CAP turns (every IEMOCAP / MELD dialogue: max_len 110) take the flat path: the warp first
// writes the row starts, speakers and per-speaker prefix counts of the dialogue to shared memory, then its lanes walk
// the dialogue's edges in their final order -- lane = edge, so every col / etype / edge_index store instruction
// covers 32 consecutive edges (128 / 256 contiguous bytes), whatever the row degrees are.  The relation counts of a
// row come from the prefix counts in closed form (sources of speaker a before / from position i inside the window),
// written as one contiguous [L, R] block.  (r1/r2: the row-at-a-time version below kept 21 of 32 lanes busy with a
// 10/10 window and spent ~100 instructions per row on shuffles and eight ballots: 26 % of the HBM roofline.)
constexpr int BUILD_CAP = 128;
constexpr int BUILD_WPB = 8;
constexpr int BUILD_MAX_SPK = 4;   // 2 n^2 <= 32 relations

__device__ __forceinline__ void graph_build_rows(const int* __restrict__ spk, int L, int64_t n0, int64_t e0, int back, int fwd,
                                                 int n_spk, int transposed, int b, int lane, int64_t* __restrict__ rowptr,
                                                 int* __restrict__ col, int* __restrict__ etype, int64_t* __restrict__ edge_index,
                                                 int64_t E, int* __restrict__ node_b, int* __restrict__ node_t,
                                                 float* __restrict__ inv_cnt) {
  const int R = 2 * n_spk * n_spk;
  int64_t carry = 0;
  for (int base = 0; base < L; base += 32) {
    const int i = base + lane;
    const bool valid = i < L;
    const int lo = max(0, i - back), hi = min(L - 1, i + fwd);
    const int deg = valid ? hi - lo + 1 : 0;
    int incl = deg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int64_t e_row = e0 + carry + (incl - deg);
    const int si = valid ? spk[n0 + i] : 0;
    if (valid) {
      rowptr[n0 + i] = e_row;
      if (node_b) { node_b[n0 + i] = b; node_t[n0 + i] = i; }
    }
    const int rows = min(32, L - base);
    for (int r = 0; r < rows; ++r) {
      const int r_i = base + r;
      const int r_lo = __shfl_sync(0xffffffffu, lo, r), r_deg = __shfl_sync(0xffffffffu, deg, r);
      const int r_si = __shfl_sync(0xffffffffu, si, r);
      const int64_t r_e = __shfl_sync(0xffffffffu, e_row, r);
      const int64_t row = n0 + r_i;
      float mycnt = 0.f;                       // lane q < R counts relation q of this row
      for (int k0 = 0; k0 < r_deg; k0 += 32) {
        const int k = k0 + lane;
        int rel = -1;
        if (k < r_deg) {
          const int j = r_lo + k;
          const int sj = spk[n0 + j];
          rel = transposed ? rel_id(r_si, sj, r_i < j, n_spk) : rel_id(sj, r_si, j < r_i, n_spk);
          const int64_t e = r_e + k;
          col[e] = (int)(n0 + j);
          etype[e] = rel;
          if (edge_index) { edge_index[e] = transposed ? row : n0 + j; edge_index[E + e] = transposed ? n0 + j : row; }
        }
        if (inv_cnt)
          for (int q = 0; q < R; ++q) {
            const int c = __popc(__ballot_sync(0xffffffffu, rel == q));
            if (lane == (q & 31)) mycnt += (float)c;   // R <= 32 is checked on the host
          }
      }
      if (inv_cnt && lane < R) inv_cnt[row * R + lane] = mycnt > 0.f ? 1.f / mycnt : 0.f;
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) rowptr[n0 + L] = e0 + carry;   // the next dialogue's first row writes the same value
}

__global__ void __launch_bounds__(BUILD_WPB * 32) graph_build_kernel(const int* __restrict__ lengths, const int* __restrict__ spk,
                                                                    const int64_t* __restrict__ node_off,
                                                                    const int64_t* __restrict__ edge_off, int B, int wp, int wf,
                                                                    int n_spk, int transposed, int64_t* __restrict__ rowptr,
                                                                    int* __restrict__ col, int* __restrict__ etype,
                                                                    int64_t* __restrict__ edge_index, int64_t E,
                                                                    int* __restrict__ node_b, int* __restrict__ node_t,
                                                                    float* __restrict__ inv_cnt) {
  __shared__ int s_rstart[BUILD_WPB][BUILD_CAP + 1];
  __shared__ int s_spk[BUILD_WPB][BUILD_CAP];
  __shared__ int s_pref[BUILD_WPB][BUILD_MAX_SPK][BUILD_CAP + 1];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * BUILD_WPB + wib;
  if (warp >= B) return;
  const int b = warp;
  const int L = lengths[b];
  const int64_t n0 = node_off[b], e0 = edge_off[b];
  const int back = transposed ? wf : wp, fwd = transposed ? wp : wf;   // window of a row: [r - back, r + fwd]
  if (L > BUILD_CAP || n_spk > BUILD_MAX_SPK || (transposed && inv_cnt)) {   // long dialogues: row-at-a-time path
    graph_build_rows(spk, L, n0, e0, back, fwd, n_spk, transposed, b, lane, rowptr, col, etype, edge_index, E, node_b, node_t,
                     inv_cnt);
    return;
  }
  const int R = 2 * n_spk * n_spk;
  int* rstart = s_rstart[wib];
  int* sp = s_spk[wib];
  // ---- pass 1: row starts (local edge offsets), speakers, per-speaker prefix counts ----
  int carry = 0;
  int pc[BUILD_MAX_SPK] = {0, 0, 0, 0};
  for (int base = 0; base < L; base += 32) {
    const int i = base + lane;
    const bool valid = i < L;
    const int deg = valid ? min(L - 1, i + fwd) - max(0, i - back) + 1 : 0;
    int incl = deg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int si = valid ? __ldg(spk + n0 + i) : -1;
    if (valid) {
      rstart[i] = carry + incl - deg;
      sp[i] = si;
      rowptr[n0 + i] = e0 + carry + incl - deg;
      if (node_b) { node_b[n0 + i] = b; node_t[n0 + i] = i; }
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
    if (inv_cnt) {
#pragma unroll
      for (int a = 0; a < BUILD_MAX_SPK; ++a) {
        if (a < n_spk) {
          const unsigned m = __ballot_sync(0xffffffffu, si == a);
          if (valid) s_pref[wib][a][i] = pc[a] + __popc(m & ((1u << lane) - 1u));   // sources of speaker a before i
          pc[a] += __popc(m);
        }
      }
    }
  }
  const int El = carry;
  if (lane == 0) {
    rstart[L] = El;
    rowptr[n0 + L] = e0 + El;   // the next dialogue's first row writes the same value
    if (inv_cnt)
      for (int a = 0; a < n_spk; ++a) s_pref[wib][a][L] = pc[a];
  }
  __syncwarp();
  // ---- pass 2: lane = edge, in the final edge order ----
  {
    int r = 0;
    for (int e = lane; e < El; e += 32) {
      // edges only move forward between iterations: a short linear scan (the row of e is the last r with rstart[r] <= e)
      int lo_r = r, hi_r = L - 1;
      if (rstart[min(r + 4, L)] <= e) {        // far ahead (short rows): binary search the rest
        lo_r = min(r + 4, L - 1);
        while (lo_r < hi_r) {
          const int mid = (lo_r + hi_r + 1) >> 1;
          if (rstart[mid] <= e) lo_r = mid; else hi_r = mid - 1;
        }
        r = lo_r;
      } else {
        while (rstart[r + 1] <= e) ++r;
      }
      const int k = e - rstart[r];
      const int j = max(0, r - back) + k;
      const int si = sp[r], sj = sp[j];
      const int rel = transposed ? rel_id(si, sj, r < j, n_spk) : rel_id(sj, si, j < r, n_spk);
      const int64_t ge = e0 + e;
      col[ge] = (int)(n0 + j);
      etype[ge] = rel;
      if (edge_index) {
        edge_index[ge] = transposed ? n0 + r : n0 + j;
        edge_index[E + ge] = transposed ? n0 + j : n0 + r;
      }
    }
  }
  // ---- pass 3: 1 / |N_r(i)| in closed form from the prefix counts, one contiguous [L, R] block ----
  if (inv_cnt) {
    float* out = inv_cnt + n0 * R;
    for (int idx = lane; idx < L * R; idx += 32) {
      const int i = idx / R, q = idx - i * R;
      const int dir = q & 1, pair = q >> 1;
      const int a = pair / n_spk, dd = pair - a * n_spk;   // relation (source speaker a, target speaker dd, dir)
      int cnt = 0;
      if (dd == sp[i]) {
        const int lo = max(0, i - wp), hi = min(L - 1, i + wf);
        cnt = dir == 0 ? s_pref[wib][a][i] - s_pref[wib][a][lo] : s_pref[wib][a][hi + 1] - s_pref[wib][a][i];
      }
      out[idx] = cnt > 0 ? 1.f / (float)cnt : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) graph_pack_kernel(const float* __restrict__ x, const int* __restrict__ node_b,
                                                         const int* __restrict__ node_t, float* __restrict__ out,
                                                         int64_t N, int B, int d4) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * d4) return;
  const int64_t n = idx / d4;
  const int c = (int)(idx % d4);
  const float4* src = reinterpret_cast<const float4*>(x) + ((int64_t)node_t[n] * B + node_b[n]) * d4 + c;
  reinterpret_cast<float4*>(out)[idx] = __ldg(src);
}

__global__ void __launch_bounds__(256) graph_unpack_kernel(const float* __restrict__ xn, const int* __restrict__ lengths,
                                                           const int64_t* __restrict__ node_off, float* __restrict__ out,
                                                           int S, int B, int d4) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)S * B * d4) return;
  const int c = (int)(idx % d4);
  const int64_t tb = idx / d4;
  const int b = (int)(tb % B), t = (int)(tb / B);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < lengths[b]) v = __ldg(reinterpret_cast<const float4*>(xn) + (node_off[b] + t) * d4 + c);
  reinterpret_cast<float4*>(out)[idx] = v;
}

// Device-side collate of the loader's per-utterance metadata (reference dataloader.py:44-58: speaker one-hots, the
// all-ones utterance mask and the labels of each dialogue, zero-padded by pad_sequence): thread = one padded slot.
__global__ void __launch_bounds__(256) collate_meta_kernel(const int* __restrict__ speakers, const int64_t* __restrict__ labels,
                                                           const int* __restrict__ lengths, const int64_t* __restrict__ node_off,
                                                           float* __restrict__ qmask, float* __restrict__ umask,
                                                           int64_t* __restrict__ label, int S, int B, int n_spk) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)S * B) return;
  const int b = (int)(idx % B), t = (int)(idx / B);          // qmask is (S,B,n_spk): consecutive threads write consecutive rows
  const bool real = t < lengths[b];
  const int64_t n = node_off[b] + t;
  const int sp = real ? speakers[n] : -1;
  for (int k = 0; k < n_spk; ++k) qmask[idx * n_spk + k] = sp == k ? 1.f : 0.f;
  umask[(int64_t)b * S + t] = real ? 1.f : 0.f;              // (B,S)
  label[(int64_t)b * S + t] = real ? labels[n] : 0;
}

constexpr int GV = 4;   // float4 per lane at most: d <= 512 (the kernels are instantiated for 1 and GV)

__device__ __forceinline__ void add4(float4& a, const float4 v, float w) {
  a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
}

// Warp per node; relation-outer, edge-inner: out[n, r, :] = inv[n, r] * sum_{e: etype = r} x[col[e], :].
template <int NV>
__global__ void __launch_bounds__(256) graph_gather_typed_kernel(const float* __restrict__ x, const int64_t* __restrict__ rowptr,
                                                                 const int* __restrict__ col, const int* __restrict__ etype,
                                                                 const float* __restrict__ inv_cnt, float* __restrict__ out,
                                                                 int64_t N, int R, int d4) {
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const int64_t beg = rowptr[n], end = rowptr[n + 1];
  const float4* xv = reinterpret_cast<const float4*>(x);
  float4* ov = reinterpret_cast<float4*>(out) + n * R * d4;
  // first (usually only) chunk of <= 32 edges lives in registers: lane e holds (col, etype) of edge beg + e
  const int deg0 = (int)((end - beg) < 32 ? (end - beg) : 32);
  const int c0 = lane < deg0 ? col[beg + lane] : 0;
  const int t0 = lane < deg0 ? etype[beg + lane] : -1;
  for (int r = 0; r < R; ++r) {
    float4 acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned m = __ballot_sync(0xffffffffu, t0 == r);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int64_t j = __shfl_sync(0xffffffffu, c0, src);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane + 32 * k;
        if (c < d4) add4(acc[k], __ldg(xv + j * d4 + c), 1.f);
      }
    }
    for (int64_t e = beg + 32; e < end; ++e) {   // windows wider than 32 edges (not the DialogueGCN defaults)
      if (etype[e] == r) {
        const int64_t j = col[e];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int c = lane + 32 * k;
          if (c < d4) add4(acc[k], __ldg(xv + j * d4 + c), 1.f);
        }
      }
    }
    const float w = inv_cnt[n * R + r];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      if (c < d4) ov[(int64_t)r * d4 + c] = make_float4(acc[k].x * w, acc[k].y * w, acc[k].z * w, acc[k].w * w);
    }
  }
}

// Warp per node: out[n, :] = sum_e w_e * in[col[e], slot_e, :].
template <bool TYPED, bool WEIGHTED, int NV>
__global__ void __launch_bounds__(256) graph_gather_sum_kernel(const float* __restrict__ in, const int64_t* __restrict__ rowptr,
                                                               const int* __restrict__ col, const int* __restrict__ etype,
                                                               const float* __restrict__ inv_cnt, float* __restrict__ out,
                                                               int64_t N, int in_slots, int R, int d4) {
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const int64_t beg = rowptr[n], end = rowptr[n + 1];
  const float4* iv = reinterpret_cast<const float4*>(in);
  float4 acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t cb = beg; cb < end; cb += 32) {
    const int cnt = (int)((end - cb) < 32 ? (end - cb) : 32);
    const int cj = lane < cnt ? col[cb + lane] : 0;
    const int tj = (TYPED || WEIGHTED) && lane < cnt ? etype[cb + lane] : 0;
    const float wj = WEIGHTED && lane < cnt ? __ldg(inv_cnt + (int64_t)cj * R + tj) : 1.f;
    for (int s = 0; s < cnt; ++s) {
      const int64_t j = __shfl_sync(0xffffffffu, cj, s);
      const int slot = TYPED ? __shfl_sync(0xffffffffu, tj, s) : 0;
      const float w = WEIGHTED ? __shfl_sync(0xffffffffu, wj, s) : 1.f;
      const float4* row = iv + (j * in_slots + slot) * d4;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane + 32 * k;
        if (c < d4) add4(acc[k], __ldg(row + c), w);
      }
    }
  }
  float4* ov = reinterpret_cast<float4*>(out) + n * d4;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane + 32 * k;
    if (c < d4) ov[c] = acc[k];
  }
}

// ---- dialogue-staged gathers ---------------------------------------------------------------------------------------
// The warp-per-node kernels above re-read every feature row from L2 once per neighbour (~21 times with a 10/10
// window): on the 1M-utterance sweep the relation-typed gather moved 11.6 GB through L2 for 3.8 GB of algorithmic
// traffic and ran at L2 speed (2.0 ms, 29 % of HBM peak).  A window never leaves its dialogue, so here one CTA owns
// one dialogue: its <= 110 feature rows are staged in shared memory once (coalesced), every neighbour read is an
// smem read, and HBM sees each input row once and each output row once.
template <bool TYPED, int NV>
__global__ void __launch_bounds__(256) graph_gather_dialogue_kernel(const float* __restrict__ x, const int64_t* __restrict__ node_off,
                                                                    const int64_t* __restrict__ rowptr, const int* __restrict__ col,
                                                                    const int* __restrict__ etype, const float* __restrict__ inv_cnt,
                                                                    float* __restrict__ out, int R, int d4) {
  extern __shared__ __align__(16) float4 xs[];   // [L][d4]
  const int b = blockIdx.x;
  const int64_t n0 = node_off[b];
  const int L = (int)(node_off[b + 1] - n0);
  const float4* xv = reinterpret_cast<const float4*>(x) + n0 * d4;
  {   // four independent loads in flight per thread: the staging is a DRAM-latency chain otherwise
    const int total = L * d4, step = blockDim.x;
    int idx = threadIdx.x;
    for (; idx + 3 * step < total; idx += 4 * step) {
      const float4 a0 = __ldg(xv + idx), a1 = __ldg(xv + idx + step), a2 = __ldg(xv + idx + 2 * step), a3 = __ldg(xv + idx + 3 * step);
      xs[idx] = a0; xs[idx + step] = a1; xs[idx + 2 * step] = a2; xs[idx + 3 * step] = a3;
    }
    for (; idx < total; idx += step) xs[idx] = __ldg(xv + idx);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < L; i += nw) {
    const int64_t n = n0 + i;
    const int64_t beg = rowptr[n], end = rowptr[n + 1];
    const int deg0 = (int)((end - beg) < 32 ? (end - beg) : 32);
    const int c0 = lane < deg0 ? col[beg + lane] - (int)n0 : 0;
    const int t0 = (TYPED && lane < deg0) ? etype[beg + lane] : -1;
    if (TYPED) {
      float4* ov = reinterpret_cast<float4*>(out) + n * R * d4;
      const float w_lane = lane < R ? __ldg(inv_cnt + n * R + lane) : 0.f;   // R <= 32: one coalesced load per row
      const bool short_row = end - beg <= 32;
      for (int r = 0; r < R; ++r) {
        unsigned m = __ballot_sync(0xffffffffu, t0 == r);
        if (m == 0 && short_row) {           // empty relation: zeros, nothing to accumulate or scale
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const int c = lane + 32 * k;
            if (c < d4) ov[(int64_t)r * d4 + c] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          continue;
        }
        float4 acc[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const int j = __shfl_sync(0xffffffffu, c0, src);
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const int c = lane + 32 * k;
            if (c < d4) add4(acc[k], xs[j * d4 + c], 1.f);
          }
        }
        for (int64_t e = beg + 32; e < end; ++e) {
          if (etype[e] == r) {
            const int j = col[e] - (int)n0;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              const int c = lane + 32 * k;
              if (c < d4) add4(acc[k], xs[j * d4 + c], 1.f);
            }
          }
        }
        const float w = R <= 32 ? __shfl_sync(0xffffffffu, w_lane, r) : inv_cnt[n * R + r];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int c = lane + 32 * k;
          if (c < d4) ov[(int64_t)r * d4 + c] = make_float4(acc[k].x * w, acc[k].y * w, acc[k].z * w, acc[k].w * w);
        }
      }
    } else {
      float4 acc[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      // Sources of a row are ascending and unique, so last - first == deg - 1 means a contiguous run (every row of a
      // window graph): the neighbour rows are then xs[first], xs[first + 1], ... -- no shuffle, no index arithmetic.
      const int c_first = __shfl_sync(0xffffffffu, c0, 0);
      const int c_last = __shfl_sync(0xffffffffu, c0, deg0 > 0 ? deg0 - 1 : 0);
      if (deg0 > 0 && c_last - c_first == deg0 - 1) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int c = lane + 32 * k;
          if (c < d4) {
            const float4* row = xs + c_first * d4 + c;
#pragma unroll 4
            for (int s = 0; s < deg0; ++s) {
              const float4 v = row[s * d4];
              acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
            }
          }
        }
      } else {
        for (int s = 0; s < deg0; ++s) {
          const int j = __shfl_sync(0xffffffffu, c0, s);
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const int c = lane + 32 * k;
            if (c < d4) add4(acc[k], xs[j * d4 + c], 1.f);
          }
        }
      }
      for (int64_t e = beg + 32; e < end; ++e) {
        const int j = col[e] - (int)n0;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int c = lane + 32 * k;
          if (c < d4) add4(acc[k], xs[j * d4 + c], 1.f);
        }
      }
      float4* ov = reinterpret_cast<float4*>(out) + n * d4;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane + 32 * k;
        if (c < d4) ov[c] = acc[k];
      }
    }
  }
}

// ---- plain (GraphConv) sum over a window graph: running window sums -------------------------------------------------
// The rows of a window graph are contiguous runs [lo_i, hi_i] with lo and hi non-decreasing in i, so
//   out_i = out_{i-1} + sum_{hi_{i-1} < j <= hi_i} x_j - sum_{lo_{i-1} <= j < lo_i} x_j :
// two row updates per node instead of wp + wf + 1 neighbour reads (r1's dialogue-staged gather was instruction-issue
// bound at 32 % of the HBM roofline doing 21 shared-memory adds per output float4).  One CTA = one dialogue staged in
// shared memory; thread = (row segment, float4 column): the first row of a segment is summed directly, the rest slide.
// A row that is not a contiguous run, or whose bounds move backwards, is summed directly from the CSR (generic graphs
// stay correct).  Sums differ from the direct sum by the rounding of <= L/segments running updates (~1e-6 relative).
__global__ void __launch_bounds__(256) graph_gather_window_kernel(const float* __restrict__ x, const int64_t* __restrict__ node_off,
                                                                  const int64_t* __restrict__ rowptr, const int* __restrict__ col,
                                                                  float* __restrict__ out, int d4) {
  extern __shared__ __align__(16) float4 xs[];   // [L][d4], then int2 bounds[L] (lo, hi; hi < lo marks a non-run row)
  const int b = blockIdx.x;
  const int64_t n0 = node_off[b];
  const int L = (int)(node_off[b + 1] - n0);
  const float4* xv = reinterpret_cast<const float4*>(x) + n0 * d4;
  int2* bounds = reinterpret_cast<int2*>(xs + (size_t)L * d4);
  {
    const int total = L * d4, step = blockDim.x;
    int idx = threadIdx.x;
    for (; idx + 3 * step < total; idx += 4 * step) {
      const float4 a0 = __ldg(xv + idx), a1 = __ldg(xv + idx + step), a2 = __ldg(xv + idx + 2 * step), a3 = __ldg(xv + idx + 3 * step);
      xs[idx] = a0; xs[idx + step] = a1; xs[idx + 2 * step] = a2; xs[idx + 3 * step] = a3;
    }
    for (; idx < total; idx += step) xs[idx] = __ldg(xv + idx);
  }
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const int64_t beg = rowptr[n0 + i], end = rowptr[n0 + i + 1];
    int lo = 0, hi = -1;
    if (end > beg) {
      lo = col[beg] - (int)n0;
      hi = col[end - 1] - (int)n0;
      if (hi - lo != (int)(end - beg) - 1 || lo < 0 || hi >= L) { lo = 1; hi = -1; }   // not a run inside the dialogue
    } else {
      lo = 0; hi = -2;                                                                  // empty row: sum = 0
    }
    bounds[i] = make_int2(lo, hi);
  }
  __syncthreads();
  const int groups = max(1, (int)blockDim.x / d4);
  const int g = threadIdx.x / d4, c = threadIdx.x % d4;
  if (g >= groups) return;
  const int seg = (L + groups - 1) / groups;
  const int i_beg = g * seg, i_end = min(L, i_beg + seg);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int plo = 0, phi = -1;
  bool have = false;
  for (int i = i_beg; i < i_end; ++i) {
    const int2 bd = bounds[i];
    const int64_t n = n0 + i;
    if (bd.y == -2) {                        // empty row
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      have = false;
    } else if (bd.y < bd.x) {                // generic row: direct sum over the CSR entries
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t e = rowptr[n]; e < rowptr[n + 1]; ++e) {
        const int64_t j = col[e];
        const float4 v = (j >= n0 && j < n0 + L) ? xs[(j - n0) * d4 + c] : __ldg(reinterpret_cast<const float4*>(x) + j * d4 + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      have = false;
    } else if (have && bd.x >= plo && bd.y >= phi && bd.x <= phi + 1) {   // slide
      for (int j = phi + 1; j <= bd.y; ++j) { const float4 v = xs[j * d4 + c]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
      for (int j = plo; j < bd.x; ++j) { const float4 v = xs[j * d4 + c]; acc.x -= v.x; acc.y -= v.y; acc.z -= v.z; acc.w -= v.w; }
      plo = bd.x; phi = bd.y;
    } else {                                 // first row of the segment (or a jump): direct sum of the run
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = bd.x; j <= bd.y; ++j) { const float4 v = xs[j * d4 + c]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
      plo = bd.x; phi = bd.y; have = true;
    }
    reinterpret_cast<float4*>(out)[n * d4 + c] = acc;
  }
}

constexpr size_t GATHER_SMEM_MAX = 200 * 1024;

int launch_gather_window(const float* x, const int64_t* node_off, int B, int max_len, const int64_t* rowptr, const int* col,
                         float* out, int d, cudaStream_t st) {
  const size_t smem = (size_t)max_len * d * sizeof(float) + (size_t)max_len * sizeof(int2);
  {
    static std::atomic<unsigned long long> done{0ull};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(done.load(std::memory_order_acquire) & bit)) {
      cudaFuncSetAttribute(graph_gather_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(GATHER_SMEM_MAX + 2048));
      cudaFuncSetAttribute(graph_gather_window_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      done.fetch_or(bit, std::memory_order_release);
    }
  }
  graph_gather_window_kernel<<<B, 256, smem, st>>>(x, node_off, rowptr, col, out, d / 4);
  GANFFN_LAUNCHED("graph_gather_window_kernel");
  return GANFFN_OK;
}

template <bool TYPED, int NV>
int launch_gather_dialogue(const float* x, const int64_t* node_off, int B, int max_len, const int64_t* rowptr, const int* col,
                           const int* etype, const float* inv_cnt, float* out, int R, int d, cudaStream_t st) {
  const size_t smem = (size_t)max_len * d * sizeof(float);
  {
    // per-device attributes (several 44 KB dialogues per SM: also ask for the largest shared-memory carve-out)
    static std::atomic<unsigned long long> done{0ull};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(done.load(std::memory_order_acquire) & bit)) {
      cudaFuncSetAttribute(graph_gather_dialogue_kernel<TYPED, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GATHER_SMEM_MAX);
      cudaFuncSetAttribute(graph_gather_dialogue_kernel<TYPED, NV>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      done.fetch_or(bit, std::memory_order_release);
    }
  }
  graph_gather_dialogue_kernel<TYPED, NV><<<B, 256, smem, st>>>(x, node_off, rowptr, col, etype, inv_cnt, out, R, d / 4);
  GANFFN_LAUNCHED("graph_gather_dialogue_kernel");
  return GANFFN_OK;
}

}  // namespace

int graph_offsets(const int* lengths, int B, int wp, int wf, int64_t* node_off, int64_t* edge_off, cudaStream_t st) {
  GANFFN_CHECK_ARG(lengths && node_off && edge_off && B >= 1 && wp >= 0 && wf >= 0, "graph_offsets: bad arguments");
  graph_offsets_kernel<<<1, 1024, 0, st>>>(lengths, B, wp, wf, node_off, edge_off);
  GANFFN_LAUNCHED("graph_offsets_kernel");
  return GANFFN_OK;
}

int graph_build(const int* lengths, const int* speakers, const int64_t* node_off, const int64_t* edge_off, int B, int wp,
                int wf, int n_speakers, int transposed, int64_t* rowptr, int* col, int* etype, int64_t* edge_index,
                int64_t n_edges, int* node_b, int* node_t, float* inv_cnt, cudaStream_t st) {
  GANFFN_CHECK_ARG(lengths && speakers && node_off && edge_off && rowptr && col && etype, "graph_build: null pointer");
  GANFFN_CHECK_ARG(B >= 1 && wp >= 0 && wf >= 0 && n_speakers >= 1 && 2 * n_speakers * n_speakers <= 32,
                   "graph_build: bad arguments (at most 4 speakers: 2 n^2 <= 32 relations)");
  GANFFN_CHECK_ARG((node_b == nullptr) == (node_t == nullptr), "graph_build: node_b and node_t go together");
  const int wpb = BUILD_WPB;
  graph_build_kernel<<<cdiv(B, wpb), wpb * 32, 0, st>>>(lengths, speakers, node_off, edge_off, B, wp, wf, n_speakers,
                                                        transposed, rowptr, col, etype, edge_index, n_edges, node_b, node_t,
                                                        inv_cnt);
  GANFFN_LAUNCHED("graph_build_kernel");
  return GANFFN_OK;
}

int graph_pack(const float* x_sbd, const int* node_b, const int* node_t, float* x_nodes, int64_t N, int B, int d,
               cudaStream_t st) {
  GANFFN_CHECK_ARG(x_sbd && node_b && node_t && x_nodes && d % 4 == 0, "graph_pack: bad arguments (d must be a multiple of 4)");
  if (N <= 0) return GANFFN_OK;
  graph_pack_kernel<<<cdiv(N * (d / 4), 256), 256, 0, st>>>(x_sbd, node_b, node_t, x_nodes, N, B, d / 4);
  GANFFN_LAUNCHED("graph_pack_kernel");
  return GANFFN_OK;
}

int graph_unpack(const float* x_nodes, const int* lengths, const int64_t* node_off, float* x_sbd, int S, int B, int d,
                 cudaStream_t st) {
  GANFFN_CHECK_ARG(x_nodes && lengths && node_off && x_sbd && d % 4 == 0, "graph_unpack: bad arguments (d must be a multiple of 4)");
  graph_unpack_kernel<<<cdiv((int64_t)S * B * (d / 4), 256), 256, 0, st>>>(x_nodes, lengths, node_off, x_sbd, S, B, d / 4);
  GANFFN_LAUNCHED("graph_unpack_kernel");
  return GANFFN_OK;
}

int collate_meta(const int* speakers, const int64_t* labels, const int* lengths, const int64_t* node_off, float* qmask,
                 float* umask, int64_t* label, int S, int B, int n_spk, cudaStream_t st) {
  GANFFN_CHECK_ARG(speakers && labels && lengths && node_off && qmask && umask && label, "collate_meta: null pointer");
  GANFFN_CHECK_ARG(S >= 1 && B >= 1 && n_spk >= 1, "collate_meta: S=%d B=%d n_speakers=%d", S, B, n_spk);
  collate_meta_kernel<<<cdiv((int64_t)S * B, 256), 256, 0, st>>>(speakers, labels, lengths, node_off, qmask, umask, label, S, B, n_spk);
  GANFFN_LAUNCHED("collate_meta_kernel");
  return GANFFN_OK;
}

int graph_gather_typed(const float* x, const int64_t* rowptr, const int* col, const int* etype, const float* inv_cnt,
                       float* out, int64_t N, int R, int d, const int64_t* node_off, int B, int max_len, cudaStream_t st) {
  GANFFN_CHECK_ARG(x && rowptr && col && etype && inv_cnt && out, "graph_gather_typed: null pointer");
  GANFFN_CHECK_ARG(d % 4 == 0 && d <= 128 * GV && R >= 1, "graph_gather_typed: d=%d must be a multiple of 4 and <= %d", d, 128 * GV);
  if (N <= 0) return GANFFN_OK;
  if (node_off && B > 0 && max_len > 0 && (size_t)max_len * d * sizeof(float) <= GATHER_SMEM_MAX)
    return d <= 128 ? launch_gather_dialogue<true, 1>(x, node_off, B, max_len, rowptr, col, etype, inv_cnt, out, R, d, st)
                    : launch_gather_dialogue<true, GV>(x, node_off, B, max_len, rowptr, col, etype, inv_cnt, out, R, d, st);
  if (d <= 128) graph_gather_typed_kernel<1><<<cdiv(N, 8), 256, 0, st>>>(x, rowptr, col, etype, inv_cnt, out, N, R, d / 4);
  else graph_gather_typed_kernel<GV><<<cdiv(N, 8), 256, 0, st>>>(x, rowptr, col, etype, inv_cnt, out, N, R, d / 4);
  GANFFN_LAUNCHED("graph_gather_typed_kernel");
  return GANFFN_OK;
}

int graph_gather_sum(const float* in, const int64_t* rowptr, const int* col, const int* etype, const float* inv_cnt,
                     float* out, int64_t N, int in_slots, int R, int d, const int64_t* node_off, int B, int max_len,
                     cudaStream_t st) {
  GANFFN_CHECK_ARG(in && rowptr && col && out, "graph_gather_sum: null pointer");
  GANFFN_CHECK_ARG(d % 4 == 0 && d <= 128 * GV && in_slots >= 1, "graph_gather_sum: d=%d must be a multiple of 4 and <= %d", d, 128 * GV);
  GANFFN_CHECK_ARG((in_slots == 1 && inv_cnt == nullptr) || etype != nullptr, "graph_gather_sum: typed / weighted gathers need etype");
  if (N <= 0) return GANFFN_OK;
  const dim3 grid(cdiv(N, 8));
  const bool typed = in_slots > 1, weighted = inv_cnt != nullptr;
  static const bool no_window = getenv("GANFFN_NO_WINDOW_SUM") != nullptr;   // A/B switch
  if (!typed && !weighted && !no_window && node_off && B > 0 && max_len > 0 && d / 4 <= 256 &&
      (size_t)max_len * d * sizeof(float) <= GATHER_SMEM_MAX)
    return launch_gather_window(in, node_off, B, max_len, rowptr, col, out, d, st);
  if (!typed && !weighted && node_off && B > 0 && max_len > 0 && (size_t)max_len * d * sizeof(float) <= GATHER_SMEM_MAX)
    return d <= 128 ? launch_gather_dialogue<false, 1>(in, node_off, B, max_len, rowptr, col, etype, inv_cnt, out, R, d, st)
                    : launch_gather_dialogue<false, GV>(in, node_off, B, max_len, rowptr, col, etype, inv_cnt, out, R, d, st);
#define GANFFN_GS(T, W)                                                                                                  \
  do {                                                                                                                   \
    if (d <= 128) graph_gather_sum_kernel<T, W, 1><<<grid, 256, 0, st>>>(in, rowptr, col, etype, inv_cnt, out, N, in_slots, R, d / 4); \
    else graph_gather_sum_kernel<T, W, GV><<<grid, 256, 0, st>>>(in, rowptr, col, etype, inv_cnt, out, N, in_slots, R, d / 4);          \
  } while (0)
  if (typed && weighted) GANFFN_GS(true, true);
  else if (typed) GANFFN_GS(true, false);
  else if (weighted) GANFFN_GS(false, true);
  else GANFFN_GS(false, false);
#undef GANFFN_GS
  GANFFN_LAUNCHED("graph_gather_sum_kernel");
  return GANFFN_OK;
}

}  // namespace ganffn

using namespace ganffn;
static inline cudaStream_t GS(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int64_t ganffn_graph_num_edges_host(const int* lengths_host, int n_dialogues, int wp, int wf, int64_t* n_nodes) {
  int64_t E = 0, N = 0;
  for (int b = 0; b < n_dialogues; ++b) {
    const int L = lengths_host[b];
    N += L;
    for (int i = 0; i < L; ++i) E += std::min(L - 1, i + wf) - std::max(0, i - wp) + 1;
  }
  if (n_nodes) *n_nodes = N;
  return E;
}

int ganffn_graph_offsets(const int* lengths, int n_dialogues, int wp, int wf, int64_t* node_off, int64_t* edge_off,
                         void* stream) {
  return graph_offsets(lengths, n_dialogues, wp, wf, node_off, edge_off, GS(stream));
}

int ganffn_graph_build(const int* lengths, const int* speakers, const int64_t* node_off, const int64_t* edge_off,
                       int n_dialogues, int wp, int wf, int n_speakers, int transposed, int64_t* rowptr, int* col, int* etype,
                       int64_t* edge_index, int64_t n_edges, int* node_b, int* node_t, float* inv_cnt, void* stream) {
  return graph_build(lengths, speakers, node_off, edge_off, n_dialogues, wp, wf, n_speakers, transposed, rowptr, col, etype,
                     edge_index, n_edges, node_b, node_t, inv_cnt, GS(stream));
}

int ganffn_graph_pack(const float* x_sbd, const int* node_b, const int* node_t, float* x_nodes, int64_t n_nodes, int B, int d,
                      void* stream) {
  return graph_pack(x_sbd, node_b, node_t, x_nodes, n_nodes, B, d, GS(stream));
}

int ganffn_graph_unpack(const float* x_nodes, const int* lengths, const int64_t* node_off, float* x_sbd, int S, int B, int d,
                        void* stream) {
  return graph_unpack(x_nodes, lengths, node_off, x_sbd, S, B, d, GS(stream));
}

int ganffn_collate_meta(const int* speakers, const int64_t* labels, const int* lengths, const int64_t* node_off, float* qmask,
                        float* umask, int64_t* label, int S, int B, int n_speakers, void* stream) {
  return collate_meta(speakers, labels, lengths, node_off, qmask, umask, label, S, B, n_speakers, GS(stream));
}

int ganffn_graph_gather_typed(const float* x, const int64_t* rowptr, const int* col, const int* etype, const float* inv_cnt,
                              float* out, int64_t n_nodes, int n_rel, int d, const int64_t* node_off, int n_dialogues,
                              int max_len, void* stream) {
  return graph_gather_typed(x, rowptr, col, etype, inv_cnt, out, n_nodes, n_rel, d, node_off, n_dialogues, max_len, GS(stream));
}

int ganffn_graph_gather_sum(const float* in, const int64_t* rowptr, const int* col, const int* etype, const float* inv_cnt,
                            float* out, int64_t n_nodes, int in_slots, int n_rel, int d, const int64_t* node_off,
                            int n_dialogues, int max_len, void* stream) {
  return graph_gather_sum(in, rowptr, col, etype, inv_cnt, out, n_nodes, in_slots, n_rel, d, node_off, n_dialogues, max_len,
                          GS(stream));
}

}  // extern "C"
