"""CPU oracle of the dialogue-graph path (north_star parts 2-3).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's CPU leg.

PARITY UNPINNED -- NO REFERENCE IMPLEMENTATION.  /root/reference contains no edge construction and no graph
convolution (`grep -ri "edge_index|edge_type|window|graph|scatter|RGCN"` matches only
README/requirements install notes).  The semantics below are this repository's own statement of a DialogueGCN-style
window graph and of the two convolutions north_star names; they are written as plain Python loops / dense index_add
so that they are obviously what include/ganffn.h says, and the CUDA kernels are held to them (edges bit-exact,
convolutions rtol 1e-4).  They are NOT checked against torch_geometric (not installed) or the upstream DialogueGCN.

  edges      for each dialogue, for each target i, sources j = max(0,i-wp) .. min(L-1,i+wf) ascending (self loop
             included); node ids are dialogue-major (offset = running sum of lengths); canonical order = by target,
             then by source.
  edge_type  ((speaker[j] * n_speakers + speaker[i]) << 1) | (0 if j < i else 1)
  RGCN       out_i = W_root x_i + b + sum_r W_r * mean_{j in N_r(i)} x_j          (mean per relation, PyG's default)
             optional basis decomposition W_r = sum_k a[r,k] V_k
  GraphConv  out_i = W_root x_i + b + W_rel * sum_{j in N(i)} x_j                 (sum aggregation, PyG's default)
"""
from typing import List, Sequence, Tuple

import numpy as np
import torch


def build_edges(lengths: Sequence[int], speakers: Sequence[Sequence[int]], wp: int, wf: int, n_speakers: int
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> edge_index [2,E] int64 (row 0 source, row 1 target), edge_type [E] int32, rowptr [N+1] int64."""
    src, dst, typ, rowptr = [], [], [], [0]
    off = 0
    for b, L in enumerate(lengths):
        spk = speakers[b]
        for i in range(L):
            for j in range(max(0, i - wp), min(L - 1, i + wf) + 1):
                src.append(off + j)
                dst.append(off + i)
                typ.append(((int(spk[j]) * n_speakers + int(spk[i])) << 1) | (0 if j < i else 1))
            rowptr.append(len(src))
        off += L
    return (np.array([src, dst], dtype=np.int64).reshape(2, -1), np.array(typ, dtype=np.int32),
            np.array(rowptr, dtype=np.int64))


def transpose_edges(edge_index: np.ndarray, edge_type: np.ndarray, n_nodes: int):
    """Rows = sources, targets ascending: the structure the backward gather walks."""
    order = np.lexsort((edge_index[1], edge_index[0]))
    rowptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.add.at(rowptr, edge_index[0] + 1, 1)
    return np.cumsum(rowptr), edge_index[1][order].astype(np.int32), edge_type[order]


def pack(x_sbd: torch.Tensor, lengths: Sequence[int]) -> torch.Tensor:
    return torch.cat([x_sbd[:L, b] for b, L in enumerate(lengths)], dim=0)


def unpack(x_nodes: torch.Tensor, lengths: Sequence[int], S: int) -> torch.Tensor:
    out = x_nodes.new_zeros(S, len(lengths), x_nodes.shape[1])
    off = 0
    for b, L in enumerate(lengths):
        out[:L, b] = x_nodes[off:off + L]
        off += L
    return out


def rgcn(x: torch.Tensor, edge_index, edge_type, n_rel: int, weight: torch.Tensor, root: torch.Tensor, bias: torch.Tensor,
         comp: torch.Tensor = None) -> torch.Tensor:
    """x [N,d]; weight [n_rel, d, h] (or bases [n_bases, d, h] with comp [n_rel, n_bases]); root [d, h]; bias [h]."""
    N, d = x.shape
    W = weight if comp is None else torch.einsum("rk,kdh->rdh", comp, weight)
    src = torch.as_tensor(edge_index[0], dtype=torch.long)
    dst = torch.as_tensor(edge_index[1], dtype=torch.long)
    et = torch.as_tensor(edge_type, dtype=torch.long)
    out = x @ root + bias
    for r in range(n_rel):
        m = et == r
        if not bool(m.any()):
            continue
        agg = torch.zeros(N, d, dtype=x.dtype).index_add(0, dst[m], x[src[m]])
        cnt = torch.zeros(N, dtype=x.dtype).index_add(0, dst[m], torch.ones(int(m.sum()), dtype=x.dtype))
        out = out + (agg / cnt.clamp(min=1).unsqueeze(1)) @ W[r]
    return out


def graph_conv(x: torch.Tensor, edge_index, w_rel: torch.Tensor, w_root: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """x [N,d]; w_rel, w_root [d, h]; bias [h]."""
    src = torch.as_tensor(edge_index[0], dtype=torch.long)
    dst = torch.as_tensor(edge_index[1], dtype=torch.long)
    agg = torch.zeros_like(x).index_add(0, dst, x[src])
    return agg @ w_rel + x @ w_root + bias
