"""Dialogue graph on the sm_100a kernels (north_star parts 2-3): window-graph construction into CSR and the
relation-typed / plain graph convolutions as segmented gathers + dense contractions.

ABSENT FROM THE REFERENCE: the reference has no graph code, so this module mirrors no reference
class; its semantics are stated in ``include/ganffn.h`` and pinned by ``oracle/graph_oracle.py`` ("parity unpinned --
no reference implementation").  Layer names follow the convolutions north_star names (DialogueGCN uses torch_geometric's
``RGCNConv`` then ``GraphConv``); neither torch_geometric nor torch_scatter is imported anywhere.

Data flow of one convolution (all device-side, no atomics):
  forward   gather over the CSR (rows = targets)           -> [N, n_rel*d] (RGCN) or [N, d] (GraphConv)
            one dense contraction with the stacked weights -> ganffn_linear_fwd (tcgen05 3xTF32 / FFMA engine)
  backward  ganffn_linear_dgrad / wgrad, then the same gather over the transposed CSR (rows = sources): the
            "scatter-add" of the message gradients is a deterministic segmented sum.
"""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import functional as GF
from ._lib import lib, ptr


class DialogueGraph:
    """The window graph of a batch of dialogues, built on the device.

    ``lengths``  real turns per dialogue; ``speakers`` (S,B) integer speaker ids of the zero-padded batch, or the
    loader's one-hot ``qmask`` (S,B,party); ``wp`` / ``wf`` past / future window.  Device tensors: ``node_off``
    [B+1], ``rowptr`` [N+1], ``col`` / ``etype`` [E] (rows = targets), ``rowptr_t`` / ``col_t`` / ``etype_t`` (rows =
    sources), ``inv_cnt`` [N, n_rel], ``edge_index`` [2,E] and ``edge_type`` [E] in the canonical order."""

    def __init__(self, lengths: Sequence[int], speakers: torch.Tensor, wp: int = 10, wf: int = 10, n_speakers: int = 2,
                 device="cuda", with_edge_index: bool = True):
        L = lib()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DialogueGraph is built by the sm_100a kernels; there is no CPU fallback (oracle/graph_oracle.py is the CPU checker)")
        self.lengths_host = [int(x) for x in lengths]
        self.B, self.S = len(self.lengths_host), max(self.lengths_host)
        self.wp, self.wf, self.n_speakers = int(wp), int(wf), int(n_speakers)
        self.n_rel = 2 * n_speakers * n_speakers
        lh = np.asarray(self.lengths_host, dtype=np.int32)
        nn_ = ctypes.c_int64()
        self.E = int(L.cdll.ganffn_graph_num_edges_host(lh.ctypes.data, self.B, self.wp, self.wf, ctypes.byref(nn_)))
        self.N = int(nn_.value)
        st = torch.cuda.current_stream(dev).cuda_stream
        self.lengths = torch.as_tensor(lh, device=dev)
        if speakers.dim() == 3:
            speakers = speakers.argmax(dim=2)
        spk_sb = speakers.to(device=dev, dtype=torch.int32).contiguous()          # (S,B)
        i32 = lambda n: torch.empty(max(int(n), 1), dtype=torch.int32, device=dev)
        i64 = lambda n: torch.empty(max(int(n), 1), dtype=torch.int64, device=dev)
        self.node_off, self.edge_off = i64(self.B + 1), i64(self.B + 1)
        L.call("ganffn_graph_offsets", ptr(self.lengths), self.B, self.wp, self.wf, ptr(self.node_off), ptr(self.edge_off), st)
        self.rowptr, self.col, self.etype = i64(self.N + 1), i32(self.E), i32(self.E)
        self.node_b, self.node_t = i32(self.N), i32(self.N)
        self.inv_cnt = torch.empty(max(self.N * self.n_rel, 1), dtype=torch.float32, device=dev)
        self.edge_index = i64(2 * self.E).view(2, -1) if with_edge_index else None
        # per-node speaker ids from the padded (S,B) batch (index plumbing; the kernels re-derive node_b / node_t)
        nb = torch.repeat_interleave(torch.arange(self.B, device=dev), self.lengths.long())
        nt = torch.arange(self.N, device=dev) - self.node_off[:-1][nb]
        self.speakers = spk_sb[nt, nb].contiguous() if self.N else torch.zeros(1, dtype=torch.int32, device=dev)
        L.call("ganffn_graph_build", ptr(self.lengths), ptr(self.speakers), ptr(self.node_off), ptr(self.edge_off), self.B,
               self.wp, self.wf, self.n_speakers, 0, ptr(self.rowptr), ptr(self.col), ptr(self.etype),
               ptr(self.edge_index) if with_edge_index else None, self.E, ptr(self.node_b), ptr(self.node_t), ptr(self.inv_cnt), st)
        self.rowptr_t, self.col_t, self.etype_t = i64(self.N + 1), i32(self.E), i32(self.E)
        L.call("ganffn_graph_build", ptr(self.lengths), ptr(self.speakers), ptr(self.node_off), ptr(self.edge_off), self.B,
               self.wp, self.wf, self.n_speakers, 1, ptr(self.rowptr_t), ptr(self.col_t), ptr(self.etype_t), None, self.E,
               None, None, None, st)
        self.edge_type = self.etype
        self.device = dev

    # (S,B,d) zero-padded batch <-> packed node features
    def pack(self, x_sbd: torch.Tensor) -> torch.Tensor:
        return _Pack.apply(x_sbd, self)

    def unpack(self, x_nodes: torch.Tensor) -> torch.Tensor:
        return _Unpack.apply(x_nodes, self)


def _pack(x, g: DialogueGraph):
    x = x.contiguous()
    out = torch.empty((g.N, x.shape[2]), dtype=torch.float32, device=x.device)
    lib().call("ganffn_graph_pack", ptr(x), ptr(g.node_b), ptr(g.node_t), ptr(out), g.N, g.B, x.shape[2], GF._stream(x))
    return out


def _unpack(xn, g: DialogueGraph):
    xn = xn.contiguous()
    out = torch.empty((g.S, g.B, xn.shape[1]), dtype=torch.float32, device=xn.device)
    lib().call("ganffn_graph_unpack", ptr(xn), ptr(g.lengths), ptr(g.node_off), ptr(out), g.S, g.B, xn.shape[1], GF._stream(xn))
    return out


class _Pack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g):
        GF._require_cuda(x, "graph pack input")
        ctx.g = g
        return _pack(x, g)

    @staticmethod
    def backward(ctx, dy):
        return _unpack(dy, ctx.g), None


class _Unpack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xn, g):
        GF._require_cuda(xn, "graph unpack input")
        ctx.g = g
        return _unpack(xn, g)

    @staticmethod
    def backward(ctx, dy):
        return _pack(dy, ctx.g), None


def _linear_fwd(x, w, b, residual=None):
    L = lib()
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    n = L.query("ganffn_gemm_scratch_floats", M, N, K)
    ws = GF.scratch(x.device, n, "graph")
    L.call("ganffn_linear_fwd", ptr(x), ptr(w), ptr(b), ptr(residual), ptr(y), None, M, N, K, 0, 0, 0.0, 0, 0, ptr(ws), n,
           GF._stream(x))
    return y


def _linear_dgrad(dy, w, K, residual=None):
    L = lib()
    M, N = dy.shape
    dx = torch.empty((M, K), dtype=torch.float32, device=dy.device)
    n = L.query("ganffn_gemm_scratch_floats", M, K, N)
    ws = GF.scratch(dy.device, n, "graph")
    L.call("ganffn_linear_dgrad", ptr(dy), ptr(w), ptr(residual), ptr(dx), M, N, K, ptr(ws), n, GF._stream(dy))
    return dx


def _linear_wgrad(dy, x, with_bias):
    L = lib()
    M, K = x.shape
    N = dy.shape[1]
    dw = torch.empty((N, K), dtype=torch.float32, device=x.device)
    db = torch.empty(N, dtype=torch.float32, device=x.device) if with_bias else None
    ws = GF.scratch(x.device, L.query("ganffn_wgrad_scratch_floats", M, N, K), "graph")
    L.call("ganffn_linear_wgrad", ptr(dy), ptr(x), ptr(dw), ptr(db), M, N, K, 0, ptr(ws), GF._stream(x))
    return dw, db


class _GraphLayer(torch.autograd.Function):
    """y = gather(x) @ w_rel^T + x @ w_root^T + b.  ``typed``: relation-typed mean gather ([N, n_rel*d], w_rel
    [h, n_rel*d]); else plain sum ([N, d], w_rel [h, d]).  The gathered features are never concatenated or copied:
    the root product is the residual of the relation product's epilogue."""

    @staticmethod
    def forward(ctx, x, w_rel, w_root, bias, g: DialogueGraph, typed: bool):
        GF._require_cuda(x, "graph convolution input")
        L = lib()
        x, w_rel, w_root, bias = x.contiguous(), w_rel.contiguous(), w_root.contiguous(), bias.contiguous()
        N, d = x.shape
        slots = g.n_rel if typed else 1
        agg = torch.empty((N, slots * d), dtype=torch.float32, device=x.device)
        st = GF._stream(x)
        if typed:
            L.call("ganffn_graph_gather_typed", ptr(x), ptr(g.rowptr), ptr(g.col), ptr(g.etype), ptr(g.inv_cnt), ptr(agg), N,
                   g.n_rel, d, ptr(g.node_off), g.B, g.S, st)
        else:
            L.call("ganffn_graph_gather_sum", ptr(x), ptr(g.rowptr), ptr(g.col), None, None, ptr(agg), N, 1, g.n_rel, d, ptr(g.node_off), g.B, g.S, st)
        y0 = _linear_fwd(x, w_root, bias)
        y = _linear_fwd(agg, w_rel, None, residual=y0)
        ctx.save_for_backward(x, agg, w_rel, w_root)
        ctx.g, ctx.typed = g, typed
        return y

    @staticmethod
    def backward(ctx, dy):
        L = lib()
        x, agg, w_rel, w_root = ctx.saved_tensors
        g, typed = ctx.g, ctx.typed
        N, d = x.shape
        slots = g.n_rel if typed else 1
        dy = dy.contiguous()
        dagg = _linear_dgrad(dy, w_rel, slots * d)
        dmsg = torch.empty((N, d), dtype=torch.float32, device=dy.device)
        st = GF._stream(dy)
        if typed:
            L.call("ganffn_graph_gather_sum", ptr(dagg), ptr(g.rowptr_t), ptr(g.col_t), ptr(g.etype_t), ptr(g.inv_cnt), ptr(dmsg),
                   N, g.n_rel, g.n_rel, d, None, 0, 0, st)
        else:
            L.call("ganffn_graph_gather_sum", ptr(dagg), ptr(g.rowptr_t), ptr(g.col_t), None, None, ptr(dmsg), N, 1, g.n_rel, d, ptr(g.node_off), g.B, g.S, st)
        dx = _linear_dgrad(dy, w_root, d, residual=dmsg)
        dw_rel, _ = _linear_wgrad(dy, agg, False)
        dw_root, db = _linear_wgrad(dy, x, True)
        return dx, dw_rel, dw_root, db, None, None


class RGCNConv(nn.Module):
    """Relation-typed graph convolution: ``out_i = root^T x_i + bias + sum_r W_r^T mean_{j in N_r(i)} x_j`` with optional
    basis decomposition ``W_r = sum_k comp[r,k] bases[k]`` (parameters named as torch_geometric names them)."""

    def __init__(self, in_channels: int, out_channels: int, num_relations: int, num_bases: Optional[int] = None):
        super().__init__()
        self.in_channels, self.out_channels, self.num_relations, self.num_bases = in_channels, out_channels, num_relations, num_bases
        nb = num_bases if num_bases else num_relations
        self.weight = nn.Parameter(torch.empty(nb, in_channels, out_channels))
        self.comp = nn.Parameter(torch.empty(num_relations, nb)) if num_bases else None
        self.root = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        bound = 1.0 / math.sqrt(in_channels)
        nn.init.uniform_(self.weight, -bound, bound)
        nn.init.uniform_(self.root, -bound, bound)
        if self.comp is not None:
            nn.init.uniform_(self.comp, -1.0 / math.sqrt(nb), 1.0 / math.sqrt(nb))

    def forward(self, x: torch.Tensor, graph: DialogueGraph) -> torch.Tensor:
        if graph.n_rel != self.num_relations:
            raise ValueError(f"graph has {graph.n_rel} relations, layer expects {self.num_relations}")
        W = self.weight if self.comp is None else torch.einsum("rk,kdh->rdh", self.comp, self.weight)   # parameters only
        w_rel = W.reshape(-1, self.out_channels).t()                                                         # [h, R d]
        return _GraphLayer.apply(x, w_rel, self.root.t(), self.bias, graph, True)


class GraphConv(nn.Module):
    """Plain graph convolution: ``out_i = lin_root(x_i) + lin_rel(sum_{j in N(i)} x_j)`` (sum aggregation)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.lin_rel = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_root = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x: torch.Tensor, graph: DialogueGraph) -> torch.Tensor:
        return _GraphLayer.apply(x, self.lin_rel.weight, self.lin_root.weight, self.lin_rel.bias, graph, False)
