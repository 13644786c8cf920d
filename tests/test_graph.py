"""Dialogue graph (north_star parts 2-3): CUDA edge construction and graph convolutions against oracle/graph_oracle.py.

PARITY UNPINNED -- NO REFERENCE IMPLEMENTATION: the reference has no graph code, so these tests
hold the kernels to this repository's own CPU statement of the semantics: edge_index / edge_type / CSR bit-exact,
convolution outputs and gradients at rtol 1e-4.  The CPU tests check the oracle against an independent brute-force
formulation (adjacency-matrix form), so the oracle is at least self-consistent."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import graph_oracle as GO

CASES = [  # lengths, wp, wf, n_speakers
    ([12, 7, 10], 10, 10, 2),
    ([1, 2, 1, 5], 3, 2, 2),
    ([94] * 4, 10, 10, 2),
    ([33, 5, 17, 110, 64, 1], 0, 4, 3),
    ([40, 9], 20, 20, 2),          # 41-edge rows: more than one 32-edge chunk
    ([6, 6], 0, 0, 2),             # self loops only
]


def _speakers(lengths, n_spk, seed=3):
    g = torch.Generator().manual_seed(seed)
    S, B = max(lengths), len(lengths)
    spk = torch.randint(0, n_spk, (S, B), generator=g)
    return spk, [spk[:L, b].tolist() for b, L in enumerate(lengths)]


@pytest.mark.parametrize("lengths,wp,wf,n_spk", CASES)
def test_oracle_edges_match_adjacency_formulation(lengths, wp, wf, n_spk):
    spk, per = _speakers(lengths, n_spk)
    ei, et, rowptr = GO.build_edges(lengths, per, wp, wf, n_spk)
    N = sum(lengths)
    assert rowptr[-1] == ei.shape[1] and len(rowptr) == N + 1
    # independent formulation: dense adjacency per dialogue
    A = np.zeros((N, N), dtype=bool)
    off = 0
    for L in lengths:
        t = np.arange(L)
        A[off:off + L, off:off + L] = (t[None, :] >= t[:, None] - wp) & (t[None, :] <= t[:, None] + wf)   # A[target, source]
        off += L
    dst, src = np.nonzero(A)                       # row-major: by target, then by source = the canonical order
    assert np.array_equal(ei[0], src) and np.array_equal(ei[1], dst)
    flat = np.concatenate([np.array(p) for p in per])
    assert np.array_equal(et, ((flat[src] * n_spk + flat[dst]) << 1) | (src >= dst))
    rp_t, col_t, et_t = GO.transpose_edges(ei, et, N)
    dst_t, src_t = np.nonzero(A.T)                 # rows = sources
    assert np.array_equal(col_t, src_t) and rp_t[-1] == ei.shape[1]


def test_oracle_convolutions_match_dense_formulation():
    lengths, wp, wf, n_spk = [5, 3], 2, 1, 2
    spk, per = _speakers(lengths, n_spk)
    ei, et, _ = GO.build_edges(lengths, per, wp, wf, n_spk)
    N, d, h, R = sum(lengths), 8, 6, 2 * n_spk * n_spk
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, d, generator=g, dtype=torch.float64)
    W, root, bias = (torch.randn(*s, generator=g, dtype=torch.float64) for s in ((R, d, h), (d, h), (h,)))
    out = GO.rgcn(x, ei, et, R, W, root, bias)
    ref = x @ root + bias
    for r in range(R):
        A = torch.zeros(N, N, dtype=torch.float64)
        m = et == r
        A[ei[1][m], ei[0][m]] = 1
        ref = ref + (A / A.sum(1, keepdim=True).clamp(min=1)) @ x @ W[r]
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
    A = torch.zeros(N, N, dtype=torch.float64)
    A[ei[1], ei[0]] = 1
    assert torch.allclose(GO.graph_conv(x, ei, W[0], root, bias), A @ x @ W[0] + x @ root + bias, rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("lengths,wp,wf,n_spk", CASES)
def test_cuda_edge_build_is_bit_exact(lengths, wp, wf, n_spk):
    from gan_ffn_b200.graph import DialogueGraph
    spk, per = _speakers(lengths, n_spk)
    ei, et, rowptr = GO.build_edges(lengths, per, wp, wf, n_spk)
    g = DialogueGraph(lengths, spk, wp, wf, n_spk)
    N, E = sum(lengths), ei.shape[1]
    assert (g.N, g.E) == (N, E)
    assert np.array_equal(g.edge_index.cpu().numpy(), ei), "edge_index"
    assert np.array_equal(g.edge_type[:E].cpu().numpy(), et), "edge_type"
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr), "rowptr"
    assert np.array_equal(g.col[:E].cpu().numpy(), ei[0].astype(np.int32)), "col"
    rp_t, col_t, et_t = GO.transpose_edges(ei, et, N)
    assert np.array_equal(g.rowptr_t.cpu().numpy(), rp_t) and np.array_equal(g.col_t[:E].cpu().numpy(), col_t)
    assert np.array_equal(g.etype_t[:E].cpu().numpy(), et_t), "transposed edge_type"
    off = np.concatenate([[0], np.cumsum(lengths)])
    assert np.array_equal(g.node_off.cpu().numpy(), off)
    assert np.array_equal(g.node_b[:N].cpu().numpy(), np.repeat(np.arange(len(lengths)), lengths))
    R = 2 * n_spk * n_spk
    cnt = np.zeros((N, R))
    np.add.at(cnt, (ei[1], et), 1)
    inv = np.where(cnt > 0, 1.0 / np.maximum(cnt, 1), 0.0).astype(np.float32)
    assert np.array_equal(g.inv_cnt[:N * R].view(N, R).cpu().numpy(), inv), "inv_cnt"


@pytest.mark.gpu
def test_cuda_edge_build_many_dialogues():
    """More dialogues than one scan block (1024) and the 1M-sweep length distribution."""
    from gan_ffn_b200 import synthetic
    from gan_ffn_b200.graph import DialogueGraph
    lengths = synthetic.ragged_lengths(2500, 1, 110, seed=21)
    spk, per = _speakers(lengths, 2)
    ei, et, rowptr = GO.build_edges(lengths, per, 10, 10, 2)
    g = DialogueGraph(lengths, spk, 10, 10, 2)
    assert np.array_equal(g.edge_index.cpu().numpy(), ei) and np.array_equal(g.edge_type[:g.E].cpu().numpy(), et)
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr)


@pytest.mark.gpu
@pytest.mark.parametrize("lengths,wp,wf,n_spk", CASES[:5])
@pytest.mark.parametrize("bases", [None, 3])
def test_cuda_graph_convolutions_match_oracle(lengths, wp, wf, n_spk, bases):
    from gan_ffn_b200.graph import DialogueGraph, GraphConv, RGCNConv
    spk, per = _speakers(lengths, n_spk)
    ei, et, _ = GO.build_edges(lengths, per, wp, wf, n_spk)
    S, B, d, h = max(lengths), len(lengths), 100, 64
    R = 2 * n_spk * n_spk
    gen = torch.Generator().manual_seed(7)
    x_sbd = torch.rand(S, B, d, generator=gen)
    torch.manual_seed(5)
    rg, gc = RGCNConv(d, h, R, num_bases=bases), GraphConv(h, h)
    cot = torch.randn(S, B, h, generator=gen)
    # oracle (fp64 on CPU)
    xs = x_sbd.double().requires_grad_(True)
    P = {n: p.detach().double().requires_grad_(True) for n, p in list(rg.named_parameters()) + [("gc." + n, p) for n, p in gc.named_parameters()]}
    xn = GO.pack(xs, lengths)
    h1 = GO.rgcn(xn, ei, et, R, P["weight"], P["root"], P["bias"], P.get("comp"))
    h2 = GO.graph_conv(torch.relu(h1), ei, P["gc.lin_rel.weight"].t(), P["gc.lin_root.weight"].t(), P["gc.lin_rel.bias"])
    out_ref = GO.unpack(h2, lengths, S)
    (out_ref * cot.double()).sum().backward()
    # CUDA
    g = DialogueGraph(lengths, spk, wp, wf, n_spk)
    rg, gc = rg.cuda(), gc.cuda()
    xc = x_sbd.cuda().requires_grad_(True)
    out = g.unpack(gc(torch.relu(rg(g.pack(xc), g)), g))
    (out * cot.cuda()).sum().backward()
    H.assert_close(out.detach().cpu().numpy(), out_ref.detach().numpy(), "graph conv output", rtol=1e-4, atol_frac=1e-5)
    H.assert_close(xc.grad.cpu().numpy(), xs.grad.numpy(), "d input", rtol=1e-4, atol_frac=1e-5)
    got = dict(list(rg.named_parameters()) + [("gc." + n, p) for n, p in gc.named_parameters()])
    for n, p in P.items():
        H.assert_close(got[n].grad.cpu().numpy(), p.grad.numpy(), f"d {n}", rtol=1e-4, atol_frac=1e-5)


@pytest.mark.gpu
def test_pack_unpack_round_trip_and_errors():
    from gan_ffn_b200.graph import DialogueGraph
    lengths = [5, 9, 1]
    spk, _ = _speakers(lengths, 2)
    g = DialogueGraph(lengths, spk, 2, 2, 2)
    x = torch.rand(9, 3, 100, device="cuda")
    xn = g.pack(x)
    assert torch.equal(xn.cpu(), GO.pack(x.cpu(), lengths))
    back = g.unpack(xn)
    mask = torch.zeros(9, 3, 1)
    for b, L in enumerate(lengths):
        mask[:L, b] = 1
    assert torch.equal(back.cpu(), x.cpu() * mask)
    with pytest.raises(RuntimeError):
        DialogueGraph(lengths, spk, 2, 2, 2, device="cpu")
    with pytest.raises((ValueError, RuntimeError)):
        g.pack(torch.rand(9, 3, 101, device="cuda"))
