"""GPU parity of every C-ABI primitive against fp64 CPU restatements (oracle/ganffn_oracle.py).

Tolerance (north_star): rtol 1e-4 in fp32.  Checked as |x - ref| <= 1e-4*|ref| + 1e-5*max|ref|
(the absolute floor covers entries that cancel to ~0)."""
import math

import numpy as np
import pytest
import torch

import helpers as H
from helpers import O

pytestmark = pytest.mark.gpu

ENGINES = [0, 1, 2]  # GANFFN_GEMM_AUTO, _SIMT, _TC (TC falls back to SIMT tiles for unsupported shapes)


@pytest.fixture(scope="module")
def L():
    from gan_ffn_b200._lib import lib
    return lib()


@pytest.fixture(autouse=True)
def _restore_engine(L):
    yield
    L.cdll.ganffn_set_gemm_engine(0)


def dev(t):
    return t.to("cuda", torch.float32).contiguous()


def P(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def close(a, e, what, rtol=1e-4, floor=1e-5):
    H.assert_close(a.detach().double().cpu().numpy(), e.detach().double().cpu().numpy(), what, rtol=rtol, atol_frac=floor)


def gemm_ws(L, M, N, K):
    n = int(L.cdll.ganffn_gemm_scratch_floats(M, N, K))
    return torch.empty(max(n, 1), device="cuda"), n


LINEAR_SHAPES = [(300, 300, 100), (282, 2048, 100), (282, 100, 2048), (50, 1, 16), (37, 6, 100), (3008, 1536, 512),
                 (3008, 100, 2048), (3008, 2048, 512), (129, 512, 512), (1, 100, 100), (36, 64, 100), (36, 16, 64),
                 # K <= 128 and N >= 256: the A-stationary tcgen05 kernel (ragged M, N and K tails, several tiles per CTA)
                 (3008, 2048, 100), (3008, 512, 100), (200, 1000, 64), (130, 260, 128), (6016, 2048, 100), (97, 4096, 36),
                 # narrow outputs (out-proj and the dX products of the d=100 networks)
                 (3008, 100, 100), (3008, 100, 300), (300, 100, 100), (6016, 128, 320), (257, 36, 64)]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("M,N,K", LINEAR_SHAPES)
def test_linear_fwd_plain(L, engine, M, N, K):
    L.cdll.ganffn_set_gemm_engine(engine)
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, generator=g, dtype=torch.float64)
    w = torch.randn(N, K, generator=g, dtype=torch.float64) / math.sqrt(K)
    b = torch.randn(N, generator=g, dtype=torch.float64)
    xd, wd, bd = dev(x), dev(w), dev(b)
    y = torch.empty(M, N, device="cuda")
    ws, n = gemm_ws(L, M, N, K)
    L.call("ganffn_linear_fwd", P(xd), P(wd), P(bd), None, P(y), None, M, N, K, 0, 0, 0.0, 0, 0, P(ws), n, stream())
    ref = xd.double().cpu() @ wd.double().cpu().T + bd.double().cpu()
    close(y, ref, f"linear {M}x{N}x{K}", rtol=1e-4, floor=2e-6)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("act,dba", [(1, 0), (2, 1), (3, 1), (0, 0)])
def test_linear_fwd_epilogues_with_dropout_and_residual(L, engine, act, dba):
    L.cdll.ganffn_set_gemm_engine(engine)
    M, N, K = 282, 512, 100
    g = torch.Generator().manual_seed(11 + act)
    x, w, b, r = (torch.randn(*s, generator=g) for s in ((M, K), (N, K), (N,), (M, N)))
    w = w / 10
    xd, wd, bd, rd = dev(x), dev(w), dev(b), dev(r)
    y = torch.empty(M, N, device="cuda")
    pre = torch.empty(M, N, device="cuda")
    p, seed, site = 0.2, 0xDEADBEEF12345, 201
    ws, n = gemm_ws(L, M, N, K)
    L.call("ganffn_linear_fwd", P(xd), P(wd), P(bd), P(rd), P(y), P(pre), M, N, K, act, dba, p, seed, site, P(ws), n,
           stream())
    from gan_ffn_b200.functional import dropout_mask
    mask = dropout_mask(M, N, p, seed, site).double().cpu()
    v = x.double() @ w.double().T + b.double()
    f = {0: lambda t: t, 1: torch.relu, 2: O.gelu, 3: torch.sigmoid}[act]
    if dba:
        pre_ref = v * mask
        ref = f(pre_ref) + r.double()
    else:
        pre_ref = v
        ref = f(v) * mask + r.double()
    close(pre, pre_ref, "pre-activation", floor=2e-6)
    close(y, ref, "epilogue output", floor=2e-6)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("M,N,K,p", [(3008, 100, 100, 0.1), (3008, 100, 100, 0.0), (3008, 100, 2048, 0.1), (6016, 100, 2048, 0.0),
                                     (130, 128, 260, 0.1), (97, 64, 100, 0.0), (300, 512, 512, 0.1), (3008, 512, 2048, 0.0),
                                     (50, 100, 36, 0.1), (282, 36, 100, 0.1)])
def test_linear_layernorm_tail(L, engine, M, N, K, p):
    """z = residual + drop(linear(x)), y = LayerNorm(z): fused in the tcgen05 epilogue / split-K fold for N <= 128,
    stand-alone LayerNorm elsewhere (torch transformer.py:944-982, post-norm)."""
    L.cdll.ganffn_set_gemm_engine(engine)
    g = torch.Generator().manual_seed(M + 3 * N + 5 * K)
    x, w, b, r = (torch.randn(*s, generator=g) for s in ((M, K), (N, K), (N,), (M, N)))
    w = w / math.sqrt(K)
    gamma, beta = torch.randn(N, generator=g), torch.randn(N, generator=g)
    xd, wd, bd, rd, gd, btd = (dev(t) for t in (x, w, b, r, gamma, beta))
    z = torch.empty(M, N, device="cuda")
    y = torch.empty(M, N, device="cuda")
    seed, site = 0xABCDEF0123, 33
    ws, n = gemm_ws(L, M, N, K)
    L.call("ganffn_linear_ln_fwd", P(xd), P(wd), P(bd), P(rd), P(gd), P(btd), P(z), P(y), M, N, K, p, seed, site, P(ws), n,
           stream())
    from gan_ffn_b200.functional import dropout_mask
    mask = dropout_mask(M, N, p, seed, site).double().cpu()
    zr = (x.double() @ w.double().T + b.double()) * mask + r.double()
    yr = torch.nn.functional.layer_norm(zr, (N,), gamma.double(), beta.double(), 1e-5)
    close(z, zr, f"pre-norm sum {M}x{N}x{K}", floor=2e-6)
    close(y, yr, f"LayerNorm output {M}x{N}x{K}", floor=2e-6)


@pytest.mark.parametrize("T,d,p", [(3008, 100, 0.2), (6016, 100, 0.0), (130, 100, 0.2), (31, 64, 0.2), (1, 128, 0.0), (257, 100, 0.2)])
def test_discriminator_head_kernels(L, T, d, p):
    """gelu -> fc1 (d->64) -> fc2 (64->16) -> fc3 (16->1) -> sigmoid with dropout before each activation, forward and
    backward in one kernel each (csrc/head.cu; reference model.py:1320-1327), against an fp64 restatement with the
    exported masks; frozen variant (no parameter gradients) gives the same dx."""
    from gan_ffn_b200.functional import dropout_mask
    g = torch.Generator().manual_seed(T + d)
    x = torch.randn(T, d, generator=g)
    w1, b1 = torch.randn(64, d, generator=g) / math.sqrt(d), torch.randn(64, generator=g) * 0.1
    w2, b2 = torch.randn(16, 64, generator=g) / 8, torch.randn(16, generator=g) * 0.1
    w3, b3 = torch.randn(16, generator=g) / 4, torch.randn(1, generator=g) * 0.1
    dp = torch.randn(T, generator=g)
    seed, site0 = 0x5EEDF00D, 200
    dv = [dev(t) for t in (x, w1, b1, w2, b2, w3, b3, dp)]
    xd, w1d, b1d, w2d, b2d, w3d, b3d, dpd = dv
    e = lambda *sh: torch.empty(*sh, device="cuda")
    g0, f1, a1, f2, a2, prob = e(T, d), e(T, 64), e(T, 64), e(T, 16), e(T, 16), e(T)
    L.call("ganffn_disc_head_fwd", P(xd), P(w1d), P(b1d), P(w2d), P(b2d), P(w3d), P(b3d), P(g0), P(f1), P(a1), P(f2), P(a2), P(prob),
           T, d, p, seed, site0, stream())
    m1 = dropout_mask(T, 64, p, seed, site0 + 1).double().cpu()
    m2 = dropout_mask(T, 16, p, seed, site0 + 2).double().cpu()
    m3 = dropout_mask(T, 1, p, seed, site0 + 3).double().cpu().view(-1)
    X = x.double().requires_grad_(True)
    W1, B1, W2, B2, W3, B3 = (t.double().requires_grad_(True) for t in (w1, b1, w2, b2, w3, b3))
    G0 = O.gelu(X)
    F1 = (G0 @ W1.T + B1) * m1
    A1 = O.gelu(F1)
    F2 = (A1 @ W2.T + B2) * m2
    A2 = O.gelu(F2)
    PR = torch.sigmoid((A2 @ W3 + B3) * m3)
    for got, ref, name in ((g0, G0, "g0"), (f1, F1, "f1"), (a1, A1, "a1"), (f2, F2, "f2"), (a2, A2, "a2"), (prob, PR, "prob")):
        close(got, ref, f"head {name}", floor=2e-6)
    (PR * dp.double()).sum().backward()
    grads = [torch.zeros_like(t) for t in (w1d, b1d, w2d, b2d, w3d, b3d)]
    dx = e(T, d)
    L.call("ganffn_disc_head_bwd", P(dpd), P(prob), P(xd), P(g0), P(f1), P(a1), P(f2), P(a2), P(w1d), P(w2d), P(w3d), P(dx),
           *[P(t) for t in grads], T, d, p, seed, site0, stream())
    close(dx, X.grad, "head dx", floor=5e-6)
    for got, ref, name in zip(grads, (W1, B1, W2, B2, W3, B3), ("dw1", "db1", "dw2", "db2", "dw3", "db3")):
        close(got, ref.grad.reshape(got.shape), f"head {name}", floor=5e-6)
    dx2 = e(T, d)
    L.call("ganffn_disc_head_bwd", P(dpd), P(prob), P(xd), None, P(f1), None, P(f2), None, P(w1d), P(w2d), P(w3d), P(dx2),
           None, None, None, None, None, None, T, d, p, seed, site0, stream())
    assert torch.equal(dx, dx2), "the frozen variant must give the same data gradient"


def test_dropout_mask_statistics_and_determinism(L):
    from gan_ffn_b200.functional import dropout_mask
    for p in (0.1, 0.2, 0.6):
        m = dropout_mask(4096, 512, p, 1234, 17)
        keep = (m > 0).float().mean().item()
        assert abs(keep - (1 - p)) < 3e-3, (p, keep)
        assert torch.allclose(m[m > 0], torch.tensor(1 / (1 - p), device="cuda"))
        assert torch.equal(m, dropout_mask(4096, 512, p, 1234, 17))
        assert not torch.equal(m, dropout_mask(4096, 512, p, 1234, 18))
        assert not torch.equal(m, dropout_mask(4096, 512, p, 1235, 17))
    assert torch.equal(dropout_mask(16, 8, 0.0, 1, 1), torch.ones(16, 8, device="cuda"))


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("M,N,K", [(282, 2048, 100), (282, 100, 2048), (3008, 512, 2048), (3008, 1536, 512), (50, 1, 16),
                                   (36, 16, 64), (3008, 100, 2048), (200, 64, 1000), (130, 128, 260), (3008, 100, 100), (3008, 300, 100),
                                   (300, 128, 36)])
def test_linear_dgrad_and_wgrad(L, engine, M, N, K):
    L.cdll.ganffn_set_gemm_engine(engine)
    g = torch.Generator().manual_seed(M + N + K)
    dy = torch.randn(M, N, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(N)
    x = torch.randn(M, K, generator=g)
    r = torch.randn(M, K, generator=g)
    dyd, wd, xd, rd = dev(dy), dev(w), dev(x), dev(r)
    dx = torch.empty(M, K, device="cuda")
    ws, n = gemm_ws(L, M, K, N)
    L.call("ganffn_linear_dgrad", P(dyd), P(wd), P(rd), P(dx), M, N, K, P(ws), n, stream())
    close(dx, dy.double() @ w.double() + r.double(), "dgrad", floor=2e-6)

    dw0 = torch.randn(N, K, generator=g)
    db0 = torch.randn(N, generator=g)
    for acc in (0, 1):
        dw, db = dev(dw0).clone(), dev(db0).clone()
        wsn = int(L.cdll.ganffn_wgrad_scratch_floats(M, N, K))
        ws2 = torch.empty(max(wsn, 1), device="cuda")
        L.call("ganffn_linear_wgrad", P(dyd), P(xd), P(dw), P(db), M, N, K, acc, P(ws2), stream())
        rw = dy.double().T @ x.double() + (dw0.double() if acc else 0)
        rb = dy.double().sum(0) + (db0.double() if acc else 0)
        close(dw, rw, f"wgrad acc={acc}", floor=2e-6)
        close(db, rb, f"bias grad acc={acc}", floor=2e-6)


def _attn_ref(qkv, S, B, d, H_, mask=None):
    hd = d // H_
    q, k, v = qkv.split(d, dim=-1)
    f = lambda t: t.reshape(S, B, H_, hd).permute(1, 2, 0, 3)
    q, k, v = f(q), f(k), f(v)
    p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), -1)
    if mask is not None:
        p = p * mask
    return (p @ v).permute(2, 0, 1, 3).reshape(S, B, d)


@pytest.mark.parametrize("S,B,d,nh,p", [(94, 4, 100, 10, 0.0), (110, 3, 512, 8, 0.0), (5, 2, 32, 4, 0.0), (1, 2, 100, 10, 0.0),
                                        (37, 3, 100, 10, 0.1), (33, 2, 512, 8, 0.1), (110, 2, 64, 4, 0.1), (7, 3, 64, 2, 0.1)])
def test_attention_fwd_bwd(L, S, B, d, nh, p):
    g = torch.Generator().manual_seed(S * 13 + d)
    qkv = torch.randn(S, B, 3 * d, generator=g).double()
    do = torch.randn(S, B, d, generator=g).double()
    qd, dod = dev(qkv), dev(do)
    o = torch.empty(S, B, d, device="cuda")
    lse = torch.empty(B * nh * S, device="cuda")
    dqkv = torch.empty(S, B, 3 * d, device="cuda")
    seed, site = 987654321, 16
    L.call("ganffn_attention_fwd", P(qd), P(o), P(lse), S, B, d, nh, p, seed, site, stream())
    L.call("ganffn_attention_bwd", P(qd), P(o), P(lse), P(dod), P(dqkv), S, B, d, nh, p, seed, site, stream())
    mask = None
    if p > 0:
        from gan_ffn_b200.functional import dropout_mask
        mask = dropout_mask(B * nh * S, S, p, seed, site, row_stride=(S + 3) // 4 * 4).double().cpu().view(B, nh, S, S)
    qr = qd.double().cpu().requires_grad_(True)
    ref = _attn_ref(qr, S, B, d, nh, mask)
    ref.backward(dod.double().cpu())
    close(o, ref, "attention out", floor=2e-6)
    close(dqkv, qr.grad, "attention dqkv", floor=1e-5)


@pytest.mark.parametrize("T,d", [(282, 100), (3008, 512), (5, 32), (1, 100)])
def test_layernorm_fwd_bwd(L, T, d):
    g = torch.Generator().manual_seed(T + d)
    z = (torch.randn(T, d, generator=g) * 2 + 0.5).double()
    gam = torch.randn(d, generator=g).double()
    bet = torch.randn(d, generator=g).double()
    dy = torch.randn(T, d, generator=g).double()
    zd, gd, bd, dyd = dev(z), dev(gam), dev(bet), dev(dy)
    y = torch.empty(T, d, device="cuda")
    L.call("ganffn_layernorm_fwd", P(zd), P(gd), P(bd), P(y), T, d, stream())
    zr = zd.double().cpu().requires_grad_(True)
    gr = gd.double().cpu().requires_grad_(True)
    br = bd.double().cpu().requires_grad_(True)
    ref = O.layer_norm(zr, gr, br)
    ref.backward(dyd.double().cpu())
    close(y, ref, "layernorm", floor=2e-6)
    ws = torch.empty(int(L.cdll.ganffn_layernorm_scratch_floats(T, d)), device="cuda")
    for acc, p in ((0, 0.0), (1, 0.1)):
        dz = torch.empty(T, d, device="cuda")
        dzd = torch.empty(T, d, device="cuda")
        dg = torch.ones(d, device="cuda")
        db = torch.ones(d, device="cuda")
        L.call("ganffn_layernorm_bwd", P(dyd), P(zd), P(gd), P(dz), P(dzd) if p > 0 else None, P(dg), P(db), T, d, acc, p,
               77, 19, P(ws), stream())
        close(dz, zr.grad, "ln dz", floor=1e-5)
        close(dg, gr.grad + acc, "ln dgamma", floor=1e-5)
        close(db, br.grad + acc, "ln dbeta", floor=1e-5)
        if p > 0:
            from gan_ffn_b200.functional import dropout_mask
            close(dzd, zr.grad * dropout_mask(T, d, p, 77, 19).double().cpu(), "ln dz*mask", floor=1e-5)


def test_posenc(L):
    S, B, d = 94, 5, 100
    x = torch.rand(S, B, d)
    pe = O.positional_table(d)
    y = torch.empty(S, B, d, device="cuda")
    x_d, pe_d = dev(x), dev(pe)
    L.call("ganffn_posenc_fwd", P(x_d), P(pe_d), P(y), S, B, d, 0.0, 0, stream())
    assert torch.equal(y.cpu(), O.positional_encoding(x))          # one fp32 add: bit exact
    with pytest.raises(ValueError, match="110"):
        L.call("ganffn_posenc_fwd", P(x_d), P(pe_d), P(y), 111, 1, d, 0.0, 0, stream())


def test_fuse_cls_and_masked_nll(L):
    S, B, dh, C = 23, 5, 100, 6
    T = S * B
    g = torch.Generator().manual_seed(5)
    a, v, t = (torch.randn(T, dh, generator=g) for _ in range(3))
    w = torch.randn(C, dh, generator=g) / 10
    b = torch.randn(C, generator=g)
    target = torch.randint(0, C, (T,), generator=g)
    mask = (torch.rand(T, generator=g) > 0.3).float()
    cw = torch.tensor(H.synthetic.IEMOCAP_LOSS_WEIGHTS)
    fusion = torch.empty(T, dh, device="cuda")
    logp = torch.empty(T, C, device="cuda")
    ad, vd, td, wd, bd = dev(a), dev(v), dev(t), dev(w), dev(b)
    L.call("ganffn_fuse_cls_fwd", P(ad), P(vd), P(td), P(wd), P(bd), P(fusion), P(logp), T, dh, C, stream())
    ar, vr, tr, wr, br = (z.double().requires_grad_(True) for z in (a, v, t, w, b))
    lp_ref = torch.log_softmax((ar + vr + tr) @ wr.T + br, -1)
    close(logp, lp_ref, "log_prob", floor=2e-6)
    for weight in (cw, None):
        out = torch.empty(2, device="cuda")
        tgt, msk = target.cuda(), mask.cuda()
        wt = None if weight is None else weight.cuda()
        L.call("ganffn_masked_nll_fwd", P(logp), P(tgt), P(msk), P(wt), P(out), T, C, 0.0, stream())
        loss_ref = O.masked_nll(lp_ref, target, mask.view(1, -1), None if weight is None else weight.double())
        close(out[0], loss_ref, "masked nll")
        dl = torch.ones(1, device="cuda")
        dpred = torch.empty(T, C, device="cuda")
        L.call("ganffn_masked_nll_bwd", P(dl), P(out), P(tgt), P(msk), P(wt), P(dpred), T, C, stream())
        (dlp_ref,) = torch.autograd.grad(loss_ref, lp_ref, retain_graph=True)
        close(dpred, dlp_ref, "masked nll grad")
    # backward of fusion+classifier given d_log_prob
    for z in (ar, vr, tr, wr, br):
        z.grad = None
    loss_ref = O.masked_nll(lp_ref, target, mask.view(1, -1), cw.double())
    loss_ref.backward()
    dfus = torch.empty(T, dh, device="cuda")
    dw = torch.empty(C, dh, device="cuda")
    db = torch.empty(C, device="cuda")
    ws = torch.empty(int(L.cdll.ganffn_fuse_cls_scratch_floats(T, dh, C)), device="cuda")
    # dpred currently holds the unweighted gradient; recompute the weighted one
    out = torch.empty(2, device="cuda")
    tgt_d, msk_d, cw_d, one_d = target.cuda(), mask.cuda(), cw.cuda(), torch.ones(1, device="cuda")   # keep alive
    L.call("ganffn_masked_nll_fwd", P(logp), P(tgt_d), P(msk_d), P(cw_d), P(out), T, C, 0.0, stream())
    L.call("ganffn_masked_nll_bwd", P(one_d), P(out), P(tgt_d), P(msk_d), P(cw_d), P(dpred), T, C, stream())
    L.call("ganffn_fuse_cls_bwd", P(dpred), P(logp), P(fusion), P(wd), P(dfus), P(dw), P(db), T, dh, C, 0, P(ws), stream())
    close(dfus, ar.grad, "d_fusion", floor=1e-5)
    close(dw, wr.grad, "d fc.weight", floor=1e-5)
    close(db, br.grad, "d fc.bias", floor=1e-5)


def test_masked_nll_den_override(L):
    T, C = 40, 7
    g = torch.Generator().manual_seed(9)
    lp = torch.log_softmax(torch.randn(T, C, generator=g), -1)
    tgt = torch.randint(0, C, (T,), generator=g)
    msk = torch.ones(T)
    out = torch.empty(2, device="cuda")
    lp_d, tgt_d, msk_d = dev(lp), tgt.cuda(), msk.cuda()
    L.call("ganffn_masked_nll_fwd", P(lp_d), P(tgt_d), P(msk_d), None, P(out), T, C, 80.0, stream())
    assert abs(out[0].item() - O.masked_nll(lp.double(), tgt, msk.view(1, -1)).item() * 0.5) < 1e-5
    assert out[1].item() == 80.0


def test_bce_fwd_bwd_including_saturation(L):
    n = 3008
    g = torch.Generator().manual_seed(3)
    prob = torch.rand(n, generator=g)
    prob[:4] = torch.tensor([0.0, 1.0, 1e-30, 1 - 1e-7])      # log clamp at -100 (torch BCELoss)
    for tval in (1.0, 0.0):
        tgt = torch.full((n,), tval)
        out = torch.empty(1, device="cuda")
        prob_d, tgt_d = dev(prob), dev(tgt)
        L.call("ganffn_bce_fwd", P(prob_d), P(tgt_d), P(out), n, 1.0, stream())
        ref = torch.nn.functional.binary_cross_entropy(prob, tgt)
        assert abs(out.item() - ref.item()) <= 1e-4 * abs(ref.item())
    pr = torch.rand(n, generator=g) * 0.98 + 0.01
    prr = pr.double().requires_grad_(True)
    ref = O.bce(prr, torch.ones(n, dtype=torch.float64))
    ref.backward()
    dp = torch.empty(n, device="cuda")
    half_d, pr_d, ones_d = torch.full((1,), 0.5, device="cuda"), dev(pr), torch.ones(n, device="cuda")
    L.call("ganffn_bce_bwd", P(half_d), P(pr_d), P(ones_d), P(dp), n, 1.0, stream())
    close(dp, prr.grad * 0.5, "bce grad")


@pytest.mark.parametrize("wd,gs", [(0.0, 1.0), (0.008, 0.5)])
def test_adam_matches_torch_optim(L, wd, gs):
    n = 100003   # exercises the n % 4 tail
    g = torch.Generator().manual_seed(21)
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * 0.01 for _ in range(3)]
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pt], lr=1e-4, betas=(0.5, 0.6), weight_decay=wd)
    p, m, v = dev(p0), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step, gr in enumerate(grads, 1):
        pt.grad = gr * gs
        opt.step()
        gr_d = dev(gr)
        L.call("ganffn_adam_step", P(p), P(gr_d), P(m), P(v), n, step, 1e-4, 0.5, 0.6, 1e-8, wd, gs, stream())
    delta = (p.cpu() - p0).double()
    ref = (pt.detach() - p0).double()
    # three steps of at most lr each, compared at rtol 1e-4 of the total update plus one fp32 ulp of the parameter
    tol = 1e-4 * 3e-4 + 2 * torch.finfo(torch.float32).eps * p0.double().abs()
    assert ((delta - ref).abs() <= tol).all(), float((delta - ref).abs().max())
    with pytest.raises(ValueError):
        L.call("ganffn_adam_step", P(p), P(p), P(m), P(v), n, 0, 1e-4, 0.5, 0.6, 1e-8, wd, gs, stream())
