"""GAN_FFN_DialogueRNN (BASELINE config 5).

CPU: the DialogueRNN head restatement (gan_ffn_b200/dialogue_rnn.py, stock PyTorch) against fixtures written by the
unmodified reference (oracle/make_golden_dialogue_rnn.py): default initialisation bit-identical under the seed,
log-probabilities, input gradient and parameter-gradient norms at rtol 1e-5 for every attention variant.
GPU: the whole model -- fused features from the sm_100a kernels feeding the head -- against the reference's
log-probabilities (rtol 1e-4, north_star) and state_dict key parity."""
import os
import sys

import numpy as np
import pytest
import torch

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "dialogue_rnn_ref_seed3407.npz"), allow_pickle=False)
SEED = 3407
DIMS = dict(D_m=100, D_g=500, D_p=500, D_e=100, D_h=100)
HEAD_CONFIGS = [("general", False), ("simple", False), ("general2", True), ("dot", False), ("concat", False)]


def _head_inputs(S=9, B=3):
    from gan_ffn_b200 import synthetic
    b = synthetic.make_batch(n_dialogues=B, lengths=[S, S - 4, S - 2], seed=SEED + 1)
    g = torch.Generator().manual_seed(SEED + 2)
    U = torch.rand(S, B, DIMS["D_m"], generator=g) * b.umask.t().unsqueeze(2)
    return U, b.qmask, b.umask


@pytest.mark.parametrize("att,listener", HEAD_CONFIGS)
def test_head_matches_reference_fixture(att, listener):
    from gan_ffn_b200.dialogue_rnn import BiModel
    key = f"head/{att}/{int(listener)}"
    U, qmask, umask = _head_inputs()
    torch.manual_seed(SEED)
    d_m = DIMS["D_m"] if att != "dot" else DIMS["D_g"]
    m = BiModel(d_m, DIMS["D_g"], DIMS["D_p"], DIMS["D_e"], DIMS["D_h"], n_classes=6, listener_state=listener,
                context_attention=att, D_a=100, dropout_rec=0.1, dropout=0.6).eval()
    p_sum = np.array([p.detach().double().sum().item() for _, p in m.named_parameters()])
    # fp64 sums of bit-identical parameters: only the summation order of torch's parallel reduction may differ (last bits)
    assert np.allclose(p_sum, GOLD[key + "/p_sum"], rtol=1e-12, atol=1e-12), "default initialisation differs from the reference's"
    Ux = U if att != "dot" else torch.cat([U] * 5, dim=2)
    Ux = Ux.clone().requires_grad_(True)
    lp, alpha, alpha_f, alpha_b = m(Ux, qmask, umask)
    w = torch.linspace(0.5, 1.5, lp.numel()).view_as(lp)
    (lp * w).sum().backward()
    H.assert_close(lp.detach().numpy(), GOLD[key + "/log_prob"], f"{key} log_prob", rtol=1e-5, atol_frac=1e-6)
    H.assert_close(Ux.grad.numpy(), GOLD[key + "/dU"], f"{key} dU", rtol=1e-4, atol_frac=1e-5)
    H.assert_close(alpha[-1].detach().numpy(), GOLD[key + "/alpha_last"], f"{key} alpha", rtol=1e-5, atol_frac=1e-6)
    names = [str(n) for n in GOLD[key + "/g_names"]]
    got = {n: p.grad.double().norm().item() for n, p in m.named_parameters() if p.grad is not None}
    assert sorted(got) == sorted(names)
    for n, ref in zip(names, GOLD[key + "/g_norm"]):
        assert abs(got[n] - ref) <= 1e-4 * ref + 1e-10, (n, got[n], ref)


def test_reverse_seq_semantics():
    from gan_ffn_b200.dialogue_rnn import BiModel
    X = torch.arange(5 * 3 * 2, dtype=torch.float32).view(5, 3, 2) + 1
    mask = torch.tensor([[1, 1, 1, 1, 1], [1, 1, 0, 0, 0], [1, 1, 1, 0, 0]], dtype=torch.float32)
    R = BiModel._reverse_seq(X, mask)
    assert R.shape == (5, 3, 2)
    assert torch.equal(R[:, 0], X[:, 0].flip(0))
    assert torch.equal(R[:2, 1], X[:2, 1].flip(0)) and torch.count_nonzero(R[2:, 1]) == 0
    assert torch.equal(R[:3, 2], X[:3, 2].flip(0)) and torch.count_nonzero(R[3:, 2]) == 0
    # all dialogues shorter than the padded length: output is max(len) long, as pad_sequence gives
    assert BiModel._reverse_seq(X, mask[1:]).shape == (3, 2, 2)


@pytest.mark.gpu
def test_gan_ffn_dialogue_rnn_matches_reference_fixture():
    import gan_ffn_b200 as G
    from gan_ffn_b200 import synthetic
    torch.manual_seed(SEED)
    ga, gv, gt = G.AcousticGenerator(100, dropout=0.2), G.VisualGenerator(100, dropout=0.2), G.TextGenerator(100, dropout=0.2)
    model = G.GAN_FFN_DialogueRNN(ga, gv, gt, DIMS["D_m"], DIMS["D_g"], DIMS["D_p"], DIMS["D_e"], DIMS["D_h"], 100, 6, False,
                                  "general", 0.1, 0.6)
    assert list(model.state_dict().keys()) == [str(k) for k in GOLD["model/state_keys"]]
    model = model.to("cuda").eval()
    b = synthetic.make_batch(n_dialogues=3, lengths=[12, 7, 10], seed=SEED).to("cuda")
    with torch.no_grad():
        fusion = model.fusion(b.acoustic, b.visual, b.text)
        lp = model(b.acoustic, b.visual, b.text, b.qmask, b.umask)[0]
    H.assert_close(fusion.cpu().numpy(), GOLD["model/fusion"], "fused features", rtol=1e-4, atol_frac=1e-5)
    H.assert_close(lp.cpu().numpy(), GOLD["model/log_prob"], "GAN_FFN_DialogueRNN log_prob", rtol=1e-4, atol_frac=1e-5)
    # and it trains: gradients reach the generators' arenas through the head
    model.train()
    lp = model(b.acoustic, b.visual, b.text, b.qmask, b.umask)[0]
    lp_ = lp.transpose(0, 1).contiguous().view(-1, 6)
    loss = G.MaskedNLLLoss()(lp_, b.label.view(-1), b.umask)
    loss.backward()
    for gen in (ga, gv, gt):
        g = gen.arena().grad
        assert torch.isfinite(g).all() and float(g.abs().sum()) > 0


@pytest.mark.gpu
def test_dialogue_rnn_trainer_step_eager_and_graph_replay_agree():
    """ClassifierTrainer over GAN_FFN_DialogueRNN (the train_IEMOCAP_DialogueRNN.py loop body): the step recorded into
    a CUDA graph (the head gets the longest dialogue from the host, so nothing is read back) equals the eager step."""
    import gan_ffn_b200 as G
    from gan_ffn_b200 import synthetic, train

    def build():
        torch.manual_seed(SEED)
        ga, gv, gt = G.AcousticGenerator(100), G.VisualGenerator(100), G.TextGenerator(100)
        m = G.GAN_FFN_DialogueRNN(ga, gv, gt, 100, 50, 50, 40, 40, 30, 6, False, "general", 0.1, 0.6).to("cuda")
        cls = train.ClassifierTrainer(m, torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device="cuda"))
        for mod in m.modules():                      # dropout off everywhere: the two runs must be comparable
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        m.train = lambda mode=True: torch.nn.Module.train(m, False)   # keep the generators' kernels in eval mode
        return m, cls

    batch = synthetic.make_batch(n_dialogues=3, lengths=[12, 7, 10], seed=SEED).to("cuda")
    m1, eager = build()
    m2, graphed = build()
    stepper = train.GraphedTrainStep(None, graphed, seed=1)
    for it in range(3):
        l1, p1, _ = eager.step(batch, train=True)
        out = stepper(batch)
        assert abs(float(l1) - float(out["loss"])) <= 1e-4 * abs(float(l1)), (it, float(l1), float(out["loss"]))
        assert torch.equal(p1, out["pred"])
    assert stepper.kernels_per_replay, "the third call must have been a graph replay"
    torch.cuda.synchronize()
    for (n1, a), (n2, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert n1 == n2
        assert float((a.detach() - b.detach()).abs().max()) <= 3 * 2e-4 * 1.001, n1      # three Adam steps of lr 1e-4 at most
