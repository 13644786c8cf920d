"""CPU: pins (i) our module constructors' default initialisation and (ii) the CPU oracle against
the fixtures the unmodified reference produced (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import O


@pytest.fixture(scope="module")
def G():
    return H.golden()


@pytest.fixture(scope="module")
def nets():
    return H.build_nets("cpu")


def test_default_init_is_bit_identical_to_reference(G, nets):
    ns, ffn = nets
    for k in H.NET_ORDER:
        names = [n for n, _ in ns[k].named_parameters()]
        assert names == [str(n) for n in G[f"{k}/w_names"]], f"{k}: parameter names/order differ from the reference"
        sums = np.array([p.detach().double().sum().item() for _, p in ns[k].named_parameters()])
        assert np.array_equal(sums, G[f"{k}/w_sum"]), f"{k}: default init differs from the reference under seed {H.SEED}"
    assert np.array_equal(np.array([ffn.fc.weight.double().sum().item(), ffn.fc.bias.double().sum().item()]),
                          G["ffn/fc_w_sum"])


def test_state_dict_keys_include_reference_quirks(nets):
    ns, ffn = nets
    sd = ns["visual_disc"].state_dict()
    assert "position_encoding.pe" in sd and "encoder_layer.linear1.weight" in sd and "object.weight" in sd
    assert "transformer_encoder.layers.7.norm2.bias" in sd
    assert {"lstm.weight_ih_l0", "smax_fc.weight", "fc.weight"} <= set(ffn.state_dict())
    n_live = sum(p.numel() for n, p in ns["acoustic_gen"].named_parameters() if not n.startswith("encoder_layer."))
    assert sum(p.numel() for p in ns["acoustic_gen"].parameters()) == 4175944
    assert n_live == 4175944 - 452548


@pytest.mark.parametrize("name", H.NET_ORDER)
def test_oracle_network_matches_reference_fixture(G, nets, name):
    ns, _ = nets
    batch = H.golden_batch()
    P = O.params_of(ns[name], requires_grad=True)
    x = H.net_inputs(batch)[name].clone().requires_grad_(True)
    y = H.oracle_forward(name, x, P)
    loss = (y * H.golden_cotangent(batch.seq_len)).sum() if name.endswith("gen") else O.bce(y, torch.ones_like(y))
    loss.backward()
    H.assert_close(y.detach(), G[f"{name}/out"], f"{name} output", atol_frac=1e-6)
    H.assert_close(loss.item(), G[f"{name}/loss"], f"{name} loss")
    H.assert_close(x.grad, G[f"{name}/dx"], f"{name} dx", atol_frac=H.RTOL)
    H.check_grads({k: v.grad for k, v in P.items() if v.grad is not None}, G, name)


def test_oracle_visual_discriminator_skips_object_on_100_wide_input(G, nets):
    ns, _ = nets
    batch = H.golden_batch()
    P = O.params_of(ns["visual_disc"], requires_grad=True)
    x = batch.acoustic.clone().requires_grad_(True)
    y = O.discriminator(x, P)
    loss = O.bce(y, torch.zeros_like(y))
    loss.backward()
    H.assert_close(y.detach(), G["visual_disc_fake/out"], "prob", atol_frac=1e-6)
    H.assert_close(loss.item(), G["visual_disc_fake/loss"], "loss")
    grads = {k: v.grad for k, v in P.items() if v.grad is not None}
    assert "object.weight" not in grads
    H.check_grads(grads, G, "visual_disc_fake")


def test_oracle_stage2_matches_reference_fixture(G, nets):
    ns, ffn = nets
    batch = H.golden_batch()
    Pa, Pv, Pt = (O.params_of(ns[k], requires_grad=True) for k in ("acoustic_gen", "visual_gen", "text_gen"))
    fw = ffn.fc.weight.detach().clone().requires_grad_(True)
    fb = ffn.fc.bias.detach().clone().requires_grad_(True)
    lp = O.gan_ffn(batch.acoustic, batch.visual, batch.text, Pa, Pv, Pt, fw, fb)
    lp_ = lp.transpose(0, 1).contiguous().view(-1, 6)
    loss = O.masked_nll(lp_, batch.label.view(-1), batch.umask, torch.tensor(H.synthetic.IEMOCAP_LOSS_WEIGHTS))
    loss.backward()
    H.assert_close(lp.detach(), G["ffn/log_prob"], "log_prob", atol_frac=1e-6)
    H.assert_close(loss.item(), G["ffn/loss"], "loss")
    grads = {"fc.weight": fw.grad, "fc.bias": fb.grad}
    for pre, P in (("acoustic_generator.", Pa), ("visual_generator.", Pv), ("text_generator.", Pt)):
        grads.update({pre + k: v.grad for k, v in P.items() if v.grad is not None})
    H.check_grads(grads, G, "ffn")


def test_oracle_stage1_substeps_and_adam_match_reference_fixture(G, nets):
    ns, _ = nets
    batch = H.golden_batch()
    S = batch.seq_len
    valid, fake = torch.ones(S, 3, 1), torch.zeros(S, 3, 1)
    Pd = O.params_of(ns["visual_disc"], requires_grad=True)
    Pg = O.params_of(ns["acoustic_gen"], requires_grad=True)
    # train_disc (train_IEMOCAP.py:217-225)
    fusion = O.generator(batch.acoustic, Pg, 10).detach()
    d_loss = (O.bce(O.discriminator(batch.visual, Pd), valid) + O.bce(O.discriminator(fusion, Pd), fake)) / 2.0
    d_loss.backward()
    H.assert_close(d_loss.item(), G["train_disc/loss"], "d_loss")
    H.check_grads({k: v.grad for k, v in Pd.items() if v.grad is not None}, G, "train_disc")
    # train_gen (train_IEMOCAP.py:246-250)
    for v in Pd.values():
        v.grad = None
    g_loss = O.bce(O.discriminator(O.generator(batch.acoustic, Pg, 10), Pd), valid)
    g_loss.backward()
    H.assert_close(g_loss.item(), G["train_gen/loss"], "g_loss")
    ggrads = {k: v.grad for k, v in Pg.items() if v.grad is not None}
    H.check_grads(ggrads, G, "train_gen")
    # Adam (train_IEMOCAP.py:292)
    names = [str(n) for n in G["adam/names"]]
    for i, n in enumerate(names):
        p = Pg[n].detach().clone()
        before = p.clone()
        O.adam_step(p, ggrads[n], torch.zeros_like(p), torch.zeros_like(p), 1, 1e-4, 0.5, 0.6)
        delta = (p - before).double().reshape(-1).numpy()[H.probe_index(p.numel())]
        H.check_adam_delta(delta, G["adam/delta_probe"][i], ggrads[n], 1e-4, n)
