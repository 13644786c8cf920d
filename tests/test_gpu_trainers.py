"""GPU parity of the two *trainers* against the reference port.

``GANTrainer.batch`` must run the twelve sub-steps of reference train_IEMOCAP.py:355-382 in the reference's order
(pairings, overwritten loss keys, six Adam optimizers at lr, lr/2 and lr*1.1, betas (0.5, 0.6)) and
``ClassifierTrainer.step`` the body of train_IEMOCAP.py:127-170 (Adam lr 1e-4, L2 0.008).  The checker is
``oracle/reference_port.PortTrainer`` -- stock ``torch.nn`` modules and ``torch.optim.Adam`` on the CPU, itself pinned
to the unmodified reference's fixtures (tests/test_port_and_host.py) -- started from identical weights.  Dropout is
forced off on both sides (modules pinned in eval mode: ``.train()`` is a no-op), because torch's generator streams
cannot be reproduced by a fused kernel; everything else (order, optimizers, learning rates, losses) is live.

Bars: the six surviving stage-1 losses and the stage-2 loss rtol 1e-4; post-step weights per network: the mean
|difference| of the *accumulated update* below 1 % of that network's learning rate, 99.9 % of the entries within
5 % of lr, and no entry further than (optimizer steps on that network) x 2 x lr (the first-step bound of Adam for
noise-level gradients, see helpers.check_adam_delta).  The PARITY lines are appended to
gpurun_out/parity_trainers.txt so that the measured distances can be committed under profiles/."""
import json
import os

import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu

LOSS_KEYS = ["acoustic_D_loss", "acoustic_G_loss", "text_D_loss", "text_G_loss", "visual_D_loss", "visual_G_loss"]
# optimizer steps per network in one batch (train_IEMOCAP.py:355-382) and its learning rate (:292-297)
STEPS = {"acoustic_gen": 2, "text_gen": 2, "visual_gen": 2, "acoustic_disc": 2, "visual_disc": 2, "text_disc": 2}
LR = {"acoustic_gen": 1e-4, "visual_gen": 1e-4, "text_gen": 1.1e-4, "acoustic_disc": 5e-5, "visual_disc": 5e-5, "text_disc": 5e-5}


def _pin_eval(module):
    module.eval()
    module.train = lambda mode=True: module      # the loop bodies call .train(): keep dropout off on both sides
    return module


def _log(line):
    print(line)
    out = os.path.join(H.ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_trainers.txt"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


def _build_pair():
    """(our trainers on the GPU, the port's trainer on the CPU) from identical weights."""
    from gan_ffn_b200 import synthetic, train
    from oracle import reference_port as RP
    pnets, pffn = RP.build()
    nets, ffn = train.build_networks(device="cuda")
    for k in nets:                                  # same attribute names, hence the same state_dict keys
        missing = nets[k].load_state_dict(pnets[k].state_dict(), strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
    ffn.fc.load_state_dict(pffn.fc.state_dict())
    for m in list(nets.values()) + [ffn]:
        _pin_eval(m)
    for m in list(pnets.values()) + [pffn]:
        _pin_eval(m)
    w = torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS)
    gan = train.GANTrainer(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], nets["acoustic_disc"],
                           nets["visual_disc"], nets["text_disc"])
    cls = train.ClassifierTrainer(ffn, w.cuda())
    port = RP.PortTrainer(pnets, pffn, w)
    return nets, ffn, gan, cls, pnets, pffn, port


def _flat(module):
    return torch.cat([p.detach().reshape(-1).double().cpu() for _, p in sorted(module.named_parameters())])


def _compare_updates(name, ours_before, ours_after, port_before, port_after, lr, steps):
    du, dp = ours_after - ours_before, port_after - port_before
    err = (du - dp).abs()
    moved = float(dp.abs().mean())
    stats = {"net": name, "lr": lr, "optimizer_steps": steps, "mean_abs_update_port": moved,
             "mean_err_over_lr": float(err.mean()) / lr, "p999_err_over_lr": float(err.quantile(0.999)) / lr
             if err.numel() < 2 ** 24 else float(err[torch.randperm(err.numel())[:2 ** 23]].quantile(0.999)) / lr,
             "max_err_over_lr": float(err.max()) / lr}
    _log("PARITY trainer-update " + json.dumps(stats))
    assert moved > 0.2 * lr, f"{name}: the port's weights barely moved ({moved:.2e}); the test would be vacuous"
    assert stats["mean_err_over_lr"] < 0.01, stats
    assert stats["p999_err_over_lr"] < 0.05, stats
    assert stats["max_err_over_lr"] <= steps * 2.0 * 1.001, stats


@pytest.mark.parametrize("freeze_disc", [False, True], ids=["reference-train_gen", "frozen-disc"])
@pytest.mark.parametrize("batch_disc", [False, True], ids=["two-pass-disc", "batched-disc"])
@pytest.mark.parametrize("overlap", [False, True], ids=["serial", "lanes+chains"])
def test_gan_batch_and_classifier_step_match_reference_port(overlap, batch_disc, freeze_disc):
    """batch_disc=False is the reference's train_disc body verbatim (two discriminator passes); True runs them as one
    pass over [real | fake] (train.train_disc_batched) -- both must match the port.  freeze_disc=False is the
    reference's train_gen body verbatim; True skips the discriminator weight gradients that train_gen computes and
    nothing reads (train.train_gen_frozen_disc) -- losses and every network's update must not change."""
    from gan_ffn_b200 import synthetic
    nets, ffn, gan, cls, pnets, pffn, port = _build_pair()
    gan.overlap = cls.overlap = overlap
    gan.batch_disc = batch_disc
    gan.freeze_disc = freeze_disc
    batch = synthetic.make_batch(n_dialogues=4, lengths=[14, 9, 12, 5], seed=11)
    cb = batch.to("cuda")

    before_o = {k: _flat(m) for k, m in nets.items()}
    before_p = {k: _flat(m) for k, m in pnets.items()}
    fc_before = ffn.fc.weight.detach().double().cpu().clone()

    # ---- stage 1: twelve sub-steps -------------------------------------------------------------------------------
    ours = gan.batch(cb)
    ref = port.gan_batch(batch)
    assert sorted(ours) == sorted(ref) == LOSS_KEYS, "the six surviving loss keys of train_IEMOCAP.py:355-382"
    for k in LOSS_KEYS:
        a, e = float(ours[k]), float(ref[k])
        _log(f"PARITY trainer-loss {k} ours={a:.8f} port={e:.8f} rel={abs(a - e) / abs(e):.2e} overlap={overlap} batch_disc={batch_disc}")
        assert abs(a - e) <= H.RTOL * abs(e), (k, a, e)
    torch.cuda.synchronize()
    for k in nets:
        _compare_updates(k, before_o[k], _flat(nets[k]), before_p[k], _flat(pnets[k]), LR[k], STEPS[k])

    # ---- stage 2 on the weights stage 1 left behind ---------------------------------------------------------------
    mid_o = {k: _flat(nets[k]) for k in ("acoustic_gen", "visual_gen", "text_gen")}
    mid_p = {k: _flat(pnets[k]) for k in mid_o}
    loss, pred, labels = cls.step(cb, train=True)
    loss_ref, pred_ref = port.classifier_step(batch, train=True)
    a, e = float(loss), float(loss_ref)
    _log(f"PARITY trainer-loss stage2 ours={a:.8f} port={e:.8f} rel={abs(a - e) / abs(e):.2e} overlap={overlap}")
    assert abs(a - e) <= 2 * H.RTOL * abs(e), (a, e)          # starts from weights that already differ by ~1e-2 lr
    agree = float((pred.cpu() == pred_ref).float().mean())
    assert agree >= 0.95, f"argmax predictions agree on {agree:.3f} of the slots"
    assert torch.equal(labels.cpu(), batch.label.view(-1))
    for k in mid_o:
        _compare_updates(k + "/stage2", mid_o[k], _flat(nets[k]), mid_p[k], _flat(pnets[k]), 1e-4, 1)
    d_fc = (ffn.fc.weight.detach().double().cpu() - fc_before) - (pffn.fc.weight.detach().double() - fc_before)
    assert float(d_fc.abs().max()) <= 2e-4 * 1.001

    # ---- an eval step afterwards returns the plain mean loss (no stale denominator override) -----------
    l_eval, _, _ = cls.step(cb, train=False)
    l_eval_ref, _ = port.classifier_step(batch, train=False)
    assert abs(float(l_eval) - float(l_eval_ref)) <= 5 * H.RTOL * abs(float(l_eval_ref))


def test_frozen_parameters_gives_the_same_input_gradient_and_leaves_the_arena_alone():
    """``frozen_parameters(net)``: ganffn_net_bwd with grads = NULL -- dx identical to the full backward pass, no
    parameter gradient written."""
    import gan_ffn_b200 as GB
    from gan_ffn_b200 import functional as GF
    torch.manual_seed(5)
    for net, width in ((GB.TextDiscriminator(100), 100), (GB.VisualDiscriminator(100), 512), (GB.AcousticGenerator(100), 100)):
        net = net.cuda().train()
        GB.manual_seed(123)
        x1 = torch.randn(21, 3, width, device="cuda", requires_grad=True)
        y1 = net(x1)
        g = torch.randn_like(y1)
        y1.backward(g)
        ref_dx = x1.grad.detach().clone()
        assert float(net.arena().grad.abs().sum()) > 0
        net.arena().grad.zero_()
        GB.manual_seed(123)                      # the same dropout masks
        x2 = x1.detach().clone().requires_grad_(True)
        with GF.frozen_parameters(net):
            y2 = net(x2)
        assert torch.equal(y1.detach(), y2.detach())
        y2.backward(g)
        assert torch.equal(x2.grad, ref_dx), "data gradient must not depend on whether parameter gradients are computed"
        assert float(net.arena().grad.abs().sum()) == 0.0, "a frozen network's gradient arena must stay untouched"
        x3 = torch.randn(21, 3, width, device="cuda")     # nothing requires grad: backward is never called
        with GF.frozen_parameters(net):
            assert not net(x3).requires_grad


def test_gan_batch_graph_replay_matches_reference_port():
    """The same comparison through ``GraphedTrainStep`` (eager call, recorded call, replayed call = three batches)."""
    from gan_ffn_b200 import synthetic, train
    nets, ffn, gan, cls, pnets, pffn, port = _build_pair()
    batch = synthetic.make_batch(n_dialogues=3, lengths=[12, 7, 10], seed=5)
    cb = batch.to("cuda")
    stepper = train.GraphedTrainStep(gan, cls, seed=1)
    before_o = {k: _flat(m) for k, m in nets.items()}
    before_p = {k: _flat(m) for k, m in pnets.items()}
    for it in range(3):
        out = stepper(cb)
        ref = port.gan_batch(batch)
        loss_ref, _ = port.classifier_step(batch, train=True)
        tol = H.RTOL * (1 + 4 * it)        # later batches start from weights that differ by a few % of lr
        for k in LOSS_KEYS:
            assert abs(float(out[k]) - float(ref[k])) <= tol * abs(float(ref[k])), (it, k, float(out[k]), float(ref[k]))
        assert abs(float(out["loss"]) - float(loss_ref)) <= 2 * tol * abs(float(loss_ref)), (it, float(out["loss"]), float(loss_ref))
    assert stepper.kernels_per_replay, "the third call must have been a graph replay"
    torch.cuda.synchronize()
    for k in nets:
        steps = 3 * (STEPS[k] + (1 if k.endswith("gen") else 0))
        du, dp = _flat(nets[k]) - before_o[k], _flat(pnets[k]) - before_p[k]
        err = (du - dp).abs()
        _log(f"PARITY trainer-graph3 {k} mean_err/lr={float(err.mean()) / LR[k]:.4f} max_err/lr={float(err.max()) / LR[k]:.3f}")
        assert float(err.mean()) < 0.03 * LR[k] and float(err.max()) <= steps * 2.0 * LR[k] * 1.001


def _train_two_steps(deterministic):
    """Two whole train steps (stage 1 + stage 2, TRAIN mode, dropout on, fixed dropout seeds) from the default init."""
    import gan_ffn_b200 as GB
    from gan_ffn_b200 import synthetic, train
    prev = GB.set_deterministic(deterministic)
    try:
        nets, ffn = train.build_networks(device="cuda")
        gan = train.GANTrainer(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], nets["acoustic_disc"],
                               nets["visual_disc"], nets["text_disc"])
        cls = train.ClassifierTrainer(ffn, torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device="cuda"))
        GB.manual_seed(77)
        cb = synthetic.make_batch(n_dialogues=8, seq_len=40, seed=3).to("cuda")
        for _ in range(2):
            gan.batch(cb)
            cls.step(cb, train=True)
        torch.cuda.synchronize()
        return torch.cat([_flat(m) for m in nets.values()] + [ffn.fc.weight.detach().reshape(-1).double().cpu()])
    finally:
        GB.set_deterministic(prev)


def test_deterministic_switch_gives_bit_identical_runs():
    """``set_deterministic(True)``: fixed-order split-K folds and single-writer bias / LayerNorm gradients instead of
    red.global.add (the reference pins determinism, train_IEMOCAP.py:46-53) -- two runs are bit-identical.  The
    default (atomic) mode is only required to stay within a few lr of it."""
    a = _train_two_steps(True)
    b = _train_two_steps(True)
    assert torch.equal(a, b), f"deterministic runs differ in {int((a != b).sum())} of {a.numel()} weights"
    c = _train_two_steps(False)
    _log(f"PARITY deterministic-vs-atomic max|dw|={float((a - c).abs().max()):.3e} mean|dw|={float((a - c).abs().mean()):.3e}")
    assert float((a - c).abs().mean()) < 0.05 * 1e-4
