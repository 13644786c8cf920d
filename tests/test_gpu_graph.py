"""CUDA-graph replay of the train step against the eager path.

With the same DeviceSeedStream state and the same initial weights, a replayed step must produce what the eager
step produces (the kernels are the same; only the launch mechanism differs).  Gradient accumulation uses
red.global.add, so the comparison is to rtol 1e-4 rather than bit-exact.  Also checks that consecutive replays draw
different dropout masks and advance Adam's bias correction (the device-resident step count)."""
import copy

import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


def _setup(seed=1234):
    from gan_ffn_b200 import synthetic, train
    torch.manual_seed(seed)
    nets, ffn = train.build_networks(device="cuda", seed=seed)
    gan = train.GANTrainer(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], nets["acoustic_disc"],
                           nets["visual_disc"], nets["text_disc"])
    cls = train.ClassifierTrainer(ffn, torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device="cuda"))
    return nets, ffn, gan, cls


def _flat(nets, ffn):
    return torch.cat([p.detach().reshape(-1) for m in list(nets.values()) + [ffn] for p in m.parameters()])


def test_graph_replay_matches_eager_steps():
    from gan_ffn_b200 import synthetic, train
    batch = synthetic.make_batch(n_dialogues=3, seq_len=17, seed=5).to("cuda")
    results = {}
    for mode in ("eager", "graph"):
        nets, ffn, gan, cls = _setup()
        stepper = train.GraphedTrainStep(gan, cls, seed=99, enabled=(mode == "graph"))
        outs = []
        for _ in range(4):   # graph mode: eager, capture+replay, replay, replay
            out = stepper(batch)
            outs.append({k: v.detach().clone() for k, v in out.items()})
        torch.cuda.synchronize()
        if mode == "graph":
            assert stepper.kernels_per_replay, "the step was never captured"
        results[mode] = (outs, _flat(nets, ffn))
    # Step 0 starts from identical weights: rtol 1e-4.  Later steps start from weights that already differ by the
    # round-off of the (order-dependent, cross-stream) red.global.add accumulation, which Adam amplifies to +-lr on
    # entries whose gradient is ~0: 2e-3.
    for i, (a, b) in enumerate(zip(results["eager"][0], results["graph"][0])):
        for k in a:
            if k in ("pred", "labels"):
                continue
            H.assert_close(b[k].cpu(), a[k].cpu(), f"step {i} {k}", rtol=1e-4 if i == 0 else 2e-3, atol_frac=1e-5)
    # Weights: Adam turns a gradient of magnitude ~0 into a step of +-lr, so the (order-dependent) round-off of the
    # red.global.add accumulation can move individual entries by up to lr per step.  On average the weights must agree
    # to a small fraction of one step, and no entry may differ by more than the Adam steps taken: a generator is
    # stepped three times per call (twice in stage 1, once in stage 2), i.e. twelve times over the four calls.
    wa, wb = results["eager"][1], results["graph"][1]
    diff = (wa - wb).abs()
    lr_max = 1.1e-4
    assert float(diff.max()) <= 12 * lr_max * 1.01, float(diff.max())
    assert float(diff.mean()) < 0.1 * lr_max, float(diff.mean())


def test_replays_draw_fresh_dropout_masks():
    from gan_ffn_b200 import synthetic, train
    batch = synthetic.make_batch(n_dialogues=2, seq_len=11, seed=7).to("cuda")
    nets, ffn, gan, cls = _setup()
    for opt in (gan.opt_acoustic_G, gan.opt_acoustic_D, gan.opt_visual_G, gan.opt_visual_D, gan.opt_text_G,
                gan.opt_text_D, cls.optimizer):
        opt.param_groups[0]["lr"] = 0.0          # freeze the weights: only the masks can change the losses
    stepper = train.GraphedTrainStep(gan, cls, seed=3)
    losses = []
    for _ in range(4):
        losses.append(float(stepper(batch)["loss"].item()))
    assert stepper.kernels_per_replay
    assert len({round(x, 7) for x in losses[1:]}) == 3, f"replays reused a dropout mask: {losses}"


def _setup_overlap(overlap, seed=1234):
    from gan_ffn_b200 import synthetic, train
    torch.manual_seed(seed)
    nets, ffn = train.build_networks(device="cuda", seed=seed)
    gan = train.GANTrainer(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], nets["acoustic_disc"],
                           nets["visual_disc"], nets["text_disc"], overlap=overlap)
    cls = train.ClassifierTrainer(ffn, torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device="cuda"), overlap=overlap)
    return nets, ffn, gan, cls


@pytest.mark.parametrize("graphed", [False, True])
def test_network_lanes_match_serial_issue(graphed):
    """Independent networks of a sub-step on concurrent streams (functional._Lanes) against the one-stream issue
    order: same seeds, same weights -> same losses (rtol 1e-4: the red.global.add order differs) and weights that
    agree to a fraction of an Adam step."""
    from gan_ffn_b200 import synthetic, train
    batch = synthetic.make_batch(n_dialogues=4, seq_len=23, seed=11).to("cuda")
    results = {}
    for overlap in (False, True):
        nets, ffn, gan, cls = _setup_overlap(overlap)
        stepper = train.GraphedTrainStep(gan, cls, seed=7, enabled=graphed)
        outs = []
        for _ in range(3):
            out = stepper(batch)
            outs.append({k: v.detach().clone() for k, v in out.items()})
        torch.cuda.synchronize()
        results[overlap] = (outs, _flat(nets, ffn))
    for i, (a, b) in enumerate(zip(results[False][0], results[True][0])):
        for k in a:
            if k in ("pred", "labels"):
                continue
            H.assert_close(b[k].cpu(), a[k].cpu(), f"step {i} {k}", rtol=1e-4 if i == 0 else 2e-3, atol_frac=1e-5)
    diff = (results[False][1] - results[True][1]).abs()
    assert float(diff.mean()) < 0.1 * 1.1e-4, float(diff.mean())


def test_network_lanes_gradients_match_serial_issue():
    """Dropout off (eval-mode networks, gradients enabled): the train_disc / train_gen / classifier bodies with the
    networks on concurrent lanes must leave the same gradients in the arenas as the one-stream issue order."""
    import gan_ffn_b200 as G
    from gan_ffn_b200 import functional as GF, synthetic
    batch = synthetic.make_batch(n_dialogues=5, seq_len=31, seed=13).to("cuda")
    w = torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device="cuda")
    grads = {}
    for overlap in (False, True):
        nets, ffn, gan, cls = _setup_overlap(overlap)
        for m in list(nets.values()) + [ffn]:
            m.eval()
        disc, gen = nets["visual_disc"], nets["acoustic_gen"]
        bce = G.BCELoss()
        out = []
        for rep in range(3):     # repeated: a missing cross-stream dependency shows up as run-to-run differences
            with GF.overlap_networks(overlap):
                gan.opt_visual_D.zero_grad()
                real_prob = disc(batch.visual)
                fusion = gen(batch.acoustic)
                fake_prob = disc(fusion.detach())
                d_loss = (bce(real_prob, torch.ones_like(real_prob)) + bce(fake_prob, torch.zeros_like(fake_prob))) / 2.0
                d_loss.backward()
                GF.join_lanes()
                g_disc = disc.arena().grad.clone()
                gan.opt_acoustic_G.zero_grad()
                prob = disc(gen(batch.acoustic))
                g_loss = bce(prob, torch.ones_like(prob))
                g_loss.backward()
                GF.join_lanes()
                g_gen = gen.arena().grad.clone()
                cls.optimizer.zero_grad()
                lp = ffn(batch.acoustic, batch.visual, batch.text)[0]
                lp_ = lp.transpose(0, 1).contiguous().view(-1, 6)
                loss = G.MaskedNLLLoss(w)(lp_, batch.label.view(-1), batch.umask)
                loss.backward()
                GF.join_lanes()
                g_ffn = torch.cat([nets[k].arena().grad for k in ("acoustic_gen", "visual_gen", "text_gen")]).clone()
            torch.cuda.synchronize()
            out.append((d_loss.detach().clone(), g_loss.detach().clone(), loss.detach().clone(), g_disc, g_gen, g_ffn))
        grads[overlap] = out
    names = ["d_loss", "g_loss", "nll", "grad(visual_disc)", "grad(acoustic_gen)", "grad(generators, stage 2)"]
    ref = grads[False][0]
    for overlap in (False, True):
        for rep, got in enumerate(grads[overlap]):
            for nm, a, e in zip(names, got, ref):
                H.assert_close(a.cpu(), e.cpu(), f"overlap={overlap} rep {rep} {nm}", rtol=1e-4,
                               atol_frac=1e-4 if nm.startswith("grad") else 1e-6)
