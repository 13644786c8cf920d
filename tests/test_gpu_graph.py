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
    for i, (a, b) in enumerate(zip(results["eager"][0], results["graph"][0])):
        for k in a:
            if k in ("pred", "labels"):
                continue
            H.assert_close(b[k].cpu(), a[k].cpu(), f"step {i} {k}", rtol=1e-4, atol_frac=1e-5)
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
