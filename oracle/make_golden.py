"""Generates tests/golden/*.npz by running the UNMODIFIED reference (`/root/reference/model.py`).

TEST INFRASTRUCTURE.  Runs only in the build container (the reference does not travel to the
GPU box); the fixtures it writes are committed.  Usage:  python oracle/make_golden.py

What a fixture holds (all produced by the reference modules in eval mode with grad enabled,
i.e. the train-path arithmetic minus dropout masks):
  * per-parameter float64 sums of the default initialisation under seed 3407 (pins that our
    constructors draw bit-identical weights),
  * outputs, a scalar loss, the input gradient, and for every parameter gradient its sum, L2 norm
    and 64 probed entries (full gradients of 4-29 M parameters would not be "small fixtures"),
  * the stage-2 path: GAN_FFN log-probabilities, MaskedNLLLoss value and gradients,
  * one torch.optim.Adam step on the acoustic generator (parameter deltas, probed).
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import model as ref  # noqa: E402  (the reference)
from gan_ffn_b200 import synthetic  # noqa: E402

SEED = 3407
NPROBE = 64


def probe_index(n: int) -> np.ndarray:
    return (np.arange(NPROBE, dtype=np.int64) * 2654435761 + 12345) % n


def build_reference_nets():
    """Construction order is part of the fixture: it fixes which RNG draws each net gets."""
    torch.manual_seed(SEED)
    nets = {
        "acoustic_gen": ref.AcousticGenerator(100, dropout=0.2),
        "visual_gen": ref.VisualGenerator(100, dropout=0.2),
        "text_gen": ref.TextGenerator(100, dropout=0.2),
        "acoustic_disc": ref.AcousticDiscriminator(100, dropout=0.2),
        "visual_disc": ref.VisualDiscriminator(100, dropout=0.2),
        "text_disc": ref.TextDiscriminator(100, dropout=0.2),
    }
    ffn = ref.GAN_FFN(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], n_classes=6, dropout=0.6)
    return nets, ffn


def grad_record(module, out: dict, key: str):
    names, sums, norms, probes = [], [], [], []
    for n, p in module.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().double().reshape(-1)
        names.append(n)
        sums.append(g.sum().item())
        norms.append(g.norm().item())
        probes.append(g[torch.from_numpy(probe_index(g.numel()))].numpy())
    out[f"{key}/g_names"] = np.array(names)
    out[f"{key}/g_sum"] = np.array(sums)
    out[f"{key}/g_norm"] = np.array(norms)
    out[f"{key}/g_probe"] = np.stack(probes)


def main():
    nets, ffn = build_reference_nets()
    for m in list(nets.values()) + [ffn]:
        m.eval()
    out = {}
    # ---- initialisation pins --------------------------------------------------------------------
    for k, m in nets.items():
        names = [n for n, _ in m.named_parameters()]
        out[f"{k}/w_names"] = np.array(names)
        out[f"{k}/w_sum"] = np.array([p.detach().double().sum().item() for _, p in m.named_parameters()])
    out["ffn/fc_w_sum"] = np.array([ffn.fc.weight.double().sum().item(), ffn.fc.bias.double().sum().item()])

    batch = synthetic.make_batch(n_dialogues=3, lengths=[12, 7, 10], seed=SEED)
    g = torch.Generator().manual_seed(SEED + 7)
    probe_w = torch.rand(batch.seq_len, 3, 100, generator=g)   # fixed cotangent for generator outputs
    inputs = {"acoustic_gen": batch.acoustic, "visual_gen": batch.visual, "text_gen": batch.text,
              "acoustic_disc": batch.acoustic, "visual_disc": batch.visual, "text_disc": batch.text}

    # ---- every network alone: forward, scalar loss, backward ------------------------------------------
    bce = torch.nn.BCELoss()
    for k, m in nets.items():
        m.zero_grad()
        x = inputs[k].clone().requires_grad_(True)
        y = m(x)
        if k.endswith("gen"):
            loss = (y * probe_w).sum()
        else:
            loss = bce(y, torch.ones_like(y))
        loss.backward()
        out[f"{k}/out"] = y.detach().numpy()
        out[f"{k}/loss"] = np.array(loss.item())
        out[f"{k}/dx"] = x.grad.numpy()
        grad_record(m, out, k)

    # visual discriminator on a 100-wide (generated) input: the `object` projection is skipped
    m = nets["visual_disc"]
    m.zero_grad()
    x = batch.acoustic.clone().requires_grad_(True)
    y = m(x)
    loss = bce(y, torch.zeros_like(y))
    loss.backward()
    out["visual_disc_fake/out"] = y.detach().numpy()
    out["visual_disc_fake/loss"] = np.array(loss.item())
    out["visual_disc_fake/dx"] = x.grad.numpy()
    grad_record(m, out, "visual_disc_fake")

    # ---- stage 2: GAN_FFN + MaskedNLLLoss (train_IEMOCAP.py:151-165) ---------------------------------------
    ffn.zero_grad()
    w = torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS)
    log_prob = ffn(batch.acoustic, batch.visual, batch.text)[0]
    lp_ = log_prob.transpose(0, 1).contiguous().view(-1, log_prob.size()[2])
    loss = ref.MaskedNLLLoss(w)(lp_, batch.label.view(-1), batch.umask)
    loss.backward()
    out["ffn/log_prob"] = log_prob.detach().numpy()
    out["ffn/loss"] = np.array(loss.item())
    grad_record(ffn, out, "ffn")

    # ---- stage 1 sub-steps: train_disc / train_gen losses (train_IEMOCAP.py:200-252), dropout off ------
    d, gen = nets["visual_disc"], nets["acoustic_gen"]
    valid = torch.ones(batch.seq_len, 3, 1)
    fake = torch.zeros(batch.seq_len, 3, 1)
    d.zero_grad(); gen.zero_grad()
    d_loss = (bce(d(batch.visual), valid) + bce(d(gen(batch.acoustic).detach()), fake)) / 2.0
    d_loss.backward()
    out["train_disc/loss"] = np.array(d_loss.item())
    grad_record(d, out, "train_disc")
    d.zero_grad(); gen.zero_grad()
    g_loss = bce(d(gen(batch.acoustic)), valid)
    g_loss.backward()
    out["train_gen/loss"] = np.array(g_loss.item())
    grad_record(gen, out, "train_gen")

    # ---- one Adam step (train_IEMOCAP.py:292: lr 1e-4, betas (0.5, 0.6)) on those generator gradients ----
    before = {n: p.detach().clone() for n, p in gen.named_parameters()}
    opt = torch.optim.Adam(gen.parameters(), lr=1e-4, betas=(0.5, 0.6))
    opt.step()
    names, deltas = [], []
    for n, p in gen.named_parameters():
        if p.grad is None:
            continue
        dlt = (p.detach() - before[n]).double().reshape(-1)
        names.append(n)
        deltas.append(dlt[torch.from_numpy(probe_index(dlt.numel()))].numpy())
    out["adam/names"] = np.array(names)
    out["adam/delta_probe"] = np.stack(deltas)

    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", "ganffn_ref_seed3407.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays; torch", torch.__version__)


if __name__ == "__main__":
    main()
