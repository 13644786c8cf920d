"""Generates tests/golden/dialogue_rnn_ref_seed3407.npz by running the UNMODIFIED reference
(`/root/reference/model.py`: BiModel and GAN_FFN_DialogueRNN, model.py:981-1062 and :1465-1528).

TEST INFRASTRUCTURE.  Runs only in the build container; the fixture it writes is committed.
Usage:  python oracle/make_golden_dialogue_rnn.py

Weights are not stored: both the reference and gan_ffn_b200 draw their default initialisation from
torch.manual_seed(SEED) in the same construction order (pinned bit-exactly by the per-parameter sums kept here).
Kept: for each head configuration the log-probabilities, the input gradient and per-parameter gradient norms of a
scalar loss, all in eval mode (dropout off); and the log-probabilities of the whole GAN_FFN_DialogueRNN on a
synthetic batch (config 5 of BASELINE.json: fused features feeding the DialogueRNN head)."""
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
DIMS = dict(D_m=100, D_g=500, D_p=500, D_e=100, D_h=100)     # train_IEMOCAP_DialogueRNN.py:635-641
HEAD_CONFIGS = [("general", False), ("simple", False), ("general2", True), ("dot", False), ("concat", False)]


def head_inputs(S=9, B=3):
    b = synthetic.make_batch(n_dialogues=B, lengths=[S, S - 4, S - 2], seed=SEED + 1)
    g = torch.Generator().manual_seed(SEED + 2)
    U = torch.rand(S, B, DIMS["D_m"], generator=g) * b.umask.t().unsqueeze(2)
    return U, b.qmask, b.umask


def main():
    out = {}
    U, qmask, umask = head_inputs()
    for att, listener in HEAD_CONFIGS:
        key = f"head/{att}/{int(listener)}"
        torch.manual_seed(SEED)
        d_m = DIMS["D_m"] if att != "dot" else DIMS["D_g"]
        m = ref.BiModel(d_m, DIMS["D_g"], DIMS["D_p"], DIMS["D_e"], DIMS["D_h"], n_classes=6, listener_state=listener,
                        context_attention=att, D_a=100, dropout_rec=0.1, dropout=0.6).eval()
        Ux = U if att != "dot" else torch.cat([U] * 5, dim=2)
        Ux = Ux.clone().requires_grad_(True)
        lp, alpha, alpha_f, alpha_b = m(Ux, qmask, umask)
        w = torch.linspace(0.5, 1.5, lp.numel()).view_as(lp)
        (lp * w).sum().backward()
        out[key + "/log_prob"] = lp.detach().numpy()
        out[key + "/dU"] = Ux.grad.numpy()
        out[key + "/alpha_last"] = alpha[-1].detach().numpy()
        names = [n for n, p in m.named_parameters() if p.grad is not None]
        out[key + "/g_names"] = np.array(names)
        out[key + "/g_norm"] = np.array([m.get_parameter(n).grad.double().norm().item() for n in names])
        out[key + "/p_sum"] = np.array([p.detach().double().sum().item() for _, p in m.named_parameters()])
    # the whole model of config 5
    torch.manual_seed(SEED)
    ga, gv, gt = ref.AcousticGenerator(100, dropout=0.2), ref.VisualGenerator(100, dropout=0.2), ref.TextGenerator(100, dropout=0.2)
    model = ref.GAN_FFN_DialogueRNN(ga, gv, gt, DIMS["D_m"], DIMS["D_g"], DIMS["D_p"], DIMS["D_e"], DIMS["D_h"], 100, 6, False,
                                    "general", 0.1, 0.6).eval()
    b = synthetic.make_batch(n_dialogues=3, lengths=[12, 7, 10], seed=SEED)
    with torch.no_grad():
        lp = model(b.acoustic, b.visual, b.text, b.qmask, b.umask)[0]
        fusion = ga(b.acoustic) + gv(b.visual) + gt(b.text)
    out["model/log_prob"] = lp.numpy()
    out["model/fusion"] = fusion.numpy()
    out["model/state_keys"] = np.array(list(model.state_dict().keys()))
    path = os.path.join(ROOT, "tests", "golden", "dialogue_rnn_ref_seed3407.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
