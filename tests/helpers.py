"""Shared test helpers: build our networks the way the golden fixture built the reference's,
run the CPU oracle with autograd, compare against fixture records."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gan_ffn_b200 import model as M          # noqa: E402
from gan_ffn_b200 import synthetic           # noqa: E402
from oracle import ganffn_oracle as O        # noqa: E402

SEED = 3407
NPROBE = 64
GOLDEN = os.path.join(ROOT, "tests", "golden", "ganffn_ref_seed3407.npz")
NET_ORDER = ["acoustic_gen", "visual_gen", "text_gen", "acoustic_disc", "visual_disc", "text_disc"]
NHEAD = {"acoustic_gen": 10, "visual_gen": 8, "text_gen": 10}

# fp32 tolerance of north_star: rtol 1e-4.  Element-wise on outputs / losses; gradients are compared
# element-wise with rtol 1e-4 plus an absolute floor of 1e-4 x the tensor's largest magnitude (entries
# that are tiny next to their tensor carry no more than fp32 round-off of the large ones).
RTOL = 1e-4


def golden():
    return np.load(GOLDEN, allow_pickle=False)


def probe_index(n):
    return (np.arange(NPROBE, dtype=np.int64) * 2654435761 + 12345) % n


def build_nets(device="cpu"):
    """Same construction order and seed as oracle/make_golden.py:build_reference_nets."""
    torch.manual_seed(SEED)
    nets = {
        "acoustic_gen": M.AcousticGenerator(100, dropout=0.2),
        "visual_gen": M.VisualGenerator(100, dropout=0.2),
        "text_gen": M.TextGenerator(100, dropout=0.2),
        "acoustic_disc": M.AcousticDiscriminator(100, dropout=0.2),
        "visual_disc": M.VisualDiscriminator(100, dropout=0.2),
        "text_disc": M.TextDiscriminator(100, dropout=0.2),
    }
    ffn = M.GAN_FFN(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], n_classes=6, dropout=0.6)
    if device != "cpu":
        for m in nets.values():
            m.to(device)
        ffn.to(device)
    return nets, ffn


def golden_batch():
    return synthetic.make_batch(n_dialogues=3, lengths=[12, 7, 10], seed=SEED)


def golden_cotangent(S):
    g = torch.Generator().manual_seed(SEED + 7)
    return torch.rand(S, 3, 100, generator=g)


def net_inputs(batch):
    return {"acoustic_gen": batch.acoustic, "visual_gen": batch.visual, "text_gen": batch.text,
            "acoustic_disc": batch.acoustic, "visual_disc": batch.visual, "text_disc": batch.text}


def oracle_forward(name, x, P, masks=None):
    if name.endswith("gen"):
        return O.generator(x, P, NHEAD[name], masks)
    return O.discriminator(x, P, masks)


def assert_close(actual, expected, what, rtol=RTOL, atol_frac=0.0):
    a = np.asarray(actual, dtype=np.float64)
    e = np.asarray(expected, dtype=np.float64)
    assert a.shape == e.shape, f"{what}: shape {a.shape} vs {e.shape}"
    scale = float(np.abs(e).max()) if e.size else 0.0
    tol = rtol * np.abs(e) + atol_frac * scale
    err = np.abs(a - e)
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{e.size} entries off; worst |err|={err.max():.3e} "
                           f"(max|ref|={scale:.3e}, worst err/ref-scale={err.max() / max(scale, 1e-30):.3e})")


def check_grads(named_grads, G, key, rtol=RTOL):
    """named_grads: dict name -> tensor; G: golden npz; key: record prefix."""
    names = [str(n) for n in G[f"{key}/g_names"]]
    assert set(names) == set(named_grads), (sorted(set(names) ^ set(named_grads)))
    for i, n in enumerate(names):
        g = named_grads[n].detach().double().cpu().reshape(-1).numpy()
        norm = float(G[f"{key}/g_norm"][i])
        assert abs(np.linalg.norm(g) - norm) <= rtol * norm + 1e-12, f"{key} grad norm {n}: {np.linalg.norm(g)} vs {norm}"
        probe = g[probe_index(g.size)]
        ref = G[f"{key}/g_probe"][i]
        scale = max(float(np.abs(g).max()), 1e-30)
        err = np.abs(probe - ref)
        assert (err <= rtol * np.abs(ref) + rtol * scale).all(), \
            f"{key} grad probe {n}: worst {err.max():.3e} at scale {scale:.3e}"
        # the sum is a cancellation-prone statistic: bound it by rtol x sum|g|
        assert abs(g.sum() - float(G[f"{key}/g_sum"][i])) <= rtol * np.abs(g).sum() + 1e-12, f"{key} grad sum {n}"


def check_adam_delta(delta, ref, grad, lr, what):
    """The first Adam step is -lr * g/(|g| + eps): well conditioned only where |g| >> eps = 1e-8 and above
    the gradient's own round-off.  Entries whose gradient is round-off noise (e.g. the key bias of a softmax
    attention, whose true gradient is exactly zero) may take any value in [-lr, lr]."""
    g = grad.detach().double().cpu().reshape(-1).numpy()
    gp = np.abs(g[probe_index(g.size)])
    solid = gp > 1e-4 * max(np.abs(g).max(), 1e-30)
    err = np.abs(np.asarray(delta) - np.asarray(ref))
    assert (err[solid] <= 1e-3 * lr).all(), f"adam {what}: worst {err[solid].max():.3e} vs lr {lr}"
    assert (err[~solid] <= 2.0 * lr * 1.001).all(), f"adam {what} (noise-level gradients)"
