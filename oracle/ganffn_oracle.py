"""CPU oracle for the GAN-FFN fusion hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain restatement, in explicit torch-on-CPU arithmetic, of what
the reference computes on the hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product path (``gan_ffn_b200``) never does.

Where the arithmetic lives.  The reference (``/root/reference/model.py``) builds
every network out of stock ``torch.nn`` modules, so the arithmetic is defined by
the third-party dependency ``torch`` (unpinned in ``requirements.txt:2``; the
README installs 1.11.0+cu113; this image carries 2.11.0+cu128, which is what
the oracle is pinned against).  The published algorithm restated below is
``torch/nn/modules/transformer.py:944-982`` (post-norm encoder layer, ReLU FFN),
``torch/nn/functional.py`` multi_head_attention_forward (packed in-proj,
softmax(QK^T/sqrt(hd)) V, out-proj), LayerNorm eps 1e-5, exact-erf GELU,
``BCELoss`` (log clamped at -100) and ``optim.Adam`` (L2 folded into the grad).

Pinning.  The reference ships no tests or golden vectors, so the
oracle is pinned against *outputs of the reference itself*: ``make_golden.py``
imports the unmodified ``/root/reference/model.py`` in the build container and
writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file
against those fixtures, and (when ``/root/reference`` is present)
``tests/test_oracle_vs_reference.py`` checks it against the live modules.

Parameters are passed as a dict keyed by the reference's ``state_dict`` names so
the same dict can come from a reference module or from ours.

Dropout.  ``masks`` is ``None`` (eval: identity) or a callable
``masks(site:int, shape:tuple) -> tensor`` returning the *scaled* keep mask
(0 or 1/(1-p)) for that site.  Site numbering is shared with the CUDA kernels
(``include/ganffn.h``: GANFFN_SITE_*), so a test can export the masks the
kernels drew and inject them here for an exact train-mode comparison.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch

Tensor = torch.Tensor
MaskFn = Optional[Callable[[int, tuple], Tensor]]

# ---- dropout site ids (mirrors include/ganffn.h) --------------------------------
SITE_PE = 0


def site_layer(layer: int, k: int) -> int:
    """k: 0 attention probabilities, 1 after out-proj, 2 FFN hidden, 3 after linear2."""
    return 16 * (layer + 1) + k


SITE_HEAD = 200  # + 0 after gelu(encoder out) [generators], +1 after fc1, +2 after fc2, +3 after fc3

# ---- hyper-parameters fixed by the reference constructors ----------------------
NLAYERS = 8          # model.py:1211-1213 (num_layers=8) and the five siblings
DFF = 2048           # torch default dim_feedforward
LN_EPS = 1e-5        # torch default layer_norm_eps
P_ENC = 0.1          # torch default TransformerEncoderLayer dropout
P_PE = 0.2           # model.py:1179 PositionalEncoding(dropout=0.2)
MAX_LEN = 110        # model.py:1179


def _drop(x: Tensor, masks: MaskFn, site: int) -> Tensor:
    if masks is None:
        return x
    return x * masks(site, tuple(x.shape)).to(x.dtype)


def gelu(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (model.py:1218)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def positional_table(d_model: int, max_len: int = MAX_LEN, dtype=torch.float32) -> Tensor:
    """model.py:1182-1188.  Returns (max_len, d_model).  The table is built in
    fp32 exactly as the reference does and then cast, so fp64 runs of the oracle
    see the same (fp32-rounded) constants the reference adds."""
    pos = torch.arange(max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.to(dtype)


def positional_encoding(x: Tensor, masks: MaskFn = None) -> Tensor:
    """model.py:1191-1197: x + pe[:S] then Dropout(0.2).  x is (S,B,d)."""
    S, _, d = x.shape
    if S > MAX_LEN:
        raise ValueError(f"seq_len {S} > max_len {MAX_LEN} (model.py:1179)")
    pe = positional_table(d, dtype=x.dtype)[:S].unsqueeze(1)
    return _drop(x + pe, masks, SITE_PE)


def layer_norm(z: Tensor, w: Tensor, b: Tensor) -> Tensor:
    mu = z.mean(-1, keepdim=True)
    var = ((z - mu) ** 2).mean(-1, keepdim=True)  # biased, as torch
    return (z - mu) * torch.rsqrt(var + LN_EPS) * w + b


def attention(x: Tensor, P: Dict[str, Tensor], pre: str, nhead: int, masks: MaskFn, layer: int) -> Tensor:
    """nn.MultiheadAttention self-attention, batch_first=False, no masks
    (functional.py multi_head_attention_forward).  x: (S,B,d) -> (S,B,d)."""
    S, B, d = x.shape
    hd = d // nhead
    qkv = x @ P[pre + "self_attn.in_proj_weight"].T + P[pre + "self_attn.in_proj_bias"]
    q, k, v = qkv.split(d, dim=-1)

    def heads(t):  # (S,B,d) -> (B,H,S,hd)
        return t.reshape(S, B, nhead, hd).permute(1, 2, 0, 3)

    q, k, v = heads(q), heads(k), heads(v)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))
    p = torch.softmax(s, dim=-1)
    p = _drop(p, masks, site_layer(layer, 0))          # (B,H,S,S)
    o = (p @ v).permute(2, 0, 1, 3).reshape(S, B, d)
    return o @ P[pre + "self_attn.out_proj.weight"].T + P[pre + "self_attn.out_proj.bias"]


def encoder_layer(x: Tensor, P: Dict[str, Tensor], layer: int, nhead: int, masks: MaskFn = None,
                  prefix: str = "transformer_encoder.layers.") -> Tensor:
    """torch/nn/modules/transformer.py:944-982 with norm_first=False, activation=relu."""
    pre = f"{prefix}{layer}."
    sa = _drop(attention(x, P, pre, nhead, masks, layer), masks, site_layer(layer, 1))
    x = layer_norm(x + sa, P[pre + "norm1.weight"], P[pre + "norm1.bias"])
    h = torch.relu(x @ P[pre + "linear1.weight"].T + P[pre + "linear1.bias"])
    h = _drop(h, masks, site_layer(layer, 2))
    ff = _drop(h @ P[pre + "linear2.weight"].T + P[pre + "linear2.bias"], masks, site_layer(layer, 3))
    return layer_norm(x + ff, P[pre + "norm2.weight"], P[pre + "norm2.bias"])


def encoder(x: Tensor, P: Dict[str, Tensor], nhead: int, masks: MaskFn = None, nlayers: int = NLAYERS) -> Tensor:
    for l in range(nlayers):
        x = encoder_layer(x, P, l, nhead, masks)
    return x


def generator(x: Tensor, P: Dict[str, Tensor], nhead: int, masks: MaskFn = None) -> Tensor:
    """AcousticGenerator/TextGenerator (nhead=10) model.py:1221-1231, 1286-1294;
    VisualGenerator (nhead=8) model.py:1255-1263.  Dropout comes *before* GELU
    after fc1/fc2 (model.py:1227-1228)."""
    y = gelu(encoder(positional_encoding(x, masks), P, nhead, masks))
    y = _drop(y, masks, SITE_HEAD + 0)
    y = gelu(_drop(y @ P["fc1.weight"].T + P["fc1.bias"], masks, SITE_HEAD + 1))
    y = gelu(_drop(y @ P["fc2.weight"].T + P["fc2.bias"], masks, SITE_HEAD + 2))
    return y


def discriminator(x: Tensor, P: Dict[str, Tensor], masks: MaskFn = None, nhead: int = 10) -> Tensor:
    """Acoustic/Text discriminator model.py:1320-1327, 1390-1397; Visual
    discriminator model.py:1354-1364 (``object`` 512->100 only when the input is
    512 wide, i.e. real visual features).  Returns probabilities (S,B,1)."""
    if x.shape[-1] == 512:
        x = x @ P["object.weight"].T + P["object.bias"]
    y = gelu(encoder(positional_encoding(x, masks), P, nhead, masks))
    y = gelu(_drop(y @ P["fc1.weight"].T + P["fc1.bias"], masks, SITE_HEAD + 1))
    y = gelu(_drop(y @ P["fc2.weight"].T + P["fc2.bias"], masks, SITE_HEAD + 2))
    y = torch.sigmoid(_drop(y @ P["fc3.weight"].T + P["fc3.bias"], masks, SITE_HEAD + 3))
    return y


def gan_ffn(acoustic: Tensor, visual: Tensor, text: Tensor, Pa, Pv, Pt, fc_w: Tensor, fc_b: Tensor,
            masks_a: MaskFn = None, masks_v: MaskFn = None, masks_t: MaskFn = None) -> Tensor:
    """GAN_FFN.forward model.py:1434-1462.  Argument order acoustic, visual, text."""
    fusion = generator(acoustic, Pa, 10, masks_a) + generator(visual, Pv, 8, masks_v) + generator(text, Pt, 10, masks_t)
    return torch.log_softmax(fusion @ fc_w.T + fc_b, dim=2)


def fusion_sum(acoustic, visual, text, Pa, Pv, Pt) -> Tensor:
    """GAN_FFN_DialogueRNN.forward model.py:1517-1524 (the part before BiModel)."""
    return generator(acoustic, Pa, 10) + generator(visual, Pv, 8) + generator(text, Pt, 10)


def masked_nll(pred: Tensor, target: Tensor, mask: Tensor, weight: Optional[Tensor] = None) -> Tensor:
    """MaskedNLLLoss.forward model.py:68-81.  pred (B*S,C) log-probs batch-major,
    target (B*S,) int64, mask (B,S)."""
    m = mask.reshape(-1).to(pred.dtype)
    picked = pred.gather(1, target.view(-1, 1)).squeeze(1)
    if weight is None:
        return -(picked * m).sum() / m.sum()
    w = weight.to(pred.dtype)[target]
    return -(w * picked * m).sum() / (w * m).sum()


def bce(prob: Tensor, target: Tensor) -> Tensor:
    """torch.nn.BCELoss() (train_IEMOCAP.py:300): mean over every element,
    log terms clamped at -100."""
    lp = torch.clamp(torch.log(prob), min=-100.0)
    l1p = torch.clamp(torch.log(1.0 - prob), min=-100.0)
    return -(target * lp + (1.0 - target) * l1p).mean()


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, b1: float, b2: float,
              eps: float = 1e-8, weight_decay: float = 0.0) -> None:
    """torch.optim.Adam single-tensor update, in place (train_IEMOCAP.py:292-297, :661).
    ``step`` is the 1-based step count after the increment."""
    if weight_decay != 0.0:
        g = g + weight_decay * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# ---- convenience: parameter dicts and autograd -----------------------------------
def params_of(module: torch.nn.Module, dtype=torch.float32, requires_grad: bool = False) -> Dict[str, Tensor]:
    """Detached CPU copy of a module's parameters keyed by state_dict name.  Strips a
    leading ``module.`` (nn.DataParallel, train_IEMOCAP.py:587-593)."""
    out = {}
    for k, v in module.state_dict().items():
        k = k[len("module."):] if k.startswith("module.") else k
        t = v.detach().to("cpu", dtype).clone()
        if requires_grad and t.is_floating_point():
            t.requires_grad_(True)
        out[k] = t
    return out
