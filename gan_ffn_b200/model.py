"""Host-side mirror of the reference's hot-path ``nn.Module`` surface.

Same class names, constructor signatures, attribute names (hence ``state_dict`` keys and
default initialisation under a fixed seed) and ``forward`` signatures as
``/root/reference/model.py:1178-1528``; the arithmetic is done by the sm_100a kernels in
``libganffn.so``.  The ``nn.TransformerEncoderLayer`` / ``nn.Linear`` sub-modules are kept
purely as *parameter containers*: their ``forward`` is never called.

Quirks kept on purpose: the prototype ``encoder_layer`` stays registered as
a ninth, never-used layer; ``GAN_FFN.lstm`` / ``smax_fc`` are constructed and unused;
padded slots are computed as real tokens (no key-padding mask).
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn as nn

from . import functional as GF

GENERATOR, DISCRIMINATOR = 0, 1


class PositionalEncoding(nn.Module):
    """reference model.py:1178-1197.  ``pe`` is a registered buffer exactly as there; the
    add and the Dropout(0.2) happen inside the network kernel sequence."""

    def __init__(self, d_model: int, dropout: float = 0.2, max_len: int = 110):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, 1, d_model)
        pe[:, 0, 0::2] = torch.sin(position * div_term)
        pe[:, 0, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


def _encoder(d_model: int, nhead: int):
    layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # "enable_nested_tensor ..." (same warning the reference prints)
        enc = nn.TransformerEncoder(encoder_layer=layer, num_layers=8)
    return layer, enc


_LAYER_KEYS = ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
               "self_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
               "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")


class _FusedNet(nn.Module):
    """Common machinery of the six networks: the flat parameter arena and the call into the
    whole-network kernel sequence."""

    _kind = GENERATOR

    def _spec(self) -> GF.NetSpec:
        lyr = self.transformer_encoder.layers[0]
        h2 = self.fc2.out_features
        return GF.NetSpec(self._kind, lyr.self_attn.embed_dim, lyr.self_attn.num_heads, lyr.linear1.out_features,
                          len(self.transformer_encoder.layers), self.fc1.out_features, h2)

    def _table(self):
        """Parameters in the canonical order of include/ganffn.h (ganffn_net_fwd)."""
        tab = []
        for lyr in self.transformer_encoder.layers:
            named = dict(lyr.named_parameters())
            tab += [named[k] for k in _LAYER_KEYS]
        tab += [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias]
        if self._kind == DISCRIMINATOR:
            tab += [self.fc3.weight, self.fc3.bias]
            obj = getattr(self, "object", None)
            tab += [obj.weight, obj.bias] if obj is not None else [None, None]
        return tab

    def arena(self) -> GF.ParamArena:
        """The flat parameter/gradient arena (built on first use on the module's device)."""
        ar = self.__dict__.get("_arena")
        dev = self.fc1.weight.device
        if ar is None or not ar.valid_for(dev):
            if dev.type != "cuda":
                raise RuntimeError(
                    f"{type(self).__name__} is on {dev}: gan_ffn_b200 computes only with its sm_100a CUDA kernels; "
                    "move the module to a CUDA device (there is no CPU fallback).")
            table = self._table()
            live = [p for p in table if p is not None]
            ar = GF.ParamArena(live, table)
            self.__dict__["_arena"] = ar
            self.__dict__["_spec_cache"] = self._spec()
        return ar

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_arena", None)  # .to()/.cuda() re-create parameter storage
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):
        st = self.__dict__.copy()
        st.pop("_arena", None)
        st.pop("_spec_cache", None)
        return st

    def _run(self, x: torch.Tensor) -> torch.Tensor:
        ar = self.arena()
        return GF.net_forward(x, ar, self.__dict__["_spec_cache"], self.position_encoding.pe, self.training,
                              self.dropout.p)


class _Generator(_FusedNet):
    _kind = GENERATOR

    def _build(self, d_in: int, nhead: int, hidden: int, D_h: int, dropout: float):
        # construction order = reference order, so default init under a seed is bit-identical
        self.position_encoding = PositionalEncoding(d_in)
        self.encoder_layer, self.transformer_encoder = _encoder(d_in, nhead)
        self.fc1 = nn.Linear(d_in, hidden)
        self.fc2 = nn.Linear(hidden, D_h)
        self.gelu = nn.GELU()
        self.dropout = nn.Dropout(dropout)


class AcousticGenerator(_Generator):
    """acoustic (seq_len, batch, 100) -> fusion (seq_len, batch, D_h).  reference model.py:1200-1231."""

    def __init__(self, D_h, dropout=0.2):
        super(AcousticGenerator, self).__init__()
        self._build(100, 10, 512, D_h, dropout)

    def forward(self, acoustic):
        return self._run(acoustic)


class VisualGenerator(_Generator):
    """visual (seq_len, batch, 512) -> fusion (seq_len, batch, D_h).  reference model.py:1234-1263."""

    def __init__(self, D_h, dropout=0.2):
        super(VisualGenerator, self).__init__()
        self._build(512, 8, 1024, D_h, dropout)

    def forward(self, acoustic):
        return self._run(acoustic)


class TextGenerator(_Generator):
    """text (seq_len, batch, 100) -> fusion (seq_len, batch, D_h).  reference model.py:1266-1294."""

    def __init__(self, D_h, dropout=0.2):
        super(TextGenerator, self).__init__()
        self._build(100, 10, 512, D_h, dropout)

    def forward(self, acoustic):
        return self._run(acoustic)


class _Discriminator(_FusedNet):
    _kind = DISCRIMINATOR

    def _build(self, D_h: int, dropout: float, with_object: bool):
        self.position_encoding = PositionalEncoding(D_h)
        self.encoder_layer, self.transformer_encoder = _encoder(D_h, 10)
        if with_object:
            self.object = nn.Linear(512, 100)  # real visual features are 512 wide (reference model.py:1344)
        self.fc1 = nn.Linear(D_h, 64)
        self.fc2 = nn.Linear(64, 16)
        self.fc3 = nn.Linear(16, 1)
        self.gelu = nn.GELU()
        self.sigmoid = nn.Sigmoid()
        self.dropout = nn.Dropout(dropout)


    def project_real(self, real: torch.Tensor) -> torch.Tensor:
        """What ``forward`` does to a *real* feature before the positional encoding: the identity, except for the
        visual discriminator whose 512-wide real input goes through ``object`` (model.py:1355-1356).  Used by the
        batched real|fake pass of ``train.train_disc_batched``."""
        obj = getattr(self, "object", None)
        if obj is None or real.size(-1) != obj.in_features:
            return real
        ar = self.arena()
        tab = ar.table
        return GF.arena_linear(real, ar, int(tab[-2]), int(tab[-1]), obj.out_features)


class AcousticDiscriminator(_Discriminator):
    """fusion (seq_len, batch, D_h) -> prob (seq_len, batch, 1).  reference model.py:1297-1327."""

    def __init__(self, D_h, dropout=0.2):
        super(AcousticDiscriminator, self).__init__()
        self._build(D_h, dropout, False)

    def forward(self, acoustic_fusion):
        return self._run(acoustic_fusion)


class VisualDiscriminator(_Discriminator):
    """fusion (seq_len, batch, D_h or 512) -> prob (seq_len, batch, 1).  reference model.py:1330-1364:
    a 512-wide input first goes through ``object`` (512 -> 100)."""

    def __init__(self, D_h, dropout=0.2):
        super(VisualDiscriminator, self).__init__()
        self._build(D_h, dropout, True)

    def forward(self, visual_fusion):
        return self._run(visual_fusion)


class TextDiscriminator(_Discriminator):
    """fusion (seq_len, batch, D_h) -> prob (seq_len, batch, 1).  reference model.py:1367-1397."""

    def __init__(self, D_h, dropout=0.2):
        super(TextDiscriminator, self).__init__()
        self._build(D_h, dropout, False)

    def forward(self, text_fusion):
        return self._run(text_fusion)


class GAN_FFN(nn.Module):
    """reference model.py:1405-1462: fusion = G_a(acoustic) + G_v(visual) + G_t(text), ``fc`` 100 -> n_classes,
    ``log_softmax(dim=2)``.  Returns ``(log_prob, [], [], [])``.  Argument order: acoustic, visual, text."""

    def __init__(self, acoustic_generator, visual_generator, text_generator, n_classes=6, dropout=0.2):
        super(GAN_FFN, self).__init__()
        self.n_classes = n_classes
        self.acoustic_generator = acoustic_generator
        self.visual_generator = visual_generator
        self.text_generator = text_generator
        self.lstm = nn.LSTM(100, n_classes, bidirectional=False)  # constructed, unused (as in the reference)
        self.gelu = nn.GELU()
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.smax_fc = nn.Linear(32 * 2, n_classes)               # constructed, unused
        self.fc = nn.Linear(100, n_classes)

    def forward(self, acoustic, visual, text):
        alpha, alpha_f, alpha_b = [], [], []
        acoustic_fusion = self.acoustic_generator(acoustic)
        visual_fusion = self.visual_generator(visual)
        text_fusion = self.text_generator(text)
        log_prob = GF.fuse_classify(acoustic_fusion, visual_fusion, text_fusion, self.fc.weight, self.fc.bias)
        return log_prob, alpha, alpha_f, alpha_b


class GAN_FFN_DialogueRNN(nn.Module):
    """reference model.py:1465-1528: the same three generators and sum as ``GAN_FFN`` (the part on the sm_100a
    kernels), feeding the DialogueRNN ``BiModel`` head (stock PyTorch, ``dialogue_rnn.py``).  ``fc1`` is constructed
    and unused, as in the reference.  ``forward(acoustic, visual, text, qmask, umask)`` returns
    ``(log_prob, alpha, alpha_f, alpha_b)``."""

    def __init__(self, acoustic_generator, visual_generator, text_generator, D_m, D_g, D_p, D_e, D_h, D_a, n_classes,
                 listener_state, context_attention, dropout_rec, dropout):
        super(GAN_FFN_DialogueRNN, self).__init__()
        from .dialogue_rnn import BiModel
        self.n_classes = n_classes
        self.acoustic_generator = acoustic_generator
        self.visual_generator = visual_generator
        self.text_generator = text_generator
        self.gelu = nn.GELU()
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.bi_model = BiModel(D_m=D_m, D_g=D_g, D_p=D_p, D_e=D_e, D_h=D_h, n_classes=n_classes,
                                listener_state=listener_state, context_attention=context_attention, D_a=D_a,
                                dropout_rec=dropout_rec, dropout=dropout)
        self.fc1 = nn.Linear(100, n_classes)

    def fusion(self, acoustic, visual, text):
        """The fused features the head consumes (model.py:1517-1524)."""
        acoustic_fusion = self.acoustic_generator(acoustic)
        visual_fusion = self.visual_generator(visual)
        text_fusion = self.text_generator(text)
        GF.join_lanes()   # the sum below is a plain torch op on the caller's stream
        return acoustic_fusion + visual_fusion + text_fusion

    def forward(self, acoustic, visual, text, qmask, umask, max_len=None):
        """``max_len`` (optional, beyond the reference's signature): the longest dialogue of the batch as a host integer
        (the loader's ``lengths``); without it the head reads it back from ``umask`` (a device synchronisation)."""
        fusion = self.fusion(acoustic, visual, text)
        log_prob, alpha, alpha_f, alpha_b = self.bi_model(fusion, qmask, umask, max_len=max_len)
        return log_prob, alpha, alpha_f, alpha_b


class MaskedNLLLoss(nn.Module):
    """reference model.py:62-81.  ``den_override`` (> 0) replaces the denominator
    sum(w[target]*mask) -- used under dialogue sharding, where it must be the *global* sum."""

    def __init__(self, weight=None):
        super(MaskedNLLLoss, self).__init__()
        self.weight = weight
        self.den_override = 0.0

    def forward(self, pred, target, mask):
        return GF.masked_nll(pred, target, mask, self.weight, self.den_override)


class BCELoss(nn.Module):
    """``torch.nn.BCELoss()`` as used at reference train_IEMOCAP.py:300 (mean reduction).
    ``scale`` (host float) and ``scale_tensor`` (device scalar, CUDA-graph safe) multiply the mean: under dialogue
    sharding a rank contributes local_mean * B_local / B_global."""

    def __init__(self):
        super().__init__()
        self.scale = 1.0
        self.scale_tensor = None

    def forward(self, input, target):
        loss = GF.bce(input, target, self.scale)
        return loss if self.scale_tensor is None else loss * self.scale_tensor
