"""CPU port of the reference's hot path on stock ``torch.nn`` modules.  TEST INFRASTRUCTURE ONLY:
used as the timed CPU baseline (``bench.py`` ``cpu_baseline`` leg and ``--impl reference``) and as
a second checker in ``tests/``.  It does what ``/root/reference/model.py:1178-1462`` and the loop
bodies ``train_IEMOCAP.py:127-170, 200-252, 355-382`` do, with the same stock operators
(``nn.TransformerEncoder``, ``nn.Linear``, ``nn.GELU``, ``nn.Dropout``, ``nn.BCELoss``,
``optim.Adam``), so its CPU time is the reference's CPU time; the reference itself is Python that
cannot travel to the GPU box.  One parametrised class replaces the reference's six near-identical
ones; attribute names (hence ``state_dict`` keys) are the reference's.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F


class _PE(nn.Module):  # model.py:1178-1197
    def __init__(self, d_model, dropout=0.2, max_len=110):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        pos = torch.arange(max_len).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, 1, d_model)
        pe[:, 0, 0::2] = torch.sin(pos * div)
        pe[:, 0, 1::2] = torch.cos(pos * div)
        self.register_buffer("pe", pe)

    def forward(self, x):
        return self.dropout(x + self.pe[: x.size(0)])


class PortNet(nn.Module):
    """kind 'gen': model.py:1200-1294; kind 'disc': model.py:1297-1397."""

    def __init__(self, kind, d_model, nhead, widths, dropout=0.2, with_object=False):
        super().__init__()
        self.kind = kind
        self.position_encoding = _PE(d_model)
        self.encoder_layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.transformer_encoder = nn.TransformerEncoder(encoder_layer=self.encoder_layer, num_layers=8)
        if with_object:
            self.object = nn.Linear(512, 100)
        dims = [d_model] + list(widths)
        for i in range(len(widths)):
            setattr(self, f"fc{i + 1}", nn.Linear(dims[i], dims[i + 1]))
        self.n_fc = len(widths)
        self.gelu = nn.GELU()
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        if self.kind == "disc" and x.size(-1) == 512:
            x = self.object(x)
        y = self.gelu(self.transformer_encoder(self.position_encoding(x)))
        if self.kind == "gen":
            y = self.dropout(y)
            y = self.gelu(self.dropout(self.fc1(y)))
            return self.gelu(self.dropout(self.fc2(y)))
        y = self.gelu(self.dropout(self.fc1(y)))
        y = self.gelu(self.dropout(self.fc2(y)))
        return torch.sigmoid(self.dropout(self.fc3(y)))


class PortGANFFN(nn.Module):  # model.py:1405-1462
    def __init__(self, ga, gv, gt, n_classes=6, dropout=0.2):
        super().__init__()
        self.acoustic_generator, self.visual_generator, self.text_generator = ga, gv, gt
        self.lstm = nn.LSTM(100, n_classes, bidirectional=False)
        self.dropout = nn.Dropout(dropout)
        self.smax_fc = nn.Linear(64, n_classes)
        self.fc = nn.Linear(100, n_classes)

    def forward(self, acoustic, visual, text):
        fusion = self.acoustic_generator(acoustic) + self.visual_generator(visual) + self.text_generator(text)
        return F.log_softmax(self.fc(fusion), 2), [], [], []


class PortMaskedNLLLoss(nn.Module):  # model.py:62-81
    def __init__(self, weight=None):
        super().__init__()
        self.weight = weight
        self.loss = nn.NLLLoss(weight=weight, reduction="sum")

    def forward(self, pred, target, mask):
        mask_ = mask.view(-1, 1)
        if self.weight is None:
            return self.loss(pred * mask_, target) / torch.sum(mask)
        return self.loss(pred * mask_, target) / torch.sum(self.weight[target] * mask_.squeeze())


def build(D_h=100, n_classes=6, seed=3407):
    """Same construction order as gan_ffn_b200.train.build_networks / train_IEMOCAP.py:580-585."""
    torch.manual_seed(seed)
    nets = dict(
        acoustic_gen=PortNet("gen", 100, 10, (512, D_h)), acoustic_disc=PortNet("disc", D_h, 10, (64, 16, 1)),
        visual_gen=PortNet("gen", 512, 8, (1024, D_h)), visual_disc=PortNet("disc", D_h, 10, (64, 16, 1), with_object=True),
        text_gen=PortNet("gen", 100, 10, (512, D_h)), text_disc=PortNet("disc", D_h, 10, (64, 16, 1)))
    ffn = PortGANFFN(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], n_classes, dropout=0.6)
    return nets, ffn


def _train_disc(disc, real_d, gen, real_g, opt, adv, valid, fake):  # train_IEMOCAP.py:200-227
    disc.train(); gen.eval()
    opt.zero_grad()
    d_loss = (adv(disc(real_d), valid) + adv(disc(gen(real_g).detach()), fake)) / 2.0
    d_loss.backward()
    opt.step()
    return d_loss.detach()


def _train_gen(gen, real_g, disc, opt, adv, valid, fake):  # train_IEMOCAP.py:230-252
    gen.train(); disc.eval()
    opt.zero_grad()
    g_loss = adv(disc(gen(real_g)), valid)
    g_loss.backward()
    opt.step()
    return g_loss.detach()


class PortTrainer:
    """Stage 1 (train_IEMOCAP.py:255-393) + stage 2 (:103-197) on the CPU."""

    def __init__(self, nets, ffn, loss_weights=None, lr=1e-4, b1=0.5, b2=0.6, ffn_lr=1e-4, l2=0.008):
        self.nets, self.ffn = nets, ffn
        A = torch.optim.Adam
        self.opts = dict(
            acoustic_gen=A(nets["acoustic_gen"].parameters(), lr=lr, betas=(b1, b2)),
            acoustic_disc=A(nets["acoustic_disc"].parameters(), lr=lr / 2, betas=(b1, b2)),
            visual_gen=A(nets["visual_gen"].parameters(), lr=lr, betas=(b1, b2)),
            visual_disc=A(nets["visual_disc"].parameters(), lr=lr / 2, betas=(b1, b2)),
            text_gen=A(nets["text_gen"].parameters(), lr=lr * 1.1, betas=(b1, b2)),
            text_disc=A(nets["text_disc"].parameters(), lr=lr / 2, betas=(b1, b2)))
        self.adv = nn.BCELoss()
        self.loss_function = PortMaskedNLLLoss(loss_weights)
        self.optimizer = A(ffn.parameters(), lr=ffn_lr, weight_decay=l2)

    def gan_batch(self, b):
        n, o, adv = self.nets, self.opts, self.adv
        S, B = b.text.size(0), b.text.size(1)
        dev = b.text.device          # CPU for the baseline legs; bench.py's gpu_eager_baseline leg runs the same port on the GPU
        valid, fake = torch.ones(S, B, 1, device=dev), torch.zeros(S, B, 1, device=dev)
        t, v, a = b.text, b.visual, b.acoustic
        L = {}
        L["visual_D_loss"] = _train_disc(n["visual_disc"], v, n["acoustic_gen"], a, o["visual_disc"], adv, valid, fake)
        L["acoustic_G_loss"] = _train_gen(n["acoustic_gen"], a, n["visual_disc"], o["acoustic_gen"], adv, valid, fake)
        L["visual_D_loss"] = _train_disc(n["visual_disc"], v, n["text_gen"], t, o["visual_disc"], adv, valid, fake)
        L["text_G_loss"] = _train_gen(n["text_gen"], t, n["visual_disc"], o["text_gen"], adv, valid, fake)
        L["text_D_loss"] = _train_disc(n["text_disc"], t, n["acoustic_gen"], a, o["text_disc"], adv, valid, fake)
        L["acoustic_G_loss"] = _train_gen(n["acoustic_gen"], a, n["text_disc"], o["acoustic_gen"], adv, valid, fake)
        L["acoustic_D_loss"] = _train_disc(n["acoustic_disc"], a, n["text_gen"], t, o["acoustic_disc"], adv, valid, fake)
        L["text_G_loss"] = _train_gen(n["text_gen"], t, n["acoustic_disc"], o["text_gen"], adv, valid, fake)
        L["text_D_loss"] = _train_disc(n["text_disc"], t, n["visual_gen"], v, o["text_disc"], adv, valid, fake)
        L["visual_G_loss"] = _train_gen(n["visual_gen"], v, n["text_disc"], o["visual_gen"], adv, valid, fake)
        L["acoustic_D_loss"] = _train_disc(n["acoustic_disc"], a, n["visual_gen"], v, o["acoustic_disc"], adv, valid, fake)
        L["visual_G_loss"] = _train_gen(n["visual_gen"], v, n["acoustic_disc"], o["visual_gen"], adv, valid, fake)
        return L

    def classifier_step(self, b, train=True):
        self.ffn.train() if train else self.ffn.eval()
        if train:
            self.optimizer.zero_grad()
        with torch.set_grad_enabled(train):
            log_prob = self.ffn(b.acoustic, b.visual, b.text)[0]
            lp_ = log_prob.transpose(0, 1).contiguous().view(-1, log_prob.size()[2])
            loss = self.loss_function(lp_, b.label.view(-1), b.umask)
        if train:
            loss.backward()
            self.optimizer.step()
        return loss.detach(), torch.argmax(lp_, 1)
