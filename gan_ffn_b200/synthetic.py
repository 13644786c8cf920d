"""Synthetic IEMOCAP- / MELD-shaped batches.

The reference's feature pickles are not shipped, so every workload here is synthetic with the
shapes ``dataloader.py`` would produce: ``text (S,B,100)``, ``visual (S,B,512)``,
``acoustic (S,B,100)`` fp32 in [0,1) (the loader min-max normalises each dialogue,
dataloader.py:20-35), zero at padded slots (``pad_sequence``, dataloader.py:57),
``qmask (S,B,2)`` one-hot speakers, ``umask (B,S)`` 1/0, ``label (B,S)`` int64 with 0 at padding.
Everything is generated on the CPU with a seeded ``torch.Generator`` so the same batch can be
rebuilt bit-for-bit anywhere (tests, golden fixtures, bench).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

TEXT_DIM, VISUAL_DIM, ACOUSTIC_DIM = 100, 512, 100   # train_IEMOCAP.py:142-144
IEMOCAP_CLASSES, MELD_CLASSES = 6, 7                  # train_IEMOCAP.py:523, train_MELD.py:139
IEMOCAP_LOSS_WEIGHTS = (1.2, 0.60072, 0.38066, 0.94019, 0.67924, 0.34332)  # train_IEMOCAP.py:653
MAX_LEN = 110                                         # model.py:1179


@dataclass
class Batch:
    text: torch.Tensor      # (S,B,100)
    visual: torch.Tensor    # (S,B,512)
    acoustic: torch.Tensor  # (S,B,100)
    qmask: torch.Tensor     # (S,B,2)
    umask: torch.Tensor     # (B,S)
    label: torch.Tensor     # (B,S) int64
    lengths: List[int]

    @property
    def seq_len(self) -> int:
        return self.text.shape[0]

    @property
    def n_dialogues(self) -> int:
        return self.text.shape[1]

    @property
    def padded_slots(self) -> int:
        return self.seq_len * self.n_dialogues

    @property
    def real_utterances(self) -> int:
        return int(sum(self.lengths))

    def to(self, device, non_blocking: bool = False) -> "Batch":
        f = lambda t: t.to(device, non_blocking=non_blocking)
        return Batch(f(self.text), f(self.visual), f(self.acoustic), f(self.qmask), f(self.umask), f(self.label),
                     self.lengths)

    def pin(self) -> "Batch":
        f = lambda t: t.pin_memory()
        return Batch(f(self.text), f(self.visual), f(self.acoustic), f(self.qmask), f(self.umask), f(self.label),
                     self.lengths)

    def dialogues(self, idx) -> "Batch":
        """Sub-batch of the given dialogue indices, *keeping the global pad length* (a dialogue's
        output depends on how far it is padded)."""
        idx = torch.as_tensor(idx, dtype=torch.long)
        return Batch(self.text[:, idx].contiguous(), self.visual[:, idx].contiguous(),
                     self.acoustic[:, idx].contiguous(), self.qmask[:, idx].contiguous(),
                     self.umask[idx].contiguous(), self.label[idx].contiguous(), [self.lengths[i] for i in idx.tolist()])

    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.text, self.visual, self.acoustic, self.qmask, self.umask, self.label))


def make_batch(n_dialogues: int = 32, seq_len: int = 94, lengths: Optional[List[int]] = None, n_classes: int = 6,
               seed: int = 3407) -> Batch:
    """Full-length batch when ``lengths`` is None (the shape the author logged, model.py:1437);
    otherwise ragged with ``S = max(lengths)``."""
    g = torch.Generator().manual_seed(seed)
    if lengths is None:
        lengths = [seq_len] * n_dialogues
    assert len(lengths) == n_dialogues
    S = max(lengths)
    if S > MAX_LEN:
        raise ValueError(f"dialogue of {S} turns exceeds PositionalEncoding max_len {MAX_LEN} (model.py:1179)")
    B = n_dialogues
    text = torch.rand(S, B, TEXT_DIM, generator=g)
    visual = torch.rand(S, B, VISUAL_DIM, generator=g)
    acoustic = torch.rand(S, B, ACOUSTIC_DIM, generator=g)
    spk = torch.randint(0, 2, (S, B), generator=g)
    label = torch.randint(0, n_classes, (B, S), generator=g)
    umask = torch.zeros(B, S)
    for b, n in enumerate(lengths):
        umask[b, :n] = 1.0
    keep = umask.t().unsqueeze(-1)                       # (S,B,1)
    qmask = torch.nn.functional.one_hot(spk, 2).float() * keep
    return Batch(text * keep, visual * keep, acoustic * keep, qmask, umask, (label * umask.long()), list(lengths))


def ragged_lengths(n_dialogues: int, lo: int = 10, hi: int = 110, seed: int = 3407) -> List[int]:
    g = torch.Generator().manual_seed(seed + 1)
    return torch.randint(lo, hi + 1, (n_dialogues,), generator=g).tolist()
