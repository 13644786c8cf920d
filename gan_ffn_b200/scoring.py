"""Batched inference scoring, sharded by dialogue (BASELINE.json configs 3 and 4; SURVEY.md §8e "inference scoring").

Dialogues are independent, so scoring needs no collective: the batches of a sweep are dealt to the ranks and every
rank scores its own.  What is *not* free is the pad length: the reference passes no key-padding mask, so a dialogue's
output depends on how far its batch is zero-padded.  A batch is therefore the unit of sharding -- a
batch is never split across ranks with different pad lengths -- and the batching policy is part of the result:

  * ``plan_batches(lengths, batch_size, sort=True)``  groups dialogues of similar length (ascending), which is what
    keeps the padded slots close to the real utterances (S*B vs sum(len)); ``sort=False`` keeps loader order.
  * ``shard_batches(batches, world, rank)``           deals whole batches round-robin (longest-first when sorted, so
    every rank gets the same mix of lengths).

"MELD-shaped" (config 3): the reference's GAN-FFN cannot take MELD's 600-d text (its generator widths are literals,
SURVEY.md D4); as SURVEY.md §8d defines it, MELD-shaped here means 7 classes (train_MELD.py:139), batches of 32
(:114), short dialogues (S <= 33) and the 100/512/100 modality widths the constructors fix.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

import torch

from .synthetic import Batch, make_batch


def plan_batches(lengths: Sequence[int], batch_size: int = 32, sort: bool = True) -> List[List[int]]:
    """Dialogue indices grouped into batches (the last one may be short)."""
    order = sorted(range(len(lengths)), key=lambda i: (lengths[i], i)) if sort else list(range(len(lengths)))
    return [order[i:i + batch_size] for i in range(0, len(order), batch_size)]


def shard_batches(batches: Sequence[List[int]], world_size: int, rank: int) -> List[List[int]]:
    """This rank's batches: round-robin over the plan, so no rank gets only the long dialogues."""
    return [b for k, b in enumerate(batches) if k % world_size == rank]


def padded_slots(batches: Iterable[List[int]], lengths: Sequence[int]) -> int:
    return sum(max(lengths[i] for i in b) * len(b) for b in batches)


@torch.no_grad()
def score_batch(model, batch: Batch) -> Dict[str, torch.Tensor]:
    """Eval-mode forward of ``GAN_FFN`` over one batch (the reference's evaluation pass, train_IEMOCAP.py:151-158 with
    ``train=False``): log-probabilities (S,B,C) and the argmax per (dialogue, turn) in the loader's (B,S) order."""
    from . import functional as GF
    model.eval()
    with GF.overlap_networks():   # the three generators are independent: one lane each
        log_prob = model(batch.acoustic, batch.visual, batch.text)[0]
    lp_ = log_prob.transpose(0, 1).contiguous().view(-1, log_prob.size(2))
    pred = torch.argmax(lp_, 1).view(batch.n_dialogues, batch.seq_len)
    return {"log_prob": log_prob, "pred": pred}


class GraphedScorer:
    """``score_batch`` replayed from one CUDA graph per batch shape (pad length x dialogues): the eval forward of the
    three generators is ~170 kernel launches, so an eager sweep is bound by the host (≈6 ms per batch of 32) while the
    device needs ≈2 ms.  First call of a shape runs eagerly (warm-up), the second records, later calls copy the batch
    into the graph's input tensors and replay.  Outputs are owned by the graph: read them before the next call of the
    same shape."""

    def __init__(self, model):
        self.model = model
        self._seen, self._graphs = set(), {}

    @torch.no_grad()
    def __call__(self, batch: Batch) -> Dict[str, torch.Tensor]:
        key = (batch.seq_len, batch.n_dialogues, batch.text.device.index)
        if key not in self._seen:
            self._seen.add(key)
            return score_batch(self.model, batch)
        if key not in self._graphs:
            static = Batch(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in vars(batch).items()})
            torch.cuda.synchronize(batch.text.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = score_batch(self.model, static)
            self._graphs[key] = (graph, static, out)
        graph, static, out = self._graphs[key]
        for name in ("text", "visual", "acoustic"):
            getattr(static, name).copy_(getattr(batch, name), non_blocking=True)
        graph.replay()
        return out


class SyntheticDialogues:
    """A corpus of synthetic dialogues addressed by index: dialogue ``i`` always has the same length and features
    (seeded by ``i``), whichever rank or batch it lands in."""

    def __init__(self, lengths: Sequence[int], n_classes: int = 6, seed: int = 3407):
        self.lengths, self.n_classes, self.seed = list(lengths), n_classes, seed

    @property
    def utterances(self) -> int:
        return sum(self.lengths)

    def batch(self, idx: Sequence[int]) -> Batch:
        """The dialogues ``idx`` zero-padded to the longest of them (``collate_fn`` semantics, dataloader.py:55-58)."""
        S = max(self.lengths[i] for i in idx)
        parts = [make_batch(n_dialogues=1, lengths=[self.lengths[i]], n_classes=self.n_classes, seed=self.seed + 7919 * i)
                 for i in idx]

        def cat(get, dim, pad_dim):
            outs = []
            for p in parts:
                t = get(p)
                pad = S - t.shape[pad_dim]
                if pad:
                    shape = list(t.shape)
                    shape[pad_dim] = pad
                    t = torch.cat([t, t.new_zeros(shape)], dim=pad_dim)
                outs.append(t)
            return torch.cat(outs, dim=dim)

        return Batch(cat(lambda p: p.text, 1, 0), cat(lambda p: p.visual, 1, 0), cat(lambda p: p.acoustic, 1, 0),
                     cat(lambda p: p.qmask, 1, 0), cat(lambda p: p.umask, 0, 1), cat(lambda p: p.label, 0, 1),
                     [self.lengths[i] for i in idx])
