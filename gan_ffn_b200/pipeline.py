"""Input pipeline of the train / scoring loops: packed host batches, device-side collate and a
one-batch-ahead host -> device prefetcher.

The reference collates on the CPU (dataloader.py:55-58: ``pad_sequence`` over the dialogues of a batch), moves the
padded batch to the GPU (train_IEMOCAP.py:137-140) and -- in stage 1 -- casts it back to CPU tensors
(``.type(torch.FloatTensor)``, :349-351).  Here the dialogues of a batch are concatenated, *not* padded, on the host
(``pack_dialogues``: bytes proportional to the real utterances), copied to the device from pinned memory on a copy
stream one batch ahead of the compute stream (``DevicePrefetcher``), and padded on the device
(``collate_on_device``: ``ganffn_graph_unpack`` for the three feature tensors, ``ganffn_collate_meta`` for
``qmask`` / ``umask`` / ``label``).  The result is the loader's batch bit for bit (``tests/test_pipeline.py`` compares
it with ``oracle/collate_oracle.py``, the ``pad_sequence`` restatement); nothing goes back to the host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch

from .synthetic import Batch


@dataclass
class PackedBatch:
    """The dialogues of one batch, concatenated in batch order (dialogue-major), as they leave the host."""
    text: torch.Tensor       # [N, 100]
    visual: torch.Tensor     # [N, 512]
    acoustic: torch.Tensor   # [N, 100]
    speaker: torch.Tensor    # [N] int32 (argmax of the loader's one-hot qmask)
    label: torch.Tensor      # [N] int64
    lengths: torch.Tensor    # [B] int32
    node_off: torch.Tensor   # [B+1] int64, exclusive prefix sums of lengths
    lengths_host: List[int]

    @property
    def seq_len(self) -> int:
        return max(self.lengths_host)

    @property
    def n_dialogues(self) -> int:
        return len(self.lengths_host)

    def tensors(self):
        return (self.text, self.visual, self.acoustic, self.speaker, self.label, self.lengths, self.node_off)

    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())

    def pin(self) -> "PackedBatch":
        return PackedBatch(*(t.pin_memory() for t in self.tensors()), self.lengths_host)

    def to(self, device, non_blocking: bool = False) -> "PackedBatch":
        return PackedBatch(*(t.to(device, non_blocking=non_blocking) for t in self.tensors()), self.lengths_host)


def pack_dialogues(items: Sequence[Sequence[torch.Tensor]], pin: bool = True) -> PackedBatch:
    """``items[i]`` is what the reference's ``IEMOCAPDataset.__getitem__`` returns for one dialogue (dataloader.py:41-51):
    ``(text [L,100], visual [L,512], acoustic [L,100], qmask [L,2], umask [L], label [L], ...)``.  Concatenation only:
    the padding the reference's ``collate_fn`` does here happens on the device."""
    lengths = [int(it[0].shape[0]) for it in items]
    if min(lengths) < 1:
        raise ValueError("every dialogue needs at least one utterance")
    cat = lambda k, dt: torch.cat([it[k].to(dt) for it in items], dim=0).contiguous()
    text, visual, acoustic = cat(0, torch.float32), cat(1, torch.float32), cat(2, torch.float32)
    speaker = torch.cat([it[3].argmax(dim=1) for it in items]).to(torch.int32).contiguous()
    label = cat(5, torch.int64)
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    pb = PackedBatch(text, visual, acoustic, speaker, label, torch.tensor(lengths, dtype=torch.int32), torch.from_numpy(off), lengths)
    return pb.pin() if pin and torch.cuda.is_available() else pb


def pack_batch(batch: Batch, pin: bool = True) -> PackedBatch:
    """A zero-padded ``Batch`` (host) -> its packed form (the inverse of ``collate_on_device``)."""
    items = []
    for b, n in enumerate(batch.lengths):
        items.append((batch.text[:n, b], batch.visual[:n, b], batch.acoustic[:n, b], batch.qmask[:n, b], batch.umask[b, :n],
                      batch.label[b, :n]))
    return pack_dialogues(items, pin=pin)


def collate_on_device(pb: PackedBatch, seq_len: Optional[int] = None, n_speakers: int = 2) -> Batch:
    """Pads a device-resident ``PackedBatch`` to ``(S,B,.)`` / ``(B,S)`` on the current stream (``seq_len`` = the
    *global* pad length under data parallelism; default: the longest dialogue of the batch)."""
    from ._lib import lib, ptr
    from . import functional as GF
    GF._require_cuda(pb.text, "packed batch")
    L = lib()
    S, B = int(pb.seq_len if seq_len is None else seq_len), pb.n_dialogues
    if S < pb.seq_len:
        raise ValueError(f"seq_len {S} is shorter than the longest dialogue ({pb.seq_len})")
    dev, st = pb.text.device, GF._stream(pb.text)
    out = {}
    for name in ("text", "visual", "acoustic"):
        src = getattr(pb, name)
        dst = torch.empty((S, B, src.shape[1]), dtype=torch.float32, device=dev)
        L.call("ganffn_graph_unpack", ptr(src), ptr(pb.lengths), ptr(pb.node_off), ptr(dst), S, B, src.shape[1], st)
        out[name] = dst
    qmask = torch.empty((S, B, n_speakers), dtype=torch.float32, device=dev)
    umask = torch.empty((B, S), dtype=torch.float32, device=dev)
    label = torch.empty((B, S), dtype=torch.int64, device=dev)
    L.call("ganffn_collate_meta", ptr(pb.speaker), ptr(pb.label), ptr(pb.lengths), ptr(pb.node_off), ptr(qmask), ptr(umask),
           ptr(label), S, B, n_speakers, st)
    return Batch(out["text"], out["visual"], out["acoustic"], qmask, umask, label, list(pb.lengths_host))


class DevicePrefetcher:
    """Iterates device ``Batch``es over an iterable of pinned ``PackedBatch``es, keeping one batch in flight: the copy of
    batch k+1 runs on a side stream while the compute stream works on batch k; the device-side collate is enqueued on
    the compute stream behind the copy's event (no host synchronisation anywhere).

    ``seq_len`` fixes the pad length (a callable ``seq_len(packed) -> int`` may derive it per batch, e.g. the global
    maximum under data parallelism)."""

    def __init__(self, packed: Iterable[PackedBatch], device, seq_len=None, n_speakers: int = 2):
        self.src, self.device = packed, torch.device(device)
        self.seq_len, self.n_speakers = seq_len, n_speakers
        self.copy_stream = torch.cuda.Stream(device=self.device)

    def _issue(self, pb: PackedBatch):
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            dpb = pb.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        for t in dpb.tensors():
            t.record_stream(cur)            # consumed by the collate kernels on the compute stream
        return dpb, ev, pb

    def __iter__(self) -> Iterator[Batch]:
        it = iter(self.src)
        try:
            nxt = self._issue(next(it))
        except StopIteration:
            return
        while nxt is not None:
            dpb, ev, pb = nxt
            try:
                nxt = self._issue(next(it))   # batch k+1 starts copying before batch k is collated and consumed
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            S = self.seq_len(pb) if callable(self.seq_len) else self.seq_len
            yield collate_on_device(dpb, S, self.n_speakers)
