"""Data parallelism by dialogue.

One process per GPU.  A global batch of dialogues is split along dim 1 of ``(S,B,d)`` -- never
along the sequence axis, which is what the reference's ``nn.DataParallel(dim=0)`` does by accident
(train_IEMOCAP.py:587-593, README.md:82-83) -- and every shard keeps the *global* pad length,
because a dialogue's output depends on how far it is zero-padded.  Gradients are
summed with one all-reduce per parameter arena per optimizer step (NCCL over NVLink on GPUs, gloo
in the CPU tests); losses are scaled so that the summed shard gradients equal the single-device
gradient on the whole global batch:

  * BCELoss is a mean over all S*B slots  -> each rank uses local_mean / world_size;
  * MaskedNLLLoss divides by sum(w[label]*umask) -> every rank divides by the *global* sum.

Inference scoring needs no collective at all.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from .synthetic import Batch


def shard_indices(n_dialogues: int, world_size: int, rank: int) -> List[int]:
    """Contiguous, balanced split of dialogue indices (first ``n % world`` ranks get one more)."""
    base, rem = divmod(n_dialogues, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def shard_batch(batch: Batch, world_size: int, rank: int) -> Batch:
    """This rank's dialogues of a global batch, padded to the global seq_len."""
    return batch.dialogues(shard_indices(batch.n_dialogues, world_size, rank))


class GradReducer:
    """Sum-all-reduce of flat gradient buffers across the data-parallel group."""

    def __init__(self, group=None, overlap_min_numel: int = 8_000_000):
        # Per-layer overlapped all-reduce only for arenas of at least this many parameters (the visual generator's
        # 25.8 M: 12.6 MB buckets).  Measured at 2 GPUs (r2): bucketing the 3.6-4.1 M-parameter d=100 networks too adds
        # ~150 small NCCL launches per step that fight the one-CTA-per-SM GEMM kernels for SMs (59.0 vs 57.8 ms/step);
        # their single 14.5 MB all-reduce already overlaps the other sub-step chain.
        self.overlap_min_numel = overlap_min_numel
        if not dist.is_initialized():
            raise RuntimeError("GradReducer needs torch.distributed to be initialised (one process per GPU)")
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.bytes_reduced = 0
        self.calls = 0
        self._streams = {}

    def reduce(self, buffers: Sequence[torch.Tensor]) -> None:
        works = []
        for b in buffers:
            works.append(dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.bytes_reduced += b.numel() * b.element_size()
            self.calls += 1
        for w in works:
            w.wait()

    def reduce_arena_by_layer(self, arena, nlayers: int, bwd_stream: torch.cuda.Stream):
        """All-reduce of one network's gradient arena in per-encoder-layer buckets, each launched as soon as that
        layer's gradients are complete (``ganffn_net_bwd_layer_wait``), so the collective of layer l overlaps the
        backward pass of layers l-1 .. 0 instead of sitting exposed in front of the optimizer.
        Called right after ``ganffn_net_bwd`` was issued on ``bwd_stream``; everything is device-side (capturable).
        Returns the async work handles (``FusedAdam.step`` waits on them) plus the communication stream to join."""
        from ._lib import lib
        comm = self._comm_stream(arena.grad.device)
        tab, flat = arena.table, arena.grad
        starts = [int(tab[l * 12]) for l in range(nlayers)] + [int(tab[nlayers * 12])]
        works = []
        L = lib()
        whole = False
        for l in range(nlayers - 1, 0, -1):
            if not whole and L.cdll.ganffn_net_bwd_layer_wait(bwd_stream.cuda_stream, l, comm.cuda_stream) != 0:
                comm.wait_stream(bwd_stream)        # no per-layer events: every bucket waits for the whole pass
                whole = True
            with torch.cuda.stream(comm):
                works.append(dist.all_reduce(flat[starts[l]:starts[l + 1]], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        comm.wait_stream(bwd_stream)                # layer 0 and the head (+ `object`) are complete with the pass itself
        with torch.cuda.stream(comm):
            works.append(dist.all_reduce(flat[starts[0]:starts[1]], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            works.append(dist.all_reduce(flat[starts[nlayers]:], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.bytes_reduced += flat.numel() * flat.element_size()
        self.calls += len(works)
        return works, comm

    def reduce_arena_whole(self, arena, after_stream: torch.cuda.Stream):
        """One all-reduce of a whole gradient arena on the communication stream, ordered after ``after_stream``."""
        comm = self._comm_stream(arena.grad.device)
        comm.wait_stream(after_stream)
        with torch.cuda.stream(comm):
            works = [dist.all_reduce(arena.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True)]
        self.bytes_reduced += arena.grad.numel() * arena.grad.element_size()
        self.calls += 1
        return works, comm

    def _comm_stream(self, device) -> torch.cuda.Stream:
        st = self._streams.get(device.index)
        if st is None:
            st = torch.cuda.Stream(device=device)
            self._streams[device.index] = st
        return st

    def global_sum(self, value: float, device) -> float:
        """Sum of a host scalar over the group (e.g. dialogues per rank -> global batch size)."""
        t = torch.tensor([float(value)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())

    def local_fraction_tensor(self, local_count: int, device) -> torch.Tensor:
        """local_count / (sum of local_count over the group) as a device scalar.  Every rank calls this for every
        batch (so the collective sequence is the same on all ranks whatever the shard sizes are); the value never
        touches the host, hence the call can be recorded into a CUDA graph."""
        mine = torch.full((1,), float(local_count), dtype=torch.float32, device=device)
        tot = mine.clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
        return (mine / tot).reshape(())

    def sum_tensor(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum of a device tensor over the group (loss logging under data parallelism)."""
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def global_nll_denominator(self, label: torch.Tensor, umask: torch.Tensor, weight: Optional[torch.Tensor]) -> float:
        """sum over the *global* batch of w[label]*umask (MaskedNLLLoss denominator, model.py:78-80)."""
        m = umask.reshape(-1).to(torch.float32)
        w = m if weight is None else weight.to(m.device)[label.reshape(-1)] * m
        den = w.sum().reshape(1)
        dist.all_reduce(den, op=dist.ReduceOp.SUM, group=self.group)
        return float(den.item())

    def global_nll_denominator_tensor(self, label: torch.Tensor, umask: torch.Tensor, weight: Optional[torch.Tensor]) -> torch.Tensor:
        """The same sum as a device scalar, never read on the host: the all-reduce and everything after it can be
        recorded into a CUDA graph (GraphedTrainStep under data parallelism)."""
        m = umask.reshape(-1).to(torch.float32)
        w = m if weight is None else weight.to(m.device)[label.reshape(-1)] * m
        den = w.sum().reshape(1)
        dist.all_reduce(den, op=dist.ReduceOp.SUM, group=self.group)
        return den.reshape(())


def init_from_env(backend: Optional[str] = None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns
    (rank, local_rank, world_size)."""
    import os
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shutdown(steppers=(), timeout_s: float = 30.0) -> bool:
    """Orderly teardown of a data-parallel process: release recorded CUDA graphs (they reference the communicator's
    kernels), drain the device, barrier, then destroy the process group.  ``destroy_process_group`` runs under a
    watchdog: if the backend does not return within ``timeout_s`` the function reports False and the caller decides
    (bench.py then exits the process, which releases everything).  Returns True on a clean teardown."""
    import threading
    for st in steppers:
        st.release()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if not dist.is_initialized():
        return True
    dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    done = threading.Event()

    def _destroy():
        try:
            dist.destroy_process_group()
        finally:
            done.set()

    th = threading.Thread(target=_destroy, daemon=True)
    th.start()
    return done.wait(timeout_s)
