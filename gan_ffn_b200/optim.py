"""Fused Adam over the flat parameter arenas.

Same semantics as ``torch.optim.Adam`` as the reference uses it (train_IEMOCAP.py:292-297 and
:661: L2 folded into the gradient, bias correction, eps outside the square root), but one
kernel launch per network arena instead of a foreach over ~150 tensors, and a natural place for
the data-parallel gradient all-reduce (one NCCL call per arena, see ``parallel.py``).
"""
from __future__ import annotations

from typing import Iterable, List

import torch

from . import functional as GF
from ._lib import lib, ptr


def _arena_modules(module: torch.nn.Module):
    from .model import _FusedNet
    return [m for m in module.modules() if isinstance(m, _FusedNet)]


class FusedAdam:
    """``FusedAdam(module_or_modules, lr, betas, eps, weight_decay)``.

    Parameters that live in a network arena are stepped with one ``ganffn_adam_step`` call per
    arena; the few loose parameters (``GAN_FFN.fc``) get one call each.  Parameters whose
    ``.grad`` is ``None`` (the reference's dead prototype layer, ``lstm``, ``smax_fc``) are
    skipped, exactly as ``torch.optim.Adam`` skips them."""

    def __init__(self, modules, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_reducer=None):
        if isinstance(modules, torch.nn.Module):
            modules = [modules]
        self.modules: List[torch.nn.Module] = list(modules)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.grad_reducer = grad_reducer          # parallel.GradReducer or None
        # overlap the gradient all-reduce with the backward pass (per-layer buckets); ``expected_backwards`` = how many
        # backward passes add into a network's arena before step() (1, or 2 for the reference's two-pass train_disc)
        self.overlap_reduce = True
        self.expected_backwards = 1
        # several passes into one arena may only be overlapped when they are ordered behind each other (the network pass
        # and the `object` projection of the batched visual discriminator are); concurrent passes on different lanes
        # (the two-pass train_disc body) fall back to one all-reduce per arena at step()
        self.sequential_backwards = False
        # data parallel: run the arena updates on the communication stream, behind their all-reduces, instead of making
        # the caller's stream wait for the collectives; the arenas' later users wait for ``update_event`` (functional.
        # ParamArena.wait_updated).  Set by the trainers of train.py, which join the events at the end of a batch.
        self.defer_step = False
        self.grad_scale = 1.0
        self.state = {}                           # id(arena or param) -> dict(step, m, v)
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]

    # -- discovery ------------------------------------------------------------------------
    def _arenas(self):
        seen, out = set(), []
        for mod in self.modules:
            for net in _arena_modules(mod):
                ar = net.arena()
                if id(ar) not in seen:
                    seen.add(id(ar))
                    out.append(ar)
        return out

    def _loose(self, arenas):
        owned = {id(p) for ar in arenas for p in ar.params}
        seen, out = set(), []
        for mod in self.modules:
            for p in mod.parameters():
                if id(p) not in owned and id(p) not in seen and p.requires_grad:
                    seen.add(id(p))
                    out.append(p)
        return out

    # -- torch.optim API surface the reference loop uses ------------------------------------
    def zero_grad(self, set_to_none: bool = True):
        GF.join_lanes()
        for mod in self.modules:
            for p in mod.parameters():
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()
        if set_to_none:
            for ar in self._arenas():
                ar.prezero()   # one memset now, on the caller's stream, instead of one inside the first backward
                ar.reduce_works = None
                # data parallel: the backward pass that completes the expected number of passes into this arena launches
                # the all-reduce itself, layer by layer (functional._NetFunction.backward -> GradReducer.reduce_arena_by_layer)
                ar.reduce_hook = ({"reducer": self.grad_reducer, "expected": int(self.expected_backwards), "seen": 0}
                                  if (self.grad_reducer is not None and self.overlap_reduce and ar.grad.is_cuda and
                                      (self.expected_backwards == 1 or self.sequential_backwards)) else None)

    def _state_for(self, key, like: torch.Tensor):
        """Adam state of an arena / loose parameter.  The step count lives on the device (``step_t``) and is bumped by
        a captured torch op, so ``step()`` can be recorded into a CUDA graph and replayed (bias correction is
        derived in the kernel from ``*step_t``); ``step`` mirrors it on the host for eager use."""
        st = self.state.get(id(key))
        if st is None:
            st = {"step": 0, "m": torch.zeros_like(like), "v": torch.zeros_like(like), "keep": key,
                  "step_t": torch.zeros(1, dtype=torch.int32, device=like.device)}
            self.state[id(key)] = st
        st["step"] += 1
        st["step_t"].add_(1)
        return st

    @torch.no_grad()
    def step(self):
        L = lib()
        GF.join_lanes()   # backward passes issued on network lanes must have landed in the gradient arenas
        lr = self.param_groups[0]["lr"]
        b1, b2 = self.betas
        arenas = self._arenas()
        live = [ar for ar in arenas if ar.grads_live()]
        loose = [p for p in self._loose(arenas) if p.grad is not None]
        deferred = (self.grad_reducer is not None and self.defer_step and bool(live) and live[0].flat.is_cuda)
        if deferred:
            # Collectives and arena updates on the communication stream: the caller's stream goes straight on to the next
            # sub-step (whose first network usually is a different one) while this network's gradients are summed and
            # its Adam step runs; its next forward pass / zero_grad waits for ``update_event``.
            cur = torch.cuda.current_stream(live[0].flat.device)
            self.grad_reducer.reduce([p.grad for p in loose])        # a few hundred floats: stays on the caller's stream
            comm = None
            for ar in live:
                ar.reduce_hook = None
                if ar.reduce_works is None:
                    ar.reduce_works = self.grad_reducer.reduce_arena_whole(ar, cur)
                works, comm = ar.reduce_works
                ar.reduce_works = None
                with torch.cuda.stream(comm):
                    for w in works:
                        w.wait()
                    st = self._state_for(ar, ar.flat)
                    GF._call(ar.flat, "ganffn_adam_step_dev", ptr(ar.flat), ptr(ar.grad), ptr(st["m"]), ptr(st["v"]), ar.numel,
                             ptr(st["step_t"]), lr, b1, b2, self.eps, self.weight_decay, self.grad_scale, comm.cuda_stream)
                    ev = torch.cuda.Event()
                    ev.record(comm)
                ar.update_event = ev
            live = []
        elif self.grad_reducer is not None:
            cur = torch.cuda.current_stream() if torch.cuda.is_available() else None
            late = [ar for ar in live if ar.reduce_works is None]
            self.grad_reducer.reduce([ar.grad for ar in late] + [p.grad for p in loose])
            for ar in live:
                ar.reduce_hook = None
                if ar.reduce_works is not None:        # launched from the backward pass, overlapped with it
                    works, comm = ar.reduce_works
                    for w in works:
                        w.wait()
                    cur.wait_stream(comm)
                    ar.reduce_works = None
        for ar in live:
            st = self._state_for(ar, ar.flat)
            GF._call(ar.flat, "ganffn_adam_step_dev", ptr(ar.flat), ptr(ar.grad), ptr(st["m"]), ptr(st["v"]), ar.numel,
                   ptr(st["step_t"]), lr, b1, b2, self.eps, self.weight_decay, self.grad_scale, GF._stream(ar.flat))
        for p in loose:
            st = self._state_for(p, p)
            g = p.grad.contiguous()
            GF._call(p, "ganffn_adam_step_dev", ptr(p), ptr(g), ptr(st["m"]), ptr(st["v"]), p.numel(), ptr(st["step_t"]), lr,
                   b1, b2, self.eps, self.weight_decay, self.grad_scale, GF._stream(p))

    def join_deferred(self, stream=None) -> None:
        """Orders ``stream`` (default: the current one) behind every deferred arena update of this optimizer and clears the
        marks (end of a train step: a captured CUDA graph must have every forked stream joined, and host code that
        reads the parameters afterwards expects them updated on the current stream)."""
        if not torch.cuda.is_available():
            return
        for ar in self._arenas():
            if ar.update_event is not None:
                (stream or torch.cuda.current_stream(ar.flat.device)).wait_event(ar.update_event)
                ar.update_event = None
