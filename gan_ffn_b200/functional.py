"""Autograd bridges between PyTorch tensors and the C ABI (``include/ganffn.h``).

PyTorch is plumbing here: it owns device memory, the stream and the autograd graph
between networks.  Every number is produced by ``libganffn.so``.  Each ``Function``
hands raw device pointers to one C entry point in ``forward`` and one in ``backward``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from ._lib import lib, ptr

# --------------------------------------------------------------------------------------
# dropout seed stream
# --------------------------------------------------------------------------------------
_MASK64 = (1 << 64) - 1
_seed_state = {"base": None, "ctr": 0}


def manual_seed(seed: int) -> None:
    """Re-seeds the dropout stream of the kernels (independent of torch's generators)."""
    _seed_state["base"] = int(seed) & _MASK64
    _seed_state["ctr"] = 0


def next_seed() -> int:
    if _seed_state["base"] is None:
        _seed_state["base"] = int(torch.initial_seed()) & _MASK64
    _seed_state["ctr"] += 1
    x = (_seed_state["base"] * 0x9E3779B97F4A7C15 + _seed_state["ctr"] * 0xD1B54A32D192ED03) & _MASK64
    x ^= x >> 31
    return x


class DeviceSeedStream:
    """Graph-safe dropout seeds.  ``slots`` is a device int64 vector that ``advance()`` refills *on the device* from
    (base, step counter) with a splitmix64-style hash, using only torch ops, so the refill can be captured into a
    CUDA graph: every replay draws fresh masks although the kernels' arguments are frozen.  Each network forward of
    a step takes the next slot (``take()``) and hands its address to the kernels (``seed_dev`` of ganffn_net_fwd);
    the backward pass of that call re-reads the same word."""

    _K = [0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, 0xD1B54A32D192ED03]

    @staticmethod
    def _s64(x: int) -> int:
        x &= _MASK64
        return x - (1 << 64) if x >= (1 << 63) else x

    def __init__(self, device, n_slots: int = 64, base: Optional[int] = None):
        self.device = torch.device(device)
        self.n = n_slots
        base = int(torch.initial_seed()) if base is None else int(base)
        self.base = torch.tensor([self._s64(base * self._K[3])], dtype=torch.int64, device=self.device)
        self.counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.idx = torch.arange(1, n_slots + 1, dtype=torch.int64, device=self.device) * self._s64(self._K[0])
        self.slots = torch.zeros(n_slots, dtype=torch.int64, device=self.device)
        self.next = 0
        self.advance()

    def advance(self) -> None:
        """Start of a step: new seeds for all slots (device-side; capturable)."""
        self.counter.add_(1)
        z = self.base + self.counter * self._s64(self._K[2]) + self.idx
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * self._s64(self._K[1])
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * self._s64(self._K[2])
        z = z ^ ((z >> 31) & ((1 << 33) - 1))
        self.slots.copy_(z)
        self.next = 0

    def take(self) -> int:
        """Device address of the next seed word of this step."""
        if self.next >= self.n:
            raise RuntimeError(f"DeviceSeedStream: more than {self.n} dropout-drawing forward calls in one step")
        p = self.slots.data_ptr() + 8 * self.next
        self.next += 1
        return p

    def value(self, slot: int) -> int:
        """Host copy of a slot as the unsigned 64-bit seed the kernels see (tests)."""
        return int(self.slots[slot].item()) & _MASK64


_seed_stream: Optional[DeviceSeedStream] = None


def set_seed_stream(stream: Optional[DeviceSeedStream]) -> Optional[DeviceSeedStream]:
    """Route the dropout seeds of all train-mode forwards through a DeviceSeedStream (None = host seeds)."""
    global _seed_stream
    prev, _seed_stream = _seed_stream, stream
    return prev


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _call(t: torch.Tensor, name: str, *args) -> None:
    """C-ABI call with ``t``'s device current: the kernels launch on (and opt their shared memory in for) the
    current device, which need not be the tensor's in a process that drives several GPUs."""
    idx = t.device.index
    if idx is not None and idx != torch.cuda.current_device():
        with torch.cuda.device(idx):
            lib().call(name, *args)
    else:
        lib().call(name, *args)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: got a {t.device} tensor. gan_ffn_b200 computes only with its sm_100a CUDA kernels; "
            "there is no CPU fallback (use oracle/ for CPU checking).")
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32, got {t.dtype}")


_scratch: Dict[Tuple[str, int], torch.Tensor] = {}
_scratch_retired: List[torch.Tensor] = []


def _scratch_tag(device: torch.device) -> str:
    """One workspace per stream: networks on different lanes / sub-step chains run concurrently."""
    h = torch.cuda.current_stream(device).cuda_stream
    return f"s{h}" if h else "main"


def scratch(device: torch.device, floats: int, tag: str = "main") -> torch.Tensor:
    """A grow-only workspace per (device, tag).  All kernels of this package are issued on
    the current stream, so one workspace per device is safe to share between calls."""
    key = (tag, device.index if device.index is not None else torch.cuda.current_device())
    buf = _scratch.get(key)
    if buf is None or buf.numel() < floats:
        if buf is not None:
            _scratch_retired.append(buf)   # a recorded CUDA graph may still hold this address: never free a workspace
        buf = torch.empty(max(int(floats), 1), dtype=torch.float32, device=device)
        _scratch[key] = buf
    return buf


# --------------------------------------------------------------------------------------
# network lanes: independent networks of one loop body on concurrent streams
# --------------------------------------------------------------------------------------
class _Lanes:
    """Inside ``overlap_networks()`` every whole-network call (``net_forward``) is issued on one of ``n`` side
    streams ("lanes", round-robin) instead of the caller's stream, so that networks that do not depend on each
    other overlap on the device: ``disc(real)`` with ``gen(real_gen)`` and the two discriminator backward passes
    of ``train_disc`` (train_IEMOCAP.py:217-225), the three generators of ``GAN_FFN.forward`` (model.py:1440-1442)
    and their backward passes.  At S*B = 3008 slots most kernels of the d=100 networks are 24..144-CTA grids on a
    148-SM device, so two or three of them fit side by side.

    Ordering rules (all device-side, capturable into a CUDA graph):
      * fork: a lane waits for everything already enqueued on the caller's stream, and for the producing lane of
        its input when that was another network of the same context;
      * join: every other entry point of this package (losses, fuse_classify, optimizer step, context exit) first
        makes the caller's stream wait for all lanes that carry un-joined work;
      * backward: autograd runs each network's backward on the lane its forward ran on and orders gradient
        hand-overs between streams itself.
    Only code that consumes network outputs through this package's own entry points may run inside the context
    (the two loop bodies of train.py do); a plain torch op on a network output would not wait for its lane."""

    def __init__(self):
        self.active = False
        self.n = 3
        self.streams: Dict[Tuple[int, int], List[torch.cuda.Stream]] = {}
        self.busy: Dict[int, torch.cuda.Stream] = {}
        self.producers: Dict[int, Tuple[torch.cuda.Event, torch.cuda.Stream]] = {}
        self.counter = 0
        self.chain_handles = set()      # streams registered as sub-step chains (register_chain_stream)

    def lane(self, device: torch.device) -> Tuple[int, torch.cuda.Stream]:
        """Next lane of the pool that belongs to the caller's current stream: sub-step chains (train.GANTrainer) each
        fork their own lanes, so kernels of concurrent chains never queue behind each other on a shared lane."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = torch.cuda.current_stream(device).cuda_stream
        key = (idx, h if h in self.chain_handles else 0)   # one shared pool for every non-chain caller stream
        pool = self.streams.get(key)
        if pool is None:
            pool = [torch.cuda.Stream(device=device) for _ in range(self.n)]
            self.streams[key] = pool
        k = self.counter % self.n
        self.counter += 1
        return k, pool[k]

    def is_lane(self, handle: int) -> bool:
        return any(st.cuda_stream == handle for pool in self.streams.values() for st in pool)

    def touch(self, stream: torch.cuda.Stream) -> None:
        self.busy[stream.cuda_stream] = stream

    def join(self) -> None:
        """The caller's current stream waits for all lanes with un-joined work."""
        if not self.busy:
            return
        main = torch.cuda.current_stream()
        for st in self.busy.values():
            if st.cuda_stream != main.cuda_stream:
                main.wait_stream(st)
        self.busy.clear()
        self.producers.clear()


_lanes = _Lanes()


class overlap_networks:
    """Context manager: run the networks called inside on concurrent lanes (see ``_Lanes``)."""

    def __init__(self, enabled: bool = True, lanes: int = 3):
        self.enabled, self.lanes = enabled, lanes

    def __enter__(self):
        self.prev = _lanes.active
        if self.enabled and not _deterministic["on"]:
            _lanes.n = self.lanes
            _lanes.active = True
        return self

    def __exit__(self, *exc):
        _lanes.join()
        _lanes.active = self.prev
        return False


def join_lanes() -> None:
    _lanes.join()


def register_chain_stream(stream: torch.cuda.Stream) -> None:
    """Marks ``stream`` as a sub-step chain: networks called on it fork their own pool of lanes."""
    _lanes.chain_handles.add(stream.cuda_stream)


_deterministic = {"on": False}


def set_deterministic(on: bool = True) -> bool:
    """Bit-reproducible training steps (the reference pins determinism, train_IEMOCAP.py:46-53): the kernels
    accumulate gradients in a fixed order (``ganffn_set_deterministic``) and the networks of a loop body run serially
    instead of on concurrent lanes (two backward passes must not add to one arena at the same time).  Costs
    throughput; off by default.  Returns the previous setting."""
    prev = _deterministic["on"]
    _deterministic["on"] = bool(on)
    lib().cdll.ganffn_set_deterministic(int(bool(on)))
    return prev


def _al(n: int, a: int = 32) -> int:
    return (n + a - 1) // a * a


# --------------------------------------------------------------------------------------
# flat parameter arena
# --------------------------------------------------------------------------------------
class ParamArena:
    """One contiguous fp32 buffer for all *live* parameters of a network and a twin buffer
    for their gradients.  The ``nn.Parameter`` objects keep their names (so ``state_dict``
    keys match the reference) but their ``.data`` become views into ``flat`` and their
    ``.grad`` views into ``grad``.  One pointer + an offset table is all the C side needs;
    Adam and the NCCL gradient all-reduce each become a single call over the arena."""

    def __init__(self, params: List[torch.nn.Parameter], table: List[Optional[torch.nn.Parameter]]):
        dev = params[0].device
        offs, total = {}, 0
        for p in params:
            offs[id(p)] = total
            total += _al(p.numel())
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.params = params
        self.grad_views = []
        with torch.no_grad():
            for p in params:
                o = offs[id(p)]
                view = self.flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self.grad_views.append(self.grad[o:o + p.numel()].view(p.shape))
        self.table = np.array([-1 if p is None else offs[id(p)] for p in table], dtype=np.int64)
        self.numel = total
        self.sentinel = params[0]
        self.requires_grad = any(p.requires_grad for p in params)
        self.prezeroed = False      # FusedAdam.zero_grad() already cleared ``grad`` on the caller's stream
        self.zero_event, self.zero_stream = None, None
        # data parallelism: FusedAdam.zero_grad() arms ``reduce_hook`` = {"reducer", "expected", "seen"}; the backward pass
        # that completes the expected count launches the per-layer all-reduces (overlapped with the rest of that
        # pass) and leaves their handles in ``reduce_works`` for FusedAdam.step() to wait on
        self.reduce_hook, self.reduce_works = None, None
        # parameter gradients switched off for the calls made inside ``frozen_parameters(module)``
        self.frozen = False
        # data parallel: FusedAdam may run this arena's update on the communication stream, behind its gradient
        # all-reduce (``defer_step``); ``update_event`` then marks the update, and every later reader / writer of the
        # arena (forward passes, zero_grad) waits for it on its own stream (``wait_updated``)
        self.update_event = None

    def wait_updated(self, stream=None) -> None:
        """Orders ``stream`` (default: the current one) behind a deferred optimizer update of this arena."""
        ev = self.update_event
        if ev is not None:
            (stream or torch.cuda.current_stream(self.flat.device)).wait_event(ev)

    def prezero(self) -> None:
        """zero_grad(set_to_none=True) for an arena: the ``.grad`` views are dropped by the caller; the buffer is
        cleared now, on the caller's stream, so that backward passes on several lanes can accumulate into it."""
        self.wait_updated()          # a deferred Adam step may still be reading the gradients
        self.grad.zero_()
        self.prezeroed = True
        self.zero_event, self.zero_stream = None, None

    def valid_for(self, dev: torch.device) -> bool:
        o = 0
        return (self.flat.device == dev and self.sentinel.data_ptr() == self.flat.data_ptr() + 4 * o
                and self.params[-1].device == dev)

    def grads_live(self) -> bool:
        g = self.sentinel.grad
        return g is not None and g.data_ptr() == self.grad_views[0].data_ptr()

    def install_grads(self) -> None:
        for p, gv in zip(self.params, self.grad_views):
            if p.requires_grad:
                p.grad = gv


# --------------------------------------------------------------------------------------
# whole-network function
# --------------------------------------------------------------------------------------
class NetSpec:
    """Static description of one generator / discriminator (mirrors the C NetDims)."""

    def __init__(self, kind: int, d: int, nhead: int, dff: int, nlayers: int, h1: int, h2: int):
        self.kind, self.d, self.nhead, self.dff, self.nlayers, self.h1, self.h2 = kind, d, nhead, dff, nlayers, h1, h2

    def dims(self, S: int, B: int, d_in: int):
        return (self.kind, S, B, d_in, self.d, self.nhead, self.dff, self.nlayers, self.h1, self.h2)


class _NetFunction(torch.autograd.Function):
    """forward: ganffn_net_fwd; backward: ganffn_net_bwd.  Parameter gradients are written
    straight into the arena's gradient buffer (accumulating when the buffer is live), so the
    only autograd edge is the input ``x``."""

    @staticmethod
    def forward(ctx, x, anchor, arena: ParamArena, spec: NetSpec, pe, train: bool, p_head: float, seed: int, seed_ptr,
                param_grads: bool = True):
        L = lib()
        ctx.param_grads = param_grads
        arena.wait_updated()         # deferred optimizer step on the communication stream (data parallel)
        S, B, d_in = x.shape
        dims = spec.dims(S, B, d_in)
        stash_n = L.query("ganffn_net_stash_floats", *dims)
        scratch_n = L.query("ganffn_net_scratch_floats", *dims)
        stash = torch.empty(stash_n, dtype=torch.float32, device=x.device)
        ws = scratch(x.device, scratch_n, _scratch_tag(x.device))
        out = torch.empty((S, B, spec.h2 if spec.kind == 0 else 1), dtype=torch.float32, device=x.device)
        _call(x, "ganffn_net_fwd", spec.kind, ptr(arena.flat), arena.table.ctypes.data, ptr(pe), ptr(x), ptr(out),
               ptr(stash), ptr(ws), S, B, d_in, spec.d, spec.nhead, spec.dff, spec.nlayers, spec.h1, spec.h2,
               int(train), float(p_head), seed, seed_ptr, _stream(x))
        ctx.arena, ctx.spec, ctx.train, ctx.p_head, ctx.seed, ctx.seed_ptr = arena, spec, train, p_head, seed, seed_ptr
        ctx.stash = stash
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, d_out):
        L = lib()
        (x, out), arena, spec = ctx.saved_tensors, ctx.arena, ctx.spec
        S, B, d_in = x.shape
        dims = spec.dims(S, B, d_in)
        d_out = d_out.contiguous()
        ws = scratch(x.device, L.query("ganffn_net_scratch_floats", *dims), _scratch_tag(x.device))
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        cur = torch.cuda.current_stream(x.device)
        if _lanes.streams and _lanes.is_lane(cur.cuda_stream):
            _lanes.touch(cur)   # a backward pass on a lane (autograd runs it where the forward ran): join before use
        if not ctx.param_grads:
            # frozen network (frozen_parameters): data gradient only, the gradient arena is not touched
            if dx is not None:
                _call(x, "ganffn_net_bwd", spec.kind, ptr(arena.flat), arena.table.ctypes.data, ptr(x), ptr(out), ptr(d_out),
                      ptr(ctx.stash), None, ptr(dx), ptr(ws), S, B, d_in, spec.d, spec.nhead, spec.dff,
                      spec.nlayers, spec.h1, spec.h2, int(ctx.train), float(ctx.p_head), ctx.seed, ctx.seed_ptr, 1, _stream(x))
            ctx.stash = None
            return dx, None, None, None, None, None, None, None, None, None
        if not arena.grads_live():
            # Parameter gradients are accumulated by the kernels (red.global.add from the wgrad GEMMs and the
            # LayerNorm backward), torch-style: a fresh backward starts from a zeroed arena (one memset) -- done
            # by FusedAdam.zero_grad() on the caller's stream (``prezeroed``), else here.
            if not arena.prezeroed:
                arena.grad.zero_()
                arena.zero_event = torch.cuda.Event()
                arena.zero_event.record(cur)
                arena.zero_stream = cur.cuda_stream
            arena.prezeroed = False
            arena.install_grads()
        elif arena.zero_event is not None and arena.zero_stream != cur.cuda_stream:
            cur.wait_event(arena.zero_event)   # another lane zeroed the arena for this backward pass
        _call(x, "ganffn_net_bwd", spec.kind, ptr(arena.flat), arena.table.ctypes.data, ptr(x), ptr(out), ptr(d_out),
               ptr(ctx.stash), ptr(arena.grad), ptr(dx), ptr(ws), S, B, d_in, spec.d, spec.nhead, spec.dff,
               spec.nlayers, spec.h1, spec.h2, int(ctx.train), float(ctx.p_head), ctx.seed, ctx.seed_ptr, 1, _stream(x))
        ctx.stash = None
        hook = arena.reduce_hook
        if hook is not None:
            hook["seen"] += 1
            if hook["seen"] == hook["expected"]:
                red = hook["reducer"]
                # large arenas: per-encoder-layer buckets overlapped with the rest of this pass; small ones: one
                # all-reduce as soon as the pass is complete (it then overlaps whatever else the step is doing)
                arena.reduce_works = (red.reduce_arena_by_layer(arena, spec.nlayers, cur)
                                      if arena.numel >= getattr(red, "overlap_min_numel", 0) else red.reduce_arena_whole(arena, cur))
                arena.reduce_hook = None
        return dx, None, None, None, None, None, None, None, None, None


class frozen_parameters:
    """``with frozen_parameters(net): y = net(x)`` -- calls of ``net`` made inside the block back-propagate to their
    input only; no weight, bias or LayerNorm gradient of ``net`` is computed or accumulated (the usual ``requires_grad_
    (False)`` freeze of a discriminator while its generator trains, without touching a hundred parameter flags per
    sub-step).  The reference's ``train_gen`` (train_IEMOCAP.py:230-252) does compute the discriminator's parameter
    gradients, but nothing ever reads them: only the generator's optimizer steps, and the discriminator's next use is
    ``train_disc``, which starts with ``opt.zero_grad()`` (:221)."""

    def __init__(self, net):
        self.arena = net.arena()

    def __enter__(self):
        self.prev = self.arena.frozen
        self.arena.frozen = True
        return self

    def __exit__(self, *exc):
        self.arena.frozen = self.prev
        return False


def net_forward(x: torch.Tensor, arena: ParamArena, spec: NetSpec, pe: torch.Tensor, train: bool, p_head: float):
    _require_cuda(x, "network input")
    if x.dim() != 3:
        raise ValueError(f"expected (seq_len, batch, dim) input, got shape {tuple(x.shape)}")
    x = x.contiguous()
    seed, seed_ptr = 0, None
    if train:
        if _seed_stream is not None:
            seed_ptr = _seed_stream.take()
        else:
            seed = next_seed()
    anchor = arena.flat
    param_grads = arena.requires_grad and not arena.frozen
    if param_grads and torch.is_grad_enabled():
        anchor = arena.flat.detach().requires_grad_(True)  # makes autograd call backward even for data inputs
    if not _lanes.active:
        return _NetFunction.apply(x, anchor, arena, spec, pe, train, p_head, seed, seed_ptr, param_grads)
    main = torch.cuda.current_stream(x.device)
    k, lane = _lanes.lane(x.device)
    lane.wait_stream(main)                                   # fork
    src = _lanes.producers.get(x.untyped_storage().data_ptr())
    if src is not None and src[1].cuda_stream != lane.cuda_stream:
        lane.wait_event(src[0])                              # input made by another network still on its lane
    x.record_stream(lane)
    with torch.cuda.stream(lane):
        out = _NetFunction.apply(x, anchor, arena, spec, pe, train, p_head, seed, seed_ptr, param_grads)
        ev = torch.cuda.Event()
        ev.record(lane)
    out.record_stream(main)
    _lanes.producers[out.untyped_storage().data_ptr()] = (ev, lane)
    _lanes.touch(lane)
    return out


# --------------------------------------------------------------------------------------
# a single nn.Linear whose parameters live in a network arena (VisualDiscriminator.object, model.py:1344, 1355-1356)
# --------------------------------------------------------------------------------------
class _ArenaLinearFunction(torch.autograd.Function):
    """y = x W^T + b with W, b at offsets (ow, ob) of ``arena``; the weight / bias gradients are accumulated straight
    into the arena's gradient buffer (like _NetFunction).  Used when the `object` projection of the visual
    discriminator has to run on its own: the batched real|fake pass of ``train_disc_batched`` projects only the
    real half."""

    @staticmethod
    def forward(ctx, x, anchor, arena: ParamArena, ow: int, ob: int, n_out: int):
        L = lib()
        arena.wait_updated()
        S, B, K = x.shape
        M = S * B
        y = torch.empty((S, B, n_out), dtype=torch.float32, device=x.device)
        nsc = L.query("ganffn_gemm_scratch_floats", M, n_out, K)
        ws = scratch(x.device, nsc, _scratch_tag(x.device) + "/lin")
        _call(x, "ganffn_linear_fwd", ptr(x), arena.flat.data_ptr() + 4 * ow, arena.flat.data_ptr() + 4 * ob, None, ptr(y),
              None, M, n_out, K, 0, 0, 0.0, 0, 0, ptr(ws), nsc, _stream(x))
        ctx.arena, ctx.ow, ctx.ob, ctx.n_out = arena, ow, ob, n_out
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = lib()
        (x,), arena = ctx.saved_tensors, ctx.arena
        S, B, K = x.shape
        M = S * B
        dy = dy.contiguous()
        cur = torch.cuda.current_stream(x.device)
        if _lanes.streams and _lanes.is_lane(cur.cuda_stream):
            _lanes.touch(cur)
        if not arena.grads_live():
            if not arena.prezeroed:
                arena.grad.zero_()
                arena.zero_event = torch.cuda.Event()
                arena.zero_event.record(cur)
                arena.zero_stream = cur.cuda_stream
            arena.prezeroed = False
            arena.install_grads()
        elif arena.zero_event is not None and arena.zero_stream != cur.cuda_stream:
            cur.wait_event(arena.zero_event)
        nsc = L.query("ganffn_wgrad_scratch_floats", M, ctx.n_out, K)
        ws = scratch(x.device, nsc, _scratch_tag(x.device) + "/lin")
        _call(x, "ganffn_linear_wgrad", ptr(dy), ptr(x), arena.grad.data_ptr() + 4 * ctx.ow,
              arena.grad.data_ptr() + 4 * ctx.ob, M, ctx.n_out, K, 1, ptr(ws), _stream(x))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            nsd = L.query("ganffn_gemm_scratch_floats", M, K, ctx.n_out)
            wsd = scratch(x.device, nsd, _scratch_tag(x.device) + "/lin")
            _call(x, "ganffn_linear_dgrad", ptr(dy), arena.flat.data_ptr() + 4 * ctx.ow, None, ptr(dx), M, ctx.n_out, K,
                  ptr(wsd), nsd, _stream(x))
        hook = arena.reduce_hook
        if hook is not None:       # this pass counts as one of the expected passes into the arena (see _NetFunction.backward)
            hook["seen"] += 1
            if hook["seen"] == hook["expected"]:
                arena.reduce_works = hook["reducer"].reduce_arena_whole(arena, cur)
                arena.reduce_hook = None
        return dx, None, None, None, None, None


def arena_linear(x: torch.Tensor, arena: ParamArena, ow: int, ob: int, n_out: int) -> torch.Tensor:
    _require_cuda(x, "linear input")
    x = x.contiguous()
    anchor = arena.flat
    if arena.requires_grad and torch.is_grad_enabled():
        anchor = arena.flat.detach().requires_grad_(True)
    _lanes.join()          # runs on the caller's stream: inputs made on lanes must have landed
    return _ArenaLinearFunction.apply(x, anchor, arena, ow, ob, n_out)


# --------------------------------------------------------------------------------------
# fusion + classifier + log-softmax   (GAN_FFN.forward, model.py:1444-1449)
# --------------------------------------------------------------------------------------
class _FuseClsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, v, t, w, b):
        L = lib()
        S, B, dh = a.shape
        C = w.shape[0]
        T = S * B
        a, v, t, w, b = a.contiguous(), v.contiguous(), t.contiguous(), w.contiguous(), b.contiguous()
        fusion = torch.empty((S, B, dh), dtype=torch.float32, device=a.device)
        logp = torch.empty((S, B, C), dtype=torch.float32, device=a.device)
        _call(a, "ganffn_fuse_cls_fwd", ptr(a), ptr(v), ptr(t), ptr(w), ptr(b), ptr(fusion), ptr(logp), T, dh, C,
               _stream(a))
        ctx.save_for_backward(fusion, logp, w)
        return logp

    @staticmethod
    def backward(ctx, d_logp):
        L = lib()
        fusion, logp, w = ctx.saved_tensors
        S, B, dh = fusion.shape
        C = w.shape[0]
        T = S * B
        d_logp = d_logp.contiguous()
        d_fusion = torch.empty_like(fusion)
        dw = torch.empty_like(w)
        db = torch.empty(C, dtype=torch.float32, device=w.device)
        ws = scratch(w.device, L.query("ganffn_fuse_cls_scratch_floats", T, dh, C), "cls")
        _call(w, "ganffn_fuse_cls_bwd", ptr(d_logp), ptr(logp), ptr(fusion), ptr(w), ptr(d_fusion), ptr(dw), ptr(db), T,
               dh, C, 0, ptr(ws), _stream(w))
        return d_fusion, d_fusion, d_fusion, dw, db


def fuse_classify(a, v, t, w, b):
    for z in (a, v, t, w, b):
        _require_cuda(z, "fuse_classify")
    _lanes.join()
    return _FuseClsFunction.apply(a, v, t, w, b)


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
class _MaskedNLLFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, weight, den_override: float):
        L = lib()
        n, C = pred.shape
        pred = pred.contiguous()
        out = torch.empty(2, dtype=torch.float32, device=pred.device)
        _call(pred, "ganffn_masked_nll_fwd", ptr(pred), ptr(target), ptr(mask), ptr(weight), ptr(out), n, C,
               float(den_override), _stream(pred))
        ctx.save_for_backward(target, mask)
        ctx.out, ctx.weight, ctx.shape = out, weight, (n, C)
        return out[0]

    @staticmethod
    def backward(ctx, d_loss):
        L = lib()
        target, mask = ctx.saved_tensors
        out = ctx.out
        n, C = ctx.shape
        d_loss = d_loss.contiguous()
        d_pred = torch.empty((n, C), dtype=torch.float32, device=out.device)
        _call(out, "ganffn_masked_nll_bwd", ptr(d_loss), ptr(out), ptr(target), ptr(mask), ptr(ctx.weight), ptr(d_pred), n,
               C, _stream(out))
        return d_pred, None, None, None, None


def masked_nll(pred, target, mask, weight=None, den_override: float = 0.0):
    _require_cuda(pred, "MaskedNLLLoss pred")
    _lanes.join()
    if pred.dim() != 2:
        raise ValueError(f"MaskedNLLLoss: pred must be (batch*seq_len, n_classes), got {tuple(pred.shape)}")
    target = target.reshape(-1).to(torch.int64).contiguous()
    mask = mask.reshape(-1).to(torch.float32).contiguous()
    if target.numel() != pred.shape[0] or mask.numel() != pred.shape[0]:
        raise ValueError("MaskedNLLLoss: pred, target and mask disagree on batch*seq_len")
    if weight is not None:
        weight = weight.to(device=pred.device, dtype=torch.float32).contiguous()
    return _MaskedNLLFunction.apply(pred, target, mask, weight, den_override)


class _BCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, target, scale: float):
        L = lib()
        prob, target = prob.contiguous(), target.contiguous()
        out = torch.empty(1, dtype=torch.float32, device=prob.device)
        _call(prob, "ganffn_bce_fwd", ptr(prob), ptr(target), ptr(out), prob.numel(), float(scale), _stream(prob))
        ctx.save_for_backward(prob, target)
        ctx.scale = scale
        return out.view(())

    @staticmethod
    def backward(ctx, d_loss):
        L = lib()
        prob, target = ctx.saved_tensors
        d_loss = d_loss.contiguous()
        d_prob = torch.empty_like(prob)
        _call(prob, "ganffn_bce_bwd", ptr(d_loss), ptr(prob), ptr(target), ptr(d_prob), prob.numel(), float(ctx.scale),
               _stream(prob))
        return d_prob, None, None


def bce(prob, target, scale: float = 1.0):
    _require_cuda(prob, "BCELoss input")
    _lanes.join()
    if prob.shape != target.shape:
        raise ValueError(f"BCELoss: input {tuple(prob.shape)} and target {tuple(target.shape)} differ")
    target = target.to(device=prob.device, dtype=torch.float32)
    return _BCEFunction.apply(prob, target, scale)


# --------------------------------------------------------------------------------------
# dropout-mask export (tests: inject the kernels' masks into the CPU oracle)
# --------------------------------------------------------------------------------------
def dropout_mask(rows: int, cols: int, p: float, seed: int, site: int, row_stride: Optional[int] = None,
                 device="cuda") -> torch.Tensor:
    out = torch.empty((rows, cols), dtype=torch.float32, device=device)
    _call(out, "ganffn_dropout_mask", ptr(out), rows, cols, cols if row_stride is None else row_stride, float(p),
               seed, site, _stream(out))
    return out
