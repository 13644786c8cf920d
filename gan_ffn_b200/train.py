"""The reference's two training loops, bodies unchanged, over our modules.

``train_disc`` / ``train_gen`` / ``gan_batch`` restate reference train_IEMOCAP.py:200-252 and the
twelve pairings of :355-382; ``classifier_step`` restates the body of ``train_or_eval_model``
(:127-170).  Only the constructors differ from the reference scripts: the loss modules are
``gan_ffn_b200.BCELoss`` / ``MaskedNLLLoss`` and the optimizers ``gan_ffn_b200.FusedAdam``
(same hyper-parameters: train_IEMOCAP.py:292-297, :602-606, :661).

The reference's ``.type(torch.FloatTensor)`` round trip that sends the stage-1 inputs back to the CPU
(train_IEMOCAP.py:349-351, :40) is a bug of the script, not part of the arithmetic, and is not
reproduced: batches stay device-resident.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import functional as GF
from . import model as M
from .optim import FusedAdam
from .synthetic import Batch

GAN_LR, GAN_B1, GAN_B2 = 1e-4, 0.5, 0.6          # train_IEMOCAP.py:602-606
FFN_LR, FFN_L2 = 1e-4, 0.008                     # train_IEMOCAP.py:456-457 (--lr, --l2 defaults)


def train_disc(disc, real_dics, gen, real_gen, opt, adversarial_loss, valid, fake):
    """reference train_IEMOCAP.py:200-227."""
    disc.train()
    gen.eval()

    opt.zero_grad()
    real_prob = disc(real_dics)
    fusion = gen(real_gen)
    fake_prob = disc(fusion.detach())
    d_loss = (adversarial_loss(real_prob, valid) + adversarial_loss(fake_prob, fake)) / 2.0
    res = d_loss.detach()
    d_loss.backward()
    opt.step()
    return res


def train_disc_batched(disc, real_dics, gen, real_gen, opt, adversarial_loss, valid, fake):
    """``train_disc`` with the discriminator's two passes (real features, generated features: the same weights) run
    as ONE pass over 2B dialogues:  prob = D([real | G(x).detach()]) and
    (BCE(prob_real, 1) + BCE(prob_fake, 0)) / 2 == BCE(prob, [1 | 0]) because both halves have S*B slots.
    Same losses, gradients and update as reference train_IEMOCAP.py:200-227 (dialogues do not interact; only the
    dropout-mask stream and the fp32 summation order differ); every kernel of the discriminator sees M = 2*S*B rows,
    so the d=100 products launch twice the CTAs for the same fixed per-launch cost."""
    disc.train()
    gen.eval()

    opt.zero_grad()
    fusion = gen(real_gen)
    real_in = disc.project_real(real_dics)               # `object` 512 -> 100 for real visual features (model.py:1355)
    GF.join_lanes()                                       # torch.cat is a plain torch op on the caller's stream
    both = torch.cat([real_in, fusion.detach()], dim=1)   # (S, 2B, D_h): dialogues side by side, global pad length kept
    prob = disc(both)
    d_loss = adversarial_loss(prob, torch.cat([valid, fake], dim=1))
    res = d_loss.detach()
    d_loss.backward()
    opt.step()
    return res


def train_gen(gen, real_gen, disc, opt, adversarial_loss, valid, fake):
    """reference train_IEMOCAP.py:230-252."""
    gen.train()
    disc.eval()

    opt.zero_grad()
    fusion = gen(real_gen)
    prob = disc(fusion)
    g_loss = adversarial_loss(prob, valid)
    res = g_loss.detach()
    g_loss.backward()
    opt.step()
    return res


def train_gen_frozen_disc(gen, real_gen, disc, opt, adversarial_loss, valid, fake):
    """``train_gen`` with the discriminator frozen for the sub-step: the generator's loss, gradients and update are
    those of reference train_IEMOCAP.py:230-252; the discriminator's weight gradients -- which the reference computes
    and never uses (its next use, ``train_disc``, starts with ``opt.zero_grad()``, :221) -- are not computed
    (a quarter of the backward GEMM work of the stage-1 batch)."""
    gen.train()
    disc.eval()

    opt.zero_grad()
    fusion = gen(real_gen)
    with GF.frozen_parameters(disc):
        prob = disc(fusion)
    g_loss = adversarial_loss(prob, valid)
    res = g_loss.detach()
    g_loss.backward()
    opt.step()
    return res


class GANTrainer:
    """Stage 1 (reference ``train_GAN``, train_IEMOCAP.py:255-393): six networks, six Adam
    optimizers (generators lr, discriminators lr/2, text generator lr*1.1), BCE adversarial loss."""

    def __init__(self, acoustic_gen, visual_gen, text_gen, acoustic_disc, visual_disc, text_disc, lr=GAN_LR, b1=GAN_B1,
                 b2=GAN_B2, grad_reducer=None, world_size: int = 1, overlap: bool = True, batch_disc: bool = True,
                 chains: int = 2, freeze_disc: bool = True):
        self.nets = dict(acoustic_gen=acoustic_gen, visual_gen=visual_gen, text_gen=text_gen, acoustic_disc=acoustic_disc,
                         visual_disc=visual_disc, text_disc=text_disc)
        mk = lambda net, rate: FusedAdam(net, lr=rate, betas=(b1, b2), grad_reducer=grad_reducer)
        self.opt_acoustic_G = mk(acoustic_gen, lr)
        self.opt_acoustic_D = mk(acoustic_disc, lr / 2)
        self.opt_visual_G = mk(visual_gen, lr)
        self.opt_visual_D = mk(visual_disc, lr / 2)
        self.opt_text_G = mk(text_gen, lr * 1.1)
        self.opt_text_D = mk(text_disc, lr / 2)
        self.adversarial_loss = M.BCELoss()
        self.grad_reducer, self.world_size = grad_reducer, world_size
        # data parallel: all-reduce + Adam of a sub-step's network run on the communication stream while the caller's
        # stream starts the next sub-step (FusedAdam.defer_step); joined at the end of the batch
        for o in (self.opt_acoustic_G, self.opt_acoustic_D, self.opt_visual_G, self.opt_visual_D, self.opt_text_G, self.opt_text_D):
            o.defer_step = grad_reducer is not None
        # independent networks of a sub-step on concurrent streams (functional._Lanes); the loop bodies are unchanged
        self.overlap = overlap
        # train_disc as one discriminator pass over [real | fake] (train_disc_batched) instead of two
        self.batch_disc = batch_disc
        # train_gen without the discriminator's never-read weight gradients (train_gen_frozen_disc); False = reference body
        self.freeze_disc = freeze_disc
        # independent sub-steps on concurrent chains (see _batch); 1 = the reference's strictly serial order
        self.chains = chains
        self._chain_streams = {}

    def batch(self, data: Batch) -> Dict[str, torch.Tensor]:
        """The twelve sub-steps of one batch, in the reference's order (train_IEMOCAP.py:355-382).
        Returns the six surviving loss values as device scalars (later sub-steps overwrite earlier ones,
        as in the reference)."""
        with GF.overlap_networks(self.overlap):
            loss = self._batch(data)
        for o in (self.opt_acoustic_G, self.opt_acoustic_D, self.opt_visual_G, self.opt_visual_D, self.opt_text_G, self.opt_text_D):
            o.join_deferred()
        return loss

    # The twelve sub-steps of train_IEMOCAP.py:355-382 in the reference's order: (kind, discriminator, generator, loss key).
    SUBSTEPS = (("D", "visual_disc", "acoustic_gen", "visual_D_loss"), ("G", "visual_disc", "acoustic_gen", "acoustic_G_loss"),
                ("D", "visual_disc", "text_gen", "visual_D_loss"), ("G", "visual_disc", "text_gen", "text_G_loss"),
                ("D", "text_disc", "acoustic_gen", "text_D_loss"), ("G", "text_disc", "acoustic_gen", "acoustic_G_loss"),
                ("D", "acoustic_disc", "text_gen", "acoustic_D_loss"), ("G", "acoustic_disc", "text_gen", "text_G_loss"),
                ("D", "text_disc", "visual_gen", "text_D_loss"), ("G", "text_disc", "visual_gen", "visual_G_loss"),
                ("D", "acoustic_disc", "visual_gen", "acoustic_D_loss"), ("G", "acoustic_disc", "visual_gen", "visual_G_loss"))

    def _batch(self, data: Batch) -> Dict[str, torch.Tensor]:
        n = self.nets
        real = {"text_gen": data.text, "visual_gen": data.visual, "acoustic_gen": data.acoustic,
                "text_disc": data.text, "visual_disc": data.visual, "acoustic_disc": data.acoustic}
        opts = {"acoustic_gen": self.opt_acoustic_G, "visual_gen": self.opt_visual_G, "text_gen": self.opt_text_G,
                "acoustic_disc": self.opt_acoustic_D, "visual_disc": self.opt_visual_D, "text_disc": self.opt_text_D}
        real_text = data.text
        seq_len, batch_size = real_text.size(0), real_text.size(1)
        valid = torch.ones(seq_len, batch_size, 1, device=real_text.device)
        fake = torch.zeros(seq_len, batch_size, 1, device=real_text.device)
        adv = self.adversarial_loss
        # BCELoss is a mean over the *global* S*B slots: each rank contributes local_mean * (B_local / B_global)
        # (= 1/world_size for equal shards), so the summed shard gradients equal the single-device gradient.
        # The global batch size is all-reduced on the device for EVERY batch by EVERY rank (never cached by the local
        # size: with uneven last batches ranks would disagree on whether a collective is due and pair a scalar
        # all-reduce with a gradient all-reduce); the scale stays a device scalar, so the step is still capturable.
        if self.grad_reducer is not None:
            adv.scale_tensor = self.grad_reducer.local_fraction_tensor(batch_size, real_text.device)
        disc_step = train_disc_batched if self.batch_disc else train_disc
        for o in (self.opt_acoustic_D, self.opt_visual_D, self.opt_text_D):
            o.expected_backwards = 1 if self.batch_disc else 2     # the two-pass body adds D(real) and D(fake) into one arena
        if self.batch_disc:                                         # the network pass, then the separate `object` projection
            self.opt_visual_D.expected_backwards, self.opt_visual_D.sequential_backwards = 2, True
        else:
            self.opt_visual_D.sequential_backwards = False

        def run(kind, d, g):
            if kind == "D":
                return disc_step(n[d], real[d], n[g], real[g], opts[d], adv, valid, fake)
            gen_step = train_gen_frozen_disc if self.freeze_disc else train_gen
            return gen_step(n[g], real[g], n[d], opts[g], adv, valid, fake)

        loss = {}
        chains = self.chains if (self.overlap and not GF._deterministic["on"]) else 1
        if chains <= 1:
            for kind, d, g, key in self.SUBSTEPS:
                loss[key] = run(kind, d, g)
            return loss

        # ---- sub-step chains --------------------------------------------------------------------
        # A sub-step touches exactly two networks (reads / updates their weights, gradient arenas and Adam state), so it
        # depends only on the previous sub-steps that touched one of the two.  In the reference's order that leaves two
        # independent runs -- sub-steps 3,4,7,8 use {D_v, G_t, D_a} while 5,6,9,10 use {D_t, G_a, G_v} -- and each
        # sub-step is one chain of dependent 10-40 us kernels that leaves most SMs idle.  Each sub-step goes to the
        # chain (stream) whose tail is one of its dependencies, else to another chain, and waits for the events of the
        # dependencies that live elsewhere: every network still sees exactly the reference's sequence of operations.
        dev = real_text.device
        main = torch.cuda.current_stream(dev)
        pool = self._chain_streams.get(dev.index)
        if pool is None:
            pool = [torch.cuda.Stream(device=dev) for _ in range(chains)]
            for st in pool:
                GF.register_chain_stream(st)
            self._chain_streams[dev.index] = pool
        tail = [None] * len(pool)            # index of the last sub-step enqueued on each chain
        last = {}                            # network -> (sub-step index, chain, event)
        used = set()
        for idx, (kind, d, g, key) in enumerate(self.SUBSTEPS):
            deps = [last[x] for x in (d, g) if x in last]
            c = next((dc for (di, dc, _) in sorted(deps, reverse=True) if tail[dc] == di), None)
            if c is None:
                c = next((k for k in range(len(pool)) if tail[k] is None), None)
                if c is None:
                    c = deps[0][1] if deps else 0
            st = pool[c]
            if c not in used:
                st.wait_stream(main)         # fork: everything already enqueued by the caller (inputs, zeroed arenas)
                used.add(c)
            for (_, dc, ev) in deps:
                if dc != c:
                    st.wait_event(ev)
            with torch.cuda.stream(st):
                loss[key] = run(kind, d, g)
                ev = torch.cuda.Event()
                ev.record(st)
            tail[c] = idx
            last[d] = last[g] = (idx, c, ev)
        for c in used:
            main.wait_stream(pool[c])        # join
        for t in loss.values():
            t.record_stream(main)
        return loss


class ClassifierTrainer:
    """Stage 2 (reference ``train_or_eval_model``, train_IEMOCAP.py:103-197) for ``GAN_FFN``, and the same loop body of
    train_IEMOCAP_DialogueRNN.py for ``GAN_FFN_DialogueRNN`` (the model is then called with ``qmask`` and ``umask`` too)."""

    def __init__(self, model: M.GAN_FFN, loss_weights=None, lr=FFN_LR, l2=FFN_L2, grad_reducer=None, overlap: bool = True):
        self.model = model
        self.overlap = overlap
        self.loss_function = M.MaskedNLLLoss(loss_weights)
        self.optimizer = FusedAdam(model, lr=lr, weight_decay=l2, grad_reducer=grad_reducer)
        self.optimizer.defer_step = grad_reducer is not None
        self.grad_reducer = grad_reducer

    def step(self, data: Batch, train: bool = True):
        """One batch of the loop body (train_IEMOCAP.py:127-170).  Returns (loss, pred_, labels_)."""
        with GF.overlap_networks(self.overlap):
            out = self._step(data, train)
        self.optimizer.join_deferred()
        return out

    def _step(self, data: Batch, train: bool = True):
        model, optimizer = self.model, self.optimizer
        model.train() if train else model.eval()
        if train:
            optimizer.zero_grad()
        textf, visuf, acouf, umask, label = data.text, data.visual, data.acoustic, data.umask, data.label
        den = None
        prev_override = self.loss_function.den_override
        if self.grad_reducer is not None and train:
            # global denominator sum(w[label]*umask) so that summed shard gradients equal the single-device gradient
            # on the whole batch.  It stays on the device (one scalar all-reduce, no host read), so the
            # step can be replayed from a CUDA graph: the kernel computes the local numerator (denominator 1) and the
            # division by the global sum is a device-side scalar op.
            den = self.grad_reducer.global_nll_denominator_tensor(label, umask, self.loss_function.weight)
        try:
            # the override is scoped to this call: a later eval step must get the plain local mean again
            self.loss_function.den_override = 1.0 if den is not None else 0.0
            with torch.set_grad_enabled(train):
                if isinstance(model, M.GAN_FFN_DialogueRNN):
                    log_prob, alpha, alpha_f, alpha_b = model(acouf, visuf, textf, data.qmask, umask, max_len=max(data.lengths))
                else:
                    log_prob, alpha, alpha_f, alpha_b = model(acouf, visuf, textf)
                lp_ = log_prob.transpose(0, 1).contiguous().view(-1, log_prob.size()[2])
                labels_ = label.view(-1)
                loss = self.loss_function(lp_, labels_, umask)
                if den is not None:
                    loss = loss / den
        finally:
            self.loss_function.den_override = prev_override
        pred_ = torch.argmax(lp_, 1)
        if train:
            loss.backward()
            optimizer.step()
        logged = loss.detach()
        if den is not None:
            # each rank holds (local numerator / global denominator): the logged value is their sum = the loss of the
            # whole global batch, identical on every rank (one more scalar all-reduce, recorded with the step)
            logged = self.grad_reducer.sum_tensor(logged.clone())
        return logged, pred_, labels_


class GraphedTrainStep:
    """One train step (stage-1 GAN batch and/or stage-2 classifier step) as a replayed CUDA graph
    (the twelve sub-steps are ~5000 kernel launches; replay removes their launch overhead).

    The loop bodies are the same Python as the eager path -- they are simply *recorded* once per batch shape:
    the first call for a shape runs eagerly (it is a real training step and doubles as warm-up), the second call
    captures, every later call copies the batch into the graph's static input tensors and replays.  What makes the
    recording replayable: dropout seeds come from a ``DeviceSeedStream`` that the graph advances on the device, the
    Adam step count lives on the device, and the kernels never allocate or synchronise.

    Returns a dict of device tensors owned by the graph (read them before the next call): the six GAN losses and,
    with a classifier, ``loss`` / ``pred`` / ``labels``."""

    def __init__(self, gan: "GANTrainer" = None, cls: "ClassifierTrainer" = None, seed=None, enabled: bool = True):
        if gan is None and cls is None:
            raise ValueError("GraphedTrainStep needs a GANTrainer and/or a ClassifierTrainer")
        self.gan, self.cls, self.enabled = gan, cls, enabled
        self.seed = seed
        self.seeds = None
        self._seen, self._graphs = set(), {}
        self.kernels_per_replay = {}   # shape key -> kernels of this library recorded in the graph
        self.last_key = None
        # Data parallelism: the gradient all-reduces (NCCL) and the scalar all-reduce of the NLL denominator are device-
        # side and are recorded with the rest of the step; if a backend cannot be captured, the step falls back to eager.
        self.distributed = (gan is not None and gan.grad_reducer is not None) or (cls is not None and cls.grad_reducer is not None)

    def _body(self, batch: Batch):
        out = {}
        if self.gan is not None:
            out.update(self.gan.batch(batch))
        if self.cls is not None:
            loss, pred, labels = self.cls.step(batch, train=True)
            out.update(loss=loss, pred=pred, labels=labels)
        return out

    def release(self) -> None:
        """Drops the recorded graphs (they hold the captured NCCL kernels and the static batches).  Call before
        tearing the process group down: drop graphs -> synchronize -> destroy group."""
        self._graphs.clear()
        import gc
        gc.collect()
        torch.cuda.synchronize()

    def __call__(self, batch: Batch):
        from . import functional as GF
        dev = batch.text.device
        if self.seeds is None:
            self.seeds = GF.DeviceSeedStream(dev, base=self.seed)
        prev = GF.set_seed_stream(self.seeds)
        try:
            # max(lengths): the DialogueRNN head bakes the longest dialogue into the recording (BiModel._reverse_seq)
            key = (tuple(batch.text.shape), tuple(batch.visual.shape), dev.index, max(batch.lengths) if batch.lengths else 0)
            self.last_key = key
            if not self.enabled or key not in self._seen:
                self._seen.add(key)
                self.seeds.advance()
                return self._body(batch)
            if key not in self._graphs:
                static = Batch(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in vars(batch).items()})
                torch.cuda.synchronize(dev)
                from ._lib import lib
                n0 = int(lib().cdll.ganffn_launch_count())
                graph = torch.cuda.CUDAGraph()
                try:
                    with torch.cuda.graph(graph):
                        self.seeds.advance()
                        out = self._body(static)
                except RuntimeError as exc:
                    # Only a failed *capture* of the collectives is survivable (a backend whose all-reduce cannot be
                    # recorded): anything else, and any failure without data parallelism, is a real bug and propagates.
                    msg = str(exc).lower()
                    if not self.distributed or not any(k in msg for k in ("captur", "nccl", "graph")):
                        raise
                    import warnings
                    warnings.warn(f"GraphedTrainStep: the data-parallel step could not be captured ({exc}); "
                                  "running eagerly from now on", RuntimeWarning)
                    self.enabled = False
                    torch.cuda.synchronize(dev)
                    self.seeds.advance()
                    return self._body(batch)
                self.kernels_per_replay[key] = int(lib().cdll.ganffn_launch_count()) - n0
                self._graphs[key] = (graph, static, out)
            graph, static, out = self._graphs[key]
            for k, v in vars(batch).items():
                if torch.is_tensor(v):
                    getattr(static, k).copy_(v, non_blocking=True)
            graph.replay()
            return out
        finally:
            GF.set_seed_stream(prev)


def build_networks(D_h: int = 100, n_classes: int = 6, device="cuda", seed: int = 3407):
    """The six networks + GAN_FFN as the reference's ``__main__`` builds them
    (train_IEMOCAP.py:580-585, :629-635: GAN dropout 0.2, GAN_FFN dropout --dropout 0.6)."""
    torch.manual_seed(seed)
    nets = dict(acoustic_gen=M.AcousticGenerator(D_h, dropout=0.2), acoustic_disc=M.AcousticDiscriminator(D_h, dropout=0.2),
                visual_gen=M.VisualGenerator(D_h, dropout=0.2), visual_disc=M.VisualDiscriminator(D_h, dropout=0.2),
                text_gen=M.TextGenerator(D_h, dropout=0.2), text_disc=M.TextDiscriminator(D_h, dropout=0.2))
    ffn = M.GAN_FFN(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], n_classes=n_classes, dropout=0.6)
    ffn.to(device)
    for k in ("acoustic_disc", "visual_disc", "text_disc"):
        nets[k].to(device)
    return nets, ffn
