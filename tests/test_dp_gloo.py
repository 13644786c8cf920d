"""CPU, world_size 2, gloo: the dialogue-sharded data-parallel plumbing (gan_ffn_b200/parallel.py).

The arithmetic here is the CPU oracle (tests may use it); what is under test is the host logic: shards keep
the global pad length, the loss scalings of SURVEY.md §8e, and one sum-all-reduce per flat gradient buffer --
together they must reproduce the single-process gradient on the whole global batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flat(grads):
    return torch.cat([g.reshape(-1) for g in grads])


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    import helpers as H
    from helpers import O
    from gan_ffn_b200 import parallel, synthetic
    rank_, _, world_ = parallel.init_from_env("gloo")
    assert (rank_, world_) == (rank, world)
    red = parallel.GradReducer()

    torch.manual_seed(H.SEED)
    from gan_ffn_b200 import model as M
    gen, disc = M.TextGenerator(100), M.TextDiscriminator(100)          # parameter containers only (CPU)
    fc = torch.nn.Linear(100, 6)
    glob = synthetic.make_batch(n_dialogues=5, lengths=[12, 7, 10, 3, 9], seed=5)   # 5 dialogues over 2 ranks: 3 + 2
    mine = parallel.shard_batch(glob, world, rank)
    assert mine.seq_len == glob.seq_len == 12, "shards must keep the global pad length"
    w = torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS)

    def stage2_loss(b, den):
        Pg = O.params_of(gen, requires_grad=True)
        fw = fc.weight.detach().clone().requires_grad_(True)
        lp = torch.log_softmax(O.generator(b.text, Pg, 10) @ fw.T + fc.bias.detach(), dim=2)
        lp_ = lp.transpose(0, 1).contiguous().view(-1, 6)
        m = b.umask.reshape(-1)
        num = -(w[b.label.view(-1)] * m * lp_.gather(1, b.label.view(-1, 1)).squeeze(1)).sum()
        return num / den, [v for v in Pg.values() if v.requires_grad] + [fw]

    def bce_loss(b, scale):
        Pd = O.params_of(disc, requires_grad=True)
        prob = O.discriminator(b.text, Pd)
        return O.bce(prob, torch.ones_like(prob)) * scale, [v for v in Pd.values() if v.requires_grad]

    # ---- sharded: local losses with global normalisers, then one all-reduce per flat gradient buffer ----------
    den = red.global_nll_denominator(mine.label, mine.umask, w)
    loss2, params2 = stage2_loss(mine, den)
    g2 = _flat(torch.autograd.grad(loss2, params2, allow_unused=True, materialize_grads=True))
    scale = mine.n_dialogues / red.global_sum(mine.n_dialogues, "cpu")
    loss1, params1 = bce_loss(mine, scale)
    g1 = _flat(torch.autograd.grad(loss1, params1, allow_unused=True, materialize_grads=True))
    red.reduce([g2, g1])
    losses = torch.stack([loss2.detach(), loss1.detach()])
    dist.all_reduce(losses)

    # ---- single process on the whole global batch ---------------------------------------------------------------
    den_ref = float((w[glob.label.view(-1)] * glob.umask.reshape(-1)).sum())
    ref2, p2 = stage2_loss(glob, den_ref)
    r2 = _flat(torch.autograd.grad(ref2, p2, allow_unused=True, materialize_grads=True))
    ref1, p1 = bce_loss(glob, 1.0)
    r1 = _flat(torch.autograd.grad(ref1, p1, allow_unused=True, materialize_grads=True))

    ok = {
        "den": abs(den - den_ref) <= 1e-5 * den_ref,
        "loss2": abs(losses[0].item() - ref2.item()) <= 1e-5 * abs(ref2.item()),
        "loss1": abs(losses[1].item() - ref1.item()) <= 1e-5 * abs(ref1.item()),
        "g2": float((g2 - r2).abs().max()) <= 1e-4 * float(r2.abs().max()),
        "g1": float((g1 - r1).abs().max()) <= 1e-4 * float(r1.abs().max()),
        "calls": red.calls == 2 and red.bytes_reduced == 4 * (g2.numel() + g1.numel()),
    }
    if rank == 0:
        torch.save(ok, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_dialogue_sharded_gradients_equal_single_process(tmp_path):
    out = str(tmp_path / "ok.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ok = torch.load(out)
    assert all(ok.values()), ok


def test_grad_reducer_requires_process_group():
    from gan_ffn_b200 import parallel
    if not dist.is_initialized():
        with pytest.raises(RuntimeError, match="one process per GPU"):
            parallel.GradReducer()
