"""GPU, >= 2 devices, NCCL: dialogue-sharded training step through the CUDA path equals the single-GPU step on
the whole global batch (gradients after the all-reduce; losses).  Skipped on a 1-GPU box."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    import gan_ffn_b200 as GB
    from gan_ffn_b200 import parallel, synthetic, train
    parallel.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    red = parallel.GradReducer()
    w = torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device=dev)
    glob = synthetic.make_batch(n_dialogues=6, lengths=[40, 17, 33, 8, 25, 40], seed=9)

    def run(batch, reducer):
        nets, ffn = train.build_networks(device=dev)
        for m in list(nets.values()) + [ffn]:
            m.eval()                                           # dropout off: the comparison must be deterministic
        b = batch.to(dev)
        # stage 2
        lossf = GB.MaskedNLLLoss(w)
        if reducer is not None:
            lossf.den_override = reducer.global_nll_denominator(b.label, b.umask, w)
        ffn.zero_grad(set_to_none=True)
        lp = ffn(b.acoustic, b.visual, b.text)[0]
        loss2 = lossf(lp.transpose(0, 1).contiguous().view(-1, 6), b.label.view(-1), b.umask)
        loss2.backward()
        # stage 1, one discriminator pass
        bce = GB.BCELoss()
        if reducer is not None:
            bce.scale = b.n_dialogues / reducer.global_sum(b.n_dialogues, dev)
        d = nets["visual_disc"]
        d.zero_grad(set_to_none=True)
        prob = d(b.visual)
        loss1 = bce(prob, torch.ones_like(prob))
        loss1.backward()
        bufs = [m.arena().grad for m in (nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], d)]
        bufs += [ffn.fc.weight.grad, ffn.fc.bias.grad]
        losses = torch.stack([loss2.detach(), loss1.detach()])
        if reducer is not None:
            reducer.reduce(bufs)
            dist.all_reduce(losses)
        return [t.clone() for t in bufs], losses

    g_dp, l_dp = run(parallel.shard_batch(glob, world, rank), red)
    g_one, l_one = run(glob, None)
    ok = {"loss": bool(torch.allclose(l_dp, l_one, rtol=1e-4, atol=0))}
    for i, (a, e) in enumerate(zip(g_dp, g_one)):
        ok[f"grad{i}"] = float((a - e).abs().max()) <= 1e-4 * float(e.abs().max()) + 1e-12
    if rank == 0:
        torch.save(ok, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_step_equals_single_gpu_step(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "ok.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ok = torch.load(out)
    assert all(ok.values()), ok


def _trainer_worker(rank, world, port, out_path):
    """ClassifierTrainer under data parallelism, replayed from a CUDA graph (NCCL all-reduces and the device-resident
    NLL denominator recorded with the step) against the single-GPU eager trainer on the whole global batch."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    from gan_ffn_b200 import parallel, synthetic, train
    parallel.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    red = parallel.GradReducer()
    w = torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device=dev)
    glob = synthetic.make_batch(n_dialogues=6, lengths=[40, 17, 33, 8, 25, 40], seed=9)

    def run(batch, reducer, graphed):
        nets, ffn = train.build_networks(device=dev)
        ffn.eval()
        ffn.train = lambda mode=True: ffn          # keep dropout off: the comparison must be deterministic
        cls = train.ClassifierTrainer(ffn, w, grad_reducer=reducer)
        stepper = train.GraphedTrainStep(None, cls, seed=5, enabled=graphed)
        b = batch.to(dev)
        losses = []
        for _ in range(4):                          # graphed: eager, record, replay, replay
            losses.append(stepper(b)["loss"].detach().clone())
        torch.cuda.synchronize()
        flat = torch.cat([p.detach().reshape(-1) for p in ffn.parameters()])
        captured = bool(stepper.kernels_per_replay)
        stepper.release()
        return torch.stack(losses), flat, captured

    l_dp, w_dp, captured = run(parallel.shard_batch(glob, world, rank), red, True)   # the trainer logs the global loss
    l_one, w_one, _ = run(glob, None, False)
    ok = {"captured_under_dp": captured,
          "loss_step0": bool(torch.allclose(l_dp[0], l_one[0], rtol=1e-4, atol=0)),
          "loss_later": bool(torch.allclose(l_dp, l_one, rtol=2e-3, atol=0)),
          "weights": float((w_dp - w_one).abs().mean()) < 0.1 * 1e-4}
    if rank == 0:
        torch.save(ok, out_path)
    if not parallel.shutdown():      # graphs released above -> drain -> barrier -> destroy (watchdog inside)
        os._exit(3)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_trainer_graph_replay_equals_single_gpu_trainer(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "ok.pt")
    ctx = mp.spawn(_trainer_worker, args=(2, _free_port(), out), nprocs=2, join=False)
    for p in ctx.processes:
        p.join(timeout=300)
        assert p.exitcode == 0, f"worker exit code {p.exitcode}"
    ok = torch.load(out)
    assert all(ok.values()), ok


def _full_trainers_worker(rank, world, port, out_path):
    """The real trainers (GANTrainer with its six FusedAdam optimizers + ClassifierTrainer) under a GradReducer,
    with UNEVEN shards and a batch size that changes from one batch to the next (5 dialogues -> 3+2, then 3 -> 2+1),
    against the single-GPU trainers on the whole global batches.  Dropout off (modules pinned in eval mode)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    from gan_ffn_b200 import parallel, synthetic, train
    parallel.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    red = parallel.GradReducer()
    w = torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device=dev)
    batches = [synthetic.make_batch(n_dialogues=5, lengths=[30, 17, 22, 8, 25], seed=9),
               synthetic.make_batch(n_dialogues=3, lengths=[12, 30, 21], seed=10)]
    keys = ["acoustic_D_loss", "acoustic_G_loss", "text_D_loss", "text_G_loss", "visual_D_loss", "visual_G_loss"]

    def run(reducer):
        nets, ffn = train.build_networks(device=dev)
        for m in list(nets.values()) + [ffn]:
            m.eval()
            m.train = (lambda mod: (lambda mode=True: mod))(m)
        gan = train.GANTrainer(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], nets["acoustic_disc"],
                               nets["visual_disc"], nets["text_disc"], grad_reducer=reducer, world_size=world)
        cls = train.ClassifierTrainer(ffn, w, grad_reducer=reducer)
        out = []
        for gb in batches:
            b = (parallel.shard_batch(gb, world, rank) if reducer is not None else gb).to(dev)
            losses = gan.batch(b)
            vals = torch.stack([losses[k] for k in keys])
            if reducer is not None:
                dist.all_reduce(vals)           # BCE: every rank holds local_mean * B_local / B_global
            l2, _, _ = cls.step(b, train=True)  # already the global loss on every rank
            out.append(torch.cat([vals, l2.reshape(1)]))
        torch.cuda.synchronize()
        flat = torch.cat([p.detach().reshape(-1) for m in nets.values() for p in m.parameters()])
        return torch.stack(out), flat

    l_dp, w_dp = run(red)
    l_one, w_one = run(None)
    ok = {"losses_batch0": bool(torch.allclose(l_dp[0], l_one[0], rtol=1e-4, atol=0)),
          "losses_batch1": bool(torch.allclose(l_dp[1], l_one[1], rtol=5e-4, atol=0)),
          "weights_mean": float((w_dp - w_one).abs().mean()) < 0.02 * 5e-5,
          "rel_loss_err": float(((l_dp - l_one).abs() / l_one.abs()).max())}
    if rank == 0:
        torch.save(ok, out_path)
    if not parallel.shutdown():
        os._exit(3)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_real_trainers_with_uneven_changing_shards_equal_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "ok.pt")
    ctx = mp.spawn(_full_trainers_worker, args=(2, _free_port(), out), nprocs=2, join=False)
    for p in ctx.processes:
        p.join(timeout=600)
        assert p.exitcode == 0, f"worker exit code {p.exitcode}"
    ok = torch.load(out)
    print("PARITY dp-trainers", ok)
    assert all(v for k, v in ok.items() if k != "rel_loss_err"), ok
