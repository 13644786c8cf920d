#!/usr/bin/env python
"""Fixed-size TRAINING sweep, data-parallel by dialogue (BASELINE.json config 4: "synthetic 1M-utterance sweep, dialogues
up to 110 turns, sharded by dialogue across 2/4/8 B200 with NCCL grad allreduce").  Strong scaling: the corpus is
fixed, every step trains on one global batch of 32 x world dialogues (32 per GPU), so 1 GPU runs ~521 steps and 8 GPUs
~65.  A step is the whole hot path: the stage-1 GAN batch (12 sub-steps) + the stage-2 classifier step, train mode,
dropout on, Adam, one NCCL all-reduce per parameter arena per optimizer step.  Not the headline bench (bench.py is the
fixed-shape weak-scaling step); prints one JSON line on rank 0.

  python tools/train_sweep.py --utterances 1000000
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_sweep.py --utterances 1000000

Batching policy (part of the result, SURVEY.md §0: a dialogue's output depends on its pad length): dialogues sorted by
length, global batches of 32 x world consecutive dialogues, each padded to the next multiple of `--bucket` turns (10) so
that the step is replayed from at most 11 recorded CUDA graphs; every rank pads to the GLOBAL batch's length.  The last,
partial global batch is dropped.  Host data: one pinned host batch per distinct (pad length, dialogues) shape per rank
(generating 1M x 712 features in Python would dominate the run); every step copies its batch host -> device and reads
the seven losses back.  Timed with CUDA events around the whole sweep, max over ranks."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import gan_ffn_b200 as G  # noqa: E402
from gan_ffn_b200 import parallel, scoring, synthetic, train  # noqa: E402
from gan_ffn_b200._lib import lib  # noqa: E402
from gan_ffn_b200.synthetic import Batch  # noqa: E402


def pad_to(b: Batch, S: int) -> Batch:
    """Zero-pads a batch along the sequence axis to S turns (collate semantics, dataloader.py:55-58)."""
    if b.seq_len == S:
        return b
    add = S - b.seq_len
    f3 = lambda t: torch.cat([t, t.new_zeros((add,) + tuple(t.shape[1:]))], dim=0)
    f2 = lambda t: torch.cat([t, t.new_zeros((t.shape[0], add))], dim=1)
    return Batch(f3(b.text), f3(b.visual), f3(b.acoustic), f3(b.qmask), f2(b.umask), f2(b.label), b.lengths)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utterances", type=int, default=1_000_000)
    ap.add_argument("--per-gpu", type=int, default=32, help="dialogues per GPU per step")
    ap.add_argument("--bucket", type=int, default=10, help="pad lengths are rounded up to a multiple of this many turns")
    ap.add_argument("--max-steps", type=int, default=0, help="time only the first K steps of the plan (0 = the whole corpus)")
    args = ap.parse_args()
    rank, local_rank, world = parallel.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lo, hi = 10, 110
    n_dialogues = int(round(args.utterances / ((lo + hi) / 2)))
    lengths = synthetic.ragged_lengths(n_dialogues, lo, hi, seed=11)
    corpus = scoring.SyntheticDialogues(lengths)
    gbs = args.per_gpu * world
    plan = [b for b in scoring.plan_batches(lengths, gbs, sort=True) if len(b) == gbs]
    if args.max_steps:
        step_ids = [int(round(k * (len(plan) - 1) / max(args.max_steps - 1, 1))) for k in range(args.max_steps)]   # spread over all lengths
        plan = [plan[i] for i in sorted(set(step_ids))]
    bucket = lambda s: min(hi, (s + args.bucket - 1) // args.bucket * args.bucket)

    reducer = parallel.GradReducer() if world > 1 else None
    nets, ffn = train.build_networks(device=dev)
    gan = train.GANTrainer(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], nets["acoustic_disc"],
                           nets["visual_disc"], nets["text_disc"], grad_reducer=reducer, world_size=world)
    cls = train.ClassifierTrainer(ffn, torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device=dev), grad_reducer=reducer)
    G.manual_seed(3407 + rank)
    stepper = train.GraphedTrainStep(gan, cls, seed=3407 + rank)

    steps = []            # (pad length, my dialogue indices, real utterances of the GLOBAL batch)
    pool = {}
    for gb in plan:
        S = bucket(max(lengths[i] for i in gb))
        mine = [gb[j] for j in parallel.shard_indices(len(gb), world, rank)]
        steps.append((S, mine, sum(lengths[i] for i in gb)))
        key = (S, len(mine))
        if key not in pool:
            pool[key] = pad_to(corpus.batch(mine), S).pin()
    keys = sorted(pool)
    LOSS_KEYS = ["acoustic_D_loss", "acoustic_G_loss", "text_D_loss", "text_G_loss", "visual_D_loss", "visual_G_loss", "loss"]

    for key in keys:                                  # warm-up: eager call, recording, one replay per shape
        for _ in range(3):
            stepper(pool[key].to(dev))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    L = lib()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    copy_stream = torch.cuda.Stream()
    t0 = time.perf_counter()
    a.record()
    with torch.cuda.stream(copy_stream):
        nxt = pool[(steps[0][0], len(steps[0][1]))].to(dev, non_blocking=True)
    real = slots = h2d = 0
    last = None
    for k, (S, mine, real_global) in enumerate(steps):
        torch.cuda.current_stream().wait_stream(copy_stream)
        cur = nxt
        for t_ in (cur.text, cur.visual, cur.acoustic, cur.qmask, cur.umask, cur.label):
            t_.record_stream(torch.cuda.current_stream())
        if k + 1 < len(steps):
            with torch.cuda.stream(copy_stream):       # double-buffered host -> device copy, one batch ahead
                nxt = pool[(steps[k + 1][0], len(steps[k + 1][1]))].to(dev, non_blocking=True)
        out = stepper(cur)
        last = torch.stack([out[k_] for k_ in LOSS_KEYS]).to("cpu", non_blocking=True)
        real += real_global
        slots += S * gbs
        h2d += pool[(S, len(mine))].h2d_bytes()
    e.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([a.elapsed_time(e), wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        line = {"metric": "gan_ffn_train_sweep_utterances_per_sec", "unit": "utterances/s", "n_gpus": world,
                "value": real / (ms / 1e3), "padded_value": slots / (ms / 1e3), "wall_value": real / (float(t[1]) / 1e3),
                "ms_total": ms, "steps": len(steps), "ms_per_step": ms / len(steps), "higher_is_better": True, "scaling": "strong",
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "train sweep: stage-1 GAN batch (12 sub-steps) + stage-2 classifier step per global batch, "
                                       "dialogues of 10..110 turns, 6 classes, dropout on",
                           "utterances": real, "padded_slots": slots, "dialogues": len(steps) * gbs, "global_batch": gbs,
                           "per_gpu_batch": args.per_gpu, "graphs": len(keys),
                           "batching": f"sorted by length, consecutive global batches, pad length rounded up to a multiple of {args.bucket}, last partial batch dropped",
                           "parallelism": f"dp{world} by dialogue, NCCL all-reduce of the gradient arenas per optimizer step",
                           "timing": "CUDA events around the whole sweep incl. the host->device copy of every batch and the loss read-back, max over ranks"},
                "h2d_bytes_per_rank": h2d, "final_losses": [float(x) for x in last],
                "gpu_launches_per_replay": sorted(set(stepper.kernels_per_replay.values()))}
        print(json.dumps(line), flush=True)
    if world > 1 and not parallel.shutdown([stepper]):
        os._exit(0)


if __name__ == "__main__":
    main()
