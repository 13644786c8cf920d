#!/usr/bin/env python
"""Inference-scoring sweeps sharded by dialogue (BASELINE.json configs 3 and 4).  Not the headline bench (bench.py is
the train step); prints one JSON line on rank 0.

  python tools/score_sweep.py --utterances 1000000                      1M-utterance sweep, lengths ~ U{10..110}
  python tools/score_sweep.py --meld --utterances 200000                MELD-shaped: 7 classes, S <= 33
  python -m torch.distributed.run --nproc-per-node N ... tools/score_sweep.py --utterances 1000000

Batches of 32 dialogues, sorted by length (stated in the output: the pad length changes both the work and the
result), dealt round-robin to the ranks; no collective on the data path.  Timed end to end per rank:
pinned host batch -> device (copy stream, one batch ahead) -> GAN_FFN eval forward -> predictions back to the host;
CUDA events around the whole sweep, max over ranks."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from gan_ffn_b200 import parallel, scoring, synthetic, train  # noqa: E402
from gan_ffn_b200._lib import lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utterances", type=int, default=1_000_000)
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--meld", action="store_true", help="MELD-shaped: 7 classes, dialogues of 1..33 turns")
    ap.add_argument("--unsorted", action="store_true", help="keep loader order instead of grouping by length")
    ap.add_argument("--eager", action="store_true", help="launch kernels eagerly instead of replaying one CUDA graph per batch shape")
    args = ap.parse_args()
    rank, local_rank, world = parallel.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n_classes, lo, hi = (7, 1, 33) if args.meld else (6, 10, 110)
    mean_len = (lo + hi) / 2
    n_dialogues = int(round(args.utterances / mean_len))
    lengths = synthetic.ragged_lengths(n_dialogues, lo, hi, seed=11)
    corpus = scoring.SyntheticDialogues(lengths, n_classes=n_classes)
    plan = scoring.plan_batches(lengths, args.batch_size, sort=not args.unsorted)
    mine = scoring.shard_batches(plan, world, rank)
    nets, ffn = train.build_networks(n_classes=n_classes, device=dev)

    # Host batches: generating 1M utterances x 712 features in Python would dominate the run, so each rank builds one
    # pinned host batch per distinct (pad length, batch size) of its plan (<= ~100 shapes when sorted by length) and
    # every planned batch copies the entry of exactly its own shape from the host: the work is the plan's work.
    pool = {}
    for idx in mine:
        key = (max(lengths[i] for i in idx), len(idx))
        if key not in pool:
            pool[key] = corpus.batch(idx).pin()
    keys = sorted(pool)

    def host_batch(idx):
        S, B = max(lengths[i] for i in idx), len(idx)
        return pool[(S, B)], S * B, sum(lengths[i] for i in idx)

    copy_stream = torch.cuda.Stream()
    L = lib()
    scorer = (lambda b: scoring.score_batch(ffn, b)) if args.eager else scoring.GraphedScorer(ffn)
    # warm-up: every distinct shape (workspace growth, lazy module state; graph mode: eager call, then the recording)
    for key in keys:
        for _ in range(1 if args.eager else 3):
            scorer(pool[key].to(dev))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    L.cdll.ganffn_reset_launch_count()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    slots = real = h2d = d2h = 0
    t0 = time.perf_counter()
    a.record()
    nxt = None
    with torch.cuda.stream(copy_stream):
        hb, s_, r_ = host_batch(mine[0])
        nxt = (hb.to(dev, non_blocking=True), s_, r_, hb.h2d_bytes())
    preds = []
    for k in range(len(mine)):
        torch.cuda.current_stream().wait_stream(copy_stream)
        cur, s_, r_, nb = nxt
        slots += s_; real += r_; h2d += nb
        if k + 1 < len(mine):
            with torch.cuda.stream(copy_stream):
                hb, s2, r2 = host_batch(mine[k + 1])
                nxt = (hb.to(dev, non_blocking=True), s2, r2, hb.h2d_bytes())
        out = scorer(cur)
        p = out["pred"].to("cpu", non_blocking=True)
        d2h += p.numel() * p.element_size()
        preds.append(p)
        if len(preds) > 8:
            preds.pop(0)
    e.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = a.elapsed_time(e)
    t = torch.tensor([ms, wall * 1e3, float(slots), float(real), float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    else:
        tmax = tsum = t
    if rank == 0:
        ms_max, wall_max = float(tmax[0]), float(tmax[1])
        line = {"metric": "gan_ffn_scoring_utterances_per_sec", "unit": "utterances/s", "n_gpus": world,
                "value": float(tsum[3]) / (ms_max / 1e3), "padded_value": float(tsum[2]) / (ms_max / 1e3),
                "wall_value": float(tsum[3]) / (wall_max / 1e3), "ms_total": ms_max, "higher_is_better": True,
                "scaling": "strong", "dtype": "f32", "data": "synthetic",
                "config": {"workload": ("meld_shaped_scoring: 7 classes, dialogues of 1..33 turns, modality widths 100/512/100 "
                                        "(the reference's GAN-FFN cannot take MELD's 600-d text, SURVEY.md D4)") if args.meld else
                           "sweep: dialogues of 10..110 turns, 6 classes",
                           "utterances": float(tsum[3]), "padded_slots": float(tsum[2]), "dialogues": n_dialogues,
                           "batches": len(plan), "batch_size": args.batch_size,
                           "batching": "loader order" if args.unsorted else "sorted by length (ascending), whole batches dealt round-robin to ranks",
                           "parallelism": f"dp{world} by dialogue, no collective on the data path",
                           "host_data": f"{len(pool)} pinned host batches per rank (one per distinct pad length x batch size); every planned batch is copied from the host",
                           "launch": "eager kernel launches" if args.eager else "one CUDA graph per batch shape, replayed",
                           "timing": "CUDA events around the whole sweep incl. host->device and device->host copies, max over ranks"},
                "h2d_bytes": float(tsum[4]), "d2h_bytes": float(tsum[5]), "gpu_launches": int(L.cdll.ganffn_launch_count())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
