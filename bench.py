#!/usr/bin/env python
"""bench.py -- GAN-FFN train-step throughput (utterances/s) on 1..8 B200.

A *step* is one pass of the whole hot path over one IEMOCAP-shaped synthetic batch per GPU
(S=94 turns x B=32 dialogues, text/visual/acoustic 100/512/100, 6 classes, train mode, dropout on):
  stage 1  the twelve adversarial sub-steps of reference train_IEMOCAP.py:355-382
           (6x train_disc + 6x train_gen: generator and discriminator fwd/bwd, BCE, Adam), then
  stage 2  one classifier step of train_IEMOCAP.py:127-170 (GAN_FFN fwd, MaskedNLLLoss, bwd, Adam).
Unit of work: padded utterance slots S*B (all are computed and, in stage 1, all enter the loss).

  python bench.py --gpus N --steps K --warmup W            our arm (torchrun for N>1)
  python bench.py --impl reference ...                     the reference's CPU path (oracle port) on host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

S_IEMOCAP, B_IEMOCAP = 94, 32
METRIC = "gan_ffn_train_step_padded_utterances_per_sec"
UNIT = "utterances/s"
WORKLOAD = "iemocap_train_step: stage-1 GAN batch (12 sub-steps) + stage-2 classifier step, S=94 B=32 per GPU, fp32, dropout on"


def flops_per_slot(S):
    """Algorithmic FLOP per padded slot: stage-1 batch + stage-2 step."""
    E = lambda d: 8 * d * d + 4 * S * d + 8192 * d
    G = lambda d, h: 8 * E(d) + 2 * d * h + 200 * h
    D = 8 * E(100) + 14880
    ffn_fwd = 2 * G(100, 512) + G(512, 1024) + 1200
    stage2 = 3 * ffn_fwd
    g100, g512 = G(100, 512), G(512, 1024)
    # train_disc: D(real) + G fwd (eval) + D(fake) fwd, bwd through both D passes (2x fwd each)
    # train_gen: G fwd + D fwd, bwd through both
    def disc_step(gf, real_extra):
        return (2 * D + real_extra + gf) + 2 * (2 * D + real_extra)
    def gen_step(gf):
        return 3 * (gf + D)
    obj = 102400
    stage1 = (disc_step(g100, obj) + gen_step(g100)) * 2 + (disc_step(g100, 0) + gen_step(g100)) * 2 \
        + (disc_step(g512, 0) + gen_step(g512)) * 2
    return stage1, stage2


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def _port_trainer(device="cpu"):
    import torch
    from gan_ffn_b200 import synthetic
    from oracle import reference_port as RP
    nets, ffn = RP.build()
    if device != "cpu":
        for m in list(nets.values()) + [ffn]:
            m.to(device)
    return RP.PortTrainer(nets, ffn, torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device=device))


def _pin_eval(tr):
    """Dropout off: modules pinned in eval mode (the loop bodies' .train() calls become no-ops)."""
    for m in list(tr.nets.values()) + [tr.ffn]:
        m.eval()
        m.train = (lambda mod: (lambda mode=True: mod))(m)


def run_reference(args):
    """The reference's CPU implementation of the path (oracle/reference_port.py: stock torch modules, the
    reference's own operators and loop bodies) on the box's host cores, on the SAME config as our arm: all
    `--dialogues` (32) dialogues at S=94 per step, dropout on, `--warmup` warm-up steps and `--steps` timed steps.
    (~10 s per step on 16 cores: 25 steps are ~4 minutes.)  Safety valve only: if the first warm-up step projects the
    whole run beyond 20 minutes, fewer steps are timed and the line says so."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gan_ffn_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr = _port_trainer()
    nd = args.ref_dialogues or args.dialogues
    b = synthetic.make_batch(n_dialogues=nd, seq_len=args.seq_len)
    def step():
        tr.gan_batch(b)
        tr.classifier_step(b)
    warm, steps = max(args.warmup, 1), args.steps
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    if first * (warm + steps) > 1200.0:
        steps = max(1, int(1200.0 / first) - warm) if int(1200.0 / first) > warm else 1
        warm = min(warm, max(1, int(1200.0 / first) - steps))
    for _ in range(warm - 1):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    # 59 % of the CPU train time is dropout RNG -- also time the same step with dropout off
    tr_off = _port_trainer()
    _pin_eval(tr_off)
    tr_off.gan_batch(b); tr_off.classifier_step(b)
    t1 = time.perf_counter()
    tr_off.gan_batch(b); tr_off.classifier_step(b)
    dt_off = time.perf_counter() - t1
    slots = b.padded_slots
    val = slots * steps / dt
    sample = (f"all {nd} dialogues at S={args.seq_len} ({slots} padded slots/step), {warm} warm-up + {steps} timed steps, dropout on"
              + ("" if steps == args.steps else f" (asked for {args.steps} steps: cut to stay within 20 minutes)"))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if (args.seq_len, nd) == (S_IEMOCAP, B_IEMOCAP) else f"train_step S={args.seq_len} B={nd}",
                       "seq_len": args.seq_len, "dialogues_per_gpu": nd, "sample": sample, "torch": torch.__version__,
                       "ms_per_step_dropout_off": 1e3 * dt_off, "value_dropout_off": slots / dt_off,
                       "dropout_off_note": "one timed step after one warm-up, modules pinned in eval mode (BASELINE.md §3: the CPU train step is dominated by dropout RNG)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline():
    """The oracle port timed on the host cores on a bounded sample of the same workload (rank 0, N=1 only): one
    warm-up and one timed step of the full 32-dialogue batch, dropout on (~10 s each on 16 cores), plus one
    dropout-off step."""
    import torch
    from gan_ffn_b200 import synthetic
    cores = os.cpu_count() or 1
    prev = torch.get_num_threads()
    torch.set_num_threads(cores)
    try:
        tr = _port_trainer()
        b = synthetic.make_batch(n_dialogues=B_IEMOCAP, seq_len=S_IEMOCAP)
        small = synthetic.make_batch(n_dialogues=2, seq_len=S_IEMOCAP)
        tr.gan_batch(small); tr.classifier_step(small)           # warm-up (thread pools, allocator) on two dialogues
        t0 = time.perf_counter()
        tr.gan_batch(b); tr.classifier_step(b)
        dt = time.perf_counter() - t0
        _pin_eval(tr)
        t1 = time.perf_counter()
        tr.gan_batch(b); tr.classifier_step(b)
        dt_off = time.perf_counter() - t1
        return {"value": b.padded_slots / dt, "unit": UNIT, "cores": cores, "kind": "port",
                "value_dropout_off": b.padded_slots / dt_off,
                "sample": f"1 step of all {B_IEMOCAP} dialogues at S={S_IEMOCAP} ({b.padded_slots} slots/step) after a 2-dialogue warm-up, "
                          f"dropout on; value_dropout_off = 1 more step with dropout off; torch {torch.__version__} CPU"}
    finally:
        torch.set_num_threads(prev)


def gpu_eager_baseline(dev, steps=3):
    """SURVEY.md §2a's comparator "the reference on the box": the stock-torch port (nn.TransformerEncoder, cuBLAS,
    torch SDPA, torch.optim.Adam; the reference's loop bodies) in eager mode on the same B200, same batch, dropout
    on -- once with fp32 matmuls (allow_tf32=False, the parity-equivalent setting) and once with TF32 allowed."""
    import torch
    from gan_ffn_b200 import synthetic
    out = {}
    b = synthetic.make_batch(n_dialogues=B_IEMOCAP, seq_len=S_IEMOCAP).to(dev)
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for tag, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            tr = _port_trainer(dev)
            for _ in range(2):
                tr.gan_batch(b); tr.classifier_step(b)
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            ms1 = ms2 = 0.0
            for _ in range(steps):
                e0.record(); tr.gan_batch(b); e1.record(); tr.classifier_step(b); e2.record()
                torch.cuda.synchronize()
                ms1 += e0.elapsed_time(e1); ms2 += e1.elapsed_time(e2)
            ms = (ms1 + ms2) / steps
            out[tag] = {"ms_per_step": ms, "stage1_ms": ms1 / steps, "stage2_ms": ms2 / steps, "value": b.padded_slots / (ms / 1e3), "unit": UNIT}
            del tr
        out["note"] = ("stock torch eager (oracle/reference_port.py on cuda): cuBLAS SGEMM / TF32 GEMM, torch SDPA, foreach Adam; "
                       f"{steps} timed steps after 2 warm-ups, CUDA events, same S=94 B=32 batch, dropout on; host-launch bound (~5 k launches per step)")
        out["torch"] = torch.__version__
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
        torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gan_ffn_b200 as G
    from gan_ffn_b200 import parallel, synthetic, train
    from gan_ffn_b200._lib import lib
    import ctypes

    rank, local_rank, world = parallel.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm"
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = lib()
    if args.engine is not None:
        L.cdll.ganffn_set_gemm_engine({"auto": 0, "simt": 1, "tc": 2, "tf32x1": 3}[args.engine])

    reducer = parallel.GradReducer() if world > 1 else None
    nets, ffn = train.build_networks(device=dev)
    gan = train.GANTrainer(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], nets["acoustic_disc"],
                           nets["visual_disc"], nets["text_disc"], grad_reducer=reducer, world_size=world,
                           overlap=not args.no_lanes, batch_disc=not args.no_batch_disc, chains=args.chains,
                           freeze_disc=not args.no_freeze_disc)
    cls = train.ClassifierTrainer(ffn, torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device=dev), grad_reducer=reducer,
                                  overlap=not args.no_lanes)
    if args.no_overlap_reduce:
        for o in (gan.opt_acoustic_G, gan.opt_acoustic_D, gan.opt_visual_G, gan.opt_visual_D, gan.opt_text_G, gan.opt_text_D, cls.optimizer):
            o.overlap_reduce = False
    G.manual_seed(3407 + rank)

    S, B = args.seq_len, args.dialogues
    host = synthetic.make_batch(n_dialogues=B, seq_len=S, seed=3407 + rank).pin()   # weak scaling: fixed work per GPU
    resident = host.to(dev)

    # The step is replayed from a CUDA graph (first call eager, second call captures): same Python loop bodies,
    # recorded once; dropout seeds and the Adam step count live on the device so every replay is a fresh step.
    stepper = train.GraphedTrainStep(gan, cls, seed=3407 + rank, enabled=not args.eager)
    LOSS_KEYS = ["acoustic_D_loss", "acoustic_G_loss", "text_D_loss", "text_G_loss", "visual_D_loss", "visual_G_loss"]

    def step(batch):
        out = stepper(batch)
        return {k: out[k] for k in LOSS_KEYS}, out["loss"], out["pred"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    # ---- warm-up -------------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 2)):      # >= 2: the first call is eager, the second records the graph
        step(resident)
    barrier()

    # ---- timed region: K steps, CUDA events per step on the launching stream, L2 flushed in between --------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    L.cdll.ganffn_reset_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b_ in ev:
        flush.zero_()
        a.record()
        step(resident)
        b_.record()
    barrier()
    launches = int(L.cdll.ganffn_launch_count())
    graphed = stepper.enabled and stepper.last_key in stepper.kernels_per_replay
    if graphed:   # kernels are launched by graph replay, not through the C entry points: count what was recorded
        launches = stepper.kernels_per_replay[stepper.last_key] * args.steps
    ms_total = sum(a.elapsed_time(b_) for a, b_ in ev)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    slots = S * B * world
    value = slots * args.steps / (ms_total / 1e3)

    # ---- stage split (untimed for the headline; reported in config) --------------------------------------------
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    barrier()
    e0.record(); gan.batch(resident); e1.record(); cls.step(resident, train=True); e2.record()
    barrier()
    stage1_ms, stage2_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)

    # ---- end to end: pinned host batch -> device every step, losses and predictions read back -----------------
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        db = host.to(dev, non_blocking=True)
        losses, loss, pred = step(db)
        vals = torch.stack([losses[k] for k in sorted(losses)] + [loss]).cpu()      # 7 scalars
        pred_host = pred.cpu()
        d2h = vals.numel() * 4 + pred_host.numel() * pred_host.element_size()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = slots * args.steps / float(t.item())

    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "quick": True,
                              "stage1_ms": stage1_ms, "stage2_ms": stage2_ms, "e2e": e2e_value, "gpu_launches": launches,
                              "lanes": not args.no_lanes, "batch_disc": not args.no_batch_disc, "freeze_disc": not args.no_freeze_disc, "chains": args.chains, "overlap_reduce": not args.no_overlap_reduce,
                              "wgrad_cap": os.environ.get("GANFFN_WGRAD_CAP"), "clocks": clocks}), flush=True)
        if world > 1 and not parallel.shutdown([stepper]):
            os._exit(0)
        return

    # ---- roofline leg: every GEMM of one step bracketed by CUDA events ------------------------------------------------
    # The step is recorded ONCE MORE as a CUDA graph in serialised form (no lanes, no sub-step chains, weight gradients
    # on the caller's stream) with the library's GEMM brackets switched on: inside a capture they become external
    # event-record nodes, so the replay timestamps every GEMM (incl. its split-K fold) with no host launch gap inside
    # a bracket.  The same replay gives the serialised step time, so the GEMM share is comparable with the ncu launch
    # list (profiles/), which serialises too.
    from gan_ffn_b200 import functional as GF
    prev_side = L.cdll.ganffn_set_side_streams(0)
    gan.overlap, cls.overlap = False, False
    torch.cuda.synchronize()
    seeds = GF.DeviceSeedStream(dev, base=99)
    prev_seeds = GF.set_seed_stream(seeds)
    table_rows, serial_ms, prof_mode = [], None, "graph replay, external event nodes"
    pg = None
    try:
        pg = torch.cuda.CUDAGraph()
        L.cdll.ganffn_gemm_profile_enable(1)
        try:
            with torch.cuda.graph(pg):
                seeds.advance()
                gan.batch(resident)
                cls.step(resident, train=True)
        finally:
            L.cdll.ganffn_gemm_profile_enable(0)
        pg.replay()
        torch.cuda.synchronize()
        flush.zero_()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(); pg.replay(); r1.record()
        torch.cuda.synchronize()
        serial_ms = r0.elapsed_time(r1)
    except Exception as exc:       # fall back to eager brackets behind a long device-side sleep
        prof_mode = f"eager brackets behind a device sleep (graph capture of the brackets failed: {type(exc).__name__}: {exc})"
        torch.cuda.synchronize()
        L.cdll.ganffn_gemm_profile_enable(1)
        torch.cuda._sleep(int(0.08 * 1.9e9))
        gan.batch(resident)
        cls.step(resident, train=True)
        torch.cuda.synchronize()
        L.cdll.ganffn_gemm_profile_enable(0)
    GF.set_seed_stream(prev_seeds)
    L.cdll.ganffn_set_side_streams(prev_side)
    gan.overlap, cls.overlap = not args.no_lanes, not args.no_lanes
    max_rows = 256
    shp = (ctypes.c_int64 * (6 * max_rows))()
    tms = (ctypes.c_double * (2 * max_rows))()
    pg = None                      # the recorded graph holds captured NCCL kernels under DP: drop it before the teardown
    import gc
    gc.collect()
    nrows = int(L.cdll.ganffn_gemm_profile_table(shp, tms, max_rows))
    for r in range(max(nrows, 0)):
        M_, N_, K_, ta, bnk, eng = (int(shp[6 * r + j]) for j in range(6))
        table_rows.append({"M": M_, "N": N_, "K": K_, "transA": ta, "b_is_nk": bnk, "engine": {1: "simt", 2: "tc"}.get(eng, str(eng)),
                           "launches": int(tms[2 * r + 1]), "ms": tms[2 * r]})
    ms_a, fl_a, n_a = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    L.call("ganffn_gemm_profile_collect", 0, ctypes.byref(ms_a), ctypes.byref(fl_a), ctypes.byref(n_a))
    gemm_ms, gemm_flops, gemm_n = ms_a.value, fl_a.value, n_a.value
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    achieved_tf = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    s1, s2 = flops_per_slot(S)
    for row in table_rows:
        fl = 2.0 * row["M"] * row["N"] * row["K"] * row["launches"]
        row["us_per_launch"] = 1e3 * row["ms"] / max(row["launches"], 1)
        row["tflops"] = fl / (row["ms"] / 1e3) / 1e12 if row["ms"] > 0 else 0.0
        row["frac_bf16_peak"] = row["tflops"] / peak_tf if peak_tf else None
        row["frac_3xtf32_ceiling"] = row["tflops"] / (peak_tf / 6.0) if peak_tf else None
    table_rows.sort(key=lambda r: -r["ms"])
    roofline = {"bound": "tensor", "kernel": "GEMM engine (all linear layers fwd/dgrad/wgrad incl. split-K fold)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
                "traffic": traffic, "peak_source": peak_src, "launches_per_step": gemm_n,
                "avg_launch_us": 1e3 * gemm_ms / gemm_n if gemm_n else None,
                "algorithmic_gflop_per_launch": gemm_flops / gemm_n / 1e9 if gemm_n else None,
                "timing": prof_mode, "serialized_step_ms": serial_ms,
                "gemm_share_of_serialized_step": gemm_ms / serial_ms if serial_ms else None,
                "gemm_ms_per_step": gemm_ms,
                "ceiling_3xtf32_tflops": peak_tf / 6.0, "frac_of_3xtf32_ceiling": achieved_tf / (peak_tf / 6.0) if peak_tf else None,
                "per_shape": table_rows[:40],
                "gemm_share_note": "GEMM time and the serialised step time come from the same replay (no lanes, no chains, weight gradients on the caller's stream); the timed headline steps overlap networks and sub-steps, so they are shorter than the serialised step",
                "note": "fp32-parity arithmetic: FFMA tiles or 3xTF32 tcgen05 (3 MMAs per product at half the bf16 rate), so frac <= ~0.17 by construction against the bf16 peak"}

    # ---- reduced-precision variant (north_star: rtol 2e-2 if offered): the same step with one TF32 MMA per product ----
    reduced = None
    if args.engine is None and not args.no_reduced:
        try:
            prev_engine = L.cdll.ganffn_set_gemm_engine(3)
            st2 = train.GraphedTrainStep(gan, cls, seed=4242 + rank, enabled=not args.eager)
            for _ in range(3):
                st2(resident)
            barrier()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
            for a, b_ in evs:
                flush.zero_()
                a.record(); st2(resident); b_.record()
            barrier()
            ms_r = sum(a.elapsed_time(b_) for a, b_ in evs) / len(evs)
            t = torch.tensor([ms_r], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_r = float(t.item())
            reduced = {"engine": "tf32x1", "dtype": "tf32 operands (10-bit mantissa), fp32 accumulate, one tcgen05 MMA per product",
                       "ms_per_step": ms_r, "value": slots / (ms_r / 1e3), "unit": UNIT, "steps": len(evs),
                       "tolerance": "rtol 2e-2 (tests/test_gpu_nets.py::test_reduced_precision_variant)",
                       "note": "separate line, never the fp32-parity headline"}
            st2.release()
        except Exception as exc:
            reduced = {"error": f"{type(exc).__name__}: {exc}"}
        finally:
            L.cdll.ganffn_set_gemm_engine(prev_engine)

    # ---- the HBM-bound kernel of the path: fused Adam over the largest arena (28 B/param), timed alone ------------------
    hbm = None
    if rank == 0:
        n_par = int(nets["visual_gen"].arena().numel)
        bufs = [torch.rand(n_par, device=dev) * 1e-2 for _ in range(4)]
        step_t = torch.ones(1, dtype=torch.int32, device=dev)
        st_ = torch.cuda.current_stream(dev).cuda_stream
        ms_ = []
        for _ in range(6):
            flush.zero_()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            L.call("ganffn_adam_step_dev", bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(), n_par,
                   step_t.data_ptr(), 1e-4, 0.5, 0.6, 1e-8, 0.0, 1.0, st_)
            b_.record()
            torch.cuda.synchronize()
            ms_.append(a_.elapsed_time(b_))
        ms_adam = sorted(ms_[1:])[len(ms_[1:]) // 2]
        peak_hbm = float(peaks.get("hbm_gbs", 6500.0))
        hbm = {"bound": "hbm", "kernel": "adam_kernel over the visual generator's arena", "params": n_par,
               "algorithmic_bytes": 28 * n_par, "ms": ms_adam, "achieved": 28 * n_par / ms_adam / 1e6, "peak": peak_hbm,
               "unit": "GB/s", "frac": 28 * n_par / ms_adam / 1e6 / peak_hbm,
               "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6500 GB/s",
               "note": "p, g, m, v read + p, m, v written = 28 B/param; L2 flushed before each timed launch"}
        del bufs

    # ---- dialogue-graph kernels (north_star parts 2-3; no reference implementation): achieved HBM GB/s -------------
    graph = None
    if rank == 0 and world == 1 and not args.no_graph:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import graph_bench
            del flush
            torch.cuda.empty_cache()
            graph = graph_bench.graph_leg(utterances=args.graph_utterances, reps=3, device=dev)
        except Exception as exc:   # the headline must not depend on the auxiliary leg
            graph = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        cpu = cpu_baseline() if (world == 1 and not args.no_cpu_baseline) else None
        eager = None
        if world == 1 and not args.no_eager_baseline:
            try:
                eager = gpu_eager_baseline(dev)
            except Exception as exc:   # an auxiliary leg must not take the headline down
                eager = {"error": f"{type(exc).__name__}: {exc}"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.engine != "tf32x1" else "tf32x1 (reduced-precision variant, not the parity path)", "data": "synthetic",
                "config": {"workload": WORKLOAD if (S, B) == (S_IEMOCAP, B_IEMOCAP) else f"train_step S={S} B={B} per GPU",
                           "seq_len": S, "dialogues_per_gpu": B, "global_dialogues": B * world, "parallelism": f"dp{world} by dialogue",
                           "stage1_ms": stage1_ms, "stage2_ms": stage2_ms,
                           "stage_split": "eager launches (informational; the timed steps are graph replays)" if graphed else "eager",
                           "launch": "CUDA graph replay of the whole step" if graphed else "eager kernel launches",
                           "algorithmic_gflop_per_step_per_gpu": (s1 + s2) * S * B / 1e9,
                           "step_tflops_per_gpu": (s1 + s2) * S * B / (ms_total / args.steps / 1e3) / 1e12,
                           "l2": "256 MiB buffer written between timed steps (L2 flush); per-step working set ~2 GB >> 126 MB L2",
                           "timing": "sum of per-step CUDA-event intervals on the launching stream, max over ranks",
                           "gemm_engine": args.engine or "auto"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": host.h2d_bytes(), "d2h_bytes_per_step": d2h},
                "gpu_launches": launches,
                "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if eager is not None:
            line["gpu_eager_baseline"] = eager
        if reduced is not None:
            line["reduced_precision"] = reduced
        if hbm is not None:
            line["roofline_hbm"] = hbm
        if graph is not None:
            line["graph"] = graph
        print(json.dumps(line), flush=True)
    if world > 1:
        # orderly teardown: drop the recorded graphs (they hold the captured NCCL kernels) -> drain -> barrier ->
        # destroy the group, under a watchdog (parallel.shutdown); only a teardown that does not return is cut short
        sys.stdout.flush()
        sys.stderr.flush()
        if not parallel.shutdown([stepper]):
            print(f"[rank {rank}] process-group teardown did not return in 30 s; exiting the process", file=sys.stderr, flush=True)
            os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seq-len", type=int, default=S_IEMOCAP)
    ap.add_argument("--dialogues", type=int, default=B_IEMOCAP, help="dialogues per GPU")
    ap.add_argument("--ref-dialogues", type=int, default=0, help="dialogues per step of the CPU reference arm (0 = the same as --dialogues)")
    ap.add_argument("--engine", default=None, choices=["auto", "simt", "tc", "tf32x1"],
                    help="GEMM engine; tf32x1 = the reduced-precision variant (one TF32 MMA per product, rtol 2e-2), never the headline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reduced", action="store_true", help="skip the reduced-precision (tf32x1) leg")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the stock-torch-eager-on-GPU comparator leg")
    ap.add_argument("--no-graph", action="store_true", help="skip the dialogue-graph kernel leg (HBM GB/s of edge build / gathers)")
    ap.add_argument("--graph-utterances", type=int, default=1_000_000)
    ap.add_argument("--no-lanes", action="store_true", help="A/B switch: run the networks of a loop body serially (no concurrent lanes)")
    ap.add_argument("--no-batch-disc", action="store_true", help="A/B switch: train_disc as two discriminator passes (reference body) instead of one [real|fake] pass")
    ap.add_argument("--no-freeze-disc", action="store_true", help="A/B switch: train_gen computes the discriminator's (never read) weight gradients, as the reference body does")
    ap.add_argument("--no-overlap-reduce", action="store_true", help="A/B switch: one gradient all-reduce per arena at optimizer.step() instead of per-layer buckets overlapped with the backward pass")
    ap.add_argument("--chains", type=int, default=2, help="A/B switch: concurrent sub-step chains of the stage-1 batch (1 = serial order)")
    ap.add_argument("--quick", action="store_true", help="headline + e2e only: skip the roofline, Adam, graph, CPU and eager-GPU legs (A/B runs)")
    ap.add_argument("--eager", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph of the step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
