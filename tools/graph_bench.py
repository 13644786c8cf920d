#!/usr/bin/env python
"""Achieved HBM GB/s of the dialogue-graph kernels (north_star: ">= 50% of HBM roofline on the ... scatter kernels").
Standalone (`python tools/graph_bench.py`) or imported by bench.py (`graph_leg`).  Sweep-shaped problem: dialogues of
10..110 turns, window 10/10, two speakers, d = 100 features; algorithmic bytes per kernel are stated next to each
number.  CUDA events on the launching stream, L2 flushed (256 MiB write) before every timed launch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def graph_leg(utterances=1_000_000, d=100, wp=10, wf=10, reps=5, device="cuda"):
    from gan_ffn_b200 import synthetic
    from gan_ffn_b200._lib import lib, ptr
    from gan_ffn_b200.graph import DialogueGraph
    L = lib()
    dev = torch.device(device)
    n_dialogues = int(round(utterances / 60))
    lengths = synthetic.ragged_lengths(n_dialogues, 10, 110, seed=11)
    S, B = max(lengths), n_dialogues
    spk = torch.randint(0, 2, (S, B), device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream

    def timed(fn):
        ms = []
        for _ in range(reps + 1):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); e.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(e))
        return sorted(ms[1:])[len(ms[1:]) // 2]

    g = DialogueGraph(lengths, spk, wp, wf, 2, device=dev)
    N, E, R = g.N, g.E, g.n_rel
    x = torch.rand(N, d, device=dev)
    agg = torch.empty(N, R * d, device=dev)
    dx = torch.empty(N, d, device=dev)
    out = {}

    def build():
        L.call("ganffn_graph_build", ptr(g.lengths), ptr(g.speakers), ptr(g.node_off), ptr(g.edge_off), g.B, wp, wf, 2, 0,
               ptr(g.rowptr), ptr(g.col), ptr(g.etype), ptr(g.edge_index), E, ptr(g.node_b), ptr(g.node_t), ptr(g.inv_cnt), st)
    ms = timed(build)
    by = E * (4 + 4 + 16) + N * (8 + 4 + 4 + 4 * R + 4) + 4 * B * 5     # col, etype, edge_index | rowptr, node_b/t, inv_cnt, speakers
    out["edge_build"] = {"ms": ms, "algorithmic_bytes": by, "GBps": by / ms / 1e6, "edges_per_s": E / ms * 1e3}

    def typed():
        L.call("ganffn_graph_gather_typed", ptr(x), ptr(g.rowptr), ptr(g.col), ptr(g.etype), ptr(g.inv_cnt), ptr(agg), N, R, d, ptr(g.node_off), g.B, g.S, st)
    ms = timed(typed)
    by = N * d * 4 + E * 8 + N * 8 + N * R * 4 + N * R * d * 4            # x once, CSR, inv_cnt, out
    out["rgcn_gather_fwd"] = {"ms": ms, "algorithmic_bytes": by, "GBps": by / ms / 1e6}

    def typed_bwd():
        L.call("ganffn_graph_gather_sum", ptr(agg), ptr(g.rowptr_t), ptr(g.col_t), ptr(g.etype_t), ptr(g.inv_cnt), ptr(dx), N, R, R, d, None, 0, 0, st)
    ms = timed(typed_bwd)
    nnz = int((g.inv_cnt[:N * R] > 0).sum().item())                        # non-empty (target, relation) rows of d_out
    by = nnz * d * 4 + E * 8 + N * 8 + N * R * 4 + N * d * 4               # every non-empty d_out row once, CSR, inv, dx
    out["rgcn_gather_bwd"] = {"ms": ms, "algorithmic_bytes": by, "GBps": by / ms / 1e6}

    def plain():
        L.call("ganffn_graph_gather_sum", ptr(x), ptr(g.rowptr), ptr(g.col), None, None, ptr(dx), N, 1, R, d, ptr(g.node_off), g.B, g.S, st)
    ms = timed(plain)
    by = N * d * 4 + E * 4 + N * 8 + N * d * 4                             # x once, CSR, out
    out["graphconv_gather"] = {"ms": ms, "algorithmic_bytes": by, "GBps": by / ms / 1e6,
                               "kernel": "running window sums (two row updates per node); reads only the first / last col of a row"}

    # ---- whole layers: the kernel-level figures above count each kernel's own input + output, and the typed gather's
    # output [N, n_rel*d] is an INTERMEDIATE of the RGCN layer (it exists because gather and contraction are two
    # launches).  Per SURVEY.md §8(d) the honest algorithmic bytes of a layer are its inputs + final output only:
    # x + CSR + inv_cnt + out[N,h].  The layer also does 2*N*(n_rel+1)*d*h FLOP of dense contraction, which makes
    # the RGCN layer tensor-bound once fused (~310 FLOP per honest byte) -- both fractions are reported.
    from gan_ffn_b200.graph import GraphConv, RGCNConv
    h = d
    torch.manual_seed(0)
    rg, gc = RGCNConv(d, h, R).to(dev), GraphConv(d, h).to(dev)
    peaks0 = {}
    try:
        peaks0 = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_ceiling = float(peaks0.get("bf16_tflops", 1600.0)) / 6.0           # 3xTF32 on the tf32 pipe
    with torch.no_grad():
        ms = timed(lambda: rg(x, g))
    by = N * d * 4 + E * 8 + N * 8 + N * R * 4 + N * h * 4
    fl = 2.0 * N * (R + 1) * d * h
    out["rgcn_layer_fwd"] = {"ms": ms, "algorithmic_bytes": by, "GBps": by / ms / 1e6, "gflop": fl / 1e9, "tflops": fl / ms / 1e9,
                             "frac_of_3xtf32_ceiling": fl / ms / 1e9 / tf_ceiling, "bound": "tensor (fp32-parity 3xTF32)",
                             "launches": "typed gather + root product + relation product"}
    with torch.no_grad():
        ms = timed(lambda: gc(x, g))
    by = N * d * 4 + E * 4 + N * 8 + N * h * 4
    fl = 2.0 * N * 2 * d * h
    out["graphconv_layer_fwd"] = {"ms": ms, "algorithmic_bytes": by, "GBps": by / ms / 1e6, "gflop": fl / 1e9, "tflops": fl / ms / 1e9,
                                  "frac_of_3xtf32_ceiling": fl / ms / 1e9 / tf_ceiling, "bound": "hbm / tensor (about even)",
                                  "launches": "window-sum gather + two products"}
    xg = x.clone().requires_grad_(True)
    def rg_fb():
        y = rg(xg, g)
        y.backward(torch.ones_like(y))
        xg.grad = None
    ms = timed(rg_fb)
    out["rgcn_layer_fwd_bwd"] = {"ms": ms, "gflop": 3 * 2.0 * N * (R + 1) * d * h / 1e9, "tflops": 3 * 2.0 * N * (R + 1) * d * h / ms / 1e9,
                                 "algorithmic_bytes": 2 * (N * d * 4 + N * h * 4) + 2 * (E * 8 + N * 8 + N * R * 4),
                                 "GBps": (2 * (N * d * 4 + N * h * 4) + 2 * (E * 8 + N * 8 + N * R * 4)) / ms / 1e6}
    del rg, gc, xg

    # ---- IEMOCAP-shaped batch (BASELINE config 2: "... + graph conv fwd/bwd"): the bench step's 32 dialogues x 94 turns, window
    # 10/10, two speakers, fused features of width 100 -> DialogueGCN's RGCNConv then GraphConv (100 -> 100 -> 100), forward
    # and backward, graph construction and pack / unpack included.  Launch-latency bound at this size (3 008 nodes).
    S2, B2 = 94, 32
    feats = torch.rand(S2, B2, d, device=dev, requires_grad=True)
    spk2 = torch.randint(0, 2, (S2, B2), device=dev)
    rg2, gc2 = RGCNConv(d, h, R).to(dev), GraphConv(h, h).to(dev)

    def iemocap_step():
        g2 = DialogueGraph([S2] * B2, spk2, wp, wf, 2, device=dev)
        xn = g2.pack(feats)
        y = gc2(torch.relu(rg2(xn, g2)), g2)
        out_sbd = g2.unpack(y)
        out_sbd.sum().backward()
        feats.grad = None
        for q in list(rg2.parameters()) + list(gc2.parameters()):
            q.grad = None
        return g2
    g2 = iemocap_step()
    ms = timed(iemocap_step)
    out["iemocap_batch_build_rgcn_graphconv_fwd_bwd"] = {
        "ms": ms, "nodes": g2.N, "edges": g2.E, "utterances_per_s": g2.N / ms * 1e3,
        "what": "graph build (CSR + transposed CSR + edge_index) + pack + RGCNConv + ReLU + GraphConv + unpack, forward and backward, "
                "eager launches, 32 dialogues x 94 turns (the bench step's batch)"}
    del rg2, gc2, feats

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6500.0))
    for v in out.values():
        if "GBps" in v:
            v["frac_of_hbm_peak"] = v["GBps"] / peak
    return {"workload": f"{N} utterances in {B} dialogues of 10..110 turns, window {wp}/{wf}, 2 speakers, d={d}: {E} edges, {R} relations",
            "hbm_peak_GBps": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6500 GB/s",
            "parity": "unpinned: no reference implementation (SURVEY.md D1/D2); checked against oracle/graph_oracle.py",
            "kernels": out}


if __name__ == "__main__":
    print(json.dumps(graph_leg()), flush=True)
