#!/usr/bin/env python
"""BASELINE config 5: the train_IEMOCAP_DialogueRNN.py variant -- GAN-FFN fused features on the sm_100a kernels feeding the
DialogueRNN head (gan_ffn_b200/dialogue_rnn.py, stock PyTorch) -- one classifier train step (forward, MaskedNLLLoss,
backward, Adam) at S=94, B=32, eager and replayed from a CUDA graph.  Prints one JSON line.  The head is ~2 x S time steps
of small torch ops, so the eager step is bound by the host; the replay shows what the device needs."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gan_ffn_b200 as G  # noqa: E402
from gan_ffn_b200 import synthetic, train  # noqa: E402


def build(dev):
    torch.manual_seed(3407)
    ga, gv, gt = G.AcousticGenerator(100, dropout=0.2), G.VisualGenerator(100, dropout=0.2), G.TextGenerator(100, dropout=0.2)
    # train_IEMOCAP_DialogueRNN.py defaults: D_m = 100 (fused features), D_g = D_p = 500, D_e = D_h = 100, general attention
    model = G.GAN_FFN_DialogueRNN(ga, gv, gt, 100, 500, 500, 100, 100, 100, 6, False, "general", 0.1, 0.6).to(dev)
    return train.ClassifierTrainer(model, torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device=dev))


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    dev = torch.device("cuda:0")
    S, B, steps = 94, 32, 10
    batch = synthetic.make_batch(n_dialogues=B, seq_len=S).to(dev)
    cls = build(dev)
    eager_ms = timed(lambda: cls.step(batch, train=True), steps, 3)
    stepper = train.GraphedTrainStep(None, build(dev), seed=1)
    graph_ms = timed(lambda: stepper(batch), steps, 4)
    assert stepper.kernels_per_replay, "the timed calls must have been graph replays"
    print(json.dumps({"metric": "gan_ffn_dialogue_rnn_train_step_padded_utterances_per_sec", "unit": "utterances/s", "n_gpus": 1,
                      "value": S * B / graph_ms * 1e3, "ms_per_step": graph_ms, "eager_ms_per_step": eager_ms,
                      "eager_value": S * B / eager_ms * 1e3, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "GAN_FFN_DialogueRNN classifier train step (three generators on the kernels + BiModel head on "
                                             "stock PyTorch), S=94 B=32, dropout on, Adam", "launch": "CUDA graph replay (value) / eager"}}), flush=True)


if __name__ == "__main__":
    main()
