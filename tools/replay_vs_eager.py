"""Experiment: eager vs CUDA-graph replay of the stage-1 batch and the stage-2 step (same dropout seeds on replay)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gan_ffn_b200 import synthetic, train  # noqa: E402

dev = torch.device("cuda:0")
nets, ffn = train.build_networks(device=dev)
gan = train.GANTrainer(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], nets["acoustic_disc"],
                       nets["visual_disc"], nets["text_disc"])
cls = train.ClassifierTrainer(ffn, torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device=dev))
batch = synthetic.make_batch(n_dialogues=32, seq_len=94).to(dev)


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    t_issue = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, t_issue


for name, fn in (("stage1", lambda: gan.batch(batch)), ("stage2", lambda: cls.step(batch, train=True))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e, ti = timeit(fn)
    print(f"{name}: eager {e:.2f} ms/iter (host issue time {ti:.2f} ms/iter)")
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        gt, _ = timeit(g.replay)
        print(f"{name}: graph replay {gt:.2f} ms/iter")
    except Exception as ex:  # noqa: BLE001
        print(f"{name}: capture failed: {type(ex).__name__}: {ex}")
