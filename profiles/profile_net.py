"""One d=100 discriminator pass (train mode: forward, BCE, backward) at S=94 B=32 between cudaProfilerStart/Stop,
for `ncu --profile-from-start off --set full` captures of every kernel of an encoder layer.  Not a benchmark."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gan_ffn_b200 as G  # noqa: E402
from gan_ffn_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--net", default="text_disc", choices=["text_disc", "visual_gen", "text_gen"])
ap.add_argument("--dialogues", type=int, default=32)
ap.add_argument("--seq-len", type=int, default=94)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(3407)
net = {"text_disc": G.TextDiscriminator, "visual_gen": G.VisualGenerator, "text_gen": G.TextGenerator}[args.net](100).to(dev).train()
batch = synthetic.make_batch(n_dialogues=args.dialogues, seq_len=args.seq_len).to(dev)
x = batch.visual if args.net == "visual_gen" else batch.text
bce = G.BCELoss()


def once():
    net.zero_grad()
    y = net(x)
    loss = bce(y, torch.ones_like(y)) if y.size(-1) == 1 else y.square().mean()
    loss.backward()


once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one pass")
