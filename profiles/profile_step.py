"""One train step (stage-1 GAN batch + stage-2 classifier step, S=94 B=32) between cudaProfilerStart/Stop,
for `ncu --profile-from-start off` launch lists and `--set full` captures.  Not a benchmark."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gan_ffn_b200 as G  # noqa: E402
from gan_ffn_b200 import synthetic, train  # noqa: E402
from gan_ffn_b200._lib import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tc"])
ap.add_argument("--stage", default="both", choices=["both", "1", "2"])
ap.add_argument("--dialogues", type=int, default=32)
ap.add_argument("--seq-len", type=int, default=94)
ap.add_argument("--no-batch-disc", action="store_true")
args = ap.parse_args()
lib().cdll.ganffn_set_gemm_engine({"auto": 0, "simt": 1, "tc": 2}[args.engine])
dev = torch.device("cuda:0")
nets, ffn = train.build_networks(device=dev)
gan = train.GANTrainer(nets["acoustic_gen"], nets["visual_gen"], nets["text_gen"], nets["acoustic_disc"],
                       nets["visual_disc"], nets["text_disc"], batch_disc=not args.no_batch_disc)
cls = train.ClassifierTrainer(ffn, torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device=dev))
batch = synthetic.make_batch(n_dialogues=args.dialogues, seq_len=args.seq_len).to(dev)


def step():
    if args.stage in ("both", "1"):
        gan.batch(batch)
    if args.stage in ("both", "2"):
        cls.step(batch, train=True)


step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step; kernels launched so far:", lib().cdll.ganffn_launch_count())
