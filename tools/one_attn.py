"""Runs the attention kernels a few times (for ncu).  Usage: one_attn.py S B d nhead p"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gan_ffn_b200._lib import lib  # noqa: E402

S, B, d, nh = (int(a) for a in sys.argv[1:5])
p = float(sys.argv[5])
L = lib()
st = torch.cuda.current_stream().cuda_stream
qkv = torch.randn(S, B, 3 * d, device="cuda")
do = torch.randn(S, B, d, device="cuda")
o = torch.empty(S, B, d, device="cuda")
lse = torch.empty(B * nh * S, device="cuda")
dqkv = torch.empty(S, B, 3 * d, device="cuda")
for _ in range(2):
    L.call("ganffn_attention_fwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), S, B, d, nh, p, 1234, 16, st)
    L.call("ganffn_attention_bwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), do.data_ptr(), dqkv.data_ptr(), S, B, d, nh, p,
           1234, 16, st)
torch.cuda.synchronize()
print("ok")
