"""Tuning aid: CUDA-event times of the attention kernels at the IEMOCAP shapes (S=94, B=32)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
from gan_ffn_b200._lib import lib  # noqa: E402
from gtime import graph_time  # noqa: E402

L = lib()
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for (S, B, d, nh, p) in ((94, 32, 100, 10, 0.1), (94, 32, 100, 10, 0.0), (94, 64, 100, 10, 0.1), (94, 32, 512, 8, 0.1), (94, 32, 512, 8, 0.0), (110, 32, 512, 8, 0.1)):
    qkv = torch.randn(S, B, 3 * d, device="cuda")
    do = torch.randn(S, B, d, device="cuda")
    o = torch.empty(S, B, d, device="cuda")
    lse = torch.empty(B * nh * S, device="cuda")
    dqkv = torch.empty(S, B, 3 * d, device="cuda")
    res = {}
    for name in ("fwd", "bwd"):
        def call():
            if name == "fwd":
                L.call("ganffn_attention_fwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), S, B, d, nh, p, 1234, 16, torch.cuda.current_stream().cuda_stream)
            else:
                L.call("ganffn_attention_bwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), do.data_ptr(), dqkv.data_ptr(),
                       S, B, d, nh, p, 1234, 16, torch.cuda.current_stream().cuda_stream)
        res[name] = graph_time(call)
    print(f"GANFFN_ATTN={os.environ.get('GANFFN_ATTN', '3')} S={S} B={B} d={d} nhead={nh} p={p}: fwd {res['fwd']:.1f} us  bwd {res['bwd']:.1f} us")
