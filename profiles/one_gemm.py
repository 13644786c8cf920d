"""Runs one GEMM shape a few times (for `ncu --set full -k regex:gemm_tc`).  Usage: one_gemm.py kind M N K [engine]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gan_ffn_b200._lib import lib  # noqa: E402

kind, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
engine = int(sys.argv[5]) if len(sys.argv) > 5 else 2
L = lib()
L.cdll.ganffn_set_gemm_engine(engine)
dev = "cuda"
x, w, dy = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.randn(M, N, device=dev)
y, dx, dw = torch.empty(M, N, device=dev), torch.empty(M, K, device=dev), torch.empty(N, K, device=dev)
ws = torch.empty(max(int(L.cdll.ganffn_gemm_scratch_floats(M, N, K)), int(L.cdll.ganffn_gemm_scratch_floats(M, K, N)),
                     int(L.cdll.ganffn_wgrad_scratch_floats(M, N, K)), 1), device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    if kind == "fwd":
        L.call("ganffn_linear_fwd", x.data_ptr(), w.data_ptr(), None, None, y.data_ptr(), None, M, N, K, 0, 0, 0.0, 0, 0,
               ws.data_ptr(), ws.numel(), st)
    elif kind == "dgrad":
        L.call("ganffn_linear_dgrad", dy.data_ptr(), w.data_ptr(), None, dx.data_ptr(), M, N, K, ws.data_ptr(), ws.numel(), st)
    else:
        L.call("ganffn_linear_wgrad", dy.data_ptr(), x.data_ptr(), dw.data_ptr(), None, M, N, K, 0, ws.data_ptr(), st)
torch.cuda.synchronize()
print("ok")
