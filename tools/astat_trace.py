"""Debug aid: in-kernel clock64 timeline of CTA (0,0,0) of the tcgen05 GEMM kernels for one forward shape
(build with `make EXTRA=-DGANFFN_TC_TRACE`; GANFFN_NO_ASTAT=1 selects the generic kernel).  Usage: astat_trace.py M N K"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gan_ffn_b200._lib import lib  # noqa: E402

L = lib()
L.cdll.ganffn_set_gemm_engine(2)
M, N, K = (int(a) for a in sys.argv[1:4])
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
x, w = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
b = torch.randn(N, device=dev)
y = torch.empty(M, N, device=dev)
ws = torch.empty(max(int(L.cdll.ganffn_gemm_scratch_floats(M, N, K)), 1), device=dev)
for rep in range(3):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    # act=1 (ReLU), dropout p=0.1 after the activation: the linear1 epilogue
    L.call("ganffn_linear_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), None, y.data_ptr(), None, M, N, K, 1, 0, 0.1, 1234, 7,
           ws.data_ptr(), ws.numel(), st)
    e.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 128)()
    L.cdll.ganffn_debug_tc_trace(buf)
    t = list(buf)
    r = lambda i: t[i] - t[0]
    print(f"M={M} N={N} K={K} rep={rep} event_us={a.elapsed_time(e) * 1e3:.1f}")
    print("  [0..7]   ", [r(i) for i in range(0, 8)])
    print("  [8..15]  ", [r(i) for i in range(8, 16)])
    print("  [24..31] ", [r(i) for i in range(24, 32)])
    print("  [48..63] ", [r(i) for i in range(48, 64)])
    print("  [70..72] ", [r(i) for i in range(70, 73)])
    print("  [80..87] ", [r(i) for i in range(80, 88)])
