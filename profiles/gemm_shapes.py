"""Times every GEMM shape of the IEMOCAP train step (T=3008) on both engines (CUDA-graph replay of 20 back-to-back
calls, so host launch overhead is excluded and L2 is warm, as inside the step).
Not a benchmark of the step: a tuning aid.  Output: one line per (kind, M, N, K, engine)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import torch  # noqa: E402
from gan_ffn_b200._lib import lib  # noqa: E402
from gtime import graph_time  # noqa: E402

L = lib()
T = int(os.environ.get("T", 3008))
dev = "cuda"
shapes = []
for d, name in ((512, "Gv"), (100, "d100")):
    dff = 2048
    for (n, k, tag) in ((3 * d, d, "in_proj"), (d, d, "out_proj"), (dff, d, "ffn1"), (d, dff, "ffn2")):
        shapes.append((f"{name}.{tag}.fwd", "fwd", T, n, k))
        shapes.append((f"{name}.{tag}.dgrad", "dgrad", T, n, k))
        shapes.append((f"{name}.{tag}.wgrad", "wgrad", T, n, k))
shapes += [("Gv.fc1.fwd", "fwd", T, 1024, 512), ("Gv.fc2.fwd", "fwd", T, 100, 1024), ("d100.fc1.fwd", "fwd", T, 512, 100)]


def run(kind, M, N, K, engine, iters=20):
    L.cdll.ganffn_set_gemm_engine(engine)
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev)
    dy = torch.randn(M, N, device=dev)
    y = torch.empty(M, N, device=dev)
    dx = torch.empty(M, K, device=dev)
    dw = torch.empty(N, K, device=dev)
    ws = torch.empty(max(int(L.cdll.ganffn_gemm_scratch_floats(M, N, K)), int(L.cdll.ganffn_gemm_scratch_floats(M, K, N)),
                         int(L.cdll.ganffn_wgrad_scratch_floats(M, N, K)), 1), device=dev)

    def call():
        st = torch.cuda.current_stream().cuda_stream
        if kind == "fwd":
            L.call("ganffn_linear_fwd", x.data_ptr(), w.data_ptr(), None, None, y.data_ptr(), None, M, N, K, 0, 0, 0.0, 0, 0,
                   ws.data_ptr(), ws.numel(), st)
        elif kind == "dgrad":
            L.call("ganffn_linear_dgrad", dy.data_ptr(), w.data_ptr(), None, dx.data_ptr(), M, N, K, ws.data_ptr(), ws.numel(), st)
        else:
            L.call("ganffn_linear_wgrad", dy.data_ptr(), x.data_ptr(), dw.data_ptr(), None, M, N, K, 0, ws.data_ptr(), st)
    return graph_time(call) * 1e-3


print(f"{'shape':24s} {'M':>5s} {'N':>5s} {'K':>5s}  simt_us  simt_TF    tc_us    tc_TF")
for name, kind, M, N, K in shapes:
    fl = 2.0 * M * N * K
    t1 = run(kind, M, N, K, 1)
    t2 = run(kind, M, N, K, 2)
    print(f"{name:24s} {M:5d} {N:5d} {K:5d} {t1*1e3:8.1f} {fl/t1/1e9:8.1f} {t2*1e3:8.1f} {fl/t2/1e9:8.1f}")
L.cdll.ganffn_set_gemm_engine(0)
