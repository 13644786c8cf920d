"""Debug aid: prints the in-kernel clock64 timeline of CTA (0,0,0) of gemm_tc_kernel (build with
`make EXTRA=-DGANFFN_TC_TRACE`).  Usage: tc_trace.py M N K [M N K ...]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gan_ffn_b200._lib import lib  # noqa: E402

L = lib()
L.cdll.ganffn_set_gemm_engine(2)
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
import os
L.cdll.ganffn_debug_tc_flags(int(os.environ.get("TCDBG", "0")))
KIND = os.environ.get("TCKIND", "fwd")      # fwd: y = x w^T (K-major A);  wgrad: dw = dy^T x (MN-major A, red.global.add epilogue)
args = [int(a) for a in sys.argv[1:]]
flush = torch.empty(64 * 1024 * 1024, device=dev)
for i in range(0, len(args), 3):
    M, N, K = args[i:i + 3]
    x, w = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
    y = torch.empty(M, N, device=dev)
    ws = torch.empty(max(int(L.cdll.ganffn_gemm_scratch_floats(M, N, K)), 1), device=dev)
    for rep in range(3):
        if rep == 2:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if KIND == "wgrad":   # here (M, N, K) are the linear layer's: the product is [N x K] over M rows
            if rep == 0:
                dyw, dw = torch.randn(M, N, device=dev), torch.zeros(N, K, device=dev)
                ws2 = torch.empty(max(int(L.cdll.ganffn_wgrad_scratch_floats(M, N, K)), 1), device=dev)
            L.call("ganffn_linear_wgrad", dyw.data_ptr(), x.data_ptr(), dw.data_ptr(), None, M, N, K, 1, ws2.data_ptr(), st)
        else:
            L.call("ganffn_linear_fwd", x.data_ptr(), w.data_ptr(), None, None, y.data_ptr(), None, M, N, K, 0, 0, 0.0, 0, 0,
                   ws.data_ptr(), ws.numel(), st)
        b.record()
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 128)()
        L.cdll.ganffn_debug_tc_trace(buf)
        t = list(buf)
        t0 = t[0]
        r = lambda i: t[i] - t0
        nkb = min(16, (K + 31) // 32)
        print(f"M={M} N={N} K={K} rep={rep} ({'cold' if rep == 2 else 'warm'} L2) event_us={a.elapsed_time(b)*1e3:.1f}")
        print(f"  setup_done={r(1)} prod_done={r(3)} acc_ready={r(4)} epi_done={r(5)} all_sync={r(6)}")
        print(f"  epilogue: ld_done={r(70)} staged={r(71)} stored={r(72)} iters=" + " ".join(str(r(73+k)) for k in range(8)))
        print("  A prod arrived: " + " ".join(f"{r(8+k)}" for k in range(nkb)))
        print("  B prod arrived: " + " ".join(f"{r(24+k)}" for k in range(nkb)))
        print("  mma  (full_seen, issued):   " + " ".join(f"({r(48+2*k)},{r(49+2*k)})" for k in range(nkb)))
