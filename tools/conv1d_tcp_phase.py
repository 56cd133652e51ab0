"""In-kernel cycle accounting of conv1d_tcp_kernel (CTA 0): where the producer, the MMA issuer and an epilogue warp spend
their time for the 1-D U-Net's layer shapes at batch B.   MSGM_TCP_PROF=1 python tools/conv1d_tcp_phase.py [B] [fast]"""
import ctypes as C
import os
import sys

os.environ["MSGM_TCP_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from sdeflow_light_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fast = int(sys.argv[2]) if len(sys.argv) > 2 else 0
Lb, h, st = _lib.lib(), _lib.ctx(dev), _lib.stream_ptr(dev)
print(f"B = {B}, fast = {fast}; cycles of CTA 0 (k = 1000 clk)")
for (C1, C2, Cout, K, L, gelu) in [(32, 0, 32, 3, 1000, 1), (32, 0, 32, 4, 1000, 0), (64, 0, 64, 3, 500, 1), (64, 0, 128, 3, 250, 1),
                                   (128, 0, 128, 3, 250, 1), (128, 0, 128, 4, 250, 0), (128, 128, 128, 3, 250, 1),
                                   (64, 64, 64, 3, 500, 1), (32, 32, 32, 3, 1000, 1)]:
    Cin = C1 + C2
    W = torch.randn(Cout, Cin, K, device=dev) / (Cin * K) ** 0.5
    bias = torch.randn(Cout, device=dev)
    img = torch.empty(Lb.msgm_conv1d_tc_pack_bytes(Cout, Cin, K), device=dev, dtype=torch.uint8)
    _lib.check(Lb.msgm_conv1d_tc_pack(h, _lib.ptr(W), Cout, Cin, Cin, K, _lib.ptr(img), st))
    pls = []
    for Cc in (C1, C2):
        if Cc:
            pl = torch.zeros(Lb.msgm_planes_bytes(B, Cc, L), device=dev, dtype=torch.uint8)
            _lib.check(Lb.msgm_planes_pack(h, _lib.ptr(torch.randn(B, Cc, L, device=dev)), _lib.ptr(pl), B, Cc, L, st))
            pls.append(pl)
    Lout = L if K == 3 else (L - 2) // 2 + 1
    outp = torch.zeros(Lb.msgm_planes_bytes(B, Cout, Lout), device=dev, dtype=torch.uint8)
    d = _lib.Conv1dTcpDesc(pls[0].data_ptr(), pls[1].data_ptr() if C2 else None, img.data_ptr(), bias.data_ptr(), None,
                           outp.data_ptr(), None, B, C1, C2, Cout, K, L, 0, gelu, 0, fast)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        _lib.check(Lb.msgm_conv1d_tcp(h, C.byref(d), st))
    e0.record()
    _lib.check(Lb.msgm_conv1d_tcp(h, C.byref(d), st))
    e1.record()
    torch.cuda.synchronize()
    c = _lib.debug_counters(dev)
    k = lambda v: f"{v / 1e3:6.1f}k"  # noqa: E731
    print(f"{C1:3d}+{C2:3d}->{Cout:3d} k{K} L={L:4d}: {e0.elapsed_time(e1) * 1e3:6.1f} us | epilogue: wait acc {k(c[0])} drain {k(c[1])} "
          f"tiles {c[2]} | mma: wait free acc {k(c[8])} wait stage {k(c[9])} issue {k(c[10])} chunks {c[11]} | "
          f"producer: wait stage {k(c[16])} issue {k(c[17])}", flush=True)
