"""A few conv_wgrad_tc launches for profiling / timing: python tools/wgrad_tc_one.py
(2-D 128 -> 128 3x3 at 16x16, 64 samples; 1-D 128 -> 128 k3 at L = 250, 128 samples; 2-D 32 -> 32 3x3 at 32x32, 64 samples)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from sdeflow_light_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
h, L = _lib.ctx(dev), _lib.lib()
st = _lib.stream_ptr(dev)
torch.manual_seed(0)
for (N, Cout, Cin, KH, KW, Hs, Ws) in ((64, 128, 128, 3, 3, 16, 16), (128, 128, 128, 1, 3, 1, 250), (64, 32, 32, 3, 3, 32, 32)):
    cot = torch.randn(N, Cout, Hs, Ws, device=dev) * 1e-3
    x = torch.randn(N, Cin, Hs, Ws, device=dev)
    gW = torch.empty(Cout, Cin, KH, KW, device=dev)
    amax = torch.empty(1, device=dev)
    _lib.check(L.msgm_amax(h, _lib.ptr(cot), cot.numel(), _lib.ptr(amax), st))
    scratch = torch.empty(L.msgm_conv_wgrad_tc_scratch_bytes(h, N, Cout, Cin, KH, KW, 1, KW // 2, 1, Hs, Ws), device=dev, dtype=torch.uint8)

    def run():
        _lib.check(L.msgm_conv_wgrad_tc(h, _lib.ptr(cot), _lib.ptr(x), None, _lib.ptr(gW), _lib.ptr(amax), None, _lib.ptr(scratch), N, Cout,
                                        Cin, 0, Cin, 0, KH, KW, 1, KW // 2, 1, Hs, Ws, 0, st))

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    flop = 2.0 * N * Hs * Ws * Cout * Cin * KH * KW
    print(f"wgrad_tc N={N} {Cin}->{Cout} {KH}x{KW} @{Hs}x{Ws}: {us:7.1f} us per call (kernel + reduce), "
          f"{flop / us / 1e6:6.1f} TFLOP/s algorithmic", flush=True)
