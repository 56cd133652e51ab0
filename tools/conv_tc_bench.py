"""Tensor-core conv (csrc/conv2d_tc.cu) in isolation: the 2-D U-Net's representative layer shapes at a batch large enough
for steady state.  Prints per shape the CUDA-event time, the algorithmic TFLOP/s (2 B H W Cout Cin K^2) and the tensor-pipe
work actually issued (x3 for the fp16 hi/lo split, x Hp Wp / (H W) for the padding ring) as a fraction of the measured
dense bf16 peak in MEASURED_PEAKS.json.  `--one` runs a single shape once (for ncu)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from sdeflow_light_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
h, L = _lib.ctx(dev), _lib.lib()
peak = 1395.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass

SHAPES = [  # B, Cin, Cout, K, H
    (1024, 32, 32, 3, 32), (1024, 96, 32, 3, 32), (1024, 64, 64, 3, 32), (2048, 192, 64, 3, 16), (2048, 128, 128, 3, 16),
    (4096, 256, 128, 3, 8), (2048, 64, 192, 1, 16), (4096, 128, 128, 1, 8),
]
if "--one" in sys.argv:
    SHAPES = [SHAPES[int(sys.argv[sys.argv.index("--one") + 1])]]

for B, Cin, Cout, K, H in SHAPES:
    x = torch.randn(B, Cin, H, H, device=dev)
    W = torch.randn(Cout, Cin, K, K, device=dev) / (Cin * K * K) ** 0.5
    bias = torch.randn(Cout, device=dev)
    gamma, beta = torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev)
    img = torch.empty(L.msgm_conv2d_tc_pack_bytes(Cout, Cin, K), device=dev, dtype=torch.uint8)
    _lib.check(L.msgm_conv2d_tc_pack(h, _lib.ptr(W), Cout, Cin, K, _lib.ptr(img), _lib.stream_ptr(dev)))
    ss = torch.empty(B, Cin, 2, device=dev)
    _lib.check(L.msgm_gn_scale_shift(h, _lib.ptr(x), Cin, None, 0, H * H, 32, B, _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(ss),
                                     _lib.stream_ptr(dev)))
    out = torch.empty(B, Cout, H, H, device=dev)
    d = _lib.Conv2dTcDesc(x.data_ptr(), None, img.data_ptr(), bias.data_ptr(), None, None, ss.data_ptr(), out.data_ptr(), B,
                          Cin, 0, Cout, K, 1, 1, H, H, 2, int("--fast" in sys.argv))
    run = lambda: _lib.check(L.msgm_conv2d_tc(h, C.byref(d), _lib.stream_ptr(dev)))  # noqa: E731
    reps = 1 if "--one" in sys.argv else 10
    for _ in range(0 if "--one" in sys.argv else 3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flop = 2.0 * B * H * H * Cout * Cin * K * K
    pad = ((H + 2) / H) ** 2 if K == 3 else 1.0
    nprod = 1 if "--fast" in sys.argv else 3
    byts = 4.0 * B * H * H * (Cin + Cout)
    print(f"conv{K}x{K} B={B} {Cin:3d}->{Cout:3d} @{H}x{H}: {ms * 1e3:8.1f} us  {flop / ms / 1e9:7.1f} TFLOP/s algorithmic "
          f"({flop / ms / 1e9 / peak:.3f} of {peak:.0f}), tensor-pipe work x{nprod * pad:.2f} -> {nprod * pad * flop / ms / 1e9 / peak:.3f}; "
          f"activation traffic {byts / ms / 1e6:7.1f} GB/s")
