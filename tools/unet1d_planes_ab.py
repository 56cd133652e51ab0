"""1-D U-Net forward (BASELINE config 3: L = 1000, base 32): fp32-staged tensor-core convs (`planes = False`) vs the
TMA-fed convs on activation planes (`planes = True`, csrc/conv1d_tcp.cu), eager launches and graph replay.

    python tools/unet1d_planes_ab.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import sdeflow_light_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = P.UNet1D(1000, premodule="NormalizeLogRadius").to(dev)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    for B in (16, 256, 1024):
        x, t = torch.randn(B, 1000, device=dev), torch.rand(B, device=dev)
        ref = None
        for mode in ("tc", "tc16"):
            for planes in (False, True):
                for graph in (False, True):
                    net.conv_mode, net.planes, net.cuda_graph = mode, planes, graph
                    y = net(x, t)
                    if ref is None:
                        ref = y
                    ms = timed(lambda: net(x, t))
                    print(f"B={B:5d} conv_mode={mode:4s} planes={int(planes)} graph={int(graph)}: {ms:7.3f} ms per forward "
                          f"({B / ms * 1e3:9.0f} samples/s)  rel diff vs first {float((y - ref).abs().max() / ref.abs().max()):.2e}",
                          flush=True)
