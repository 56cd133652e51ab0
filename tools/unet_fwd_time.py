"""Forward time of both U-Net score nets on the kernel path (one line each): python tools/unet_fwd_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import sdeflow_light_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
nets = [("unet1d", P.UNet1D(1000, premodule="NormalizeLogRadius").to(dev), 256, 1000),
        ("unet2d", P.VorticityUNet(32, (1, 2, 4), 2, premodule="NormalizeLogRadius", in_space=32, attention_resolutions=(2, 4),
                                   flatten_order="F").to(dev), 128, 1024)]
with torch.no_grad():
    for name, net, B, d in nets:
        for p_ in net.parameters():
            if p_.dim() > 1 and float(p_.abs().sum()) == 0.0:
                p_.normal_(0, 0.02)
        x, t = torch.randn(B, d, device=dev), torch.rand(B, device=dev)
        for _ in range(5):
            net(x, t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            net(x, t)
        e1.record()
        torch.cuda.synchronize()
        print(f"{os.environ.get('MSGM_LIB_VARIANT', 'default'):8s} {name} batch {B}: {e0.elapsed_time(e1) / 30:.3f} ms per forward", flush=True)
