import os, sys
sys.path.insert(0, "/root/repo")
import torch
import sdeflow_light_b200 as P
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = P.VorticityUNet(32, (1, 2, 4), 2, premodule="NormalizeLogRadius", in_space=32, attention_resolutions=(2, 4), flatten_order="F").to(dev)
with torch.no_grad():
    for p_ in net.parameters():
        if p_.dim() > 1 and float(p_.abs().sum()) == 0.0:
            p_.normal_(0, 0.02)
    for B in (16, 128, 512):
        x, t = torch.randn(B, 1024, device=dev), torch.rand(B, device=dev)
        ref = None
        for br in (False, True, False, True):
            net.core.graph_branches = br
            net.__dict__.pop("_graphs", None)
            for _ in range(5):
                y = net(x, t)
            if ref is None:
                ref = y.clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40):
                net(x, t)
            e1.record()
            torch.cuda.synchronize()
            print(f"B={B} graph_branches={br}: {e0.elapsed_time(e1) / 40:.3f} ms  identical={bool(torch.equal(y, ref))}", flush=True)
