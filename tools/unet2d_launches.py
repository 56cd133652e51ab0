import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeflow_light_b200 as P
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = P.VorticityUNet(32, (1, 2, 4), 2, premodule="NormalizeLogRadius", in_space=32, attention_resolutions=(2, 4), flatten_order="F").to(dev)
x, t = torch.randn(128, 1024, device=dev), torch.rand(128, device=dev)
net.cuda_graph = False
with torch.no_grad():
    net(x, t); net(x, t)
torch.cuda.synchronize()
