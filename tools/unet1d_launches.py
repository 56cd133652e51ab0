import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeflow_light_b200 as P
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = P.UNet1D(1000, premodule="NormalizeLogRadius").to(dev)
net.cuda_graph = False  # one kernel node per launch for the profiler
net.planes = os.environ.get("MSGM_PLANES", "1") != "0"
x, t = torch.randn(256, 1000, device=dev), torch.rand(256, device=dev)
with torch.no_grad():
    net(x, t); net(x, t)
torch.cuda.synchronize()
