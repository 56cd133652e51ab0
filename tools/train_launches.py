"""Replay the graphed SSM train iteration a few times (for an ncu launch list: which kernels make up an iteration)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from oracle import msgm_oracle as O  # noqa: E402
from sdeflow_light_b200.train import GraphedSsmStep  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
prob = bench.build_problem(d)
P, gen = bench.package_objects(prob, dev)
data = O.gaussian_mixture(100_000, d, seed=0).to(dev)
step = GraphedSsmStep(gen, (B, d), lr=1e-3, warmup=1)
x = data[:B].clone()
torch.cuda.synchronize()
print("REPLAYS BEGIN", flush=True)
for _ in range(2):
    step(x)
torch.cuda.synchronize()
print("loss", float(step.loss))
