"""Minimal check of csrc/p2p.cu under torchrun: handles exchange, one eager all-reduce+Adam call, expected update."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from sdeflow_light_b200 import _lib  # noqa: E402
from sdeflow_light_b200.dist import P2PAllreduceAdam  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t0 = time.time()


def say(*a):
    print(f"[rank {rank} +{time.time() - t0:5.1f}s]", *a, flush=True)


n = 1000
say("creating")
h = P2PAllreduceAdam(dev, n)
say("connected")
p = torch.zeros(n, device=dev)
tab = torch.tensor([[p.data_ptr(), 0]], dtype=torch.int64, device=dev)
flat = torch.full((1004,), float(rank + 1), device=dev)[:n]
m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
lr, step = torch.tensor(1e-3, device=dev), torch.zeros(1, device=dev, dtype=torch.int64)
for it in range(3):
    h.step(tab, 1, n, flat, m, v, lr, step)
    torch.cuda.synchronize()
    say("step", it, "param[0]", float(p[0]), "adam_step", int(step.item()), "async", _lib.debug_flags(dev))
# mean gradient = (1 + ... + world) / world: Adam's first steps move by lr each
say("expected param ~", -3e-3)
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side):
    h.step(tab, 1, n, flat, m, v, lr, step)
torch.cuda.current_stream(dev).wait_stream(side)
torch.cuda.synchronize()
say("capturing")
with torch.cuda.graph(g):
    h.step(tab, 1, n, flat, m, v, lr, step)
say("captured")
for it in range(3):
    g.replay()
    torch.cuda.synchronize()
    say("replay", it, "param[0]", float(p[0]), "adam_step", int(step.item()), "async", _lib.debug_flags(dev))
if os.environ.get("P2P_TRAINER"):
    import bench
    from sdeflow_light_b200.train import GraphedSsmStep
    prob = bench.build_problem(8)
    P, gen = bench.package_objects(prob, dev)
    say("building trainer")
    st = GraphedSsmStep(gen, (256, 8), lr=1e-3, seed=11, p2p=True)
    say("trainer built, p2p", st.p2p is not None)
    x = prob["data"][rank * 256:(rank + 1) * 256].to(dev)
    for it in range(3):
        l = st(x)
        torch.cuda.synchronize()
        say("train iter", it, float(l), "async", _lib.debug_flags(dev))
dist.destroy_process_group()
