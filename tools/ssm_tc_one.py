"""A few eager SSM training iterations for the ncu launch list: python tools/ssm_tc_one.py [batch] [precision] [dim]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from sdeflow_light_b200 import ssm_fused  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
prec = sys.argv[2] if len(sys.argv) > 2 else "f16tc"
d = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = torch.device("cuda", 0)
prob = bench.build_problem(d)
P, gen = bench.package_objects(prob, dev)
gen.ssm_precision = prec
gen.train()
gen.device_rng = True
x = prob["data"][:batch].to(dev) if batch <= 100000 else None
n = sum(p.numel() for p in gen.a.parameters())
flat = torch.zeros(n, device=dev)
gout = torch.full((batch,), 1.0 / batch, device=dev)
for _ in range(4):
    t_, y, v = gen._prepare(x)
    loss = ssm_fused.fused_loss_and_grads(gen, t_, y, v, gout, flat)
torch.cuda.synchronize()
print("ok", float(loss.mean()), float(flat.abs().max()), P._lib.debug_flags(dev))
if os.environ.get("MSGM_TC_PROF"):
    c = P._lib.debug_counters(dev)
    ntile = max(1, -(-batch // 64) // 148 + (1 if (-(-batch // 64)) % 148 else 0)) if batch > 64 * 148 else 1
    names = ["prologue", "F1 mma", "F1 epi", "F2 mma", "F2 epi", "F3 mma", "F3 epi", "F4 mma", "loss epi", "gW4 mma",
             "B3 epi(+dgrad4)", "B3 mma", "B2 epi", "B2 mma", "B1 epi", "gW1 mma", "gW1 flush", "setup", "flush gW3", "flush gW2"]
    print("cycles of CTA 0 thread 0 (whole launch; CTA 0 ran", ntile, "tiles):")
    for n_, v_ in zip(names, c):
        print(f"  {n_:18s} {v_:10d}")
    print("  total", sum(c[:20]))
