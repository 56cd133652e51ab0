"""Where one SSM train iteration spends its time (CUDA events around each phase)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from oracle import msgm_oracle as O  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
sde, mlp = bench.build_problem(d)
P, gen = bench.package_objects(sde, mlp, dev)
data = O.gaussian_mixture(100_000, d, seed=0).to(dev)
opt = torch.optim.Adam(gen.parameters(), lr=1e-3)
gen.train()
for B in (256, 16384):
    acc = {}
    for it in range(12):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
        ev[0].record()
        opt.zero_grad(set_to_none=False)
        x = data[torch.randint(0, data.shape[0], (B,), device=dev)]
        ev[1].record()
        t_, x, y = gen.sample_txy(x)
        ev[2].record()
        v = P.SDEs.sample_v(x.shape, vtype="rademacher", device=dev)
        ev[3].record()
        loss = gen.ssm_loss(t_, x, y, v).mean()
        ev[4].record()
        loss.backward()
        ev[5].record()
        opt.step()
        ev[6].record()
        torch.cuda.synchronize()
        if it >= 2:
            for k, name in enumerate(["batch", "sample_txy (noising)", "probe v", "ssm forward", "ssm backward+wgrad", "adam"]):
                acc[name] = acc.get(name, 0.0) + ev[k].elapsed_time(ev[k + 1]) / 10
    print(f"B={B}: " + "  ".join(f"{k}={v:.3f}ms" for k, v in acc.items()) + f"  total={sum(acc.values()):.3f}ms")
