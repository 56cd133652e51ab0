"""Where one SSM train iteration spends its time: CUDA events around each phase of the eager loop, then the same
iteration replayed as a CUDA graph (train.GraphedSsmStep)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from oracle import msgm_oracle as O  # noqa: E402
from sdeflow_light_b200.train import GraphedSsmStep  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
prob = bench.build_problem(d)
P, gen = bench.package_objects(prob, dev)
data = O.gaussian_mixture(100_000, d, seed=0).to(dev)
opt = torch.optim.Adam(gen.parameters(), lr=1e-3, fused=True)
gen.train()
for B in (256, 4096, 16384, 65536):
    for dev_rng in (False, True):
        gen.device_rng = dev_rng
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
        print(f"B={B} eager device_rng={dev_rng}: " + "  ".join(f"{k}={v:.3f}ms" for k, v in acc.items()) + f"  total={sum(acc.values()):.3f}ms", flush=True)
    gen.device_rng = False
    step = GraphedSsmStep(gen, (B, d), lr=1e-3)
    xs = [data[torch.randint(0, data.shape[0], (B,), device=dev)] for _ in range(4)]
    for _ in range(5):
        step(xs[0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(200):
        step(xs[i & 3])
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B} cuda graph: {e0.elapsed_time(e1) / 200:.4f} ms/iter, own launches/iter {step.launches_per_iter}, loss {float(step.loss):.4f}", flush=True)
