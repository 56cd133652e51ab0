"""Timing of one graphed SSM training iteration, fp32 kernels vs the tensor-core step: python tools/ssm_tc_bench.py [dim]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from sdeflow_light_b200.train import GraphedSsmStep  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
prob = bench.build_problem(d)
data = prob["data"].to(dev)
for prec in ("fp32", "f16tc"):
    for batch in (256, 4096, 16384, 65536):
        P, gen = bench.package_objects(prob, dev)
        gen.ssm_precision = prec
        step = GraphedSsmStep(gen, (batch, d), lr=1e-3, seed=1)
        x = data[:batch] if batch <= data.shape[0] else data[torch.randint(0, data.shape[0], (batch,), device=dev)]
        for _ in range(5):
            step(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            loss = step(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
        print(json.dumps({"dim": d, "precision": prec, "batch": batch, "ms_per_iter": ms, "samples_per_s": batch / ms * 1e3,
                          "launches": step.launches_per_iter, "loss": float(loss), "flags": P._lib.debug_flags(dev)}),
              flush=True)
