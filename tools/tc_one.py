"""One tensor-core sampler call for profiling: python tools/tc_one.py <dim> [particles] [steps] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
N = int(sys.argv[3]) if len(sys.argv) > 3 else 128
prec = sys.argv[4] if len(sys.argv) > 4 else "f16tc"
dev = torch.device("cuda", 0)
prob = bench.build_problem(d)
P, gen = bench.package_objects(prob, dev)
torch.manual_seed(1)
x0 = (torch.randn(B, d) * 1.5).to(dev)
for i in range(2):
    out = P.rk4_stratonovich_sampler(gen, x0, N, seed=i, precision=prec, lmbd=0.0, keep_all_samples=False,
                                     norm_correction=True, device_out=True)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(out).all()), P._lib.debug_flags(dev))
