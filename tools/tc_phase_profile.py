"""Cycle accounting of the tensor-core sampler (CTA 0): MSGM_TC_PROF=1 python tools/tc_phase_profile.py [dim]"""
import os
import sys

os.environ["MSGM_TC_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = 32
B = 148 * (4 if d <= 4 else 3) * 128  # one tile per slot
prob = bench.build_problem(d)
P, gen = bench.package_objects(prob, torch.device("cuda", 0))
x0 = (torch.randn(B, d) * 1.5).cuda()
for _ in range(2):
    P.rk4_stratonovich_sampler(gen, x0, N, keep_all_samples=False, norm_correction=True, precision="f16tc", seed=1,
                               device_out=True)
c = P._lib.debug_counters("cuda:0")
stages = 4 * N
names = {6: "particle: step boundary (re-pin, noise) /4", 7: "particle: L1 operand build + stores",
         0: "particle: fence + arrive", 1: "particle: wait for accumulators", 2: "particle: 3 epilogues",
         3: "particle: output layer (tensor)", 4: "particle: G.y read back", 5: "particle: stage increment + RK",
         8: "issuer: waiting for an operand", 9: "issuer: issue"}
print(f"d={d}: cycles per RK4 stage (CTA 0, flags={P._lib.debug_flags('cuda:0')})")
for k, n in names.items():
    print(f"  {n:34s} {c[k] / stages:9.0f}")
print(f"  particle total {sum(c[0:8]) / stages:9.0f}   issuer total {sum(c[8:13]) / stages:9.0f}")
