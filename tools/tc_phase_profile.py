"""Cycle accounting of the tensor-core sampler (CTA 0): MSGM_TC_PROF=1 python tools/tc_phase_profile.py [dim]"""
import os
import sys

os.environ["MSGM_TC_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N, B = 32, 148 * 2 * 128
sde, mlp = bench.build_problem(d)
P, gen = bench.package_objects(sde, mlp, torch.device("cuda", 0))
x0 = (torch.randn(B, d) * 1.5).cuda()
for _ in range(2):
    P.rk4_stratonovich_sampler(gen, x0, N, keep_all_samples=False, norm_correction=True, precision="f16tc", seed=1,
                               device_out=True)
c = P._lib.debug_counters("cuda:0")
stages = 4 * N
names = {0: "owner: SDE update + L1 operand", 1: "owner: wait d0 (L1)", 2: "owner: wait d0 (L2)", 3: "owner: wait d0 (L3)",
         4: "owner: 3 half-epilogues", 5: "owner: output layer hand-off", 8: "mma: wait L1 operand", 9: "mma: issue",
         10: "mma: wait a0 (L2,L3)", 11: "mma: wait a1 (L2,L3)", 12: "mma: wait act3", 16: "helper: wait d1 (3x)",
         17: "helper: 3 half-epilogues"}
print(f"d={d}: cycles per RK4 stage (CTA 0, flags={P._lib.debug_flags('cuda:0')})")
for k, n in names.items():
    print(f"  {n:34s} {c[k] / stages:9.0f}")
print(f"  owner total {sum(c[0:6]) / stages:9.0f}   mma total {sum(c[8:13]) / stages:9.0f}   helper total {sum(c[16:18]) / stages:9.0f}")
