"""A/B timing of sampler_tc build variants: MSGM_LIB_VARIANT=<name> python tools/tc_variants.py [dims...]

Prints one JSON line per dimension: particle-steps/s at 2^20 particles x 128 RK4 steps (device-resident, CUDA events,
best and mean of 3 after 2 warm-ups) and the max deviation from the fp32 kernel on 4096 particles x 16 steps with
injected noise (sanity only; the parity tests are in tests/)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

dims = [int(a) for a in sys.argv[1:]] or [2, 8]
dev = torch.device("cuda", 0)
for d in dims:
    prob = bench.build_problem(d)
    P, gen = bench.package_objects(prob, dev)
    B, N = int(os.environ.get("TCV_B", 1 << 20)), int(os.environ.get("TCV_N", 128))
    torch.manual_seed(1)
    x0 = (torch.randn(B, d) * 1.5).to(dev)
    kw = dict(lmbd=0.0, keep_all_samples=False, norm_correction=True, device_out=True)
    xs, nz = x0[:4096].clone(), torch.randn(16, 4096, d, device=dev)
    ref = P.rk4_stratonovich_sampler(gen, xs, 16, noise=nz, precision="fp32", **kw)
    got = P.rk4_stratonovich_sampler(gen, xs, 16, noise=nz, precision="f16tc", **kw)
    err = float((ref - got).abs().max())
    ms = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = P.rk4_stratonovich_sampler(gen, x0, N, seed=i, precision="f16tc", **kw)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = ms[2:]
    print(json.dumps({"variant": os.environ.get("MSGM_LIB_VARIANT", "default"), "d": d,
                      "psteps_per_s_best": B * N / min(ms) * 1e3, "psteps_per_s_mean": B * N / (sum(ms) / len(ms)) * 1e3,
                      "ms": ms, "max_abs_vs_fp32_16steps": err, "finite": bool(torch.isfinite(out).all()),
                      "flags": P._lib.debug_flags(dev)}), flush=True)
