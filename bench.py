#!/usr/bin/env python
"""Headline benchmark: reverse-SDE sampling throughput (particle-steps/s) of the MSGM hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dim D] [--precision P]
                    [--particles B] [--sde-steps N]

Workload (default = BASELINE.json configs[4], the configuration the metric's 1/2/4/8-GPU sweep is quoted on; it fits one
GPU): MLP score net (NN.py, hidden 128, NormalizeLogRadius) on synthetic higher-dimensional Gaussian-mixture data (d = 8),
dense multiplicative SDE, 10^7 particles per GPU, RK4-Stratonovich with 1000 reverse steps, lambda = 0, radius correction
on, final state only (keep_all_samples=False).  One bench "step" is one full sampler call over the batch.  Weights are
random-init (no checkpoints offline), data synthetic.  BASELINE.json configs[1] (2^20 particles x 128 steps) is timed in
the same run and reported under ``config2``; ``--particles 1048576 --sde-steps 128`` makes it the headline instead.

* ``value``      whole-job particle-steps/s with x_0 already resident in HBM (kernel launch -> completion).
* ``e2e``        the same through the public API ``rk4_stratonovich_sampler`` with HOST buffers: pinned x_0 H2D
                 and the final states D2H are inside the timed region.  ``e2e_stock``: the unmodified drop-in call
                 (CPU tensor in, CPU tensor out) as the reference driver issues it (MSGM_higherDim.py:903-906).
* ``roofline``   tensor-pipe roofline of the sampler kernel: algorithmic FLOP / CUDA-event duration vs the
                 measured cuBLAS bf16 peak in MEASURED_PEAKS.json (sustained figure: the kernel runs for seconds).
* ``fp32_parity`` throughput of the fp32 parity mode (``precision="fp32"``, CUDA cores, reference arithmetic) on a bounded
                 sample of the same workload, next to the f16 tensor-core headline.
* ``cpu_baseline`` the reference's own CPU implementation (baseline/_ref when present, else the oracle port, which
                 issues the same ATen op sequence) timed on this host on a bounded sample of the same workload.
* ``--impl reference`` times that CPU path alone, on the same config, as the reference arm.

Multi-GPU: launched by torchrun, one rank per GPU; particles are sharded (weak scaling, B per GPU fixed), no
data-path collective; Philox noise is keyed by the global particle index.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SDE_STEPS = 1000               # BASELINE.json configs[4]
PARTICLES_PER_GPU = 10_000_000
CFG2_STEPS, CFG2_PARTICLES = 128, 1 << 20   # BASELINE.json configs[1]
STAGES = {"em": 1, "heun": 2, "rk4": 4}


def flop_per_particle_step(d: int, pre: int, dense: bool, scheme: str = "rk4") -> float:
    """S * (F_net + [dense] (2 d^3 + 4 d^2)), F_net = 2*128*(2d+1+pre) + 65536."""
    f_net = 2 * 128 * (2 * d + 1 + pre) + 65536
    return STAGES[scheme] * (f_net + ((2 * d ** 3 + 4 * d ** 2) if dense else 0))


def load_traffic(workload: str, precision: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this exact workload."""
    for name in ("traffic_r02.json", "traffic_r01.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(p):
            ent = json.load(open(p)).get(f"{workload}/{precision}")
            if ent:
                return ent["traffic_bytes"]
    return None  # no capture of this exact workload: say so rather than quote another one's bytes


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        j = json.load(open(p))
        return dict(tflops=float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), hbm=float(j["hbm_gbs"]),
                    src="measured(sustained)")
    return dict(tflops=1590.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    The sampler is started well before the region (nvidia-smi needs up to a second to enumerate an 8-GPU box) and every
    row is stamped on arrival; `begin()` / `end()` bracket the timed region and `summary()` reports the rows inside it
    (or, for a region shorter than the sampling period, the rows closest to it -- `samples_in_region` says which)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.rows, self.proc, self.index, self.t0, self.t1 = [], None, index, None, None
        if not enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            time.sleep(0.12)  # let the last in-region sample arrive
            self.proc.terminate()
            self.t.join(timeout=2)
            self.proc = None

    def summary(self):
        rows = [(ts, r) for ts, r in self.rows if r and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        inside = [r for ts, r in rows if self.t0 is not None and self.t0 <= ts <= (self.t1 or ts) + 0.06]
        n_in = len(inside)
        if not inside:  # region shorter than the sampling period: the samples nearest to it
            mid = 0.5 * ((self.t0 or rows[-1][0]) + (self.t1 or rows[-1][0]))
            inside = [r for _, r in sorted(rows, key=lambda e: abs(e[0] - mid))[:3]]
        sm = sorted(int(r[0]) for r in inside)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in inside)]
        mx = [int(r[1]) for r in inside if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "samples_in_region": n_in}


def build_problem(d: int, seed: int = 0) -> dict:
    """Synthetic config-2/5 problem, identical on every rank (CPU tensors): Gaussian-mixture data (means 3 randn(K,d), shared
    correlation A = randn(d,d) as the reference's Gaussian(correlation=True), data.py:751-778), the log-radius table, a
    dense skew-symmetric G scaled to tr(L_G) = -d/2 (SDEs.py:315-341) and nn.Linear-style random weights with the output
    layer widened x6 so that the untrained net produces an O(1) drift."""
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(d, d, generator=g)
    mu = 3.0 * torch.randn(8, d, generator=g)
    c = torch.randint(0, 8, (100_000,), generator=g)
    data = (torch.randn(100_000, d, generator=g) @ A.T) * 0.3 + mu[c]
    F = torch.stack([torch.randn(d, d, generator=g) for _ in range(d)], dim=2)
    G = 0.5 * (F - F.transpose(0, 1))
    G = torch.sqrt(-0.5 * d / torch.trace(0.5 * torch.einsum("ijk,jmk->im", G, G))) * G
    sizes = [d + 2, 128, 128, 128, d]
    W, b = [], []
    for l in range(4):
        k = 1.0 / sizes[l] ** 0.5
        W.append((torch.rand(sizes[l + 1], sizes[l], generator=g) * 2 - 1) * k * (6.0 if l == 3 else 1.0))
        b.append((torch.rand(sizes[l + 1], generator=g) * 2 - 1) * k * (6.0 if l == 3 else 1.0))
    return dict(dim=d, data=data, G=G, L_G=0.5 * torch.einsum("ijk,jmk->im", G, G),
                r_T=torch.log(torch.linalg.norm(data, dim=1) + 1e-6), W=W, b=b, beta_min=0.1, beta_max=20.0, T=1.0,
                t_epsilon=1e-3, num_steps_forward=16)


def package_objects(prob, device):
    import sdeflow_light_b200 as P
    d = prob["dim"]
    T = torch.nn.Parameter(torch.FloatTensor([prob["T"]]), requires_grad=False)
    base = P.MSGMsde(torch.randn(8, d), beta_min=prob["beta_min"], beta_max=prob["beta_max"], T=T,
                     t_epsilon=prob["t_epsilon"], denseTensor=True, norm_sampler="ecdf", norm_map="log",
                     num_steps_forward=prob["num_steps_forward"], device=device, estim_cst_norm_dens_r_T=False)
    base.G, base.L_G, base.r_T = prob["G"].to(device), prob["L_G"].to(device), prob["r_T"].to(device)
    net = P.MLP(input_dim=d, premodule="NormalizeLogRadius")
    with torch.no_grad():
        for i, l in enumerate(net.linears()):
            l.weight.copy_(prob["W"][i])
            l.bias.copy_(prob["b"][i])
    gen = P.PluginReverseSDE(base, net.to(device), T, deviceReverseSDE=device).to(device)
    return P, gen


# ---- CPU legs (cpu_baseline, --impl reference): the only places that touch oracle/ or baseline/_ref ----------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def cpu_objects(prob):
    """(kind, sampler callable, train-step factory) of the reference's CPU path on this problem: the UNMODIFIED reference
    from baseline/_ref when it travelled with the repo (plot / IO modules stubbed), else the oracle port."""
    d = prob["dim"]
    if os.path.isfile(os.path.join(REF_DIR, "SDEs.py")):
        os.environ["MSGM_REFERENCE_ROOT"] = REF_DIR
        from oracle import ref_live
        ref = ref_live.load()
        base, gen, net = ref_live.build(ref, "msgm_dense", d, prob["data"][:4096], "NormalizeLogRadius",
                                        beta_min=prob["beta_min"], beta_max=prob["beta_max"], t_eps=prob["t_epsilon"],
                                        n_fwd=prob["num_steps_forward"], T0=prob["T"])
        base.G, base.L_G, base.r_T = prob["G"], prob["L_G"], prob["r_T"]
        with torch.no_grad():
            lin = [m for m in net.main if isinstance(m, torch.nn.Linear)]
            for i, l in enumerate(lin):
                l.weight.copy_(prob["W"][i])
                l.bias.copy_(prob["b"][i])

        def sample(B, n):
            x0 = base.latent_sample(B, d)
            t0 = time.perf_counter()
            ref.sde_scheme.rk4_stratonovich_sampler(gen, x0, num_steps=n, lmbd=0., keep_all_samples=False,
                                                    norm_correction=True)
            return time.perf_counter() - t0

        def train(batch, iters):
            opt = torch.optim.Adam(gen.parameters(), lr=1e-3)
            gen.train()
            t0 = None
            for it in range(iters + 1):
                if it == 1:
                    t0 = time.perf_counter()
                opt.zero_grad()
                x = prob["data"][torch.randint(0, prob["data"].shape[0], (batch,))]
                gen.ssm(x).mean().backward()
                opt.step()
            return batch * iters / (time.perf_counter() - t0)

        return "reference", sample, train
    from oracle import msgm_oracle as O
    sde = O.OSde("msgm_dense", d, prob["beta_min"], prob["beta_max"], prob["T"], prob["t_epsilon"],
                 prob["num_steps_forward"], G=prob["G"], L_G=prob["L_G"], r_T=prob["r_T"], norm_map="log")
    mlp = O.OMlp(prob["W"], prob["b"], True, d)

    def sample(B, n):
        x0 = O.latent_sample(sde, B)
        t0 = time.perf_counter()
        O.integrate(O.OReverse(sde, mlp), x0, n, "rk4", 0.0, keep_all_samples=False, norm_correction=True)
        return time.perf_counter() - t0

    def train(batch, iters):
        params = [p.clone().requires_grad_(True) for p in mlp.parameters()]
        rev = O.OReverse(sde, O.OMlp(params[0::2], params[1::2], True, d))
        opt = torch.optim.Adam(params, lr=1e-3)
        t0 = None
        for it in range(iters + 1):
            if it == 1:
                t0 = time.perf_counter()
            opt.zero_grad()
            x = prob["data"][torch.randint(0, prob["data"].shape[0], (batch,))]
            loss, _ = O.ssm(rev, x)
            loss.mean().backward()
            opt.step()
        return batch * iters / (time.perf_counter() - t0)

    return "port", sample, train


def time_gpu_train(P, gen, data_dev, batch: int, iters: int, world: int, dev, graphed: bool = False,
                   precision: str = "fp32"):
    """samples/s (all ranks) of gen.ssm(x).mean().backward(); [all-reduce]; Adam.step() on the fused SSM kernels,
    eagerly (the reference's loop verbatim) or replayed as a CUDA graph (sdeflow_light_b200.train.GraphedSsmStep)."""
    import torch.distributed as dist
    from sdeflow_light_b200 import dist as D
    params = [p for p in gen.parameters() if p.requires_grad]
    gen.ssm_precision = precision  # "fp32": reference arithmetic (CUDA cores); "f16tc": one tcgen05 launch per iteration
    if graphed:
        from sdeflow_light_b200.train import GraphedSsmStep
        gstep = GraphedSsmStep(gen, (batch, data_dev.shape[1]), lr=1e-3)

        def step():
            return gstep(data_dev[torch.randint(0, data_dev.shape[0], (batch,), device=dev)])
    else:
        opt = torch.optim.Adam(params, lr=1e-3, fused=True)

        def step():
            opt.zero_grad(set_to_none=False)
            x = data_dev[torch.randint(0, data_dev.shape[0], (batch,), device=dev)]
            loss = gen.ssm(x).mean()
            loss.backward()
            D.allreduce_grads_(params)
            opt.step()
            return loss

    for _ in range(5):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    l0 = P._lib.launch_count(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        loss = step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    n_launch = gstep.launches_per_iter if graphed else (P._lib.launch_count(dev) - l0) / iters
    if graphed:
        gstep.close()  # collective teardown of the peer-memory all-reduce buffers (no-op on one GPU)
    return world * batch * iters / (ms / 1e3), ms / iters, n_launch, float(loss.detach())


def time_unet_forward(P, dev):
    """Score-net forward throughput of BASELINE configs 3 and 4 (one net evaluation = one Runge-Kutta stage of the generic
    sampler): this repo's kernel path (tcgen05 convs / attention, split fp16 x3 = fp32-level parity) next to torch's fp32
    library path on the same module and weights."""
    out = {"metric": "score_net_forward_samples_per_sec", "unit": "samples/s", "precision": "fp16x3 split (fp32-level parity)",
           "runs": []}

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    torch.manual_seed(0)
    nets = [("unet1d_L1000_base32", P.UNet1D(1000, premodule="NormalizeLogRadius").to(dev), 256, 1000, 0.445),
            ("vorticity_unet2d_32x32_base32", P.VorticityUNet(32, (1, 2, 4), 2, premodule="NormalizeLogRadius", in_space=32,
                                                              attention_resolutions=(2, 4), flatten_order="F").to(dev), 128,
             1024, 1.204)]
    large = {"unet1d_L1000_base32": 1024, "vorticity_unet2d_32x32_base32": 512}  # second, larger particle batch per net
    with torch.no_grad():
        for name, net, B, d, gflop in nets:
            for p_ in net.parameters():  # the reference zero-initialises some convs: give them weights so that work is real
                if p_.dim() > 1 and float(p_.abs().sum()) == 0.0:
                    p_.normal_(0, 0.02)
            x, t = torch.randn(B, d, device=dev), torch.rand(B, device=dev)
            ms = timeit(lambda: net(x, t), 10)
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                prev = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = False
                ms_t = timeit(lambda: net._forward(x, t), 5)
                ref = net._forward(x, t)
                torch.backends.cuda.matmul.allow_tf32 = prev
            err = float((net(x, t) - ref).abs().max() / ref.abs().max())
            xl, tl = torch.randn(large[name], d, device=dev), torch.rand(large[name], device=dev)
            ms_l = timeit(lambda: net(xl, tl), 5)
            out["runs"].append({"net": name, "batch": B, "ms_per_forward": ms, "value": B / ms * 1e3,
                                "algorithmic_tflops": gflop * B / ms, "torch_fp32_ms": ms_t,
                                "rel_diff_vs_torch_fp32": err,
                                "large_batch": {"batch": large[name], "ms_per_forward": ms_l, "value": large[name] / ms_l * 1e3,
                                                "algorithmic_tflops": gflop * large[name] / ms_l}})
    # (i) RK4 reverse sampling of BASELINE configs 3 / 4 through the drop-in sampler (sparse multiplicative SDE, norm correction,
    # in-kernel Philox): one CUDA graph per step (4 net evaluations + 4 stage updates), against SURVEY 8d's tensor bounds;
    # (ii) one SSM training iteration of the same nets on the hand-written training path (unet_train.py: forward-mode pairs on the
    # tcgen05 convs, hand-derived backward), with the one-launch noising and Adam, replayed as CUDA graphs by train.GraphedSsmStep
    from sdeflow_light_b200 import unet_train
    from sdeflow_light_b200.train import GraphedSsmStep
    bounds = {"unet1d_L1000_base32": 7.8e5, "vorticity_unet2d_32x32_base32": 2.9e5}
    for (name, net, Bs, d, _), (Bt, nfwd) in zip(nets, ((64, 16), (32, 128))):
        try:
            data = torch.randn(512, d)
            T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
            base = P.MSGMsde(data, beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=False, norm_map="log",
                             num_steps_forward=nfwd, device=dev, estim_cst_norm_dens_r_T=False)
            gen = P.PluginReverseSDE(base, net, T, deviceReverseSDE=dev).to(dev)
            x0, nst = gen.latent_sample(Bs, d), 16
            ms_s = timeit(lambda: P.rk4_stratonovich_sampler(gen, x0, nst, keep_all_samples=False, norm_correction=True,
                                                             seed=1, device_out=True), 3)
            entry = {"sampling_batch": Bs, "sampling_steps": nst, "sampling_ms_per_call": ms_s,
                     "sampling_particle_steps_per_sec": Bs * nst / ms_s * 1e3,
                     "sampling_frac_of_tensor_bound": Bs * nst / ms_s * 1e3 / bounds[name]}
            Bl = large[name]
            xl0 = gen.latent_sample(Bl, d)
            ms_l = timeit(lambda: P.rk4_stratonovich_sampler(gen, xl0, nst, keep_all_samples=False, norm_correction=True,
                                                             seed=1, device_out=True), 2)
            entry["sampling_large_batch"] = {"batch": Bl, "steps": nst, "ms_per_call": ms_l,
                                             "particle_steps_per_sec": Bl * nst / ms_l * 1e3,
                                             "frac_of_tensor_bound": Bl * nst / ms_l * 1e3 / bounds[name]}
            # the same call with the convs in the single-product mode (`conv_mode = "tc16"`: one fp16 product per contraction
            # instead of the fp16 x 3 split; forward 4e-4 (1-D) / 2e-3 (2-D) of max|y| from the split mode, sampling only)
            holder = net.core if hasattr(net, "core") else net
            holder.conv_mode = "tc16"
            try:
                ms_h = timeit(lambda: P.rk4_stratonovich_sampler(gen, xl0, nst, keep_all_samples=False, norm_correction=True,
                                                                 seed=1, device_out=True), 2)
            finally:
                holder.conv_mode = "tc"
            entry["sampling_large_batch_tc16"] = {"batch": Bl, "steps": nst, "ms_per_call": ms_h,
                                                  "particle_steps_per_sec": Bl * nst / ms_h * 1e3,
                                                  "frac_of_tensor_bound": Bl * nst / ms_h * 1e3 / bounds[name]}
            xs = data[:Bt].to(dev)
            entry["train_path"] = "hand-written kernels" if unet_train.supported(gen, xs) else "library autograd"
            step = GraphedSsmStep(gen, (Bt, d), lr=1e-4)
            ms = timeit(lambda: step(xs), 10)
            entry.update({"train_batch": Bt, "num_steps_forward": nfwd, "train_ms_per_iter": ms,
                          "train_samples_per_sec": Bt / ms * 1e3, "train_loss": float(step.loss)})
        except Exception as exc:
            entry = {"train_error": f"{type(exc).__name__}: {exc}"}
        for r in out["runs"]:
            if r["net"] == name:
                r.update(entry)
    return out


def cpu_sample_size(args):
    """Bounded sample of the workload for the CPU legs: ~10-30 s of CPU work per measurement."""
    return 50_000, 4


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)  # torchrun sets OMP_NUM_THREADS=1; this arm may use every host core
    prob = build_problem(args.dim)
    kind, sample, _ = cpu_objects(prob)
    B, n = cpu_sample_size(args)
    times = []
    for i in range(args.warmup + args.steps):
        t = sample(B, n)
        if i >= args.warmup:
            times.append(t)
    ms = 1e3 * sum(times) / len(times)
    value = B * n / (ms / 1e3)
    what = "the unmodified reference (baseline/_ref, sde_scheme.rk4_stratonovich_sampler, torch CPU fp32)" \
        if kind == "reference" else "oracle port of the reference's op sequence (torch CPU fp32)"
    emit({
        "impl": "reference", "metric": "reverse_sde_particle_steps_per_sec", "value": value,
        "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, "cpu"),
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": torch.get_num_threads(),
                         "kind": kind,
                         "sample": f"{B} particles x {n} RK4 steps per bench step, same net/SDE as the GPU arm; {what}"},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


def workload_name(dim, steps, particles):
    return f"gaussmix_d{dim}_msgm_dense_mlp128_rk4_n{steps}_{particles}particles_per_gpu"


def workload_config(args, where):
    return {"workload": workload_name(args.dim, args.sde_steps, args.particles),
            "dim": args.dim, "particles_per_gpu": args.particles, "sde_steps": args.sde_steps, "scheme": "rk4",
            "lmbd": 0.0, "norm_correction": True, "precision": args.precision if where == "gpu" else "fp32",
            "l2_policy": "working set is on-chip (weights in smem, state in registers); x_0/x_N (2 x 4*B*d bytes) "
                         "streamed once per call; a 256 MB buffer is rewritten between timed calls to flush L2"}


_JSON_FD = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else written to fd 1 during the run (NCCL's version banner, library
    warnings) is routed to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dim", type=int, default=8)
    ap.add_argument("--precision", default="f16tc", choices=["fp32", "f16tc"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the SSM training leg (profiling runs)")
    ap.add_argument("--no-unet", action="store_true", help="skip the U-Net score-net forward leg (configs 3 and 4)")
    ap.add_argument("--no-extra", action="store_true", help="skip config 2 / fp32 parity / per-dimension side measurements")
    ap.add_argument("--sde-steps", type=int, default=N_SDE_STEPS, help="reverse-SDE steps per sampler call (default 1000 = "
                    "BASELINE config 5; config 2 uses 128)")
    ap.add_argument("--particles", type=int, default=PARTICLES_PER_GPU, help="particles per GPU (default 10^7 = BASELINE "
                    "config 5; config 2 uses 2^20)")
    args = ap.parse_args()
    _claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (sdeflow_light_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    prob = build_problem(args.dim)
    P, gen = package_objects(prob, dev)
    B, N = args.particles, args.sde_steps
    torch.manual_seed(1234 + rank)
    x0_host = (torch.randn(B, args.dim) * 1.5).pin_memory()
    x0_dev = x0_host.to(dev)
    flush = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32)  # 256 MB > 126 MB L2
    kw = dict(lmbd=0.0, keep_all_samples=False, norm_correction=True, precision=args.precision,
              particle_offset=rank * B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_call(seed):
        return P.rk4_stratonovich_sampler(gen, x0_dev, N, seed=seed, device_out=True, **kw)

    def time_calls(fn, reps, warm):
        """mean CUDA-event milliseconds of `reps` calls after `warm` untimed ones (secondary measurements)."""
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize(dev)
        ms = []
        for i in range(reps):
            flush.fill_(float(i))
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            fn(50 + i)
            b_.record()
            torch.cuda.synchronize(dev)
            ms.append(a_.elapsed_time(b_))
        return sum(ms) / len(ms)

    # ---- the other named configuration, the fp32 parity mode and the other dimensions, same run (taken first: seconds of work) ----
    def sampler_rate(d_, nb, ns, prec, reps=2):
        pr = prob if d_ == args.dim else build_problem(d_)
        g_ = gen if d_ == args.dim else package_objects(pr, dev)[1]
        xs = torch.randn(nb, d_, device=dev) * 1.5
        ms_ = time_calls(lambda sd: P.rk4_stratonovich_sampler(
            g_, xs, ns, seed=sd, device_out=True, lmbd=0.0, keep_all_samples=False, norm_correction=True, precision=prec),
            reps, 1)
        return nb * ns / (ms_ / 1e3), ms_

    peaks = load_peaks()
    extra = {}
    if rank == 0 and not args.no_extra:
        v2, ms2 = sampler_rate(args.dim, CFG2_PARTICLES, CFG2_STEPS, args.precision, reps=3)
        extra["config2"] = {"workload": workload_name(args.dim, CFG2_STEPS, CFG2_PARTICLES), "value": v2,
                            "unit": "particle-steps/s", "ms_per_call": ms2, "precision": args.precision,
                            "roofline_frac": v2 * flop_per_particle_step(args.dim, 1, True) / 1e12 / peaks["tflops"]}
        nb32 = 1 << 17
        v32, ms32 = sampler_rate(args.dim, nb32, 128, "fp32")
        extra["fp32_parity"] = {"precision": "fp32", "kernel": "sample_fp32_kernel", "value": v32,
                                "unit": "particle-steps/s", "ms_per_call": ms32,
                                "sample": f"{nb32} particles x 128 RK4 steps, same net/SDE (fp32 CUDA-core parity mode: "
                                          "states within 5e-5 + 5e-5 |x| of the reference)",
                                "fp32_fma_frac": v32 * flop_per_particle_step(args.dim, 1, True) / 1e12 / 72.0}
        by_dim = []
        for d_ in (2, 4, 8, 16):
            vd, msd = sampler_rate(d_, CFG2_PARTICLES, CFG2_STEPS, "f16tc")
            by_dim.append({"dim": d_, "value": vd, "ms_per_call": msd,
                           "roofline_frac": vd * flop_per_particle_step(d_, 1, True) / 1e12 / peaks["tflops"]})
        extra["by_dim_f16tc_2p20_x128"] = by_dim

    clk = ClockSampler(local, enabled=(rank == 0))  # started before the warm-up so that it is sampling by the timed region
    for i in range(args.warmup):
        one_call(i)
    barrier()

    # ---- device-resident timing (value, roofline) ----------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = P._lib.launch_count(dev)
    barrier()
    clk.begin()
    for i in range(args.steps):
        flush.fill_(float(i))
        ev[i][0].record()
        one_call(100 + i)
        ev[i][1].record()
    barrier()
    clk.end()
    clk.stop()
    launches = P._lib.launch_count(dev) - launches0
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    t_dev = torch.tensor([sum(kern_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())
    value = world * B * N * args.steps / (total_ms / 1e3)
    P._lib.check_async(dev)

    # ---- end-to-end through the public API with host buffers -------------------------------------------------
    # Every step copies its input from pinned host memory and its result back to pinned host memory inside the timed
    # region; the copies run on two side streams so that step i+1's upload and step i-1's download overlap step i's kernel
    # (what a host-side caller of the drop-in API would do with a stream of batches).
    out_host = [torch.empty(B, args.dim).pin_memory() for _ in range(2)]
    xbuf = [torch.empty_like(x0_dev) for _ in range(2)]
    s_in, s_out, main = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(s_in):
            if i >= 2:
                s_in.wait_event(ev_free[i % 2])  # the sampler call that read this buffer has finished
            xbuf[i % 2].copy_(x0_host, non_blocking=True)
            ev_in[i % 2].record(s_in)

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s_in.wait_stream(main)   # nothing of the copies starts before the timed region does
    s_out.wait_stream(main)
    upload(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            upload(i + 1)
        main.wait_event(ev_in[i % 2])
        res = P.rk4_stratonovich_sampler(gen, xbuf[i % 2], N, seed=200 + i, device_out=True, **kw)
        ev_free[i % 2].record(main)
        ev_res = torch.cuda.Event()
        ev_res.record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_res)
            out_host[i % 2].copy_(res, non_blocking=True)
        res.record_stream(s_out)
    main.wait_stream(s_out)  # the last download is inside the timed region
    e1.record()
    barrier()
    t_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * N * args.steps / (float(t_e2e.item()) / 1e3)
    assert all(bool(torch.isfinite(o).all()) for o in out_host[:min(2, args.steps)])
    del out_host, xbuf

    # ---- the stock drop-in call: CPU tensor in, CPU tensor out, nothing but the reference's own arguments ---------------
    # (a) as the reference driver issues it (MSGM_higherDim.py:903-906): 10^4 particles, 128 steps, every step kept;
    # (b) a 2^20-particle batch, final state only.  Wall clock around the blocking call (it returns a host tensor).
    stock = []
    if rank == 0:
        for nb, ns, keep in ((10_000, 128, True), (1 << 20, 128, False)):
            xc = torch.randn(nb, args.dim) * 1.5
            best = float("inf")
            for _ in range(3):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                out = P.rk4_stratonovich_sampler(gen, xc, ns, lmbd=0., keep_all_samples=keep, include_t0=keep,
                                                 norm_correction=True, precision=args.precision)
                best = min(best, time.perf_counter() - t0)
            assert out.device.type == "cpu" and bool(torch.isfinite(out).all())
            stock.append({"particles": nb, "sde_steps": ns, "keep_all_samples": keep, "include_t0": keep,
                          "seconds": best, "value": nb * ns / best, "unit": "particle-steps/s",
                          "h2d_bytes": xc.numel() * 4, "d2h_bytes": out.numel() * 4})

    # ---- secondary metric of BASELINE.json: score-matching train samples/s (SSM is the reference's live loss) -----
    data_host = prob["data"]
    data_dev = data_host.to(dev)
    gen.train()
    train = {"metric": "ssm_train_samples_per_sec", "unit": "samples/s",
             "precision": "fp32 = reference arithmetic on CUDA cores (loss 1e-7 / gradients 1e-6 of the reference); f16tc = "
                          "tcgen05 step, fp16 operands / fp32 accumulation (loss 2e-3 / gradients 3e-3, stated tolerance)",
             "step": "gen.ssm(x).mean().backward(); flat-grad all-reduce (N>1); Adam step", "runs": []}
    legs = ((256, False, "fp32"), (256, True, "fp32"), (16384, False, "fp32"), (16384, True, "fp32"),
            (256, True, "f16tc"), (16384, True, "f16tc"), (65536, True, "f16tc"))
    for batch, graphed, prec in (() if args.no_train else legs):
        v_, ms_, launches_, loss_ = time_gpu_train(P, gen, data_dev, batch, 100 if graphed else 20, world, dev, graphed,
                                                   prec)
        train["runs"].append({"batch_per_gpu": batch, "mode": "cuda_graph" if graphed else "eager", "precision": prec,
                              "value": v_, "ms_per_iter": ms_, "gpu_launches_per_iter": launches_, "loss": loss_})
    gen.ssm_precision = "fp32"
    train["value"] = max((r["value"] for r in train["runs"]), default=None)
    train["value_fp32"] = max((r["value"] for r in train["runs"] if r["precision"] == "fp32"), default=None)

    if rank == 0:
        fl = flop_per_particle_step(args.dim, 1, True) * B * N  # per launch
        kms = sum(kern_ms) / len(kern_ms)
        achieved = fl / (kms / 1e3) / 1e12
        line = {
            "metric": "reverse_sde_particle_steps_per_sec", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f16",
            "data": "synthetic", "config": workload_config(args, "gpu"),
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": B * args.dim * 4,
                    "d2h_bytes_per_step": B * args.dim * 4},
            "e2e_stock": stock,
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"],
                         "traffic": load_traffic(workload_config(args, "gpu")["workload"], args.precision),
                         "algorithmic_bytes": 8 * args.dim * B, "peak_source": peaks["src"],
                         "kernel": "sample_fp32_kernel" if args.precision == "fp32" else "sample_tc_kernel",
                         "flop_per_launch": fl, "kernel_ms": kms,
                         # the unit that actually binds the tcgen05 sampler: one MUFU.TANH per hidden activation
                         # (3 x 128 per net evaluation, 4 evaluations per RK4 step) at 16 / clk / SM; measured XU-pipe
                         # utilisation in profiles/ncu_raw_sample_tc_d*_r02_baseline_pipes.csv
                         "binding_unit": None if args.precision == "fp32" else {
                             "pipe": "xu (MUFU)", "ops_per_particle_step": 4 * 384,
                             "achieved_ops_per_s": 4 * 384 * B * N / (kms / 1e3),
                             "peak_ops_per_s": 148 * 16 * 1.965e9,
                             "frac": 4 * 384 * B * N / (kms / 1e3) / (148 * 16 * 1.965e9)}},
        }
        line.update(extra)
        line["train"] = train
        if world == 1 and not args.no_unet:
            try:
                line["unet"] = time_unet_forward(P, dev)
            except Exception as exc:  # a secondary leg must never cost the headline line
                line["unet"] = {"error": f"{type(exc).__name__}: {exc}"}
        if world > 1:  # the CPU baselines are timed at N = 1 only (torchrun also pins every rank to one host thread)
            line["cpu_baseline"] = {"value": None, "unit": "particle-steps/s", "cores": None, "kind": "port",
                                    "sample": "timed at N=1 only"}
        elif not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            kind, cpu_sample, cpu_train = cpu_objects(prob)
            what = "unmodified reference from baseline/_ref" if kind == "reference" else "oracle port of the reference's op sequence"
            train["cpu_baseline"] = {"value": None if args.no_train else cpu_train(256, 10), "unit": "samples/s",
                                     "cores": torch.get_num_threads(), "kind": kind,
                                     "sample": f"batch 256, 10 iterations after 1 warm-up ({what}, torch CPU fp32)"}
            nb, ns = cpu_sample_size(args)
            best = min(cpu_sample(nb, ns) for _ in range(3))
            line["cpu_baseline"] = {"value": nb * ns / best, "unit": "particle-steps/s", "cores": torch.get_num_threads(),
                                    "kind": kind,
                                    "sample": f"{nb} particles x {ns} RK4 steps, best of 3, same net/SDE ({what}, "
                                              "torch CPU fp32)"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
