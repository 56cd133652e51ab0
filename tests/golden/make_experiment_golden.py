"""Golden numbers for a REDUCED run of the reference's default experiment (MSGM_higherDim.py defaults: swissroll, d = 2,
dense multiplicative SDE, MLP score net with NormalizeLogRadius, batch 256, Adam lr 1e-3, N_fwd = 16; sampling: 10 000
particles, RK4, 128 steps, lmbd 0, norm_correction) -- TEST INFRASTRUCTURE, run in the build container only:

    python -m tests.golden.make_experiment_golden            # ~8 min of CPU: trains the UNMODIFIED reference twice
    MSGM_EXPERIMENT_KIND=sgm python -m tests.golden.make_experiment_golden    # the additive baseline (x02_...)

The reference's 2^20 iterations are out of reach for a fixture, so the run is cut to K iterations; the fixture stores the
initial weights, G / L_G, and for two training seeds the moments of the generated samples, their MMD to held-out data
(quantitative_comparison.compute_mmd) and the loss on a fixed evaluation set, so that the GPU test can train the
drop-in package for the same K iterations and require agreement within the reference's own seed-to-seed spread.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import msgm_oracle as O  # noqa: E402
from oracle import ref_live  # noqa: E402

K = int(os.environ.get("MSGM_EXPERIMENT_ITERS", "3000"))
N_INIT, N_TEST, N_GEN, N_BACK, BATCH = 20000, 10000, 10000, 128, 256


def run(ref, x_init, xtest, sd0, G, LG, seed, msgm=True):
    torch.manual_seed(seed)
    random.seed(seed)
    if msgm:
        base, gen, net = ref_live.build(ref, "msgm_dense", 2, x_init=x_init, premodule="NormalizeLogRadius")
        base.G, base.L_G = G.clone(), LG.clone()
    else:  # the driver's additive baseline: SGMsde, MLP without premodule, no radius correction (MSGM_higherDim.py:716-746)
        base, gen, net = ref_live.build(ref, "sgm", 2, premodule=None)
    net.load_state_dict(sd0)
    opt = torch.optim.Adam(gen.parameters(), lr=1e-3)
    gen.train()
    xe = xtest[:2048]
    torch.manual_seed(1000 + seed)
    te, _, ye = gen.sample_txy(xe)

    def eval_loss():  # the reference draws the Hutchinson probe inside ssm_loss: fix it through the RNG state
        torch.manual_seed(4242)
        return float(gen.ssm_loss(te, xe, ye.clone().requires_grad_()).mean().detach())

    l0 = eval_loss()
    torch.manual_seed(seed)
    for it in range(K):
        opt.zero_grad()
        x = x_init[torch.randint(0, x_init.shape[0], (BATCH,))]
        gen.ssm(x).mean().backward()
        opt.step()
    l1 = eval_loss()
    gen.eval()
    with torch.no_grad():
        x0 = gen.latent_sample(N_GEN, 2)
        xs = ref.sde_scheme.rk4_stratonovich_sampler(gen, x0, N_BACK, lmbd=0.0, keep_all_samples=False,
                                                     norm_correction=msgm)
    mmd = float(ref.qc.compute_mmd(xs[:4000], xtest[:4000]))
    return dict(mean=xs.mean(0).numpy(), cov=torch.cov(xs.T).numpy(), mmd=mmd, loss0=l0, loss1=l1,
                radius_q=torch.quantile(xs.norm(dim=1), torch.tensor([0.1, 0.5, 0.9])).numpy())


def main():
    ref = ref_live.load()
    np.random.seed(0)
    torch.manual_seed(0)
    random.seed(0)
    x_init, xtest = O.swiss_roll(N_INIT), O.swiss_roll(N_TEST)
    base, gen, net = ref_live.build(ref, "msgm_dense", 2, x_init=x_init, premodule="NormalizeLogRadius")
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    G, LG = base.G.clone(), base.L_G.clone()
    out = {"meta_K": np.array(K), "G": G.numpy(), "L_G": LG.numpy(), "xinit_sum": np.array(float(x_init.double().sum())),
           "xtest_sum": np.array(float(xtest.double().sum())),
           "data_mean": xtest.mean(0).numpy(), "data_cov": torch.cov(xtest.T).numpy(),
           "mmd_data_data": np.array(float(ref.qc.compute_mmd(x_init[:4000], xtest[:4000])))}
    for k, v in sd0.items():
        out["sd0." + k] = v.numpy()
    if os.environ.get("MSGM_EXPERIMENT_KIND", "msgm") == "sgm":
        torch.manual_seed(0)
        _, _, net_s = ref_live.build(ref, "sgm", 2, premodule=None)
        sd0 = {k: v.clone() for k, v in net_s.state_dict().items()}
        for k in [k for k in out if k.startswith("sd0.")] + ["G", "L_G"]:
            out.pop(k, None)
        for k, v in sd0.items():
            out["sd0." + k] = v.numpy()
        name, msgm = "x02_experiment_swissroll_sgm.npz", False
    else:
        name, msgm = "x01_experiment_swissroll_msgm.npz", True
    for i, seed in enumerate((1, 2)):
        r = run(ref, x_init, xtest, sd0, G, LG, seed, msgm)
        print(f"seed {seed}: mmd {r['mmd']:.5f} loss {r['loss0']:.4f} -> {r['loss1']:.4f} mean {r['mean']} cov {r['cov'].ravel()}",
              flush=True)
        for k, v in r.items():
            out[f"run{i}.{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name), **out)
    print("written", name)


if __name__ == "__main__":
    main()
