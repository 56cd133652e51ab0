"""A reduced run of the reference's DEFAULT experiment (MSGM_higherDim.py defaults: swissroll, d = 2, dense multiplicative
SDE, MLP score net + NormalizeLogRadius, batch 256, Adam lr 1e-3, N_fwd = 16; 10 000 particles, RK4, 128 reverse steps,
norm_correction) end to end on the GPU path -- train K iterations with the fused SSM kernels, sample with the fused
sampler (fp32 and tcgen05 modes), evaluate the moments and the MMD (quantitative_comparison.compute_mmd) -- against the
same K iterations of the UNMODIFIED reference on CPU for two training seeds (tests/golden/make_experiment_golden.py).

The RNG streams differ (in-kernel Philox vs torch's CPU generator), so the comparison is statistical: every metric of the
GPU run must lie within the reference's own seed-to-seed spread widened by the stated margins.
"""
import os

import numpy as np
import pytest
import torch

import sdeflow_light_b200 as P
from sdeflow_light_b200 import quantitative_comparison as QC
from oracle import msgm_oracle as O
from tests import _build as Bd

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("kind", ["msgm", "sgm"])  # the driver runs both: MSGMs = [0, 1] (MSGM_higherDim.py:154-157)
def test_reduced_default_experiment_matches_reference(kind):
    fix = "x01_experiment_swissroll_msgm.npz" if kind == "msgm" else "x02_experiment_swissroll_sgm.npz"
    arr = {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(GOLD, fix)).items()}
    msgm = kind == "msgm"
    K = int(arr["meta_K"])
    np.random.seed(0)
    x_init, xtest = O.swiss_roll(20000), O.swiss_roll(10000)
    assert abs(float(x_init.double().sum()) - float(arr["xinit_sum"])) < 1e-6, "training data differ from the fixture's"
    assert abs(float(xtest.double().sum()) - float(arr["xtest_sum"])) < 1e-6
    T = Bd.T_param(1.0)
    if msgm:
        base = P.MSGMsde(x_init, beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=True, norm_sampler="ecdf",
                         norm_map="log", num_steps_forward=16, device=DEV, estim_cst_norm_dens_r_T=False)
        base.G, base.L_G = arr["G"].to(DEV), arr["L_G"].to(DEV)
    else:
        base = P.SGMsde(beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, num_steps_forward=16, device=DEV)
    net = P.MLP(2, premodule="NormalizeLogRadius" if msgm else None)
    net.load_state_dict({k[4:]: v for k, v in arr.items() if k.startswith("sd0.")})
    gen = P.PluginReverseSDE(base, net.to(DEV), T, deviceReverseSDE=DEV).to(DEV)
    opt = torch.optim.Adam(gen.parameters(), lr=1e-3)
    gen.train()
    torch.manual_seed(1)
    data_dev = x_init.to(DEV)
    for it in range(K):  # the reference driver's loop verbatim (MSGM_higherDim.py:803-809)
        opt.zero_grad()
        x = data_dev[torch.randint(0, data_dev.shape[0], (256,), device=DEV)]
        gen.ssm(x).mean().backward()
        opt.step()
    gen.eval()
    xt = xtest[:4000].to(DEV)
    ref_mmd = torch.stack([arr["run0.mmd"], arr["run1.mmd"]]).float()
    ref_mean = torch.stack([arr["run0.mean"], arr["run1.mean"]]).float()
    ref_cov = torch.stack([arr["run0.cov"], arr["run1.cov"]]).float()
    ref_rq = torch.stack([arr["run0.radius_q"], arr["run1.radius_q"]]).float()
    spread = lambda t_: (t_[0] - t_[1]).abs()  # noqa: E731
    for prec in ("fp32", "f16tc"):
        torch.manual_seed(5)
        x0 = gen.latent_sample(10000, 2)
        xs = P.rk4_stratonovich_sampler(gen, x0, 128, lmbd=0.0, keep_all_samples=False, norm_correction=msgm, precision=prec,
                                        seed=17)
        assert torch.isfinite(xs).all()
        mmd = float(QC.compute_mmd(xs[:4000].to(DEV), xt))
        mean, cov = xs.mean(0), torch.cov(xs.T)
        rq = torch.quantile(xs.norm(dim=1), torch.tensor([0.1, 0.5, 0.9]))
        Bd.report(test=f"experiment-{kind}-{prec}", K=K, mmd=mmd, ref_mmd=[float(v) for v in ref_mmd],
                  mean=[float(v) for v in mean], cov=[float(v) for v in cov.flatten()])
        # generated samples are as close to the data as the reference's (MMD of two data halves = the noise floor)
        assert mmd < float(ref_mmd.max()) + 3 * float(spread(ref_mmd)) + 2e-3
        # first and second moments, radial quantiles: inside the reference's seed-to-seed spread + 10 % of the data scale
        scale = float(arr["data_cov"].diagonal().max()) ** 0.5
        assert float((mean - ref_mean.mean(0)).abs().max()) < 3 * float(spread(ref_mean).max()) + 0.1 * scale
        assert float((cov - ref_cov.mean(0)).abs().max()) < 3 * float(spread(ref_cov).max()) + 0.1 * scale ** 2
        assert float((rq - ref_rq.mean(0)).abs().max()) < 3 * float(spread(ref_rq).max()) + 0.1 * scale
