"""GPU parity of prior sampling and the MMD metric against the reference's golden fixture and the oracle."""
import pytest
import torch

import sdeflow_light_b200 as P
from sdeflow_light_b200 import quantitative_comparison as QC
from oracle import msgm_oracle as O
from tests import _build as Bd
from tests import _golden as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_latent_sample_replays_reference_draws():
    meta, arr = G.load("misc_latent_mmd")
    T = Bd.T_param(1.0)
    base = P.MSGMsde(torch.randn(8, 2), T=T, norm_map="log", device=DEV, estim_cst_norm_dens_r_T=False)
    base.r_T = arr["r_T"].to(DEV)
    x0 = base.latent_sample(512, 2, U=arr["U"], Z=arr["Z"])
    err = float((x0.cpu() - arr["x0"]).abs().max())
    Bd.report(test="latent-sample-replay", max_abs=err)
    assert err < 2e-5 * float(arr["x0"].abs().max())


def test_latent_sample_statistics_and_sharding():
    torch.manual_seed(0)
    d, B = 5, 400_000
    data = torch.randn(20_000, d) * torch.linspace(0.5, 2.0, d)
    T = Bd.T_param(1.0)
    base = P.MSGMsde(data, T=T, norm_map="log", device=DEV, estim_cst_norm_dens_r_T=False)
    x = base.latent_sample(B, d, seed=11)
    a = base.latent_sample(B // 4, d, seed=11)
    b = base.latent_sample(B - B // 4, d, seed=11, particle_offset=B // 4)
    assert torch.equal(x, torch.cat([a, b]))  # keyed by the global particle index
    r = x.norm(dim=1)
    ref_r = data.norm(dim=1)
    qs = torch.tensor([0.1, 0.5, 0.9])
    assert float((torch.quantile(r.cpu(), qs) - torch.quantile(ref_r, qs)).abs().max()) < 0.02  # same radius law
    s = x / r[:, None]
    assert float(s.mean(0).abs().max()) < 5e-3 and float((torch.cov(s.T) - torch.eye(d, device=DEV) / d).abs().max()) < 5e-3
    sg = P.SGMsde(T=T, device=DEV)
    z = sg.latent_sample(B, d, seed=3)
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.var()) - 1) < 1e-2


def test_mmd_matches_reference_and_oracle():
    meta, arr = G.load("misc_latent_mmd")
    got = float(QC.compute_mmd(arr["mmd_a"].to(DEV), arr["mmd_b"].to(DEV)))
    assert abs(got - float(arr["mmd"])) < 2e-6
    torch.manual_seed(1)
    for n, m, d in [(1000, 777, 2), (300, 300, 40), (65, 130, 1)]:
        a, b = torch.randn(n, d), torch.randn(m, d) * 1.3 + 0.2
        ref = float(O.compute_mmd(a, b))
        got = float(QC.compute_mmd(a.to(DEV), b.to(DEV)))
        Bd.report(test=f"mmd-{n}x{m}x{d}", ref=ref, got=got)
        assert abs(got - ref) < 2e-6
    big = torch.randn(20_000, 2, device=DEV)  # the reference would need 2 x 3.2 GB broadcasts here
    assert abs(float(QC.compute_mmd(big, big))) < 1e-6


def test_kde_log_latent_pdf_and_elbo():
    """MSGMsde.log_latent_pdf / cst_log_dens on the GPU KDE kernel against the reference's sklearn values
    (SDEs.py:240,255-265,503-509), and NN.evaluate (ELBO, NN.py:123-128) against the reference's mean on the same
    test set (a Monte-Carlo estimate: agreement within 5 standard errors)."""
    meta, arr = G.load("misc_elbo")
    T = Bd.T_param(1.0)
    base = P.MSGMsde(arr["x_init"], beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=True, norm_map="log",
                     num_steps_forward=16, device=DEV, estim_cst_norm_dens_r_T=True)
    assert abs(base._bandwidth - float(arr["bandwidth"])) < 1e-7
    assert float((base.r_T.cpu() - arr["r_T"]).abs().max()) < 1e-6
    cst = float(base.cst_log_dens)
    lp = base.log_latent_pdf(arr["yT"].to(DEV)).cpu()
    err = float((lp - arr["logpdf"]).abs().max())
    Bd.report(test="kde-logpdf", max_abs=err, ref_max=float(arr["logpdf"].abs().max()), cst=cst, cst_ref=float(arr["cst_log_dens"]))
    assert abs(cst - float(arr["cst_log_dens"])) < 5e-6
    assert err < 2e-5 * (1 + float(arr["logpdf"].abs().max()))
    # oracle on fresh queries, including far tails (log-sum-exp must not underflow)
    q = torch.cat([torch.randn(500, 2) * 3, torch.tensor([[1e-4, 0.0], [40.0, 40.0]])])
    ref = O.log_latent_pdf(G.oracle_objects(meta, arr)[0], q, float(arr["bandwidth"]), float(arr["cst_log_dens"]))
    got = base.log_latent_pdf(q.to(DEV)).cpu()
    assert bool(torch.isfinite(got).all()) and float(((got - ref).abs() / (1 + ref.abs())).max()) < 2e-5
    # ELBO through the package's kernels
    base.G, base.L_G = arr["G"].to(DEV), arr["L_G"].to(DEV)
    net = P.MLP(2, premodule="NormalizeLogRadius").to(DEV)
    with torch.no_grad():
        for i, l in enumerate(net.linears()):
            l.weight.copy_(arr[f"W{i}"])
            l.bias.copy_(arr[f"b{i}"])
    gen = P.PluginReverseSDE(base, net, T, deviceReverseSDE=DEV).to(DEV)
    torch.manual_seed(5)
    mean, stderr = P.NN.evaluate(gen, arr["x_test"].to(DEV))
    ref_se = float(arr["elbo_std"]) / 4096 ** 0.5
    Bd.report(test="elbo-evaluate", mean=float(mean), stderr=float(stderr), ref_mean=float(arr["elbo_mean"]), ref_stderr=ref_se)
    assert abs(float(mean) - float(arr["elbo_mean"])) < 5 * (ref_se ** 2 + float(stderr) ** 2) ** 0.5
    assert 0.5 < float(stderr) / ref_se < 2.0
