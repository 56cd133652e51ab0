"""GPU parity of the evaluation metrics of SURVEY.md 8f2 (survival curve, tail exponent, covariance / energy report)
against fixture m01 written by the unmodified reference's own_plotting.py (tests/golden/make_metrics_golden.py), and of
the one-launch Adam update (msgm_adam_step) against torch.optim.Adam."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from sdeflow_light_b200 import _lib, sample_metrics as M

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture():
    z = np.load(os.path.join(GOLD, "m01_survival_cov_energy.npz"))
    return json.loads(str(z["meta"])), z


@pytest.mark.parametrize("tag", ["plain", "scaled"])
def test_survival_curve_and_tail_exponent_against_reference(tag):
    meta, z = _fixture()
    x_ref, x_gen = torch.from_numpy(z["x_ref"]).to(DEV), torch.from_numpy(z["x_gen"]).to(DEV)
    sn = torch.from_numpy(z["std_norm"]) if tag == "scaled" else None
    out = M.survival_curves(x_gen, x_ref, std_norm=sn, n_points=meta["n_points"], tail_frac=meta["tail_frac"])
    # norms: fp32, within one ulp of torch.norm
    norms, lo, hi = M._norms(x_gen, sn)
    assert float((norms.cpu() - torch.from_numpy(z[f"{tag}_norms_gen"])).abs().max() / hi) < 2e-7
    # the grid is built from fp32 extrema through float64 log10 / logspace: a last-bit difference of an extremum moves it
    # by 1e-7 relative
    np.testing.assert_allclose(out["R_grid"], z[f"{tag}_R"], rtol=3e-7)
    for name, key in (("reference", "ref"), ("generated", "gen")):
        counts, ref_counts = out[name]["counts"], z[f"{tag}_c_{key}"]
        # integer counts: identical except where a norm sits within an ulp of a grid point (at most one particle, rarely)
        diff = np.abs(counts - ref_counts)
        assert diff.max() <= 1 and (diff > 0).mean() < 0.03, (diff.max(), (diff > 0).mean())
        np.testing.assert_allclose(out[name]["S"], z[f"{tag}_S_{key}"], atol=1.0 / out[name]["N"] + 1e-12)
    a_ref, a_gen, _ = z[f"{tag}_alpha"]
    assert out["fits"]["ref"]["k"] == int(z[f"{tag}_k"][0]) and out["fits"]["gen"]["k"] == int(z[f"{tag}_k"][1])
    assert abs(out["fits"]["ref"]["alpha"] - a_ref) < 2e-3 * abs(a_ref)
    assert abs(out["fits"]["gen"]["alpha"] - a_gen) < 2e-3 * abs(a_gen)
    outk = M.survival_curves(x_gen, x_ref, std_norm=sn, n_points=meta["n_points"], tail_k=meta["tail_k"])
    assert outk["fits"]["gen"]["k"] == int(z[f"{tag}_k"][2])
    assert abs(outk["fits"]["gen"]["alpha"] - z[f"{tag}_alpha"][2]) < 2e-3 * abs(z[f"{tag}_alpha"][2])


def test_survival_counts_bit_exact_on_given_norms():
    """Same norms, same grid -> the integer counts are those of numpy's sort + searchsorted, bit for bit; includes ties on
    grid points, zeros and a grid point beyond the maximum."""
    rng = np.random.default_rng(0)
    norms = np.abs(rng.standard_normal(300_001)).astype(np.float32)
    norms[:50] = 0.0
    R = np.logspace(-3, 1, 257)
    R[40] = float(norms[1000])   # exact ties
    R[41] = float(norms[2000])
    R = np.sort(R)
    ref = norms.size - np.searchsorted(np.sort(norms), R, side="right")
    S, counts = M._empirical_survival_from_norms(torch.from_numpy(norms).to(DEV), R)
    assert np.array_equal(counts, ref)
    assert counts[-1] == 0 and counts[0] <= norms.size - 50


def test_covariance_energy_report_against_reference():
    meta, z = _fixture()
    x_ref = torch.from_numpy(z["x_ref"]).to(DEV)
    x_fwd = torch.from_numpy(z["x_gen"][:meta["n_ref"]]).to(DEV)
    rep = M.covariance_energy_report(x_ref, x_fwd)
    assert float((rep["cov_xtest"] - torch.from_numpy(z["cov_ref"]).double()).abs().max()) < 2e-5
    assert float((rep["cov_xgen_forward"] - torch.from_numpy(z["cov_gen"]).double()).abs().max()) < 2e-4
    for k, v in meta["printed"].items():
        assert abs(rep[k] - v) < 2e-6 * max(1.0, abs(v)) + 2e-6, (k, rep[k], v)


def test_moments_wide_state():
    """d = 1024 (U-Net configurations): Gram / mean against float64 torch on a small batch."""
    torch.manual_seed(1)
    x = torch.randn(700, 1024) * torch.linspace(0.5, 2.0, 1024) + 0.3
    n, mean, cov, energy = M._moments(x.to(DEV))
    xd = x.double()
    assert float((mean - xd.mean(0)).abs().max()) < 1e-6
    assert float((cov - torch.cov(xd.T)).abs().max()) < 1e-5
    assert abs(energy - float((xd ** 2).sum(1).mean())) < 1e-5 * energy


def test_adam_step_matches_torch():
    torch.manual_seed(2)
    shapes = [(128, 10), (128,), (128, 128), (128,), (7, 128), (7,)]
    params = [torch.randn(*s, device=DEV) for s in shapes]
    ref_p = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.Adam(ref_p, lr=1e-3)
    total = sum(p.numel() for p in params)
    flat = torch.zeros(total, device=DEV)
    m, v = torch.zeros(total, device=DEV), torch.zeros(total, device=DEV)
    lr, step = torch.tensor(1e-3, device=DEV), torch.zeros(1, device=DEV, dtype=torch.int64)
    table = torch.zeros(len(params), 2, dtype=torch.int64)
    o = 0
    for i, p in enumerate(params):
        table[i, 0], table[i, 1] = p.data_ptr(), o
        o += p.numel()
    table = table.to(DEV)
    h, L = _lib.ctx(DEV), _lib.lib()
    for it in range(5):
        g = [torch.randn_like(p) * (0.1 + it) for p in params]
        flat.copy_(torch.cat([t.reshape(-1) for t in g]) * 4.0)   # as if summed over 4 ranks
        _lib.check(L.msgm_adam_step(h, _lib.ptr(table), len(params), total, _lib.ptr(flat), _lib.ptr(m), _lib.ptr(v),
                                    _lib.ptr(lr), _lib.ptr(step), 0.9, 0.999, 1e-8, 0.25, _lib.stream_ptr(torch.device(DEV))))
        for q, t in zip(ref_p, g):
            q.grad = t.clone()
        opt.step()
    assert int(step.item()) == 5
    for p, q in zip(params, ref_p):
        assert float((p - q.detach()).abs().max()) < 2e-6


def test_trainer_state_dict_is_torch_adam_compatible(tmp_path):
    """The trainer's one-launch Adam keeps flat moment buffers; its state_dict() must load into torch.optim.Adam (what the
    reference's load_checkpoint builds) and back, through NN.save_checkpoint / load_checkpoint."""
    import sdeflow_light_b200 as P
    from sdeflow_light_b200 import NN
    from sdeflow_light_b200.train import GraphedSsmStep
    torch.manual_seed(3)
    d = 4
    T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
    data = torch.randn(2048, d) * 1.5

    def build():
        base = P.MSGMsde(data, beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=True, norm_map="log",
                         num_steps_forward=16, device=DEV, estim_cst_norm_dens_r_T=False)
        return P.PluginReverseSDE(base, P.MLP(d, premodule="NormalizeLogRadius").to(DEV), T, deviceReverseSDE=DEV).to(DEV)

    gen = build()
    step = GraphedSsmStep(gen, (128, d), lr=1e-3, seed=9)
    x = data[:128].to(DEV)
    for _ in range(4):
        step(x)
    sd = step.state_dict()
    assert all(int(float(s["step"])) == 4 for s in sd["state"].values()) and len(sd["state"]) == 8
    opt = torch.optim.Adam(gen.parameters(), lr=5e-4)
    opt.load_state_dict(sd)   # the reference's loader does exactly this (NN.py:27)
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-3)
    path = str(tmp_path / "ck.pt")
    NN.save_checkpoint(path, gen, step, 3, trainer=step)
    gen2 = build()
    step2 = GraphedSsmStep(gen2, (128, d), lr=7e-4, seed=9)
    assert NN.load_checkpoint(path, gen2, step2, DEV, trainer=step2) == 3
    assert torch.equal(step2.exp_avg, step.exp_avg) and torch.equal(step2.exp_avg_sq, step.exp_avg_sq)
    assert int(step2.adam_step.item()) == 4 and int(step2._iter.item()) == 4
    assert torch.equal(gen2.base_sde.G, gen.base_sde.G)
    # both continue identically: same weights, same moments, same Philox stream position
    l1, l2 = float(step(x)), float(step2(x))
    assert l1 == l2
    for p, q in zip(gen.a.parameters(), gen2.a.parameters()):
        assert float((p - q).abs().max()) < 1e-6  # the weight-gradient reduction uses atomics: last-bit differences
