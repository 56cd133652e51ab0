"""CPU: the oracle restatement replays every golden fixture of the reference (bit-level fp32 agreement).

The fixtures are outputs of the unmodified reference (tests/golden/make_golden.py); this is what pins the oracle
on machines where /root/reference does not exist.
"""
import pytest
import torch

from oracle import msgm_oracle as O
from tests import _golden as G

TOL = 2e-6  # same ATen ops in the same order; observed 0.0 in the build container


@pytest.mark.parametrize("name", G.names("s"))
def test_sampler_fixture(name):
    meta, arr = G.load(name)
    sde, mlp = G.oracle_objects(meta, arr)
    proc = O.OForward(sde) if meta["forward"] else O.OReverse(sde, mlp)
    keep = arr.get("samplesToKeep")
    out = O.integrate(proc, arr["x0"], meta["num_steps"], meta["scheme"], meta["lmbd"],
                      keep_all_samples=meta["keep_all"], samplesToKeep=keep, include_t0=meta["include_t0"],
                      T_=meta["T_"] if meta["T_"] is None else torch.tensor([meta["T_"]]).item(),
                      norm_correction=meta["norm_correction"], noise=arr["noise"])
    assert out.shape == arr["out"].shape
    assert float((out - arr["out"]).abs().max()) <= TOL


@pytest.mark.parametrize("name", G.names("t"))
def test_ssm_fixture(name):
    meta, arr = G.load(name)
    sde, mlp = G.oracle_objects(meta, arr)
    for p in mlp.parameters():
        p.requires_grad_(True)
    rev = O.OReverse(sde, mlp)
    y = arr["y"].clone().requires_grad_()
    loss = O.ssm_loss(rev, arr["t"], y, arr["v"])
    assert float((loss.detach() - arr["loss"]).abs().max()) <= TOL * max(1.0, float(arr["loss"].abs().max()))
    grads = torch.autograd.grad(loss.mean(), mlp.parameters())
    for i in range(4):
        for g, key in ((grads[2 * i], f"gW{i}"), (grads[2 * i + 1], f"gb{i}")):
            assert float((g - arr[key]).abs().max()) <= TOL * max(1.0, float(arr[key].abs().max()))


def test_latent_and_mmd_fixture():
    meta, arr = G.load("misc_latent_mmd")
    r = torch.quantile(arr["r_T"], arr["U"]).reshape(-1, 1)
    r = torch.exp(r) - 1e-6
    Z = arr["Z"]
    x0 = r * (Z / torch.linalg.norm(Z, dim=1).reshape(-1, 1))
    assert float((x0 - arr["x0"]).abs().max()) <= TOL
    assert abs(float(O.compute_mmd(arr["mmd_a"], arr["mmd_b"])) - float(arr["mmd"])) <= TOL


def test_noise_forward_replay():
    """Forward noising with the reference's RNG order (SDEs.py:78-122): replaying the stored draws gives y."""
    meta, arr = G.load("t01_ssm_msgm_d2")
    sde, _ = G.oracle_objects(meta, arr)
    fwd = O.OForward(sde)
    t, x, n_fwd = arr["t"], arr["x"], meta["num_steps_forward"]
    n_int = torch.trunc(n_fwd * t / sde.T_tensor).to(torch.int)
    cap = O.integrate(fwd, x, n_fwd, "rk4", 0.0, False, n_int, True, noise=arr["fwd_noise"])
    j = 0
    for k in range(x.shape[0]):
        if n_int[k] == 0:
            cap[k] = O.integrate(fwd, x[k][None], 1, "rk4", 0.0, False, None, False, T_=float(t[k]),
                                 noise=arr["singles"][j][None, None])[0]
            j += 1
    assert float((cap - arr["y"]).abs().max()) <= TOL


def test_kde_log_density_fixture():
    """The oracle's exact Gaussian-kernel sum against the reference's sklearn KernelDensity.score_samples and its
    normalising constant (SDEs.py:240,255-265,503-509)."""
    meta, arr = G.load("misc_elbo")
    sde, _ = G.oracle_objects(meta, arr)
    h, cst = float(arr["bandwidth"]), float(arr["cst_log_dens"])
    assert abs(h - 0.1 * float(torch.std(arr["r_T"]))) <= 1e-7
    assert abs(float(O.kde_log_normaliser(arr["r_T"], h)) - cst) <= 2e-6
    got = O.log_latent_pdf(sde, arr["yT"], h, cst)
    assert float((got - arr["logpdf"]).abs().max()) <= 2e-5 * (1 + float(arr["logpdf"].abs().max()))
