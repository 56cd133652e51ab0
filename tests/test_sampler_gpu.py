"""GPU parity of the fused samplers (through the C ABI) against the reference's golden fixtures and the oracle.

Tolerance (fp32 mode): |x_cuda - x_ref| <= ATOL + RTOL * max_c |x_ref[.., c]| on every captured state (the error of
a state vector is measured against that particle's own magnitude: the SDE couples the components).  The kernel
sums the 128-term dot products in a different order than MKL and merges g.a and g.dW into one contraction, so
agreement is at accumulated-rounding level (observed <= 2e-6 relative), not bitwise.
"""
import pytest
import torch

import sdeflow_light_b200 as P
from oracle import msgm_oracle as O
from tests import _build as Bd
from tests import _golden as G

pytestmark = pytest.mark.gpu
ATOL, RTOL = 5e-5, 5e-5
DEV = "cuda:0"


def _close(name, out, ref, atol=ATOL, rtol=RTOL):
    out, ref = out.cpu(), ref.cpu()
    assert out.shape == ref.shape, (out.shape, ref.shape)
    err = (out - ref).abs()
    Bd.report(test=name, max_abs=float(err.max()), ref_max=float(ref.abs().max()), atol=atol, rtol=rtol)
    assert torch.isfinite(out).all()
    scale = ref.abs().amax(dim=-1, keepdim=True) if ref.dim() > 1 else ref.abs()
    assert bool((err <= atol + rtol * scale).all()), f"{name}: max abs err {float(err.max()):.3e}"


@pytest.mark.parametrize("name", G.names("s"))
def test_golden_fixture(name):
    meta, arr = G.load(name)
    if meta["forward"]:
        base, T = Bd.base_from(meta, arr, DEV)
        proc = P.forward_SDE(base, T.to(DEV))
    else:
        _, _, proc = Bd.gen_from(meta, arr, DEV)
    keep = arr.get("samplesToKeep")
    out = Bd.SAMPLERS[meta["scheme"]](
        proc, arr["x0"].to(DEV), meta["num_steps"], lmbd=meta["lmbd"], keep_all_samples=meta["keep_all"],
        samplesToKeep=keep, include_t0=meta["include_t0"], T_=-1 if meta["T_"] is None else torch.tensor([meta["T_"]]),
        norm_correction=meta["norm_correction"], noise=arr["noise"])
    assert out.device.type == "cpu"  # reference samplers return CPU tensors
    _close(name, out, arr["out"])


CASES = [  # kind, d, premodule, scheme, lmbd, norm_correction, B (ragged vs the 64-particle tile), N
    ("msgm_dense", 2, True, "rk4", 0.0, True, 1000, 32),
    ("msgm_dense", 3, True, "heun", 0.25, True, 130, 16),
    ("msgm_dense", 5, False, "em", 0.5, True, 77, 16),   # nc=False explodes to 1e12 in the reference itself
    ("msgm_dense", 8, True, "rk4", 0.0, True, 200, 12),
    ("msgm_dense", 16, True, "rk4", 0.0, True, 129, 8),
    ("msgm_dense", 24, True, "rk4", 0.5, True, 65, 4),
    ("msgm_dense", 32, True, "heun", 0.0, True, 64, 4),
    ("msgm_sparse", 2, True, "rk4", 0.0, True, 100, 16),
    ("msgm_sparse", 7, True, "em", 0.3, True, 100, 16),
    ("msgm_sparse", 32, True, "rk4", 0.0, True, 70, 8),
    ("sgm", 2, False, "rk4", 0.0, False, 1000, 32),
    ("sgm", 32, False, "em", 0.5, False, 63, 16),
    ("sgm", 1, False, "heun", 0.0, False, 5, 16),
]


@pytest.mark.parametrize("kind,d,pre,scheme,lmbd,nc,B,N", CASES)
def test_against_oracle(kind, d, pre, scheme, lmbd, nc, B, N):
    torch.manual_seed(1000 + d)
    if kind == "sgm":
        sde = O.make_sgm(d)
    else:
        sde = O.make_msgm(torch.randn(256, d) * 1.5, dense=(kind == "msgm_dense"))
    mlp = O.init_mlp(d, pre, seed=d, scale=6.0)
    x0 = torch.randn(B, d) * 1.3
    noise = torch.randn(N, B, d)
    ref = O.integrate(O.OReverse(sde, mlp), x0, N, scheme, lmbd, True, None, True, None, nc, noise=noise)
    _, _, gen = Bd.from_oracle(sde, mlp, DEV)
    out = Bd.SAMPLERS[scheme](gen, x0.to(DEV), N, lmbd=lmbd, keep_all_samples=True, include_t0=True,
                              norm_correction=nc, noise=noise)
    _close(f"oracle-{kind}-d{d}-{scheme}", out, ref)


@pytest.mark.parametrize("kind,d", [("msgm_dense", 4), ("msgm_sparse", 9), ("sgm", 3)])
@pytest.mark.parametrize("scheme", ["em", "heun", "rk4"])
def test_forward_adapter_against_oracle(kind, d, scheme):
    torch.manual_seed(5)
    sde = O.make_sgm(d) if kind == "sgm" else O.make_msgm(torch.randn(64, d), dense=(kind == "msgm_dense"))
    x0, noise = torch.randn(90, d), torch.randn(16, 90, d)
    ref = O.integrate(O.OForward(sde), x0, 16, scheme, 0.0, False, None, False, None, False, noise=noise)
    _, _, fwd = Bd.from_oracle(sde, None, DEV)
    out = Bd.SAMPLERS[scheme](fwd, x0.to(DEV), 16, keep_all_samples=False, noise=noise)
    _close(f"fwd-{kind}-{scheme}", out, ref)


def test_empty_batch_and_errors():
    sde, mlp = O.make_sgm(2), O.init_mlp(2, False, seed=1)
    _, _, gen = Bd.from_oracle(sde, mlp, DEV)
    out = P.rk4_stratonovich_sampler(gen, torch.zeros(0, 2, device=DEV), 4, keep_all_samples=False)
    assert out.shape == (0, 2)
    with pytest.raises(ValueError, match="len\\(samplesToKeep\\) must correspond to batch size"):
        P.rk4_stratonovich_sampler(gen, torch.zeros(5, 2, device=DEV), 4, keep_all_samples=False,
                                   samplesToKeep=torch.zeros(3, dtype=torch.int))


def test_philox_sharding_invariance_and_radius():
    """Size-independent properties at a realistic size: (i) any split of the particle set over ranks gives
    bit-identical results because noise is keyed by the global particle id; (ii) norm_correction pins |x|."""
    torch.manual_seed(3)
    d, B, N = 2, 100_003, 24
    sde = O.make_msgm(torch.randn(512, d), dense=True)
    mlp = O.init_mlp(d, True, seed=3, scale=6.0)
    _, _, gen = Bd.from_oracle(sde, mlp, DEV)
    x0 = (torch.randn(B, d) * 2).to(DEV)
    kw = dict(keep_all_samples=False, norm_correction=True, seed=1234, device_out=True)
    full = P.rk4_stratonovich_sampler(gen, x0, N, **kw)
    cut = 37_111
    a = P.rk4_stratonovich_sampler(gen, x0[:cut], N, particle_offset=0, **kw)
    b = P.rk4_stratonovich_sampler(gen, x0[cut:], N, particle_offset=cut, **kw)
    assert torch.equal(full, torch.cat([a, b]))
    r0, r1 = x0.norm(dim=1), full.norm(dim=1)
    assert float(((r1 - r0).abs() / r0).max()) < 1e-5
    other = P.rk4_stratonovich_sampler(gen, x0, N, **{**kw, "seed": 99})
    assert not torch.equal(full, other)


def test_philox_noise_is_standard_normal():
    """Forward SGM with beta -> sqrt(beta) dW only: one EM step of the forward SDE exposes the in-kernel normals."""
    d, B = 4, 400_000
    base = P.SGMsde(beta_min=1.0, beta_max=1.0, T=Bd.T_param(1.0), device=DEV)
    fwd = P.forward_SDE(base, base.T.to(DEV))
    x = P.euler_maruyama_sampler(fwd, torch.zeros(B, d, device=DEV), 1, keep_all_samples=False, seed=7,
                                 device_out=True)  # x = sqrt(beta) sqrt(delta) xi = xi
    assert abs(float(x.mean())) < 5e-3 and abs(float(x.var()) - 1.0) < 1e-2
    assert abs(float((x ** 4).mean()) - 3.0) < 0.1
    c = torch.corrcoef(x.T)
    assert float((c - torch.eye(d, device=DEV)).abs().max()) < 1e-2


def test_mlp_forward_kernel():
    for d, pre in [(2, True), (16, True), (32, False), (5, False)]:
        mlp = O.init_mlp(d, pre, seed=d, scale=3.0)
        y, s = torch.randn(333, d) * 2, torch.rand(333)
        ref = mlp(y, s)
        meta = dict(dim=d, premodule=pre)
        net = Bd.net_from(meta, {**{f"W{i}": mlp.W[i] for i in range(4)}, **{f"b{i}": mlp.b[i] for i in range(4)}}, DEV)
        with torch.no_grad():
            out = net(y.to(DEV), s.to(DEV))
        _close(f"mlp-fwd-d{d}", out, ref, 2e-5, 2e-5)


# ---- tensor-core (tcgen05, fp16 operands / fp32 accumulate) mode ---------------------------------------------------
# Stated tolerance of the f16tc mode: the hidden layers round weights and activations to fp16 (11-bit mantissa), so
# per-stage score values agree to ~1e-3 relative; trajectories are compared at TC_ATOL + TC_RTOL max_c|x_c| after N steps.
TC_ATOL, TC_RTOL = 3e-3, 3e-3  # observed: <= 7e-4 abs on |x| ~ 5, 1.4e-4 relative on |x| ~ 1e3
TC_CASES = [
    ("msgm_dense", 2, True, "rk4", 0.0, True, 1000, 16),
    ("msgm_dense", 2, True, "em", 0.0, True, 300, 1),
    ("msgm_dense", 4, True, "heun", 0.25, True, 130, 16),
    ("msgm_dense", 8, True, "rk4", 0.0, True, 257, 12),
    ("msgm_dense", 5, False, "em", 0.5, True, 77, 16),
    ("msgm_dense", 16, True, "rk4", 0.0, True, 200, 8),
    ("msgm_dense", 11, True, "heun", 0.25, True, 130, 8),
    ("msgm_sparse", 2, True, "rk4", 0.0, True, 100, 16),
    ("msgm_sparse", 16, True, "rk4", 0.0, True, 129, 8),
    ("msgm_sparse", 7, True, "em", 0.3, True, 100, 16),
    ("sgm", 2, False, "rk4", 0.0, False, 1000, 16),
    ("sgm", 16, False, "heun", 0.5, False, 63, 8),
]


@pytest.mark.parametrize("kind,d,pre,scheme,lmbd,nc,B,N", TC_CASES)
def test_tc_against_oracle(kind, d, pre, scheme, lmbd, nc, B, N):
    torch.manual_seed(2000 + d)
    sde = O.make_sgm(d) if kind == "sgm" else O.make_msgm(torch.randn(256, d) * 1.5, dense=(kind == "msgm_dense"))
    mlp = O.init_mlp(d, pre, seed=d, scale=6.0)
    x0 = torch.randn(B, d) * 1.3
    noise = torch.randn(N, B, d)
    ref = O.integrate(O.OReverse(sde, mlp), x0, N, scheme, lmbd, True, None, True, None, nc, noise=noise)
    _, _, gen = Bd.from_oracle(sde, mlp, DEV)
    out = Bd.SAMPLERS[scheme](gen, x0.to(DEV), N, lmbd=lmbd, keep_all_samples=True, include_t0=True,
                              norm_correction=nc, noise=noise, precision="f16tc")
    assert P._lib.debug_flags(DEV) == 0, "tensor-core sampler: a bounded mbarrier wait timed out"
    _close(f"tc-{kind}-d{d}-{scheme}-N{N}", out, ref, TC_ATOL, TC_RTOL)


def test_tc_matches_fp32_statistics():
    """f16tc vs fp32 kernels with the same Philox noise: per-particle agreement and identical ensemble moments."""
    torch.manual_seed(11)
    d, B, N = 2, 200_000, 32
    sde = O.make_msgm(O.swiss_roll(4000), dense=True)
    mlp = O.init_mlp(d, True, seed=11, scale=6.0)
    _, _, gen = Bd.from_oracle(sde, mlp, DEV)
    x0 = O.latent_sample(sde, B).to(DEV)
    kw = dict(keep_all_samples=False, norm_correction=True, seed=5, device_out=True)
    a = P.rk4_stratonovich_sampler(gen, x0, N, precision="fp32", **kw)
    b = P.rk4_stratonovich_sampler(gen, x0, N, precision="f16tc", **kw)
    assert P._lib.debug_flags(DEV) == 0
    err = (a - b).abs()
    Bd.report(test="tc-vs-fp32-200k", max_abs=float(err.max()), mean_abs=float(err.mean()), ref_max=float(a.abs().max()))
    assert float(err.mean()) < 5e-3
    assert float((a.mean(0) - b.mean(0)).abs().max()) < 2e-3
    assert float((torch.cov(a.T) - torch.cov(b.T)).abs().max()) < 5e-3


def _bench_problem(d):
    """The bench's own synthetic problem (bench.build_problem) as (oracle OSde, OMlp, package gen)."""
    import bench
    prob = bench.build_problem(d)
    sde = O.OSde("msgm_dense", d, prob["beta_min"], prob["beta_max"], prob["T"], prob["t_epsilon"],
                 prob["num_steps_forward"], G=prob["G"], L_G=prob["L_G"], r_T=prob["r_T"], norm_map="log")
    mlp = O.OMlp(prob["W"], prob["b"], True, d)
    return sde, mlp, bench.package_objects(prob, DEV)[1]


@pytest.mark.parametrize("d", [8, 2])
def test_tc_headline_configuration_against_oracle(d):
    """The bench's headline configuration itself -- its weights, G and data, dense multiplicative SDE, RK4, 128 reverse
    steps, lambda = 0, radius correction -- on 4096 particles with injected noise: EVERY one of the 128 states of the
    tcgen05 (f16tc) path against the CPU oracle at the stated f16tc tolerance, and the fp32 parity mode at its own."""
    sde, mlp, gen = _bench_problem(d)
    torch.manual_seed(77)
    B, N = 4096, 128
    x0 = torch.randn(B, d) * 1.5
    noise = torch.randn(N, B, d)
    ref = O.integrate(O.OReverse(sde, mlp), x0, N, "rk4", 0.0, True, None, True, None, True, noise=noise)
    for prec, atol, rtol in (("f16tc", TC_ATOL, TC_RTOL), ("fp32", 5e-5, 5e-5)):
        out = P.rk4_stratonovich_sampler(gen, x0.to(DEV), N, lmbd=0.0, keep_all_samples=True, include_t0=True,
                                         norm_correction=True, noise=noise, precision=prec)
        assert P._lib.debug_flags(DEV) == 0
        assert tuple(out.shape) == (N + 1, B, d)
        _close(f"headline-d{d}-N128-{prec}", out, ref, atol, rtol)
        # the error must not grow into the tolerance over the 128 steps: last state on its own
        _close(f"headline-d{d}-N128-{prec}-final", out[-1], ref[-1], atol, rtol)


def test_tc_vs_fp32_at_bench_size():
    """2^20 particles x 128 RK4 steps, in-kernel Philox (the bench's config 2): the f16tc and fp32 kernels draw the same
    noise, so particles agree one by one within the f16tc tolerance, and the ensembles agree in mean, covariance, radius
    (pinned by norm_correction) and MMD."""
    from sdeflow_light_b200 import quantitative_comparison as Q
    d, B, N = 8, 1 << 20, 128
    _, _, gen = _bench_problem(d)
    torch.manual_seed(5)
    x0 = (torch.randn(B, d) * 1.5).to(DEV)
    kw = dict(keep_all_samples=False, norm_correction=True, seed=2024, device_out=True)
    b = P.rk4_stratonovich_sampler(gen, x0, N, precision="f16tc", **kw)
    a = P.rk4_stratonovich_sampler(gen, x0, N, precision="fp32", **kw)
    assert P._lib.debug_flags(DEV) == 0
    err = (a - b).abs()
    tol = TC_ATOL + TC_RTOL * a.abs().max(dim=1, keepdim=True)[0]
    frac_out = float((err > tol).float().mean())
    Bd.report(test="tc-vs-fp32-2p20-N128", max_abs=float(err.max()), mean_abs=float(err.mean()),
              ref_max=float(a.abs().max()), frac_outside_tol=frac_out)
    assert frac_out < 1e-4, frac_out          # chaotic outliers only (none observed)
    assert float(err.mean()) < 5e-4
    assert float((a.mean(0) - b.mean(0)).abs().max()) < 1e-3
    assert float((torch.cov(a.T) - torch.cov(b.T)).abs().max()) < 2e-3
    assert float(((b.norm(dim=1) - x0.norm(dim=1)).abs() / x0.norm(dim=1)).max()) < 1e-5
    idx = torch.randperm(B, device=DEV)[:20000]
    mmd_ab = float(Q.compute_mmd(a[idx], b[idx]))
    mmd_aa = float(Q.compute_mmd(a[idx], a[torch.randperm(B, device=DEV)[:20000]]))
    Bd.report(test="tc-vs-fp32-2p20-N128-mmd", mmd_tc_vs_fp32=mmd_ab, mmd_fp32_vs_fp32_resample=mmd_aa)
    assert mmd_ab < 1e-5 and mmd_ab < mmd_aa
