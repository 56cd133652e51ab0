"""Generate the golden fixtures in this directory from the LIVE, UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Every ``*.npz`` holds the inputs (weights, noise tensor G, initial particles, injected standard-normal draws)
and the reference's outputs for one case of the hot path, plus a ``meta`` JSON string.  The reference has no
golden vectors of its own; these are outputs of the reference itself.

Noise is recorded by RNG replay: the reference samplers draw exactly one ``randn_like(x)`` per step
(sde_scheme.py:84,144,227), so re-seeding and drawing the same shapes reproduces what they consumed.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_live  # noqa: E402


def _boost(net, k):
    with torch.no_grad():
        net.main[6].weight.mul_(k)
        net.main[6].bias.mul_(k)


def _net_arrays(net):
    lin = [m for m in net.main if isinstance(m, torch.nn.Linear)]
    out = {}
    for i, l in enumerate(lin):
        out[f"W{i}"] = l.weight.detach().numpy().copy()
        out[f"b{i}"] = l.bias.detach().numpy().copy()
    return out


def _sde_arrays(base):
    out = {}
    if type(base).__name__ == "MSGMsde":
        out["r_T"] = base.r_T.numpy().copy()
        if not base.sparseTensor:
            out["G"] = base.G.numpy().copy()
            out["L_G"] = base.L_G.numpy().copy()
    return out


def _save(name, meta, **arrays):
    arrays = {k: np.asarray(v) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta), **arrays)
    print(f"wrote {name}.npz  ({sum(a.nbytes for a in arrays.values()) / 1024:.0f} KiB raw)")


def sampler_case(ref, name, kind, dim, pre, scheme, N, B, lmbd, nc, seed, forward=False, boost=8.0,
                 beta=(0.1, 20.0), keep=None, T_=None):
    torch.manual_seed(seed)
    x_init = torch.randn(1024, dim) * 1.5
    base, gen, net = ref_live.build(ref, kind, dim, x_init, pre, beta_min=beta[0], beta_max=beta[1])
    _boost(net, boost)
    proc = ref.SDEs.forward_SDE(base, base.T) if forward else gen
    x0 = torch.randn(B, dim) * 1.2
    fn = {"em": ref.sde_scheme.euler_maruyama_sampler, "heun": ref.sde_scheme.heun_sampler,
          "rk4": ref.sde_scheme.rk4_stratonovich_sampler}[scheme]
    kw = dict(lmbd=lmbd, include_t0=True, norm_correction=nc)
    if keep is not None:
        g = torch.Generator().manual_seed(seed + 1)
        keep_t = torch.randint(0, N + 1, (B, 1), generator=g).to(torch.int)
        kw.update(keep_all_samples=False, samplesToKeep=keep_t)
    else:
        kw.update(keep_all_samples=True)
    if T_ is not None:
        kw.update(T_=torch.tensor([T_]))
    torch.manual_seed(seed + 2)
    xs = fn(proc, x0, N, **kw)
    torch.manual_seed(seed + 2)
    noise = torch.stack([torch.randn_like(x0) for _ in range(N)])
    meta = dict(kind=kind, dim=dim, premodule=pre is not None, scheme=scheme, num_steps=N, lmbd=lmbd,
                norm_correction=nc, include_t0=True, forward=forward, beta_min=beta[0], beta_max=beta[1],
                T=1.0, T_=T_, keep_all=keep is None)
    arrays = dict(x0=x0.numpy(), noise=noise.numpy(), out=xs.numpy(), **_sde_arrays(base))
    if not forward:
        arrays.update(_net_arrays(net))
    if keep is not None:
        arrays["samplesToKeep"] = keep_t.numpy()
    _save(name, meta, **arrays)


def ssm_case(ref, name, kind, dim, pre, B, seed, n_fwd=16, boost=3.0):
    torch.manual_seed(seed)
    x_init = torch.randn(1024, dim) * 1.5
    base, gen, net = ref_live.build(ref, kind, dim, x_init, pre, n_fwd=n_fwd)
    _boost(net, boost)
    x = torch.randn(B, dim)
    # --- forward noising with replayable noise (SDEs.py:648-682, 78-122)
    torch.manual_seed(seed + 1)
    t_, _, y = gen.sample_txy(x)
    state = torch.get_rng_state()
    v = ref.SDEs.sample_rademacher(x.shape, "cpu")
    torch.manual_seed(seed + 1)
    u_t = torch.rand(B, 1)
    if kind == "sgm":
        fwd_noise = torch.randn(1, B, dim)
        singles = torch.zeros(0, dim)
    else:
        fwd_noise = torch.stack([torch.randn(B, dim) for _ in range(n_fwd)])
        n_int = torch.trunc(n_fwd * t_ / base.T).to(torch.int).flatten()
        singles = torch.cat([torch.randn(1, dim) for k in range(B) if n_int[k] == 0] + [torch.zeros(0, dim)])
    # --- loss and gradients with the same v (SDEs.py:616-646)
    torch.set_rng_state(state)
    y = y.detach().clone().requires_grad_()
    gen.train()
    loss = gen.ssm_loss(t_, x, y)
    gen.zero_grad()
    loss.mean().backward()
    arrays = dict(x=x.numpy(), u_t=u_t.numpy(), t=t_.numpy(), y=y.detach().numpy(), v=v.numpy(),
                  fwd_noise=fwd_noise.numpy(), singles=singles.numpy(), loss=loss.detach().numpy(),
                  **_sde_arrays(base), **_net_arrays(net))
    lin = [m for m in net.main if isinstance(m, torch.nn.Linear)]
    for i, l in enumerate(lin):
        arrays[f"gW{i}"] = l.weight.grad.numpy().copy()
        arrays[f"gb{i}"] = l.bias.grad.numpy().copy()
    meta = dict(kind=kind, dim=dim, premodule=pre is not None, beta_min=0.1, beta_max=20.0, T=1.0,
                t_epsilon=1e-3, num_steps_forward=n_fwd, vtype="rademacher")
    _save(name, meta, **arrays)


def misc_case(ref):
    torch.manual_seed(31)
    x_init = torch.randn(4096, 2) * torch.tensor([1.5, 0.7])
    base, gen, net = ref_live.build(ref, "msgm_dense", 2, x_init, "NormalizeLogRadius")
    torch.manual_seed(32)
    x0 = gen.latent_sample(512, 2)
    torch.manual_seed(32)
    U = torch.rand(512)
    Z = torch.randn(512, 2)
    a, b = torch.randn(300, 4), torch.randn(257, 4) * 1.2 + 0.3
    mmd = ref.qc.compute_mmd(a, b)
    _save("misc_latent_mmd", dict(norm_map="log"), r_T=base.r_T.numpy(), U=U.numpy(), Z=Z.numpy(),
          x0=x0.numpy(), mmd_a=a.numpy(), mmd_b=b.numpy(), mmd=np.float32(mmd))


def elbo_case(ref):
    """ELBO pieces (NN.py:123-128, SDEs.py:495-509,708-721): the KDE log-density of the radii with its normalising
    constant (sklearn on the reference side), and the reference's ELBO mean over a large test set."""
    torch.manual_seed(61)
    x_init = torch.randn(3000, 2) * torch.tensor([1.5, 0.7])
    T = ref_live.T_param(1.0)
    base = ref.SDEs.MSGMsde(x_init, beta_min=0.1, beta_max=20.0, t_epsilon=1e-3, T=T, num_steps_forward=16,
                            device="cpu", estim_cst_norm_dens_r_T=True, norm_sampler="ecdf", norm_map="log",
                            denseTensor=True, plot_validate=False)
    net = ref.NN.MLP(input_dim=2, index_dim=1, hidden_dim=128, premodule="NormalizeLogRadius")
    gen = ref.SDEs.PluginReverseSDE(base, net, T, vtype="rademacher", debias=False, ssm_intT=False,
                                    deviceReverseSDE="cpu")
    yT = torch.randn(300, 2) * 2.0
    logpdf = base.log_latent_pdf(yT)
    x_test = torch.randn(4096, 2) * torch.tensor([1.5, 0.7])
    torch.manual_seed(62)
    gen.eval()
    with torch.enable_grad():
        elbo = gen.elbo_random_t_slice(x_test).detach()
    arrays = dict(x_init=x_init.numpy(), yT=yT.numpy(), logpdf=logpdf.numpy(), x_test=x_test.numpy(),
                  cst_log_dens=np.float32(base.cst_log_dens), bandwidth=np.float32(base.kde.bandwidth),
                  elbo_mean=np.float32(elbo.mean()), elbo_std=np.float32(elbo.std()), **_sde_arrays(base),
                  **_net_arrays(net))
    _save("misc_elbo", dict(kind="msgm_dense", dim=2, premodule=True, beta_min=0.1, beta_max=20.0, T=1.0,
                            t_epsilon=1e-3, num_steps_forward=16, norm_map="log"), **arrays)


def unet1d_case(ref, name, kind, L, B, N, seed, pre="NormalizeLogRadius"):
    """UNet1D score net (NNUnet1D.py) on 1-D signals of length L: forward, RK4 reverse sampling, SSM loss + gradients."""
    torch.manual_seed(seed)
    sig = torch.sin(torch.linspace(0, 6.28, L)[None] * torch.randint(1, 4, (256, 1)) + 6.28 * torch.rand(256, 1)) \
        + 0.1 * torch.randn(256, L)
    net = ref.NNUnet1D.UNet1D(input_dim=L, base_channels=8, channel_mults=(1, 2, 4), num_res_blocks=2,
                              premodule=pre, emb_dim=16)
    with torch.no_grad():
        net.final.weight.mul_(4.0)
    base, gen, _ = ref_live.build(ref, kind, L, sig, pre, net=net)
    x0 = sig[:B].clone() + 0.3 * torch.randn(B, L)
    s = torch.rand(B)
    with torch.no_grad():
        fwd = net(x0, s)
    torch.manual_seed(seed + 1)
    xs = ref.sde_scheme.rk4_stratonovich_sampler(gen, x0, N, lmbd=0., keep_all_samples=True, include_t0=True,
                                                 norm_correction=(kind != "sgm"))
    torch.manual_seed(seed + 1)
    noise = torch.stack([torch.randn_like(x0) for _ in range(N)])
    # SSM loss and gradients for given (t, y, v)
    t_ = torch.rand(B, 1).clamp_min(1e-3)
    y = (x0 + 0.2 * torch.randn(B, L)).requires_grad_()
    state = torch.get_rng_state()
    v = ref.SDEs.sample_rademacher(x0.shape, "cpu")
    torch.set_rng_state(state)
    gen.train()
    loss = gen.ssm_loss(t_, x0, y)
    gen.zero_grad()
    loss.mean().backward()
    arrays = dict(x0=x0.numpy(), s=s.numpy(), fwd=fwd.numpy(), noise=noise.numpy(), out=xs.numpy(), t=t_.numpy(),
                  y=y.detach().numpy(), v=v.numpy(), loss=loss.detach().numpy(), **_sde_arrays(base))
    for k, p_ in net.state_dict().items():
        arrays["sd." + k] = p_.numpy().copy()
    for k, p_ in net.named_parameters():
        arrays["grad." + k] = p_.grad.numpy().copy()
    meta = dict(kind=kind, dim=L, premodule=pre is not None, scheme="rk4", num_steps=N, lmbd=0.0,
                norm_correction=(kind != "sgm"), include_t0=True, beta_min=0.1, beta_max=20.0, T=1.0,
                base_channels=8, emb_dim=16)
    _save(name, meta, **arrays)


def build_unet1d_full(module, L, pre, seed):
    """Default-size UNet1D (base 32, embedding 128: BASELINE config 3) from a seed, for fixtures that do not store weights."""
    torch.manual_seed(seed)
    net = module.UNet1D(input_dim=L, premodule=pre)
    with torch.no_grad():
        net.final.weight.mul_(4.0)
    return net


def unet1d_full_case(ref, name, L, B, N, seed, pre="NormalizeLogRadius"):
    """UNet1D at the full configuration-3 size (L = 1000, default widths) with the sparse multiplicative SDE: forward, RK4
    reverse sampling, SSM loss and gradients.  Weights are reproduced from the seed (checksums stored); gradients are stored as
    per-tensor norms and leading entries."""
    net = build_unet1d_full(ref.NNUnet1D, L, pre, seed)
    torch.manual_seed(seed + 1)
    sig = torch.sin(torch.linspace(0, 6.28, L)[None] * torch.randint(1, 4, (64, 1)) + 6.28 * torch.rand(64, 1)) \
        + 0.1 * torch.randn(64, L)
    base, gen, _ = ref_live.build(ref, "msgm_sparse", L, sig, pre, net=net)
    x0 = sig[:B].clone() + 0.3 * torch.randn(B, L)
    s = torch.rand(B)
    with torch.no_grad():
        fwd = net(x0, s)
    torch.manual_seed(seed + 2)
    xs = ref.sde_scheme.rk4_stratonovich_sampler(gen, x0, N, lmbd=0., keep_all_samples=True, include_t0=True,
                                                 norm_correction=True)
    torch.manual_seed(seed + 2)
    noise = torch.stack([torch.randn_like(x0) for _ in range(N)])
    t_ = torch.rand(B, 1).clamp_min(1e-3)
    y = (x0 + 0.2 * torch.randn(B, L)).requires_grad_()
    state = torch.get_rng_state()
    v = ref.SDEs.sample_rademacher(x0.shape, "cpu")
    torch.set_rng_state(state)
    gen.train()
    loss = gen.ssm_loss(t_, x0, y)
    gen.zero_grad()
    loss.mean().backward()
    arrays = dict(x0=x0.numpy(), s=s.numpy(), fwd=fwd.numpy(), noise=noise.numpy(), out=xs.numpy(), t=t_.numpy(),
                  y=y.detach().numpy(), v=v.numpy(), loss=loss.detach().numpy(), **_sde_arrays(base))
    arrays["wsum"] = np.array([float(p_.double().sum()) for p_ in net.state_dict().values()])
    names = [k for k, _ in net.named_parameters()]
    arrays["gradnorm"] = np.array([float(p_.grad.norm()) for _, p_ in net.named_parameters()], dtype=np.float32)
    arrays["gradhead"] = np.stack([torch.nn.functional.pad(p_.grad.flatten()[:8], (0, max(0, 8 - p_.numel()))).numpy()
                                   for _, p_ in net.named_parameters()])
    meta = dict(kind="msgm_sparse", dim=L, premodule=pre is not None, scheme="rk4", num_steps=N, lmbd=0.0,
                norm_correction=True, include_t0=True, beta_min=0.1, beta_max=20.0, T=1.0, seed=seed, param_names=names)
    _save(name, meta, **arrays)


def build_unet2d(module, S, pre, order, seed):
    """Seeded construction shared by the fixture writer (reference module) and the tests (drop-in module): the same
    constructor order consumes the global RNG identically, so the 4.04 M weights need not be stored."""
    torch.manual_seed(seed)
    net = module.VorticityUNet(base_channels=32, channel_mults=(1, 2, 4), num_res_blocks=2, premodule=pre, in_space=S,
                               attention_resolutions=(2, 4), flatten_order=order)
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():  # the reference zero-initialises these; randomise them so that every path is exercised
        for k, p_ in net.named_parameters():
            if p_.abs().sum() == 0 and p_.dim() > 1:
                p_.copy_(torch.randn(p_.shape, generator=g) * (0.5 / p_[0].numel() ** 0.5))
    return net


def unet2d_case(ref, name, S, B, N, seed, pre="NormalizeLogRadius", order="F"):
    """VorticityUNet (NNUnet.py + model/unet.py, the driver's 4.04 M-parameter configuration) on flattened SxS images
    with the sparse multiplicative SDE.  Weights are reproduced from the seed; gradients are stored as per-tensor norms
    and leading entries."""
    L = S * S
    net = build_unet2d(ref.NNUnet, S, pre, order, seed)
    torch.manual_seed(seed + 1)
    img = torch.nn.functional.avg_pool2d(torch.randn(64, 1, S + 4, S + 4), 5, stride=1).reshape(64, L) * 4.0
    base, gen, _ = ref_live.build(ref, "msgm_sparse", L, img, pre, net=net)
    x0 = img[:B].clone() + 0.2 * torch.randn(B, L)
    s = torch.rand(B)
    with torch.no_grad():
        fwd = net(x0, s)
    torch.manual_seed(seed + 2)
    xs = ref.sde_scheme.rk4_stratonovich_sampler(gen, x0, N, lmbd=0., keep_all_samples=True, include_t0=True,
                                                 norm_correction=True)
    torch.manual_seed(seed + 2)
    noise = torch.stack([torch.randn_like(x0) for _ in range(N)])
    t_ = torch.rand(B, 1).clamp_min(1e-3)
    y = (x0 + 0.2 * torch.randn(B, L)).requires_grad_()
    state = torch.get_rng_state()
    v = ref.SDEs.sample_rademacher(x0.shape, "cpu")
    torch.set_rng_state(state)
    gen.train()
    loss = gen.ssm_loss(t_, x0, y)
    gen.zero_grad()
    loss.mean().backward()
    arrays = dict(x0=x0.numpy(), s=s.numpy(), fwd=fwd.numpy(), noise=noise.numpy(), out=xs.numpy(), t=t_.numpy(),
                  y=y.detach().numpy(), v=v.numpy(), loss=loss.detach().numpy(), **_sde_arrays(base))
    arrays["wsum"] = np.array([float(p_.double().sum()) for p_ in net.state_dict().values()])
    names = [k for k, _ in net.named_parameters()]
    arrays["gradnorm"] = np.array([float(p_.grad.norm()) for _, p_ in net.named_parameters()], dtype=np.float32)
    arrays["gradhead"] = np.stack([torch.nn.functional.pad(p_.grad.flatten()[:8], (0, max(0, 8 - p_.numel()))).numpy()
                                   for _, p_ in net.named_parameters()])
    meta = dict(kind="msgm_sparse", dim=L, in_space=S, premodule=pre is not None, scheme="rk4", num_steps=N, lmbd=0.0,
                norm_correction=True, include_t0=True, beta_min=0.1, beta_max=20.0, T=1.0, flatten_order=order,
                seed=seed, param_names=names)
    _save(name, meta, **arrays)


def main():
    ref = ref_live.load()
    if "--elbo-only" in sys.argv:
        elbo_case(ref)
        return
    if "--unet2d-only" in sys.argv:
        unet2d_case(ref, "w01_unet2d_sparse_16x16", 16, 4, 2, 51)
        return
    if "--full-size-only" in sys.argv:  # BASELINE configs 3 and 4 at their full sizes, a few samples
        unet1d_full_case(ref, "v01_unet1d_sparse_L1000", 1000, 3, 2, 61)
        unet2d_case(ref, "w02_unet2d_sparse_32x32", 32, 2, 1, 71)
        return
    if "--unet-only" in sys.argv:
        unet1d_case(ref, "u01_unet1d_sparse_L64", "msgm_sparse", 64, 6, 4, 41)
        unet1d_case(ref, "u02_unet1d_sgm_L48", "sgm", 48, 5, 3, 42, pre=None)
        return
    P = "NormalizeLogRadius"
    sampler_case(ref, "s01_msgm_d2_rk4", "msgm_dense", 2, P, "rk4", 16, 64, 0.0, True, 1)
    sampler_case(ref, "s02_sgm_d2_rk4", "sgm", 2, None, "rk4", 16, 64, 0.0, False, 2)
    sampler_case(ref, "s03_msgm_d16_rk4", "msgm_dense", 16, P, "rk4", 8, 32, 0.0, True, 3)
    sampler_case(ref, "s04_sparse_d32_rk4", "msgm_sparse", 32, P, "rk4", 8, 32, 0.0, True, 4)
    sampler_case(ref, "s05_msgm_d2_heun", "msgm_dense", 2, P, "heun", 16, 64, 0.0, True, 5)
    sampler_case(ref, "s06_msgm_d4_em_l05", "msgm_dense", 4, P, "em", 16, 48, 0.5, False, 6)
    sampler_case(ref, "s07_sparse_d8_em_l05", "msgm_sparse", 8, P, "em", 16, 48, 0.5, True, 7)
    sampler_case(ref, "s08_fwd_msgm_d2_rk4", "msgm_dense", 2, P, "rk4", 16, 64, 0.0, True, 8, forward=True)
    sampler_case(ref, "s09_fwd_keep_d2_rk4", "msgm_dense", 2, P, "rk4", 16, 64, 0.0, False, 9, forward=True,
                 keep=True)
    sampler_case(ref, "s10_msgm_d32_rk4", "msgm_dense", 32, P, "rk4", 4, 16, 0.0, True, 10)
    sampler_case(ref, "s11_sgm_d16_heun_l03", "sgm", 16, None, "heun", 8, 32, 0.3, False, 11)
    sampler_case(ref, "s12_msgm_d2_rk4_Tov", "msgm_dense", 2, P, "rk4", 3, 40, 0.0, False, 12, T_=0.37)
    sampler_case(ref, "s13_fwd_sgm_d4_rk4", "sgm", 4, None, "rk4", 8, 32, 0.0, False, 13, forward=True)
    ssm_case(ref, "t01_ssm_msgm_d2", "msgm_dense", 2, P, 32, 21)
    ssm_case(ref, "t02_ssm_sgm_d2", "sgm", 2, None, 32, 22)
    ssm_case(ref, "t03_ssm_sparse_d8", "msgm_sparse", 8, P, 24, 23)
    ssm_case(ref, "t04_ssm_msgm_d16", "msgm_dense", 16, P, 16, 24)
    misc_case(ref)
    elbo_case(ref)
    unet1d_case(ref, "u01_unet1d_sparse_L64", "msgm_sparse", 64, 6, 4, 41)
    unet1d_case(ref, "u02_unet1d_sgm_L48", "sgm", 48, 5, 3, 42, pre=None)
    unet2d_case(ref, "w01_unet2d_sparse_16x16", 16, 4, 2, 51)
    unet1d_full_case(ref, "v01_unet1d_sparse_L1000", 1000, 3, 2, 61)
    unet2d_case(ref, "w02_unet2d_sparse_32x32", 32, 2, 1, 71)


if __name__ == "__main__":
    main()
