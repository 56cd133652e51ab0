"""Pin the oracle: run the live reference and the oracle restatement on identical seeded inputs.

Run in the build container only:  ``python -m oracle.check_against_reference``  (exit 0 == pinned).
TEST INFRASTRUCTURE.  Tolerances are fp32 rounding only (the two sides issue the same ATen ops).
"""
from __future__ import annotations

import sys

import torch

from . import msgm_oracle as O
from . import ref_live

TOL = 2e-6


def _maxdiff(a, b):
    return float((a - b).abs().max())


def _boost(net, k=8.0):
    """Make an untrained MLP produce an O(1) drift so that the score term is exercised."""
    with torch.no_grad():
        net.main[6].weight.mul_(k)
        net.main[6].bias.mul_(k)


def check_samplers(ref, kind, dim, premodule, B=48, N=12):
    torch.manual_seed(100 + dim)
    x_init = torch.randn(512, dim) * 1.5
    base, gen, net = ref_live.build(ref, kind, dim, x_init, premodule)
    _boost(net)
    sde, mlp = ref_live.to_oracle(base, net)
    rev = O.OReverse(sde, mlp)
    x0 = torch.randn(B, dim)
    worst = 0.0
    fns = {"em": ref.sde_scheme.euler_maruyama_sampler, "heun": ref.sde_scheme.heun_sampler,
           "rk4": ref.sde_scheme.rk4_stratonovich_sampler}
    for scheme, fn in fns.items():
        for lmbd in (0.0, 0.5):
            for nc in (False, True):
                torch.manual_seed(7)
                r = fn(gen, x0, N, lmbd=lmbd, keep_all_samples=True, include_t0=True, norm_correction=nc)
                torch.manual_seed(7)
                o = O.integrate(rev, x0, N, scheme, lmbd, True, None, True, None, nc)
                worst = max(worst, _maxdiff(r, o))
    # capture modes + T_ override + forward adapter
    keep = torch.randint(0, N + 1, (B, 1)).to(torch.int)
    torch.manual_seed(8)
    r = ref.sde_scheme.rk4_stratonovich_sampler(ref.SDEs.forward_SDE(base, base.T), x0, N, lmbd=0,
                                                keep_all_samples=False, samplesToKeep=keep, include_t0=True)
    torch.manual_seed(8)
    o = O.integrate(O.OForward(sde), x0, N, "rk4", 0.0, False, keep, True)
    worst = max(worst, _maxdiff(r, o))
    torch.manual_seed(9)
    r = ref.sde_scheme.rk4_stratonovich_sampler(gen, x0, 3, keep_all_samples=False, T_=torch.tensor([0.37]))
    torch.manual_seed(9)
    o = O.integrate(rev, x0, 3, "rk4", keep_all_samples=False, T_=torch.tensor([0.37]).item())
    worst = max(worst, _maxdiff(r, o))
    return worst


def check_ssm(ref, kind, dim, premodule, B=40):
    torch.manual_seed(200 + dim)
    x_init = torch.randn(512, dim) * 1.5
    base, gen, net = ref_live.build(ref, kind, dim, x_init, premodule)
    _boost(net, 3.0)
    sde, mlp = ref_live.to_oracle(base, net)
    for p in mlp.parameters():
        p.requires_grad_(True)
    rev = O.OReverse(sde, mlp)
    x = torch.randn(B, dim)
    torch.manual_seed(11)
    lr = gen.ssm(x)
    gen.zero_grad()
    lr.mean().backward()
    gr = [p.grad.clone() for p in net.parameters()]
    torch.manual_seed(11)
    lo, _ = O.ssm(rev, x)
    go = torch.autograd.grad(lo.mean(), mlp.parameters())
    worst = _maxdiff(lr.detach(), lo.detach()) / max(1.0, float(lr.abs().max()))
    for a, b in zip(gr, go):
        worst = max(worst, _maxdiff(a, b) / max(1.0, float(a.abs().max())))
    return worst


def check_misc(ref):
    worst = 0.0
    torch.manual_seed(3)
    x_init = torch.randn(2000, 2)
    base, gen, net = ref_live.build(ref, "msgm_dense", 2, x_init, "NormalizeLogRadius")
    sde, _ = ref_live.to_oracle(base, net)
    torch.manual_seed(4)
    r = gen.latent_sample(300, 2)
    torch.manual_seed(4)
    o = O.latent_sample(sde, 300)
    worst = max(worst, _maxdiff(r, o))
    a, b = torch.randn(200, 3), torch.randn(150, 3) + 0.5
    worst = max(worst, abs(float(ref.qc.compute_mmd(a, b)) - float(O.compute_mmd(a, b))))
    torch.manual_seed(5)
    Gr = ref.SDEs.MSGMsde(x_init, denseTensor=True, estim_cst_norm_dens_r_T=False, norm_map="log",
                          T=ref_live.T_param()).G
    torch.manual_seed(5)
    worst = max(worst, _maxdiff(Gr, O.draw_dense_G(2)))
    return worst


def main() -> int:
    ref = ref_live.load()
    rows = []
    for kind, dim, pre in [("sgm", 2, None), ("msgm_dense", 2, "NormalizeLogRadius"),
                           ("msgm_dense", 8, "NormalizeLogRadius"), ("msgm_sparse", 12, "NormalizeLogRadius"),
                           ("msgm_dense", 3, None)]:
        rows.append((f"samplers {kind} d={dim}", check_samplers(ref, kind, dim, pre)))
    for kind, dim, pre in [("sgm", 2, None), ("msgm_dense", 2, "NormalizeLogRadius"),
                           ("msgm_sparse", 6, "NormalizeLogRadius"), ("msgm_dense", 5, "NormalizeLogRadius")]:
        rows.append((f"ssm+grads {kind} d={dim}", check_ssm(ref, kind, dim, pre)))
    rows.append(("latent_sample / mmd / draw_G", check_misc(ref)))
    bad = 0
    for name, w in rows:
        ok = w <= TOL
        bad += not ok
        print(f"{'ok ' if ok else 'BAD'}  {name:40s} max|diff| = {w:.3e}")
    print("oracle pinned against live reference" if not bad else "ORACLE MISMATCH")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
