"""GPU parity of the fused SSM train step (forward-mode loss + hand-derived backward) against the reference's golden
fixtures (loss and every parameter gradient from autograd double-backward) and against the oracle on other shapes.

Tolerance: |loss_cuda - loss_ref| <= 2e-5 max|loss|, |grad_cuda - grad_ref| <= 1e-4 max|grad| per parameter tensor (fp32,
different summation order; the reference also carries the analytically vanishing v^T g(s,v) a term at rounding level).
"""
import pytest
import torch

import sdeflow_light_b200 as P
from oracle import msgm_oracle as O
from tests import _build as Bd
from tests import _golden as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _check(name, gen, t, x, y, v, loss_ref, grads_ref, ltol=2e-5, gtol=1e-4):
    net = gen.a
    gen.zero_grad()
    loss = gen.ssm_loss(t.to(DEV), x.to(DEV), y.to(DEV), v.to(DEV))
    assert loss.shape == loss_ref.shape
    lerr = float((loss.detach().cpu() - loss_ref).abs().max()) / max(1.0, float(loss_ref.abs().max()))
    loss.mean().backward()
    worst = 0.0
    for i, l in enumerate(net.linears()):
        for p, ref in ((l.weight, grads_ref[2 * i]), (l.bias, grads_ref[2 * i + 1])):
            assert p.grad is not None and p.grad.shape == ref.shape
            worst = max(worst, float((p.grad.cpu() - ref).abs().max()) / max(1e-12, float(ref.abs().max())))
    Bd.report(test=name, loss_rel_err=lerr, grad_rel_err=worst)
    assert lerr <= ltol, f"{name}: loss rel err {lerr:.3e}"
    assert worst <= gtol, f"{name}: grad rel err {worst:.3e}"


@pytest.mark.parametrize("name", G.names("t"))
def test_golden_ssm(name):
    meta, arr = G.load(name)
    _, _, gen = Bd.gen_from(meta, arr, DEV)
    grads = [arr[k] for i in range(4) for k in (f"gW{i}", f"gb{i}")]
    _check(name, gen, arr["t"], arr["x"], arr["y"], arr["v"], arr["loss"], grads)


@pytest.mark.parametrize("kind,d,pre,B", [("msgm_dense", 3, True, 257), ("msgm_dense", 32, True, 33),
                                          ("msgm_sparse", 5, False, 100), ("sgm", 7, False, 64),
                                          ("msgm_dense", 2, True, 3000)])
def test_ssm_against_oracle(kind, d, pre, B):
    torch.manual_seed(300 + d)
    sde = O.make_sgm(d) if kind == "sgm" else O.make_msgm(torch.randn(256, d) * 1.5, dense=(kind == "msgm_dense"))
    mlp = O.init_mlp(d, pre, seed=d, scale=3.0)
    for p in mlp.parameters():
        p.requires_grad_(True)
    t = torch.rand(B, 1).clamp_min(1e-3)
    y = (torch.randn(B, d) * 1.4).requires_grad_()
    v = O.sample_rademacher((B, d))
    loss = O.ssm_loss(O.OReverse(sde, mlp), t, y, v)
    grads = torch.autograd.grad(loss.mean(), mlp.parameters())
    for p in mlp.parameters():
        p.requires_grad_(False)
    _, _, gen = Bd.from_oracle(sde, mlp, DEV)
    _check(f"ssm-oracle-{kind}-d{d}-B{B}", gen, t, y.detach(), y.detach(), v, loss.detach(), [g.detach() for g in grads])


def test_forward_noising_replays_reference_rng():
    """base_sde.sample(t, x) with the reference's draws injected reproduces the reference's y_t (SDEs.py:78-122)."""
    meta, arr = G.load("t01_ssm_msgm_d2")
    base, _ = Bd.base_from(meta, arr, DEV)
    y = base.sample_scheme(arr["t"].to(DEV), arr["x"].to(DEV), keep_all_samples=False, noise=arr["fwd_noise"],
                           noise_rows=arr["singles"])
    err = float((y.cpu() - arr["y"]).abs().max())
    Bd.report(test="forward-noising-replay", max_abs=err)
    assert err <= 5e-5


def test_ssm_end_to_end_trains():
    """gen.ssm(x).mean().backward(); Adam.step() -- the reference driver's loop (MSGM_higherDim.py:803-809) -- runs on
    the fused kernels and reduces the loss on a fixed batch."""
    torch.manual_seed(0)
    d = 2
    data = O.swiss_roll(4096)
    T = Bd.T_param(1.0)
    for make in ("msgm", "sgm"):
        if make == "msgm":
            base = P.MSGMsde(data, beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=True, norm_map="log",
                             num_steps_forward=16, device=DEV, estim_cst_norm_dens_r_T=False)
            net = P.MLP(d, premodule="NormalizeLogRadius").to(DEV)
        else:
            base = P.SGMsde(beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, num_steps_forward=16, device=DEV)
            net = P.MLP(d).to(DEV)
        gen = P.PluginReverseSDE(base, net, T, deviceReverseSDE=DEV).to(DEV)
        opt = torch.optim.Adam(gen.parameters(), lr=2e-3)
        # fixed evaluation set (same t, y, v before and after): the per-batch loss itself is a noisy Hutchinson estimate
        xe = data[:4096].to(DEV)
        te, _, ye = gen.sample_txy(xe)
        ve = P.SDEs.sample_v(xe.shape, DEV)
        with torch.no_grad():
            before = float(gen.ssm_loss(te, xe, ye, ve).mean())
        for it in range(200):
            opt.zero_grad()
            x = data[torch.randint(0, 4096, (256,))].to(DEV)
            loss = gen.ssm(x).mean()
            loss.backward()
            opt.step()
            assert bool(torch.isfinite(loss))
        with torch.no_grad():
            after = float(gen.ssm_loss(te, xe, ye, ve).mean())
        Bd.report(test=f"train-{make}", eval_loss_before=before, eval_loss_after=after)
        assert after < before - 0.005, (make, before, after)  # observed: msgm 0.003 -> -0.02, sgm 10.2 -> 0.7


def test_graphed_train_step_matches_eager_and_trains():
    """train.GraphedSsmStep replays the reference loop (MSGM_higherDim.py:803-809) as one CUDA graph: (1) on the same
    Philox stream a replay produces the same loss and parameter gradient as the eager autograd iteration;
    (2) successive replays draw fresh t / noise / v; (3) training through replays reduces the held-out loss."""
    from sdeflow_light_b200.train import GraphedSsmStep
    d = 2
    data = O.swiss_roll(4096).to(DEV)
    T = Bd.T_param(1.0)
    for make in ("msgm", "sgm"):
        torch.manual_seed(1)

        def build():
            torch.manual_seed(1)
            if make == "msgm":
                base = P.MSGMsde(data.cpu(), beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=True,
                                 norm_map="log", num_steps_forward=16, device=DEV, estim_cst_norm_dens_r_T=False)
                net = P.MLP(d, premodule="NormalizeLogRadius").to(DEV)
            else:
                base = P.SGMsde(beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, num_steps_forward=16, device=DEV)
                net = P.MLP(d).to(DEV)
            return P.PluginReverseSDE(base, net, T, deviceReverseSDE=DEV).to(DEV)

        gen_g, gen_e = build(), build()
        if make == "msgm":  # G is random: share it so both replicas integrate the same SDE
            gen_e.base_sde.G, gen_e.base_sde.L_G = gen_g.base_sde.G, gen_g.base_sde.L_G
        gen_e.load_state_dict(gen_g.state_dict())
        step = GraphedSsmStep(gen_g, (256, d), lr=0.0, seed=123)
        step.set_lr(2e-3)
        opt_e = torch.optim.Adam(gen_e.parameters(), lr=2e-3)
        gen_e.train()
        gen_e.device_rng = True
        gen_e._rng = (123, torch.zeros(1, device=DEV, dtype=torch.int64), 0)  # the trainer's stream at iteration 0
        x = data[:256]
        l_g = float(step(x))
        g_graph = torch.cat([p.grad.reshape(-1) for p in gen_g.a.parameters()]).clone()
        opt_e.zero_grad()
        l = gen_e.ssm(x).mean()
        l.backward()
        l_e = float(l.detach())
        g_eager = torch.cat([p.grad.reshape(-1) for p in gen_e.a.parameters()])
        gerr = float((g_graph - g_eager).abs().max() / g_eager.abs().max())
        Bd.report(test=f"graphed-vs-eager-grad-{make}", max_rel=gerr)
        assert gerr <= 1e-4, (make, gerr)
        Bd.report(test=f"graphed-vs-eager-{make}", loss_graph=l_g, loss_eager=l_e)
        assert abs(l_g - l_e) <= 1e-5 + 1e-4 * abs(l_e), (make, l_g, l_e)  # observed: msgm identical, sgm 1.6e-5 rel
        # (2) fresh randomness per replay
        l2 = float(step(x))
        assert l2 != l_g
        # (3) trains
        xe = data
        te, _, ye = gen_g.sample_txy(xe)
        ve = P.SDEs.sample_v(xe.shape, DEV)
        with torch.no_grad():
            before = float(gen_g.ssm_loss(te, xe, ye, ve).mean())
        for it in range(200):
            loss = step(data[torch.randint(0, 4096, (256,), device=DEV)])
        assert bool(torch.isfinite(loss))
        with torch.no_grad():
            after = float(gen_g.ssm_loss(te, xe, ye, ve).mean())
        Bd.report(test=f"graphed-train-{make}", eval_loss_before=before, eval_loss_after=after,
                  own_launches_per_iter=step.launches_per_iter)
        assert after < before - 0.005, (make, before, after)


@pytest.mark.parametrize("kind,d", [("msgm_dense", 2), ("msgm_dense", 8), ("msgm_dense", 19), ("msgm_sparse", 11),
                                    ("sgm", 5)])
def test_ssm_prepare_kernel(kind, d):
    """msgm_ssm_prepare (t, v, y_t in one launch; SDEs.py:648-693, 514-536): t ~ U(0,T) floored at t_epsilon, v in
    {-1,+1}, and y_t identical to msgm_noise_forward run with the drawn t on the same Philox stream."""
    import ctypes as C
    from sdeflow_light_b200 import _lib
    torch.manual_seed(3)
    B = 20000
    T = Bd.T_param(1.0)
    if kind == "sgm":
        base = P.SGMsde(beta_min=0.1, beta_max=20., T=T, t_epsilon=0.05, num_steps_forward=16, device=DEV)
        net = P.MLP(d).to(DEV)
    else:
        base = P.MSGMsde(torch.randn(64, d), beta_min=0.1, beta_max=20., T=T, t_epsilon=0.05,
                         denseTensor=(kind == "msgm_dense"), norm_map="log", num_steps_forward=16, device=DEV,
                         estim_cst_norm_dens_r_T=False)
        net = P.MLP(d, premodule="NormalizeLogRadius").to(DEV)
    gen = P.PluginReverseSDE(base, net, T, deviceReverseSDE=DEV).to(DEV)
    gen.device_rng = True
    gen._rng = (99, None, 0)
    x = torch.randn(B, d, device=DEV)
    t, y, v = gen._prepare(x)
    t = t.reshape(-1)
    assert float(t.min()) >= 0.05 - 1e-7 and float(t.max()) <= 1.0
    frac_floor = float((t == 0.05).float().mean())
    assert abs(frac_floor - 0.05) < 0.01 and abs(float(t.mean()) - (0.5 + 0.05 ** 2 / 2)) < 0.01
    assert bool(((v == 1) | (v == -1)).all()) and abs(float(v.mean())) < 0.02
    assert abs(float((v[:, 0] * v[:, -1]).mean())) < 0.03 or d == 1
    assert bool(torch.isfinite(y).all())
    if kind != "sgm":
        y2 = x.clone()
        sd, keep = base.desc(torch.device(DEV))
        _lib.check(_lib.lib().msgm_noise_forward(_lib.ctx(torch.device(DEV)), C.byref(sd), _lib.ptr(t.contiguous()),
                                                 _lib.ptr(y2), 16, _lib.ptr(base._fwd_grid[1]), None, None, 99, 0, B,
                                                 _lib.stream_ptr(torch.device(DEV))))
        assert float((y - y2).abs().max()) == 0.0
    else:  # closed form: E[y | x, t] = mean_weight(t) x, so the residual is N(0, var(t))
        mw, var = base.mean_weight(t.reshape(-1, 1)), base.var(t.reshape(-1, 1))
        z = (y - mw * x) / var.sqrt()
        assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
    # sharding invariance: rows [B/2, B) drawn with a row offset equal the second half of the full draw
    gen._rng = (99, None, B // 2)
    t2, y2h, v2 = gen._prepare(x[B // 2:])
    assert float((t2.reshape(-1) - t[B // 2:]).abs().max()) == 0.0 and float((v2 - v[B // 2:]).abs().max()) == 0.0
    assert float((y2h - y[B // 2:]).abs().max()) == 0.0
    Bd.report(test=f"ssm-prepare-{kind}-d{d}", t_floor_frac=frac_floor, v_mean=float(v.mean()))


# ---- tensor-core SSM step (csrc/ssm_tc.cu): fp16 operands / fp32 accumulation, one launch for loss + backward + wgrad ----------
# Stated tolerance of this mode: loss 2e-3 of max|loss|, gradients 3e-3 of max|grad| per parameter tensor.
STC_LTOL, STC_GTOL = 2e-3, 3e-3  # observed: loss <= 9e-4, gradients <= 8e-4


@pytest.mark.parametrize("name", [n for n in G.names("t") if "d32" not in n])
def test_golden_ssm_tensor_core(name):
    """The reference's own loss / gradient fixtures (autograd double backward) against the tcgen05 step."""
    meta, arr = G.load(name)
    if meta["dim"] > 16:
        pytest.skip("f16tc SSM step covers d <= 16")
    _, _, gen = Bd.gen_from(meta, arr, DEV)
    gen.ssm_precision = "f16tc"
    grads = [arr[k] for i in range(4) for k in (f"gW{i}", f"gb{i}")]
    _check(name + "-f16tc", gen, arr["t"], arr["x"], arr["y"], arr["v"], arr["loss"], grads, STC_LTOL, STC_GTOL)
    assert P._lib.debug_flags(DEV) == 0


@pytest.mark.parametrize("kind,d,pre,B", [("msgm_dense", 8, True, 64), ("msgm_dense", 8, True, 1000),
                                          ("msgm_dense", 3, True, 257), ("msgm_dense", 16, True, 333),
                                          ("msgm_sparse", 5, False, 100), ("msgm_sparse", 12, True, 129),
                                          ("sgm", 7, False, 64), ("sgm", 16, False, 4100), ("msgm_dense", 2, True, 20000)])
def test_ssm_tensor_core_against_oracle(kind, d, pre, B):
    """Ragged batches (partial tiles, several tiles per CTA), every SDE kind, both layer-1 widths."""
    torch.manual_seed(700 + d)
    sde = O.make_sgm(d) if kind == "sgm" else O.make_msgm(torch.randn(256, d) * 1.5, dense=(kind == "msgm_dense"))
    mlp = O.init_mlp(d, pre, seed=d, scale=3.0)
    for p in mlp.parameters():
        p.requires_grad_(True)
    t = torch.rand(B, 1).clamp_min(1e-3)
    y = (torch.randn(B, d) * 1.4).requires_grad_()
    v = O.sample_rademacher((B, d))
    loss = O.ssm_loss(O.OReverse(sde, mlp), t, y, v)
    grads = torch.autograd.grad(loss.mean(), mlp.parameters())
    for p in mlp.parameters():
        p.requires_grad_(False)
    _, _, gen = Bd.from_oracle(sde, mlp, DEV)
    gen.ssm_precision = "f16tc"
    _check(f"ssm-tc-oracle-{kind}-d{d}-B{B}", gen, t, y.detach(), y.detach(), v, loss.detach(), [g.detach() for g in grads],
           STC_LTOL, STC_GTOL)
    assert P._lib.debug_flags(DEV) == 0


def test_ssm_tensor_core_graphed_training_tracks_fp32():
    """train.GraphedSsmStep with ssm_precision = "f16tc": the whole iteration (prologue, fused tensor-core step, partial
    sum, Adam) replays as one graph and, on the same seeds and Philox streams, follows the fp32 mode's training trajectory
    (per-iteration losses and the loss on a fixed held-out (t, y, v) after 200 updates)."""
    from sdeflow_light_b200.train import GraphedSsmStep
    d = 2
    data = O.swiss_roll(8192).to(DEV)
    T = Bd.T_param(1.0)
    hist, evals = {}, {}
    for prec in ("fp32", "f16tc"):
        torch.manual_seed(4)
        base = P.MSGMsde(data.cpu(), beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=True, norm_map="log",
                         num_steps_forward=16, device=DEV, estim_cst_norm_dens_r_T=False)
        gen = P.PluginReverseSDE(base, P.MLP(d, premodule="NormalizeLogRadius").to(DEV), T, deviceReverseSDE=DEV).to(DEV)
        if prec == "f16tc":  # same random G and initial weights as the fp32 run
            gen.base_sde.G.copy_(G0)
            gen.base_sde.L_G.copy_(LG0)
            gen.load_state_dict(sd0)
        else:
            G0, LG0, sd0 = gen.base_sde.G.clone(), gen.base_sde.L_G.clone(), {k: v.clone() for k, v in gen.state_dict().items()}
        gen.ssm_precision = prec
        gen.device_rng = True
        gen._rng = (5, None, 0)
        with torch.no_grad():
            te, ye, ve = gen._prepare(data[:2048])
            before = float(gen.ssm_loss(te, data[:2048], ye, ve).mean())
        step = GraphedSsmStep(gen, (512, d), lr=2e-3, seed=77)
        torch.manual_seed(9)
        hist[prec] = [float(step(data[torch.randint(0, data.shape[0], (512,), device=DEV)])) for _ in range(200)]
        with torch.no_grad():
            evals[prec] = (before, float(gen.ssm_loss(te, data[:2048], ye, ve).mean()))
    dmax = max(abs(a - b) for a, b in zip(hist["fp32"], hist["f16tc"]))
    Bd.report(test="ssm-tc-graphed-training", max_loss_diff=dmax, eval_fp32=evals["fp32"], eval_f16tc=evals["f16tc"])
    assert P._lib.debug_flags(DEV) == 0
    assert dmax < 5e-3, dmax
    assert abs(evals["f16tc"][1] - evals["fp32"][1]) < 5e-3 and evals["f16tc"][1] != evals["f16tc"][0]
