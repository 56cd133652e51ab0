"""Sliced score matching through the fused CUDA kernels (msgm_ssm_mlp_forward / _backward).

``ssm(gen, x)`` follows ``PluginReverseSDE.ssm`` of the reference (SDEs.py:607-646): draw t, noise x forward to y_t,
draw the Hutchinson probe v (same RNG call order as the reference), and return the per-sample loss (B,).  The loss
is a ``torch.autograd.Function`` node: ``loss.mean().backward()`` fills ``.grad`` of the score net's parameters from
the hand-derived backward kernels, so ``torch.optim.Adam(gen.parameters())`` trains exactly as in the reference driver
(MSGM_higherDim.py:803-809).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib


class _SsmMlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gen, t, y, v, *params):
        dev = y.device
        handle = _lib.ctx(dev)
        base, net = gen.base_sde, gen.a
        B = y.shape[0]
        sd, keep = base.desc(dev)
        sd.dim = y.shape[1]  # SGMsde has no dim of its own
        md, k2 = net.desc(dev)
        yc, vc, tc = _lib.f32c(y, dev), _lib.f32c(v, dev), _lib.f32c(t.reshape(-1), dev)
        loss = torch.empty(B, device=dev, dtype=torch.float32)
        scratch = torch.empty(int(_lib.lib().msgm_ssm_scratch_bytes(B)) // 4, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().msgm_ssm_mlp_forward(handle, C.byref(sd), C.byref(md), _lib.ptr(yc), _lib.ptr(vc),
                                                   _lib.ptr(tc), _lib.ptr(loss), _lib.ptr(scratch), B,
                                                   _lib.stream_ptr(dev)))
        ctx.gen, ctx.saved = gen, (yc, vc, tc, scratch, keep + k2)
        ctx.shapes = [p.shape for p in params]
        return loss

    @staticmethod
    def backward(ctx, gout):
        yc, vc, tc, scratch, _ = ctx.saved
        dev = yc.device
        gen = ctx.gen
        base, net = gen.base_sde, gen.a
        sd, keep = base.desc(dev)
        sd.dim = yc.shape[1]
        md, k2 = net.desc(dev)
        n = sum(int(torch.Size(s).numel()) for s in ctx.shapes)
        flat = torch.empty(n, device=dev, dtype=torch.float32)
        g = _lib.f32c(gout, dev)
        _lib.check(_lib.lib().msgm_ssm_mlp_backward(_lib.ctx(dev), C.byref(sd), C.byref(md), _lib.ptr(yc),
                                                    _lib.ptr(vc), _lib.ptr(tc), _lib.ptr(g), _lib.ptr(scratch),
                                                    _lib.ptr(flat), yc.shape[0], _lib.stream_ptr(dev)))
        grads, o = [], 0
        for s in ctx.shapes:
            k = int(torch.Size(s).numel())
            grads.append(flat[o:o + k].view(s))
            o += k
        return (None, None, None, None, *grads)


def tc_ok(gen, d) -> bool:
    """``gen.ssm_precision == "f16tc"`` and the tensor-core SSM kernel covers this problem (csrc/ssm_tc.cu: d <= 16)."""
    return getattr(gen, "ssm_precision", "fp32") == "f16tc" and d <= 16


def fused_loss_and_grads(gen, t, y, v, gout, grad_flat):
    """Per-sample loss (B,) AND the parameter gradient of ``sum_b gout_b loss_b`` written straight into ``grad_flat``
    (torch parameter order of the MLP), without an autograd graph: what train.GraphedSsmStep records for MLP score nets.
    ``gen.ssm_precision = "fp32"`` (default, reference arithmetic): the forward and backward kernels back to back;
    ``"f16tc"``: ONE tcgen05 launch for loss, backward and weight gradients (+ the partial-sum launch), stated tolerance in
    include/msgm_b200.h."""
    dev = y.device
    base, net = gen.base_sde, gen.a
    B = y.shape[0]
    sd, keep = base.desc(dev)
    sd.dim = y.shape[1]
    md, k2 = net.desc(dev)
    yc, vc, tc = _lib.f32c(y, dev), _lib.f32c(v, dev), _lib.f32c(t.reshape(-1), dev)
    loss = torch.empty(B, device=dev, dtype=torch.float32)
    if tc_ok(gen, y.shape[1]):
        h, L = _lib.ctx(dev), _lib.lib()
        nbytes = int(L.msgm_ssm_tc_scratch_bytes(h, int(sd.dim), int(md.premodule), B))
        part = torch.empty(nbytes // 4, device=dev, dtype=torch.float32)
        _lib.check(L.msgm_ssm_mlp_fwd_bwd_tc(h, C.byref(sd), C.byref(md), _lib.ptr(yc), _lib.ptr(vc), _lib.ptr(tc),
                                             _lib.ptr(_lib.f32c(gout, dev)), _lib.ptr(loss), _lib.ptr(grad_flat),
                                             _lib.ptr(part), float(2 ** round(math.log2(max(B, 1)))), B,
                                             _lib.stream_ptr(dev)))
        return loss
    scratch = torch.empty(int(_lib.lib().msgm_ssm_scratch_bytes(B)) // 4, device=dev, dtype=torch.float32)
    h = _lib.ctx(dev)
    _lib.check(_lib.lib().msgm_ssm_mlp_forward(h, C.byref(sd), C.byref(md), _lib.ptr(yc), _lib.ptr(vc), _lib.ptr(tc),
                                               _lib.ptr(loss), _lib.ptr(scratch), B, _lib.stream_ptr(dev)))
    _lib.check(_lib.lib().msgm_ssm_mlp_backward(h, C.byref(sd), C.byref(md), _lib.ptr(yc), _lib.ptr(vc), _lib.ptr(tc),
                                                _lib.ptr(gout), _lib.ptr(scratch), _lib.ptr(grad_flat), B,
                                                _lib.stream_ptr(dev)))
    return loss


class _SsmMlpTc(torch.autograd.Function):
    """Eager-mode wrapper of the fused tensor-core step: the kernel needs the upstream gradient up front, so ``forward``
    runs it for the loss and ``backward`` runs it again with the real ``gout`` (the graphed trainer calls
    ``fused_loss_and_grads`` directly, once per iteration)."""

    @staticmethod
    def forward(ctx, gen, t, y, v, *params):
        n = sum(p.numel() for p in params)
        scratch_grad = torch.empty(n, device=y.device, dtype=torch.float32)
        ctx.gen, ctx.saved, ctx.shapes = gen, (t, y, v), [p.shape for p in params]
        return fused_loss_and_grads(gen, t, y, v, torch.zeros(y.shape[0], device=y.device), scratch_grad)

    @staticmethod
    def backward(ctx, gout):
        t, y, v = ctx.saved
        n = sum(int(torch.Size(s).numel()) for s in ctx.shapes)
        flat = torch.empty(n, device=y.device, dtype=torch.float32)
        fused_loss_and_grads(ctx.gen, t, y, v, gout, flat)
        grads, o = [], 0
        for s in ctx.shapes:
            k = int(torch.Size(s).numel())
            grads.append(flat[o:o + k].view(s))
            o += k
        return (None, None, None, None, *grads)


def _mu_and_a(gen, t_, y):
    """mu_to_div = g.a (+ beta y / 2 for SGM) in its cancelled form, and a, with the net evaluated once."""
    base = gen.base_sde
    a = gen.a(y, t_.squeeze())
    g = base.g(t_, y, base.sparseTensor)
    if base.sparseTensor:
        I, _, K = base.IJK()
        mu = torch.zeros_like(y).scatter_add(1, I.unsqueeze(0).expand(y.shape[0], -1), g * a[:, K])
    elif g.dim() > 2:
        mu = torch.einsum('bij, bj -> bi', g, a)
    else:
        mu = g * a + 0.5 * base.beta(t_) * y
    return mu, a


def _ssm_loss_autograd(gen, t_, y, v):
    """SSM loss for score nets without a fused kernel (U-Nets): the reference's quantity v^T d(mu_to_div)/dy v + |a|^2 / 2
    (SDEs.py:616-646) on GPU tensors.  Library autograd, not a hand-written kernel.

    Default (``gen.ssm_forward_mode``, on): the directional derivative J v comes from ONE forward-mode pass
    (torch.func.jvp) and the parameter gradient from one reverse pass over it, instead of the reference's recipe (VJP with
    create_graph, then a double backward through cuDNN's fp32 double-backward convolutions): same value to fp32 rounding
    (loss 2e-7, gradients 7e-7 from the reference), 3-5x faster per iteration.  The reference's recipe remains as the
    fallback (an op without a forward-mode rule) and as ``ssm_forward_mode = False``.  (torch.autograd.forward_ad's core
    dual tensors were tried instead of torch.func: their softmax rule writes in place and breaks the reverse pass.)"""
    if not y.is_cuda:
        raise RuntimeError("sdeflow_light_b200 runs on CUDA only (no CPU fallback)")
    with torch.enable_grad():
        if getattr(gen, "ssm_forward_mode", True):
            try:
                (mu, a), (jv, _) = torch.func.jvp(lambda y_: _mu_and_a(gen, t_, y_), (y.detach(),), (v,))
                return (jv * v).reshape(y.size(0), -1).sum(1) + (a ** 2).reshape(y.size(0), -1).sum(1) / 2
            except (RuntimeError, NotImplementedError) as exc:
                if "forward" not in str(exc).lower():
                    raise
                gen.ssm_forward_mode = False  # no forward-mode rule somewhere in this net: use the reference's recipe
        y = y.detach().requires_grad_()
        mu, a = _mu_and_a(gen, t_, y)
        jv = torch.autograd.grad(mu, y, v, create_graph=gen.training)[0]
        return (jv * v).reshape(y.size(0), -1).sum(1) + (a ** 2).reshape(y.size(0), -1).sum(1) / 2


def ssm_loss(gen, t_, x, y, v=None):
    """Per-sample SSM loss for given (t, y); ``v`` defaults to a fresh probe like the reference (SDEs.py:637-638)."""
    from . import NN, SDEs
    net = gen.a
    if v is None:
        with torch.no_grad():
            v = SDEs.sample_v(x.shape, vtype=gen.vtype, device=gen.deviceReverseSDE)
        if v is None:
            raise ValueError(f"vtype {gen.vtype} not supported")
    v = v.to(y)
    if not (isinstance(net, NN.MLP) and net.fused_ok()):
        from . import unet_train
        if unet_train.supported(gen, y):  # U-Nets: hand-written forward-mode kernels, torch.autograd only as the tape
            return unet_train.ssm_loss(gen, t_.to(y), y, v)
        return _ssm_loss_autograd(gen, t_.to(y), y, v)
    params = [p for l in net.linears() for p in (l.weight, l.bias)]
    if tc_ok(gen, y.shape[1]):
        return _SsmMlpTc.apply(gen, t_, y.detach(), v, *params)
    return _SsmMlp.apply(gen, t_, y.detach(), v, *params)


def ssm(gen, x):
    if getattr(gen, "device_rng", False) and x.is_cuda and x.dim() == 2:
        from . import SDEs
        if gen.vtype in SDEs._VTYPES and isinstance(gen.base_sde, (SDEs.MSGMsde, SDEs.SGMsde)) \
                and SDEs._prepare_dim_ok(gen.base_sde, x.shape[1]):
            t_, y, v = gen._prepare(x)  # t, y_t and the probe from one launch
            return ssm_loss(gen, t_, x, y, v)
    t_, x, y = gen.sample_txy(x)
    return ssm_loss(gen, t_, x, y)
