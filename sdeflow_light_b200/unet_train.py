"""Hand-written score-matching training path of the U-Net score nets (no cuDNN / cuBLAS / ATen math on the path).

``PluginReverseSDE.ssm_loss`` (SDEs.py:616-646) needs the net output ``a(y, s)`` and its directional derivative
``adot = (da/dy) v``; the reference gets them from a VJP with ``create_graph`` and then a double backward through cuDNN.
Here the net is evaluated in FORWARD mode on *pairs*: every activation is a tensor of 2B samples, [0,B) primal and [B,2B)
tangent.  Linear layers (convolutions, Linear) act on both halves alike, so they run on this repo's inference kernels
(tcgen05 convs of csrc/conv2d_tc.cu, CUDA-core fall-backs for 1-channel convs) -- and so do their data gradients, which
are the same kernels on flipped / transposed weights.  Nonlinearities, weight / bias gradients, the embedding MLPs, the
premodule and the loss are the kernels of csrc/unet_train.cu.

``torch.autograd.Function`` is used as the TAPE only: each Function below launches hand-written kernels in ``forward`` and in
``backward``; between them torch does bookkeeping (views, one add of the two embeddings), no arithmetic of the net.

Entry points: ``unet1d_ssm_loss`` (NNUnet1D.UNet1D, NNUnet1D.py:110-179).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib

GELU, SILU = 0, 1


def _h(dev):
    return _lib.ctx(dev), _lib.lib(), _lib.stream_ptr(dev)


_zero_cache: dict = {}


def _zeros(dev, n):
    key = (str(dev), n)
    z = _zero_cache.get(key)
    if z is None:
        z = _zero_cache[key] = torch.zeros(n, device=dev, dtype=torch.float32)
    return z


# ---- raw kernel calls -------------------------------------------------------------------------------------------------------
def conv1d_raw(x1, x2, W, E, stride, pad, fast=False):
    """out = conv1d([x1, x2], W[:, :C1+C2]) (+ folded embedding table E), no bias.  W: (Cout, Cw, K), Cw >= C1 + C2."""
    dev = x1.device
    h, L, st = _h(dev)
    N, C1, Lin = x1.shape
    C2 = 0 if x2 is None else x2.shape[1]
    Cout, Cw, K = W.shape
    Cin = C1 + C2
    Lout = (Lin + 2 * pad - K) // stride + 1
    out = torch.empty((N, Cout, Lout), device=dev, dtype=torch.float32)
    zb = _zeros(dev, Cout)
    if (Cout % 32 == 0 and Cin % 16 == 0 and C1 % 16 == 0 and
            ((K == 3 and stride == 1 and pad == 1) or (K == 4 and stride == 2 and pad == 1 and Lin >= 2) or
             (K == 1 and stride == 1 and pad == 0))):
        img = torch.empty(L.msgm_conv1d_tc_pack_bytes(Cout, Cin, K), device=dev, dtype=torch.uint8)
        _lib.check(L.msgm_conv1d_tc_pack(h, _lib.ptr(W), Cout, Cw, Cin, K, _lib.ptr(img), st))
        d = _lib.Conv1dTcDesc(x1.data_ptr(), None if x2 is None else x2.data_ptr(), img.data_ptr(), zb.data_ptr(),
                              None if E is None else E.data_ptr(), out.data_ptr(), N, C1, C2, Cout, K, stride, Lin, 0, int(fast))
        _lib.check(L.msgm_conv1d_tc(h, C.byref(d), st))
        return out
    d = _lib.Conv1dDesc(x1.data_ptr(), None if x2 is None else x2.data_ptr(), W.data_ptr(), zb.data_ptr(),
                        None if E is None else E.data_ptr(), out.data_ptr(), N, C1, C2, Cw - Cin if E is not None else 0, Cout, K,
                        stride, pad, Lin, Lout, 0)
    if E is None and Cw != Cin:  # the CUDA-core kernel reads W with row length Cw = C1 + C2 + Cemb: pass a compact copy
        Wc = W[:, :Cin, :].contiguous()
        d.W = Wc.data_ptr()
    _lib.check(L.msgm_conv1d(h, C.byref(d), st))
    return out


def convt1d_raw(x, W, Lout):
    """ConvTranspose1d(k4, s2, p1) without bias; W: (Cin, Cout, 4); positions >= 2 Lin (padding to the skip length) are 0."""
    dev = x.device
    h, L, st = _h(dev)
    N, Cin, Lin = x.shape
    Cout = W.shape[1]
    zb = _zeros(dev, Cout)
    alloc = torch.zeros if Lout > 2 * Lin else torch.empty
    out = alloc((N, Cout, Lout), device=dev, dtype=torch.float32)
    if Cin % 16 == 0 and Cout % 16 == 0:
        img = torch.empty(24 * Cin * Cout, device=dev, dtype=torch.uint8)
        _lib.check(L.msgm_convt1d_tc_pack(h, _lib.ptr(W), Cout, Cin, _lib.ptr(img), st))
        _lib.check(L.msgm_convt1d_tc(h, _lib.ptr(x), _lib.ptr(img), _lib.ptr(zb), _lib.ptr(out), N, Cin, Cout, Lin, Lout, 0, st))
    else:
        _lib.check(L.msgm_convt1d_k4s2(h, _lib.ptr(x), _lib.ptr(W), _lib.ptr(zb), _lib.ptr(out), N, Cin, Cout, Lin, Lout, st))
    return out


def gemm(A, B, M, N, K, lda, ldb, ta=False, tb=False, out=None, accumulate=False):
    dev = A.device
    h, L, st = _h(dev)
    if out is None:
        out = torch.empty((M, N), device=dev, dtype=torch.float32)
    _lib.check(L.msgm_gemm_f32(h, _lib.ptr(A), _lib.ptr(B), _lib.ptr(out), M, N, K, lda, ldb, out.stride(0), int(ta), int(tb),
                               int(accumulate), st))
    return out


def rows_bias_add(x, bias, nrows):
    h, L, st = _h(x.device)
    _lib.check(L.msgm_rows_bias_add(h, _lib.ptr(x), _lib.ptr(bias), nrows, x.shape[1], x[0, 0].numel(), st))


def channel_sums(x, nrows):
    h, L, st = _h(x.device)
    out = torch.zeros(x.shape[1], device=x.device, dtype=torch.float32)
    _lib.check(L.msgm_channel_sums(h, _lib.ptr(x), _lib.ptr(out), nrows, x.shape[1], x[0, 0].numel(), st))
    return out


def ranged(conv, g):
    """``conv(g)`` with g brought to max|g| = 2^12 first and the result scaled back (powers of two: exact).  The tensor-core
    convs split operands into fp16 hi + lo; cotangents of the deep layers (~1e-7) would fall into the fp16 subnormal range."""
    h, L, st = _h(g.device)
    amax = torch.empty(1, device=g.device, dtype=torch.float32)
    _lib.check(L.msgm_amax(h, _lib.ptr(g), g.numel(), _lib.ptr(amax), st))
    gs = torch.empty_like(g)
    _lib.check(L.msgm_pow2_scale(h, _lib.ptr(g), _lib.ptr(gs), g.numel(), _lib.ptr(amax), 12, 0, st))
    out = conv(gs)
    _lib.check(L.msgm_pow2_scale(h, _lib.ptr(out), _lib.ptr(out), out.numel(), _lib.ptr(amax), 12, 1, st))
    return out


def conv_wgrad(cot, in1, in2, gW, coff, KH, KW, stride, pad, up, Hi, Wi, Ho, Wo):
    h, L, st = _h(cot.device)
    C2 = 0 if in2 is None else in2.shape[1]
    _lib.check(L.msgm_conv_wgrad(h, _lib.ptr(cot), _lib.ptr(in1), _lib.ptr(in2), _lib.ptr(gW), cot.shape[0], cot.shape[1],
                                 in1.shape[1], C2, gW.shape[1], coff, KH, KW, stride, pad, up, Hi, Wi, Ho, Wo, st))


# ---- autograd Functions = tape entries; all arithmetic is in the kernels -------------------------------------------------------
class PairAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, act):
        z = z.contiguous()
        h, L, st = _h(z.device)
        out = torch.empty_like(z)
        _lib.check(L.msgm_pair_act(h, _lib.ptr(z), None, _lib.ptr(out), z.numel() // 2, act, st))
        ctx.save_for_backward(z)
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, g):
        (z,) = ctx.saved_tensors
        g = g.contiguous()
        h, L, st = _h(z.device)
        out = torch.empty_like(z)
        _lib.check(L.msgm_pair_act(h, _lib.ptr(z), _lib.ptr(g), _lib.ptr(out), z.numel() // 2, ctx.act, st))
        return out, None


class LinearPair(torch.autograd.Function):
    """nn.Linear on a pair (2B, in): y = x W^T, bias on the primal half."""

    @staticmethod
    def forward(ctx, x, W, b):
        x = x.contiguous()
        N, K = x.shape
        out = gemm(x, W, N, W.shape[0], K, K, W.shape[1], tb=True)
        rows_bias_add(out.view(N, -1, 1), b, N // 2)
        ctx.save_for_backward(x, W)
        return out

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        g = g.contiguous()
        N, K = x.shape
        O = W.shape[0]
        gx = gemm(g, W, N, K, O, O, K) if ctx.needs_input_grad[0] else None
        gW = gemm(g, x, O, K, N, O, K, ta=True)
        gb = channel_sums(g.view(N, O, 1), N // 2)
        return gx, gW, gb


class Conv1dPair(torch.autograd.Function):
    """Conv1d over the channel concat [x1, x2, emb broadcast along the signal] on a pair (NNUnet1D.py:13-24,156-176).

    The 128 embedding channels are constant along the signal, so they are folded into a per-(sample, out-channel, tap) table
    (msgm_emb_fold) instead of being concatenated; the fold is linear in emb, so the tangent half uses the tangent embedding."""

    @staticmethod
    def forward(ctx, x1, x2, emb, W, b, stride, pad):
        dev = x1.device
        h, L, st = _h(dev)
        x1 = x1.contiguous()
        x2 = None if x2 is None else x2.contiguous()
        N, C1, Lin = x1.shape
        C2 = 0 if x2 is None else x2.shape[1]
        Cout, Cw, K = W.shape
        E = None
        if emb is not None:
            emb = emb.contiguous()
            E = torch.empty((N, Cout, K), device=dev, dtype=torch.float32)
            _lib.check(L.msgm_emb_fold(h, _lib.ptr(W), _lib.ptr(emb), _lib.ptr(E), Cw, C1 + C2, Cw - C1 - C2, Cout, K, N, st))
        out = conv1d_raw(x1, x2, W, E, stride, pad)
        rows_bias_add(out, b, N // 2)
        ctx.save_for_backward(x1, x2, emb, W)
        ctx.geom = (stride, pad, Lin, out.shape[-1])
        return out

    @staticmethod
    def backward(ctx, g):
        x1, x2, emb, W = ctx.saved_tensors
        stride, pad, Lin, Lout = ctx.geom
        g = g.contiguous()
        dev = g.device
        h, L, st = _h(dev)
        N, C1, _ = x1.shape
        C2 = 0 if x2 is None else x2.shape[1]
        Cin = C1 + C2
        Cout, Cw, K = W.shape
        gx1 = gx2 = gemb = None
        if ctx.needs_input_grad[0] or (x2 is not None and ctx.needs_input_grad[1]):
            if stride == 1:   # data gradient = the same conv with flipped taps and swapped channel roles
                Wd = W[:, :Cin, :].flip(-1).transpose(0, 1).contiguous()
                gx = ranged(lambda t_: conv1d_raw(t_, None, Wd, None, 1, K - 1 - pad), g)
            else:             # k4 s2 p1: data gradient = ConvTranspose1d(k4, s2, p1) with the weight read as (in = Cout, out = Cin)
                if Lin != 2 * Lout:
                    raise NotImplementedError("hand-written U-Net training: odd signal length at a stride-2 conv")
                Wt = W[:, :Cin, :].contiguous()
                gx = ranged(lambda t_: convt1d_raw(t_, Wt, Lin), g)
            gx1 = gx[:, :C1]
            gx2 = gx[:, C1:] if x2 is not None else None
        gW = torch.zeros_like(W)
        conv_wgrad(g, x1, x2, gW, 0, 1, K, stride, pad, 1, 1, Lin, 1, Lout)
        if emb is not None:
            Cemb = Cw - Cin
            Eb = torch.empty((N, Cout, K), device=dev, dtype=torch.float32)  # cotangent of the folded table
            _lib.check(L.msgm_tap_sums_1d(h, _lib.ptr(g), _lib.ptr(Eb), N, Cout, K, stride, pad, Lin, Lout, st))
            Wemb = W[:, Cin:, :].permute(0, 2, 1).reshape(Cout * K, Cemb).contiguous()
            gemb = gemm(Eb.view(N, Cout * K), Wemb, N, Cemb, Cout * K, Cout * K, Cemb)
            gWe = gemm(Eb.view(N, Cout * K), emb, Cout * K, Cemb, N, Cout * K, Cemb, ta=True)
            gW[:, Cin:, :] = gWe.view(Cout, K, Cemb).permute(0, 2, 1)
        gb = channel_sums(g, N // 2)
        return gx1, gx2, gemb, gW, gb, None, None


class ConvT1dPair(torch.autograd.Function):
    """ConvTranspose1d(k4, s2, p1) on a pair, output length Lout >= 2 Lin (zero-padded on the right like the reference's F.pad,
    NNUnet1D.py:170-173)."""

    @staticmethod
    def forward(ctx, x, W, b, Lout):
        x = x.contiguous()
        N, _, Lin = x.shape
        if Lout != 2 * Lin:
            raise NotImplementedError("hand-written U-Net training: signal lengths must stay even down the encoder "
                                      "(no right padding of the up-sampled signal)")
        out = convt1d_raw(x, W, Lout)
        rows_bias_add(out, b, N // 2)
        ctx.save_for_backward(x, W)
        ctx.Lout = Lout
        return out

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        g = g.contiguous()
        N, Cin, Lin = x.shape
        # data gradient: Conv1d(k4, s2, p1) with the weight read as (out = Cin, in = Cout), no flip
        gx = ranged(lambda t_: conv1d_raw(t_, None, W, None, 2, 1), g) if ctx.needs_input_grad[0] else None
        gW = torch.zeros_like(W)  # gW[ci][co][k] = sum x[ci][p] g[co][2p - 1 + k]: a conv weight gradient with the roles swapped
        conv_wgrad(x, g, None, gW, 0, 1, 4, 2, 1, 1, 1, ctx.Lout, 1, Lin)
        gb = channel_sums(g, N // 2)
        return gx, gW, gb, None


class SparseSsmLoss(torch.autograd.Function):
    """loss_b = q . adot + |a|^2 / 2 (+ beta |v|^2 / 2 for the additive SDE) from the output pair (2B, d)."""

    @staticmethod
    def forward(ctx, a_pair, y, v, t, sd):
        a_pair = a_pair.contiguous()
        h, L, st = _h(a_pair.device)
        B = y.shape[0]
        loss = torch.empty(B, device=y.device, dtype=torch.float32)
        _lib.check(L.msgm_sparse_ssm_loss(h, C.byref(sd), _lib.ptr(a_pair), _lib.ptr(y), _lib.ptr(v), _lib.ptr(t), None,
                                          _lib.ptr(loss), B, st))
        ctx.save_for_backward(a_pair, y, v, t)
        ctx.sd = sd
        return loss

    @staticmethod
    def backward(ctx, gout):
        a_pair, y, v, t = ctx.saved_tensors
        h, L, st = _h(a_pair.device)
        cot = torch.empty_like(a_pair)
        _lib.check(L.msgm_sparse_ssm_loss(h, C.byref(ctx.sd), _lib.ptr(a_pair), _lib.ptr(y), _lib.ptr(v), _lib.ptr(t),
                                          _lib.ptr(gout.contiguous()), _lib.ptr(cot), y.shape[0], st))
        return cot, None, None, None, None


def _embed_mlp_pair(mlp, x_pair):
    """nn.Sequential(Linear(1, E), GELU, Linear(E, E)) (NNUnet1D.py:26-27) on a pair (2B, 1)."""
    l1, l2 = mlp[0], mlp[2]
    return LinearPair.apply(PairAct.apply(LinearPair.apply(x_pair, l1.weight, l1.bias), GELU), l2.weight, l2.bias)


def unet1d_pair_forward(net, y, v, s):
    """(a; adot) of NNUnet1D.UNet1D at (y, s) along v: (2B, L)."""
    dev = y.device
    h, L, st = _h(dev)
    B, Ls = y.shape
    s_pair = torch.cat([s.reshape(B, 1), torch.zeros(B, 1, device=dev)], 0)  # the time input has no tangent
    emb = _embed_mlp_pair(net.time_mlp, s_pair)
    if net.premodule is not None:
        x_pair = torch.empty((2 * B, Ls), device=dev, dtype=torch.float32)
        logn = torch.empty(2 * B, device=dev, dtype=torch.float32)
        _lib.check(L.msgm_premodule_pair(h, _lib.ptr(y), _lib.ptr(v), _lib.ptr(x_pair), _lib.ptr(logn), B, Ls,
                                         float(torch.sqrt(torch.tensor(float(Ls)))), st))
        emb = emb + _embed_mlp_pair(net.scale_embed, logn.view(2 * B, 1))
    else:
        x_pair = torch.cat([y, v], 0)
    cur, skips = x_pair.view(2 * B, 1, Ls), []

    def block(blk, x1, x2):
        c1, c2 = blk.net[0], blk.net[2]
        z = Conv1dPair.apply(x1, x2, emb, c1.weight, c1.bias, 1, 1)
        z = Conv1dPair.apply(PairAct.apply(z, GELU), None, None, c2.weight, c2.bias, 1, 1)
        return PairAct.apply(z, GELU)

    for blk, down in zip(net.enc_blocks, net.downs):
        cur = block(blk, cur, None)
        skips.append(cur)
        cur = Conv1dPair.apply(cur, None, None, down.weight, down.bias, 2, 1)
    cur = block(net.middle, cur, None)
    for up, blk in zip(net.up_convs, net.dec_blocks):
        skip = skips.pop()
        cur = ConvT1dPair.apply(cur, up.weight, up.bias, skip.shape[-1])
        cur = block(blk, cur, skip)
    out = Conv1dPair.apply(cur, None, None, net.final.weight, net.final.bias, 1, 0)
    return out.view(2 * B, Ls)


def unet1d_ssm_loss(gen, t_, y, v):
    """Per-sample SSM loss (B,) of a UNet1D score net on the hand-written kernels; differentiable w.r.t. the net parameters."""
    base, net = gen.base_sde, gen.a
    dev = y.device
    sd, keep = base.desc(dev)
    sd.dim = y.shape[1]
    yc, vc = _lib.f32c(y, dev), _lib.f32c(v, dev)
    tc = _lib.f32c(t_.reshape(-1), dev)
    a_pair = unet1d_pair_forward(net, yc, vc, tc)
    return SparseSsmLoss.apply(a_pair, yc, vc, tc, sd)


def supported(gen, y) -> bool:
    """True when the hand-written training path covers this score net / SDE / batch (``gen.unet_train_kernels = False``
    forces the library path, which is kept for comparison and for configurations outside this list)."""
    from . import NNUnet1D, SDEs
    if not getattr(gen, "unet_train_kernels", True) or not y.is_cuda or y.dim() != 2:
        return False
    base, net = gen.base_sde, gen.a
    if not (isinstance(base, SDEs.SGMsde) or getattr(base, "sparseTensor", False)):
        return False
    if isinstance(net, NNUnet1D.UNet1D):
        L, n = y.shape[1], len(net.downs)
        return L % (1 << n) == 0 and net.input_dim == L
    return False


def ssm_loss(gen, t_, y, v):
    from . import NNUnet1D
    if isinstance(gen.a, NNUnet1D.UNet1D):
        return unet1d_ssm_loss(gen, t_, y, v)
    raise NotImplementedError
