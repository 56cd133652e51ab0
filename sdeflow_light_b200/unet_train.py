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

import os

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


# ---- weight gradients as a parallel branch of a captured iteration graph ----------------------------------------------------
# A conv's weight gradient is a LEAF of the backward pass: nothing but the optimiser reads it, while the data gradient next
# to it is on the critical path of the chain.  While train.GraphedSsmStep captures its iteration graph the tensor-core
# weight-gradient launches therefore go on a side stream -- forked after the cotangent and its range words exist, joined once
# in front of the optimiser (join_leaf_stream) -- and run beside the data-gradient chain when the graph is replayed.  Every
# tensor the branch reads is handed to the allocator with record_stream, so its block is not reused for the rest of the
# capture (the main stream could otherwise overwrite it before the branch has run).  Eager iterations stay on one stream.
LEAF_STREAM = os.environ.get("MSGM_LEAF_STREAM", "1") != "0"
LEAF_CAPTURE = False  # set by train.GraphedSsmStep around its own captures: only a capture that joins the branch may fork it
_leaf_streams: dict = {}
_leaf_forked: set = set()


class _leaf_branch:
    def __init__(self, dev, *reads):
        self.dev, self.reads, self.cm = dev, [t for t in reads if t is not None], None

    def __enter__(self):
        if LEAF_STREAM and LEAF_CAPTURE and torch.cuda.is_current_stream_capturing():
            side = _leaf_streams.get(self.dev.index)
            if side is None:
                side = _leaf_streams[self.dev.index] = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            for t in self.reads:
                t.record_stream(side)
            _leaf_forked.add(self.dev.index)
            self.cm = torch.cuda.stream(side)
            self.cm.__enter__()
        return self

    def __exit__(self, *exc):
        if self.cm is not None:
            self.cm.__exit__(*exc)
        return False


def join_leaf_stream(dev):
    """The current stream waits for the weight-gradient branch (no-op when nothing was forked)."""
    if dev.index in _leaf_forked:
        torch.cuda.current_stream(dev).wait_stream(_leaf_streams[dev.index])
        _leaf_forked.discard(dev.index)


# ---- raw kernel calls -------------------------------------------------------------------------------------------------------
def conv1d_raw(x1, x2, W, E, stride, pad, fast=False, dgrad=False):
    """out = conv1d([x1, x2], W[:, :C1+C2]) (+ folded embedding table E), no bias.  W: (Cout, Cw, K), Cw >= C1 + C2.
    dgrad=True: W is the weight of the FORWARD conv (stride 1, "same") and the call evaluates that conv's data gradient, i.e. the
    conv with weight W[:, :C, :] transposed and flipped, on x1 (a cotangent, W.shape[0] channels) -> (N, C, L); C = number of
    real (non-embedding) input channels is given as E (an int) in that case."""
    dev = x1.device
    h, L, st = _h(dev)
    N, C1, Lin = x1.shape
    if dgrad:
        Cf_out, Cw, K = W.shape
        Cf_in = int(E)
        out = torch.empty((N, Cf_in, Lin), device=dev, dtype=torch.float32)
        zb = _zeros(dev, Cf_in)
        if Cf_in % 32 == 0 and Cf_out % 16 == 0 and K in (1, 3) and pad == K // 2:
            img = torch.empty(L.msgm_conv1d_tc_pack_bytes(Cf_in, Cf_out, K), device=dev, dtype=torch.uint8)
            _lib.check(L.msgm_conv_tc_pack_dgrad(h, _lib.ptr(W), Cf_out, Cw, Cf_in, K, _lib.ptr(img), st))
            d = _lib.Conv1dTcDesc(x1.data_ptr(), None, img.data_ptr(), zb.data_ptr(), None, out.data_ptr(), N, Cf_out, 0, Cf_in, K,
                                  1, Lin, 0, int(fast))
            _lib.check(L.msgm_conv1d_tc(h, C.byref(d), st))
            return out
        Wd = W[:, :Cf_in, :].flip(-1).transpose(0, 1).contiguous()
        return conv1d_raw(x1, None, Wd, None, 1, K - 1 - pad, fast)
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


def bias_grad(g, nrows):
    """Bias gradient = channel sums of the cotangent: a leaf of the backward pass (side branch of a captured graph)."""
    with _leaf_branch(g.device, g):
        return channel_sums(g, nrows)


def amax_of(g):
    """Device word holding max|g| (float bits): the range scaling of the tensor-core data / weight gradient kernels."""
    h, L, st = _h(g.device)
    amax = torch.empty(1, device=g.device, dtype=torch.float32)
    _lib.check(L.msgm_amax(h, _lib.ptr(g), g.numel(), _lib.ptr(amax), st))
    return amax


def amax_pair(cot, in1, in2=None):
    """(max|cot|, max|[in1, in2]|) as two device words from ONE launch: the range words of a conv's data and weight gradients."""
    h, L, st = _h(cot.device)
    out = torch.empty(2, device=cot.device, dtype=torch.float32)
    _lib.check(L.msgm_amax2(h, _lib.ptr(cot), cot.numel(), _lib.ptr(in1), in1.numel(), _lib.ptr(in2),
                            0 if in2 is None else in2.numel(), _lib.ptr(out), st))
    return out[0:1], out[1:2]


def ranged(conv, g, amax=None, tc=True):
    """``conv(g)`` evaluated on g times the power of two that brings max|g| into [2^14, 2^15), result scaled back (exact).  The tensor-core
    convs split operands into fp16 hi + lo; cotangents of the deep layers (~1e-7) would fall into the fp16 subnormal range.
    tc=True: ``conv`` is ONE tensor-core conv launch, which scales while staging and unscales in its epilogue
    (msgm_tc_range_scale, one-shot); otherwise (CUDA-core fall-back shapes: fp32, no scaling needed) the conv runs as is."""
    h, L, st = _h(g.device)
    if not tc:
        return conv(g)
    if amax is None:
        amax = amax_of(g)
    _lib.check(L.msgm_tc_range_scale(h, _lib.ptr(amax)))
    try:
        return conv(g)
    finally:
        _lib.check(L.msgm_tc_range_scale(h, None))


WGRAD_TC = True  # weight gradients on tcgen05 (csrc/conv_wgrad_tc.cu) where the shape allows; False: fp32 CUDA-core kernel


def conv_wgrad(cot, in1, in2, Wshape, KH, KW, stride, pad, up, Hi, Wi, Ho, Wo, amax=None, amax_in=None):
    """Weight gradient (tensor of shape Wshape = (Cout, Cw, ...)) of a conv over the channels [0, C1 + C2) of its input axis; any
    further input channels of the weight (the 1-D U-Net's folded embedding channels) are left for the caller to fill."""
    h, L, st = _h(cot.device)
    C1, C2 = in1.shape[1], 0 if in2 is None else in2.shape[1]
    N, Cout, Cw = cot.shape[0], cot.shape[1], Wshape[1]
    if WGRAD_TC and L.msgm_conv_wgrad_tc_ok(N, Cout, C1, C2, KH, KW, stride, pad, up, Hi, Wi):
        if amax is None or amax_in is None:  # both operands are range-scaled: an activation tensor of small magnitude would
            amax, amax_in = amax_pair(cot, in1, in2)  # otherwise lose the low part of its fp16 split in the subnormals
        nb = L.msgm_conv_wgrad_tc_scratch_bytes(h, N, Cout, C1 + C2, KH, KW, stride, pad, up, Hi, Wi)
        with _leaf_branch(cot.device, cot, in1, in2, amax, amax_in):
            gW = torch.empty(Wshape, device=cot.device, dtype=torch.float32)  # the kernel overwrites its block
            scratch = torch.empty(nb, device=cot.device, dtype=torch.uint8)
            _lib.check(L.msgm_conv_wgrad_tc(h, _lib.ptr(cot), _lib.ptr(in1), _lib.ptr(in2), _lib.ptr(gW), _lib.ptr(amax),
                                            _lib.ptr(amax_in), _lib.ptr(scratch), N, Cout, C1, C2, Cw, 0, KH, KW, stride, pad, up,
                                            Hi, Wi, 0, _lib.stream_ptr(cot.device)))
        return gW
    if Cw == C1 + C2:  # the whole tensor is this kernel's: zero fill + accumulation on the side branch as well
        with _leaf_branch(cot.device, cot, in1, in2):
            gW = torch.zeros(Wshape, device=cot.device, dtype=torch.float32)
            _lib.check(L.msgm_conv_wgrad(h, _lib.ptr(cot), _lib.ptr(in1), _lib.ptr(in2), _lib.ptr(gW), N, Cout, C1, C2, Cw, 0, KH,
                                         KW, stride, pad, up, Hi, Wi, Ho, Wo, _lib.stream_ptr(cot.device)))
        return gW
    gW = torch.zeros(Wshape, device=cot.device, dtype=torch.float32)  # the caller fills the other channels: one stream
    _lib.check(L.msgm_conv_wgrad(h, _lib.ptr(cot), _lib.ptr(in1), _lib.ptr(in2), _lib.ptr(gW), N, Cout, C1, C2, Cw, 0, KH, KW,
                                 stride, pad, up, Hi, Wi, Ho, Wo, st))
    return gW


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
        amax, amax_in = amax_pair(g, x1, x2)
        # the weight gradient first: in a captured graph it is a side branch (conv_wgrad) beside the data gradient below
        gW = conv_wgrad(g, x1, x2, W.shape, 1, K, stride, pad, 1, 1, Lin, 1, Lout, amax, amax_in)
        if ctx.needs_input_grad[0] or (x2 is not None and ctx.needs_input_grad[1]):
            if stride == 1:   # data gradient = the same conv with flipped taps and swapped channel roles
                gx = ranged(lambda t_: conv1d_raw(t_, None, W, Cin, 1, pad, dgrad=True), g, amax)
            else:             # k4 s2 p1: data gradient = ConvTranspose1d(k4, s2, p1) with the weight read as (in = Cout, out = Cin)
                if Lin != 2 * Lout:
                    raise NotImplementedError("hand-written U-Net training: odd signal length at a stride-2 conv")
                Wt = W[:, :Cin, :].contiguous()
                gx = ranged(lambda t_: convt1d_raw(t_, Wt, Lin), g, amax)
            gx1 = gx[:, :C1]
            gx2 = gx[:, C1:] if x2 is not None else None
        if emb is not None:
            Cemb = Cw - Cin
            Eb = torch.empty((N, Cout, K), device=dev, dtype=torch.float32)  # cotangent of the folded table
            _lib.check(L.msgm_tap_sums_1d(h, _lib.ptr(g), _lib.ptr(Eb), N, Cout, K, stride, pad, Lin, Lout, st))
            Wemb = W[:, Cin:, :].permute(0, 2, 1).reshape(Cout * K, Cemb).contiguous()
            gemb = gemm(Eb.view(N, Cout * K), Wemb, N, Cemb, Cout * K, Cout * K, Cemb)
            gWe = gemm(Eb.view(N, Cout * K), emb, Cout * K, Cemb, N, Cout * K, Cemb, ta=True)
            gW[:, Cin:, :] = gWe.view(Cout, K, Cemb).permute(0, 2, 1)
        gb = bias_grad(g, N // 2)
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
        amax_x, amax = amax_pair(x, g)
        # gW[ci][co][k] = sum x[ci][p] g[co][2p - 1 + k]: a conv weight gradient with the roles swapped: the kernel's "input"
        # operand is the cotangent here, so the range scaling goes to that side
        gW = conv_wgrad(x, g, None, W.shape, 1, 4, 2, 1, 1, 1, ctx.Lout, 1, Lin, amax=amax_x, amax_in=amax)
        gx = ranged(lambda t_: conv1d_raw(t_, None, W, None, 2, 1), g, amax) if ctx.needs_input_grad[0] else None
        gb = bias_grad(g, N // 2)
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
    from . import NNUnet
    if isinstance(net, NNUnet.VorticityUNet):
        S, n = net.in_space, len(net.core.channel_mult) - 1
        return y.shape[1] == S * S and S % (1 << n) == 0 and net.core.dropout == 0
    return False


def ssm_loss(gen, t_, y, v):
    from . import NNUnet1D
    if isinstance(gen.a, NNUnet1D.UNet1D):
        return unet1d_ssm_loss(gen, t_, y, v)
    return unet2d_ssm_loss(gen, t_, y, v)


# =====================================================================================================================
# 2-D U-Net (NNUnet.VorticityUNet over model/unet.py UNetModel): GroupNorm, SiLU, 3x3 / 1x1 convs, attention, resampling
# =====================================================================================================================
def conv2d_raw(x, W, ebias, stride, up, dgrad=False):
    """out = conv2d(nearest-upsample^{up}(x), W, padding K//2, stride) + ebias[n, co]; no per-channel bias.
    dgrad=True (stride 1, up 1): W is the weight of the FORWARD conv and the call evaluates that conv's data gradient on the
    cotangent x, i.e. the conv with W transposed (channel roles) and flipped (taps), packed straight from W."""
    from .model.unet import _tc_shape_ok
    dev = x.device
    h, L, st = _h(dev)
    N, Cin, Hs, Ws = x.shape
    p = lambda t_: None if t_ is None else t_.data_ptr()  # noqa: E731
    if dgrad:
        Cf_out, Cf_in, K = W.shape[0], W.shape[1], W.shape[-1]
        if _tc_shape_ok(Cf_in, Cf_out, 0, K, 1, Hs, Ws):
            out = torch.empty((N, Cf_in, Hs, Ws), device=dev, dtype=torch.float32)
            img = torch.empty(L.msgm_conv2d_tc_pack_bytes(Cf_in, Cf_out, K), device=dev, dtype=torch.uint8)
            _lib.check(L.msgm_conv_tc_pack_dgrad(h, _lib.ptr(W), Cf_out, Cf_in, Cf_in, K * K, _lib.ptr(img), st))
            d = _lib.Conv2dTcDesc(p(x), None, p(img), None, None, None, None, p(out), N, Cf_out, 0, Cf_in, K, 1, 1, Hs, Ws, 0, 0)
            _lib.check(L.msgm_conv2d_tc(h, C.byref(d), st))
            return out
        return conv2d_raw(x, W.flip(2, 3).transpose(0, 1).contiguous(), None, 1, 1)
    Cout, K = W.shape[0], W.shape[-1]
    pad = K // 2
    Ho, Wo = (Hs * up + 2 * pad - K) // stride + 1, (Ws * up + 2 * pad - K) // stride + 1
    out = torch.empty((N, Cout, Ho, Wo), device=dev, dtype=torch.float32)
    if _tc_shape_ok(Cout, Cin, 0, K, stride, Hs * up, Ws * up):
        img = torch.empty(L.msgm_conv2d_tc_pack_bytes(Cout, Cin, K), device=dev, dtype=torch.uint8)
        _lib.check(L.msgm_conv2d_tc_pack(h, _lib.ptr(W), Cout, Cin, K, _lib.ptr(img), st))
        d = _lib.Conv2dTcDesc(p(x), None, p(img), None, p(ebias), None, None, p(out), N, Cin, 0, Cout, K, stride, up, Hs, Ws, 0, 0)
        _lib.check(L.msgm_conv2d_tc(h, C.byref(d), st))
        return out
    d = _lib.Conv2dDesc(p(x), None, p(W), None, p(ebias), None, None, None, None, p(out), N, Cin, 0, Cout, K, stride, up, Hs, Ws,
                        0, 0)
    _lib.check(L.msgm_conv2d(h, C.byref(d), st))
    return out


def resample2(x, mode):
    h, L, st = _h(x.device)
    N, Cc, H, W = x.shape
    out = torch.empty((N, Cc, 2 * H, 2 * W) if mode == 0 else (N, Cc, H // 2, W // 2), device=x.device, dtype=torch.float32)
    _lib.check(L.msgm_resample2(h, _lib.ptr(x), _lib.ptr(out), N * Cc, H if mode == 0 else H // 2, W if mode == 0 else W // 2,
                                mode, st))
    return out


def sample_channel_sums(g):
    """(N, C) sums of g (N, C, ...) over the positions."""
    h, L, st = _h(g.device)
    N, Cc = g.shape[:2]
    P_ = g[0, 0].numel()
    out = torch.empty((N, Cc), device=g.device, dtype=torch.float32)
    _lib.check(L.msgm_tap_sums_1d(h, _lib.ptr(g), _lib.ptr(out), N, Cc, 1, 1, 0, P_, P_, st))
    return out


class Conv2dPair(torch.autograd.Function):
    """conv_nd(2, ...) of model/unet.py on a pair: 3x3 (padding 1) or 1x1, stride 1 or 2, optionally on the nearest-upsampled
    input (Upsample), plus a per-sample, per-channel term e (the ResBlock's embedding projection); bias on the primal half."""

    @staticmethod
    def forward(ctx, x, W, b, e, stride, up):
        x = x.contiguous()
        N = x.shape[0]
        Cout = W.shape[0]
        eb = torch.zeros((N, Cout), device=x.device, dtype=torch.float32) if e is None else e.contiguous().clone()
        if b is not None:
            rows_bias_add(eb.view(N, Cout, 1), b, N // 2)
        out = conv2d_raw(x, W, eb, stride, up)
        ctx.save_for_backward(x, W)
        ctx.cfg = (stride, up, b is not None, e is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        stride, up, has_b, has_e = ctx.cfg
        g = g.contiguous()
        N, Cin, Hs, Ws = x.shape
        Cout, K = W.shape[0], W.shape[-1]
        Ho, Wo = g.shape[-2:]
        gx = None
        amax, amax_in = amax_pair(g, x)
        gW = conv_wgrad(g, x, None, W.shape, K, K, stride, K // 2, up, Hs, Ws, Ho, Wo, amax, amax_in)  # side branch when captured
        if ctx.needs_input_grad[0]:
            # data gradient: the same conv with flipped taps and swapped channel roles (image packed straight from W)
            src = resample2(g, 0) if stride == 2 else g     # stride 2: cotangent back on the input grid (zeros in between)
            gx = ranged(lambda t_: conv2d_raw(t_, W, None, 1, 1, dgrad=True), src, amax)
            if up == 2:                                      # adjoint of the nearest-neighbour upsampling
                gx = resample2(gx, 1)
        gb = bias_grad(g, N // 2) if has_b else None
        ge = sample_channel_sums(g) if has_e else None
        return gx, gW, gb, ge, None, None


class GnPair(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, G):
        x = x.contiguous()
        h, L, st = _h(x.device)
        N, Cc = x.shape[:2]
        HW = x[0, 0].numel()
        out = torch.empty_like(x)
        stats = torch.empty((N // 2, G, 4), device=x.device, dtype=torch.float32)
        _lib.check(L.msgm_gn_pair(h, _lib.ptr(x), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(stats), None, _lib.ptr(out), None,
                                  None, N // 2, Cc, G, HW, st))
        ctx.save_for_backward(x, gamma, stats)
        ctx.G = G
        return out

    @staticmethod
    def backward(ctx, g):
        x, gamma, stats = ctx.saved_tensors
        g = g.contiguous()
        h, L, st = _h(x.device)
        N, Cc = x.shape[:2]
        gx = torch.empty_like(x)
        gg, gb = torch.zeros_like(gamma), torch.zeros_like(gamma)
        _lib.check(L.msgm_gn_pair(h, _lib.ptr(x), _lib.ptr(gamma), None, _lib.ptr(stats), _lib.ptr(g), _lib.ptr(gx), _lib.ptr(gg),
                                  _lib.ptr(gb), N // 2, Cc, ctx.G, x[0, 0].numel(), st))
        return gx, gg, gb, None


def _prob(segs, Cptr, ldc, sc, alpha=1.0):
    """One problem of a product group: C = alpha * sum over segs of op(A) op(B); seg = (A, B, lda, ldb, sa, sb, ta, tb)."""
    q = _lib.GemmProblem()
    for i, (A, B, lda, ldb, sa, sb, ta, tb) in enumerate(segs):
        q.A[i], q.B[i] = A, B
        q.lda[i], q.ldb[i], q.stride_a[i], q.stride_b[i], q.trans_a[i], q.trans_b[i] = lda, ldb, sa, sb, int(ta), int(tb)
    q.C, q.ldc, q.stride_c, q.nseg, q.alpha, q.accumulate = Cptr, ldc, sc, len(segs), float(alpha), 0
    return q


def _gemm_group(dev, probs, M, N, K, batch):
    h, L, st = _h(dev)
    arr = (_lib.GemmProblem * len(probs))(*probs)
    _lib.check(L.msgm_gemm_group_f32(h, arr, len(probs), M, N, K, batch, st))


class AttnPair(torch.autograd.Function):
    """QKVAttention (model/unet.py:236-250, one head) on a pair qkv (2B, 3C, T): S = s^2 q^T k, P = softmax(S), O = v P^T and
    the tangents Sdot = s^2 (qdot^T k + q^T kdot), Pdot = P (Sdot - rowsum(P Sdot)), Odot = vdot P^T + v Pdot^T; the backward is
    the hand-derived adjoint.  The 21 small batched products run as FOUR grouped launches (msgm_gemm_group_f32: problems of one
    shape side by side, sums of two products inside one problem) around two row kernels (msgm_softmax_pair)."""

    @staticmethod
    def forward(ctx, qkv):
        qkv = qkv.contiguous()
        dev = qkv.device
        N, C3, T = qkv.shape
        B, Cc = N // 2, C3 // 3
        s2 = 1.0 / math.sqrt(Cc)  # (C^-1/4)^2
        f, sq, TT, so = 4, C3 * T, T * T, Cc * T  # bytes per float; sample strides (floats) of qkv, of a logit matrix, of the output
        base = qkv.data_ptr()

        def ptr(sample0, part):
            return base + f * (sample0 * sq + part * Cc * T)

        q, k, v, qd, kd, vd = ptr(0, 0), ptr(0, 1), ptr(0, 2), ptr(B, 0), ptr(B, 1), ptr(B, 2)
        S = torch.empty((B, T, T), device=dev, dtype=torch.float32)
        Sd = torch.empty_like(S)
        _gemm_group(dev, [_prob([(q, k, T, T, sq, sq, 1, 0)], S.data_ptr(), T, TT, s2),
                          _prob([(qd, k, T, T, sq, sq, 1, 0), (q, kd, T, T, sq, sq, 1, 0)], Sd.data_ptr(), T, TT, s2)], T, T, Cc, B)
        Pm, Pd = torch.empty_like(S), torch.empty_like(S)
        h, L, st = _h(dev)
        _lib.check(L.msgm_softmax_pair(h, _lib.ptr(S), _lib.ptr(Sd), None, None, _lib.ptr(Pm), _lib.ptr(Pd), B * T, T, st))
        out = torch.empty((N, Cc, T), device=dev, dtype=torch.float32)
        o, od, P_, Pd_ = out.data_ptr(), out.data_ptr() + f * B * so, Pm.data_ptr(), Pd.data_ptr()
        _gemm_group(dev, [_prob([(v, P_, T, T, sq, TT, 0, 1)], o, T, so),
                          _prob([(vd, P_, T, T, sq, TT, 0, 1), (v, Pd_, T, T, sq, TT, 0, 1)], od, T, so)], Cc, T, T, B)
        ctx.save_for_backward(qkv, Pm, Pd, Sd)
        return out

    @staticmethod
    def backward(ctx, g):
        qkv, Pm, Pd, Sd = ctx.saved_tensors
        g = g.contiguous()
        dev = qkv.device
        N, C3, T = qkv.shape
        B, Cc = N // 2, C3 // 3
        s2 = 1.0 / math.sqrt(Cc)
        f, sq, so, TT = 4, C3 * T, Cc * T, T * T
        base = qkv.data_ptr()

        def ptr(sample0, part):
            return base + f * (sample0 * sq + part * Cc * T)

        q, k, v, qd, kd, vd = ptr(0, 0), ptr(0, 1), ptr(0, 2), ptr(B, 0), ptr(B, 1), ptr(B, 2)
        ob, odb = g.data_ptr(), g.data_ptr() + f * B * so  # cotangents of O and Odot
        A = torch.empty((B, T, T), device=dev, dtype=torch.float32)
        Pdb = torch.empty_like(A)
        # direct cotangents of P and Pdot: A = Obar^T v + Odotbar^T vdot, Pdotbar = Odotbar^T v
        _gemm_group(dev, [_prob([(ob, v, T, T, so, sq, 1, 0), (odb, vd, T, T, so, sq, 1, 0)], A.data_ptr(), T, TT),
                          _prob([(odb, v, T, T, so, sq, 1, 0)], Pdb.data_ptr(), T, TT)], T, T, Cc, B)
        Sb, Sdb = torch.empty_like(A), torch.empty_like(A)
        h, L, st = _h(dev)
        _lib.check(L.msgm_softmax_pair(h, _lib.ptr(Pm), _lib.ptr(Sd), _lib.ptr(A), _lib.ptr(Pdb), _lib.ptr(Sb), _lib.ptr(Sdb),
                                       B * T, T, st))
        gq = torch.empty_like(qkv)
        gbase = gq.data_ptr()

        def gptr(sample0, part):
            return gbase + f * (sample0 * sq + part * Cc * T)

        P_, Pd_, Sb_, Sdb_ = Pm.data_ptr(), Pd.data_ptr(), Sb.data_ptr(), Sdb.data_ptr()
        _gemm_group(dev, [
            # vbar = Obar P + Odotbar Pdot ; vdotbar = Odotbar P
            _prob([(ob, P_, T, T, so, TT, 0, 0), (odb, Pd_, T, T, so, TT, 0, 0)], gptr(0, 2), T, sq),
            _prob([(odb, P_, T, T, so, TT, 0, 0)], gptr(B, 2), T, sq),
            # qbar = s^2 (k Sbar^T + kdot Sdotbar^T) ; qdotbar = s^2 k Sdotbar^T
            _prob([(k, Sb_, T, T, sq, TT, 0, 1), (kd, Sdb_, T, T, sq, TT, 0, 1)], gptr(0, 0), T, sq, s2),
            _prob([(k, Sdb_, T, T, sq, TT, 0, 1)], gptr(B, 0), T, sq, s2),
            # kbar = s^2 (q Sbar + qdot Sdotbar) ; kdotbar = s^2 q Sdotbar
            _prob([(q, Sb_, T, T, sq, TT, 0, 0), (qd, Sdb_, T, T, sq, TT, 0, 0)], gptr(0, 1), T, sq, s2),
            _prob([(q, Sdb_, T, T, sq, TT, 0, 0)], gptr(B, 1), T, sq, s2)], Cc, T, T, B)
        return gq


def _gn_silu(norm, x, silu=True):
    y = GnPair.apply(x, norm.weight, norm.bias, norm.num_groups)
    return PairAct.apply(y, SILU) if silu else y


def _resblock_pair(blk, x, emb_act):
    conv1, conv2 = blk.in_layers[2], blk.out_layers[3]
    lin = blk.emb_layers[1]
    e = LinearPair.apply(emb_act, lin.weight, lin.bias)                    # Linear(SiLU(emb)) on the pair
    hdn = Conv2dPair.apply(_gn_silu(blk.in_layers[0], x), conv1.weight, conv1.bias, e, 1, 1)
    hdn = Conv2dPair.apply(_gn_silu(blk.out_layers[0], hdn), conv2.weight, conv2.bias, None, 1, 1)
    if isinstance(blk.skip_connection, torch.nn.Identity):
        return x + hdn
    sc = blk.skip_connection
    return Conv2dPair.apply(x, sc.weight, sc.bias, None, 1, 1) + hdn


def _attention_pair(blk, x):
    N, Cc, Hh, Ww = x.shape
    qkv = Conv2dPair.apply(_gn_silu(blk.norm, x, silu=False), blk.qkv.weight.unsqueeze(-1), blk.qkv.bias, None, 1, 1)
    a = AttnPair.apply(qkv.view(N, 3 * Cc, Hh * Ww)).view(N, Cc, Hh, Ww)
    return x + Conv2dPair.apply(a, blk.proj_out.weight.unsqueeze(-1), blk.proj_out.bias, None, 1, 1)


def _run_layer_pair(layer, x, emb_act):
    from .model import unet as U
    if isinstance(layer, U.ResBlock):
        return _resblock_pair(layer, x, emb_act)
    if isinstance(layer, U.AttentionBlock):
        return _attention_pair(layer, x)
    if isinstance(layer, U.Upsample):
        if not layer.use_conv or layer.odd_size:
            raise NotImplementedError("hand-written U-Net training: Upsample without conv / odd sizes")
        return Conv2dPair.apply(x, layer.conv.weight, layer.conv.bias, None, 1, 2)
    if isinstance(layer, U.Downsample):
        if not isinstance(layer.op, torch.nn.Conv2d):
            raise NotImplementedError("hand-written U-Net training: Downsample without conv")
        return Conv2dPair.apply(x, layer.op.weight, layer.op.bias, None, 2, 1)
    if isinstance(layer, torch.nn.Conv2d):
        return Conv2dPair.apply(x, layer.weight, layer.bias, None, layer.stride[0], 1)
    raise NotImplementedError(f"hand-written U-Net training: layer {type(layer).__name__}")


def _sincos_mlp_pair(mlp, val_pair, dim):
    """time_embed / scale_embed (Linear, SiLU, Linear) of the sinusoidal embedding of a pair of scalars."""
    dev = val_pair.device
    h, L, st = _h(dev)
    N = val_pair.numel()
    emb = torch.empty((N, dim), device=dev, dtype=torch.float32)
    _lib.check(L.msgm_sincos_pair(h, _lib.ptr(val_pair), _lib.ptr(emb), N // 2, dim, st))
    l1, l2 = mlp[0], mlp[2]
    return LinearPair.apply(PairAct.apply(LinearPair.apply(emb, l1.weight, l1.bias), SILU), l2.weight, l2.bias)


def unet2d_pair_forward(net, y, v, s):
    """(a; adot) of NNUnet.VorticityUNet at (y, s) along v: (2B, d).  Follows VorticityUNet._forward / UNetModel.run_blocks."""
    from .NNUnet import scale_image
    dev = y.device
    h, L, st = _h(dev)
    B, d = y.shape
    S = net.in_space
    core = net.core
    s_pair = torch.cat([s.reshape(B), torch.zeros(B, device=dev)], 0).contiguous()
    emb = _sincos_mlp_pair(core.time_embed, s_pair, core.model_channels)
    if net.pre is not None:
        x_pair = torch.empty((2 * B, d), device=dev, dtype=torch.float32)
        logn = torch.empty(2 * B, device=dev, dtype=torch.float32)
        _lib.check(L.msgm_premodule_pair(h, _lib.ptr(y), _lib.ptr(v), _lib.ptr(x_pair), _lib.ptr(logn), B, d,
                                         float(torch.sqrt(torch.tensor(float(d)))) / scale_image, st))
        emb = emb + _sincos_mlp_pair(core.scale_embed, logn, core.model_channels)
    else:
        x_pair = torch.cat([y, v], 0) / scale_image
    img = x_pair.view(2 * B, 1, S, S) if net.flatten_order == "C" else x_pair.view(2 * B, 1, S, S).transpose(2, 3).contiguous()
    emb_act = PairAct.apply(emb, SILU)  # every ResBlock starts its embedding branch with the same SiLU(emb)
    skips, cur = [], img
    for blk in core.input_blocks:
        for layer in blk:
            cur = _run_layer_pair(layer, cur, emb_act)
        skips.append(cur)
    for layer in core.middle_block:
        cur = _run_layer_pair(layer, cur, emb_act)
    for blk in core.output_blocks:
        cur = torch.cat([cur, skips.pop()], dim=1)
        for layer in blk:
            cur = _run_layer_pair(layer, cur, emb_act)
    oc = core.out[2]
    out = Conv2dPair.apply(_gn_silu(core.out[0], cur), oc.weight, oc.bias, None, 1, 1) * float(scale_image)
    return (out.reshape(2 * B, d) if net.flatten_order == "C" else out.transpose(2, 3).contiguous().view(2 * B, d))


def unet2d_ssm_loss(gen, t_, y, v):
    base, net = gen.base_sde, gen.a
    dev = y.device
    sd, keep = base.desc(dev)
    sd.dim = y.shape[1]
    yc, vc = _lib.f32c(y, dev), _lib.f32c(v, dev)
    tc = _lib.f32c(t_.reshape(-1), dev)
    a_pair = unet2d_pair_forward(net, yc, vc, tc)
    return SparseSsmLoss.apply(a_pair, yc, vc, tc, sd)
