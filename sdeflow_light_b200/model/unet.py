"""Guided-diffusion style 2-D U-Net: drop-in for the reference's model/unet.py + model/nn_utils.py, restricted to what the
driver configures (MSGM_higherDim.py:703-716 via NNUnet.VorticityUNet: dims=2, no class conditioning, no gradient
checkpointing, no scale-shift norm, no potential parameterisation).  Module / parameter names are identical to the
reference so that its checkpoints load: ``input_blocks.i.j.*``, ``middle_block.*``, ``output_blocks.i.j.*``, ``out.*``,
``time_embed.*``; ResBlock = ``in_layers(GN,SiLU,conv3x3) + emb_layers(SiLU,Linear) -> out_layers(GN,SiLU,Dropout,
zero conv3x3)`` with identity / 1x1 skip; AttentionBlock = GN -> 1x1 qkv -> single-head softmax(q k / sqrt(C)) v -> zero
1x1 proj (+ residual); Up/Downsample = nearest x2 + conv3x3 / stride-2 conv3x3.

Inference (no autograd: the sampling hot path) runs on hand-written kernels (csrc/unet2d_fp32.cu, ``run_blocks_kernels``):
GroupNorm statistics -> conv with normalise+SiLU fused into the input staging, embedding term / bias / residual fused
into the epilogue, concatenations and nearest upsampling read in place, a small single-head attention kernel.  With
autograd enabled (training) the same modules run through torch's fp32 library path so that autograd can trace them.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib


class SiLU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


class GroupNorm32(nn.GroupNorm):
    """GroupNorm evaluated in fp32 (reference model/nn_utils.py:39-41)."""

    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


def normalization(channels):
    return GroupNorm32(min(channels, 32), channels)


def zero_module(m):
    for p in m.parameters():
        p.detach().zero_()
    return m


_FREQS = {}


def timestep_embedding(timesteps, dim, max_period=10000):
    """[cos(t w_k), sin(t w_k)], w_k = max_period^(-k/half) (reference model/nn_utils.py:130-148)."""
    half = dim // 2
    key = (half, max_period, str(timesteps.device))
    freqs = _FREQS.get(key)
    if freqs is None:  # computed on the host exactly like the reference, copied once (no H2D per call: CUDA-graph safe)
        freqs = _FREQS[key] = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(
            timesteps.device)
    ang = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)
    return torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1) if dim % 2 else emb


class TimestepBlock(nn.Module):
    """Marker: forward(x, emb)."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    def forward(self, x, emb):
        for layer in self:
            x = layer(x, emb) if isinstance(layer, TimestepBlock) else layer(x)
        return x


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, odd_size=False):
        super().__init__()
        self.channels, self.use_conv, self.dims, self.odd_size = channels, use_conv, dims, odd_size
        if use_conv:
            self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if self.use_conv:
            x = self.conv(x)
        return x[..., :-1, :-1] if self.odd_size else x


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        self.channels = channels
        self.op = nn.Conv2d(channels, channels, 3, stride=2, padding=1) if use_conv else nn.AvgPool2d(2)

    def forward(self, x):
        return self.op(x)


class ResBlock(TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        co = self.out_channels
        self.in_layers = nn.Sequential(normalization(channels), SiLU(), nn.Conv2d(channels, co, 3, padding=1))
        self.emb_layers = nn.Sequential(SiLU(), nn.Linear(emb_channels, co))
        self.out_layers = nn.Sequential(normalization(co), SiLU(), nn.Dropout(p=dropout),
                                        zero_module(nn.Conv2d(co, co, 3, padding=1)))
        self.skip_connection = nn.Identity() if co == channels else nn.Conv2d(channels, co, 1)

    def forward(self, x, emb):
        h = self.in_layers(x) + self.emb_layers(emb)[:, :, None, None]
        return self.skip_connection(x) + self.out_layers(h)


class QKVAttention(nn.Module):
    def forward(self, qkv):
        ch = qkv.shape[1] // 3
        q, k, v = torch.split(qkv, ch, dim=1)
        scale = 1 / math.sqrt(math.sqrt(ch))
        w = torch.softmax(torch.einsum("bct,bcs->bts", q * scale, k * scale).float(), dim=-1).type(qkv.dtype)
        return torch.einsum("bts,bcs->bct", w, v)


class AttentionBlock(nn.Module):
    def __init__(self, channels, num_heads=1):
        super().__init__()
        self.channels, self.num_heads = channels, num_heads
        self.norm = normalization(channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.attention = QKVAttention()
        self.proj_out = zero_module(nn.Conv1d(channels, channels, 1))

    def forward(self, x):
        b, c, *spatial = x.shape
        x = x.reshape(b, c, -1)
        qkv = self.qkv(self.norm(x))
        h = self.attention(qkv.reshape(b * self.num_heads, -1, qkv.shape[2])).reshape(b, -1, qkv.shape[2])
        return (x + self.proj_out(h)).reshape(b, c, *spatial)


class UNetModel(nn.Module):
    def __init__(self, in_channels, model_channels, out_channels, in_space, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
                 num_heads=1, num_heads_upsample=-1, use_scale_shift_norm=False, learn_potential=False):
        super().__init__()
        if dims != 2 or num_classes is not None or use_scale_shift_norm or learn_potential:
            raise NotImplementedError("only the configuration the reference driver uses is built "
                                      "(dims=2, unconditional, additive embedding, score output)")
        heads_up = num_heads if num_heads_upsample == -1 else num_heads_upsample
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions = num_res_blocks, attention_resolutions
        self.dropout, self.channel_mult, self.conv_resample = dropout, channel_mult, conv_resample
        self.num_classes, self.use_checkpoint, self.num_heads = None, False, num_heads
        self.num_heads_upsample, self.learn_potential = heads_up, False
        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), SiLU(), nn.Linear(ted, ted))
        sizes = [in_space]
        for _ in channel_mult:
            sizes.append(sizes[-1] // 2)
        ch = model_channels * channel_mult[0]
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, ch, 3, padding=1))])
        skip_chans, ds = [ch], 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, ted, dropout, out_channels=mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, num_heads=num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample)))
                skip_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(ResBlock(ch, ted, dropout), AttentionBlock(ch, num_heads=num_heads),
                                                    ResBlock(ch, ted, dropout))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [ResBlock(ch + skip_chans.pop(), ted, dropout, out_channels=model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, num_heads=heads_up))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, odd_size=sizes[level] % 2))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), SiLU(),
                                 zero_module(nn.Conv2d(model_channels * channel_mult[0], out_channels, 3, padding=1)))

    def embedding(self, timesteps):
        return self.time_embed(timestep_embedding(timesteps, self.model_channels))

    def run_blocks(self, x, emb):
        skips, h = [], x
        for blk in self.input_blocks:
            h = blk(h, emb)
            skips.append(h)
        h = self.middle_block(h, emb)
        for blk in self.output_blocks:
            h = blk(torch.cat([h, skips.pop()], dim=1), emb)
        return self.out(h)

    def forward(self, x, timesteps, y=None):
        return self.run_blocks(x, self.embedding(timesteps))

    # ---- hand-written kernel path (inference) -----------------------------------------------------------------------
    @torch.no_grad()
    def embedding_kernels(self, timesteps, log_norm=None, scale_embed=None):
        dev = timesteps.device
        h, L = _lib.ctx(dev), _lib.lib()
        B, E = timesteps.shape[0], self.time_embed[2].weight.shape[0]
        emb = torch.empty((B, E), device=dev, dtype=torch.float32)
        for mlp, val, acc in ((self.time_embed, timesteps, 0), (scale_embed, log_norm, 1)):
            if mlp is None:
                continue
            l1, l2 = mlp[0], mlp[2]
            _lib.check(L.msgm_sincos_embed_mlp(h, _lib.ptr(_lib.f32c(val.reshape(-1), dev)), _lib.ptr(_lib.f32c(l1.weight, dev)),
                                               _lib.ptr(_lib.f32c(l1.bias, dev)), _lib.ptr(_lib.f32c(l2.weight, dev)),
                                               _lib.ptr(_lib.f32c(l2.bias, dev)), _lib.ptr(emb), B, self.model_channels, E, acc,
                                               _lib.stream_ptr(dev)))
        return emb

    def _k_conv(self, dev, conv, x1, x2=None, gn=None, silu=True, ebias=None, res=None, up=1):
        """conv over [x1, x2] with optional fused GroupNorm(+SiLU) prologue, embedding term and residual."""
        h, L = _lib.ctx(dev), _lib.lib()
        B, C1, Hs, Ws = x1.shape
        C2 = 0 if x2 is None else x2.shape[1]
        W = _lib.f32c(conv.weight, dev)
        Cout, K = W.shape[0], W.shape[-1]
        bias = None if conv.bias is None else _lib.f32c(conv.bias, dev)
        stride = conv.stride[0]
        if self.conv_mode in ("tc", "tc16") and _tc_shape_ok(Cout, C1, C2, K, stride, Hs * up, Ws * up):
            return self._k_conv_tc(dev, conv, W, bias, x1, x2, gn, silu, ebias, res, up, stride)
        stats = gamma = beta = None
        G = 0
        if gn is not None:
            G = gn.num_groups
            stats = torch.empty((B, G, 2), device=dev, dtype=torch.float32)
            _lib.check(L.msgm_gn_stats(h, _lib.ptr(x1), C1, _lib.ptr(x2), C2, Hs * Ws, G, B, _lib.ptr(stats),
                                       _lib.stream_ptr(dev)))
            gamma, beta = _lib.f32c(gn.weight, dev), _lib.f32c(gn.bias, dev)
        pad = 1 if K == 3 else 0
        Ho, Wo = (Hs * up + 2 * pad - K) // stride + 1, (Ws * up + 2 * pad - K) // stride + 1
        out = torch.empty((B, Cout, Ho, Wo), device=dev, dtype=torch.float32)
        p = lambda t_: None if t_ is None else t_.data_ptr()  # noqa: E731
        d = _lib.Conv2dDesc(p(x1), p(x2), p(W), p(bias), p(ebias), p(res), p(stats), p(gamma), p(beta), p(out), B, C1, C2,
                            Cout, K, stride, up, Hs, Ws, G, 0 if gn is None else (2 if silu else 1))
        _lib.check(L.msgm_conv2d(h, C.byref(d), _lib.stream_ptr(dev)))
        return out

    # "tc": tcgen05 implicit GEMM with split fp16 x3 operands (fp32-level parity) where the shape allows; "tc16": the same
    # with ONE fp16 product (3x fewer MMAs, ~1e-3 relative: sampling only); "fp32": CUDA-core kernels only
    conv_mode = "tc"
    fuse_attention_proj = True  # AttentionBlock: attention, proj_out and the residual in one launch (attention_tc.cu, step 4)

    def _k_conv_tc(self, dev, conv, W, bias, x1, x2, gn, silu, ebias, res, up, stride):
        """Tensor-core conv (csrc/conv2d_tc.cu): packed weights cached per weight version, GroupNorm folded to scale/shift."""
        h, L = _lib.ctx(dev), _lib.lib()
        B, C1, Hs, Ws = x1.shape
        C2 = 0 if x2 is None else x2.shape[1]
        Cout, K = W.shape[0], W.shape[-1]
        cache = self.__dict__.setdefault("_tc_wimg", {})
        key = (W.data_ptr(), conv.weight._version, _lib.weight_epoch(), tuple(W.shape), dev.index)
        ent = cache.get(W.data_ptr())
        if ent is None or ent[0] != key:
            nbytes = L.msgm_conv2d_tc_pack_bytes(Cout, C1 + C2, K)
            img = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            _lib.check(L.msgm_conv2d_tc_pack(h, _lib.ptr(W), Cout, C1 + C2, K, _lib.ptr(img), _lib.stream_ptr(dev)))
            ent = cache[W.data_ptr()] = (key, img)
        ss = None
        if gn is not None:
            ss = torch.empty((B, C1 + C2, 2), device=dev, dtype=torch.float32)
            _lib.check(L.msgm_gn_scale_shift(h, _lib.ptr(x1), C1, _lib.ptr(x2), C2, Hs * Ws, gn.num_groups, B,
                                             _lib.ptr(_lib.f32c(gn.weight, dev)), _lib.ptr(_lib.f32c(gn.bias, dev)),
                                             _lib.ptr(ss), _lib.stream_ptr(dev)))
        pad = K // 2
        Ho, Wo = (Hs * up + 2 * pad - K) // stride + 1, (Ws * up + 2 * pad - K) // stride + 1
        out = torch.empty((B, Cout, Ho, Wo), device=dev, dtype=torch.float32)
        p = lambda t_: None if t_ is None else t_.data_ptr()  # noqa: E731
        d = _lib.Conv2dTcDesc(p(x1), p(x2), p(ent[1]), p(bias), p(ebias), p(res), p(ss), p(out), B, C1, C2, Cout, K, stride,
                              up, Hs, Ws, 0 if gn is None else (2 if silu else 1), int(self.conv_mode == "tc16"))
        _lib.check(L.msgm_conv2d_tc(h, C.byref(d), _lib.stream_ptr(dev)))
        return out

    def _k_all_emb_proj(self, dev, emb):
        """Linear(SiLU(emb)) of every ResBlock in one launch (stacked weights cached per weight version)."""
        h, L = _lib.ctx(dev), _lib.lib()
        blocks = [m for m in self.modules() if isinstance(m, ResBlock)]
        if not blocks or len(blocks) > 64 or emb.shape[1] > 8192:
            return None
        lins = [b_.emb_layers[1] for b_ in blocks]
        ver = tuple((l.weight._version, l.bias._version, l.weight.data_ptr()) for l in lins) + (dev.index, _lib.weight_epoch())
        ent = self.__dict__.get("_emb_stack")
        if ent is None or ent[0] != ver:
            Wc = torch.cat([_lib.f32c(l.weight, dev) for l in lins], 0).contiguous()
            bc = torch.cat([_lib.f32c(l.bias, dev) for l in lins], 0).contiguous()
            starts = [0]
            for l in lins:
                starts.append(starts[-1] + l.weight.shape[0])
            ent = self.__dict__["_emb_stack"] = (ver, Wc, bc, starts, (C.c_int32 * len(starts))(*starts))
        _, Wc, bc, starts, cstarts = ent
        B, E = emb.shape
        out = torch.empty(B * starts[-1], device=dev, dtype=torch.float32)
        _lib.check(L.msgm_emb_proj_multi(h, _lib.ptr(emb), _lib.ptr(Wc), _lib.ptr(bc), _lib.ptr(out), E, starts[-1], B,
                                         len(lins), cstarts, _lib.stream_ptr(dev)))
        return {id(b_): out[B * starts[i]:B * starts[i + 1]].view(B, -1) for i, b_ in enumerate(blocks)}

    def _k_resblock(self, dev, blk, x1, x2, emb):
        h, L = _lib.ctx(dev), _lib.lib()
        B, E = emb.shape
        lin = blk.emb_layers[1]
        pre = self.__dict__.get("_emb_cur")
        if pre is not None and id(blk) in pre:
            eb = pre[id(blk)]
        else:
            eb = torch.empty((B, blk.out_channels), device=dev, dtype=torch.float32)
            _lib.check(L.msgm_emb_proj(h, _lib.ptr(emb), _lib.ptr(_lib.f32c(lin.weight, dev)),
                                       _lib.ptr(_lib.f32c(lin.bias, dev)), _lib.ptr(eb), E, blk.out_channels, B,
                                       _lib.stream_ptr(dev)))
        # While a CUDA graph is being captured, the 1x1 skip conv of a channel-changing block (it reads the block input only)
        # goes on a side stream: in the replayed graph it is a parallel branch beside GroupNorm -> conv1 -> GroupNorm, joined
        # in front of conv2, which adds it as the residual.  Eager launches stay on one stream.
        has_skip = not isinstance(blk.skip_connection, nn.Identity)
        fork = has_skip and self.graph_branches and torch.cuda.is_current_stream_capturing()
        skip = x1  # channels unchanged: the block input is a single tensor
        if fork:
            main, side = torch.cuda.current_stream(dev), self._side_stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                skip = self._k_conv(dev, blk.skip_connection, x1, x2)
        h1 = self._k_conv(dev, blk.in_layers[2], x1, x2, gn=blk.in_layers[0], ebias=eb)
        if fork:
            main.wait_stream(side)
        elif has_skip:
            skip = self._k_conv(dev, blk.skip_connection, x1, x2)
        return self._k_conv(dev, blk.out_layers[3], h1, gn=blk.out_layers[0], res=skip)

    graph_branches = True  # independent kernels of a block as parallel branches of a captured graph (see _k_resblock)

    def _side_stream(self, dev):
        streams = self.__dict__.setdefault("_side_streams", {})
        if dev.index not in streams:
            streams[dev.index] = torch.cuda.Stream(device=dev)
        return streams[dev.index]

    def _k_attention(self, dev, blk, x):
        h, L = _lib.ctx(dev), _lib.lib()
        B, Cc, Hh, Ww = x.shape
        qkv = self._k_conv(dev, _as2d(blk.qkv), x, gn=blk.norm, silu=False)
        if (self.fuse_attention_proj and self.conv_mode in ("tc", "tc16") and
                L.msgm_attention_proj_tc_supported(Cc, Hh * Ww)):
            # attention + output projection + residual in one launch (packed 1x1 weights cached like every conv's)
            proj = _as2d(blk.proj_out)
            W = _lib.f32c(proj.weight, dev)
            cache = self.__dict__.setdefault("_tc_wimg", {})
            key = (W.data_ptr(), proj.weight._version, _lib.weight_epoch(), tuple(W.shape), dev.index)
            ent = cache.get(W.data_ptr())
            if ent is None or ent[0] != key:
                img = torch.empty(L.msgm_conv2d_tc_pack_bytes(Cc, Cc, 1), device=dev, dtype=torch.uint8)
                _lib.check(L.msgm_conv2d_tc_pack(h, _lib.ptr(W), Cc, Cc, 1, _lib.ptr(img), _lib.stream_ptr(dev)))
                ent = cache[W.data_ptr()] = (key, img)
            out = torch.empty((B, Cc, Hh, Ww), device=dev, dtype=torch.float32)
            bias = None if proj.bias is None else _lib.f32c(proj.bias, dev)
            _lib.check(L.msgm_attention_proj_tc(h, _lib.ptr(qkv), _lib.ptr(ent[1]), _lib.ptr(bias), _lib.ptr(x), _lib.ptr(out), B, Cc,
                                                Hh * Ww, _lib.stream_ptr(dev)))
            return out
        att = torch.empty((B, Cc, Hh, Ww), device=dev, dtype=torch.float32)
        fn = L.msgm_attention_tc if self.conv_mode in ("tc", "tc16") and L.msgm_attention_tc_supported(Cc, Hh * Ww) else L.msgm_attention
        _lib.check(fn(h, _lib.ptr(qkv), _lib.ptr(att), B, Cc, Hh * Ww, _lib.stream_ptr(dev)))
        return self._k_conv(dev, _as2d(blk.proj_out), att, res=x)

    def _k_sequential(self, dev, seq, x1, x2, emb):
        cur, second = x1, x2
        for layer in seq:
            if isinstance(layer, ResBlock):
                cur = self._k_resblock(dev, layer, cur, second, emb)
            elif isinstance(layer, AttentionBlock):
                cur = self._k_attention(dev, layer, cur)
            elif isinstance(layer, Downsample):
                cur = self._k_conv(dev, layer.op, cur)
            elif isinstance(layer, Upsample):
                cur = self._k_conv(dev, layer.conv, cur, up=2)
                if layer.odd_size:
                    cur = cur[..., :-1, :-1].contiguous()
            elif isinstance(layer, nn.Conv2d):
                cur = self._k_conv(dev, layer, cur, second)
            else:
                raise NotImplementedError(type(layer).__name__)
            second = None
        return cur

    @torch.no_grad()
    def run_blocks_kernels(self, x, emb):
        dev = x.device
        if self.num_heads != 1 or not self.conv_resample:
            raise NotImplementedError("kernel path is built for the driver's configuration (1 head, learned resampling)")
        skips, cur = [], _lib.f32c(x, dev)
        self.__dict__["_emb_cur"] = self._k_all_emb_proj(dev, emb)
        for blk in self.input_blocks:
            cur = self._k_sequential(dev, blk, cur, None, emb)
            skips.append(cur)
        cur = self._k_sequential(dev, self.middle_block, cur, None, emb)
        for blk in self.output_blocks:
            cur = self._k_sequential(dev, blk, cur, skips.pop(), emb)
        out = self._k_conv(dev, self.out[2], cur, gn=self.out[0])
        self.__dict__["_emb_cur"] = None
        return out


def _tc_shape_ok(Cout, C1, C2, K, stride, Hi, Wi):
    if K not in (1, 3) or Cout % 32 or (C1 + C2) % 16 or C1 % 16:
        return False
    if K == 3 and 128 + 2 * (Wi + 3) > 768:  # one 128-position block plus its halo must fit the stager's item table
        return False
    return stride == 1 or (stride == 2 and K == 3 and Hi % 2 == 0 and Wi % 2 == 0)


class _Conv1x1View:
    """Conv1d(k=1) over (B,C,T) presented as a 1x1 Conv2d to the conv kernel wrapper."""

    def __init__(self, conv1d):
        self.weight = conv1d.weight.unsqueeze(-1)
        self.bias = conv1d.bias
        self.stride = (1, 1)


def _as2d(conv1d):
    return _Conv1x1View(conv1d)
