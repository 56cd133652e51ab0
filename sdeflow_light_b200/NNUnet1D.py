"""1-D U-Net score net: drop-in for the reference's NNUnet1D.py (same constructor, parameter names, forward contract).

Topology (reference NNUnet1D.py:27-108): a 2-layer time MLP (1 -> emb -> emb, exact GELU) [+ an identical MLP of
log|x| when premodule="NormalizeLogRadius"]; three encoder blocks (conv3-GELU-conv3-GELU) each followed by a stride-2
k=4 convolution; a middle block; three decoder stages (transposed k=4 s=2 conv, concat skip, block); a 1x1 projection.
The embedding vector is concatenated as ``emb_dim`` extra channels in front of EVERY block.  No normalisation layers.

Inference (no autograd: the sampling hot path) runs entirely on hand-written kernels (csrc/conv1d_fp32.cu): the
embedding MLPs, the premodule, every Conv1d / ConvTranspose1d with fused exact-GELU epilogue.  The embedding channels,
constant along the signal, are folded into a per-(sample, out-channel, tap) table instead of being concatenated, and the
decoder's skip concatenation is read from two tensors in place.  With autograd enabled (training) the same layers run
as torch modules so that autograd can trace them.
"""
from __future__ import annotations

from typing import Optional

import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .NN import NormalizeLogRadius, evaluate  # noqa: F401


def _embed_mlp(emb_dim):
    return nn.Sequential(nn.Linear(1, emb_dim), nn.GELU(), nn.Linear(emb_dim, emb_dim))


class ConvBlock1D(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.net = nn.Sequential(nn.Conv1d(in_ch, out_ch, kernel_size=3, padding=1), nn.GELU(),
                                 nn.Conv1d(out_ch, out_ch, kernel_size=3, padding=1), nn.GELU())

    def forward(self, x):
        return self.net(x)


class UNet1D(nn.Module):
    def __init__(self, input_dim, base_channels=32, channel_mults=(1, 2, 4), num_res_blocks=2,
                 premodule: Optional[str] = None, emb_dim=128):
        super().__init__()
        assert premodule in (None, "NormalizeLogRadius")
        self.input_dim = input_dim
        self.premodule = NormalizeLogRadius() if premodule == "NormalizeLogRadius" else None
        self.time_mlp = _embed_mlp(emb_dim)
        self.scale_embed = _embed_mlp(emb_dim) if self.premodule is not None else None
        widths = [base_channels * m for m in channel_mults]
        self.enc_blocks, self.downs = nn.ModuleList(), nn.ModuleList()
        c_in = 1
        for w in widths:
            self.enc_blocks.append(ConvBlock1D(c_in + emb_dim, w))
            self.downs.append(nn.Conv1d(w, w, kernel_size=4, stride=2, padding=1))
            c_in = w
        self.middle = ConvBlock1D(c_in + emb_dim, c_in)
        self.up_convs, self.dec_blocks = nn.ModuleList(), nn.ModuleList()
        for w in reversed(widths):
            self.up_convs.append(nn.ConvTranspose1d(c_in, w, kernel_size=4, stride=2, padding=1))
            self.dec_blocks.append(ConvBlock1D(2 * w + emb_dim, w))
            c_in = w
        self.final = nn.Conv1d(c_in, 1, kernel_size=1)

    def forward(self, x, t):
        if not x.is_cuda:
            raise RuntimeError("sdeflow_light_b200.NNUnet1D.UNet1D runs on CUDA only (no CPU fallback)")
        needs_graph = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not needs_graph:
            return self._forward_kernels(x, t)
        # autograd path: the reference is fp32 end to end, keep cuDNN from silently using TF32 for the convolutions of the
        # forward AND of the backward / double backward the caller triggers later (process-wide switch, see NNUnet)
        from .NNUnet import set_library_precision
        set_library_precision(self.train_tf32)
        return self._forward(x, t)

    # ---- hand-written kernel path (inference) -------------------------------------------------------------------
    # "tc": convs on tcgen05 with split fp16 x3 operands (fp32-level parity); "tc16": one fp16 product (~1e-3 relative,
    # sampling only); "fp32": CUDA-core kernels only.  `planes` (with "tc" / "tc16", nets whose widths allow it): the
    # activations between the convs stay in the tensor cores' operand format and the convs are TMA-fed, persistent,
    # epilogue-overlapped kernels (csrc/conv1d_tcp.cu) -- same arithmetic as "tc", no fp32 staging.
    conv_mode = "tc"
    planes = True
    cuda_graph = True  # replay one forward as one CUDA graph per (batch size, weight version), as NNUnet does
    max_batch = 4096  # larger batches are evaluated in chunks of this many samples
    train_tf32 = False  # TF32 for the library convs of the autograd (training) path; off = fp32 parity with the reference

    def _conv(self, h, dev, conv, x1, x2=None, emb=None, gelu=False):
        L = _lib.lib()
        B, C1, Lin = x1.shape
        C2 = 0 if x2 is None else x2.shape[1]
        W, bias = _lib.f32c(conv.weight, dev), _lib.f32c(conv.bias, dev)
        Cout, Cw, K = W.shape
        stride, pad = conv.stride[0], conv.padding[0]
        Lout = (Lin + 2 * pad - K) // stride + 1
        Cemb = Cw - C1 - C2
        E = None
        if Cemb > 0:
            E = torch.empty((B, Cout, K), device=dev, dtype=torch.float32)
            _lib.check(L.msgm_emb_fold(h, _lib.ptr(W), _lib.ptr(emb), _lib.ptr(E), Cw, C1 + C2, Cemb, Cout, K, B,
                                       _lib.stream_ptr(dev)))
        out = torch.empty((B, Cout, Lout), device=dev, dtype=torch.float32)
        Cin = C1 + C2
        if (self.conv_mode in ("tc", "tc16") and Cout % 32 == 0 and Cin % 16 == 0 and C1 % 16 == 0 and Cin > 0 and
                ((K == 3 and stride == 1 and pad == 1) or (K == 4 and stride == 2 and pad == 1 and Lin >= 2) or
                 (K == 1 and stride == 1 and pad == 0))):
            cache = self.__dict__.setdefault("_tc_wimg", {})
            key = (conv.weight._version, _lib.weight_epoch(), tuple(W.shape), dev.index)
            ent = cache.get(W.data_ptr())
            if ent is None or ent[0] != key:
                img = torch.empty(L.msgm_conv1d_tc_pack_bytes(Cout, Cin, K), device=dev, dtype=torch.uint8)
                _lib.check(L.msgm_conv1d_tc_pack(h, _lib.ptr(W), Cout, Cw, Cin, K, _lib.ptr(img), _lib.stream_ptr(dev)))
                ent = cache[W.data_ptr()] = (key, img)
            d = _lib.Conv1dTcDesc(x1.data_ptr(), None if x2 is None else x2.data_ptr(), ent[1].data_ptr(), bias.data_ptr(),
                                  None if E is None else E.data_ptr(), out.data_ptr(), B, C1, C2, Cout, K, stride, Lin,
                                  int(gelu), int(self.conv_mode == "tc16"))
            _lib.check(L.msgm_conv1d_tc(h, C.byref(d), _lib.stream_ptr(dev)))
            return out
        d = _lib.Conv1dDesc(x1.data_ptr(), None if x2 is None else x2.data_ptr(), W.data_ptr(), bias.data_ptr(),
                            None if E is None else E.data_ptr(), out.data_ptr(), B, C1, C2, Cemb, Cout, K, stride, pad,
                            Lin, Lout, int(gelu))
        _lib.check(L.msgm_conv1d(h, C.byref(d), _lib.stream_ptr(dev)))
        return out

    def _block(self, h, dev, block, x1, x2, emb):
        y = self._conv(h, dev, block.net[0], x1, x2, emb, gelu=True)
        return self._conv(h, dev, block.net[2], y, gelu=True)

    def _embed(self, h, dev, mlp, t, out, accumulate):
        L = _lib.lib()
        l1, l2 = mlp[0], mlp[2]
        _lib.check(L.msgm_embed_mlp(h, _lib.ptr(t), _lib.ptr(_lib.f32c(l1.weight, dev)), _lib.ptr(_lib.f32c(l1.bias, dev)),
                                    _lib.ptr(_lib.f32c(l2.weight, dev)), _lib.ptr(_lib.f32c(l2.bias, dev)), _lib.ptr(out),
                                    t.shape[0], l2.weight.shape[0], int(accumulate), _lib.stream_ptr(dev)))

    @torch.no_grad()
    def _forward_kernels(self, x, t):
        dev = x.device
        h, L = _lib.ctx(dev), _lib.lib()
        if not torch.cuda.is_current_stream_capturing():
            _lib.check_async(dev)  # an earlier tensor-core launch that gave up surfaces here (no synchronisation)
        xs = _lib.f32c(x.reshape(x.shape[0], -1), dev)
        B, Lsig = xs.shape
        tt = _lib.f32c(t.reshape(-1), dev)
        if tt.numel() == 1 and B != 1:
            tt = tt.expand(B).contiguous()
        if B > self.max_batch:  # bound the activation working set (~10 MB of fp32 activations per L = 1000 sample)
            return torch.cat([self._forward_kernels(xs[i:i + self.max_batch], tt[i:i + self.max_batch])
                              for i in range(0, B, self.max_batch)], 0)
        if B > 0 and self.cuda_graph and not torch.cuda.is_current_stream_capturing():
            return self._forward_graphed(xs, tt)
        return self._forward_kernels_eager(xs, tt)

    def _forward_graphed(self, xs, tt):
        """The launch sequence of one forward is static for a batch size: capture it once, replay it afterwards (inputs
        copied into the graph's buffers, result cloned out).  Any in-place weight update, re-allocation or graph-replayed
        optimiser step (weight epoch) changes the version key and triggers a re-capture."""
        dev = xs.device
        B = xs.shape[0]
        plist = self.__dict__.get("_plist")
        if plist is None:
            plist = self.__dict__["_plist"] = list(self.parameters())
        ver = hash(tuple((p_._version, p_.data_ptr()) for p_ in plist) + (_lib.weight_epoch(),))
        cache = self.__dict__.setdefault("_graphs", {})
        key = (B, xs.shape[1], dev.index, self.conv_mode, self.planes)
        ent = cache.get(key)
        if ent is None or ent[0] != ver:
            sx, st = xs.clone(), tt.clone()
            self._forward_kernels_eager(sx, st)  # warm-up outside the capture: packs weights, allocates plane buffers
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                sout = self._forward_kernels_eager(sx, st)
            if len(cache) >= 4:
                cache.pop(next(iter(cache)))
            # the entry also pins the plane buffers the captured kernels write to: they may be dropped from `_plane_bufs`
            # (which keeps the most recent batch sizes only) while this graph is still replayed
            pinned = self.__dict__.get("_plane_bufs", {}).get((B, xs.shape[1], dev.index))
            ent = cache[key] = (ver, graph, sx, st, sout, pinned)
        _, graph, sx, st, sout, _ = ent
        sx.copy_(xs)
        st.copy_(tt)
        graph.replay()
        return sout.clone()

    @torch.no_grad()
    def _forward_kernels_eager(self, xs, tt):
        dev = xs.device
        h, L = _lib.ctx(dev), _lib.lib()
        B, Lsig = xs.shape
        E = self.time_mlp[2].weight.shape[0]
        emb = torch.empty((B, E), device=dev, dtype=torch.float32)
        logn = None
        if self.premodule is not None:
            xn, logn = torch.empty_like(xs), torch.empty(B, device=dev, dtype=torch.float32)
            _lib.check(L.msgm_normalize_log_radius(h, _lib.ptr(xs), _lib.ptr(xn), _lib.ptr(logn), B, Lsig,
                                                   _lib.stream_ptr(dev)))
            xs = xn
        if E <= 224:  # both embedding MLPs in one launch (bit-identical to the two-launch form below)
            wa = [_lib.f32c(p_, dev) for p_ in (self.time_mlp[0].weight, self.time_mlp[0].bias, self.time_mlp[2].weight,
                                                self.time_mlp[2].bias)]
            wb = [None] * 4 if logn is None else [_lib.f32c(p_, dev) for p_ in (
                self.scale_embed[0].weight, self.scale_embed[0].bias, self.scale_embed[2].weight, self.scale_embed[2].bias)]
            _lib.check(L.msgm_embed_mlp2(h, _lib.ptr(tt), *[_lib.ptr(w) for w in wa], _lib.ptr(logn), *[_lib.ptr(w) for w in wb],
                                         _lib.ptr(emb), B, E, _lib.stream_ptr(dev)))
        else:
            self._embed(h, dev, self.time_mlp, tt, emb, False)
            if logn is not None:
                self._embed(h, dev, self.scale_embed, logn, emb, True)
        if B > 0 and self.planes and self.conv_mode in ("tc", "tc16") and self._planes_ok(Lsig):
            return self._forward_planes(h, dev, xs, emb)
        cur, skips = xs.view(B, 1, Lsig), []
        for block, down in zip(self.enc_blocks, self.downs):
            cur = self._block(h, dev, block, cur, None, emb)
            skips.append(cur)
            cur = self._conv(h, dev, down, cur)
        cur = self._block(h, dev, self.middle, cur, None, emb)
        for up, block in zip(self.up_convs, self.dec_blocks):
            skip = skips.pop()
            Bc, Cin, Lin = cur.shape
            W, bias = _lib.f32c(up.weight, dev), _lib.f32c(up.bias, dev)
            Cup, Lup = W.shape[1], skip.shape[-1]
            if self.conv_mode in ("tc", "tc16") and Cin % 16 == 0 and Cup % 16 == 0 and Lup >= 2 * Lin:
                cache = self.__dict__.setdefault("_tc_wimg", {})
                key = (up.weight._version, _lib.weight_epoch(), tuple(W.shape), dev.index)
                ent = cache.get(W.data_ptr())
                if ent is None or ent[0] != key:
                    img = torch.empty(24 * Cin * Cup, device=dev, dtype=torch.uint8)
                    _lib.check(L.msgm_convt1d_tc_pack(h, _lib.ptr(W), Cup, Cin, _lib.ptr(img), _lib.stream_ptr(dev)))
                    ent = cache[W.data_ptr()] = (key, img)
                alloc = torch.zeros if Lup > 2 * Lin else torch.empty
                upo = alloc((Bc, Cup, Lup), device=dev, dtype=torch.float32)
                _lib.check(L.msgm_convt1d_tc(h, _lib.ptr(cur), _lib.ptr(ent[1]), _lib.ptr(bias), _lib.ptr(upo), Bc, Cin, Cup,
                                             Lin, Lup, int(self.conv_mode == "tc16"), _lib.stream_ptr(dev)))
            else:
                upo = torch.empty((Bc, Cup, Lup), device=dev, dtype=torch.float32)
                _lib.check(L.msgm_convt1d_k4s2(h, _lib.ptr(cur), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(upo), Bc, Cin, Cup,
                                               Lin, Lup, _lib.stream_ptr(dev)))
            cur = self._block(h, dev, block, upo, skip, emb)
        out = self._conv(h, dev, self.final, cur)
        return out.squeeze(1)

    # ---- planes path: activations in operand format between TMA-fed tensor-core convs (csrc/conv1d_tcp.cu) ----------------
    def _planes_ok(self, Lsig) -> bool:
        """Every conv between the first and the last one has Cin % 16 == 0 (both halves of a concat) and Cout % 32 == 0."""
        ws = [b.net[2].weight.shape[0] for b in self.enc_blocks]
        first = self.enc_blocks[0].net[0]
        return (all(w % 32 == 0 for w in ws) and ws[0] <= 128 and Lsig >= 2 ** len(ws) and
                first.weight.shape[1] - self.time_mlp[2].weight.shape[0] == 1 and self.final.weight.shape[0] == 1)

    def _plane_buf(self, bufs, name, B, Cc, Lc, dev):
        """Zero-filled ONCE: the kernels only write rows of real positions, padding and guard rows stay zero for good."""
        t = bufs.get(name)
        if t is None:
            t = bufs[name] = (torch.zeros(_lib.lib().msgm_planes_bytes(B, Cc, Lc), device=dev, dtype=torch.uint8), Cc, Lc)
        assert t[1] == Cc and t[2] == Lc
        return t[0]

    def _wimg(self, h, dev, weight, pack):
        cache = self.__dict__.setdefault("_tc_wimg", {})
        W = _lib.f32c(weight, dev)
        key = (weight._version, _lib.weight_epoch(), tuple(W.shape), dev.index)
        ent = cache.get(W.data_ptr())
        if ent is None or ent[0] != key:
            ent = cache[W.data_ptr()] = (key, pack(W))
        return ent[1]

    def _conv_p(self, h, dev, conv, B, x1, C1, x2, C2, Lin, E, gelu, outp=None, outf=None, transposed=False, Lout=0):
        """E: the conv's folded embedding table (B, Cout, K) from _emb_tables, or None."""
        L = _lib.lib()
        W, bias = _lib.f32c(conv.weight, dev), _lib.f32c(conv.bias, dev)
        Cin = C1 + C2
        if transposed:
            Cout, K = W.shape[1], 3

            def pack(Wc):
                img = torch.empty(24 * Cin * Cout, device=dev, dtype=torch.uint8)
                _lib.check(L.msgm_convt1d_tc_pack(h, _lib.ptr(Wc), Cout, Cin, _lib.ptr(img), _lib.stream_ptr(dev)))
                return img
        else:
            Cout, Cw, K = W.shape
            assert (E is not None) == (Cw > Cin)

            def pack(Wc):
                img = torch.empty(L.msgm_conv1d_tc_pack_bytes(Cout, Cin, K), device=dev, dtype=torch.uint8)
                _lib.check(L.msgm_conv1d_tc_pack(h, _lib.ptr(Wc), Cout, Cw, Cin, K, _lib.ptr(img), _lib.stream_ptr(dev)))
                return img
        img = self._wimg(h, dev, conv.weight, pack)
        d = _lib.Conv1dTcpDesc(x1.data_ptr(), None if x2 is None else x2.data_ptr(), img.data_ptr(), bias.data_ptr(),
                               None if E is None else E.data_ptr(), None if outp is None else outp.data_ptr(),
                               None if outf is None else outf.data_ptr(), B, C1, C2, Cout, K, Lin, Lout, int(gelu),
                               int(transposed), int(self.conv_mode == "tc16"))
        _lib.check(L.msgm_conv1d_tcp(h, C.byref(d), _lib.stream_ptr(dev)))
        return Cout

    def _emb_tables(self, h, dev, emb, convs):
        """Folded embedding tables E (B, Cout, K) of every block's first conv in one launch per 16 convs: the embedding
        channels are the LAST Cemb input channels of those convs (NNUnet1D.py:156-175) and constant along the signal."""
        L = _lib.lib()
        B, Cemb = emb.shape
        out = {}
        for g0 in range(0, len(convs), 16):
            D = _lib.EmbFoldMultiDesc()
            grp = convs[g0:g0 + 16]
            for i, conv in enumerate(grp):
                W = _lib.f32c(conv.weight, dev)
                Cout, Cw, K = W.shape
                E = out[conv] = torch.empty((B, Cout, K), device=dev, dtype=torch.float32)
                D.W[i], D.E[i], D.Cw[i], D.Coff[i], D.Cout[i], D.K[i] = W.data_ptr(), E.data_ptr(), Cw, Cw - Cemb, Cout, K
            D.n, D.Cemb, D.B, D.emb = len(grp), Cemb, B, emb.data_ptr()
            _lib.check(L.msgm_emb_fold_multi(h, C.byref(D), _lib.stream_ptr(dev)))
        return out

    def _forward_planes(self, h, dev, xs, emb):
        L = _lib.lib()
        B, Lsig = xs.shape
        Et = self._emb_tables(h, dev, emb, [b_.net[0] for b_ in self.enc_blocks] + [self.middle.net[0]] +
                              [b_.net[0] for b_ in self.dec_blocks])
        # one set of zero-initialised plane buffers per (batch size, length, device), the 8 most recent kept.  A CUDA graph
        # captured by a CALLER around this forward must keep using the same batch size among those, or call the net once
        # eagerly before capturing (the captured kernels write into these buffers); the net's own graphs pin theirs.
        allb = self.__dict__.setdefault("_plane_bufs", {})
        bkey = (B, Lsig, dev.index)
        if bkey not in allb and len(allb) >= 8 and not torch.cuda.is_current_stream_capturing():
            allb.pop(next(iter(allb)))
        bufs = allb.setdefault(bkey, {})
        buf = lambda name, Cc, Lc: self._plane_buf(bufs, name, B, Cc, Lc, dev)  # noqa: E731
        # first conv: one real channel + folded embedding -> planes
        first = self.enc_blocks[0].net[0]
        W0, b0 = _lib.f32c(first.weight, dev), _lib.f32c(first.bias, dev)
        C0, Cw0, _ = W0.shape
        E0 = Et[first]
        cur, Cc, Lc = buf("e0a", C0, Lsig), C0, Lsig
        _lib.check(L.msgm_conv1d_first_planes(h, _lib.ptr(xs), _lib.ptr(W0), Cw0, _lib.ptr(b0), _lib.ptr(E0), _lib.ptr(cur), B, C0,
                                              Lsig, 1, _lib.stream_ptr(dev)))
        skips = []
        for i, (block, down) in enumerate(zip(self.enc_blocks, self.downs)):
            if i > 0:
                Cn = block.net[0].weight.shape[0]
                nxt = buf(f"e{i}a", Cn, Lc)
                self._conv_p(h, dev, block.net[0], B, cur, Cc, None, 0, Lc, Et[block.net[0]], True, outp=nxt)
                cur, Cc = nxt, Cn
            nxt = buf(f"e{i}b", Cc, Lc)
            self._conv_p(h, dev, block.net[2], B, cur, Cc, None, 0, Lc, None, True, outp=nxt)
            skips.append((nxt, Cc, Lc))
            Ld = (Lc + 2 - 4) // 2 + 1
            dn = buf(f"d{i}", Cc, Ld)
            self._conv_p(h, dev, down, B, nxt, Cc, None, 0, Lc, None, False, outp=dn)
            cur, Lc = dn, Ld
        ma, mb = buf("ma", Cc, Lc), buf("mb", Cc, Lc)
        self._conv_p(h, dev, self.middle.net[0], B, cur, Cc, None, 0, Lc, Et[self.middle.net[0]], True, outp=ma)
        self._conv_p(h, dev, self.middle.net[2], B, ma, Cc, None, 0, Lc, None, True, outp=mb)
        cur = mb
        outf = None
        nd = len(self.dec_blocks)
        for i, (up, block) in enumerate(zip(self.up_convs, self.dec_blocks)):
            skip, Cs, Ls = skips.pop()
            Cup = up.weight.shape[1]
            upo = buf(f"u{i}", Cup, Ls)
            self._conv_p(h, dev, up, B, cur, Cc, None, 0, Lc, None, False, outp=upo, transposed=True, Lout=Ls)
            Cn = block.net[0].weight.shape[0]
            da = buf(f"x{i}a", Cn, Ls)
            self._conv_p(h, dev, block.net[0], B, upo, Cup, skip, Cs, Ls, Et[block.net[0]], True, outp=da)
            if i + 1 < nd:
                db = buf(f"x{i}b", Cn, Ls)
                self._conv_p(h, dev, block.net[2], B, da, Cn, None, 0, Ls, None, True, outp=db)
                cur = db
            else:  # the last block feeds the 1x1 projection (one output channel, HBM-bound CUDA-core kernel): fp32 NCL
                outf = torch.empty((B, Cn, Ls), device=dev, dtype=torch.float32)
                self._conv_p(h, dev, block.net[2], B, da, Cn, None, 0, Ls, None, True, outf=outf)
            Cc, Lc = Cn, Ls
        out = self._conv(h, dev, self.final, outf)
        return out.squeeze(1)

    def _forward(self, x, t):
        h = x.unsqueeze(1) if x.ndim == 2 else x
        emb = self.time_mlp(t.view(-1, 1))
        if self.premodule is not None:
            h, log_norm = self.premodule(h)
            h = h * float(torch.sqrt(torch.tensor(float(h.shape[-1]))))  # fp32 sqrt(d) as in the reference, without a device copy (CUDA-graph safe)
            emb = emb + self.scale_embed(log_norm.view(log_norm.shape[0], -1).to(emb.dtype))
        emb = emb.unsqueeze(-1)

        def with_emb(*feats):
            return torch.cat([*feats, emb.expand(-1, -1, feats[0].shape[-1])], dim=1)

        skips = []
        for block, down in zip(self.enc_blocks, self.downs):
            h = block(with_emb(h))
            skips.append(h)
            h = down(h)
        h = self.middle(with_emb(h))
        for up, block in zip(self.up_convs, self.dec_blocks):
            h, skip = up(h), skips.pop()
            if h.shape[-1] != skip.shape[-1]:
                h = F.pad(h, (0, skip.shape[-1] - h.shape[-1]))
            h = block(with_emb(h, skip))
        return self.final(h).squeeze(1)
