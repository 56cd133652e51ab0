"""1-D U-Net score net: drop-in for the reference's NNUnet1D.py (same constructor, parameter names, forward contract).

Topology (reference NNUnet1D.py:27-108): a 2-layer time MLP (1 -> emb -> emb, exact GELU) [+ an identical MLP of
log|x| when premodule="NormalizeLogRadius"]; three encoder blocks (conv3-GELU-conv3-GELU) each followed by a stride-2
k=4 convolution; a middle block; three decoder stages (transposed k=4 s=2 conv, concat skip, block); a 1x1 projection.
The embedding vector is concatenated as ``emb_dim`` extra channels in front of EVERY block.  No normalisation layers.

Round-1 status: the layer arithmetic runs through torch's convolution library calls on the GPU; the hand-written part
of this path is the per-stage SDE update (csrc/stage_ops.cu).  Hand-written conv kernels are the next step.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

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
        # the reference is fp32 end to end: keep cuDNN from silently using TF32 for the convolutions
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return self._forward(x, t)

    def _forward(self, x, t):
        h = x.unsqueeze(1) if x.ndim == 2 else x
        emb = self.time_mlp(t.view(-1, 1))
        if self.premodule is not None:
            h, log_norm = self.premodule(h)
            h = h * torch.sqrt(torch.tensor(h.shape[-1], dtype=log_norm.dtype, device=log_norm.device))
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
