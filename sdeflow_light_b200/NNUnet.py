"""2-D U-Net score net on flattened square images: drop-in for the reference's NNUnet.py (VorticityUNet,
UNetModelWithLogNorm, flat_to_img / img_to_flat; reference NNUnet.py:19-77,80-142,145-245).

Conventions kept from the reference: inputs are divided by 5 on the way in and outputs multiplied by 5 on the way out;
(B, H*W) vectors are reshaped in C or Fortran order; with premodule="NormalizeLogRadius" the image is x/(|x|+eps)*sqrt(d)
and the embedding is time_embed(sin-emb(t)) + scale_embed(sin-emb(log(|x|+eps))).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .model.unet import SiLU, UNetModel, timestep_embedding
from .NN import NormalizeLogRadius, evaluate  # noqa: F401

scale_image = 5


def set_library_precision(tf32: bool):
    """Precision of torch's library convolutions / matmuls on the autograd (training) path of the U-Nets.

    The backward and double-backward convolutions run when the USER calls ``loss.backward()``, outside any context this
    module could open, and torch's default lets cuDNN use TF32 there (10-bit mantissa: gradients 3e-4 off an fp32
    run).  The reference is fp32 end to end, so the training forward sets the process-wide switches to fp32 -- a
    deliberate, documented side effect -- unless the net opts in with ``train_tf32 = True`` (what the reference itself
    would silently get on a GPU)."""
    tf32 = bool(tf32)
    if torch.backends.cudnn.allow_tf32 != tf32:
        torch.backends.cudnn.allow_tf32 = tf32
    if torch.backends.cuda.matmul.allow_tf32 != tf32:
        torch.backends.cuda.matmul.allow_tf32 = tf32


def flat_to_img(x, H, W, order="C"):
    B, d = x.shape
    assert d == H * W, f"Expected d={H*W}, got {d}"
    x = x / scale_image
    return x.view(B, 1, H, W) if order == "C" else x.view(B, 1, W, H).transpose(2, 3).contiguous()


def img_to_flat(y, order="C"):
    B, C, H, W = y.shape
    assert C == 1, f"Expected 1 channel, got {C}"
    y = scale_image * y
    return y.reshape(B, H * W) if order == "C" else y.transpose(2, 3).contiguous().view(B, H * W)


class UNetModelWithLogNorm(UNetModel):
    def __init__(self, *args, use_log_norm: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        self.use_log_norm = use_log_norm
        if use_log_norm:
            ted = self.model_channels * 4
            self.scale_embed = nn.Sequential(nn.Linear(self.model_channels, ted), SiLU(), nn.Linear(ted, ted))

    def forward_kernels(self, x, timesteps, log_norm=None):
        emb = self.embedding_kernels(timesteps, log_norm, self.scale_embed if self.use_log_norm else None)
        return self.run_blocks_kernels(x, emb)

    def forward(self, x, timesteps, y=None, log_norm: Optional[torch.Tensor] = None):
        emb = self.embedding(timesteps)
        if self.use_log_norm:
            assert log_norm is not None, "log_norm must be provided when use_log_norm=True"
            emb = emb + self.scale_embed(timestep_embedding(log_norm.view(-1), self.model_channels))
        return self.run_blocks(x, emb)


class VorticityUNet(nn.Module):
    def __init__(self, base_channels: int = 32, channel_mults=(1, 2, 4), num_res_blocks: int = 2,
                 emb_dim_ignored: int = 128, dropout: float = 0.0, premodule: Optional[str] = None, in_space: int = 16,
                 attention_resolutions=(2, 4), conv_resample: bool = True, num_heads: int = 1,
                 use_checkpoint: bool = False, learn_potential: bool = False, flatten_order="C"):
        super().__init__()
        assert premodule in (None, "NormalizeLogRadius") and flatten_order in ("C", "F")
        self.pre = NormalizeLogRadius() if premodule == "NormalizeLogRadius" else None
        self.in_space, self.flatten_order = int(in_space), flatten_order
        self.core = UNetModelWithLogNorm(
            in_channels=1, model_channels=base_channels, out_channels=1, in_space=int(in_space),
            num_res_blocks=num_res_blocks, attention_resolutions=attention_resolutions, dropout=dropout,
            channel_mult=tuple(channel_mults), conv_resample=conv_resample, dims=2, num_classes=None,
            use_checkpoint=False, num_heads=num_heads, use_scale_shift_norm=False, learn_potential=learn_potential,
            use_log_norm=(premodule == "NormalizeLogRadius"))

    def forward(self, x, t):
        if not x.is_cuda:
            raise RuntimeError("sdeflow_light_b200.NNUnet.VorticityUNet runs on CUDA only (no CPU fallback)")
        needs_graph = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not needs_graph and x.dim() == 2:
            return self._forward_kernels(x, t.view(-1))
        # autograd (training) path through torch's library layers
        set_library_precision(self.train_tf32)
        return self._forward(x, t.view(-1))

    train_tf32 = False  # TF32 tensor cores for the library convs of the autograd (training) path; off = fp32 parity
    cuda_graph = True  # replay the ~150 launches of one forward as one CUDA graph per (batch size, weight version)
    max_batch = 1024   # larger batches are evaluated in chunks of this many samples (one graph per chunk size)

    @torch.no_grad()
    def _forward_kernels(self, x, t):
        """Inference on the hand-written kernels: wrapper (normalise, x sqrt(d), / 5, reshape) -> U-Net -> x 5, flatten.

        The launch sequence is static for a given batch size, so it is captured once and replayed as a CUDA graph
        (inputs copied into the graph's buffers, result cloned out); any in-place weight update or re-allocation
        changes the version key and triggers a re-capture (which also re-packs the tensor-core weight images)."""
        from . import _lib
        dev = x.device
        if not torch.cuda.is_current_stream_capturing():
            _lib.check_async(dev)  # an earlier tensor-core launch that gave up surfaces here (no synchronisation)
        B = x.shape[0]
        xs = _lib.f32c(x, dev)
        tt = _lib.f32c(t, dev)
        if tt.numel() == 1 and B != 1:
            tt = tt.expand(B).contiguous()
        if B > self.max_batch:  # bound the activation working set (29.5 MB of fp32 activations per 32x32 sample)
            return torch.cat([self._forward_kernels(xs[i:i + self.max_batch], tt[i:i + self.max_batch])
                              for i in range(0, B, self.max_batch)], 0)
        if not self.cuda_graph or B == 0 or torch.cuda.is_current_stream_capturing():
            return self._forward_kernels_eager(xs, tt)
        plist = self.__dict__.get("_plist")
        if plist is None:
            plist = self.__dict__["_plist"] = list(self.parameters())
        ver = hash(tuple((p_._version, p_.data_ptr()) for p_ in plist) + (_lib.weight_epoch(),))
        cache = self.__dict__.setdefault("_graphs", {})
        key = (B, dev.index, self.core.conv_mode)
        ent = cache.get(key)
        if ent is None or ent[0] != ver:
            sx, st = xs.clone(), tt.clone()
            self._forward_kernels_eager(sx, st)  # warm-up outside the capture: packs weights, sets kernel attributes
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                sout = self._forward_kernels_eager(sx, st)
            if len(cache) >= 4:
                cache.pop(next(iter(cache)))
            ent = cache[key] = (ver, graph, sx, st, sout)
        _, graph, sx, st, sout = ent
        sx.copy_(xs)
        st.copy_(tt)
        graph.replay()
        return sout.clone()

    def _forward_kernels_eager(self, xs, tt):
        from . import _lib
        dev = xs.device
        h, L = _lib.ctx(dev), _lib.lib()
        B, d = xs.shape
        S = self.in_space
        assert d == S * S, f"Flat dim {d} != {S}*{S}"
        pre = self.pre is not None
        img = torch.empty((B, 1, S, S), device=dev, dtype=torch.float32)
        logn = torch.empty(B, device=dev, dtype=torch.float32) if pre else None
        forder = int(self.flatten_order == "F")
        _lib.check(L.msgm_vort_pre(h, _lib.ptr(xs), _lib.ptr(img), _lib.ptr(logn), B, S, S, forder, int(pre),
                                   _lib.stream_ptr(dev)))
        y_img = self.core.forward_kernels(img, tt, logn)
        out = torch.empty((B, d), device=dev, dtype=torch.float32)
        _lib.check(L.msgm_vort_post(h, _lib.ptr(y_img), _lib.ptr(out), B, S, S, forder, _lib.stream_ptr(dev)))
        return out

    def _forward(self, x, t):
        log_norm = None
        if self.pre is not None:
            x, log_norm = self.pre(x)
            x = x * float(torch.sqrt(torch.tensor(float(x.shape[-1]))))  # fp32 sqrt(d) as in the reference, without a device copy (CUDA-graph safe)
        flat = x.dim() == 2
        if flat:
            img = flat_to_img(x, self.in_space, self.in_space, order=self.flatten_order)
        elif x.dim() == 4 and x.size(1) == 1:
            img = x
        else:
            raise ValueError(f"Unexpected input shape {tuple(x.shape)}")
        out = self.core(img, timesteps=t, log_norm=log_norm) if self.pre is not None else self.core(img, timesteps=t)
        return img_to_flat(out, order=self.flatten_order) if flat else out
