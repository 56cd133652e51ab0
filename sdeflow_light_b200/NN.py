"""Score networks: drop-in for the reference's NN.py (same class names, constructor arguments, parameter names).

``MLP.forward`` without autograd runs the hand-written sm_100a forward kernel (msgm_mlp_forward).  Inside the
samplers and the SSM train step the net is never called through ``forward``: its weights are handed to the fused
kernels.  With autograd enabled (someone differentiating through the net outside the fused train step) the
forward is expressed in torch ops on the GPU so that autograd can trace it -- that path is not on the hot path.
"""
from __future__ import annotations

import ctypes as C
import random

import numpy as np
import torch
import torch.nn as nn

from . import _lib


_SDE_TENSORS = ("G", "L_G", "r_T", "G_I", "G_J", "G_K", "G_V")


def save_checkpoint(path, gen_sde, optim, iteration, *, trainer=None, persist_sde=True):
    """Same dictionary layout as the reference (NN.py:13-22) so checkpoints load both ways.

    Two extra keys the reference's loader ignores (it reads its six keys by name): ``msgm_sde`` holds the base SDE's
    tensors that ``state_dict()`` does not -- the random skew-symmetric ``G`` (drawn from the global RNG at construction,
    SDEs.py:315-321), ``L_G``, the radius table ``r_T`` or the sparse ``G_I/G_J/G_K/G_V`` -- so that a resumed multiplicative
    SDE is well defined; ``msgm_trainer_rng`` holds the Philox seed / iteration counter of a
    ``train.GraphedSsmStep`` so that a resumed run continues its random stream."""
    ck = {"iteration": iteration, "model": gen_sde.state_dict(), "optimizer": optim.state_dict(),
          "torch_rng": torch.get_rng_state().cpu(), "numpy_rng": np.random.get_state(),
          "python_rng": random.getstate()}
    base = getattr(gen_sde, "base_sde", None)
    if persist_sde and base is not None:
        extra = {n: getattr(base, n).detach().cpu() for n in _SDE_TENSORS if torch.is_tensor(getattr(base, n, None))}
        if extra:
            ck["msgm_sde"] = extra
    if trainer is not None:
        ck["msgm_trainer_rng"] = trainer.rng_state()
    torch.save(ck, path)


def load_checkpoint(path, gen_sde, optim, device, *, trainer=None):
    """NN.py:24-42; additionally restores the ``msgm_sde`` / ``msgm_trainer_rng`` extras when the file has them (a file
    written by the reference has neither and loads exactly as in the reference)."""
    ck = torch.load(path, map_location=device, weights_only=False)
    gen_sde.load_state_dict(ck["model"])
    optim.load_state_dict(ck["optimizer"])
    rng = ck["torch_rng"]
    torch.set_rng_state((rng if rng.dtype == torch.uint8 else rng.to(torch.uint8)).cpu())
    np.random.set_state(ck["numpy_rng"])
    random.setstate(ck["python_rng"])
    base = getattr(gen_sde, "base_sde", None)
    if base is not None and "msgm_sde" in ck:
        for n, t in ck["msgm_sde"].items():
            cur = getattr(base, n, None)
            if torch.is_tensor(cur) and cur.shape == t.shape and cur.dtype == t.dtype:
                with torch.no_grad():
                    cur.copy_(t)  # in place: captured CUDA graphs (train.GraphedSsmStep) hold these addresses
            else:
                setattr(base, n, t.to(base.device))
        base.__dict__.pop("_rT_sorted", None)  # derived caches
    if trainer is not None and "msgm_trainer_rng" in ck:
        trainer.load_rng_state(ck["msgm_trainer_rng"])
    print(f"Resuming from iteration {ck['iteration'] + 1}")
    return ck["iteration"]


class Swish(nn.Module):
    """x * sigmoid(x) (NN.py:48-53)."""

    def forward(self, x):
        return torch.sigmoid(x) * x


class NormalizeLogRadius(nn.Module):
    """x -> (x / (|x| + eps), log(|x| + eps)), non-learnable (NN.py:56-70)."""

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        norm = torch.norm(x, dim=-1, keepdim=True) + self.eps
        return x / norm, torch.log(norm)


class MLP(nn.Module):
    """(d [+1] + index_dim) -> h -> h -> h -> d with Swish (NN.py:73-120)."""

    def __init__(self, input_dim=2, index_dim=1, hidden_dim=128, act=None, premodule=None):
        super().__init__()
        act = Swish() if act is None else act
        self.input_dim, self.index_dim, self.hidden_dim, self.act = input_dim, index_dim, hidden_dim, act
        self.output_dim = input_dim
        assert premodule is None or premodule in ["NormalizeLogRadius"]
        self.premodule = premodule
        self.pre = NormalizeLogRadius() if premodule == "NormalizeLogRadius" else None
        self.learnable_network_input_dim = input_dim + (1 if self.pre is not None else 0)
        self.main = nn.Sequential(
            nn.Linear(self.learnable_network_input_dim + index_dim, hidden_dim), act,
            nn.Linear(hidden_dim, hidden_dim), act,
            nn.Linear(hidden_dim, hidden_dim), act,
            nn.Linear(hidden_dim, self.output_dim))

    # ---- what the fused kernels need ---------------------------------------------------------------------
    def fused_ok(self) -> bool:
        """True when the hand-written kernels cover this net (the reference driver's configuration)."""
        return (self.hidden_dim == _lib_hidden() and self.index_dim == 1 and isinstance(self.act, Swish)
                and self.input_dim <= 32)

    def linears(self):
        return [m for m in self.main if isinstance(m, nn.Linear)]

    def desc(self, device):
        """(MlpDesc, keep-alive list) pointing at this module's fp32 device weights."""
        keep, d = [], _lib.MlpDesc()
        d.input_dim, d.premodule = self.input_dim, 1 if self.pre is not None else 0
        for i, l in enumerate(self.linears()):
            w, b = _lib.f32c(l.weight, device), _lib.f32c(l.bias, device)
            keep += [w, b]
            d.W[i], d.b[i] = w.data_ptr(), b.data_ptr()
        return d, keep

    def forward(self, input, t):
        sz = input.size()
        x = input.reshape(-1, self.input_dim)
        t = t.reshape(-1, self.index_dim).float()
        needs_graph = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not needs_graph and self.fused_ok():
            dev = x.device
            handle = _lib.ctx(dev)  # raises on CPU tensors: no CPU fallback
            d, keep = self.desc(dev)
            xc, tc = _lib.f32c(x, dev), _lib.f32c(t.reshape(-1), dev)
            if tc.numel() == 1 and xc.shape[0] != 1:
                tc = tc.expand(xc.shape[0]).contiguous()
            out = torch.empty_like(xc)
            _lib.check(_lib.lib().msgm_mlp_forward(handle, C.byref(d), _lib.ptr(xc), _lib.ptr(tc), _lib.ptr(out),
                                                   xc.shape[0], _lib.stream_ptr(dev)))
            return out.view(*sz)
        if not x.is_cuda:
            raise RuntimeError("sdeflow_light_b200.NN.MLP runs on CUDA only (no CPU fallback)")
        if self.pre is not None:
            h, ln = self.pre(x)
            x = torch.cat([h, ln], dim=-1)
        return self.main(torch.cat([x, t], dim=1)).view(*sz)


def _lib_hidden() -> int:
    return 128


@torch.no_grad()
def evaluate(gen_sde, x_test):
    """ELBO mean / stderr (NN.py:123-128)."""
    gen_sde.eval()
    n = x_test.size(0)
    elbo = gen_sde.elbo_random_t_slice(x_test)
    gen_sde.train()
    return elbo.mean(), elbo.std() / n ** 0.5
