"""Integrators: drop-in for the reference's sde_scheme.py (same names, arguments, return shapes, errors).

Each call is ONE persistent kernel launch that runs all ``num_steps`` steps on the GPU (msgm_sample_mlp) and ONE
device->host copy of the result, instead of the reference's Python loop with ~100 ATen launches and a blocking
D2H copy per step (sde_scheme.py:223-262).  Extra keyword-only arguments (absent from the reference):

``noise``      (num_steps,B,d) standard normals to use instead of in-kernel Philox (parity tests).
``precision``  "fp32" (default, CUDA-core parity mode) or "f16tc" (tcgen05 tensor cores).
``seed``/``particle_offset``  Philox key; by default the seed is drawn from torch's global generator so that
               ``torch.manual_seed`` controls reproducibility; particle_offset makes sharded runs draw the
               noise of the global particle index.
``device_out`` return the result on the GPU instead of the reference's CPU tensor.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

__all__ = ["EMstep", "euler_maruyama_sampler", "heun_sampler", "rk4_stratonovich_sampler"]


@torch.no_grad()
def EMstep(mu, delta, sigma, dW, sparse=False, I=None, K=None):
    """mu*delta + sigma.dW for the three sigma layouts of the reference (sde_scheme.py:18-40).

    Kept for API compatibility; the fused samplers never materialise sigma and do not call this.
    """
    if sparse:
        dx = torch.zeros_like(dW)
        dx.scatter_add_(1, I.unsqueeze(0).expand(dW.size(0), -1), sigma * dW[:, K])
    elif sigma.dim() > 2:
        dx = torch.einsum("bij,bj->bi", sigma, dW)
    else:
        dx = sigma * dW
    return mu * delta + dx


def _describe(sde):
    """Map a sampler-protocol object onto (base_sde, net or None, forward_only)."""
    from . import SDEs
    if isinstance(sde, SDEs.forward_SDE):
        return sde.base_sde, None, True
    if isinstance(sde, SDEs.PluginReverseSDE):
        return sde.base_sde, sde.a, False
    raise TypeError(f"unsupported sde object {type(sde).__name__}: expected forward_SDE or PluginReverseSDE")


def _run(scheme, sde, x_0, num_steps, lmbd, keep_all_samples, samplesToKeep, include_t0, T_, norm_correction,
         noise, precision, seed, particle_offset, device_out, T_rows=None):
    from . import NN
    base, net, fwd = _describe(sde)
    device = sde.T.device
    handle = _lib.ctx(device)  # RuntimeError on CPU: there is no CPU fallback
    _lib.check_async(device)   # a tensor-core launch that gave up earlier surfaces here (no synchronisation)
    B, d = x_0.size(0), x_0.size(1)
    T_run = _lib.host_float(sde, "T") if (not torch.is_tensor(T_) and T_ == -1) else T_.item()
    if keep_all_samples is False and samplesToKeep is not None and len(samplesToKeep) != B:
        raise ValueError("Error: len(samplesToKeep) must correspond to batch size.")
    if (fwd and d > 32) or (not fwd and not (isinstance(net, NN.MLP) and net.fused_ok())):
        from . import generic_sampler
        return generic_sampler.run(scheme, sde, x_0, num_steps, lmbd, keep_all_samples, samplesToKeep, include_t0,
                                   T_run, norm_correction, noise, seed, particle_offset, device_out)

    x = _lib.f32c(x_0, device).clone()
    if B == 0:  # nothing to launch; shapes as the reference would return them
        n_out = num_steps + (1 if include_t0 else 0)
        out = x.new_zeros((n_out, 0, d)) if keep_all_samples else x
        return out if device_out else out.to("cpu")
    sd, keep_alive = base.desc(device)
    sd.dim = d  # SGMsde, like the reference's, has no `dim` attribute: the state width is the batch's
    a = _lib.SampleArgs()
    a.scheme, a.num_steps, a.lmbd = scheme, int(num_steps), float(lmbd)
    a.norm_correction, a.include_t0, a.forward_only = int(bool(norm_correction)), int(bool(include_t0)), int(fwd)
    a.precision = {"fp32": _lib.PREC_FP32, "f16tc": _lib.PREC_F16TC}[precision]
    a.T_ = float(T_run)
    if T_rows is None:
        ts = (torch.linspace(0, 1, num_steps + 1) * T_run).to(device)  # the reference's fp32 grid (:201)
    else:
        ts = torch.linspace(0, 1, num_steps + 1).to(device)  # unit grid, scaled per row in the kernel
        T_rows = _lib.f32c(T_rows.reshape(-1), device)
        a.T_rows = T_rows.data_ptr()
    a.ts = ts.data_ptr()
    if noise is not None:
        noise = _lib.f32c(noise, device)
        if tuple(noise.shape) != (num_steps, B, d):
            raise ValueError(f"noise must have shape {(num_steps, B, d)}")
        a.noise = noise.data_ptr()
    else:
        a.seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
        a.particle_offset = int(particle_offset)
    traj = keep_out = keep_step = None
    if keep_all_samples:
        traj = torch.empty((num_steps + (1 if include_t0 else 0), B, d), device=device, dtype=torch.float32)
        a.traj = traj.data_ptr()
    elif samplesToKeep is not None:
        keep_step = torch.as_tensor(samplesToKeep).reshape(-1).to(device=device, dtype=torch.int32).contiguous()
        keep_out = torch.zeros((B, d), device=device, dtype=torch.float32)
        a.keep_step, a.keep_out = keep_step.data_ptr(), keep_out.data_ptr()
    md = None
    if not fwd:
        mdesc, k2 = net.desc(device)
        keep_alive += k2
        md = C.byref(mdesc)
    _lib.check(_lib.lib().msgm_sample_mlp(handle, C.byref(sd), md, C.byref(a), _lib.ptr(x), B,
                                          _lib.stream_ptr(device)))
    out = traj if keep_all_samples else (keep_out if samplesToKeep is not None else x)
    if device_out:
        return out
    out = out.to("cpu")  # reference samplers always return CPU tensors (:99,172,269)
    _lib.check_async(device)  # the copy synchronised: this call's own kernel has reported by now
    return out


@torch.no_grad()
def _run_rows(sde, y0, t_rows, noise=None, seed=None):
    """One RK4 step of size t_rows[k] for every row k: the batched form of the reference's per-row calls
    ``rk4_stratonovich_sampler(forward_SDE, y0[k][None], 1, T_=t[k])`` (SDEs.py:114-116).  Returns a device tensor."""
    if y0.shape[1] > 32:  # stage-kernel path takes one horizon per launch: few rows, one launch group each
        out = torch.empty_like(y0)
        for k in range(y0.shape[0]):
            out[k:k + 1] = _run(_lib.SCHEME_RK4, sde, y0[k:k + 1], 1, 0., False, None, False, t_rows[k:k + 1], False,
                                None if noise is None else noise[:, k:k + 1], "fp32", seed, k, True)
        return out
    return _run(_lib.SCHEME_RK4, sde, y0, 1, 0., False, None, False, -1, False, noise, "fp32", seed, 0, True,
                T_rows=t_rows)


@torch.no_grad()
def euler_maruyama_sampler(sde, x_0, num_steps=1000, lmbd=0., keep_all_samples=True, samplesToKeep=None,
                           include_t0=False, T_=-1, norm_correction=False, *, noise=None, precision="fp32",
                           seed=None, particle_offset=0, device_out=False):
    """Ito Euler-Maruyama (sde_scheme.py:43-99)."""
    return _run(_lib.SCHEME_EM, sde, x_0, num_steps, lmbd, keep_all_samples, samplesToKeep, include_t0, T_,
                norm_correction, noise, precision, seed, particle_offset, device_out)


@torch.no_grad()
def heun_sampler(sde, x_0, num_steps=1000, lmbd=0., keep_all_samples=True, samplesToKeep=None,
                 include_t0=False, T_=-1, norm_correction=False, *, noise=None, precision="fp32", seed=None,
                 particle_offset=0, device_out=False):
    """Stratonovich Heun / RK2 (sde_scheme.py:101-172)."""
    return _run(_lib.SCHEME_HEUN, sde, x_0, num_steps, lmbd, keep_all_samples, samplesToKeep, include_t0, T_,
                norm_correction, noise, precision, seed, particle_offset, device_out)


@torch.no_grad()
def rk4_stratonovich_sampler(sde, x_0, num_steps=1000, lmbd=0., keep_all_samples=True, samplesToKeep=None,
                             include_t0=False, T_=-1, norm_correction=False, *, noise=None, precision="fp32",
                             seed=None, particle_offset=0, device_out=False):
    """Stratonovich RK4 with one shared Wiener increment per step (sde_scheme.py:174-269)."""
    return _run(_lib.SCHEME_RK4, sde, x_0, num_steps, lmbd, keep_all_samples, samplesToKeep, include_t0, T_,
                norm_correction, noise, precision, seed, particle_offset, device_out)
