"""Sampler loop for score nets that are not fused into the persistent kernels (U-Nets, d up to 4096).

The net is called once per Runge-Kutta stage (4 per RK4 step, sde_scheme.py:230-250); everything else of the step --
the reference's three g() evaluations, two f() evaluations, gather/scatter_add with atomics, the RK bookkeeping and
the radius re-pin (sde_scheme.py:18-40,223-255; SDEs.py:556-588) -- is ONE hand-written kernel per stage
(msgm_stage_update), plus in-kernel Philox noise and a row-norm kernel.  State, noise and trajectory stay on the GPU;
there is one device->host copy at the end instead of one per step.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_NSTAGE = {_lib.SCHEME_EM: 1, _lib.SCHEME_HEUN: 2, _lib.SCHEME_RK4: 4}

# One CUDA graph per STEP: calls with at least GRAPH_MIN_STEPS steps capture the launches of one whole step once (noise draw,
# nstage net evaluations = ~130 kernels each for the U-Nets, nstage stage updates) and replay that graph num_steps times; the
# step index and the stage times are read from device memory (msgm_step_clock), so the graph is step-independent.  Measured
# on a B200 (tools/unet_sampling_bench.py, profiles/unet_sampling_bench_r02.txt): the eager loop is already GPU-bound (the
# per-stage time IS the net forward, 0.5 ms at batch 16 and 0.98 ms at batch 256 for the 1-D U-Net), so the graph changes
# throughput by less than the ~12 ms its capture costs per call; what it buys is a host thread that issues one launch per step
# instead of ~520.  Hence the threshold: only long calls (the drivers' N = 1000) take this path.
STEP_GRAPH = True
GRAPH_MIN_STEPS = 256


def _run_step_graph(handle, L, stream, device, sd, scheme, nstage, net, fwd, lmbd, norm_correction, delta, stage_time, num_steps,
                    noise, seed, particle_offset, include_t0, x, y, ks, dW, r0, s_vec, traj, keep_step, keep_out):
    f32 = np.float32
    B, d = x.shape
    s_tab = np.zeros(num_steps * nstage + 1, dtype=np.float32)
    for i in range(num_steps):
        for st in range(nstage):
            s_tab[i * nstage + st] = stage_time(i, st)
    s_dev = torch.from_numpy(s_tab).to(device)
    clock = torch.zeros(1, device=device, dtype=torch.int32)
    keep32 = None if keep_step is None else keep_step.to(torch.int32).contiguous()
    clk = _lib.StepClock(clock.data_ptr(), s_dev.data_ptr(), s_vec.data_ptr(), None if traj is None else traj.data_ptr(),
                         None if keep32 is None else keep32.data_ptr(), None if keep_out is None else keep_out.data_ptr(),
                         int(bool(include_t0)), 0)

    def step_body():
        stream = _lib.stream_ptr(device)  # the capture runs on torch's capture stream, not on the caller's
        _lib.check(L.msgm_philox_normal_clocked(handle, _lib.ptr(dW), d, B, float(f32(delta ** 0.5)), int(seed or 0),
                                                int(particle_offset), C.byref(clk), nstage, _lib.ptr(noise), stream))
        for st in range(nstage):
            a = None
            if not fwd:
                a = _lib.f32c(net(x if st == 0 else y, s_vec), device).reshape(B, d)
            _lib.check(L.msgm_stage_update_clocked(handle, C.byref(sd), scheme, st, float(lmbd), int(bool(norm_correction)),
                                                   int(fwd), C.byref(clk), float(f32(delta)), _lib.ptr(a), _lib.ptr(dW),
                                                   _lib.ptr(r0), _lib.ptr(x), _lib.ptr(y), _lib.ptr(ks), B, stream))
        _lib.check(L.msgm_clock_advance(handle, _lib.ptr(clock), stream))

    # the first step runs eagerly: it is step 0 of the result AND the warm-up that packs weight images / sets kernel
    # attributes before the capture
    step_body()
    torch.cuda.synchronize(device)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step_body()
    for _ in range(1, num_steps):
        graph.replay()


def run(scheme, sde, x_0, num_steps, lmbd, keep_all_samples, samplesToKeep, include_t0, T_run, norm_correction, noise,
        seed, particle_offset, device_out):
    from .sde_scheme import _describe
    base, net, fwd = _describe(sde)
    device = sde.T.device
    handle, L = _lib.ctx(device), _lib.lib()
    _lib.check_async(device)
    stream = _lib.stream_ptr(device)
    B, d = x_0.size(0), x_0.size(1)
    sd, keep_alive = base.desc(device)
    if sd.kind == _lib.SDE_MSGM_DENSE:
        raise NotImplementedError("dense G with a non-fused score net is not built (O(d^3) per particle); "
                                  "use denseTensor=False as the reference does for d >= 256")
    sd.dim = d
    x = _lib.f32c(x_0, device).clone()
    y, ks, dW = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    r0 = None
    if norm_correction:
        r0 = torch.empty(B, device=device, dtype=torch.float32)
        _lib.check(L.msgm_row_norm(handle, _lib.ptr(x), _lib.ptr(r0), d, B, stream))
    # time grid and step sizes exactly as the reference rounds them (sde_scheme.py:200-201,224,236,248)
    delta = T_run / num_steps
    ts = (torch.linspace(0, 1, num_steps + 1) * T_run).numpy().astype(np.float32)
    f32 = np.float32
    T_sde = f32(sde.T.item())
    nstage = _NSTAGE[scheme]
    if noise is not None:
        noise = _lib.f32c(noise, device)
        if tuple(noise.shape) != (num_steps, B, d):
            raise ValueError(f"noise must have shape {(num_steps, B, d)}")
    elif seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    traj = keep_out = keep_step = None
    if keep_all_samples:
        traj = torch.empty((num_steps + (1 if include_t0 else 0), B, d), device=device, dtype=torch.float32)
        if include_t0:
            traj[0].copy_(x)
    elif samplesToKeep is not None:
        keep_step = torch.as_tensor(samplesToKeep).reshape(-1).to(device)
        keep_out = torch.zeros((B, d), device=device, dtype=torch.float32)
    s_vec = torch.empty(B, device=device, dtype=torch.float32)

    def stage_time(i, st):
        t_s = ts[i]
        if st > 0:
            t_s = f32(t_s + f32(delta / 2)) if (nstage == 4 and st < 3) else f32(t_s + f32(delta))
        return float(t_s) if fwd else float(f32(T_sde - t_s))

    if STEP_GRAPH and num_steps >= GRAPH_MIN_STEPS and not torch.cuda.is_current_stream_capturing():
        _run_step_graph(handle, L, stream, device, sd, scheme, nstage, net, fwd, lmbd, norm_correction, delta, stage_time,
                        num_steps, noise, seed, particle_offset, include_t0, x, y, ks, dW, r0, s_vec, traj, keep_step, keep_out)
    else:
        for i in range(num_steps):
            if noise is not None:
                torch.mul(noise[i], float(f32(delta ** 0.5)), out=dW)
            else:
                _lib.check(L.msgm_philox_normal(handle, _lib.ptr(dW), d, B, float(f32(delta ** 0.5)), int(seed),
                                                int(particle_offset), i, stream))
            for st in range(nstage):
                s = stage_time(i, st)
                a = None
                if not fwd:
                    s_vec.fill_(s)
                    a = _lib.f32c(net(x if st == 0 else y, s_vec), device).reshape(B, d)
                _lib.check(L.msgm_stage_update(handle, C.byref(sd), scheme, st, float(lmbd), int(bool(norm_correction)),
                                               int(fwd), s, float(f32(delta)), _lib.ptr(a), _lib.ptr(dW), _lib.ptr(r0),
                                               _lib.ptr(x), _lib.ptr(y), _lib.ptr(ks), B, stream))
            if traj is not None:
                traj[i + (1 if include_t0 else 0)].copy_(x)
            elif keep_step is not None:
                m = keep_step == (i + (1 if include_t0 else 0))
                torch.where(m[:, None], x, keep_out, out=keep_out)  # no nonzero(): no host sync inside the loop
    out = traj if keep_all_samples else (keep_out if samplesToKeep is not None else x)
    if device_out:
        return out
    out = out.to("cpu")
    _lib.check_async(device)  # the copy synchronised: every conv / attention launch of this call has reported by now
    return out
