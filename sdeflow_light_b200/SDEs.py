"""SDE definitions: drop-in for the reference's SDEs.py (same classes, constructor arguments, methods, state_dict).

Hot entry points run hand-written sm_100a kernels through libmsgm_b200.so:

* ``SDE.sample_scheme`` / ``sample_scheme_allt`` / ``MSGMsde.sample``  -> fused forward-noising sampler launch
* ``PluginReverseSDE.ssm``                                             -> fused SSM forward/backward (ssm_fused.py)
* the samplers of ``sde_scheme`` read ``base_sde.desc()`` and never call the coefficient methods below.

The coefficient methods (``beta, f, g, div_Sigma, mu, sigma, ...``) are kept because they are the reference's
public surface (sampler <-> SDE protocol); they are thin tensor expressions evaluated on
whatever device their inputs live on and are not on the hot path of this package.
"""
from __future__ import annotations

import gc

import numpy as np
import torch

from . import _lib
from .sde_scheme import euler_maruyama_sampler, heun_sampler, rk4_stratonovich_sampler  # noqa: F401

Log2PI = float(np.log(2 * np.pi))


class forward_SDE(torch.nn.Module):
    """Sampler-protocol view of a base SDE running forward in time (reference SDEs.py:30-47)."""

    def __init__(self, base_sde, T):
        super().__init__()
        self.base_sde = base_sde
        self.T = T

    def mu(self, s, y, lmbd=0.):
        return self.mu_Strato(s, y) + 0.5 * self.base_sde.div_Sigma(s, y)

    def mu_Strato(self, s, y, lmbd=0.):
        return self.base_sde.f_strato(s, y)

    def sigma(self, s, y, lmbd=0., sparse=False):
        return self.base_sde.g(s, y, sparse=sparse)


class SDE(torch.nn.Module):
    """Common base (reference SDEs.py:49-155)."""

    def __init__(self, beta_min=0.1, beta_max=20.0, T=1.0, t_epsilon=0.001, num_steps_forward=100, device="cpu"):
        super().__init__()
        self.device = torch.device(device)
        self.T = T
        self.beta_min, self.beta_max, self.t_epsilon = beta_min, beta_max, t_epsilon
        self.num_steps_forward = num_steps_forward
        self.norm_correction = False
        self.sparseTensor = False

    def to(self, device):
        new = super().to(device)
        new.device = torch.device(device)
        new.T = self.T.to(device)
        return new

    def beta(self, t):
        return self.beta_min + (self.beta_max - self.beta_min) * t

    def IJK(self):
        return None, None, None

    # ---- descriptor for the C ABI ---------------------------------------------------------------------------
    def _kind(self):
        raise NotImplementedError

    def desc(self, device):
        d = _lib.SdeDesc()
        d.kind, d.dim = self._kind(), int(getattr(self, "dim", 0))
        d.beta_min = float(self.beta_min)
        d.beta_delta = float(self.beta_max - self.beta_min)  # difference in double, as beta() does
        d.T = _lib.host_float(self, "T")
        return d, []

    # ---- forward noising by simulation (reference SDEs.py:78-132) ------------------------------------------------
    fused_noising = True  # large sparse states: one launch for the whole forward noising (False: per-stage kernels)

    @torch.no_grad()
    def sample_scheme(self, t, y0, keep_all_samples, return_noise=False, *, noise=None, noise_rows=None,
                      _single=None):
        """y_t | y_0: the state after trunc(N_fwd t/T) RK4 steps; rows with 0 steps take ONE step of size t.

        ONE launch (msgm_noise_forward) replaces the reference's batch sampler call plus its per-row Python loop of
        one-sample sampler calls, for the dense tensor up to d = 32 and for the sparse tensor up to d = 4096 (the U-Net
        configurations); ``fused_noising = False`` selects the per-stage kernels in two passes instead (N_fwd x 5 launches).  ``noise`` (N_fwd,B,d) and
        ``noise_rows`` (n_zero_step_rows,d) inject the standard normals the reference would have drawn (parity tests).
        (The reference's 'warning : t >= T' print is dropped: it would cost a host synchronisation per call.)
        """
        if return_noise:
            raise NotImplementedError('See the official repository.')
        n_tot = self.num_steps_forward
        dev = self.device
        t = t.to(dev)
        y0 = y0.to(dev)
        d = y0.shape[1]
        if d <= 32 or (getattr(self, "sparseTensor", False) and d <= 4096 and self.fused_noising):
            # one launch: the lanes-per-particle kernel (d <= 32) or the one-CTA-per-row kernel (sparse tensor, d <= 4096)
            import ctypes as C
            handle = _lib.ctx(dev)
            y = _lib.f32c(y0, dev).clone()
            tt = _lib.f32c(t.reshape(-1), dev)
            sd, keep = self.desc(dev)
            sd.dim = d  # SGMsde has no dim of its own
            if noise is not None:
                noise = _lib.f32c(noise, dev)
            single = None if _single is None else _lib.f32c(_single, dev)  # (B,d): row k's draw if it takes 0 steps
            if noise_rows is not None:  # (m,d) draws of the zero-step rows in index order -> (B,d) by row
                n_int = torch.trunc(n_tot * t / self.T.to(dev)).to(torch.int).reshape(-1)
                single = torch.zeros_like(y)
                single[n_int == 0] = _lib.f32c(noise_rows, dev)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            grid = getattr(self, "_fwd_grid", None)  # the reference's fp32 time grid (sde_scheme.py:201), cached on device
            T_host = _lib.host_float(self, "T")
            if grid is None or grid[0] != (n_tot, T_host, str(dev)):
                self._fwd_grid = ((n_tot, T_host, str(dev)), (torch.linspace(0, 1, n_tot + 1) * T_host).to(dev))
            _lib.check(_lib.lib().msgm_noise_forward(handle, C.byref(sd), _lib.ptr(tt), _lib.ptr(y), int(n_tot),
                                                     _lib.ptr(self._fwd_grid[1]), _lib.ptr(noise), _lib.ptr(single), seed,
                                                     0, y.shape[0], _lib.stream_ptr(dev)))
            return y
        # large states (U-Net configs): per-stage kernels, two passes
        n_int = torch.trunc(n_tot * t / self.T.to(dev)).to(torch.int).reshape(-1)
        n_int[(t >= self.T.to(dev)).reshape(-1)] = n_tot
        yt = self.sample_scheme_allt(y0, include_t0=True, keep_all_samples=False, samplesToKeep=n_int,
                                     _device_out=True, _noise=noise)
        small = (n_int == 0).nonzero().reshape(-1)
        if small.numel():
            from . import sde_scheme
            yt[small] = sde_scheme._run_rows(forward_SDE(self, self.T), y0[small], t.reshape(-1)[small],
                                             noise=None if noise_rows is None else noise_rows.reshape(1, -1, y0.shape[1]))
        return yt.to(self.device)

    @torch.no_grad()
    def sample_scheme_allt(self, y0, include_t0=True, keep_all_samples=True, samplesToKeep=None, _device_out=False,
                           _noise=None):
        return rk4_stratonovich_sampler(forward_SDE(self, self.T).to(self.device), y0,
                                        num_steps=self.num_steps_forward, lmbd=0, keep_all_samples=keep_all_samples,
                                        samplesToKeep=samplesToKeep, include_t0=include_t0, device_out=_device_out,
                                        noise=_noise)

    def sample_Song_et_al(self, t, y0, return_noise=False):
        """Closed-form VP marginal (reference SDEs.py:134-146)."""
        mu = self.mean_weight(t) * y0
        std = self.var(t) ** 0.5
        epsilon = torch.randn_like(y0)
        yt = epsilon * std + mu
        return (yt, epsilon, std, self.g(t, yt)) if return_noise else yt

    def sample_debiasing_t(self, shape):
        raise NotImplementedError('See the official repository.')


class SGMsde(SDE):
    """Additive variance-preserving SDE of Song et al. (reference SDEs.py:161-215)."""

    def __init__(self, beta_min=0.1, beta_max=20.0, T=1.0, t_epsilon=0.001, num_steps_forward=100, device='cpu'):
        super().__init__(beta_min, beta_max, T, t_epsilon, num_steps_forward, device)
        self.name_SDE = "SGM"

    def _kind(self):
        return _lib.SDE_SGM

    @property
    def logvar_mean_T(self):
        return torch.zeros(1), torch.zeros(1)

    def mean_weight(self, t):
        return torch.exp(-0.25 * t ** 2 * (self.beta_max - self.beta_min) - 0.5 * t * self.beta_min)

    def var(self, t):
        return 1. - torch.exp(-0.5 * t ** 2 * (self.beta_max - self.beta_min) - t * self.beta_min)

    def f(self, t, y):
        return -0.5 * self.beta(t) * y

    f_strato = f

    def div_Sigma(self, t, y):
        return torch.zeros_like(y)

    def g(self, t, y, sparse=False):
        return torch.ones_like(y) * self.beta(t) ** 0.5

    @torch.no_grad()
    def sample(self, t, y0, return_noise=False):
        return self.sample_Song_et_al(t, y0, return_noise)

    def latent_sample(self, num_samples, n, *, seed=None, particle_offset=0):
        """x_0 ~ N(0, I) (reference SDEs.py:201-203), drawn in-kernel (Philox keyed by the global particle index)."""
        return _latent(self, num_samples, n, None, False, seed, particle_offset)

    def cond_latent_sample(self, t_, T, x):
        return self.sample(torch.ones_like(t_) * T, x)

    def log_latent_pdf(self, yT):
        return self.log_normal(yT, torch.zeros_like(yT), torch.zeros_like(yT))

    def log_normal(self, x, mean, log_var, eps=0.00001):
        return -(x - mean) ** 2 / (2. * torch.exp(log_var) + eps) - log_var / 2. - 0.5 * Log2PI


class MSGMsde(SDE):
    """Multiplicative SDE dY = G(Y) o dB with skew-symmetric slices (reference SDEs.py:221-509)."""

    def __init__(self, y0, beta_min=0.1, beta_max=20.0, T=1.0, t_epsilon=0.001, denseTensor=True,
                 norm_sampler="ecdf", norm_map=None, kernel='gaussian', plot_validate=False,
                 num_steps_forward=100, device='cpu', estim_cst_norm_dens_r_T=True):
        super().__init__(beta_min, beta_max, T, t_epsilon, num_steps_forward, device)
        self.sparseTensor = not denseTensor
        self.norm_correction = True
        self.norm_map, self.norm_sampler, self._kernel = norm_map, norm_sampler, kernel
        r = torch.linalg.norm(y0, dim=1)
        if norm_map == "log":
            r = torch.log(r + 1e-6)
        self._bandwidth = 0.1 * torch.std(r.reshape(-1, 1)).item()
        self._kde = None  # sklearn KDE of the radii, fitted lazily (only the ELBO path needs it)
        self.r_T = r.to(self.device)
        self.dim = y0.shape[1]
        self.name_SDE = "MSGM"
        if denseTensor:
            self.new_G(self.dim)
            self.L_G = 0.5 * torch.einsum('ijk, jmk -> im', self.G, self.G)
        else:
            self.name_SDE += "_sparseTens"
            self.sparse_G(self.dim)
            self.L_G = 0.5 * torch.eye(self.dim, device=self.device)
        if norm_sampler != "ecdf":
            self.name_SDE += norm_sampler + kernel
        if norm_map == "log":
            self.name_SDE += "logNorm"
        self.cst_log_dens = 0
        if estim_cst_norm_dens_r_T or plot_validate:
            # log of the rectangle-rule integral of the KDE over 1000 points spanning the radii (SDEs.py:255-265)
            grid = torch.linspace(float(self.r_T.min()), float(self.r_T.max()), 1000, device=self.r_T.device)
            dens = torch.exp(self._kde_logpdf(grid))
            self.cst_log_dens = torch.log(torch.sum(dens, dim=0) * (grid[1] - grid[0])).to(self.device)
        gc.collect()

    @property
    def kde(self):
        """sklearn KernelDensity object like the reference's attribute (SDEs.py:240), fitted lazily for callers that
        want the estimator itself; the package's own density evaluations run on the GPU (``_kde_logpdf``)."""
        if self._kde is None:
            from sklearn.neighbors import KernelDensity
            self._kde = KernelDensity(kernel=self._kernel, bandwidth=self._bandwidth).fit(
                self.r_T.reshape(-1, 1).detach().cpu())
        return self._kde

    def _kde_logpdf(self, r):
        """log KDE density of the radii at ``r`` (msgm_kde_logpdf: exact Gaussian-kernel sum, one CTA per query) --
        replaces ``kde.score_samples`` (SDEs.py:261,509)."""
        if self._kernel != "gaussian":
            raise NotImplementedError(f"kernel '{self._kernel}': only the reference's default 'gaussian' is built")
        dev = r.device
        rT = _lib.f32c(self.r_T.reshape(-1), dev)
        q = _lib.f32c(r.reshape(-1), dev)
        out = torch.empty_like(q)
        _lib.check(_lib.lib().msgm_kde_logpdf(_lib.ctx(dev), _lib.ptr(rT), rT.numel(), float(self._bandwidth), _lib.ptr(q),
                                              _lib.ptr(out), q.numel(), _lib.stream_ptr(dev)))
        return out

    def _kind(self):
        return _lib.SDE_MSGM_SPARSE if self.sparseTensor else _lib.SDE_MSGM_DENSE

    def desc(self, device):
        d, keep = super().desc(device)
        if not self.sparseTensor:
            G, LG = _lib.f32c(self.G, device), _lib.f32c(self.L_G, device)
            keep += [G, LG]
            d.G, d.L_G = G.data_ptr(), LG.data_ptr()
        return d, keep

    def to(self, device):
        new = super().to(device)
        new.r_T = self.r_T.to(device)
        new.L_G = self.L_G.to(device)
        if self.sparseTensor:
            for n in ("G_I", "G_J", "G_K", "G_V"):
                setattr(new, n, getattr(self, n).to(device))
        else:
            new.G = self.G.to(device)
        return new

    def new_G(self, n):
        """n random skew-symmetric slices, scaled so that tr(L_G) = -n/2 (reference SDEs.py:315-341).

        The slices are drawn on the CPU from torch's global generator in the reference's order, so a seeded run
        builds the same tensor as the reference, then moved to ``self.device``.
        """
        F = torch.stack([torch.randn(n, n) for _ in range(n)], dim=2)
        G = 0.5 * (F - F.transpose(0, 1))
        tr_L = torch.trace(0.5 * torch.einsum('ijk, jmk -> im', G, G))
        self.G = (torch.sqrt(-0.5 * n / tr_L) * G).to(self.device)

    def sparse_G(self, n):
        """COO arrays of the cyclic tensor (reference SDEs.py:369-399); the kernels use the stencil form."""
        k = torch.arange(n)
        kp = (k + 1) % n
        self.G_I = torch.stack([k, kp], 1).reshape(-1).to(self.device)
        self.G_J = torch.stack([kp, k], 1).reshape(-1).to(self.device)
        self.G_K = torch.stack([k, k], 1).reshape(-1).to(self.device)
        c = 0.5 * torch.sqrt(torch.tensor(2, dtype=torch.float32))
        self.G_V = torch.stack([c.expand(n), -c.expand(n)], 1).reshape(-1).to(self.device)
        self.G_sparse_cpu = None

    def IJK(self):
        return (self.G_I, self.G_J, self.G_K) if self.sparseTensor else (None, None, None)

    def f(self, t, y):
        b = self.beta(t)
        return 0.5 * b * y if self.sparseTensor else torch.einsum('ij, bj -> bi', self.L_G, b * y)

    def f_strato(self, t, y):
        return torch.zeros_like(y)

    def div_Sigma(self, t, y):
        return 2 * self.f(t, y)

    def g(self, t, y, sparse=False):
        rb = self.beta(t) ** 0.5
        if sparse:
            return self.G_V.unsqueeze(0) * (rb * y[:, self.G_J])
        return torch.einsum('ijk, bj -> bik', self.G, rb * y)

    def sample(self, t, y0, return_noise=False):
        return self.sample_scheme(t, y0, return_noise=return_noise, keep_all_samples=False).to(self.device)

    def gen_radial_distribution(self, num_samples):
        U = torch.rand(num_samples, device=self.device)
        if self.norm_sampler == "ecdf":
            r_gen = torch.quantile(self.r_T, U).reshape(num_samples, 1)
        else:
            # a draw from the Gaussian KDE = a random radius sample + N(0, h^2), on the device (the reference's
            # kde.sample branch raises NameError at SDEs.py:444; same law, no host round trip)
            if self._kernel != "gaussian":
                raise NotImplementedError(f"kernel '{self._kernel}': only the reference's default 'gaussian' is built")
            idx = torch.randint(0, self.r_T.numel(), (num_samples,), device=self.device)
            r_gen = (self.r_T.reshape(-1)[idx] + self._bandwidth * torch.randn(num_samples, device=self.device)
                     ).reshape(num_samples, 1)
            if self.norm_map != "log":
                r_gen = r_gen.clamp_min(0.)
        if self.norm_map == "log":
            r_gen = torch.exp(r_gen) - 1e-6
        return r_gen

    def latent_sample(self, num_samples, n, *, seed=None, particle_offset=0, U=None, Z=None):
        """x_0 = r s with r from the empirical radius law and s uniform on the sphere (reference SDEs.py:438-493): one
        fused kernel (sorted radius table, in-kernel Philox keyed by the global particle index).  ``U`` / ``Z`` inject the
        reference's uniform / normal draws."""
        if self.norm_sampler != "ecdf":
            return self.gen_radial_distribution(num_samples) * randu_on_sphere((num_samples, self.dim), device=self.device)
        return _latent(self, num_samples, self.dim, self._sorted_radii(), self.norm_map == "log", seed, particle_offset, U, Z)

    def _sorted_radii(self):
        cache = getattr(self, "_rT_sorted", None)
        if cache is None or cache[0] is not self.r_T:
            self._rT_sorted = (self.r_T, torch.sort(self.r_T.to(self.device).float().contiguous())[0])
        return self._rT_sorted[1]

    def cond_latent_sample(self, t_, T, x):
        r_x = torch.linalg.norm(x.detach().to(self.device), dim=1).reshape(x.shape[0], 1)
        return r_x * randu_on_sphere((x.shape[0], self.dim), device=self.device)

    def log_latent_pdf(self, yT):
        r = torch.linalg.norm(yT.detach().to(self.device), dim=1)
        return self._kde_logpdf(r) - self.cst_log_dens


def _latent(sde, num_samples, d, r_sorted, log_map, seed, particle_offset, U=None, Z=None):
    import ctypes as C
    dev = sde.device
    handle = _lib.ctx(dev)
    out = torch.empty((num_samples, d), device=dev, dtype=torch.float32)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    U = None if U is None else _lib.f32c(U.reshape(-1), dev)
    Z = None if Z is None else _lib.f32c(Z, dev)
    _lib.check(_lib.lib().msgm_latent_sample(handle, _lib.ptr(r_sorted), 0 if r_sorted is None else r_sorted.numel(),
                                             int(bool(log_map)), int(r_sorted is not None), _lib.ptr(U), _lib.ptr(Z),
                                             _lib.ptr(out), d, num_samples, int(seed), int(particle_offset),
                                             _lib.stream_ptr(dev)))
    return out


# ---- Hutchinson probes (reference SDEs.py:514-536) ---------------------------------------------------------------
def sample_rademacher(shape, device):
    return (torch.rand(*shape, device=device).ge(0.5)).float() * 2 - 1


def sample_gaussian(shape, device):
    return torch.randn(*shape, device=device)


def randu_on_sphere(shape, device):
    X = torch.randn(*shape, device=device)
    return X / torch.linalg.norm(X, dim=1).reshape(shape[0], 1)


def sample_v(shape, device, vtype='rademacher'):
    if vtype == 'rademacher':
        return sample_rademacher(shape, device=device)
    if vtype in ('normal', 'gaussian'):
        return sample_gaussian(shape, device=device)
    if vtype == 'uniform':
        return randu_on_sphere(shape, device=device)
    return None  # the reference builds an Exception without raising it (SDEs.py:535-536)


_VTYPES = {"rademacher": 0, "normal": 1, "gaussian": 1, "uniform": 2}  # msgm_vtype


def _prepare_dim_ok(base, d):
    """msgm_ssm_prepare covers d <= 32 (any SDE) and d <= 4096 for the sparse tensor / the additive SDE."""
    return d <= 32 or (d <= 4096 and (isinstance(base, SGMsde) or getattr(base, "sparseTensor", False)))


class PluginReverseSDE(torch.nn.Module):
    """Reverse-time SDE from a base SDE and a score net ``a`` (reference SDEs.py:538-729).

    ``state_dict()`` keys are ``T``, ``base_sde.T`` and ``a.*`` exactly as in the reference.
    """

    def __init__(self, base_sde, drift_a, T, vtype='rademacher', debias=False, ssm_intT=False,
                 deviceReverseSDE='cpu'):
        super().__init__()
        self.base_sde = base_sde.to(deviceReverseSDE)
        self.a = drift_a
        self.T = T.to(deviceReverseSDE)
        self.vtype, self.ssm_intT, self.debias = vtype, ssm_intT, debias
        self.deviceReverseSDE = deviceReverseSDE

    def mu(self, t, y, lmbd=0.):
        return self.ga_m_drift(self.T - t, y, lmbd)

    def ga_m_drift(self, s, y, lmbd=0.):
        b = self.base_sde
        return (1. - 0.5 * lmbd) * self.ga(s, y) - b.f(s, y) + (1. - lmbd) * b.div_Sigma(s, y)

    def ga(self, s, y):
        b = self.base_sde
        a = self.a(y, s.squeeze())
        g = b.g(s, y, b.sparseTensor)
        if b.sparseTensor:
            I, _, K = b.IJK()
            dx = torch.zeros(y.shape[0], y.shape[1], device=y.device)
            dx.scatter_add_(1, I.unsqueeze(0).expand(y.shape[0], -1), g * a[:, K])
            return dx
        return torch.einsum('bij, bj -> bi', g, a) if g.dim() > 2 else g * a

    def mu_Strato(self, t, y, lmbd=0.):
        return self.mu(t, y, lmbd) - 0.5 * (1. - lmbd) * self.base_sde.div_Sigma(self.T - t, y)

    def sigma(self, t, y, lmbd=0., sparse=False):
        return (1. - lmbd) ** 0.5 * self.base_sde.g(self.T - t, y, sparse)

    # ---- training loss ------------------------------------------------------------------------------------------
    def ssm(self, x):
        """Per-sample sliced-score-matching loss (B,), differentiable w.r.t. the net (reference SDEs.py:607-614)."""
        from . import ssm_fused
        return ssm_fused.ssm(self, x)

    def ssm_loss(self, t_, x, y, v=None):
        from . import ssm_fused
        return ssm_fused.ssm_loss(self, t_, x, y, v)

    def sample_txy(self, x):
        """(t, x, y_t) with t ~ U(0,T) floored at t_epsilon (reference SDEs.py:648-693)."""
        if self.ssm_intT:
            raise NotImplementedError("ssm_intT=True raises NameError in the reference (SDEs.py:700); not built")
        if getattr(self, "device_rng", False) and x.is_cuda and x.dim() == 2 and _prepare_dim_ok(self.base_sde, x.shape[1]) \
                and self.vtype in _VTYPES and isinstance(self.base_sde, (MSGMsde, SGMsde)):
            t_, y, _ = self._prepare(x, with_v=False)
            return t_, x, y
        with torch.no_grad():
            t_ = self.sample_t(x)
            y = self.base_sde.sample(t_, x)
        return t_, x, y

    def _prepare(self, x, with_v=True):
        """(t, y_t, v) from ONE launch (msgm_ssm_prepare): every draw is in-kernel Philox keyed by (seed + device
        counter, global row), so the call neither touches the host generator nor synchronises -- the path that
        ``device_rng = True`` selects and that train.GraphedSsmStep records.  ``_rng`` = (seed, counter tensor or None,
        row offset) is set by the trainer; without it each call draws a fresh seed from the host generator."""
        import ctypes as C
        base, dev = self.base_sde, x.device
        B, d = x.shape
        xc = _lib.f32c(x, dev)
        t_ = torch.empty(B, 1, device=dev, dtype=torch.float32)
        v = torch.empty(B, d, device=dev, dtype=torch.float32)
        y = torch.empty(B, d, device=dev, dtype=torch.float32)
        seed, counter, offset = getattr(self, "_rng", None) or (int(torch.randint(0, 2 ** 62, (1,)).item()), None, 0)
        sd, keep = base.desc(dev)
        sd.dim = d  # SGMsde has no dim of its own
        n_tot = int(base.num_steps_forward)
        T_host = _lib.host_float(base, "T")
        grid = getattr(base, "_fwd_grid", None)
        if grid is None or grid[0] != (n_tot, T_host, str(dev)):
            base._fwd_grid = ((n_tot, T_host, str(dev)), (torch.linspace(0, 1, n_tot + 1) * T_host).to(dev))
        _lib.check(_lib.lib().msgm_ssm_prepare(
            _lib.ctx(dev), C.byref(sd), _lib.ptr(xc), _lib.ptr(t_), _lib.ptr(v), _lib.ptr(y), n_tot,
            _lib.ptr(base._fwd_grid[1]), float(base.t_epsilon), _VTYPES[self.vtype], seed, _lib.ptr(counter), int(offset), B,
            _lib.stream_ptr(dev)))
        self._last_v = v if with_v else None
        return t_, y, v

    def sample_t(self, x):
        shape = [x.size(0), ] + [1 for _ in range(x.ndim - 1)]
        if getattr(self, "device_rng", False):
            t_ = torch.rand(shape, device=x.device, dtype=x.dtype) * self.T
        else:  # the reference draws on the host (SDEs.py:686); kept so that seeded runs consume the same CPU stream
            t_ = torch.rand(shape).to(x) * self.T
        m = (t_ <= self.base_sde.t_epsilon).float()
        return m * self.base_sde.t_epsilon + (1. - m) * t_

    def elbo_random_t_slice(self, x):
        qt = 1 / self.T
        loss_ssm = self.ssm(x) / qt
        t_, x, y = self.sample_txy(x)
        yT = self.cond_latent_sample(t_, self.base_sde.T, x)
        lp = self.base_sde.log_latent_pdf(yT).view(x.size(0), -1).sum(1)
        return lp - loss_ssm

    def latent_sample(self, num_samples, n):
        return self.base_sde.latent_sample(num_samples, n)

    def cond_latent_sample(self, t_, T, x):
        return self.base_sde.cond_latent_sample(t_, T, x)
