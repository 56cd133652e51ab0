"""CPU oracle for the MSGM / SGM hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  The shipped path (``sdeflow_light_b200``) never does; it fails loudly
when the CUDA library is missing.

This is a functional restatement, in plain PyTorch-on-CPU fp32 ops, of the reference algorithm
(vressegu/sdeflow-light).  Each function cites the reference ``file:line`` it follows.  It deliberately issues
the same tensor-level operations in the same order as the reference (including its redundant evaluations of
``g`` and ``f``) for two reasons: fp32 results then agree with the reference to rounding, and its run time is a
fair stand-in for the reference's CPU path when the reference itself cannot travel to the GPU box.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the oracle is
pinned against *outputs of the reference itself run in the build container*: ``oracle/check_against_reference.py``
imports ``/root/reference`` live and compares every function here on seeded inputs, and
``tests/golden/make_golden.py`` stores reference outputs as fixtures that ``tests/test_oracle_golden.py``
replays on every CPU test run.

State (``OSde``) is a plain namespace, not an nn.Module; nets are passed as callables ``a(y, s) -> (B, d)``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import torch

# ----------------------------------------------------------------------------------------------------------
# SDE description
# ----------------------------------------------------------------------------------------------------------


@dataclass
class OSde:
    """What the reference keeps on ``SGMsde`` / ``MSGMsde`` objects (SDEs.py:54-64, 226-251)."""

    kind: str  # "sgm" | "msgm_dense" | "msgm_sparse"
    dim: int
    beta_min: float = 0.1
    beta_max: float = 20.0
    T: float = 1.0
    t_epsilon: float = 1e-3
    num_steps_forward: int = 16
    G: Optional[torch.Tensor] = None  # (d,d,d), dense only
    L_G: Optional[torch.Tensor] = None  # (d,d)
    G_I: Optional[torch.Tensor] = None  # sparse COO rows, (2d,)
    G_J: Optional[torch.Tensor] = None
    G_K: Optional[torch.Tensor] = None
    G_V: Optional[torch.Tensor] = None
    r_T: Optional[torch.Tensor] = None  # sorted or unsorted (log-)radii of the training data
    norm_map: Optional[str] = None

    @property
    def sparse(self) -> bool:
        return self.kind == "msgm_sparse"

    @property
    def T_tensor(self) -> torch.Tensor:
        return torch.tensor([self.T], dtype=torch.float32)


def draw_dense_G(n: int) -> torch.Tensor:
    """Random skew-symmetric noise tensor, normalised so tr(L_G) = -n/2 (SDEs.py:315-341).

    Consumes ``n`` draws of ``torch.randn(n, n)`` from the global generator, like the reference.
    """
    G = torch.zeros(n, n, n)
    for k in range(n):
        F = torch.randn(n, n)
        G[:, :, k] = 0.5 * (F - F.T)
    L = 0.5 * torch.einsum("ijk,jmk->im", G, G)
    return torch.sqrt(-0.5 * n / torch.trace(L)) * G


def cyclic_sparse_G(n: int):
    """COO lists of the cyclic 2-nonzeros-per-slice tensor (SDEs.py:369-396).

    Entry 2k is G[k, k+1, k] = +c, entry 2k+1 is G[k+1, k, k] = -c with c = sqrt(2)/2, indices mod n.
    """
    c = 0.5 * torch.sqrt(torch.tensor(2, dtype=torch.float32))
    I, J, K, V = [], [], [], []
    for k in range(n):
        kp = (k + 1) % n
        I += [k, kp]
        J += [kp, k]
        K += [k, k]
        V += [c, -c]
    return (torch.tensor(I), torch.tensor(J), torch.tensor(K), torch.tensor(V, dtype=torch.float32))


def make_sgm(dim, beta_min=0.1, beta_max=20.0, T=1.0, t_epsilon=1e-3, num_steps_forward=16) -> OSde:
    return OSde("sgm", dim, beta_min, beta_max, T, t_epsilon, num_steps_forward)


def make_msgm(y0: torch.Tensor, dense=True, beta_min=0.1, beta_max=20.0, T=1.0, t_epsilon=1e-3,
              num_steps_forward=16, norm_map="log", G: Optional[torch.Tensor] = None) -> OSde:
    """Mirror of the parts of ``MSGMsde.__init__`` the hot path needs (SDEs.py:226-251)."""
    d = y0.shape[1]
    r_T = torch.linalg.norm(y0, dim=1)
    if norm_map == "log":
        r_T = torch.log(r_T + 1e-6)
    sde = OSde("msgm_dense" if dense else "msgm_sparse", d, beta_min, beta_max, T, t_epsilon,
               num_steps_forward, r_T=r_T, norm_map=norm_map)
    if dense:
        sde.G = draw_dense_G(d) if G is None else G
        sde.L_G = 0.5 * torch.einsum("ijk,jmk->im", sde.G, sde.G)
    else:
        sde.G_I, sde.G_J, sde.G_K, sde.G_V = cyclic_sparse_G(d)
        sde.L_G = 0.5 * torch.eye(d)
    return sde


# ----------------------------------------------------------------------------------------------------------
# coefficient functions
# ----------------------------------------------------------------------------------------------------------


def beta(sde: OSde, s):
    """SDEs.py:72-73."""
    return sde.beta_min + (sde.beta_max - sde.beta_min) * s


def coef_f(sde: OSde, s, y):
    """Ito drift of the forward SDE (SDEs.py:183-184 SGM, 410-415 MSGM incl. the sparse '+' sign quirk)."""
    b = beta(sde, s)
    if sde.kind == "sgm":
        return -0.5 * b * y
    if sde.sparse:
        return 0.5 * b * y
    return torch.einsum("ij,bj->bi", sde.L_G, b * y)


def coef_f_strato(sde: OSde, s, y):
    """SDEs.py:186-187 (SGM), 417-418 (MSGM)."""
    if sde.kind == "sgm":
        return -0.5 * beta(sde, s) * y
    return torch.zeros_like(y)


def coef_div_sigma(sde: OSde, s, y):
    """SDEs.py:189-190 (SGM: 0), 420-421 (MSGM: 2 f)."""
    if sde.kind == "sgm":
        return torch.zeros_like(y)
    return 2 * coef_f(sde, s, y)


def coef_g(sde: OSde, s, y, sparse=False):
    """Diffusion coefficient in the reference's three shapes (SDEs.py:192-194, 423-432).

    SGM -> (B,d) ; MSGM dense -> (B,d,d) materialised ; MSGM sparse -> (B,2d) aligned with G_V.
    """
    b = beta(sde, s)
    if sde.kind == "sgm":
        return torch.ones_like(y) * b ** 0.5
    if sparse:
        yJ = (b ** 0.5) * y[:, sde.G_J]
        return sde.G_V.unsqueeze(0) * yJ
    return torch.einsum("ijk,bj->bik", sde.G, (b ** 0.5) * y)


def apply_sigma(sde: OSde, sigma, w):
    """``sigma . w`` the way EMstep / ga contract it (sde_scheme.py:27-38, SDEs.py:568-578)."""
    if sde.sparse:
        prod = sigma * w[:, sde.G_K]
        out = torch.zeros_like(w)
        out.scatter_add_(1, sde.G_I.unsqueeze(0).expand(w.size(0), -1), prod)
        return out
    if sigma.dim() > 2:
        return torch.einsum("bij,bj->bi", sigma, w)
    return sigma * w


def em_increment(sde: OSde, mu, delta, sigma, dW):
    """EMstep (sde_scheme.py:18-40)."""
    return mu * delta + apply_sigma(sde, sigma, dW)


# ----------------------------------------------------------------------------------------------------------
# score networks
# ----------------------------------------------------------------------------------------------------------


def swish(x):
    """NN.py:52-53."""
    return torch.sigmoid(x) * x


def normalize_log_radius(x, eps=1e-6):
    """NN.py:64-70."""
    norm = torch.norm(x, dim=-1, keepdim=True) + eps
    return x / norm, torch.log(norm)


@dataclass
class OMlp:
    """Weights of NN.MLP in torch.nn.Linear layout: W[l] is (out, in), b[l] is (out,) (NN.py:98-106)."""

    W: Sequence[torch.Tensor]
    b: Sequence[torch.Tensor]
    premodule: bool
    input_dim: int

    def __call__(self, y, s):
        """NN.py:108-120."""
        y = y.view(-1, self.input_dim)
        s = s.view(-1, 1).float()
        if self.premodule:
            h, ln = normalize_log_radius(y)
            y = torch.cat([h, ln], dim=-1)
        h = torch.cat([y, s], dim=1)
        for l in range(3):
            h = swish(torch.addmm(self.b[l], h, self.W[l].t()))
        return torch.addmm(self.b[3], h, self.W[3].t())

    def parameters(self):
        out = []
        for w, b in zip(self.W, self.b):
            out += [w, b]
        return out


def init_mlp(dim: int, premodule: bool, hidden=128, seed: Optional[int] = None, scale: float = 1.0) -> OMlp:
    """nn.Linear default init (kaiming-uniform(a=sqrt5) == U(-1/sqrt(in), 1/sqrt(in)) for W and b).

    ``scale`` widens the last layer so that an untrained net still produces an O(1) drift in tests.
    """
    g = torch.Generator().manual_seed(seed) if seed is not None else None
    sizes = [dim + (1 if premodule else 0) + 1, hidden, hidden, hidden, dim]
    W, b = [], []
    for l in range(4):
        k = 1.0 / math.sqrt(sizes[l])
        W.append((torch.rand(sizes[l + 1], sizes[l], generator=g) * 2 - 1) * k)
        b.append((torch.rand(sizes[l + 1], generator=g) * 2 - 1) * k)
    W[3] = W[3] * scale
    b[3] = b[3] * scale
    return OMlp(W, b, premodule, dim)


# ----------------------------------------------------------------------------------------------------------
# reverse (generative) SDE built from base SDE + net  (PluginReverseSDE, SDEs.py:556-588)
# ----------------------------------------------------------------------------------------------------------


@dataclass
class OReverse:
    sde: OSde
    a: Callable

    def ga(self, s, y):
        """SDEs.py:563-579: g evaluated, net evaluated, dense case evaluates g a second time."""
        sde = self.sde
        g = coef_g(sde, s, y, sde.sparse)
        if sde.sparse:
            return apply_sigma(sde, g, self.a(y, s.squeeze()))
        if g.dim() > 2:
            return torch.einsum("bij,bj->bi", coef_g(sde, s, y), self.a(y, s.squeeze()))
        return g * self.a(y, s.squeeze())

    def ga_m_drift(self, s, y, lmbd=0.0):
        """SDEs.py:560-561."""
        return (1.0 - 0.5 * lmbd) * self.ga(s, y) - coef_f(self.sde, s, y) \
            + (1.0 - lmbd) * coef_div_sigma(self.sde, s, y)

    def mu(self, t, y, lmbd=0.0):
        """SDEs.py:556-557."""
        return self.ga_m_drift(self.sde.T_tensor - t, y, lmbd)

    def mu_strato(self, t, y, lmbd=0.0):
        """SDEs.py:583-584."""
        return self.mu(t, y, lmbd) - 0.5 * (1.0 - lmbd) * coef_div_sigma(self.sde, self.sde.T_tensor - t, y)

    def sigma(self, t, y, lmbd=0.0, sparse=False):
        """SDEs.py:587-588."""
        return (1.0 - lmbd) ** 0.5 * coef_g(self.sde, self.sde.T_tensor - t, y, sparse)


@dataclass
class OForward:
    """forward_SDE adapter (SDEs.py:30-47): Stratonovich drift f_strato, diffusion g, in forward time."""

    sde: OSde

    def mu(self, s, y, lmbd=0.0):
        return self.mu_strato(s, y) + 0.5 * coef_div_sigma(self.sde, s, y)

    def mu_strato(self, s, y, lmbd=0.0):
        return coef_f_strato(self.sde, s, y)

    def sigma(self, s, y, lmbd=0.0, sparse=False):
        return coef_g(self.sde, s, y, sparse)


# ----------------------------------------------------------------------------------------------------------
# integrators  (sde_scheme.py:43-269)
# ----------------------------------------------------------------------------------------------------------


def _row_rescale(x, r0):
    return x * (r0 / torch.norm(x, dim=1))[:, None]


@torch.no_grad()
def integrate(proc, x_0, num_steps, scheme="rk4", lmbd=0.0, keep_all_samples=True, samplesToKeep=None,
              include_t0=False, T_=None, norm_correction=False, noise: Optional[torch.Tensor] = None):
    """EM / Heun / RK4-Stratonovich loops with the reference's capture modes.

    ``proc`` is an OReverse or OForward.  ``noise`` (num_steps,B,d), if given, supplies the standard-normal
    draw of each step; otherwise ``torch.randn_like`` is called once per step exactly where the reference
    calls it (sde_scheme.py:84,144,227), so seeding the global generator replays the reference's noise.
    Output shapes follow sde_scheme.py:94-99,264-269.
    """
    sde = proc.sde
    T_ = float(torch.tensor([sde.T], dtype=torch.float32).item()) if T_ is None else float(T_)
    B = x_0.size(0)
    delta = T_ / num_steps
    ts = torch.linspace(0, 1, num_steps + 1) * T_
    sp = sde.sparse
    x = x_0.detach().clone()
    r0 = torch.norm(x, dim=1) if norm_correction else None
    xs = None
    if keep_all_samples:
        xs = torch.zeros(B, x_0.shape[1], num_steps + (1 if include_t0 else 0))
        if include_t0:
            xs[:, :, 0] = x
    elif samplesToKeep is not None:
        if len(samplesToKeep) != B:
            raise ValueError("Error: len(samplesToKeep) must correspond to batch size.")
        xs = torch.zeros(B, x_0.shape[1])
    t = torch.zeros(B, 1)
    sqrt_delta = delta ** 0.5
    for i in range(num_steps):
        t.fill_(ts[i].item())
        xi = noise[i] if noise is not None else torch.randn_like(x)
        if scheme == "em":  # sde_scheme.py:80-86, Ito drift
            mu = proc.mu(t, x, lmbd=lmbd)
            sg = proc.sigma(t, x, lmbd=lmbd, sparse=sp)
            x = x + em_increment(sde, mu, delta, sg, delta ** 0.5 * xi)
        elif scheme == "heun":  # sde_scheme.py:138-158
            mu1 = proc.mu_strato(t, x, lmbd=lmbd)
            sg1 = proc.sigma(t, x, lmbd=lmbd, sparse=sp)
            dW = delta ** 0.5 * xi
            xp = x + em_increment(sde, mu1, delta, sg1, dW)
            mu2 = proc.mu_strato(t + delta, xp, lmbd=lmbd)
            sg2 = proc.sigma(t + delta, xp, lmbd=lmbd, sparse=sp)
            x = x + em_increment(sde, mu1 + mu2, delta / 2, sg1 + sg2, dW / 2)
        elif scheme == "rk4":  # sde_scheme.py:223-253
            dW = sqrt_delta * xi
            k1 = em_increment(sde, proc.mu_strato(t, x, lmbd=lmbd), delta,
                              proc.sigma(t, x, lmbd=lmbd, sparse=sp), dW)
            xm = x + k1 / 2
            k2 = em_increment(sde, proc.mu_strato(t + delta / 2, xm, lmbd=lmbd), delta,
                              proc.sigma(t + delta / 2, xm, lmbd=lmbd, sparse=sp), dW)
            xm = x + k2 / 2
            k3 = em_increment(sde, proc.mu_strato(t + delta / 2, xm, lmbd=lmbd), delta,
                              proc.sigma(t + delta / 2, xm, lmbd=lmbd, sparse=sp), dW)
            xe = x + k3
            k4 = em_increment(sde, proc.mu_strato(t + delta, xe, lmbd=lmbd), delta,
                              proc.sigma(t + delta, xe, lmbd=lmbd, sparse=sp), dW)
            x = x + (k1 + 2 * k2 + 2 * k3 + k4) / 6
        else:
            raise ValueError(scheme)
        if norm_correction:
            x = _row_rescale(x, r0)
        if keep_all_samples:
            xs[:, :, i + include_t0] = x
        elif samplesToKeep is not None:
            if (i + include_t0) in samplesToKeep:
                m = (samplesToKeep == (i + include_t0)).flatten()
                xs[m, :] = x[m, :]
    if keep_all_samples:
        return torch.permute(xs, (2, 0, 1))
    if samplesToKeep is None:
        return x.clone()
    return xs


# ----------------------------------------------------------------------------------------------------------
# forward noising used by training  (SDEs.py:78-132, 134-146, 196-199, 434-436)
# ----------------------------------------------------------------------------------------------------------


def sgm_mean_weight(sde, t):
    """SDEs.py:177-178."""
    return torch.exp(-0.25 * t ** 2 * (sde.beta_max - sde.beta_min) - 0.5 * t * sde.beta_min)


def sgm_var(sde, t):
    """SDEs.py:180-181."""
    return 1.0 - torch.exp(-0.5 * t ** 2 * (sde.beta_max - sde.beta_min) - t * sde.beta_min)


@torch.no_grad()
def noise_forward(sde: OSde, t, y0):
    """``base_sde.sample(t, y0)``: closed form for SGM, simulation for MSGM.

    MSGM (SDEs.py:78-122): run N_fwd RK4 steps on the whole batch capturing row k at step
    n_k = trunc(N_fwd t_k / T); rows with n_k == 0 are re-simulated alone with ONE step of size t_k.
    RNG order is the reference's: N_fwd x randn(B,d), then one randn(1,d) per n_k==0 row in index order.
    """
    if sde.kind == "sgm":
        mu = sgm_mean_weight(sde, t) * y0
        std = sgm_var(sde, t) ** 0.5
        return torch.randn_like(y0) * std + mu
    n_tot = sde.num_steps_forward
    n_int = torch.trunc(n_tot * t / sde.T_tensor).to(torch.int)
    n_int[(t >= sde.T_tensor)] = n_tot
    fwd = OForward(sde)
    cap = integrate(fwd, y0, n_tot, "rk4", 0.0, keep_all_samples=False, samplesToKeep=n_int, include_t0=True)
    yt = torch.zeros_like(y0)
    for k in range(y0.shape[0]):
        if n_int[k] > 0:
            yt[k] = cap[k]
        else:
            yt[k] = integrate(fwd, y0[k][None], 1, "rk4", 0.0, keep_all_samples=False, include_t0=False,
                              T_=float(t[k]))[0]
    return yt


def sample_t(sde: OSde, x):
    """SDEs.py:684-693: U(0,T) floored to t_epsilon through a float mask."""
    t_ = torch.rand([x.size(0)] + [1] * (x.ndim - 1)).to(x) * sde.T_tensor
    m = (t_ <= sde.t_epsilon).float()
    return m * sde.t_epsilon + (1.0 - m) * t_


def sample_rademacher(shape):
    """SDEs.py:514-515."""
    return (torch.rand(*shape).ge(0.5)).float() * 2 - 1


def randu_on_sphere(shape):
    """SDEs.py:520-526."""
    X = torch.randn(*shape)
    return X / torch.linalg.norm(X, dim=1).reshape(shape[0], 1)


# ----------------------------------------------------------------------------------------------------------
# sliced score matching loss  (SDEs.py:607-646)
# ----------------------------------------------------------------------------------------------------------


def ssm_loss(rev: OReverse, t_, y, v, create_graph=True):
    """Per-sample SSM loss v^T d(mu_to_div)/dy v + |a|^2/2 with the Hutchinson probe ``v`` supplied.

    ``y`` must require grad.  Follows SDEs.py:624-646 including the algebraically-cancelling f / div terms.
    """
    a = rev.a(y, t_.squeeze())
    mu = rev.ga_m_drift(t_, y, 0.0)
    mu_to_div = mu - 0.5 * coef_div_sigma(rev.sde, t_, y)
    jv = torch.autograd.grad(mu_to_div, y, v, create_graph=create_graph)[0]
    mMu = (jv * v).view(y.size(0), -1).sum(1)
    mNu = (a ** 2).view(y.size(0), -1).sum(1) / 2
    return mMu + mNu


def ssm(rev: OReverse, x):
    """``PluginReverseSDE.ssm`` incl. its RNG order: t, forward noise, v (SDEs.py:607-614,648-682)."""
    with torch.no_grad():
        t_ = sample_t(rev.sde, x)
        y = noise_forward(rev.sde, t_, x)
    y.requires_grad_()
    with torch.no_grad():
        v = sample_rademacher(x.shape)
    return ssm_loss(rev, t_, y, v), (t_, y, v)


# ----------------------------------------------------------------------------------------------------------
# prior sampling and metrics
# ----------------------------------------------------------------------------------------------------------


def latent_sample(sde: OSde, num_samples: int):
    """SDEs.py:201-203 (SGM) and 438-493 (MSGM, ecdf radial sampler)."""
    if sde.kind == "sgm":
        return torch.randn(num_samples, sde.dim)
    U = torch.rand(num_samples)
    r = torch.quantile(sde.r_T, U).reshape(num_samples, 1)
    if sde.norm_map == "log":
        r = torch.exp(r) - 1e-6
    return r * randu_on_sphere((num_samples, sde.dim))


def rbf_kernel_mean(x, y):
    """Mean of exp(-|x-y|^2 / dim^2) over all pairs (quantitative_comparison.py:22-36)."""
    dim = x.size(1)
    k = (x.unsqueeze(1) - y.unsqueeze(0)).pow(2).mean(2) / float(dim)
    return torch.exp(-k).mean()


def compute_mmd(x, y):
    """quantitative_comparison.py:38-47."""
    return rbf_kernel_mean(x, x) + rbf_kernel_mean(y, y) - 2 * rbf_kernel_mean(x, y)


def kde_logpdf(samples, bandwidth: float, queries):
    """log density of the 1-D Gaussian KDE of ``samples`` at ``queries``: what the reference asks
    sklearn.neighbors.KernelDensity(kernel='gaussian', bandwidth=h).fit(r_T).score_samples(r) for (SDEs.py:240,261,509;
    scikit-learn is a dependency of the reference, its gaussian kernel is exp(-u^2/2)/(h sqrt(2 pi)), exact sum with the
    default rtol=atol=0).  float64 logsumexp, returned as fp32 like the reference's torch.tensor(...).to(float32)."""
    import numpy as np
    s = np.asarray(samples, dtype=np.float64).reshape(1, -1)
    q = np.asarray(queries, dtype=np.float64).reshape(-1, 1)
    e = -0.5 * ((q - s) / bandwidth) ** 2
    m = e.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(e - m).sum(axis=1))
    return torch.from_numpy((lse - np.log(s.shape[1] * bandwidth * np.sqrt(2.0 * np.pi))).astype("float32"))


def kde_log_normaliser(samples, bandwidth: float):
    """cst_log_dens of MSGMsde.__init__ (SDEs.py:255-265): log of the rectangle-rule integral of the KDE over 1000
    points spanning [min r_T, max r_T]."""
    r = torch.as_tensor(samples, dtype=torch.float32).reshape(-1)
    grid = torch.linspace(float(r.min()), float(r.max()), 1000)
    dens = torch.exp(kde_logpdf(r, bandwidth, grid))
    return torch.log(dens.sum() * (grid[1] - grid[0]))


def log_latent_pdf(sde, yT, bandwidth: float, cst_log_dens):
    """MSGMsde.log_latent_pdf (SDEs.py:503-509): KDE log-density of |y_T| minus the normalising constant."""
    r = torch.linalg.norm(yT, dim=1)
    return kde_logpdf(sde.r_T, bandwidth, r) - cst_log_dens


# ----------------------------------------------------------------------------------------------------------
# synthetic data of the reference driver (data.py:702-778); numpy/sklearn RNG like the reference
# ----------------------------------------------------------------------------------------------------------


def swiss_roll(n: int, noise: float = 0.5) -> torch.Tensor:
    """data.py:711-715."""
    from sklearn.datasets import make_swiss_roll
    return torch.from_numpy(make_swiss_roll(n, noise=noise)[0][:, [0, 2]].astype("float32") / 5.0)


def gaussian_mixture(n: int, dim: int, n_comp: int = 8, seed: int = 0) -> torch.Tensor:
    """Synthetic 'higher-dim Gaussian mixture' of BASELINE.json config 2 (no reference sampler exists;
    SURVEY.md section 8d): means 3 randn(K,d), shared correlation A = randn(d,d) as in data.py:760."""
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(dim, dim, generator=g)
    mu = 3.0 * torch.randn(n_comp, dim, generator=g)
    c = torch.randint(0, n_comp, (n,), generator=g)
    return (torch.randn(n, dim, generator=g) @ A.T) * 0.3 + mu[c]
