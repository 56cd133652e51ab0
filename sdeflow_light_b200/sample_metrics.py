"""Sample-quality numbers of the reference's evaluation step, on device-resident particles.

The reference computes them inside plotting routines on CPU/numpy copies of the samples:

* ``plot_survival_simple`` (own_plotting.py:689-857): survival curve S(R) = P(|x| > R) of test and generated samples on a
  shared log-spaced radius grid, and the tail exponent alpha of a log-log fit  ->  ``survival_curves`` (same dictionary as
  its ``return_survival=True`` result, without the figure);
* ``preprocessing`` (own_plotting.py:339-394): covariances, per-dimension variances, their distances to white noise and the
  energies E|x|^2  ->  ``covariance_energy_report``.

The O(N d) / O(N d^2) passes over the particles are hand-written kernels (msgm_row_norm_stats, msgm_survival_counts,
msgm_moments); what remains on the host is arithmetic on the 200-point grid and on d x d matrices, written to follow the
reference line by line.  No sort is needed: ``searchsorted(sort(norms), R, side='right')`` is a count, and the tail
threshold test ``R_g >= sorted[-k-1]`` is ``counts[g] <= k``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _norms(x, std_norm):
    """(norms (n,) device fp32, smallest positive norm, largest norm) of ``x * std_norm`` (own_plotting.py:646-653,729-736)."""
    dev = x.device
    xc = _lib.f32c(x, dev)
    n, d = xc.shape
    scale = None
    if std_norm is not None:
        scale = _lib.f32c(torch.as_tensor(std_norm, dtype=torch.float32).reshape(-1), dev)
        if scale.numel() == 1:
            scale = scale.expand(d).contiguous()
    norms = torch.empty(n, device=dev, dtype=torch.float32)
    mm = torch.empty(2, device=dev, dtype=torch.float32)
    _lib.check(_lib.lib().msgm_row_norm_stats(_lib.ctx(dev), _lib.ptr(xc), _lib.ptr(scale), _lib.ptr(norms), _lib.ptr(mm),
                                              d, n, _lib.stream_ptr(dev)))
    lo, hi = mm.view(torch.int32).tolist()
    minpos = None if lo == -1 else float(np.array([lo], dtype=np.int32).view(np.float32)[0])  # 0xFFFFFFFF: no positive norm
    return norms, minpos, float(np.array([hi], dtype=np.int32).view(np.float32)[0])


def _compute_common_R_grid(spans, n_points: int = 200) -> np.ndarray:
    """own_plotting.py:616-632 on the (smallest positive norm, largest norm) pairs of the data sets."""
    mins = [np.float32(lo) for lo, _ in spans if lo is not None]
    maxs = [np.float32(hi) for _, hi in spans]
    if len(maxs) == 0:
        raise ValueError("No data provided to build R grid.")
    min_pos = min(mins) if len(mins) > 0 else 1e-12
    max_val = max(maxs)
    upper = max_val if max_val > min_pos else min_pos * 10.0
    return np.logspace(np.log10(min_pos * 0.9), np.log10(upper), num=n_points)


def _empirical_survival_from_norms(norms: torch.Tensor, R_grid: np.ndarray):
    """(S, counts) with counts[g] = #{norms > R_g} (own_plotting.py:635-640), one pass over device-resident norms."""
    dev = norms.device
    grid = torch.from_numpy(np.ascontiguousarray(R_grid, dtype=np.float64)).to(dev)
    counts = torch.empty(grid.numel(), device=dev, dtype=torch.int64)
    scratch = torch.empty(grid.numel() + 1, device=dev, dtype=torch.int64)
    _lib.check(_lib.lib().msgm_survival_counts(_lib.ctx(dev), _lib.ptr(norms), norms.numel(), _lib.ptr(grid), grid.numel(),
                                               _lib.ptr(counts), _lib.ptr(scratch), _lib.stream_ptr(dev)))
    counts = counts.cpu().numpy()
    return counts.astype(float) / float(norms.numel()), counts


def _tail_fit_loglog(R_grid, S_vals, counts, n, tail_frac: float = 0.05, tail_k=None):
    """own_plotting.py:656-700 with the order-statistic threshold expressed through the counts."""
    if n < 10:
        return None, None, None
    if tail_k is None:
        k = max(10, int(np.clip(np.ceil(n * tail_frac), 10, n - 1)))
    else:
        k = int(min(max(1, tail_k), n - 1))
    mask = counts <= k  # R_grid >= sorted_norms[-k-1]
    if not np.any(mask):
        return None, k, None
    R_tail, S_tail = R_grid[mask], S_vals[mask]
    positive_mask = S_tail > 0
    if np.sum(positive_mask) < 3:
        return None, k, None
    R_tail, S_tail = R_tail[positive_mask], S_tail[positive_mask]
    b, a = np.polyfit(np.log(R_tail), np.log(S_tail), 1)
    return float(-b), int(k), np.exp(a) * (R_grid ** b)


@torch.no_grad()
def survival_curves(x=None, x_ref=None, std_norm=None, n_points: int = 200, tail_frac: float = 0.05, tail_k=None):
    """The ``survival_dict`` of ``plot_survival_simple(..., return_survival=True)`` (own_plotting.py:846-853)."""
    if x is None and x_ref is None:
        raise ValueError("At least one of x or x_ref must be provided.")
    data = {}
    for name, t in (("reference", x_ref), ("generated", x)):
        if t is not None:
            assert t.ndim == 2
            data[name] = _norms(t, std_norm)
    R_grid = _compute_common_R_grid([(lo, hi) for _, lo, hi in data.values()], n_points=n_points)
    out = {"R_grid": R_grid, "fits": {}}
    for name, short in (("reference", "ref"), ("generated", "gen")):
        if name in data:
            norms = data[name][0]
            S, counts = _empirical_survival_from_norms(norms, R_grid)
            alpha, k, _ = _tail_fit_loglog(R_grid, S, counts, norms.numel(), tail_frac=tail_frac, tail_k=tail_k)
            out[name] = {"S": S, "counts": counts, "N": norms.numel()}
            out["fits"][short] = {"alpha": alpha, "k": k}
        else:
            out[name] = {"S": None, "counts": None, "N": 0}
            out["fits"][short] = {"alpha": None, "k": None}
    return out


def _moments(x):
    """(n, mean (d,), cov (d,d) with correction 1, energy E|x|^2) in float64 from one pass (msgm_moments)."""
    dev = x.device
    xc = _lib.f32c(x, dev)
    n, d = xc.shape
    colsum = torch.empty(d, device=dev, dtype=torch.float64)
    gram = torch.empty(d, d, device=dev, dtype=torch.float64)
    _lib.check(_lib.lib().msgm_moments(_lib.ctx(dev), _lib.ptr(xc), n, d, _lib.ptr(colsum), _lib.ptr(gram),
                                       _lib.stream_ptr(dev)))
    gram = gram.cpu()
    gram = torch.triu(gram) + torch.triu(gram, 1).T  # the kernel fills the upper 32x32 tiles
    mean = colsum.cpu() / n
    cov = (gram - n * torch.outer(mean, mean)) / (n - 1)
    return n, mean, cov, float(torch.trace(gram) / n)


@torch.no_grad()
def covariance_energy_report(xtest, xgen_forward):
    """The numbers ``preprocessing`` prints (own_plotting.py:339-394): covariances of the test set and of the noised set,
    their distances to the converged / weak-white-noise covariances, and the energies."""
    d = xtest.shape[1]
    _, _, cov_xtest, energy_xtest = _moments(xtest)
    _, _, cov_xgen, energy_xgen = _moments(xgen_forward)
    xgen_var_mean = torch.diagonal(cov_xgen).mean()
    xtest_var_mean = torch.diagonal(cov_xtest).mean()
    eye = torch.eye(d, dtype=torch.float64)
    conv = xtest_var_mean * eye
    wwn = xgen_var_mean * eye
    den = lambda c: torch.sqrt(d * torch.trace(c ** 2))  # noqa: E731  (elementwise square, as in the reference)
    return {
        "cov_xtest": cov_xtest, "cov_xgen_forward": cov_xgen,
        "d_cov_xtest": float(torch.norm(cov_xtest - conv) / den(conv)),
        "d_cov_xgen_forward": float(torch.norm(cov_xgen - conv) / den(conv)),
        "d_cov_xgen_forward_wwn": float(torch.norm(cov_xgen - wwn) / den(wwn)),
        "energy_xtest": energy_xtest, "energy_xgen_forward": energy_xgen,
        "energy_ratio": energy_xgen / energy_xtest,
    }
