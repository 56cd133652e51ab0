"""MMD metric: drop-in for the reference's quantitative_comparison.py (compute_kernel, compute_mmd).

``compute_mmd`` is one tiled pairwise-reduction kernel (msgm_mmd_sums, double accumulation); the reference builds an
(N, M, d) broadcast per kernel matrix (quantitative_comparison.py:22-36).  ``compute_kernel`` returns the full matrix and
is kept as a tensor expression for API compatibility.
"""
from __future__ import annotations

import torch

from . import _lib


@torch.no_grad()
def compute_kernel(x, y):
    dim = x.size(1)
    return torch.exp(-(x.unsqueeze(1) - y.unsqueeze(0)).pow(2).mean(2) / float(dim))


@torch.no_grad()
def compute_mmd(x, y):
    dev = x.device
    handle = _lib.ctx(dev)  # RuntimeError on CPU tensors: no CPU fallback
    xc, yc = _lib.f32c(x, dev), _lib.f32c(y, dev)
    sums = torch.empty(3, device=dev, dtype=torch.float64)
    _lib.check(_lib.lib().msgm_mmd_sums(handle, _lib.ptr(xc), xc.shape[0], _lib.ptr(yc), yc.shape[0], xc.shape[1],
                                        _lib.ptr(sums), _lib.stream_ptr(dev)))
    n, m = xc.shape[0], yc.shape[0]
    return (sums[0] / (n * n) + sums[1] / (m * m) - 2 * sums[2] / (n * m)).to(torch.float32)
