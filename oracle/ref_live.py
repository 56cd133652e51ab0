"""Import the UNMODIFIED reference from /root/reference (build container only) -- TEST INFRASTRUCTURE.

The reference is pure Python and cannot travel to the GPU box, so nothing on the ``-m gpu`` / bench / smoke
paths imports this module.  It is used by ``oracle/check_against_reference.py`` (pins the oracle) and by
``tests/golden/make_golden.py`` (writes the committed fixtures).

matplotlib / seaborn / netCDF4 are absent from the image and only used for plots and file IO, so they are
stubbed in ``sys.modules`` before the import.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace
from unittest.mock import MagicMock

import torch

REF_ROOT = os.environ.get("MSGM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "SDEs.py"))


def load() -> SimpleNamespace:
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "seaborn", "netCDF4"):
        sys.modules.setdefault(m, MagicMock())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import SDEs
    import sde_scheme
    import NN
    import NNUnet1D
    import NNUnet
    import quantitative_comparison as qc
    return SimpleNamespace(SDEs=SDEs, sde_scheme=sde_scheme, NN=NN, NNUnet1D=NNUnet1D, NNUnet=NNUnet, qc=qc)


def T_param(T0=1.0):
    """MSGM_higherDim.py:728."""
    return torch.nn.Parameter(torch.FloatTensor([T0]), requires_grad=False)


def build(ref, kind: str, dim: int, x_init=None, premodule=None, beta_min=0.1, beta_max=20.0, t_eps=1e-3,
          n_fwd=16, T0=1.0, net=None):
    """Construct (base_sde, gen_sde, net) the way the driver does (MSGM_higherDim.py:716-746)."""
    T = T_param(T0)
    if net is None:
        net = ref.NN.MLP(input_dim=dim, index_dim=1, hidden_dim=128, premodule=premodule)
    if kind == "sgm":
        base = ref.SDEs.SGMsde(beta_min=beta_min, beta_max=beta_max, t_epsilon=t_eps, T=T,
                               num_steps_forward=n_fwd, device="cpu")
    else:
        base = ref.SDEs.MSGMsde(x_init, beta_min=beta_min, beta_max=beta_max, t_epsilon=t_eps, T=T,
                                num_steps_forward=n_fwd, device="cpu", estim_cst_norm_dens_r_T=False,
                                norm_sampler="ecdf", norm_map="log", denseTensor=(kind == "msgm_dense"),
                                plot_validate=False)
    gen = ref.SDEs.PluginReverseSDE(base, net, T, vtype="rademacher", debias=False, ssm_intT=False,
                                    deviceReverseSDE="cpu")
    return base, gen, net


def to_oracle(base, net=None):
    """Translate reference objects into the oracle's plain structures (same tensors, no copies of code)."""
    from . import msgm_oracle as O
    name = type(base).__name__
    common = dict(beta_min=base.beta_min, beta_max=base.beta_max, T=float(base.T.item()),
                  t_epsilon=base.t_epsilon, num_steps_forward=base.num_steps_forward)
    if name == "SGMsde":
        sde = O.OSde("sgm", dim=-1, **common)
    elif base.sparseTensor:
        sde = O.OSde("msgm_sparse", dim=base.dim, G_I=base.G_I, G_J=base.G_J, G_K=base.G_K, G_V=base.G_V,
                     L_G=base.L_G, r_T=base.r_T, norm_map=base.norm_map, **common)
    else:
        sde = O.OSde("msgm_dense", dim=base.dim, G=base.G, L_G=base.L_G, r_T=base.r_T,
                     norm_map=base.norm_map, **common)
    mlp = None
    if net is not None and type(net).__name__ == "MLP":
        lin = [m for m in net.main if isinstance(m, torch.nn.Linear)]
        mlp = O.OMlp([l.weight.detach() for l in lin], [l.bias.detach() for l in lin],
                     net.premodule is not None, net.input_dim)
        if sde.dim < 0:
            sde.dim = net.input_dim
    return sde, mlp
