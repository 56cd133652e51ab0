"""Write tests/golden/m01_survival_cov_energy.npz from the UNMODIFIED reference's own_plotting.py (build container only).

* ``_compute_common_R_grid``, ``_empirical_survival_from_norms``, ``_tail_fit_loglog`` (own_plotting.py:616-700) are called
  directly, on the norms ``plot_survival_simple`` would compute (``torch.norm(x * std_norm, dim=1)``, :729-736).
* ``preprocessing`` (own_plotting.py:333-423) is run as is with matplotlib / seaborn stubbed (``noising_plots=False``) and
  the numbers it prints (covariance distances, energies) are parsed from its standard output.
"""
import contextlib
import importlib
import io
import json
import os
import re
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "seaborn", "netCDF4"):
    sys.modules.setdefault(m, MagicMock())
sys.modules["matplotlib.pyplot"].subplots.return_value = (MagicMock(), MagicMock())
sys.path.insert(0, "/root/reference")
op = importlib.import_module("own_plotting")
op.plt.subplots.return_value = (MagicMock(), MagicMock())

torch.manual_seed(0)
np.random.seed(0)
d, n_ref, n_gen = 6, 5000, 7000
A = torch.randn(d, d)
x_ref = torch.randn(n_ref, d) @ A.T * 0.7
# generated set: heavier tail (student-like radial factor) so that the tail fit is non-trivial
x_gen = (torch.randn(n_gen, d) @ A.T * 0.7) * (1.0 + 0.5 * torch.rand(n_gen, 1) ** -0.4)
std_norm = torch.rand(d) + 0.5

out = {}
for tag, sn in (("plain", None), ("scaled", std_norm)):
    norms_ref = torch.norm(op._apply_std_norm(x_ref, sn), dim=1).cpu().numpy()
    norms_gen = torch.norm(op._apply_std_norm(x_gen, sn), dim=1).cpu().numpy()
    R = op._compute_common_R_grid([norms_ref, norms_gen], n_points=200)
    S_ref, c_ref = op._empirical_survival_from_norms(norms_ref, R)
    S_gen, c_gen = op._empirical_survival_from_norms(norms_gen, R)
    a_ref, k_ref, _ = op._tail_fit_loglog(R, S_ref, norms_ref, tail_frac=0.05, tail_k=None)
    a_gen, k_gen, _ = op._tail_fit_loglog(R, S_gen, norms_gen, tail_frac=0.05, tail_k=None)
    a_gen_k, k_gen_k, _ = op._tail_fit_loglog(R, S_gen, norms_gen, tail_k=500)
    out.update({f"{tag}_R": R, f"{tag}_S_ref": S_ref, f"{tag}_c_ref": c_ref, f"{tag}_S_gen": S_gen, f"{tag}_c_gen": c_gen,
                f"{tag}_alpha": np.array([a_ref, a_gen, a_gen_k]), f"{tag}_k": np.array([k_ref, k_gen, k_gen_k]),
                f"{tag}_norms_ref": norms_ref, f"{tag}_norms_gen": norms_gen})

xs_forward = torch.stack([x_ref, x_gen[:n_ref]])  # preprocessing uses xs_forward[-1]
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    op.preprocessing(x_ref, xs_forward, 1, "m01", 0, False, False, "/tmp", 1.0, std_norm, std_norm, "cpu")
txt = buf.getvalue()
num = r"= ([-+0-9.eE]+)"
vals = {
    "d_cov_xtest": float(re.search(r"dist cov_xtest to  cov_xgen_forward_converged.*" + num, txt).group(1)),
    "d_cov_xgen_forward": float(re.search(r"dist cov_xgen_forward  to  cov_xgen_forward_converged " + num, txt).group(1)),
    "d_cov_xgen_forward_wwn": float(re.search(r"dist cov_xgen_forward  to  weak white noise.*" + num, txt).group(1)),
    "energy_xtest": float(re.search(r"energy_xtest " + num, txt).group(1)),
    "energy_xgen_forward": float(re.search(r"energy_xgen_forward " + num, txt).group(1)),
    "energy_ratio": float(re.search(r"energy_xgen_forward / energy_xtest " + num, txt).group(1)),
}
meta = dict(d=d, n_ref=n_ref, n_gen=n_gen, n_points=200, tail_frac=0.05, tail_k=500, printed=vals)
np.savez_compressed(os.path.join(HERE, "m01_survival_cov_energy.npz"), meta=json.dumps(meta), x_ref=x_ref.numpy(),
                    x_gen=x_gen.numpy(), std_norm=std_norm.numpy(), cov_ref=torch.cov(x_ref.T).numpy(),
                    cov_gen=torch.cov(x_gen[:n_ref].T).numpy(), **out)
print(json.dumps(meta, indent=1))
