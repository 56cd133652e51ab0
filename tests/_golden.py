"""Load tests/golden/*.npz fixtures (written by tests/golden/make_golden.py from the live reference)."""
import glob
import json
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, prefix + "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    arr = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files if k != "meta"}
    return meta, arr


def oracle_objects(meta, arr):
    """Build the oracle's OSde / OMlp from a fixture."""
    from oracle import msgm_oracle as O
    kind, d = meta["kind"], meta["dim"]
    common = dict(beta_min=meta["beta_min"], beta_max=meta["beta_max"], T=meta["T"],
                  t_epsilon=meta.get("t_epsilon", 1e-3), num_steps_forward=meta.get("num_steps_forward", 16))
    if kind == "sgm":
        sde = O.OSde("sgm", d, **common)
    elif kind == "msgm_sparse":
        I, J, K, V = O.cyclic_sparse_G(d)
        sde = O.OSde(kind, d, G_I=I, G_J=J, G_K=K, G_V=V, L_G=0.5 * torch.eye(d), r_T=arr.get("r_T"),
                     norm_map="log", **common)
    else:
        sde = O.OSde(kind, d, G=arr["G"], L_G=arr["L_G"], r_T=arr.get("r_T"), norm_map="log", **common)
    mlp = None
    if "W0" in arr:
        mlp = O.OMlp([arr[f"W{i}"] for i in range(4)], [arr[f"b{i}"] for i in range(4)], meta["premodule"], d)
    return sde, mlp
