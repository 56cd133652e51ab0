"""Build sdeflow_light_b200 objects (the product under test) from golden fixtures or oracle structures."""
import json
import os

import torch

import sdeflow_light_b200 as P

REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def report(**kw):
    """Append an observed-error record for the profiles/ summaries (only when gpurun_out/ exists)."""
    if os.path.isdir(REPORT):
        with open(os.path.join(REPORT, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(kw) + "\n")


def T_param(T):
    return torch.nn.Parameter(torch.FloatTensor([T]), requires_grad=False)


def base_from(meta, arr, device, n_fwd=None):
    kind, d = meta["kind"], meta["dim"]
    T = T_param(meta["T"])
    kw = dict(beta_min=meta["beta_min"], beta_max=meta["beta_max"], T=T, t_epsilon=meta.get("t_epsilon", 1e-3),
              num_steps_forward=n_fwd or meta.get("num_steps_forward", 16), device=device)
    if kind == "sgm":
        base = P.SGMsde(**kw)
    else:
        base = P.MSGMsde(torch.randn(8, d), denseTensor=(kind == "msgm_dense"), norm_sampler="ecdf", norm_map="log",
                         estim_cst_norm_dens_r_T=False, **kw)
        if "r_T" in arr:
            base.r_T = arr["r_T"].to(device)
        if kind == "msgm_dense":
            base.G = arr["G"].to(device)
            base.L_G = arr["L_G"].to(device)
    return base, T


def net_from(meta, arr, device):
    net = P.MLP(input_dim=meta["dim"], index_dim=1, hidden_dim=128,
                premodule="NormalizeLogRadius" if meta["premodule"] else None)
    with torch.no_grad():
        for i, l in enumerate(net.linears()):
            l.weight.copy_(arr[f"W{i}"])
            l.bias.copy_(arr[f"b{i}"])
    return net.to(device)


def gen_from(meta, arr, device, n_fwd=None):
    base, T = base_from(meta, arr, device, n_fwd)
    net = net_from(meta, arr, device)
    gen = P.PluginReverseSDE(base, net, T, vtype="rademacher", debias=False, ssm_intT=False,
                             deviceReverseSDE=device).to(device)
    return base, net, gen


def from_oracle(sde, mlp, device):
    """Package objects holding the same tensors as an oracle (OSde, OMlp) pair."""
    meta = dict(kind=sde.kind, dim=sde.dim, beta_min=sde.beta_min, beta_max=sde.beta_max, T=sde.T,
                t_epsilon=sde.t_epsilon, num_steps_forward=sde.num_steps_forward,
                premodule=bool(mlp.premodule) if mlp is not None else False)
    arr = {}
    if sde.kind == "msgm_dense":
        arr.update(G=sde.G, L_G=sde.L_G)
    if sde.r_T is not None:
        arr["r_T"] = sde.r_T
    if mlp is None:
        base, T = base_from(meta, arr, device)
        return base, None, P.forward_SDE(base, T.to(device))
    for i in range(4):
        arr[f"W{i}"], arr[f"b{i}"] = mlp.W[i], mlp.b[i]
    return gen_from(meta, arr, device)


SAMPLERS = {"em": P.euler_maruyama_sampler, "heun": P.heun_sampler, "rk4": P.rk4_stratonovich_sampler}
