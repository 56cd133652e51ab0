"""Per-parameter gradient comparison: hand-written U-Net training path vs torch autograd through the library layers.
python tools/unet_train_check.py [1d|2d] [L or in_space] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import sdeflow_light_b200 as P  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "1d"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
DEV = "cuda:0"
torch.manual_seed(3)
if which == "1d":
    d = size
    net = P.UNet1D(d, premodule="NormalizeLogRadius").to(DEV)
else:
    d = size * size
    net = P.VorticityUNet(32, (1, 2, 4), 2, premodule="NormalizeLogRadius", in_space=size, attention_resolutions=(2, 4),
                          flatten_order="F").to(DEV)
    with torch.no_grad():
        for p_ in net.parameters():
            if p_.dim() > 1 and float(p_.abs().sum()) == 0.0:
                p_.normal_(0, 0.02)
T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
base = P.MSGMsde(torch.randn(64, d), beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=False, norm_map="log",
                 num_steps_forward=4, device=DEV, estim_cst_norm_dens_r_T=False)
gen = P.PluginReverseSDE(base, net, T, deviceReverseSDE=DEV).to(DEV)
gen.train()
y = (torch.randn(B, d) * 1.3).to(DEV)
v = (torch.rand(B, d).ge(0.5).float() * 2 - 1).to(DEV)
t = (torch.rand(B, 1) * 0.9 + 0.05).to(DEV)
res = {}
for own in (False, True):
    gen.unet_train_kernels = own
    gen.zero_grad()
    loss = gen.ssm_loss(t, y, y, v)
    loss.mean().backward()
    res[own] = (loss.detach().clone(), {k: (None if p.grad is None else p.grad.clone()) for k, p in net.named_parameters()})
print("loss rel", float((res[True][0] - res[False][0]).abs().max() / res[False][0].abs().max()))
for k in res[False][1]:
    a, b = res[True][1][k], res[False][1][k]
    if a is None or b is None:
        print(f"{k:45s} own={'None' if a is None else 'ok'} ref={'None' if b is None else 'ok'}")
        continue
    print(f"{k:45s} {tuple(a.shape)!s:20s} rel {float((a - b).abs().max() / b.abs().max().clamp_min(1e-20)):.2e}  |ref| {float(b.abs().max()):.2e}")
