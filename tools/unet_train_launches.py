"""One hand-written SSM training iteration (loss + backward) of a U-Net score net, for an ncu launch list:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/unet_train_launches.py [1d|2d]

Only the second iteration is inside cudaProfilerStart/Stop (the first packs nothing that is cached, but it pays allocator
warm-up)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import sdeflow_light_b200 as P  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "2d"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if which == "1d":
    d, B = 1000, 64
    net = P.UNet1D(d, premodule="NormalizeLogRadius").to(dev)
else:
    d, B = 1024, 32
    net = P.VorticityUNet(32, (1, 2, 4), 2, premodule="NormalizeLogRadius", in_space=32, attention_resolutions=(2, 4),
                          flatten_order="F").to(dev)
    with torch.no_grad():
        for p_ in net.parameters():
            if p_.dim() > 1 and float(p_.abs().sum()) == 0.0:
                p_.normal_(0, 0.02)
T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
base = P.MSGMsde(torch.randn(64, d), beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=False, norm_map="log",
                 num_steps_forward=4, device=dev, estim_cst_norm_dens_r_T=False)
gen = P.PluginReverseSDE(base, net, T, deviceReverseSDE=dev).to(dev)
gen.train()
y = (torch.randn(B, d) * 1.3).to(dev)
v = (torch.rand(B, d).ge(0.5).float() * 2 - 1).to(dev)
t = (torch.rand(B, 1) * 0.9 + 0.05).to(dev)
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    gen.zero_grad()
    gen.ssm_loss(t, y, y, v).mean().backward()
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", which)
