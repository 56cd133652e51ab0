"""RK4 reverse sampling through the generic per-stage sampler with the U-Net score nets: eager loop vs one CUDA graph per step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import sdeflow_light_b200 as P  # noqa: E402
from sdeflow_light_b200 import generic_sampler as GS  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
for which, d, batches in (("unet1d", 1000, (16, 256)), ("unet2d", 1024, (16, 128))):
    net = (P.UNet1D(d, premodule="NormalizeLogRadius") if which == "unet1d" else
           P.VorticityUNet(32, (1, 2, 4), 2, premodule="NormalizeLogRadius", in_space=32, attention_resolutions=(2, 4),
                           flatten_order="F")).to(dev)
    with torch.no_grad():
        for p_ in net.parameters():
            if p_.dim() > 1 and float(p_.abs().sum()) == 0.0:
                p_.normal_(0, 0.02)
    T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
    base = P.MSGMsde(torch.randn(512, d), beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=False, norm_map="log",
                     num_steps_forward=16, device=dev, estim_cst_norm_dens_r_T=False)
    gen = P.PluginReverseSDE(base, net, T, deviceReverseSDE=dev).to(dev)
    for B in batches:
        x0 = gen.latent_sample(B, d)
        for graphed in (False, True):
            GS.STEP_GRAPH = graphed
            N = 32
            for _ in range(2):
                P.rk4_stratonovich_sampler(gen, x0, N, keep_all_samples=False, norm_correction=True, seed=1, device_out=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                P.rk4_stratonovich_sampler(gen, x0, N, keep_all_samples=False, norm_correction=True, seed=1, device_out=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"{which} d={d} batch={B} steps={N} {'step-graph' if graphed else 'eager     '}: {ms:8.1f} ms/call  "
                  f"{B * N / ms * 1e3:9.0f} particle-steps/s", flush=True)
