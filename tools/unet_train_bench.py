"""SSM training iteration of the U-Net score nets (BASELINE configs 3 and 4): eager loop vs train.GraphedSsmStep, on the
hand-written kernel path (sdeflow_light_b200/unet_train.py, where it covers the net) and on the library path (torch autograd
through cuDNN: fp32 = reference parity, and the opt-in TF32 policy `net.train_tf32 = True`)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import sdeflow_light_b200 as P  # noqa: E402
from sdeflow_light_b200.train import GraphedSsmStep  # noqa: E402

dev = torch.device("cuda", 0)
for which, d, B, nfwd in (("unet1d", 1000, 64, 16), ("unet2d", 1024, 32, 128)):
    for mode in ("kernels", "library-fp32", "library-tf32"):
        tf32 = mode == "library-tf32"
        torch.manual_seed(0)
        data = torch.randn(512, d)
        T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
        base = P.MSGMsde(data, beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=False, norm_map="log",
                         num_steps_forward=nfwd, device=dev, estim_cst_norm_dens_r_T=False)
        net = (P.UNet1D(d, premodule="NormalizeLogRadius") if which == "unet1d" else
               P.VorticityUNet(32, (1, 2, 4), 2, premodule="NormalizeLogRadius", in_space=32, attention_resolutions=(2, 4),
                               flatten_order="F")).to(dev)
        net.train_tf32 = tf32
        gen = P.PluginReverseSDE(base, net, T, deviceReverseSDE=dev).to(dev)
        gen.unet_train_kernels = mode == "kernels"
        from sdeflow_light_b200 import unet_train
        if mode == "kernels" and not unet_train.supported(gen, data[:B].to(dev)):
            print(f"{which}: hand-written training path not built for this net", flush=True)
            continue
        xs = data[:B].to(dev)
        opt = torch.optim.Adam(gen.parameters(), lr=1e-4)
        gen.train()

        def eager():
            opt.zero_grad()
            gen.ssm(xs).mean().backward()
            opt.step()

        for _ in range(2):
            eager()
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(5):
            eager()
        torch.cuda.synchronize()
        ms_e = (time.time() - t0) / 5 * 1e3
        step = GraphedSsmStep(gen, (B, d), lr=1e-4)
        for _ in range(3):
            step(xs)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(10):
            step(xs)
        torch.cuda.synchronize()
        ms_g = (time.time() - t0) / 10 * 1e3
        print(f"{which} d={d} batch={B} N_fwd={nfwd} path={mode}: eager {ms_e:.1f} ms/iter, "
              f"graphed {ms_g:.1f} ms/iter ({B / ms_g * 1e3:.0f} samples/s)", flush=True)
