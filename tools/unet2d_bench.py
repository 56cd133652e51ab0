"""VorticityUNet (config 4: 32x32, base 32, (1,2,4), 2 res blocks, attention at 16x16 / 8x8) forward and RK4 sampling
throughput: hand-written kernels vs torch's fp32 library path for the same module."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import sdeflow_light_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
S = 32
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
net = P.VorticityUNet(32, (1, 2, 4), 2, premodule="NormalizeLogRadius", in_space=S, attention_resolutions=(2, 4),
                      flatten_order="F").to(dev)
with torch.no_grad():
    for k, p_ in net.named_parameters():
        if p_.abs().sum() == 0 and p_.dim() > 1:
            p_.normal_(0, 0.02)
x, t = torch.randn(B, S * S, device=dev), torch.rand(B, device=dev)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = P._lib.launch_count(dev)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (P._lib.launch_count(dev) - l0) / reps


with torch.no_grad():
    net.core.conv_mode = "fp32"
    net.cuda_graph = False
    ms_f, n_f = timeit(lambda: net(x, t))
    y_f = net(x, t)
    net.core.conv_mode = "tc"
    net.cuda_graph = False
    ms_e, n_e = timeit(lambda: net(x, t))
    print(f"tc kernels, eager launches: {ms_e:.2f} ms ({n_e:.0f} launches)")
    net.cuda_graph = True
    ms_k, n_k = timeit(lambda: net(x, t))
    y_k = net(x, t)
    net.core.conv_mode = "tc16"
    ms_16, _ = timeit(lambda: net(x, t))
    print(f"tc16 (single fp16 product) kernels: {ms_16:.2f} ms; rel diff vs tc {float((net(x, t) - y_k).abs().max() / y_k.abs().max()):.2e}")
    net.core.conv_mode = "tc"
    print(f"tc vs fp32 kernels: rel diff {float((y_k - y_f).abs().max() / y_f.abs().max()):.2e}; fp32 CUDA-core kernels "
          f"{ms_f:.2f} ms ({n_f:.0f} launches)")
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        torch.backends.cuda.matmul.allow_tf32 = False
        ms_t, _ = timeit(lambda: net._forward(x, t))
print(f"VorticityUNet forward B={B} {S}x{S}: kernels {ms_k:.2f} ms ({B / ms_k * 1e3:.0f} samples/s, {n_k:.0f} launches, "
      f"{1.204 * B / ms_k:.1f} TFLOP/s of the reference's 1.204 GFLOP/sample) | torch fp32 {ms_t:.2f} ms ({B / ms_t * 1e3:.0f} samples/s)")

# RK4 reverse sampling (config 4: sparse multiplicative SDE, d = 1024) through the generic per-stage sampler
img = torch.nn.functional.avg_pool2d(torch.randn(512, 1, S + 4, S + 4), 5, stride=1).reshape(512, S * S) * 4.0
T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
base = P.MSGMsde(img, beta_min=0.8, beta_max=160., T=T, t_epsilon=8e-3, denseTensor=False, norm_map="log",
                 num_steps_forward=128, device=dev, estim_cst_norm_dens_r_T=False)
gen = P.PluginReverseSDE(base, net, T, deviceReverseSDE=dev).to(dev)
x0 = gen.latent_sample(B, S * S)
N = 4
ms_s, n_s = timeit(lambda: P.rk4_stratonovich_sampler(gen, x0, N, keep_all_samples=False, norm_correction=True, seed=1,
                                                      device_out=True), reps=3)
print(f"RK4 sampling B={B} N={N}: {ms_s:.1f} ms per call = {B * N / ms_s * 1e3:.0f} particle-steps/s "
      f"({ms_s / (4 * N):.2f} ms per stage)")
