"""UNet1D (config 3: L = 1000, base 32, emb 128) forward and RK4 sampling throughput: hand-written kernels vs torch's
fp32 library path for the same module."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import sdeflow_light_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
L = 1000
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
net = P.UNet1D(L, premodule="NormalizeLogRadius").to(dev)
x, t = torch.randn(B, L, device=dev), torch.rand(B, device=dev)


def timeit(fn, reps=10):
    for _ in range(3):
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
    net.conv_mode = "fp32"
    ms_f, n_f = timeit(lambda: net(x, t))
    y_f = net(x, t)
    net.conv_mode = "tc"
    ms_k, n_k = timeit(lambda: net(x, t))
    y_k = net(x, t)
    net.conv_mode = "tc16"
    ms_16, _ = timeit(lambda: net(x, t))
    print(f"tc16 (single fp16 product) kernels: {ms_16:.2f} ms; rel diff vs tc {float((net(x, t) - y_k).abs().max() / y_k.abs().max()):.2e}")
    net.conv_mode = "tc"
    print(f"tc vs fp32 kernels: rel diff {float((y_k - y_f).abs().max() / y_f.abs().max()):.2e}; fp32 CUDA-core kernels "
          f"{ms_f:.2f} ms ({n_f:.0f} launches)")
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        ms_t, _ = timeit(lambda: net._forward(x, t))
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=True):
        ms_tf32, _ = timeit(lambda: net._forward(x, t))
gflop = 0.445 * B
print(f"UNet1D forward B={B} L={L}: kernels {ms_k:.2f} ms ({B / ms_k * 1e3:.0f} samples/s, {n_k:.0f} launches, "
      f"{gflop / ms_k:.1f} TFLOP/s of the reference's 0.445 GFLOP/sample) | torch fp32 {ms_t:.2f} ms | torch tf32 {ms_tf32:.2f} ms")

sig = torch.sin(torch.linspace(0, 6.28, L)[None] * torch.randint(1, 4, (512, 1))) + 0.1 * torch.randn(512, L)
T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
base = P.MSGMsde(sig, T=T, denseTensor=False, norm_map="log", num_steps_forward=16, device=dev, estim_cst_norm_dens_r_T=False)
gen = P.PluginReverseSDE(base, net, T, deviceReverseSDE=dev).to(dev)
x0 = gen.latent_sample(B, L)
N = 4
ms_s, n_s = timeit(lambda: P.rk4_stratonovich_sampler(gen, x0, N, keep_all_samples=False, norm_correction=True, seed=1,
                                                      device_out=True), reps=3)
print(f"RK4 sampling B={B} N={N}: {ms_s:.1f} ms per call = {B * N / ms_s * 1e3:.0f} particle-steps/s, {n_s:.0f} own launches/call")
