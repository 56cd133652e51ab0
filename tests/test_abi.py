"""CPU: the C-ABI library loads and exports every symbol include/msgm_b200.h declares; the Python mirror keeps the
reference's surface; the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

import sdeflow_light_b200 as P
from sdeflow_light_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    h = open(os.path.join(ROOT, "include", "msgm_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(msgm_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decl = _declared_symbols()
    assert len(decl) >= 7
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in msgm_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == decl
    assert lib.msgm_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes follow from the field lists in the header on LP64
    assert ctypes.sizeof(_lib.SdeDesc) == 40
    assert ctypes.sizeof(_lib.MlpDesc) == 8 + 8 * 8
    assert ctypes.sizeof(_lib.SampleArgs) == 32 + 8 * 8
    assert ctypes.sizeof(_lib.Conv1dDesc) == 6 * 8 + 11 * 4 + 4   # 11 int32 + tail padding to 8
    assert ctypes.sizeof(_lib.Conv2dDesc) == 10 * 8 + 11 * 4 + 4
    assert ctypes.sizeof(_lib.Conv2dTcDesc) == 8 * 8 + 11 * 4 + 4   # ... prologue, fast
    assert ctypes.sizeof(_lib.Conv1dTcDesc) == 6 * 8 + 9 * 4 + 4    # ... gelu, fast


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    h = ctypes.c_void_p()
    rc = _lib.lib().msgm_create(ctypes.byref(h), 0)
    assert rc == _lib.ERR_NO_DEVICE and b"no CPU fallback" in _lib.lib().msgm_last_error()
    T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
    base = P.SGMsde(T=T, device="cpu")
    gen = P.PluginReverseSDE(base, P.MLP(2), T)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.rk4_stratonovich_sampler(gen, torch.zeros(4, 2), 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            P.MLP(2)(torch.zeros(4, 2), torch.zeros(4))


def test_surface_matches_reference_signatures():
    """Argument names/defaults of the reference's public callables."""
    ref_args = ["sde", "x_0", "num_steps", "lmbd", "keep_all_samples", "samplesToKeep", "include_t0", "T_",
                "norm_correction"]
    for fn in (P.euler_maruyama_sampler, P.heun_sampler, P.rk4_stratonovich_sampler):
        sig = inspect.signature(fn)
        pos = [p.name for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]
        assert pos == ref_args
        assert sig.parameters["num_steps"].default == 1000 and sig.parameters["keep_all_samples"].default is True
        assert sig.parameters["T_"].default == -1 and sig.parameters["include_t0"].default is False
    assert list(inspect.signature(P.MSGMsde.__init__).parameters)[1:] == [
        "y0", "beta_min", "beta_max", "T", "t_epsilon", "denseTensor", "norm_sampler", "norm_map", "kernel",
        "plot_validate", "num_steps_forward", "device", "estim_cst_norm_dens_r_T"]
    assert list(inspect.signature(P.PluginReverseSDE.__init__).parameters)[1:] == [
        "base_sde", "drift_a", "T", "vtype", "debias", "ssm_intT", "deviceReverseSDE"]
    assert list(inspect.signature(P.MLP.__init__).parameters)[1:] == [
        "input_dim", "index_dim", "hidden_dim", "act", "premodule"]
    for m in ("mu", "ga_m_drift", "ga", "mu_Strato", "sigma", "ssm", "ssm_loss", "sample_txy", "sample_t",
              "elbo_random_t_slice", "latent_sample", "cond_latent_sample"):
        assert hasattr(P.PluginReverseSDE, m)
    for m in ("beta", "f", "f_strato", "g", "div_Sigma", "IJK", "sample", "sample_scheme", "sample_scheme_allt",
              "latent_sample", "cond_latent_sample", "log_latent_pdf", "to"):
        assert hasattr(P.MSGMsde, m)


def test_state_dict_keys_match_reference():
    T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
    torch.manual_seed(0)
    base = P.MSGMsde(torch.randn(64, 2), T=T, norm_map="log", estim_cst_norm_dens_r_T=False)
    gen = P.PluginReverseSDE(base, P.MLP(2, premodule="NormalizeLogRadius"), T)
    keys = sorted(gen.state_dict().keys())
    assert keys == sorted(["T", "base_sde.T"] + [f"a.main.{i}.{w}" for i in (0, 2, 4, 6) for w in ("weight", "bias")])


def test_coefficient_methods_agree_with_oracle_on_cpu_tensors():
    """The thin coefficient methods are plain tensor expressions; check them against the oracle's."""
    from oracle import msgm_oracle as O
    torch.manual_seed(1)
    y0 = torch.randn(128, 4)
    T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
    for dense in (True, False):
        torch.manual_seed(2)
        base = P.MSGMsde(y0, T=T, denseTensor=dense, norm_map="log", estim_cst_norm_dens_r_T=False)
        torch.manual_seed(2)
        sde = O.make_msgm(y0, dense=dense, num_steps_forward=100)
        if dense:
            assert torch.equal(base.G, sde.G)  # same RNG consumption order as the reference's new_G
        s, y = torch.rand(16, 1), torch.randn(16, 4)
        assert torch.allclose(base.f(s, y), O.coef_f(sde, s, y))
        assert torch.allclose(base.div_Sigma(s, y), O.coef_div_sigma(sde, s, y))
        assert torch.allclose(base.g(s, y, sparse=not dense), O.coef_g(sde, s, y, sparse=not dense))


def test_tensor_core_layer_shape_queries_need_no_gpu():
    """The shape / size queries of the tensor-core layer entry points are pure host functions: which layers of the
    reference's U-Nets (NNUnet1D.py:81-102, model/unet.py:40-250) the tcgen05 path takes, and the packed-weight sizes."""
    L = _lib.lib()
    # 2-D: (Cout, Cin, K) -> bytes = Cout Cin K^2 x (hi + lo) x 2 B
    assert L.msgm_conv2d_tc_pack_bytes(128, 256, 3) == 128 * 256 * 9 * 4
    assert L.msgm_conv2d_tc_pack_bytes(192, 64, 1) == 192 * 64 * 4
    assert L.msgm_conv2d_tc_pack_bytes(1, 32, 3) == -1 and L.msgm_conv2d_tc_pack_bytes(32, 1, 3) == -1  # first / last conv
    assert L.msgm_conv1d_tc_pack_bytes(64, 64, 4) == 64 * 64 * 4 * 4
    assert L.msgm_conv1d_tc_pack_bytes(32, 32, 5) == -1
    # attention: the reference's two shapes (C=64, T=256) and (C=128, T=64) are covered; larger ones fall back
    assert L.msgm_attention_tc_supported(64, 256) == 1 and L.msgm_attention_tc_supported(128, 64) == 1
    assert L.msgm_attention_tc_supported(128, 256) == 0 and L.msgm_attention_tc_supported(64, 1024) == 0
    from sdeflow_light_b200.model.unet import _tc_shape_ok
    assert _tc_shape_ok(64, 64, 32, 3, 1, 16, 16) and _tc_shape_ok(32, 32, 0, 3, 2, 32, 32)
    assert not _tc_shape_ok(32, 1, 0, 3, 1, 32, 32) and not _tc_shape_ok(1, 32, 0, 3, 1, 32, 32)
    assert not _tc_shape_ok(32, 32, 0, 3, 2, 31, 32) and not _tc_shape_ok(64, 40, 24, 1, 1, 8, 8)


def test_weight_gradient_shape_query_and_structs_need_no_gpu():
    """Which convolutions of the U-Net training path get their weight gradient from the tcgen05 kernel (csrc/conv_wgrad_tc.cu) is a
    pure host decision; the step-clock and product-group descriptors have the sizes the header's field lists give on LP64."""
    L = _lib.lib()
    ok = L.msgm_conv_wgrad_tc_ok  # (N, Cout, C1, C2, KH, KW, stride, pad, up, Hs, Ws)
    assert ok(64, 128, 128, 0, 3, 3, 1, 1, 1, 16, 16) == 1          # ResBlock 3x3
    assert ok(64, 64, 64, 32, 3, 3, 1, 1, 1, 32, 32) == 1           # decoder concat
    assert ok(64, 64, 64, 0, 3, 3, 1, 1, 2, 8, 8) == 1              # Upsample conv (nearest x2 input)
    assert ok(64, 192, 64, 0, 1, 1, 1, 0, 1, 16, 16) == 1           # qkv 1x1
    assert ok(128, 64, 32, 32, 1, 3, 1, 1, 1, 1, 1000) == 1         # 1-D k3 on a concat
    assert ok(128, 32, 32, 0, 1, 4, 2, 1, 1, 1, 1000) == 1          # 1-D k4 stride 2 (and ConvTranspose1d, roles swapped)
    assert ok(64, 64, 64, 0, 3, 3, 2, 1, 1, 32, 32) == 1            # Downsample 3x3 stride 2
    assert ok(128, 32, 1, 0, 1, 3, 1, 1, 1, 1, 1000) == 0           # first conv: one input channel -> CUDA-core kernel
    assert ok(128, 32, 24, 8, 1, 3, 1, 1, 1, 1, 1000) == 0          # concat boundary not a multiple of 16
    assert ok(64, 64, 64, 0, 3, 3, 2, 1, 1, 31, 32) == 0            # odd size at a stride-2 conv
    assert ok(64, 64, 64, 0, 5, 5, 1, 2, 1, 16, 16) == 0
    assert ctypes.sizeof(_lib.StepClock) == 6 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.GemmProblem) == 5 * 8 + 5 * 8 + 5 * 4 + 4 * 4 + 2 * 4 + 4 + 4 + 4  # 136 with tail padding


def test_training_path_host_policies():
    """Host-side policies of the U-Net training path: which states the one-launch prologue covers, and the process-wide
    library precision switch (fp32 unless a net opts into TF32)."""
    from sdeflow_light_b200 import SDEs
    from sdeflow_light_b200.NNUnet import set_library_precision
    T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
    sgm = P.SGMsde(T=T, device="cpu")

    class _Sparse:
        sparseTensor = True

    class _Dense:
        sparseTensor = False

    assert SDEs._prepare_dim_ok(sgm, 2) and SDEs._prepare_dim_ok(sgm, 1024) and not SDEs._prepare_dim_ok(sgm, 5000)
    assert SDEs._prepare_dim_ok(_Sparse(), 1000) and SDEs._prepare_dim_ok(_Dense(), 32) and not SDEs._prepare_dim_ok(_Dense(), 33)
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        set_library_precision(False)
        assert not torch.backends.cudnn.allow_tf32 and not torch.backends.cuda.matmul.allow_tf32
        set_library_precision(True)
        assert torch.backends.cudnn.allow_tf32 and torch.backends.cuda.matmul.allow_tf32
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def test_plane_format_and_late_round2_queries_need_no_gpu():
    """Host-only facts of the 1-D U-Net's activation-plane path (include/msgm_b200.h, csrc/conv1d_tcp.cu): buffer size of
    the plane format, descriptor layouts, which nets take the path, the fused attention projection's shape query."""
    L = _lib.lib()
    # planes of (B, C, L): fp16 [hi | lo][C / 8][2 * 640 + B (L + 3)][8]
    for B, Cc, Ln in ((1, 8, 1), (256, 32, 1000), (1024, 128, 125), (0, 16, 7)):
        assert L.msgm_planes_bytes(B, Cc, Ln) == 2 * (Cc // 8) * (2 * 640 + B * (Ln + 3)) * 16
    assert L.msgm_planes_bytes(4, 12, 10) == -1 and L.msgm_planes_bytes(4, 16, 0) == -1   # C % 8, L >= 1
    assert ctypes.sizeof(_lib.Conv1dTcpDesc) == 7 * 8 + 10 * 4            # 7 pointers, 10 int32
    assert ctypes.sizeof(_lib.EmbFoldMultiDesc) == 2 * 16 * 8 + 4 * 16 * 4 + 3 * 4 + 4 + 8   # arrays, n / Cemb / B (+ pad), emb
    # attention + projection in one launch: C must be one N tile of the packed 1x1 image (32, 64, 128)
    assert L.msgm_attention_proj_tc_supported(64, 256) == 1 and L.msgm_attention_proj_tc_supported(128, 64) == 1
    assert L.msgm_attention_proj_tc_supported(96, 192) == 0 and L.msgm_attention_proj_tc_supported(128, 256) == 0
    # which UNet1D configurations run on planes: every width a multiple of 32, first width <= 128, one real input channel
    assert P.UNet1D(1000)._planes_ok(1000) and P.UNet1D(64, premodule="NormalizeLogRadius")._planes_ok(64)
    assert not P.UNet1D(64, base_channels=8, channel_mults=(1, 2))._planes_ok(64)      # widths 8, 16
    assert not P.UNet1D(64, base_channels=48)._planes_ok(64)                           # widths 48, 96, 192
    assert not P.UNet1D(4)._planes_ok(4)                                               # shorter than 2^levels
