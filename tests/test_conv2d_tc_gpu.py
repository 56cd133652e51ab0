"""GPU parity of the tcgen05 convolution (csrc/conv2d_tc.cu) that carries the 2-D U-Net's 3x3 / 1x1 convs
(model/unet.py:40-250 in the reference): every fused feature (GroupNorm + SiLU prologue, channel concat, nearest x2
upsampling, stride 2, bias + embedding term + residual) against the same layer evaluated by torch in float64.
The operands are split fp16 hi + lo (three tensor-core products), so the stated tolerance is fp32-level: 2e-5 relative
to max|ref| (observed ~1e-6)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from sdeflow_light_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run_tc(x1, x2, W, bias, ebias, res, gn, silu, up, stride, fast=0):
    dev = torch.device(DEV)
    h, L = _lib.ctx(dev), _lib.lib()
    B, C1, Hs, Ws = x1.shape
    C2 = 0 if x2 is None else x2.shape[1]
    Cout, K = W.shape[0], W.shape[-1]
    nbytes = L.msgm_conv2d_tc_pack_bytes(Cout, C1 + C2, K)
    assert nbytes == Cout * (C1 + C2) * K * K * 4
    img = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    _lib.check(L.msgm_conv2d_tc_pack(h, _lib.ptr(W), Cout, C1 + C2, K, _lib.ptr(img), _lib.stream_ptr(dev)))
    ss = None
    if gn is not None:
        G, gamma, beta = gn
        ss = torch.empty((B, C1 + C2, 2), device=dev, dtype=torch.float32)
        _lib.check(L.msgm_gn_scale_shift(h, _lib.ptr(x1), C1, _lib.ptr(x2), C2, Hs * Ws, G, B, _lib.ptr(gamma),
                                         _lib.ptr(beta), _lib.ptr(ss), _lib.stream_ptr(dev)))
    pad = K // 2
    Ho, Wo = (Hs * up + 2 * pad - K) // stride + 1, (Ws * up + 2 * pad - K) // stride + 1
    out = torch.full((B, Cout, Ho, Wo), float("nan"), device=dev, dtype=torch.float32)
    p = lambda t_: None if t_ is None else t_.data_ptr()  # noqa: E731
    d = _lib.Conv2dTcDesc(p(x1), p(x2), p(img), p(bias), p(ebias), p(res), p(ss), p(out), B, C1, C2, Cout, K, stride, up,
                          Hs, Ws, 0 if gn is None else (2 if silu else 1), fast)
    _lib.check(L.msgm_conv2d_tc(h, C.byref(d), _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    flag = C.c_int32(0)
    _lib.check(L.msgm_debug_flags(h, C.byref(flag)))
    assert flag.value == 0, "a tcgen05 wait timed out"
    return out


def _ref(x1, x2, W, bias, ebias, res, gn, silu, up, stride):
    x = x1 if x2 is None else torch.cat([x1, x2], 1)
    x = x.double()
    if gn is not None:
        G, gamma, beta = gn
        x = F.group_norm(x, G, gamma.double(), beta.double(), eps=1e-5)
        if silu:
            x = x * torch.sigmoid(x)
    if up == 2:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    y = F.conv2d(x, W.double(), None if bias is None else bias.double(), stride=stride, padding=W.shape[-1] // 2)
    if ebias is not None:
        y = y + ebias.double()[:, :, None, None]
    if res is not None:
        y = y + res.double()
    return y


CASES = [
    # B, C1, C2, Cout, K, H, W, up, stride, gn(0 none,1 norm,2 norm+silu), bias, ebias, res
    (3, 32, 0, 32, 3, 32, 32, 1, 1, 2, 1, 1, 0),     # ResBlock in_layers at 32x32
    (2, 64, 32, 64, 3, 16, 16, 1, 1, 2, 1, 1, 0),    # decoder ResBlock over the concat [h, skip]
    (2, 128, 128, 128, 3, 8, 8, 1, 1, 2, 1, 0, 1),   # widest contraction (256 -> 128), residual epilogue
    (5, 64, 0, 192, 1, 16, 16, 1, 1, 1, 1, 0, 0),    # attention qkv: GroupNorm without SiLU, 1x1, 3 N tiles
    (4, 128, 0, 128, 1, 8, 8, 1, 1, 0, 1, 0, 1),     # attention proj_out + residual
    (3, 64, 32, 64, 1, 16, 16, 1, 1, 0, 1, 0, 0),    # ResBlock 1x1 skip over a concat
    (2, 64, 0, 64, 3, 8, 8, 2, 1, 0, 1, 0, 0),       # Upsample: nearest x2 folded into the staging
    (3, 32, 0, 32, 3, 32, 32, 1, 2, 0, 1, 0, 0),     # Downsample: stride 2
    (1, 16, 0, 32, 3, 5, 7, 1, 1, 2, 0, 0, 0),       # ragged: odd sizes, one 16-channel chunk, no bias
    (130, 32, 0, 64, 3, 8, 8, 1, 1, 2, 1, 1, 1),     # many images per tile, every epilogue term
]


@pytest.mark.parametrize("fast", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_conv2d_tc_matches_float64(case, fast):
    """fast = 0: three split fp16 products, tolerance 2e-5 (fp32 level).  fast = 1 (`conv_mode="tc16"`): one fp16 product,
    stated tolerance 3e-3 relative to max|ref| (fp16 operand rounding 2^-11 per factor; observed ~5e-4)."""
    B, C1, C2, Cout, K, H, W_, up, stride, gnm, hb, he, hr = case
    torch.manual_seed(B * 1000 + Cout + K)
    dev = DEV
    x1 = torch.randn(B, C1, H, W_, device=dev) * 1.7 + 0.3
    x2 = torch.randn(B, C2, H, W_, device=dev) * 0.6 if C2 else None
    Wt = torch.randn(Cout, C1 + C2, K, K, device=dev) / ((C1 + C2) * K * K) ** 0.5
    bias = torch.randn(Cout, device=dev) if hb else None
    ebias = torch.randn(B, Cout, device=dev) if he else None
    gn = None
    if gnm:
        G = min(32, C1 + C2)
        gn = (G, torch.rand(C1 + C2, device=dev) + 0.5, torch.randn(C1 + C2, device=dev) * 0.2)
    pad = K // 2
    Ho, Wo = (H * up + 2 * pad - K) // stride + 1, (W_ * up + 2 * pad - K) // stride + 1
    res = torch.randn(B, Cout, Ho, Wo, device=dev) if hr else None
    got = _run_tc(x1, x2, Wt, bias, ebias, res, gn, gnm == 2, up, stride, fast)
    ref = _ref(x1, x2, Wt, bias, ebias, res, gn, gnm == 2, up, stride)
    assert got.shape == ref.shape
    assert torch.isfinite(got).all(), "positions left unwritten"
    err = float((got.double() - ref).abs().max()) / float(ref.abs().max())
    assert err < (3e-3 if fast else 2e-5), f"conv2d_tc rel err {err:.3e}"


def test_conv2d_tc_rejects_unsupported_shapes():
    dev = torch.device(DEV)
    L = _lib.lib()
    assert L.msgm_conv2d_tc_pack_bytes(33, 32, 3) == -1     # Cout % 32
    assert L.msgm_conv2d_tc_pack_bytes(32, 24, 3) == -1     # Cin % 16
    assert L.msgm_conv2d_tc_pack_bytes(32, 32, 5) == -1     # kernel size
    x = torch.zeros(1, 32, 4, 4, device=dev)
    d = _lib.Conv2dTcDesc(x.data_ptr(), None, x.data_ptr(), None, None, None, None, x.data_ptr(), 1, 32, 0, 32, 3, 3, 1, 4,
                          4, 0)
    with pytest.raises(ValueError):
        _lib.check(L.msgm_conv2d_tc(_lib.ctx(dev), C.byref(d), _lib.stream_ptr(dev)))


@pytest.mark.parametrize("B,Cc,T", [(5, 64, 256), (3, 128, 64), (2, 32, 128), (4, 128, 128), (2, 64, 64)])
def test_attention_with_fused_projection_matches_float64(B, Cc, T):
    """AttentionBlock tail in one launch (model/unet.py:228-234): x + proj_out(attention(qkv)) vs float64."""
    dev = torch.device(DEV)
    h, L = _lib.ctx(dev), _lib.lib()
    assert L.msgm_attention_proj_tc_supported(Cc, T) == 1
    assert L.msgm_attention_proj_tc_supported(96, 192) == 0  # the 1x1 image of 96 channels is three N tiles
    torch.manual_seed(T + Cc + 1)
    qkv = torch.randn(B, 3 * Cc, T, device=dev) * 1.5
    W = torch.randn(Cc, Cc, 1, 1, device=dev) / Cc ** 0.5
    bias, x = torch.randn(Cc, device=dev), torch.randn(B, Cc, T, device=dev)
    img = torch.empty(L.msgm_conv2d_tc_pack_bytes(Cc, Cc, 1), device=dev, dtype=torch.uint8)
    _lib.check(L.msgm_conv2d_tc_pack(h, _lib.ptr(W), Cc, Cc, 1, _lib.ptr(img), _lib.stream_ptr(dev)))
    out = torch.full((B, Cc, T), float("nan"), device=dev)
    _lib.check(L.msgm_attention_proj_tc(h, _lib.ptr(qkv), _lib.ptr(img), _lib.ptr(bias), _lib.ptr(x), _lib.ptr(out), B, Cc, T,
                                        _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    assert _lib.debug_flags(dev) == 0
    q, k, v = qkv.double().split(Cc, dim=1)
    w = torch.softmax(torch.einsum("bct,bcs->bts", q, k) / Cc ** 0.5, dim=-1)
    att = torch.einsum("bts,bcs->bct", w, v)
    ref = torch.einsum("oc,bct->bot", W.double()[:, :, 0, 0], att) + bias.double()[None, :, None] + x.double()
    err = float((out.double() - ref).abs().max()) / float(ref.abs().max())
    assert torch.isfinite(out).all() and err < 2e-5, f"attention + projection rel err {err:.3e}"
    # bias / residual are optional
    _lib.check(L.msgm_attention_proj_tc(h, _lib.ptr(qkv), _lib.ptr(img), None, None, _lib.ptr(out), B, Cc, T, _lib.stream_ptr(dev)))
    ref0 = torch.einsum("oc,bct->bot", W.double()[:, :, 0, 0], att)
    assert float((out.double() - ref0).abs().max()) / float(ref0.abs().max()) < 2e-5


@pytest.mark.parametrize("B,Cc,T", [(5, 64, 256), (3, 128, 64), (2, 32, 128), (2, 96, 192)])
def test_attention_tc_matches_float64(B, Cc, T):
    """QKVAttention (model/unet.py:236-250) on the tensor pipe vs float64: softmax((q s)^T (k s)) v, s = C^-1/4."""
    dev = torch.device(DEV)
    h, L = _lib.ctx(dev), _lib.lib()
    assert L.msgm_attention_tc_supported(Cc, T) == 1
    torch.manual_seed(T + Cc)
    qkv = torch.randn(B, 3 * Cc, T, device=dev) * 1.5
    out = torch.full((B, Cc, T), float("nan"), device=dev)
    _lib.check(L.msgm_attention_tc(h, _lib.ptr(qkv), _lib.ptr(out), B, Cc, T, _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    flag = C.c_int32(0)
    _lib.check(L.msgm_debug_flags(h, C.byref(flag)))
    assert flag.value == 0
    q, k, v = qkv.double().split(Cc, dim=1)
    w = torch.softmax(torch.einsum("bct,bcs->bts", q, k) / Cc ** 0.5, dim=-1)
    ref = torch.einsum("bts,bcs->bct", w, v)
    err = float((out.double() - ref).abs().max()) / float(ref.abs().max())
    assert torch.isfinite(out).all() and err < 2e-5, f"attention_tc rel err {err:.3e}"
    assert L.msgm_attention_tc_supported(128, 512) == 0 and L.msgm_attention_tc_supported(40, 64) == 0


# ---- 1-D form (NNUnet1D.py: ConvBlock1D k3 convs, the k4 stride-2 down convs, folded embedding channels, GELU) ----------
CASES_1D = [
    # B, C1, C2, Cemb, Cout, K, stride, L, gelu
    (3, 32, 0, 0, 32, 3, 1, 1000, 1),      # ConvBlock1D second conv
    (2, 32, 32, 128, 32, 3, 1, 1000, 1),   # decoder block over [up, skip] + 128 folded embedding channels
    (4, 128, 128, 128, 128, 3, 1, 125, 1),  # widest decoder contraction, odd length
    (3, 64, 0, 0, 64, 4, 2, 500, 0),       # down-sampling conv
    (3, 128, 0, 0, 128, 4, 2, 125, 0),     # down-sampling conv, odd length
    (2, 16, 0, 8, 32, 3, 1, 7, 1),         # ragged: shorter than one tile, embedding table at both borders
    (2, 32, 0, 0, 64, 1, 1, 300, 0),       # pointwise
]


@pytest.mark.parametrize("fast", [0, 1])
@pytest.mark.parametrize("case", CASES_1D)
def test_conv1d_tc_matches_float64(case, fast):
    B, C1, C2, Cemb, Cout, K, stride, Lin, gelu = case
    dev = torch.device(DEV)
    h, L = _lib.ctx(dev), _lib.lib()
    torch.manual_seed(Lin + Cout + K)
    x1 = torch.randn(B, C1, Lin, device=dev) * 1.3
    x2 = torch.randn(B, C2, Lin, device=dev) if C2 else None
    emb = torch.randn(B, Cemb, device=dev) if Cemb else None
    Cw = C1 + C2 + Cemb
    W = torch.randn(Cout, Cw, K, device=dev) / (Cw * K) ** 0.5
    bias = torch.randn(Cout, device=dev)
    pad = 0 if K == 1 else 1
    Lout = (Lin + 2 * pad - K) // stride + 1
    E = None
    if Cemb:
        E = torch.empty((B, Cout, K), device=dev)
        _lib.check(L.msgm_emb_fold(h, _lib.ptr(W), _lib.ptr(emb), _lib.ptr(E), Cw, C1 + C2, Cemb, Cout, K, B,
                                   _lib.stream_ptr(dev)))
    img = torch.empty(L.msgm_conv1d_tc_pack_bytes(Cout, C1 + C2, K), device=dev, dtype=torch.uint8)
    _lib.check(L.msgm_conv1d_tc_pack(h, _lib.ptr(W), Cout, Cw, C1 + C2, K, _lib.ptr(img), _lib.stream_ptr(dev)))
    out = torch.full((B, Cout, Lout), float("nan"), device=dev)
    d = _lib.Conv1dTcDesc(x1.data_ptr(), None if x2 is None else x2.data_ptr(), img.data_ptr(), bias.data_ptr(),
                          None if E is None else E.data_ptr(), out.data_ptr(), B, C1, C2, Cout, K, stride, Lin, gelu, fast)
    _lib.check(L.msgm_conv1d_tc(h, C.byref(d), _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    feats = [x1] + ([x2] if C2 else []) + ([emb[:, :, None].expand(-1, -1, Lin)] if Cemb else [])
    ref = F.conv1d(torch.cat(feats, 1).double(), W.double(), bias.double(), stride=stride, padding=pad)
    if gelu:
        ref = F.gelu(ref)
    assert out.shape == ref.shape and torch.isfinite(out).all()
    err = float((out.double() - ref).abs().max()) / float(ref.abs().max())
    assert err < (3e-3 if fast else 2e-5), f"conv1d_tc rel err {err:.3e}"


@pytest.mark.parametrize("B,Cin,Cout,Lin,Lout", [(3, 128, 128, 125, 250), (2, 128, 64, 250, 500), (4, 64, 32, 62, 125),
                                                 (2, 16, 16, 3, 7)])
@pytest.mark.parametrize("fast", [0, 1])
def test_convt1d_tc_matches_float64(B, Cin, Cout, Lin, Lout, fast):
    """nn.ConvTranspose1d(k4, s2, p1) + right zero padding (NNUnet1D.py:98,165-169) as a 3-tap tensor-core conv."""
    dev = torch.device(DEV)
    h, L = _lib.ctx(dev), _lib.lib()
    torch.manual_seed(Lin + Cout)
    x = torch.randn(B, Cin, Lin, device=dev)
    W = torch.randn(Cin, Cout, 4, device=dev) / (Cin * 2) ** 0.5
    bias = torch.randn(Cout, device=dev)
    img = torch.empty(24 * Cin * Cout, device=dev, dtype=torch.uint8)
    _lib.check(L.msgm_convt1d_tc_pack(h, _lib.ptr(W), Cout, Cin, _lib.ptr(img), _lib.stream_ptr(dev)))
    out = torch.zeros((B, Cout, Lout), device=dev)
    out[:, :, :2 * Lin] = float("nan")
    _lib.check(L.msgm_convt1d_tc(h, _lib.ptr(x), _lib.ptr(img), _lib.ptr(bias), _lib.ptr(out), B, Cin, Cout, Lin, Lout,
                                 fast, _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    ref = F.pad(F.conv_transpose1d(x.double(), W.double(), bias.double(), stride=2, padding=1), (0, Lout - 2 * Lin))
    assert torch.isfinite(out).all()
    err = float((out.double() - ref).abs().max()) / float(ref.abs().max())
    assert err < (3e-3 if fast else 2e-5), f"convt1d_tc rel err {err:.3e}"


def test_tensor_core_kernels_write_inside_their_outputs():
    """Guard bands: every tensor-core kernel writes its output view and nothing around it (ragged shapes, where tiles hang
    over the end of the position space)."""
    dev = torch.device(DEV)
    h, L = _lib.ctx(dev), _lib.lib()
    torch.manual_seed(0)
    PAD, CAN = 4096, 12345.0

    def banded(shape):
        n = 1
        for s_ in shape:
            n *= s_
        buf = torch.full((n + 2 * PAD,), CAN, device=dev)
        return buf, buf[PAD:PAD + n].view(shape)

    def intact(buf, n):
        return bool((buf[:PAD] == CAN).all()) and bool((buf[PAD + n:] == CAN).all())

    # 2-D conv, odd sizes, one sample
    B, Cin, Cout, H, W_ = 3, 16, 32, 5, 7
    x = torch.randn(B, Cin, H, W_, device=dev)
    Wt = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.1
    img = torch.empty(L.msgm_conv2d_tc_pack_bytes(Cout, Cin, 3), device=dev, dtype=torch.uint8)
    _lib.check(L.msgm_conv2d_tc_pack(h, _lib.ptr(Wt), Cout, Cin, 3, _lib.ptr(img), _lib.stream_ptr(dev)))
    buf, out = banded((B, Cout, H, W_))
    d = _lib.Conv2dTcDesc(x.data_ptr(), None, img.data_ptr(), None, None, None, None, out.data_ptr(), B, Cin, 0, Cout, 3, 1, 1,
                          H, W_, 0, 0)
    _lib.check(L.msgm_conv2d_tc(h, C.byref(d), _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), Wt.double(), padding=1)
    assert intact(buf, out.numel()) and float((out.double() - ref).abs().max()) < 1e-4
    # transposed 1-D conv with right padding
    Bc, Ci, Co, Lin, Lout = 2, 16, 16, 9, 21
    xt = torch.randn(Bc, Ci, Lin, device=dev)
    Wc = torch.randn(Ci, Co, 4, device=dev) * 0.2
    bias = torch.randn(Co, device=dev)
    imgt = torch.empty(24 * Ci * Co, device=dev, dtype=torch.uint8)
    _lib.check(L.msgm_convt1d_tc_pack(h, _lib.ptr(Wc), Co, Ci, _lib.ptr(imgt), _lib.stream_ptr(dev)))
    buf, out = banded((Bc, Co, Lout))
    out.zero_()
    _lib.check(L.msgm_convt1d_tc(h, _lib.ptr(xt), _lib.ptr(imgt), _lib.ptr(bias), _lib.ptr(out), Bc, Ci, Co, Lin, Lout, 0,
                                 _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    ref = F.pad(F.conv_transpose1d(xt.double(), Wc.double(), bias.double(), stride=2, padding=1), (0, Lout - 2 * Lin))
    assert intact(buf, out.numel()) and float((out.double() - ref).abs().max()) < 1e-4
    # attention, T = 64 (half of the 128-query tile is padding)
    qkv = torch.randn(3, 192, 64, device=dev)
    buf, out = banded((3, 64, 64))
    _lib.check(L.msgm_attention_tc(h, _lib.ptr(qkv), _lib.ptr(out), 3, 64, 64, _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    assert intact(buf, out.numel()) and torch.isfinite(out).all()


@pytest.mark.parametrize("N,Cout,C1,C2,KH,KW,up,Hs,Ws,cscale,stride", [
    (6, 32, 32, 0, 3, 3, 1, 12, 12, 1.0, 1),     # the 32-channel full-resolution layers (M padded from 32 to 128 channels)
    (4, 64, 64, 32, 3, 3, 1, 8, 8, 1e-7, 1),     # decoder concat; cotangents far below the fp16 range (range scaling)
    (3, 128, 256, 0, 3, 3, 1, 8, 8, 3e-5, 1),    # two input-channel tiles of 128
    (5, 64, 64, 0, 3, 3, 2, 7, 5, 1.0, 1),       # Upsample conv: the input is read through the nearest x2 index
    (4, 192, 64, 0, 1, 1, 1, 16, 16, 1e-3, 1),   # qkv 1x1 conv, two output-channel tiles
    (3, 128, 128, 64, 1, 1, 1, 9, 9, 1.0, 1),    # skip_connection 1x1 on a concat: 192 input channels -> 2 tiles of 96
    (7, 64, 32, 32, 1, 3, 1, 1, 125, 1.0, 1),    # 1-D k3 conv on a concat
    (2, 1, 32, 0, 1, 1, 1, 1, 1000, 1e-6, 1),    # 1-D final pointwise conv: one output channel
    (64, 32, 32, 0, 3, 3, 1, 32, 32, 1.0, 1),    # full config-4 size of one layer: 74 k padded positions, many slices
    (6, 64, 32, 0, 1, 4, 1, 1, 250, 1.0, 2),     # 1-D k4 stride-2 down conv: two phase windows (even / odd input columns)
    (5, 128, 64, 0, 1, 4, 1, 1, 126, 1e-5, 2),   # the same structure as the ConvTranspose1d weight gradient (roles swapped)
    (4, 64, 64, 0, 3, 3, 1, 16, 16, 1.0, 2),     # Downsample: 3x3 stride 2
    (3, 128, 128, 0, 3, 3, 1, 8, 6, 1e-4, 2),
    (12, 16, 16, 0, 1, 4, 1, 1, 32, 1.0, 2),     # tiny 1-D layers of the small fixture net
    (12, 32, 32, 0, 1, 4, 1, 1, 16, 1.0, 2),
    (12, 32, 16, 0, 1, 4, 1, 1, 16, 1.0, 2),
    (12, 16, 16, 0, 1, 3, 1, 1, 32, 1.0, 1),
])
def test_conv_wgrad_tc_matches_float64(N, Cout, C1, C2, KH, KW, up, Hs, Ws, cscale, stride):
    """csrc/conv_wgrad_tc.cu: the conv weight gradient as a tcgen05 product over positions (both operands MN-major from the
    staged tile layout, split fp16 x 3, power-of-two range scaling of the cotangent) against torch's float64 weight gradient;
    the result is accumulated into gW at a channel offset, and entries outside the written block must stay untouched."""
    torch.manual_seed(N * 1000 + Cout + C1 + KH)
    dev = torch.device(DEV)
    h, L = _lib.ctx(dev), _lib.lib()
    Cin, pad = C1 + C2, (1 if KW == 4 else KW // 2)
    assert L.msgm_conv_wgrad_tc_ok(N, Cout, C1, C2, KH, KW, stride, pad, up, Hs, Ws) == 1
    x1 = torch.randn(N, C1, Hs, Ws)
    x2 = torch.randn(N, C2, Hs, Ws) if C2 else None
    xin = x1 if x2 is None else torch.cat([x1, x2], 1)
    if up == 2:
        xin = F.interpolate(xin, scale_factor=2, mode="nearest")
    Wd = torch.zeros(Cout, Cin, KH, KW, dtype=torch.float64, requires_grad=True)
    out = F.conv2d(xin.double(), Wd, stride=(1 if KH == 1 else stride, stride), padding=(KH // 2, pad))
    cot = torch.randn(out.shape) * cscale
    (out * cot.double()).sum().backward()
    ref = Wd.grad
    coff, Cw = 3, Cin + 5                       # written block inside a wider weight tensor (the 1-D U-Net's folded channels)
    g0 = torch.randn(Cout, Cw, KH, KW) * float(ref.abs().max())  # what is already in the gradient buffer
    gW = g0.clone().to(dev)
    cd, x1d = cot.to(dev), x1.to(dev)
    x2d = None if x2 is None else x2.to(dev)
    amax = torch.empty(1, device=dev, dtype=torch.float32)
    _lib.check(L.msgm_amax(h, _lib.ptr(cd), cd.numel(), _lib.ptr(amax), _lib.stream_ptr(dev)))
    scratch = torch.empty(L.msgm_conv_wgrad_tc_scratch_bytes(h, N, Cout, Cin, KH, KW, stride, pad, up, Hs, Ws), device=dev, dtype=torch.uint8)
    _lib.check(L.msgm_conv_wgrad_tc(h, _lib.ptr(cd), _lib.ptr(x1d), _lib.ptr(x2d), _lib.ptr(gW), _lib.ptr(amax), None, _lib.ptr(scratch),
                                    N, Cout, C1, C2, Cw, coff, KH, KW, stride, pad, up, Hs, Ws, 1, _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    _lib.check_async(dev)
    got = gW.cpu()
    err = float((got[:, coff:coff + Cin].double() - g0[:, coff:coff + Cin].double() - ref).abs().max() / ref.abs().max())
    untouched = torch.equal(got[:, :coff], g0[:, :coff]) and torch.equal(got[:, coff + Cin:], g0[:, coff + Cin:])
    assert untouched
    print(f"wgrad_tc N={N} {C1}+{C2}->{Cout} {KH}x{KW} s{stride} up{up} {Hs}x{Ws} cscale={cscale:g}: rel err {err:.2e}")
    assert err < 2e-5, err


def test_conv_wgrad_tc_range_scales_the_input_operand():
    """ConvTranspose1d(k4, s2, p1) weight gradient = the stride-2 weight gradient with the two tensors' roles swapped: the kernel's
    "input" operand is then the cotangent (~1e-7 in the deep layers), so the power-of-two range scaling must act on that side
    (amax_in); without it the fp16 hi/lo split of the small operand loses its low part (observed 9e-3 on the fixture net)."""
    torch.manual_seed(77)
    dev = torch.device(DEV)
    h, L = _lib.ctx(dev), _lib.lib()
    N, Cx, Cg, Lin = 6, 32, 32, 40                      # x: (N, Cx, Lin) activations; g: (N, Cg, 2 Lin) cotangent of the up-sampled signal
    x = torch.randn(N, Cx, 1, Lin)
    g = torch.randn(N, Cg, 1, 2 * Lin) * 3e-7
    Wd = torch.zeros(Cx, Cg, 1, 4, dtype=torch.float64, requires_grad=True)
    out = F.conv2d(g.double(), Wd, stride=(1, 2), padding=(0, 1))     # (N, Cx, 1, Lin): gW[cx][cg][k] = sum x[cx][p] g[cg][2p - 1 + k]
    (out * x.double()).sum().backward()
    ref = Wd.grad
    xd, gd = x.to(dev), g.to(dev)
    amax = torch.empty(1, device=dev, dtype=torch.float32)
    _lib.check(L.msgm_amax(h, _lib.ptr(gd), gd.numel(), _lib.ptr(amax), _lib.stream_ptr(dev)))
    scratch = torch.empty(L.msgm_conv_wgrad_tc_scratch_bytes(h, N, Cx, Cg, 1, 4, 2, 1, 1, 1, 2 * Lin), device=dev, dtype=torch.uint8)
    gW = torch.empty(Cx, Cg, 1, 4, device=dev)
    _lib.check(L.msgm_conv_wgrad_tc(h, _lib.ptr(xd), _lib.ptr(gd), None, _lib.ptr(gW), None, _lib.ptr(amax), _lib.ptr(scratch),
                                    N, Cx, Cg, 0, Cg, 0, 1, 4, 2, 1, 1, 1, 2 * Lin, 0, _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    err = float((gW.cpu().double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err
