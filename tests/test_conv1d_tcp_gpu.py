"""GPU parity of the TMA-fed 1-D convs on activation planes (csrc/conv1d_tcp.cu) against torch's float64 convolutions:
Conv1d k3 p1 / k4 s2 p1 / ConvTranspose1d k4 s2 p1 as the 1-D U-Net uses them (NNUnet1D.py:13-33,81-102,165-169), with the
concat read in place, the folded embedding table, bias and exact GELU, planes and fp32 outputs.  Tolerance 2e-5 of max|ref|
for the split fp16 x 3 products (fp32-level parity), 5e-3 for the single-product mode.  Also checked: the zero rows of an
output planes buffer (padding ring, guard rows) are never written.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from sdeflow_light_b200 import _lib
from tests import _build as Bd

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _planes(B, C_, L):
    n = _lib.lib().msgm_planes_bytes(B, C_, L)
    assert n > 0
    return torch.zeros(n, device=DEV, dtype=torch.uint8)


def _pack(x):
    B, C_, L = x.shape
    pl = _planes(B, C_, L)
    _lib.check(_lib.lib().msgm_planes_pack(_lib.ctx(DEV), _lib.ptr(x), _lib.ptr(pl), B, C_, L, _lib.stream_ptr(DEV)))
    return pl


def _unpack(pl, B, C_, L):
    x = torch.empty((B, C_, L), device=DEV, dtype=torch.float32)
    _lib.check(_lib.lib().msgm_planes_unpack(_lib.ctx(DEV), _lib.ptr(pl), _lib.ptr(x), B, C_, L, _lib.stream_ptr(DEV)))
    return x


def _ring_is_zero(pl, B, C_, L):
    """every row of the planes buffer that holds no position is zero"""
    rows = 2 * 640 + B * (L + 3)
    v = pl.view(torch.int16).view(2, C_ // 8, rows, 8)
    mask = torch.ones(rows, dtype=torch.bool, device=DEV)
    idx = (640 + torch.arange(B, device=DEV)[:, None] * (L + 3) + 1 + torch.arange(L, device=DEV)[None, :]).reshape(-1)
    mask[idx] = False
    return bool((v[:, :, mask, :] == 0).all())


def test_planes_roundtrip():
    torch.manual_seed(0)
    x = torch.randn(5, 24, 37, device=DEV) * 3.0
    pl = _pack(x)
    y = _unpack(pl, 5, 24, 37)
    assert float((x - y).abs().max()) <= 2e-6 * float(x.abs().max())
    assert _ring_is_zero(pl, 5, 24, 37)


CASES = [
    # kind, B, C1, C2, Cout, L, emb, gelu
    ("k3", 3, 32, 0, 32, 70, True, True),
    ("k3", 2, 64, 64, 64, 125, True, True),
    ("k3", 5, 128, 128, 128, 61, False, True),
    ("k3", 2, 32, 32, 32, 1000, True, True),
    ("k3", 4, 16, 0, 192, 33, False, False),
    ("k4", 3, 32, 0, 32, 70, False, False),
    ("k4", 2, 128, 0, 128, 251, False, False),
    ("k4", 3, 64, 0, 64, 1000, False, False),
    ("t", 3, 128, 0, 128, 31, False, False),
    ("t", 2, 64, 0, 32, 250, False, False),
    ("t", 2, 32, 0, 16, 125, False, False),
]


@pytest.mark.parametrize("kind,B,C1,C2,Cout,L,emb,gelu", CASES)
@pytest.mark.parametrize("fast", [0, 1])
def test_conv1d_tcp_matches_float64(kind, B, C1, C2, Cout, L, emb, gelu, fast):
    torch.manual_seed(1)
    Lb, h = _lib.lib(), _lib.ctx(DEV)
    st = _lib.stream_ptr(DEV)
    Cin = C1 + C2
    x1 = torch.randn(B, C1, L, device=DEV)
    x2 = torch.randn(B, C2, L, device=DEV) if C2 else None
    xin = x1 if x2 is None else torch.cat([x1, x2], 1)
    Cemb = 24 if emb else 0
    if kind == "t":
        W = torch.randn(Cin, Cout, 4, device=DEV) / (Cin * 2) ** 0.5
        bias = torch.randn(Cout, device=DEV)
        Lout = 2 * L + (3 if L % 2 else 0)  # odd case: the reference right-pads the up-sampled signal with zeros
        ref = F.conv_transpose1d(xin.double(), W.double(), bias.double(), stride=2, padding=1)
        ref = F.pad(ref, (0, Lout - ref.shape[-1]))
        img = torch.empty(24 * Cin * Cout, device=DEV, dtype=torch.uint8)
        _lib.check(Lb.msgm_convt1d_tc_pack(h, _lib.ptr(W), Cout, Cin, _lib.ptr(img), st))
        K, E = 3, None
    else:
        K = 3 if kind == "k3" else 4
        W = torch.randn(Cout, Cin + Cemb, K, device=DEV) / ((Cin + Cemb) * K) ** 0.5
        bias = torch.randn(Cout, device=DEV)
        evec = torch.randn(B, Cemb, device=DEV) if emb else None
        full = xin if not emb else torch.cat([xin, evec[:, :, None].expand(-1, -1, L)], 1)
        ref = F.conv1d(full.double(), W.double(), bias.double(), stride=1 if K == 3 else 2, padding=1)
        if gelu:
            ref = F.gelu(ref)
        Lout = ref.shape[-1]
        img = torch.empty(Lb.msgm_conv1d_tc_pack_bytes(Cout, Cin, K), device=DEV, dtype=torch.uint8)
        _lib.check(Lb.msgm_conv1d_tc_pack(h, _lib.ptr(W), Cout, Cin + Cemb, Cin, K, _lib.ptr(img), st))
        E = None
        if emb:
            E = torch.empty((B, Cout, K), device=DEV, dtype=torch.float32)
            _lib.check(Lb.msgm_emb_fold(h, _lib.ptr(W), _lib.ptr(evec), _lib.ptr(E), Cin + Cemb, Cin, Cemb, Cout, K, B, st))
    p1, p2 = _pack(x1), (None if x2 is None else _pack(x2))
    outp = _planes(B, Cout, Lout)
    outf = torch.zeros((B, Cout, Lout), device=DEV, dtype=torch.float32)
    d = _lib.Conv1dTcpDesc(p1.data_ptr(), None if p2 is None else p2.data_ptr(), img.data_ptr(), bias.data_ptr(),
                           None if E is None else E.data_ptr(), outp.data_ptr(), outf.data_ptr(), B, C1, C2, Cout, K, L, Lout,
                           int(gelu), int(kind == "t"), fast)
    _lib.check(Lb.msgm_conv1d_tcp(h, C.byref(d), st))
    torch.cuda.synchronize()
    assert _lib.debug_flags(DEV) == 0
    tol = 5e-3 if fast else 2e-5
    scale = float(ref.abs().max())
    e_f = float((outf.double() - ref).abs().max()) / scale
    got_p = _unpack(outp, B, Cout, Lout)
    e_p = float((got_p.double() - ref).abs().max()) / scale
    Bd.report(test=f"conv1d_tcp-{kind}-B{B}-C{C1}+{C2}->{Cout}-L{L}-fast{fast}", rel_f32=e_f, rel_planes=e_p)
    assert e_f <= tol and e_p <= tol, (e_f, e_p)
    assert float((got_p - outf).abs().max()) <= 4e-6 * scale  # the two outputs of one call are the same values
    assert _ring_is_zero(outp, B, Cout, Lout)


def test_first_conv_writes_planes():
    torch.manual_seed(2)
    Lb, h, st = _lib.lib(), _lib.ctx(DEV), _lib.stream_ptr(DEV)
    B, Cout, L, Cemb = 4, 32, 203, 16
    x = torch.randn(B, L, device=DEV)
    W = torch.randn(Cout, 1 + Cemb, 3, device=DEV) / 3.0
    bias = torch.randn(Cout, device=DEV)
    evec = torch.randn(B, Cemb, device=DEV)
    full = torch.cat([x[:, None, :], evec[:, :, None].expand(-1, -1, L)], 1)
    ref = F.gelu(F.conv1d(full.double(), W.double(), bias.double(), padding=1))
    E = torch.empty((B, Cout, 3), device=DEV, dtype=torch.float32)
    _lib.check(Lb.msgm_emb_fold(h, _lib.ptr(W), _lib.ptr(evec), _lib.ptr(E), 1 + Cemb, 1, Cemb, Cout, 3, B, st))
    pl = _planes(B, Cout, L)
    _lib.check(Lb.msgm_conv1d_first_planes(h, _lib.ptr(x), _lib.ptr(W), 1 + Cemb, _lib.ptr(bias), _lib.ptr(E), _lib.ptr(pl), B,
                                           Cout, L, 1, st))
    got = _unpack(pl, B, Cout, L)
    assert float((got.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    assert _ring_is_zero(pl, B, Cout, L)


def test_one_launch_embedding_kernels_are_bit_identical():
    """msgm_emb_fold_multi / msgm_embed_mlp2 against the per-layer / per-MLP launches they replace."""
    torch.manual_seed(3)
    Lb, h, st = _lib.lib(), _lib.ctx(DEV), _lib.stream_ptr(DEV)
    B, Cemb = 37, 128
    emb = torch.randn(B, Cemb, device=DEV)
    shapes = [(32, 1, 3), (64, 32, 3), (128, 64, 3), (128, 256, 3), (32, 64, 4)]
    D = _lib.EmbFoldMultiDesc()
    Ws, got, ref = [], [], []
    for i, (Cout, Cin, K) in enumerate(shapes):
        W = torch.randn(Cout, Cin + Cemb, K, device=DEV)
        E1 = torch.empty((B, Cout, K), device=DEV)
        E2 = torch.full((B, Cout, K), float("nan"), device=DEV)
        _lib.check(Lb.msgm_emb_fold(h, _lib.ptr(W), _lib.ptr(emb), _lib.ptr(E1), Cin + Cemb, Cin, Cemb, Cout, K, B, st))
        D.W[i], D.E[i], D.Cw[i], D.Coff[i], D.Cout[i], D.K[i] = W.data_ptr(), E2.data_ptr(), Cin + Cemb, Cin, Cout, K
        Ws.append(W); ref.append(E1); got.append(E2)
    D.n, D.Cemb, D.B, D.emb = len(shapes), Cemb, B, emb.data_ptr()
    _lib.check(Lb.msgm_emb_fold_multi(h, C.byref(D), st))
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    for E in (128, 96):
        t, u = torch.rand(B, device=DEV), torch.randn(B, device=DEV)
        pa = [torch.randn(E, 1, device=DEV), torch.randn(E, device=DEV), torch.randn(E, E, device=DEV) / E ** 0.5,
              torch.randn(E, device=DEV)]
        pb = [torch.randn(E, 1, device=DEV), torch.randn(E, device=DEV), torch.randn(E, E, device=DEV) / E ** 0.5,
              torch.randn(E, device=DEV)]
        for both in (False, True):
            o1, o2 = torch.empty(B, E, device=DEV), torch.empty(B, E, device=DEV)
            _lib.check(Lb.msgm_embed_mlp(h, _lib.ptr(t), *[_lib.ptr(p) for p in pa], _lib.ptr(o1), B, E, 0, st))
            if both:
                _lib.check(Lb.msgm_embed_mlp(h, _lib.ptr(u), *[_lib.ptr(p) for p in pb], _lib.ptr(o1), B, E, 1, st))
            _lib.check(Lb.msgm_embed_mlp2(h, _lib.ptr(t), *[_lib.ptr(p) for p in pa], _lib.ptr(u) if both else None,
                                          *[(_lib.ptr(p) if both else None) for p in pb], _lib.ptr(o2), B, E, st))
            assert torch.equal(o1, o2), float((o1 - o2).abs().max())


@pytest.mark.parametrize("kind,B,Cin,Cout,L", [("k3", 3, 64, 64, 301), ("k4", 5, 32, 64, 130), ("t", 2, 64, 32, 77)])
def test_conv1d_tcp_writes_stay_inside_its_outputs(kind, B, Cin, Cout, L):
    """Both outputs sit in the middle of larger allocations filled with a byte pattern: no byte outside them changes, and
    nothing but rows of real positions changes inside the planes (compute-sanitizer is not available on the GPU pool)."""
    torch.manual_seed(4)
    Lb, h, st = _lib.lib(), _lib.ctx(DEV), _lib.stream_ptr(DEV)
    x = torch.randn(B, Cin, L, device=DEV)
    bias = torch.randn(Cout, device=DEV)
    if kind == "t":
        W = torch.randn(Cin, Cout, 4, device=DEV)
        img = torch.empty(24 * Cin * Cout, device=DEV, dtype=torch.uint8)
        _lib.check(Lb.msgm_convt1d_tc_pack(h, _lib.ptr(W), Cout, Cin, _lib.ptr(img), st))
        K, Lout = 3, 2 * L + 5
    else:
        K = 3 if kind == "k3" else 4
        W = torch.randn(Cout, Cin, K, device=DEV)
        img = torch.empty(Lb.msgm_conv1d_tc_pack_bytes(Cout, Cin, K), device=DEV, dtype=torch.uint8)
        _lib.check(Lb.msgm_conv1d_tc_pack(h, _lib.ptr(W), Cout, Cin, Cin, K, _lib.ptr(img), st))
        Lout = L if K == 3 else (L - 2) // 2 + 1
    pad = 1 << 16
    nb = Lb.msgm_planes_bytes(B, Cout, Lout)
    big_p = torch.full((nb + 2 * pad,), 0xAB, device=DEV, dtype=torch.uint8)
    big_p[pad:pad + nb] = 0
    nf = B * Cout * Lout * 4
    big_f = torch.full((nf + 2 * pad,), 0xCD, device=DEV, dtype=torch.uint8)
    d = _lib.Conv1dTcpDesc(_pack(x).data_ptr(), None, img.data_ptr(), bias.data_ptr(), None, big_p.data_ptr() + pad,
                           big_f.data_ptr() + pad, B, Cin, 0, Cout, K, L, Lout, 1, int(kind == "t"), 0)
    _lib.check(Lb.msgm_conv1d_tcp(h, C.byref(d), st))
    torch.cuda.synchronize()
    assert _lib.debug_flags(DEV) == 0
    for big, n, pat in ((big_p, nb, 0xAB), (big_f, nf, 0xCD)):
        assert bool((big[:pad] == pat).all()) and bool((big[pad + n:] == pat).all())
    assert _ring_is_zero(big_p[pad:pad + nb], B, Cout, Lout)
    covered = 2 * L if kind == "t" else Lout  # a transposed conv leaves the reference's right padding untouched
    fo = big_f[pad:pad + nf].view(torch.float32).view(B, Cout, Lout)
    assert bool(torch.isfinite(fo[:, :, :covered]).all())
    assert bool((big_f[pad:pad + nf].view(B, Cout, Lout, 4)[:, :, covered:] == 0xCD).all())


def test_conv1d_tcp_full_size_properties():
    """BASELINE config 3 sizes (L = 1000, batch 256, 32 -> 32 and 128 + 128 -> 128 channels), where a float64 reference is too
    slow to be worth it: size-independent properties of a bias-free conv without activation -- linearity in the input,
    sample independence (a sample's output does not depend on its neighbours in the batch: tiles straddle samples), and
    agreement of the planes output with the fp32 output of the same call."""
    torch.manual_seed(5)
    Lb, h, st = _lib.lib(), _lib.ctx(DEV), _lib.stream_ptr(DEV)
    for (B, C1, C2, Cout, L) in [(256, 32, 0, 32, 1000), (256, 128, 128, 128, 250)]:
        Cin = C1 + C2
        W = torch.randn(Cout, Cin, 3, device=DEV) / (3 * Cin) ** 0.5
        img = torch.empty(Lb.msgm_conv1d_tc_pack_bytes(Cout, Cin, 3), device=DEV, dtype=torch.uint8)
        _lib.check(Lb.msgm_conv1d_tc_pack(h, _lib.ptr(W), Cout, Cin, Cin, 3, _lib.ptr(img), st))

        def conv(xa, xb, nb):
            outp = _planes(nb, Cout, L)
            outf = torch.empty((nb, Cout, L), device=DEV, dtype=torch.float32)
            d = _lib.Conv1dTcpDesc(_pack(xa).data_ptr(), None if xb is None else _pack(xb).data_ptr(), img.data_ptr(), None, None,
                                   outp.data_ptr(), outf.data_ptr(), nb, C1, C2, Cout, 3, L, 0, 0, 0, 0)
            _lib.check(Lb.msgm_conv1d_tcp(h, C.byref(d), st))
            return outf, outp

        xa, ya = torch.randn(B, C1, L, device=DEV), torch.randn(B, C1, L, device=DEV)
        xb = torch.randn(B, C2, L, device=DEV) if C2 else None
        yb = torch.randn(B, C2, L, device=DEV) if C2 else None
        fx, px = conv(xa, xb, B)
        fy, _ = conv(ya, yb, B)
        fz, _ = conv(2.0 * xa - 0.5 * ya, None if xb is None else 2.0 * xb - 0.5 * yb, B)
        scale = float(fx.abs().max())
        assert float((fz - (2.0 * fx - 0.5 * fy)).abs().max()) <= 2e-5 * scale           # linearity
        assert float((_unpack(px, B, Cout, L) - fx).abs().max()) <= 4e-6 * scale          # planes == fp32 output
        sub = slice(100, 117)                                                             # 17 samples out of the middle
        fs, _ = conv(xa[sub].contiguous(), None if xb is None else xb[sub].contiguous(), 17)
        assert torch.equal(fs, fx[sub])                                                   # sample independence, bit for bit
        assert _lib.debug_flags(DEV) == 0
