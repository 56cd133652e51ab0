"""Coefficients of the packed-half polynomial branch of the tensor-core sampler's Swish epilogue (csrc/sampler_tc.cu).

Swish(z) = z sigmoid(z) = h + h tanh(h), h = z/2 = 2 max(h,0) - e(|h|) with e(u) = u (1 - tanh u), a bump that is
< 1.2e-3 beyond u = 4.5.  e is fitted on [0, U] by a degree-DEG Chebyshev interpolant, converted to the monomial basis in
x = 2u/U - 1 in [-1,1] (coefficients O(1): safe for fp16 Horner), and checked exhaustively over every fp16 input.

    python tools/swish_poly_fit.py          # prints the __half2 constants and the error report
"""
import numpy as np
from numpy.polynomial import chebyshev as C

U, DEG = 4.5, 9
f16 = np.float16


def fit():
    xs = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000)
    u = (xs + 1) / 2 * U
    return C.cheb2poly(C.chebfit(xs, u * (1 - np.tanh(u)), DEG))


def fma16(a, b, c):
    return (a.astype(np.float32) * b.astype(np.float32) + c.astype(np.float32)).astype(f16)


def emulate(h, mono):
    u = np.minimum(np.abs(h), f16(U))
    x = fma16(u, np.full_like(u, f16(2 / U)), np.full_like(u, f16(-1)))
    p = np.full_like(x, f16(mono[-1]))
    for c in mono[-2::-1]:
        p = fma16(p, x, np.full_like(x, f16(c)))
    return fma16(np.maximum(h, f16(0)), np.full_like(h, f16(2)), -p)


if __name__ == "__main__":
    mono = fit()
    h = np.arange(65536, dtype=np.uint16).view(f16)
    h = h[np.isfinite(h)]
    h = h[np.abs(h.astype(np.float64)) <= 24]
    t = h.astype(np.float64) * (1 + np.tanh(h.astype(np.float64)))
    err = np.abs(emulate(h, mono).astype(np.float64) - t)
    rnd = np.abs(t.astype(f16).astype(np.float64) - t)
    print(f"// U = {U}, degree {DEG}: max |err| = {err.max():.2e} over all fp16 inputs with |h| <= 24; max excess over "
          f"the fp16 rounding of the exact value = {(err - rnd).max():.2e}; mean |err| for |h| < 4 = "
          f"{err[np.abs(h.astype(float)) < 4].mean():.2e}")
    for k, c in enumerate(mono):
        bits = int(np.array([c], dtype=f16).view(np.uint16)[0])
        print(f"//   c{k} = {c:+.8f}  fp16 0x{bits:04X}")
    print("constexpr uint32_t SWISH_POLY[] = {" + ", ".join(
        f"0x{int(np.array([c], dtype=f16).view(np.uint16)[0]) * 0x10001:08X}u" for c in mono) + "};")
    print(f"// 2/U = 0x{int(np.array([2 / U], dtype=f16).view(np.uint16)[0]) * 0x10001:08X}, "
          f"U = 0x{int(np.array([U], dtype=f16).view(np.uint16)[0]) * 0x10001:08X}")
