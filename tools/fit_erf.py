"""Derive the coefficients of shx_erff (simplehydrology_b200/csrc/shx_math.cuh) and
measure its error against scipy's double erf with every operation rounded to fp32.

The path computes erf(0.4*discharge) once per particle step (reference
cellpool.h:242-244 calls libm erf).  The kernels use their own erf so that the CUDA
path and the CPU lock-step oracle can be compared bit for bit (CUDA's erff and
glibc's erff differ in the last ulp); this script documents where its constants
come from and how far it is from the exact function.
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy import special

f32 = np.float32
XA = 0.875


def cheb_to_mono(ch):
    return ch.convert(kind=P.Polynomial).convert(domain=[-1, 1]).coef


def fit():
    n = 4000
    u = np.cos(np.pi * (np.arange(n) + 0.5) / n) * 0.5 + 0.5
    t = u * XA * XA
    x = np.sqrt(t)
    a = cheb_to_mono(C.Chebyshev.fit(t, special.erf(x) / x - 1.0, 6, domain=[0, XA * XA]))
    lo, hi = 1 / 5.0, 1 / (XA + 1.0)
    z = lo + u * (hi - lo)
    b = cheb_to_mono(C.Chebyshev.fit(z, special.erfcx(1 / z - 1), 8, domain=[lo, hi]))
    return a.astype(f32), b.astype(f32)


EXP_C = np.array([1.0, 1.0, 1 / 2, 1 / 6, 1 / 24, 1 / 120, 1 / 720], dtype=f32)
LOG2E = f32(1.4426950408889634)
LN2_HI = f32(0.693145751953125)       # 12 significant bits: k*LN2_HI is exact for |k|<2^12
LN2_LO = f32(1.42860682030941723212e-6)


def expf_neg(y):
    """exp(y) for y in [-17, 0], fp32 op by op."""
    y = y.astype(f32)
    k = np.rint(y * LOG2E).astype(f32)
    r = (y - k * LN2_HI).astype(f32)
    r = (r - k * LN2_LO).astype(f32)
    p = np.full_like(r, EXP_C[6])
    for c in EXP_C[5::-1]:
        p = (p * r).astype(f32)
        p = (p + c).astype(f32)
    scale = ((k.astype(np.int32) + 127) << 23).view(f32)
    return (p * scale).astype(f32)


def erff(x, a, b):
    x = x.astype(f32)
    ax = np.abs(x)
    t = (ax * ax).astype(f32)
    # region A
    pa = np.full_like(t, a[6])
    for c in a[5::-1]:
        pa = (pa * t).astype(f32)
        pa = (pa + c).astype(f32)
    ra = (ax + (ax * pa).astype(f32)).astype(f32)
    # region B
    z = (f32(1.0) / (ax + f32(1.0)).astype(f32)).astype(f32)
    pb = np.full_like(z, b[8])
    for c in b[7::-1]:
        pb = (pb * z).astype(f32)
        pb = (pb + c).astype(f32)
    e = expf_neg(-np.minimum(t, f32(17.0)))
    rb = (f32(1.0) - (e * pb).astype(f32)).astype(f32)
    r = np.where(ax < f32(XA), ra, np.where(ax < f32(4.0), rb, f32(1.0)))
    return np.copysign(r, x).astype(f32)


if __name__ == "__main__":
    a, b = fit()
    xs = np.concatenate([np.linspace(0, 4.5, 2_000_001), np.logspace(-30, 0, 100001)]).astype(f32)
    got = erff(xs, a, b).astype(np.float64)
    want = special.erf(xs.astype(np.float64))
    ulp = np.spacing(np.abs(want).astype(f32)).astype(np.float64)
    err = np.abs(got - want) / ulp
    print("max ulp error", err.max(), "at x =", xs[err.argmax()])
    print("max abs error", np.abs(got - want).max())
    for name, arr in (("A", a), ("B", b)):
        print(name, ", ".join(f"{float(c).hex()}f" for c in arr))
        print(name, ", ".join(f"{c:.9e}f" for c in arr))
