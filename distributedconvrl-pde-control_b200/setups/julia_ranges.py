"""Julia Float64 range semantics needed to rebuild the reference's sensor bases.

`collect(a:s:b)` for Float64 does NOT always have floor((b-a)/s)+1 elements; the
reference's Gaussians depend on it (`dx-50dx : dx : Lx+50dx`
has 291 samples for KS22 and 340 for KS200).  This follows
Base.(:)(::Float64, ::Float64, ::Float64) of Julia 1.9 (base/twiceprecision.jl).
"""
from fractions import Fraction
from math import gcd

import numpy as np

_MAXINT_F32 = 1 << 24


def _rat(x):
    """Base.rat: continued-fraction approximation, |num|,|den| <= 2^24."""
    y = float(x)
    a, b, c, d = 1, 0, 0, 1
    while abs(y) <= _MAXINT_F32:
        f = int(y)
        y -= f
        a, c = f * a + c, a
        b, d = f * b + d, b
        if max(abs(a), abs(b)) > _MAXINT_F32:
            return c, d
        if b != 0 and a / b == x:
            break
        if y == 0.0:
            break
        y = 1.0 / y
    return a, b


def _exact(x):
    n, d = _rat(x)
    return (n, d) if d != 0 and n / d == x else None


def float_range(start, step, stop):
    start, step, stop = float(start), float(step), float(stop)
    rs, rt, re_ = _exact(start), _exact(step), _exact(stop)
    if rs and rt and re_:
        (sn, sd), (tn, td), (en, ed) = rs, rt, re_
        den = sd // gcd(sd, td) * td
        if den and abs(start * den) <= 2.0 ** 53 and abs(step * den) <= 2.0 ** 53:
            a, s = round(start * den), round(step * den)
            num, dd = den * en - ed * a, s * ed
            q = num // dd if (num >= 0) == (dd > 0) else -((-num) // dd)
            n = max(0, q) + 1
            inside = lambda lo, x, hi: lo <= x <= hi or hi <= x <= lo
            if inside(start, start + (n - 1) * step, stop + step / 2) and not inside(start, start + n * step, stop):
                return np.array([float(Fraction(a + i * s, den)) for i in range(n)])
    lf = (stop - start) / step
    if lf < 0:
        n = 0
    elif lf == 0:
        n = 1
    else:
        n = int(round(lf)) + 1
        last = start + (n - 1) * step
        n -= int(start < stop < last) + int(start > stop > last)
    return start + step * np.arange(n, dtype=np.float64)
