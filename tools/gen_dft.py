#!/usr/bin/env python3
"""Generate straight-line in-register DFT butterflies for the batched FFTs.

Output: distributedconvrl-pde-control_b200/csrc/dft_gen.cuh

Each generated function
    template <typename T, int S> __device__ void dft_R(T* xr, T* xi)
computes X[k] = sum_n x[n] * exp(S * 2*pi*i * n*k / R) in place, natural order
in and out (S = -1 forward, S = +1 inverse, unnormalised), with all twiddles as
literals.  Composite sizes use Cooley-Tukey inside the register file
(R = P*Q: Q transforms of size P on stride-Q inputs, twiddle, P transforms of
size Q); the output permutation costs nothing because it is only a renaming of
SSA temporaries.  Base cases 2, 3, 4, 5 use the symmetric (cos/sin) form.

The batched 1-D FFT of length N = N1*N2 in ks_step.cuh / fft2d.cuh is then
"DFT_N1 in registers -> twiddle -> transpose through shared memory -> DFT_N2 in
registers" (four-step FFT), so sizes 192=12x16, 240=15x16, 256=16x16,
384=16x24, 600=24x25 need R in {12, 15, 16, 24, 25}.
"""
import math
import sys
from pathlib import Path

SIZES = [2, 3, 4, 5, 6, 8, 10, 12, 15, 16, 20, 24, 25, 32]


class Emitter:
    def __init__(self):
        self.lines = []
        self.n = 0

    def new(self):
        self.n += 1
        return "v%d" % self.n

    def emit(self, s):
        self.lines.append("    " + s)

    def cplx(self, re_expr, im_expr):
        v = self.new()
        self.emit("const T %sr = %s, %si = %s;" % (v, re_expr, v, im_expr))
        return v


def lit(x):
    return "T(%.17g)" % x


def add(e, a, b):
    return e.cplx("%sr + %sr" % (a, b), "%si + %si" % (a, b))


def sub(e, a, b):
    return e.cplx("%sr - %sr" % (a, b), "%si - %si" % (a, b))


def mul_i_sigma(e, a):
    """a * (S*i): (ar + i ai) * (S i) = -S ai + i S ar."""
    return e.cplx("-(T(S) * %si)" % a, "T(S) * %sr" % a)


def twiddle(e, a, m, R):
    """a * exp(S * 2*pi*i * m / R)."""
    m %= R
    if m == 0:
        return a
    g = math.gcd(m, R)
    num, den = m // g, R // g
    if den == 2:                                  # -1
        return e.cplx("-%sr" % a, "-%si" % a)
    if den == 4:
        if num == 1:                              # S*i
            return mul_i_sigma(e, a)
        return e.cplx("T(S) * %si" % a, "-(T(S) * %sr)" % a)   # -S*i
    if den == 8:
        h = lit(math.sqrt(0.5))
        # (1 + S i)/sqrt2 etc.: exp(S*i*pi*num/4)
        cr = {1: 1, 3: -1, 5: -1, 7: 1}[num]
        ci = {1: 1, 3: 1, 5: -1, 7: -1}[num]      # times S
        # (ar + i ai)(cr + i S ci) h = h*(cr ar - S ci ai) + i h*(S ci ar + cr ai)
        re = "%s * (%s%sr - T(S) * %s%si)" % (h, "" if cr > 0 else "-", a, "" if ci > 0 else "-", a)
        im = "%s * (T(S) * %s%sr + %s%si)" % (h, "" if ci > 0 else "-", a, "" if cr > 0 else "-", a)
        return e.cplx(re, im)
    th = 2 * math.pi * num / den
    c, s = lit(math.cos(th)), lit(math.sin(th))
    return e.cplx("%sr * %s - %si * (T(S) * %s)" % (a, c, a, s),
                  "%sr * (T(S) * %s) + %si * %s" % (a, s, a, c))


def dft_prime(e, x):
    """Symmetric form for odd prime R (3, 5)."""
    R = len(x)
    h = (R - 1) // 2
    p = [None] + [add(e, x[j], x[R - j]) for j in range(1, h + 1)]
    m = [None] + [sub(e, x[j], x[R - j]) for j in range(1, h + 1)]
    out = [None] * R
    s0r = " + ".join(["%sr" % x[0]] + ["%sr" % p[j] for j in range(1, h + 1)])
    s0i = " + ".join(["%si" % x[0]] + ["%si" % p[j] for j in range(1, h + 1)])
    out[0] = e.cplx(s0r, s0i)
    for k in range(1, h + 1):
        ar = "%sr" % x[0]
        ai = "%si" % x[0]
        br = bi = None
        for j in range(1, h + 1):
            c = math.cos(2 * math.pi * j * k / R)
            s = math.sin(2 * math.pi * j * k / R)
            ar += " + %s * %sr" % (lit(c), p[j])
            ai += " + %s * %si" % (lit(c), p[j])
            tr = "%s * %sr" % (lit(s), m[j])
            ti = "%s * %si" % (lit(s), m[j])
            br = tr if br is None else br + " + " + tr
            bi = ti if bi is None else bi + " + " + ti
        A = e.cplx(ar, ai)
        B = e.cplx(br, bi)
        # X_k = A + S*i*B ; X_{R-k} = A - S*i*B ;  S*i*B = (-S Bi, S Br)
        out[k] = e.cplx("%sr - T(S) * %si" % (A, B), "%si + T(S) * %sr" % (A, B))
        out[R - k] = e.cplx("%sr + T(S) * %si" % (A, B), "%si - T(S) * %sr" % (A, B))
    return out


def dft(e, x):
    R = len(x)
    if R == 1:
        return list(x)
    if R == 2:
        return [add(e, x[0], x[1]), sub(e, x[0], x[1])]
    if R == 4:
        a = add(e, x[0], x[2])
        b = sub(e, x[0], x[2])
        c = add(e, x[1], x[3])
        d = mul_i_sigma(e, sub(e, x[1], x[3]))
        return [add(e, a, c), add(e, b, d), sub(e, a, c), sub(e, b, d)]
    if R in (3, 5):
        return dft_prime(e, x)
    if R % 4 == 0:
        P = 4
    else:
        P = next(f for f in (2, 3, 5) if R % f == 0)
    Q = R // P
    # n = Q*n1 + n2, k = k1 + P*k2
    Y = [[None] * P for _ in range(Q)]
    for n2 in range(Q):
        sub_in = [x[Q * n1 + n2] for n1 in range(P)]
        yk = dft(e, sub_in)
        for k1 in range(P):
            Y[n2][k1] = twiddle(e, yk[k1], n2 * k1, R)
    out = [None] * R
    for k1 in range(P):
        zk = dft(e, [Y[n2][k1] for n2 in range(Q)])
        for k2 in range(Q):
            out[k1 + P * k2] = zk[k2]
    return out


def gen(R):
    e = Emitter()
    x = []
    for n in range(R):
        v = e.new()
        e.emit("const T %sr = xr[%d], %si = xi[%d];" % (v, n, v, n))
        x.append(v)
    out = dft(e, x)
    for k in range(R):
        e.emit("xr[%d] = %sr; xi[%d] = %si;" % (k, out[k], k, out[k]))
    head = ("template <typename T, int S>\n"
            "__device__ __forceinline__ void dft_%d(T* __restrict__ xr, T* __restrict__ xi) {\n" % R)
    return head + "\n".join(e.lines) + "\n}\n"


def main():
    out = Path(__file__).resolve().parent.parent / "distributedconvrl-pde-control_b200" / "csrc" / "dft_gen.cuh"
    parts = ["// GENERATED by tools/gen_dft.py -- do not edit.\n"
             "// In-register DFT butterflies, X[k] = sum_n x[n] exp(S*2*pi*i*n*k/R).\n"
             "#pragma once\n\nnamespace pdeb200 {\n"]
    for R in SIZES:
        parts.append(gen(R))
    parts.append("template <int R, typename T, int S>\n"
                 "__device__ __forceinline__ void dft_r(T* __restrict__ xr, T* __restrict__ xi) {\n")
    for i, R in enumerate(SIZES):
        parts.append("    %sif constexpr (R == %d) dft_%d<T, S>(xr, xi);\n" % ("else " if i else "", R, R))
    parts.append("    else static_assert(R < 0, \"no generated DFT for this radix\");\n}\n")
    parts.append("}  // namespace pdeb200\n")
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text("".join(parts))
    print("wrote", out, sum(p.count("\n") for p in parts), "lines")


if __name__ == "__main__":
    sys.exit(main())
