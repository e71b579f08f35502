#!/usr/bin/env python3
"""Straight-line FFT codelet generator (a tiny genfft) for the B200 STFT kernels.

Emits fully unrolled, constant-folded, CSE'd fp32 DFT codelets as C functions that
compile unchanged under nvcc (``__device__ __forceinline__``) and gcc (``static
inline``; used by the CPU unit tests to check every codelet against numpy.fft).

Codelets operate on register arrays with static indices only:

    B2A_CODELET void b2a_rdft16(const float (&x)[16], float (&yr)[9], float (&yi)[9]);   // real -> half spectrum
    B2A_CODELET void b2a_cdft25(const float (&xr)[25], const float (&xi)[25], float (&yr)[25], float (&yi)[25]);

The algorithm is recursive decimation-in-time Cooley-Tukey over the factors {4, 2, 5, 3}
with (a) real-input awareness at every recursion level (sub-DFTs of real sequences only
compute bins 0..m/2, the rest by conjugation), (b) algebraic simplification of trivial
twiddles (1, -1, +-i, (1+-i)/sqrt2) and (c) hash-consing so identical sub-expressions
are computed once.  Not derived from any reference source: the reference has no FFT
code at all (it calls MLX).

Usage:  python tools/gen_codelets.py > mlx_swift_audio_b200/csrc/codelets.h
"""
from __future__ import annotations

import math
import sys

# ------------------------------------------------------------------------------------
# symbolic real expressions with hash-consing
# ------------------------------------------------------------------------------------


class G:
    """Expression graph.  Nodes are ints; node table maps id -> tuple."""

    def __init__(self):
        self.nodes = []
        self.memo = {}
        self.ZERO = self._mk(("const", 0.0))

    def _mk(self, key):
        n = self.memo.get(key)
        if n is None:
            n = len(self.nodes)
            self.nodes.append(key)
            self.memo[key] = n
        return n

    def inp(self, name):
        return self._mk(("in", name))

    def const(self, v):
        if v == 0.0:
            v = 0.0
        return self._mk(("const", float(v)))

    def is_const(self, a):
        return self.nodes[a][0] == "const"

    def cval(self, a):
        return self.nodes[a][1]

    def neg(self, a):
        k = self.nodes[a]
        if k[0] == "const":
            return self.const(-k[1])
        if k[0] == "neg":
            return k[1]
        if k[0] == "mul":
            return self.mul(-k[1], k[2])
        if k[0] == "sub":
            return self.sub(k[2], k[1])
        return self._mk(("neg", a))

    def add(self, a, b):
        ka, kb = self.nodes[a], self.nodes[b]
        if ka[0] == "const" and kb[0] == "const":
            return self.const(ka[1] + kb[1])
        if ka[0] == "const" and ka[1] == 0.0:
            return b
        if kb[0] == "const" and kb[1] == 0.0:
            return a
        if kb[0] == "neg":
            return self.sub(a, kb[1])
        if ka[0] == "neg":
            return self.sub(b, ka[1])
        if a > b:
            a, b = b, a
        return self._mk(("add", a, b))

    def sub(self, a, b):
        ka, kb = self.nodes[a], self.nodes[b]
        if ka[0] == "const" and kb[0] == "const":
            return self.const(ka[1] - kb[1])
        if kb[0] == "const" and kb[1] == 0.0:
            return a
        if ka[0] == "const" and ka[1] == 0.0:
            return self.neg(b)
        if a == b:
            return self.ZERO
        if kb[0] == "neg":
            return self.add(a, kb[1])
        return self._mk(("sub", a, b))

    def mul(self, c, a):
        """constant * node"""
        c = float(c)
        k = self.nodes[a]
        if c == 0.0:
            return self.ZERO
        if k[0] == "const":
            return self.const(c * k[1])
        if c == 1.0:
            return a
        if c == -1.0:
            return self.neg(a)
        if k[0] == "neg":
            return self.mul(-c, k[1])
        if k[0] == "mul":
            return self.mul(c * k[1], k[2])
        return self._mk(("mul", c, a))


def _snap(v):
    """Snap twiddle components that are exactly representable special values."""
    for s in (0.0, 1.0, -1.0, 0.5, -0.5):
        if abs(v - s) < 1e-15:
            return s
    return v


class C:
    """Complex symbolic value."""
    __slots__ = ("g", "re", "im")

    def __init__(self, g, re, im):
        self.g, self.re, self.im = g, re, im

    def __add__(self, o):
        return C(self.g, self.g.add(self.re, o.re), self.g.add(self.im, o.im))

    def __sub__(self, o):
        return C(self.g, self.g.sub(self.re, o.re), self.g.sub(self.im, o.im))

    def conj(self):
        return C(self.g, self.re, self.g.neg(self.im))

    def mul_i(self):  # * (+i)
        return C(self.g, self.g.neg(self.im), self.re)

    def mul_neg_i(self):  # * (-i)
        return C(self.g, self.im, self.g.neg(self.re))

    def scale(self, c):
        return C(self.g, self.g.mul(c, self.re), self.g.mul(c, self.im))

    def mul_const(self, wr, wi):
        g = self.g
        wr, wi = _snap(wr), _snap(wi)
        if wi == 0.0:
            return self.scale(wr)
        if wr == 0.0:
            return C(g, g.mul(-wi, self.im), g.mul(wi, self.re))
        if abs(abs(wr) - abs(wi)) < 1e-15:
            # (a+ib) * c(sr + i si) with |sr|=|si|=1:  c*((sr a - si b) + i (si a + sr b))
            c = abs(wr)
            sr = 1.0 if wr > 0 else -1.0
            si = 1.0 if wi > 0 else -1.0
            re = g.sub(g.mul(sr, self.re), g.mul(si, self.im))
            im = g.add(g.mul(si, self.re), g.mul(sr, self.im))
            return C(g, g.mul(c, re), g.mul(c, im))
        re = g.sub(g.mul(wr, self.re), g.mul(wi, self.im))
        im = g.add(g.mul(wi, self.re), g.mul(wr, self.im))
        return C(g, re, im)


USE_PFA = True


def _is_real_seq(g, xs):
    return all(x.im == g.ZERO for x in xs)


def _factor(n):
    for r in (4, 2, 5, 3):
        if n % r == 0 and n > r:
            return r
    return None


def _butterfly(g, xs, sign):
    """Direct small DFT of len(xs) in {2,3,4,5} (or any odd prime) with pairing."""
    r = len(xs)
    if r == 1:
        return list(xs)
    if r == 2:
        return [xs[0] + xs[1], xs[0] - xs[1]]
    if r == 4:
        a, b = xs[0] + xs[2], xs[0] - xs[2]
        c, d = xs[1] + xs[3], xs[1] - xs[3]
        dj = d.mul_neg_i() if sign < 0 else d.mul_i()
        return [a + c, b + dj, a - c, b - dj]
    assert r % 2 == 1
    h = (r - 1) // 2
    s = [xs[j] + xs[r - j] for j in range(1, h + 1)]
    d = [xs[j] - xs[r - j] for j in range(1, h + 1)]
    y0 = xs[0]
    for t in s:
        y0 = y0 + t
    out = [None] * r
    out[0] = y0
    for k in range(1, h + 1):
        a = xs[0]
        b = None
        for j in range(1, h + 1):
            ang = 2.0 * math.pi * ((j * k) % r) / r
            a = a + s[j - 1].scale(_snap(math.cos(ang)))
            t = d[j - 1].scale(_snap(math.sin(ang)))
            b = t if b is None else b + t
        bi = b.mul_neg_i() if sign < 0 else b.mul_i()   # forward: -i * sum sin * d
        out[k] = a + bi
        out[r - k] = a - bi
    return out


def _coprime_split(n):
    """(r, m) with n = r*m, gcd(r, m) = 1 and r in {4, 5, 3} (a radix the butterflies handle directly), or None."""
    for r in (4, 5, 3):
        if n % r == 0 and n > r and math.gcd(r, n // r) == 1:
            return r, n // r
    return None


def _dft_pfa(g, xs, sign, r, m):
    """Good-Thomas prime-factor step: n = r*m with gcd(r, m) = 1 becomes an r x m two-dimensional DFT with NO twiddle
    factors.  Input map n = (m*n1 + r*n2) mod N, output map k = (m*(m^-1 mod r)*k1 + r*(r^-1 mod m)*k2) mod N."""
    n = r * m
    real = _is_real_seq(g, xs)
    rows = [dft(g, [xs[(m * n1 + r * n2) % n] for n2 in range(m)], sign) for n1 in range(r)]   # DFTs of size m over n2
    u = pow(m, -1, r)
    v = pow(r, -1, m)
    ys = [None] * n
    # real input: rows are Hermitian in k2, so the columns k2 > m/2 are conjugates: X[-k] = conj(X[k])
    k2max = m // 2 + 1 if real else m
    for k2 in range(k2max):
        b = _butterfly(g, [rows[n1][k2] for n1 in range(r)], sign)
        for k1 in range(r):
            ys[(m * u * k1 + r * v * k2) % n] = b[k1]
    if real:
        for k in range(n):
            if ys[k] is None:
                ys[k] = ys[(n - k) % n].conj()
    assert all(y is not None for y in ys)
    return ys


def dft(g, xs, sign=-1):
    """Symbolic DFT.  sign=-1 forward (e^{-2 pi i nk/N})."""
    n = len(xs)
    real = _is_real_seq(g, xs)
    r = _factor(n)
    cp = _coprime_split(n) if USE_PFA else None
    if r is None:
        ys = _butterfly(g, xs, sign)
    elif cp is not None:
        ys = _dft_pfa(g, xs, sign, cp[0], cp[1])
    else:
        m = n // r
        subs = [dft(g, xs[q::r], sign) for q in range(r)]
        ys = [None] * n
        # Real input: the sub-DFTs are Hermitian, so column m-k is the conjugate of column k rotated by one
        # radix step -- butterfly(m-k)[j] = conj(butterfly(k)[r-1-j]) -- and only columns 0..m/2 are computed.
        kmax = m // 2 + 1 if real else m
        for k in range(kmax):
            col = []
            for q in range(r):
                ang = sign * 2.0 * math.pi * ((q * k) % n) / n
                col.append(subs[q][k].mul_const(math.cos(ang), math.sin(ang)))
            b = _butterfly(g, col, sign)
            for j in range(r):
                ys[k + m * j] = b[j]
        for k in range(kmax, m):
            for j in range(r):
                ys[k + m * j] = ys[(m - k) + m * (r - 1 - j)].conj()
    if real:
        # enforce exact Hermitian structure so upstream CSE sees conj pairs as the same nodes
        for k in range(n // 2 + 1, n):
            ys[k] = ys[n - k].conj()
        ys[0] = C(g, ys[0].re, g.ZERO)
        if n % 2 == 0:
            ys[n // 2] = C(g, ys[n // 2].re, g.ZERO)
    return ys


# ------------------------------------------------------------------------------------
# emission
# ------------------------------------------------------------------------------------

def _fmt(v):
    s = repr(float(v))
    if "e" in s or "." in s or "inf" in s or "nan" in s:
        pass
    else:
        s += ".0"
    return s + "f"


def emit(g, name, args, outputs):
    """outputs: list of (lvalue_string, node)."""
    lines = []
    done = {}
    counter = [0]

    def ref(n):
        k = g.nodes[n]
        if k[0] == "const":
            return _fmt(k[1])
        if k[0] == "in":
            return k[1]
        return done[n]

    def visit(root):
        stack = [(root, False)]
        while stack:
            n, expanded = stack.pop()
            k = g.nodes[n]
            if k[0] in ("const", "in") or n in done:
                continue
            if not expanded:
                stack.append((n, True))
                ch = [k[1]] if k[0] == "neg" else ([k[2]] if k[0] == "mul" else [k[1], k[2]])
                for c in reversed(ch):
                    stack.append((c, False))
            else:
                v = "t%d" % counter[0]
                counter[0] += 1
                if k[0] == "add":
                    e = "%s + %s" % (ref(k[1]), ref(k[2]))
                elif k[0] == "sub":
                    e = "%s - %s" % (ref(k[1]), ref(k[2]))
                elif k[0] == "mul":
                    e = "%s * %s" % (_fmt(k[1]), ref(k[2]))
                elif k[0] == "neg":
                    e = "-%s" % ref(k[1])
                lines.append("  const float %s = %s;" % (v, e))
                done[n] = v

    for lv, n in outputs:
        visit(n)
        lines.append("  %s = %s;" % (lv, ref(n)))
    nadd = sum(1 for n in done if g.nodes[n][0] in ("add", "sub"))
    nmul = sum(1 for n in done if g.nodes[n][0] == "mul")
    nneg = sum(1 for n in done if g.nodes[n][0] == "neg")
    hdr = "// %s: %d add/sub, %d mul, %d neg\nB2A_CODELET void %s(%s) {" % (name, nadd, nmul, nneg, name, args)
    return hdr + "\n" + "\n".join(lines) + "\n}\n", (nadd, nmul, nneg)


def gen_rdft(n):
    """real forward DFT, bins 0..n/2."""
    g = G()
    xs = [C(g, g.inp("x[%d]" % i), g.ZERO) for i in range(n)]
    ys = dft(g, xs, -1)
    h = n // 2 + 1
    outs = []
    for k in range(h):
        outs.append(("yr[%d]" % k, ys[k].re))
        outs.append(("yi[%d]" % k, ys[k].im))
    return emit(g, "b2a_rdft%d" % n, "const float (&x)[%d], float (&yr)[%d], float (&yi)[%d]" % (n, h, h), outs)


def gen_cdft(n):
    g = G()
    xs = [C(g, g.inp("xr[%d]" % i), g.inp("xi[%d]" % i)) for i in range(n)]
    ys = dft(g, xs, -1)
    outs = []
    for k in range(n):
        outs.append(("yr[%d]" % k, ys[k].re))
        outs.append(("yi[%d]" % k, ys[k].im))
    return emit(g, "b2a_cdft%d" % n,
                "const float (&xr)[%d], const float (&xi)[%d], float (&yr)[%d], float (&yi)[%d]" % (n, n, n, n), outs)


def gen_c2r(n):
    """Unnormalised inverse real DFT: half spectrum (n/2+1 bins; imag of DC/Nyquist ignored)
    -> n real samples, x[t] = sum_k X[k] e^{+2 pi i tk/n} over the Hermitian extension.
    Uses the Hartley identity x[t] = Re F[t] + Im F[t], F = DFT(Re X + Im X)."""
    g = G()
    h = n // 2 + 1
    c = [None] * n
    for k in range(h):
        xr = g.inp("xr[%d]" % k)
        xi = g.inp("xi[%d]" % k)
        if k == 0 or (n % 2 == 0 and k == n // 2):
            c[k] = C(g, xr, g.ZERO)
        else:
            c[k] = C(g, g.add(xr, xi), g.ZERO)
            c[n - k] = C(g, g.sub(xr, xi), g.ZERO)
    f = dft(g, c, -1)
    outs = []
    for t in range(n):
        if t <= n // 2:
            node = g.add(f[t].re, f[t].im)
        else:
            node = g.sub(f[n - t].re, f[n - t].im)
        outs.append(("y[%d]" % t, node))
    return emit(g, "b2a_c2r%d" % n, "const float (&xr)[%d], const float (&xi)[%d], float (&y)[%d]" % (h, h, n), outs)


def gen_rdft_odd(n):
    """Odd-frequency DFT of a real sequence: U[k] = sum_j y[j] e^{-2 pi i j (k + 1/2)/n}, k = 0..(n-1)//2.
    This is the stage-B item k1 = N1/2 of the two-stage real FFT (inputs are real after stage A;
    the inter-stage twiddle e^{-pi i j/n} is folded in here as compile-time constants).  The other
    bins are conjugate mirrors: U[n-1-k] = conj(U[k]).
    Two formulations are generated and the cheaper one is kept: (a) twiddle then complex DFT of size n,
    (b) the odd bins of the real DFT of size 2n of the zero-extended sequence (real-input savings at every level)."""
    h = (n - 1) // 2 + 1
    name = "b2a_rdftodd%d" % n
    args = "const float (&x)[%d], float (&yr)[%d], float (&yi)[%d]" % (n, h, h)

    def form_a():
        g = G()
        xs = []
        for j in range(n):
            ang = -math.pi * j / n
            xs.append(C(g, g.inp("x[%d]" % j), g.ZERO).mul_const(math.cos(ang), math.sin(ang)))
        ys = dft(g, xs, -1)
        return g, [ys[k] for k in range(h)]

    def form_b():
        g = G()
        xs = [C(g, g.inp("x[%d]" % j), g.ZERO) for j in range(n)] + [C(g, g.ZERO, g.ZERO) for _ in range(n)]
        ys = dft(g, xs, -1)
        return g, [ys[2 * k + 1] for k in range(h)]

    best = None
    for form in (form_a, form_b):
        g, ys = form()
        outs = []
        for k in range(h):
            outs.append(("yr[%d]" % k, ys[k].re))
            outs.append(("yi[%d]" % k, ys[k].im))
        code, st = emit(g, name, args, outs)
        if best is None or sum(st) < sum(best[1]):
            best = (code, st)
    return best


RDFT = [16, 20, 25, 32, 40, 60, 64]
RDFT_ODD = [16, 20, 25, 32]
CDFT = [16, 20, 25, 30, 32]
C2R = [16, 20]


def main():
    out = []
    out.append("// GENERATED by tools/gen_codelets.py -- do not edit.\n"
               "// Straight-line fp32 DFT codelets (forward sign e^{-2 pi i nk/N}); c2r is the unnormalised inverse.\n"
               "#pragma once\n"
               "#if defined(__CUDACC__)\n#define B2A_CODELET __device__ __forceinline__\n#else\n"
               "#define B2A_CODELET static inline\n#endif\n")
    stats = []
    for n in RDFT:
        code, st = gen_rdft(n)
        out.append(code)
        stats.append(("rdft%d" % n, st))
    for n in RDFT_ODD:
        code, st = gen_rdft_odd(n)
        out.append(code)
        stats.append(("rdftodd%d" % n, st))
    for n in CDFT:
        code, st = gen_cdft(n)
        out.append(code)
        stats.append(("cdft%d" % n, st))
    for n in C2R:
        code, st = gen_c2r(n)
        out.append(code)
        stats.append(("c2r%d" % n, st))
    sys.stdout.write("\n".join(out))
    for nm, st in stats:
        sys.stderr.write("%-8s add=%4d mul=%4d neg=%3d total=%d\n" % (nm, st[0], st[1], st[2], sum(st)))


if __name__ == "__main__":
    main()
