"""Generate straight-line FFT codelets for the fused STFT kernel.

    python tools/gen_codelets.py            # rewrites deep_audio_mixer_b200/csrc/fft_codelets.cuh

A tiny symbolic engine builds the butterfly network, folds trivial twiddles (1, -i, the
(1 -+ i)/sqrt2 pair), fuses every constant multiply into an FMA (general twiddles use the
tangent-factored butterfly: 6 FMAs instead of 4 mul/fma + 4 add), hash-conses common
sub-expressions and eliminates dead code.  It then checks the network numerically against
numpy.fft before writing any C++.  The emitted functions are templates on the real type, so
the same codelet serves the float32 and the float64 kernels.

Codelets
  dam_cfft<N>(re, im)           in-place complex DFT, natural order, forward sign e^{-2 pi i nk/N}
  dam_rfft<2N>_packed(x, sr, si)  2N real samples -> N complex slots:
        slot 0      = (Y[0], Y[N])           both real, unscaled
        slot k      = 2*Y[k], k = 1..N-1     (the factor 2 is folded into the stage-2 twiddles)
     where Y = DFT_{2N}(x).
"""
import math
import os
import sys
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'deep_audio_mixer_b200', 'csrc', 'fft_codelets.cuh')


def exact_cos_sin(num: int, den: int):
    """(cos, sin)(2*pi*num/den) correctly rounded to double (50-digit mpmath evaluation)."""
    import mpmath
    mpmath.mp.dps = 50
    fr = Fraction(num, den) % 1
    ang = 2 * mpmath.pi * mpmath.mpf(fr.numerator) / mpmath.mpf(fr.denominator)
    c, s = float(mpmath.cos(ang)), float(mpmath.sin(ang))
    if fr.denominator in (1, 2, 4):                 # exact zeros instead of 1e-51
        c, s = float(round(c)), float(round(s))
    return c, s


class Graph:
    def __init__(self):
        self.nodes = []          # (op, args)
        self.memo = {}
        self.consts = {}

    def _mk(self, op, *args):
        key = (op,) + args
        if key in self.memo:
            return self.memo[key]
        self.nodes.append(key)
        idx = len(self.nodes) - 1
        self.memo[key] = idx
        return idx

    def inp(self, name):
        return V(self, self._mk('in', name), 1)


class V:
    """signed reference to a node; node None means exact zero."""
    __slots__ = ('g', 'n', 's')

    def __init__(self, g, n, s):
        self.g, self.n, self.s = g, n, s

    def is_zero(self):
        return self.n is None

    def __neg__(self):
        return V(self.g, self.n, -self.s)

    def __add__(self, o):
        if self.is_zero():
            return o
        if o.is_zero():
            return self
        g = self.g
        if self.s == o.s:
            a, b = sorted((self.n, o.n))
            return V(g, g._mk('add', a, b), self.s)
        if self.s > 0:
            return V(g, g._mk('sub', self.n, o.n), 1)
        return V(g, g._mk('sub', o.n, self.n), 1)

    def __sub__(self, o):
        return self + (-o)

    def scale(self, c: float):
        if self.is_zero() or c == 0.0:
            return V(self.g, None, 1)
        if c == 1.0:
            return self
        if c == -1.0:
            return -self
        s = self.s if c > 0 else -self.s
        return V(self.g, self.g._mk('mul', abs(c), self.n), s)

    def fma(self, c: float, x):
        """self + c * x as one fused node."""
        if x.is_zero() or c == 0.0:
            return self
        if self.is_zero():
            return x.scale(c)
        if abs(c) == 1.0:
            return self + x.scale(c)
        g = self.g
        ce = c * x.s                        # self.s * (|self| + (ce*self.s) * |x|)
        return V(g, g._mk('fma', ce * self.s, x.n, self.n), self.s)


def zero(g):
    return V(g, None, 1)


def butterfly(a, b, num, den):
    """(a + w*b, a - w*b) with w = exp(-2 pi i num/den); a, b are (re, im) pairs of V."""
    ar, ai = a
    br, bi = b
    fr = Fraction(num, den) % 1
    c, s = exact_cos_sin(num, den)
    wr, wi = c, -s
    if fr == 0:
        tr, ti = br, bi
        return (ar + tr, ai + ti), (ar - tr, ai - ti)
    if fr == Fraction(1, 2):
        return (ar - br, ai - bi), (ar + br, ai + bi)
    if fr == Fraction(1, 4):                       # w = -i : t = (bi, -br)
        return (ar + bi, ai - br), (ar - bi, ai + br)
    if fr == Fraction(3, 4):                       # w = +i : t = (-bi, br)
        return (ar - bi, ai + br), (ar + bi, ai - br)
    if fr.denominator == 8:                        # |wr| == |wi| == sqrt(1/2)
        h = math.sqrt(0.5)
        sr_, si_ = (1 if wr > 0 else -1), (1 if wi > 0 else -1)
        # t = h * ((sr*br - si*bi), (sr*bi + si*br))
        ur = br.scale(sr_) - bi.scale(si_)
        ui = bi.scale(sr_) + br.scale(si_)
        return (ar.fma(h, ur), ai.fma(h, ui)), (ar.fma(-h, ur), ai.fma(-h, ui))
    if abs(wr) >= abs(wi):
        t = wi / wr                                # w*b = wr * (br - t*bi, bi + t*br)
        ur = br.fma(-t, bi)
        ui = bi.fma(t, br)
        return (ar.fma(wr, ur), ai.fma(wr, ui)), (ar.fma(-wr, ur), ai.fma(-wr, ui))
    t = wr / wi                                    # w*b = wi * (t*br - bi, t*bi + br)
    ur = (-bi).fma(t, br)
    ui = br.fma(t, bi)
    return (ar.fma(wi, ur), ai.fma(wi, ui)), (ar.fma(-wi, ur), ai.fma(-wi, ui))


def fft_rec(x):
    """Radix-2 decimation in time over a list of (re, im) V pairs; natural order out."""
    n = len(x)
    if n == 1:
        return list(x)
    ev = fft_rec(x[0::2])
    od = fft_rec(x[1::2])
    out = [None] * n
    for k in range(n // 2):
        out[k], out[k + n // 2] = butterfly(ev[k], od[k], k, n)
    return out


def cmul_const(v, num, den):
    """v * exp(-2 pi i num/den) for constant twiddle (used in the real-FFT post-pass)."""
    c, s = exact_cos_sin(num, den)
    wr, wi = c, -s
    vr, vi = v
    if abs(wr) >= abs(wi):
        t = wi / wr
        return (vr.fma(-t, vi)).scale(wr), (vi.fma(t, vr)).scale(wr)
    t = wr / wi
    return ((-vi).fma(t, vr)).scale(wi), (vr.fma(t, vi)).scale(wi)


def build_cfft(n):
    g = Graph()
    x = [(g.inp('re[%d]' % i), g.inp('im[%d]' % i)) for i in range(n)]
    out = fft_rec(x)
    outs = []
    for k in range(n):
        outs.append(('re[%d]' % k, out[k][0]))
        outs.append(('im[%d]' % k, out[k][1]))
    return g, outs


def build_rfft_packed(n2):
    """n2 real inputs x[0..n2) -> n2/2 complex slots as described in the module docstring."""
    n = n2 // 2
    g = Graph()
    z = [(g.inp('x[%d]' % (2 * m)), g.inp('x[%d]' % (2 * m + 1))) for m in range(n)]
    Z = fft_rec(z)
    outs = []
    zr, zi = Z[0]
    outs.append(('sr[0]', zr + zi))                 # Y[0]
    outs.append(('si[0]', zr - zi))                 # Y[n]
    slots = {}
    for k in range(1, n // 2 + 1):
        A = Z[k]
        B = (Z[n - k][0], -Z[n - k][1])             # conj Z[n-k]
        S = (A[0] + B[0], A[1] + B[1])
        D = (A[0] - B[0], A[1] - B[1])
        if 2 * k == n:                              # W = -i, D = 2i Im Z, P = 2 Im Z (real)
            slots[k] = (A[0] + A[0], -(A[1] + A[1]))
            continue
        # 2Y[k] = S - i W D ; 2Y[n-k] = conj(S + i W D), W = exp(-2 pi i k / n2)
        c, s = exact_cos_sin(k, n2)
        wr, wi = c, -s
        # P = W*D computed unfused from S so each output is one FMA-pair deep
        Dr, Di = D
        # Y2[k].re = S.re + P.im = S.re + wr*Di + wi*Dr ; Y2[k].im = S.im - P.re
        y_k = (S[0].fma(wr, Di).fma(wi, Dr), S[1].fma(-wr, Dr).fma(wi, Di))
        # S + iP = (S.re - P.im, S.im + P.re); conj -> (S.re - P.im, -(S.im + P.re))
        y_nk = (S[0].fma(-wr, Di).fma(-wi, Dr), -(S[1].fma(wr, Dr).fma(-wi, Di)))
        slots[k], slots[n - k] = y_k, y_nk
    for k in range(1, n):
        outs.append(('sr[%d]' % k, slots[k][0]))
        outs.append(('si[%d]' % k, slots[k][1]))
    return g, outs


def evaluate(g, outs, inputs):
    """Numerically evaluate the graph (float64) for verification."""
    val = [None] * len(g.nodes)
    for i, nd in enumerate(g.nodes):
        op = nd[0]
        if op == 'in':
            val[i] = inputs[nd[1]]
        elif op == 'add':
            val[i] = val[nd[1]] + val[nd[2]]
        elif op == 'sub':
            val[i] = val[nd[1]] - val[nd[2]]
        elif op == 'mul':
            val[i] = nd[1] * val[nd[2]]
        elif op == 'fma':
            val[i] = nd[1] * val[nd[2]] + val[nd[3]]
    res = {}
    for name, v in outs:
        res[name] = 0.0 if v.is_zero() else v.s * val[v.n]
    return res


def emit(g, outs, in_place_arrays):
    """Returns (lines, op_counts).  Dead nodes are dropped; temporaries are SSA consts."""
    live = set()
    stack = [v.n for _, v in outs if not v.is_zero()]
    while stack:
        i = stack.pop()
        if i in live:
            continue
        live.add(i)
        nd = g.nodes[i]
        if nd[0] in ('add', 'sub'):
            stack += [nd[1], nd[2]]
        elif nd[0] == 'mul':
            stack.append(nd[2])
        elif nd[0] == 'fma':
            stack += [nd[2], nd[3]]
    counts = {'add': 0, 'mul': 0, 'fma': 0}
    name = {}
    lines = []
    # inputs that are also outputs (in-place) must be read into temporaries first
    for i in sorted(live):
        nd = g.nodes[i]
        if nd[0] == 'in':
            if in_place_arrays:
                name[i] = 'i%d' % i
                lines.append('const R i%d = %s;' % (i, nd[1]))
            else:
                name[i] = nd[1]

    def lit(c):
        return 'R(%s)' % repr(float(c))

    for i in sorted(live):
        nd = g.nodes[i]
        op = nd[0]
        if op == 'in':
            continue
        name[i] = 't%d' % i
        if op == 'add':
            lines.append('const R t%d = %s + %s;' % (i, name[nd[1]], name[nd[2]]))
            counts['add'] += 1
        elif op == 'sub':
            lines.append('const R t%d = %s - %s;' % (i, name[nd[1]], name[nd[2]]))
            counts['add'] += 1
        elif op == 'mul':
            lines.append('const R t%d = %s * %s;' % (i, lit(nd[1]), name[nd[2]]))
            counts['mul'] += 1
        elif op == 'fma':
            lines.append('const R t%d = dam_fma(%s, %s, %s);' % (i, lit(nd[1]), name[nd[2]], name[nd[3]]))
            counts['fma'] += 1
    for nm, v in outs:
        if v.is_zero():
            lines.append('%s = R(0);' % nm)
        else:
            lines.append('%s = %s%s;' % (nm, '-' if v.s < 0 else '', name[v.n]))
    return lines, counts


def emit_packed(g, outs, in_place_arrays):
    """Same network on PACKED pairs (two independent transforms per thread, one per half of a 64-bit
    register): every op becomes one sm_100 f32x2 instruction (FADD2 / FMUL2 / FFMA2 with a scalar
    immediate), i.e. half the issue slots per transform."""
    live = set()
    stack = [v.n for _, v in outs if not v.is_zero()]
    while stack:
        i = stack.pop()
        if i in live:
            continue
        live.add(i)
        nd = g.nodes[i]
        if nd[0] in ('add', 'sub'):
            stack += [nd[1], nd[2]]
        elif nd[0] == 'mul':
            stack.append(nd[2])
        elif nd[0] == 'fma':
            stack += [nd[2], nd[3]]
    counts = {'add': 0, 'mul': 0, 'fma': 0}
    name = {}
    lines = []
    for i in sorted(live):
        nd = g.nodes[i]
        if nd[0] == 'in':
            if in_place_arrays:
                name[i] = 'i%d' % i
                lines.append('const pk2 i%d = %s;' % (i, nd[1]))
            else:
                name[i] = nd[1]

    def lit(c):
        return 'pk_c(%sf)' % repr(float(np.float32(c)))

    for i in sorted(live):
        nd = g.nodes[i]
        op = nd[0]
        if op == 'in':
            continue
        name[i] = 't%d' % i
        if op == 'add':
            lines.append('const pk2 t%d = pk_add(%s, %s);' % (i, name[nd[1]], name[nd[2]]))
            counts['add'] += 1
        elif op == 'sub':
            lines.append('const pk2 t%d = pk_sub(%s, %s);' % (i, name[nd[1]], name[nd[2]]))
            counts['add'] += 1
        elif op == 'mul':
            lines.append('const pk2 t%d = pk_mul(%s, %s);' % (i, name[nd[2]], lit(nd[1])))
            counts['mul'] += 1
        elif op == 'fma':
            lines.append('const pk2 t%d = pk_fma(%s, %s, %s);' % (i, name[nd[2]], lit(nd[1]), name[nd[3]]))
            counts['fma'] += 1
    for nm, v in outs:
        if v.is_zero():
            lines.append('%s = pk_c(0.0f);' % nm)
        elif v.s < 0:
            lines.append('%s = pk_mul(%s, pk_c(-1.0f));' % (nm, name[v.n]))
            counts['mul'] += 1
        else:
            lines.append('%s = %s;' % (nm, name[v.n]))
    return lines, counts


# ---------------------------------------------------------------------------------------------------
# Packed-COMPLEX codelets: one complex value = one 64-bit register (re in the low half, im in the high
# half), arithmetic on sm_100's packed FP32 instructions.  ptxas folds half swaps and per-half signs
# into FFMA2/FADD2 operand modifiers (.LO_HI, .NP/.PN) and equal-magnitude constant pairs into a scalar
# immediate, so
#   a +- b                       1 FADD2 each
#   a +- (-i) b, a +- i b        1 FFMA2 each   (swap(b) * (+-1, -+1) + a)
#   a +- w b, general twiddle    3 FFMA2        (tangent-factored: u = b + t*(-i or i)b ; a +- c*u')
# i.e. 194 issue slots for a 32-point complex DFT instead of 388, with the SAME 64 data registers.
class PC:
    """Emitter that also evaluates every emitted value numerically (complex128) for verification."""

    def __init__(self):
        self.lines = []
        self.n = 0
        self.val = {}

    def new(self, expr, value):
        name = 'p%d' % self.n
        self.n += 1
        self.lines.append('const pk2 %s = %s;' % (name, expr))
        self.val[name] = value
        return name

    def inp(self, name, value):
        self.val[name] = value
        return name

    @staticmethod
    def C(a, b):
        return 'pk_make(%sf, %sf)' % (repr(float(np.float32(a))), repr(float(np.float32(b))))

    def add(self, a, b):
        return self.new('pk_add(%s, %s)' % (a, b), self.val[a] + self.val[b])

    def sub(self, a, b):
        return self.new('pk_sub(%s, %s)' % (a, b), self.val[a] - self.val[b])

    def fma(self, x, ca, cb, y, swap=False, swap_y=False):
        """(x or swap(x)) * (ca, cb) + (y or swap(y)), per half."""
        xv = self.val[x]
        xs = complex(xv.imag, xv.real) if swap else xv
        yv = self.val[y]
        ys = complex(yv.imag, yv.real) if swap_y else yv
        value = complex(xs.real * ca, xs.imag * cb) + ys
        xe = 'pk_swap(%s)' % x if swap else x
        ye = 'pk_swap(%s)' % y if swap_y else y
        return self.new('pk_fma(%s, %s, %s)' % (xe, self.C(ca, cb), ye), value)

    def mul(self, x, ca, cb, swap=False):
        xv = self.val[x]
        xs = complex(xv.imag, xv.real) if swap else xv
        xe = 'pk_swap(%s)' % x if swap else x
        return self.new('pk_mul(%s, %s)' % (xe, self.C(ca, cb)), complex(xs.real * ca, xs.imag * cb))


def pc_butterfly(e, a, b, num, den):
    """(a + w b, a - w b), w = exp(-2 pi i num/den), on packed complex values."""
    fr = Fraction(num, den) % 1
    c, s = exact_cos_sin(num, den)
    wr, wi = c, -s
    if fr == 0:
        return e.add(a, b), e.sub(a, b)
    if fr == Fraction(1, 2):
        return e.sub(a, b), e.add(a, b)
    if fr == Fraction(1, 4):                       # w b = (bi, -br)
        return e.fma(b, 1.0, -1.0, a, swap=True), e.fma(b, -1.0, 1.0, a, swap=True)
    if fr == Fraction(3, 4):                       # w b = (-bi, br)
        return e.fma(b, -1.0, 1.0, a, swap=True), e.fma(b, 1.0, -1.0, a, swap=True)
    if abs(wr) >= abs(wi):
        t = wi / wr                                # w b = wr (b + t (-bi, br))
        u = e.fma(b, -t, t, b, swap=True)
        return e.fma(u, wr, wr, a), e.fma(u, -wr, -wr, a)
    t = wr / wi                                    # u = b + t (bi, -br); w b = wi (-ui, ur)
    u = e.fma(b, t, -t, b, swap=True)
    return e.fma(u, -wi, wi, a, swap=True), e.fma(u, wi, -wi, a, swap=True)


def pc_fft(e, x):
    n = len(x)
    if n == 1:
        return list(x)
    ev = pc_fft(e, x[0::2])
    od = pc_fft(e, x[1::2])
    out = [None] * n
    for k in range(n // 2):
        out[k], out[k + n // 2] = pc_butterfly(e, ev[k], od[k], k, n)
    return out


def build_pc_cfft(n):
    rng = np.random.default_rng(n)
    e = PC()
    zin = rng.normal(size=n) + 1j * rng.normal(size=n)
    x = [e.inp('z[%d]' % i, zin[i]) for i in range(n)]
    pre = ['const pk2 z%d_ = z[%d];' % (i, i) for i in range(n)]
    e2 = PC()
    x = [e2.inp('z%d_' % i, zin[i]) for i in range(n)]
    out = pc_fft(e2, x)
    ref = np.fft.fft(zin)
    got = np.array([e2.val[o] for o in out])
    assert np.max(np.abs(got - ref)) < 1e-5 * n, np.max(np.abs(got - ref))       # constants are float32
    lines = pre + e2.lines + ['z[%d] = %s;' % (k, out[k]) for k in range(n)]
    return lines, e2.n


def build_pc_rfft(n2):
    """n2 real samples given as n2/2 packed pairs xp[m] = (x[2m], x[2m+1]) -> slots:
       y0n (lo = Y[0], hi = Y[n2/2], both real) and s[k] = 2*Y[k], k = 1..n2/2-1 (packed complex)."""
    n = n2 // 2
    rng = np.random.default_rng(n2)
    xin = rng.normal(size=n2)
    e = PC()
    pre = ['const pk2 z%d_ = xp[%d];' % (m, m) for m in range(n)]
    z = [e.inp('z%d_' % m, complex(xin[2 * m], xin[2 * m + 1])) for m in range(n)]
    Z = pc_fft(e, z)
    ref = np.fft.rfft(xin)
    outs = []
    # slot 0: (Zr + Zi, Zr - Zi) = Z0 * (1, -1) + swap(Z0)
    y0n = e.fma(Z[0], 1.0, -1.0, Z[0], swap_y=True)
    assert abs(e.val[y0n].real - ref[0].real) < 1e-4 and abs(e.val[y0n].imag - ref[n].real) < 1e-4
    outs.append('y0n = %s;' % y0n)
    for k in range(1, n // 2 + 1):
        A, B = Z[k], Z[n - k]
        if 2 * k == n:
            yk = e.mul(A, 2.0, -2.0)                 # 2 conj(Z[n/2])
            assert abs(e.val[yk] - 2 * ref[k]) < 1e-4
            outs.append('s[%d] = %s;' % (k, yk))
            continue
        c, sn = exact_cos_sin(k, n2)
        wr, wi = c, -sn
        S = e.fma(B, 1.0, -1.0, A)                   # A + conj(B)
        D = e.fma(B, -1.0, 1.0, A)                   # A - conj(B)
        Sc = e.fma(A, 1.0, -1.0, B)                  # conj(A) + B = conj(S)
        P0 = e.mul(D, wr, wr)
        P = e.fma(D, -wi, wi, P0, swap=True)         # W * D
        yk = e.fma(P, 1.0, -1.0, S, swap=True)       # S - i P
        ynk = e.fma(P, -1.0, -1.0, Sc, swap=True)    # conj(S + i P) = conj(S) + (-Pi, -Pr)
        assert abs(e.val[yk] - 2 * ref[k]) < 1e-4 and abs(e.val[ynk] - 2 * ref[n - k]) < 1e-4
        outs.append('s[%d] = %s;' % (k, yk))
        outs.append('s[%d] = %s;' % (n - k, ynk))
    return pre + e.lines + outs, e.n


def verify():
    rng = np.random.default_rng(0)
    for n in (8, 16, 32):
        g, outs = build_cfft(n)
        zr, zi = rng.normal(size=n), rng.normal(size=n)
        inputs = {('re[%d]' % i): zr[i] for i in range(n)}
        inputs.update({('im[%d]' % i): zi[i] for i in range(n)})
        res = evaluate(g, outs, inputs)
        ref = np.fft.fft(zr + 1j * zi)
        got = np.array([res['re[%d]' % k] + 1j * res['im[%d]' % k] for k in range(n)])
        assert np.max(np.abs(got - ref)) < 1e-13 * n, (n, np.max(np.abs(got - ref)))
    for n2 in (16, 32, 64):
        g, outs = build_rfft_packed(n2)
        x = rng.normal(size=n2)
        res = evaluate(g, outs, {('x[%d]' % i): x[i] for i in range(n2)})
        ref = np.fft.rfft(x)
        n = n2 // 2
        assert abs(res['sr[0]'] - ref[0].real) < 1e-12 and abs(res['si[0]'] - ref[n].real) < 1e-12
        got = np.array([res['sr[%d]' % k] + 1j * res['si[%d]' % k] for k in range(1, n)])
        assert np.max(np.abs(got - 2 * ref[1:n])) < 1e-12, n2


def main():
    verify()
    parts = []
    parts.append('// GENERATED by tools/gen_codelets.py -- do not edit by hand.\n'
                 '// Straight-line FFT codelets (radix-2 DIT network, trivial twiddles folded,\n'
                 '// tangent-factored FMA butterflies).  Verified against numpy.fft at generation time.\n'
                 '#pragma once\n\n'
                 '__device__ __forceinline__ float  dam_fma(float a, float b, float c)   { return __fmaf_rn(a, b, c); }\n'
                 '__device__ __forceinline__ double dam_fma(double a, double b, double c) { return __fma_rn(a, b, c); }\n')
    summary = []
    for n in (32,):                      # the only sizes the kernels instantiate (stage 2: 32-point complex)
        g, outs = build_cfft(n)
        lines, counts = emit(g, outs, in_place_arrays=True)
        summary.append('// dam_cfft%d: %d add, %d mul, %d fma = %d ops' %
                       (n, counts['add'], counts['mul'], counts['fma'], sum(counts.values())))
        parts.append('// In-place forward complex DFT of %d points, natural order in and out.\n'
                     'template <typename R>\n__device__ __forceinline__ void dam_cfft%d(R (&re)[%d], R (&im)[%d]) {\n    %s\n}\n'
                     % (n, n, n, n, '\n    '.join(lines)))
    for n2 in (64,):                     # stage 1: 64 real samples
        g, outs = build_rfft_packed(n2)
        lines, counts = emit(g, outs, in_place_arrays=False)
        summary.append('// dam_rfft%d_packed: %d add, %d mul, %d fma = %d ops' %
                       (n2, counts['add'], counts['mul'], counts['fma'], sum(counts.values())))
        parts.append('// %d real samples -> %d complex slots: slot 0 = (Y[0], Y[%d]); slot k = 2*Y[k].\n'
                     'template <typename R>\n__device__ __forceinline__ void dam_rfft%d_packed(const R (&x)[%d], R (&sr)[%d], R (&si)[%d]) {\n    %s\n}\n'
                     % (n2, n2 // 2, n2 // 2, n2, n2, n2 // 2, n2 // 2, '\n    '.join(lines)))
    # ---- packed (f32x2) float-only helpers
    parts.append('// ---- sm_100 packed FP32 (f32x2) helpers: a 64-bit register holds two floats ----\n'
                 'typedef unsigned long long pk2;\n'
                 '__device__ __forceinline__ pk2 pk_make(float lo, float hi) { pk2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }\n'
                 '__device__ __forceinline__ pk2 pk_c(float c) { return pk_make(c, c); }   // ptxas folds this into a scalar immediate / broadcast operand\n'
                 '__device__ __forceinline__ void pk_split(pk2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }\n'
                 '__device__ __forceinline__ pk2 pk_add(pk2 a, pk2 b) { pk2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }\n'
                 '__device__ __forceinline__ pk2 pk_sub(pk2 a, pk2 b) { pk2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }\n'
                 '__device__ __forceinline__ pk2 pk_mul(pk2 a, pk2 b) { pk2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }\n'
                 '__device__ __forceinline__ pk2 pk_fma(pk2 a, pk2 b, pk2 c) { pk2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }\n')
    # ---- packed-complex variants (one complex value per 64-bit register)
    parts.append('__device__ __forceinline__ pk2 pk_swap(pk2 v) { float lo, hi; pk_split(v, lo, hi); return pk_make(hi, lo); }   // folded into an operand modifier\n')
    lines, nops = build_pc_cfft(32)
    summary.append('// dam_cfft32_pc: %d packed ops (one 32-point complex transform, packed complex registers)' % nops)
    parts.append('// In-place forward complex DFT of 32 points; z[k] = (re, im) packed.\n'
                 '__device__ __forceinline__ void dam_cfft32_pc(pk2 (&z)[32]) {\n    %s\n}\n' % '\n    '.join(lines))
    lines, nops = build_pc_rfft(64)
    summary.append('// dam_rfft64_pc: %d packed ops (64 real samples, packed complex registers)' % nops)
    parts.append('// 64 real samples as 32 packed pairs xp[m] = (x[2m], x[2m+1]) -> y0n = (Y[0], Y[32]) and s[k] = 2*Y[k] (packed complex), k = 1..31.\n'
                 '__device__ __forceinline__ void dam_rfft64_pc(const pk2 (&xp)[32], pk2& y0n, pk2 (&s)[32]) {\n    %s\n}\n' % '\n    '.join(lines))
    text = parts[0] + '\n' + '\n'.join(summary) + '\n\n' + '\n'.join(parts[1:])
    with open(OUT, 'w') as f:
        f.write(text)
    print('\n'.join(summary))
    print('wrote', OUT)


if __name__ == '__main__':
    sys.exit(main())
