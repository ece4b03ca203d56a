#!/usr/bin/env python
"""Generates the straight-line FFT codelets of the log-mel kernels (csrc/codelets.cuh).

The kernels keep one frame per lane and split a frame's real FFT over the 16 warps of a CTA
(DESIGN.md §5): every warp runs one small transform entirely in registers, so the transforms are
emitted as straight-line code with zeros, real inputs and unused outputs pruned at generation time:

  recipe K (512-point real FFT of 400 samples, n = 16 n1 + n2, k = k1 + 32 k2)
    k_pass1        real DFT-32 over n1 (25 non-zero inputs)         -> k1 = 0..16
    dft16          complex DFT-16 over n2 (rows k1 = 1..15)         -> X[k1 + 32 k2]
    k_pass2_edge   rows k1 = 0 and k1 = 16 together (real inputs)   -> X[32 k2] (k2 = 1..7), X[16 + 32 k2] (k2 = 0..7)
  recipe W (400-point real FFT, n = 16 n1 + n2, k = k1 + 25 k2)
    w_pass1        real DFT-25 over n1                              -> k1 = 0..12
    w_pass2_edge   row k1 = 0 (real inputs)                         -> X[25 k2], k2 = 0..8
    (rows k1 = 1..12 use dft16)

A tiny expression DAG (add / sub / mul-by-constant / neg, with exact-zero tracking and common
sub-expression sharing) is built by a mixed-radix Cooley-Tukey recursion; only what the requested
outputs reach is emitted.  Every codelet is checked here against numpy.fft before it is written, and
again on the host through g++ by tests/test_codelets.py.

    python tools/gen_codelets.py            # rewrites speech_transcript_embeddings_b200/csrc/codelets.cuh
"""
from __future__ import annotations

import cmath
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "speech_transcript_embeddings_b200" / "csrc" / "codelets.cuh"


class Graph:
    """Real-valued expression DAG.  A value is a node id, or None for an exact zero."""

    def __init__(self):
        self.nodes = []          # (op, a, b)  op in {in, add, sub, mul, neg}
        self.memo = {}

    def _mk(self, op, a, b=None):
        key = (op, a, b)
        if op == "add" and a > b:
            key = (op, b, a)
        hit = self.memo.get(key)
        if hit is not None:
            return hit
        self.nodes.append(key)
        self.memo[key] = len(self.nodes) - 1
        return len(self.nodes) - 1

    def inp(self, name):
        return self._mk("in", name)

    def is_neg(self, a):
        return a is not None and self.nodes[a][0] == "neg"

    def neg(self, a):
        if a is None:
            return None
        if self.is_neg(a):
            return self.nodes[a][1]
        if self.nodes[a][0] == "mul":
            return self.mul(-self.nodes[a][1], self.nodes[a][2])
        if self.nodes[a][0] == "sub":
            return self._mk("sub", self.nodes[a][2], self.nodes[a][1])
        return self._mk("neg", a)

    def add(self, a, b):
        if a is None:
            return b
        if b is None:
            return a
        if self.is_neg(b):
            return self.sub(a, self.nodes[b][1])
        if self.is_neg(a):
            return self.sub(b, self.nodes[a][1])
        return self._mk("add", a, b)

    def sub(self, a, b):
        if b is None:
            return a
        if a is None:
            return self.neg(b)
        if a == b:
            return None
        if self.is_neg(b):
            return self.add(a, self.nodes[b][1])
        if self.is_neg(a):
            return self.neg(self.add(self.nodes[a][1], b))
        return self._mk("sub", a, b)

    def mul(self, c, a):
        if a is None or c == 0.0:
            return None
        if c == 1.0:
            return a
        if c == -1.0:
            return self.neg(a)
        if self.is_neg(a):
            return self.mul(-c, self.nodes[a][1])
        if self.nodes[a][0] == "mul":
            return self.mul(c * self.nodes[a][1], self.nodes[a][2])
        return self._mk("mul", float(c), a)


class Cx:
    """Complex value over a Graph: (re, im) node ids."""
    __slots__ = ("g", "re", "im")

    def __init__(self, g, re, im):
        self.g, self.re, self.im = g, re, im

    def __add__(self, o):
        return Cx(self.g, self.g.add(self.re, o.re), self.g.add(self.im, o.im))

    def __sub__(self, o):
        return Cx(self.g, self.g.sub(self.re, o.re), self.g.sub(self.im, o.im))

    def conj(self):
        return Cx(self.g, self.re, self.g.neg(self.im))

    def mul_mj(self):     # times -j
        return Cx(self.g, self.im, self.g.neg(self.re))

    def mul_j(self):
        return Cx(self.g, self.g.neg(self.im), self.re)

    def scale(self, c):
        return Cx(self.g, self.g.mul(c, self.re), self.g.mul(c, self.im))

    def cmul(self, w: complex):
        c, s = _snap(w.real), _snap(w.imag)
        g = self.g
        if s == 0.0:
            return self.scale(c)
        if c == 0.0:
            return Cx(g, g.mul(-s, self.im), g.mul(s, self.re))
        if abs(abs(c) - abs(s)) < 1e-15:
            # c (a + jb)(1 + j s/c): one add and one multiply per component
            sg = 1.0 if (c > 0) == (s > 0) else -1.0
            if sg > 0:
                re, im = g.sub(self.re, self.im), g.add(self.re, self.im)
            else:
                re, im = g.add(self.re, self.im), g.sub(self.im, self.re)
            return Cx(g, g.mul(c, re), g.mul(c, im))
        return Cx(g, g.sub(g.mul(c, self.re), g.mul(s, self.im)), g.add(g.mul(s, self.re), g.mul(c, self.im)))


def _snap(v):
    for t in (0.0, 1.0, -1.0):
        if abs(v - t) < 1e-15:
            return t
    return v


def w(n, k):
    """exp(-2 pi j k / n), exact at the multiples of pi/4."""
    k %= n
    z = cmath.exp(-2j * math.pi * k / n)
    if (8 * k) % n == 0:
        e = (8 * k // n) % 8
        h = math.sqrt(0.5)
        z = [1, h - 1j * h, -1j, -h - 1j * h, -1, -h + 1j * h, 1j, h + 1j * h][e]
    return complex(z)


def is_real(xs):
    return all(x.im is None for x in xs)


def butterfly(xs):
    """Small forward DFT of len(xs) in {2, 3, 4, 5} complex values."""
    r = len(xs)
    g = xs[0].g
    if r == 2:
        return [xs[0] + xs[1], xs[0] - xs[1]]
    if r == 4:
        t0, t1 = xs[0] + xs[2], xs[0] - xs[2]
        t2, t3 = xs[1] + xs[3], xs[1] - xs[3]
        return [t0 + t2, t1 + t3.mul_mj(), t0 - t2, t1 + t3.mul_j()]
    if r == 5:
        c1, c2 = math.cos(2 * math.pi / 5), math.cos(4 * math.pi / 5)
        s1, s2 = math.sin(2 * math.pi / 5), math.sin(4 * math.pi / 5)
        t1, t2 = xs[1] + xs[4], xs[2] + xs[3]
        t3, t4 = xs[1] - xs[4], xs[2] - xs[3]
        m1 = xs[0] + t1.scale(c1) + t2.scale(c2)
        m2 = xs[0] + t1.scale(c2) + t2.scale(c1)
        n1 = t3.scale(s1) + t4.scale(s2)
        n2 = t3.scale(s2) - t4.scale(s1)
        return [xs[0] + t1 + t2, m1 + n1.mul_mj(), m2 + n2.mul_mj(), m2 + n2.mul_j(), m1 + n1.mul_j()]
    raise ValueError(r)


def dft(xs):
    """Forward DFT of a list of Cx by decimation in time; real inputs produce shared conjugate halves."""
    n = len(xs)
    if n == 1:
        return list(xs)
    if n in (2, 4, 5):
        out = butterfly(xs)
    else:
        r = 4 if n % 4 == 0 else 2 if n % 2 == 0 else 5
        assert n % r == 0, n
        m = n // r
        subs = [dft(xs[q::r]) for q in range(r)]
        out = [None] * n
        real = is_real(xs)
        tw = {}
        for k in range(m):
            if real and r in (2, 4) and k > m - k:
                # real input: sub[q][m-k'] = conj(sub[q][k'])  =>  twiddled value = W_r^q conj(twiddled value of column m-k)
                tw[k] = [tw[m - k][q].conj().cmul(w(r, q)) for q in range(r)]
            else:
                tw[k] = [subs[q][k].cmul(w(n, q * k)) for q in range(r)]
            col = butterfly(tw[k])
            for j in range(r):
                out[k + m * j] = col[j]
    if is_real(xs):
        for k in range(n // 2 + 1, n):
            out[k] = out[n - k].conj()
    return out


class Codelet:
    def __init__(self, name, doc):
        self.name, self.doc = name, doc
        self.g = Graph()
        self.params = []         # (c type prefix, name, length, is_output)
        self.outputs = []        # (lvalue, node or None)

    def real_in(self, name, count, nonzero=None):
        self.params.append(("const T", name, count, False))
        nz = count if nonzero is None else nonzero
        return [Cx(self.g, self.g.inp(f"{name}[{i}]") if i < nz else None, None) for i in range(count)]

    def complex_in(self, re, im, count):
        self.params.append(("const T", re, count, False))
        self.params.append(("const T", im, count, False))
        return [Cx(self.g, self.g.inp(f"{re}[{i}]"), self.g.inp(f"{im}[{i}]")) for i in range(count)]

    def out_arrays(self, *names_counts):
        for name, count in names_counts:
            self.params.append(("T", name, count, True))

    def emit_out(self, lvalue, node):
        self.outputs.append((lvalue, node))

    # ---- evaluation / emission ----
    def _reachable(self):
        need = set()
        stack = [n for _, n in self.outputs if n is not None]
        while stack:
            i = stack.pop()
            if i in need:
                continue
            need.add(i)
            op, a, b = self.g.nodes[i]
            if op in ("add", "sub"):
                stack += [a, b]
            elif op == "mul":
                stack.append(b)
            elif op == "neg":
                stack.append(a)
        return sorted(need)

    def op_count(self):
        return sum(1 for i in self._reachable() if self.g.nodes[i][0] != "in")

    def fused_op_count(self):
        """Ops after mul+add contraction (a multiply whose only consumer is one add/sub becomes an FMA)."""
        reach = self._reachable()
        uses = {}
        for i in reach:
            op, a, b = self.g.nodes[i]
            for s_ in ((a, b) if op in ("add", "sub") else (b,) if op == "mul" else (a,) if op == "neg" else ()):
                uses.setdefault(s_, []).append(i)
        for _, n_ in self.outputs:
            if n_ is not None:
                uses.setdefault(n_, []).append(-1)
        fused, taken = 0, set()
        for i in reach:
            if self.g.nodes[i][0] == "mul" and len(uses.get(i, [])) == 1 and uses[i][0] >= 0:
                u = uses[i][0]
                if self.g.nodes[u][0] in ("add", "sub") and u not in taken:
                    taken.add(u)
                    fused += 1
        return self.op_count() - fused

    def evaluate(self, env):
        """env: {"name[i]": value}; returns {lvalue: value} in float64 (numpy scalars)."""
        val = {}
        for i in self._reachable():
            op, a, b = self.g.nodes[i]
            if op == "in":
                val[i] = np.float64(env[a])
            elif op == "add":
                val[i] = val[a] + val[b]
            elif op == "sub":
                val[i] = val[a] - val[b]
            elif op == "mul":
                val[i] = np.float64(a) * val[b]
            else:
                val[i] = -val[a]
        return {lv: (val[n] if n is not None else np.float64(0.0)) for lv, n in self.outputs}

    def source(self):
        lines = [f"// {self.doc}  [{self.op_count()} arithmetic ops]",
                 "template <typename T>",
                 f"__host__ __device__ __forceinline__ void {self.name}(" +
                 ", ".join(f"{ty} (&{nm})[{cnt}]" for ty, nm, cnt, _ in self.params) + ") {"]
        name = {}
        for i in self._reachable():
            op, a, b = self.g.nodes[i]
            if op == "in":
                name[i] = a
                continue
            name[i] = f"t{i}"
            if op == "add":
                rhs = f"{name[a]} + {name[b]}"
            elif op == "sub":
                rhs = f"{name[a]} - {name[b]}"
            elif op == "mul":
                rhs = f"T({a!r}) * {name[b]}"
            else:
                rhs = f"-{name[a]}"
            lines.append(f"    const T t{i} = {rhs};")
        for lv, n in self.outputs:
            lines.append(f"    {lv} = {name[n] if n is not None else 'T(0)'};")
        lines.append("}")
        return "\n".join(lines)


# ---------------------------------------------------------------------------------------------
# the codelets
# ---------------------------------------------------------------------------------------------
def make_dft16():
    c = Codelet("dft16", "forward complex DFT-16, natural order: y[k] = sum_n x[n] exp(-2 pi j n k / 16)")
    xs = c.complex_in("xr", "xi", 16)
    c.out_arrays(("yr", 16), ("yi", 16))
    ys = dft(xs)
    for k in range(16):
        c.emit_out(f"yr[{k}]", ys[k].re)
        c.emit_out(f"yi[{k}]", ys[k].im)

    def check(rng):
        x = rng.standard_normal(16) + 1j * rng.standard_normal(16)
        env = {f"xr[{i}]": x[i].real for i in range(16)} | {f"xi[{i}]": x[i].imag for i in range(16)}
        got = c.evaluate(env)
        ref = np.fft.fft(x)
        return max(abs(complex(got[f"yr[{k}]"], got[f"yi[{k}]"]) - ref[k]) for k in range(16))
    return c, check


def make_k_pass1():
    c = Codelet("k_pass1", "recipe K pass 1: real DFT-32 of y[0..24] (y[25..31] = 0) -> re/im[k1], k1 = 0..16 "
                           "(im[0] = im[16] = 0 are not written)")
    xs = c.real_in("y", 32, nonzero=25)
    c.params[-1] = ("const T", "y", 25, False)
    c.out_arrays(("re", 17), ("im", 17))
    ys = dft(xs)
    for k in range(17):
        c.emit_out(f"re[{k}]", ys[k].re)
        if k not in (0, 16):
            c.emit_out(f"im[{k}]", ys[k].im)

    def check(rng):
        y = rng.standard_normal(25)
        got = c.evaluate({f"y[{i}]": y[i] for i in range(25)})
        ref = np.fft.fft(np.concatenate([y, np.zeros(7)]))
        err = 0.0
        for k in range(17):
            im = got.get(f"im[{k}]", 0.0)
            err = max(err, abs(complex(got[f"re[{k}]"], im) - ref[k]))
        return err
    return c, check


def make_k_pass2_edge():
    c = Codelet("k_pass2_edge",
                "recipe K pass 2, rows k1 = 0 and k1 = 16 (both real): a[n2] = R[0][n2], r[n2] = R[16][n2]; "
                "e0[k2-1] = X[32 k2] (k2 = 1..7), e16[k2] = X[16 + 32 k2] (k2 = 0..7)")
    a = c.real_in("a", 16)
    r = c.real_in("r", 16)
    c.out_arrays(("e0r", 7), ("e0i", 7), ("e16r", 8), ("e16i", 8))
    ya = dft(a)
    # X[16 + 32 k2] = sum_n r[n] W32^(n (2 k2 + 1)); with n = m, m + 8:  W32^(8 (2 k2 + 1)) = -j (-1)^k2, so the even
    # outputs are a DFT-8 of u[m] = W32^m (r[m] - j r[m + 8]) and the odd ones follow from X[15 - k2] = conj X[k2]
    u = [Cx(c.g, r[m].re, c.g.neg(r[m + 8].re)).cmul(w(32, m)) for m in range(8)]
    ue = dft(u)
    yb = [None] * 16
    for q in range(8):
        yb[2 * q] = ue[q]
    for k2 in range(1, 8, 2):
        yb[k2] = yb[15 - k2].conj()
    for k2 in range(1, 8):
        c.emit_out(f"e0r[{k2 - 1}]", ya[k2].re)
        c.emit_out(f"e0i[{k2 - 1}]", ya[k2].im)
    for k2 in range(8):
        c.emit_out(f"e16r[{k2}]", yb[k2].re)
        c.emit_out(f"e16i[{k2}]", yb[k2].im)

    def check(rng):
        av, rv = rng.standard_normal(16), rng.standard_normal(16)
        got = c.evaluate({f"a[{i}]": av[i] for i in range(16)} | {f"r[{i}]": rv[i] for i in range(16)})
        ra = np.fft.fft(av)
        rb = np.fft.fft(rv * np.exp(-2j * np.pi * 16 * np.arange(16) / 512))
        err = 0.0
        for k2 in range(1, 8):
            err = max(err, abs(complex(got[f"e0r[{k2 - 1}]"], got[f"e0i[{k2 - 1}]"]) - ra[k2]))
        for k2 in range(8):
            err = max(err, abs(complex(got[f"e16r[{k2}]"], got[f"e16i[{k2}]"]) - rb[k2]))
        return err
    return c, check


def make_w_pass1():
    c = Codelet("w_pass1", "recipe W pass 1: real DFT-25 of y[0..24] -> re/im[k1], k1 = 0..12 (im[0] = 0 is not written)")
    xs = c.real_in("y", 25)
    c.out_arrays(("re", 13), ("im", 13))
    ys = dft(xs)
    for k in range(13):
        c.emit_out(f"re[{k}]", ys[k].re)
        if k:
            c.emit_out(f"im[{k}]", ys[k].im)

    def check(rng):
        y = rng.standard_normal(25)
        got = c.evaluate({f"y[{i}]": y[i] for i in range(25)})
        ref = np.fft.fft(y)
        return max(abs(complex(got[f"re[{k}]"], got.get(f"im[{k}]", 0.0)) - ref[k]) for k in range(13))
    return c, check


def make_w_pass2_edge():
    c = Codelet("w_pass2_edge", "recipe W pass 2, row k1 = 0 (real): a[n2] = R[0][n2]; er/ei[k2] = X[25 k2], k2 = 0..8 "
                                "(ei[0] = ei[8] = 0 are written as zeros)")
    a = c.real_in("a", 16)
    c.out_arrays(("er", 9), ("ei", 9))
    ya = dft(a)
    for k2 in range(9):
        c.emit_out(f"er[{k2}]", ya[k2].re)
        c.emit_out(f"ei[{k2}]", ya[k2].im)

    def check(rng):
        av = rng.standard_normal(16)
        got = c.evaluate({f"a[{i}]": av[i] for i in range(16)})
        ref = np.fft.fft(av)
        return max(abs(complex(got[f"er[{k}]"], got[f"ei[{k}]"]) - ref[k]) for k in range(9))
    return c, check


HEADER = """// GENERATED by tools/gen_codelets.py -- do not edit; re-run the generator instead.
//
// Straight-line small-transform codelets for the frame-per-lane log-mel kernels (fbank_k.cu,
// logmel_w.cu).  __host__ __device__ so that tests/test_codelets.py can run them through g++.
#pragma once
#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace stx {
namespace codelets {

"""


def main():
    rng = np.random.default_rng(0)
    parts = [HEADER]
    for make in (make_dft16, make_k_pass1, make_k_pass2_edge, make_w_pass1, make_w_pass2_edge):
        c, check = make()
        err = max(check(rng) for _ in range(8))
        assert err < 1e-13, (c.name, err)
        print(f"{c.name:14s} {c.op_count():4d} ops ({c.fused_op_count()} after FMA contraction)   max |err| vs numpy.fft = {err:.1e}", file=sys.stderr)
        parts.append(c.source())
        parts.append("\n\n")
    parts.append("}  // namespace codelets\n}  // namespace stx\n")
    OUT.write_text("".join(parts))
    print(OUT)


if __name__ == "__main__":
    main()
