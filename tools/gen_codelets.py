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
    """Real-valued expression DAG whose values carry a LAZY SCALE: a value is None (an exact zero) or a pair
    (scale, node) standing for scale * node.  Multiplying by a constant only changes the scale; an addition of two
    values with different scales becomes ONE fused multiply-add, s1 * (n1 + (s2 / s1) * n2), whose scale stays pending
    and is absorbed by the next addition (or paid as one multiply at an output).  A complex twiddle multiplication
    followed by a butterfly therefore costs 6 FMAs instead of 4 + 4 operations (the tangent form of the twiddle falls
    out of the rule), and negations never cost anything."""

    def __init__(self):
        self.nodes = []          # ("in", name) | ("add", a, b) | ("sub", a, b) | ("fma", c, a, b) = c * a + b
        self.memo = {}

    def _mk(self, *key):
        if key[0] == "add" and key[1] > key[2]:
            key = ("add", key[2], key[1])
        hit = self.memo.get(key)
        if hit is not None:
            return hit
        self.nodes.append(key)
        self.memo[key] = len(self.nodes) - 1
        return len(self.nodes) - 1

    def inp(self, name):
        return (1.0, self._mk("in", name))

    @staticmethod
    def neg(a):
        return None if a is None else (-a[0], a[1])

    @staticmethod
    def mul(c, a):
        if a is None or c == 0.0:
            return None
        return (c * a[0], a[1])

    def add(self, a, b):
        if a is None:
            return b
        if b is None:
            return a
        (s1, n1), (s2, n2) = a, b
        if n1 == n2:
            return None if s1 + s2 == 0.0 else (s1 + s2, n1)
        if s1 == s2:
            return (s1, self._mk("add", n1, n2))
        if s1 == -s2:
            return (s1, self._mk("sub", n1, n2))
        # pivot: a unit scale if there is one (the result then needs no multiply at an output), else the larger one
        # (|ratio| <= 1: the tangent, not the cotangent)
        if abs(s2) == 1.0 and abs(s1) != 1.0 or (abs(s1) != 1.0 and abs(s2) > abs(s1)):
            s1, n1, s2, n2 = s2, n2, s1, n1
        return (s1, self._mk("fma", s2 / s1, n2, n1))

    def sub(self, a, b):
        return self.add(a, self.neg(b))


class Cx:
    """Complex value over a Graph: (re, im) lazily scaled values."""
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
    if r == 3:
        c, sn = -0.5, math.sin(2 * math.pi / 3)
        t1, t2 = xs[1] + xs[2], (xs[1] - xs[2]).scale(sn)
        m1 = xs[0] + t1.scale(c)
        return [xs[0] + t1, m1 + t2.mul_mj(), m1 + t2.mul_j()]
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
    """Forward DFT of a list of Cx by decimation in time.  Real inputs: only the columns k <= m/2 of the last stage are
    computed (with all their outputs); every other output is the conjugate of one of those."""
    n = len(xs)
    if n == 1:
        return list(xs)
    real = is_real(xs)
    if n in (2, 3, 4, 5):
        out = butterfly(xs)
    else:
        r = 4 if n % 4 == 0 else 2 if n % 2 == 0 else 5 if n % 5 == 0 else 3
        assert n % r == 0, n
        m = n // r
        subs = [dft(xs[q::r]) for q in range(r)]
        out = [NOT_COMPUTED] * n
        for k in range(m):
            if real and 2 * k > m:
                continue
            col = butterfly([subs[q][k].cmul(w(n, q * k)) for q in range(r)])
            for j in range(r):
                out[k + m * j] = col[j]
        for i in range(n):
            if out[i] is NOT_COMPUTED:
                out[i] = out[(n - i) % n].conj()
    if real:
        # exact symmetry: share the nodes (and make the imaginary parts of bins 0 and n/2 exact zeros)
        for k in range(n // 2 + 1, n):
            out[k] = out[n - k].conj()
        out[0] = Cx(out[0].g, out[0].re, None)
        if n % 2 == 0:
            out[n // 2] = Cx(out[0].g, out[n // 2].re, None)
    return out


NOT_COMPUTED = object()


class Codelet:
    def __init__(self, name, doc):
        self.name, self.doc = name, doc
        self.g = Graph()
        self.params = []         # (c type prefix, name, length, is_output)
        self.outputs = []        # (lvalue, node or None)
        self.exported = {}       # array name -> [scale per index]: pending output scales the CALLER folds into its own
                                 # constants (the twiddle table) instead of the codelet paying a multiply for them

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

    def emit_out_scaled(self, array, index, count, node):
        """Output array[index] WITHOUT its pending scale; the scale goes to the constexpr table <codelet>_<array>_scale."""
        tab = self.exported.setdefault(array, [1.0] * count)
        if node is not None:
            tab[index] = abs(node[0])
            node = (1.0 if node[0] > 0 else -1.0, node[1])
        self.outputs.append((f"{array}[{index}]", node))

    # ---- evaluation / emission ----
    def _reachable(self):
        need = set()
        stack = [v[1] for _, v in self.outputs if v is not None]
        while stack:
            i = stack.pop()
            if i in need:
                continue
            need.add(i)
            node = self.g.nodes[i]
            if node[0] in ("add", "sub"):
                stack += [node[1], node[2]]
            elif node[0] == "fma":
                stack += [node[2], node[3]]
        return sorted(need)

    def op_count(self):
        """Arithmetic instructions: one per add / sub / fma node plus one multiply per output whose pending scale is
        not +-1 (a sign is folded into the consumer)."""
        ops = sum(1 for i in self._reachable() if self.g.nodes[i][0] != "in")
        return ops + sum(1 for _, v in self.outputs if v is not None and abs(v[0]) != 1.0)

    def fma_count(self):
        return sum(1 for i in self._reachable() if self.g.nodes[i][0] == "fma")

    def evaluate(self, env):
        """env: {"name[i]": value}; returns {lvalue: value} in float64 (numpy scalars)."""
        val = {}
        for i in self._reachable():
            node = self.g.nodes[i]
            if node[0] == "in":
                val[i] = np.float64(env[node[1]])
            elif node[0] == "add":
                val[i] = val[node[1]] + val[node[2]]
            elif node[0] == "sub":
                val[i] = val[node[1]] - val[node[2]]
            else:
                val[i] = np.float64(node[1]) * val[node[2]] + val[node[3]]
        res = {lv: (np.float64(v[0]) * val[v[1]] if v is not None else np.float64(0.0)) for lv, v in self.outputs}
        for array, tab in self.exported.items():
            for i, sc in enumerate(tab):
                if f"{array}[{i}]" in res:
                    res[f"{array}[{i}]"] = res[f"{array}[{i}]"] * np.float64(sc)
        return res

    def source(self):
        lines = [f"// {self.doc}  [{self.op_count()} arithmetic ops, {self.fma_count()} of them FMAs]",
                 "template <typename T>",
                 f"__host__ __device__ __forceinline__ void {self.name}(" +
                 ", ".join(f"{ty} (&{nm})[{cnt}]" for ty, nm, cnt, _ in self.params) + ") {"]
        name = {}
        for i in self._reachable():
            node = self.g.nodes[i]
            if node[0] == "in":
                name[i] = node[1]
                continue
            name[i] = f"t{i}"
            if node[0] == "add":
                rhs = f"{name[node[1]]} + {name[node[2]]}"
            elif node[0] == "sub":
                rhs = f"{name[node[1]]} - {name[node[2]]}"
            else:
                rhs = f"fma_(T({node[1]!r}), {name[node[2]]}, {name[node[3]]})"
            lines.append(f"    const T t{i} = {rhs};")
        for lv, v in self.outputs:
            if v is None:
                rhs = "T(0)"
            elif v[0] == 1.0:
                rhs = name[v[1]]
            elif v[0] == -1.0:
                rhs = f"-{name[v[1]]}"
            else:
                rhs = f"T({v[0]!r}) * {name[v[1]]}"
            lines.append(f"    {lv} = {rhs};")
        lines.append("}")
        for array, tab in self.exported.items():
            lines.append(f"// true value of {array}[k] = {self.name}_{array}_scale[k] * (what {self.name} writes): fold it into the next constant")
            lines.append(f"constexpr double {self.name}_{array}_scale[{len(tab)}] = {{" + ", ".join(repr(v) for v in tab) + "};")
            lines.append(f"__host__ __device__ constexpr double {self.name}_{array}_scale_of(int k) {{   // usable in device code")
            lines.append(f"    constexpr double t[{len(tab)}] = {{" + ", ".join(repr(v) for v in tab) + "};")
            lines.append("    return t[k];")
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
    c = Codelet("w_pass1", "recipe W pass 1: real DFT-25 of y[0..24] -> re[k1], im[k1] / w_pass1_im_scale[k1], k1 = 0..12 (im[0] = 0 is not written)")
    xs = c.real_in("y", 25)
    c.out_arrays(("re", 13), ("im", 13))
    ys = dft(xs)
    for k in range(13):
        c.emit_out(f"re[{k}]", ys[k].re)
        if k:
            c.emit_out_scaled("im", k, 13, ys[k].im)

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

#include <cmath>

namespace stx {
namespace codelets {

__host__ __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
__host__ __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }

"""


def main():
    rng = np.random.default_rng(0)
    parts = [HEADER]
    for make in (make_dft16, make_k_pass1, make_k_pass2_edge, make_w_pass1, make_w_pass2_edge):
        c, check = make()
        err = max(check(rng) for _ in range(8))
        assert err < 1e-13, (c.name, err)
        print(f"{c.name:14s} {c.op_count():4d} ops ({c.fma_count()} FMAs)   max |err| vs numpy.fft = {err:.1e}", file=sys.stderr)
        parts.append(c.source())
        parts.append("\n\n")
    parts.append("}  // namespace codelets\n}  // namespace stx\n")
    OUT.write_text("".join(parts))
    print(OUT)


if __name__ == "__main__":
    main()
