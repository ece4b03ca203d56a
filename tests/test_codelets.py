"""The generated FFT codelets (csrc/codelets.cuh), compiled for the HOST with g++ and checked against numpy.fft.

The same header is what the CUDA kernels include; here it only proves that the straight-line code the
generator (tools/gen_codelets.py) wrote computes the transforms the kernels assume, in float64 and float32.
"""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HDR = ROOT / "speech_transcript_embeddings_b200" / "csrc" / "codelets.cuh"

HARNESS = r"""
#include "codelets.cuh"
using namespace stx::codelets;
template <typename T> static void run_dft16(const T* in, T* out) {
    T xr[16], xi[16], yr[16], yi[16];
    for (int i = 0; i < 16; ++i) { xr[i] = in[2 * i]; xi[i] = in[2 * i + 1]; }
    dft16<T>(xr, xi, yr, yi);
    for (int i = 0; i < 16; ++i) { out[2 * i] = yr[i]; out[2 * i + 1] = yi[i]; }
}
template <typename T> static void run_k_pass1(const T* in, T* out) {
    T y[25], re[17], im[17];
    for (int i = 0; i < 25; ++i) y[i] = in[i];
    for (int i = 0; i < 17; ++i) re[i] = im[i] = 0;
    k_pass1<T>(y, re, im);
    for (int i = 0; i < 17; ++i) { out[2 * i] = re[i]; out[2 * i + 1] = im[i]; }
}
template <typename T> static void run_k_pass2_edge(const T* in, T* out) {
    T a[16], r[16], e0r[7], e0i[7], e16r[8], e16i[8];
    for (int i = 0; i < 16; ++i) { a[i] = in[i]; r[i] = in[16 + i]; }
    k_pass2_edge<T>(a, r, e0r, e0i, e16r, e16i);
    for (int i = 0; i < 7; ++i) { out[2 * i] = e0r[i]; out[2 * i + 1] = e0i[i]; }
    for (int i = 0; i < 8; ++i) { out[14 + 2 * i] = e16r[i]; out[15 + 2 * i] = e16i[i]; }
}
template <typename T> static void run_w_pass1(const T* in, T* out) {
    T y[25], re[13], im[13];
    for (int i = 0; i < 25; ++i) y[i] = in[i];
    for (int i = 0; i < 13; ++i) re[i] = im[i] = 0;
    w_pass1<T>(y, re, im);
    for (int i = 0; i < 13; ++i) { out[2 * i] = re[i]; out[2 * i + 1] = im[i] * T(w_pass1_im_scale[i]); }   // exported scale
}
template <typename T> static void run_w_pass2_edge(const T* in, T* out) {
    T a[16], er[9], ei[9];
    for (int i = 0; i < 16; ++i) a[i] = in[i];
    w_pass2_edge<T>(a, er, ei);
    for (int i = 0; i < 9; ++i) { out[2 * i] = er[i]; out[2 * i + 1] = ei[i]; }
}
extern "C" {
#define BOTH(name) \
    void name##_f64(const double* in, double* out) { run_##name<double>(in, out); } \
    void name##_f32(const float* in, float* out) { run_##name<float>(in, out); }
BOTH(dft16) BOTH(k_pass1) BOTH(k_pass2_edge) BOTH(w_pass1) BOTH(w_pass2_edge)
}
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("codelets")
    src = d / "harness.cpp"
    src.write_text(HARNESS)
    so = d / "libcodelets_host.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", f"-I{HDR.parent}", str(src), "-o", str(so)],
                   check=True)
    return C.CDLL(str(so))


def _call(lib, name, x, n_out, dtype):
    suffix = "f64" if dtype == np.float64 else "f32"
    x = np.ascontiguousarray(x, dtype)
    out = np.zeros(n_out, dtype)
    getattr(lib, f"{name}_{suffix}")(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out.astype(np.float64)


def _cx(v):
    return v[0::2] + 1j * v[1::2]


@pytest.mark.parametrize("dtype,tol", [(np.float64, 2e-14), (np.float32, 2e-5)])
def test_codelets_match_numpy_fft(lib, dtype, tol):
    rng = np.random.default_rng(7)
    for _ in range(4):
        z = rng.standard_normal(16) + 1j * rng.standard_normal(16)
        got = _cx(_call(lib, "dft16", np.stack([z.real, z.imag], 1).ravel(), 32, dtype))
        assert np.abs(got - np.fft.fft(z)).max() <= tol

        y = rng.standard_normal(25)
        got = _cx(_call(lib, "k_pass1", y, 34, dtype))
        assert np.abs(got - np.fft.fft(np.concatenate([y, np.zeros(7)]))[:17]).max() <= tol

        a, r = rng.standard_normal(16), rng.standard_normal(16)
        got = _cx(_call(lib, "k_pass2_edge", np.concatenate([a, r]), 30, dtype))
        fa = np.fft.fft(a)
        fr = np.fft.fft(r * np.exp(-2j * np.pi * 16 * np.arange(16) / 512))
        assert np.abs(got[:7] - fa[1:8]).max() <= tol          # X[32 k2], k2 = 1..7
        assert np.abs(got[7:] - fr[:8]).max() <= tol           # X[16 + 32 k2], k2 = 0..7

        got = _cx(_call(lib, "w_pass1", y, 26, dtype))
        assert np.abs(got - np.fft.fft(y)[:13]).max() <= tol

        got = _cx(_call(lib, "w_pass2_edge", a, 18, dtype))
        assert np.abs(got - fa[:9]).max() <= tol


def test_two_pass_factorisation_is_the_real_fft():
    """The index maps the kernels use: n = 16 n1 + n2, k = k1 + 32 k2 (K) / k1 + 25 k2 (W), mirrored bins."""
    rng = np.random.default_rng(3)
    for nfft, n1_len in ((512, 32), (400, 25)):
        x = np.zeros(nfft)
        x[:400] = rng.standard_normal(400)
        ref = np.fft.fft(x)
        R = np.stack([np.fft.fft(x[n2::16]) for n2 in range(16)], 1)          # [k1, n2], pass 1 per n2
        tw = np.exp(-2j * np.pi * np.outer(np.arange(n1_len), np.arange(16)) / nfft)
        X = np.fft.fft(R * tw, axis=1)                                         # [k1, k2] = X[k1 + n1_len k2]
        for k1 in range(n1_len // 2 + 1):
            for k2 in range(16):
                k = k1 + n1_len * k2
                assert abs(X[k1, k2] - ref[k]) < 1e-9
                assert abs(abs(X[k1, k2]) - abs(ref[(nfft - k) % nfft])) < 1e-9   # mirrored bin, same power


def test_generated_header_is_current(tmp_path):
    """codelets.cuh is exactly what tools/gen_codelets.py writes."""
    sys.path.insert(0, str(ROOT / "tools"))
    import gen_codelets
    saved = gen_codelets.OUT
    try:
        gen_codelets.OUT = tmp_path / "codelets.cuh"
        gen_codelets.main()
        assert gen_codelets.OUT.read_text() == HDR.read_text()
    finally:
        gen_codelets.OUT = saved
