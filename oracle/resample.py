"""Resampling oracle: the step before the log-mel path.

The reference resamples on the host with ``librosa.resample(audio_array, orig_sr=orig_sr, target_sr=16000)``
(R/processor.py:82-86; librosa 0.10.1 per R/pyproject.toml:10).  librosa's default ``res_type="soxr_hq"`` runs inside
the ``soxr`` C library, which is neither vendored in the reference nor installed offline (and ``librosa`` itself is
absent), so that arithmetic can be neither imported nor pinned here: PARITY WITH soxr_hq IS UNPINNED.  What is restated
and pinned instead is librosa's documented ``res_type="polyphase"`` path, i.e. ``scipy.signal.resample_poly(y,
target_sr // gcd, orig_sr // gcd)`` followed by ``fix_length`` to ``ceil(n * target_sr / orig_sr)`` samples
(librosa/core/audio.py, ``resample``: the polyphase branch and ``fix=True``); scipy IS importable (1.18.1), so this
restatement is checked against it live and through ``tests/golden/resample.npz``.

Algorithm (scipy/signal/_signaltools.py ``resample_poly`` with its defaults window=("kaiser", 5.0),
padtype="constant"; scipy/signal/_fir_filter_design.py ``firwin``; scipy/signal/_upfirdn.py):

  g = gcd(up, down); up //= g; down //= g; n_out = ceil(n * up / down)
  max_rate = max(up, down); f_c = 1 / max_rate; half_len = 10 * max_rate
  h = firwin(2 * half_len + 1, f_c, window=("kaiser", 5.0))      float64: sinc low-pass * Kaiser, unit DC gain
  h = float32(h) * float32(up)                                     the filter takes the dtype of x
  n_pre_pad = down - half_len % down;  n_pre_remove = (half_len + n_pre_pad) // down
  y_full[t] = sum_k h_pad[k] * x_up[t * down - k]                  x_up = x with up - 1 zeros after every sample
  y = y_full[n_pre_remove : n_pre_remove + n_out]

scipy accumulates in float32 in tap order; this oracle accumulates in float64 (the exact value of the float32 filter
applied to the float32 signal) and rounds once, so it differs from scipy by float32 summation noise only (<= 1e-6 for
|x| <= 1), which is the tolerance of the parity tests.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import math

import numpy as np


def plan(orig_sr: int, target_sr: int):
    """(up, down) after removing the common factor, like resample_poly."""
    g = math.gcd(int(orig_sr), int(target_sr))
    return int(target_sr) // g, int(orig_sr) // g


def out_length(n: int, up: int, down: int) -> int:
    return (n * up + down - 1) // down


def design(up: int, down: int):
    """(h_pad float32 [len], n_pre_remove): the padded, scaled filter resample_poly hands to upfirdn."""
    max_rate = max(up, down)
    f_c = 1.0 / max_rate
    half_len = 10 * max_rate
    numtaps = 2 * half_len + 1
    alpha = 0.5 * (numtaps - 1)
    m = np.arange(numtaps, dtype=np.float64) - alpha
    h = f_c * np.sinc(f_c * m)                                     # firwin, pass_zero low-pass: right * sinc(right * m)
    n = np.arange(numtaps, dtype=np.float64)
    win = np.i0(5.0 * np.sqrt(np.maximum(0.0, 1.0 - ((n - alpha) / alpha) ** 2))) / np.i0(5.0)   # windows.kaiser(M, 5, sym=True)
    h = h * win
    h = h / h.sum()                                                # unit gain at DC (scale=True)
    h32 = h.astype(np.float32) * np.float32(up)
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    h_pad = np.concatenate([np.zeros(n_pre_pad, np.float32), h32])
    return h_pad, n_pre_remove


def resample_poly(x: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """float32 [ceil(n * target_sr / orig_sr)] = librosa.resample(x, orig_sr, target_sr, res_type="polyphase")."""
    x = np.asarray(x, dtype=np.float32).reshape(-1)
    up, down = plan(orig_sr, target_sr)
    if up == down:
        return x.copy()
    n = x.size
    n_out = out_length(n, up, down)
    h_pad, n_pre_remove = design(up, down)
    L = h_pad.size
    J = (L + up - 1) // up
    hb = np.zeros((up, J), np.float64)                              # polyphase bank: hb[phase][j] = h_pad[phase + j * up]
    for ph in range(up):
        taps = h_pad[ph::up]
        hb[ph, :taps.size] = taps
    t = np.arange(n_out, dtype=np.int64) + n_pre_remove
    p = t * down
    i0 = p // up
    ph = (p - i0 * up).astype(np.int64)
    xp = np.concatenate([np.zeros(J, np.float64), x.astype(np.float64), np.zeros(J + 1, np.float64)])
    y = np.zeros(n_out, np.float64)
    for j in range(J):                                              # y[t] = sum_j hb[ph][j] * x[i0 - j]
        idx = np.clip(i0 - j, -J, n) + J
        y += hb[ph, j] * xp[idx]
    return y.astype(np.float32)
