"""Seeded synthetic 16 kHz mono PCM and embedding generators.

Signal classes (float32, generated on the CPU with ``torch.Generator`` so every
host produces identical bits):

  G      0.1 * N(0,1)                          U      Uniform(-1, 1)
  AM     0.3 * N(0,1) * (0.5+0.5 sin 2pi 3t)^2  HS     G in the first half, digital zeros after
  small  1e-4 * N(0,1)                         loud   3 * N(0,1)  (triggers peak-normalise)
  tone   0.5 * sin(2pi 440 t)                  chirp  linear 50 Hz -> 7 kHz sweep

``tone`` and ``chirp`` are ill-conditioned in the CPU reference itself (stationary
spectra make the per-bin variance vanish) and are reported, not parity-gated.
"""
from __future__ import annotations

import numpy as np
import torch

SAMPLE_RATE = 16000
GATED_CLASSES = ("G", "U", "AM", "HS", "small", "loud")
REPORTED_CLASSES = ("tone", "chirp")


def clip(kind: str, n: int, seed: int) -> np.ndarray:
    g = torch.Generator().manual_seed(int(seed))
    t = torch.arange(n, dtype=torch.float64) / SAMPLE_RATE
    if kind == "G":
        x = 0.1 * torch.randn(n, generator=g)
    elif kind == "U":
        x = torch.rand(n, generator=g) * 2.0 - 1.0
    elif kind == "AM":
        env = (0.5 + 0.5 * torch.sin(2 * np.pi * 3.0 * t)) ** 2
        x = 0.3 * torch.randn(n, generator=g) * env.float()
    elif kind == "HS":
        x = 0.1 * torch.randn(n, generator=g)
        x[n // 2:] = 0.0
    elif kind == "small":
        x = 1e-4 * torch.randn(n, generator=g)
    elif kind == "loud":
        x = 3.0 * torch.randn(n, generator=g)
    elif kind == "tone":
        x = (0.5 * torch.sin(2 * np.pi * 440.0 * t)).float()
    elif kind == "chirp":
        dur = max(n / SAMPLE_RATE, 1e-9)
        x = (0.5 * torch.sin(2 * np.pi * (50.0 * t + 0.5 * (7000.0 - 50.0) / dur * t * t))).float()
    else:
        raise ValueError(f"unknown signal class {kind!r}")
    return x.to(torch.float32).numpy()


def batch_fixed(num_clips: int, seconds: float = 30.0, kind: str = "G", seed0: int = 0):
    """cfg1/cfg2: ``num_clips`` clips of equal length, seeds seed0..seed0+num_clips-1."""
    n = int(round(seconds * SAMPLE_RATE))
    return [clip(kind, n, seed0 + i) for i in range(num_clips)]


def variable_lengths(num_clips: int, seed: int = 1234, whole_seconds: bool = True,
                     min_s: int = 1, max_s: int = 30) -> np.ndarray:
    """cfg3 lengths in samples: 16000*Uniform{1..30}, or arbitrary (odd-T included)."""
    g = torch.Generator().manual_seed(int(seed))
    if whole_seconds:
        return (torch.randint(min_s, max_s + 1, (num_clips,), generator=g) * SAMPLE_RATE).numpy()
    return torch.randint(min_s * SAMPLE_RATE, max_s * SAMPLE_RATE + 1, (num_clips,), generator=g).numpy()


def batch_variable(num_clips: int, seed: int = 1234, whole_seconds: bool = True, kind: str = "G",
                   min_s: int = 1, max_s: int = 30):
    lens = variable_lengths(num_clips, seed, whole_seconds, min_s, max_s)
    return [clip(kind, int(n), seed + 1 + i) for i, n in enumerate(lens)]


def embedding_pairs(n: int, d: int, seed: int = 0, noise: float = 0.5):
    """cfg5: A = normalize(N(0,1)), B = normalize(A + noise*N(0,1)) (correlated pairs)."""
    g = torch.Generator().manual_seed(int(seed))
    a = torch.randn(n, d, generator=g)
    a = a / a.norm(dim=1, keepdim=True)
    b = a + noise * torch.randn(n, d, generator=g)
    b = b / b.norm(dim=1, keepdim=True)
    return a.numpy().astype(np.float32), b.numpy().astype(np.float32)
