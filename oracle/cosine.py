"""Cosine-scoring oracle.

  pairwise   R/processor.py:148-159 (AudioTextProcessor.compute_similarity: re-normalise
             only when ||e|| deviates from 1 by more than 1e-4, then sum(e1*e2, dim=1)),
             same arithmetic at R/inference.py:121, R/cv_inference.py:105,
             R/training/trainer_unfreeze.py:1073-1074 (after F.normalize, R/model.py:326-327)
  matrix     north_star's N x M superset: S = normalize(A) @ normalize(B).T in float64;
             its diagonal must equal the pairwise scores.

F.normalize(x, p=2, dim=1) = x / max(||x||_2, 1e-12).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np


def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    return x / np.maximum(np.sqrt((x * x).sum(axis=1, keepdims=True)), eps)


def _maybe_normalize_f32(e: np.ndarray) -> np.ndarray:
    """The reference's conditional re-normalisation, in float32 like torch does it."""
    e = np.asarray(e, dtype=np.float32)
    nrm = np.sqrt((e * e).sum(axis=1, dtype=np.float32))
    if not np.all(np.abs(nrm - np.float32(1.0)) <= np.float32(1e-4) + np.float32(1e-5) * np.float32(1.0)):
        e = e / np.maximum(nrm, np.float32(1e-12))[:, None]
    return e


def pairwise_reference(e1: np.ndarray, e2: np.ndarray) -> np.ndarray:
    """float32 [N], following compute_similarity step by step."""
    a = _maybe_normalize_f32(e1)
    b = _maybe_normalize_f32(e2)
    return (a * b).sum(axis=1, dtype=np.float32)


def pairwise_f64(e1: np.ndarray, e2: np.ndarray) -> np.ndarray:
    return (l2_normalize(e1) * l2_normalize(e2)).sum(axis=1)


def matrix_f64(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return l2_normalize(a) @ l2_normalize(b).T


def pos_neg_reference(aud: np.ndarray, pos: np.ndarray, neg: np.ndarray, temperature: float = 0.1,
                      corrupt_gamma: float = 0.35, alignment_factor: np.ndarray | None = None) -> dict:
    """The evaluation-time scoring consumers, float64:
      s_pos / s_neg     (aud * txt).sum(dim=1) after F.normalize      R/training/trainer_unfreeze.py:561-563, 1206-1207
      hr_pos / hr_neg   to_human_readable(s, temperature, "prob") = sigmoid(s / temperature)      :924-939, 1215-1216
      per_sample        F.cross_entropy(stack([s_pos, s_neg], 1) / temperature, target 0, reduction="none")  :722-726
                        times the optional alignment factor 1 - sigmoid(mean_align) * alignment_weight      :729-733
      loss              per_sample.mean() + corrupt_gamma * relu(s_neg).mean()                    :735-739
    """
    a, p, n = l2_normalize(aud), l2_normalize(pos), l2_normalize(neg)
    s_pos, s_neg = (a * p).sum(axis=1), (a * n).sum(axis=1)
    z = (s_neg - s_pos) / temperature
    per = np.logaddexp(0.0, z)                                  # log(1 + exp(l_neg - l_pos))
    if alignment_factor is not None:
        per = per * np.asarray(alignment_factor, np.float64)
    loss = per.mean() + (corrupt_gamma * np.maximum(s_neg, 0.0).mean() if corrupt_gamma > 0 else 0.0)
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    return {"s_pos": s_pos, "s_neg": s_neg, "hr_pos": sig(s_pos / temperature), "hr_neg": sig(s_neg / temperature),
            "per_sample": per, "loss": float(loss)}
