"""Probe of the N x M cosine kernel on WELL-MATCHED pairs (cos = 1, 0.96, 0.82, 0.45): diagonal and off-diagonal error against
float64, and the pairwise (CUDA-core) path beside it.  The tensor core rounds toward zero after every MMA; with a single
accumulator that biased high scores by -1.8e-8 * D * score (cosine section).

    python tools/probe_cosine_bias.py
"""
import sys, numpy as np, torch
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200 import ops
dev = torch.device('cuda', 0)
rng = np.random.default_rng(0)
for D in (256, 768, 1024):
    a = rng.standard_normal((512, D)).astype(np.float32)
    for noise in (0.0, 0.3, 0.7, 2.0):
        b = (a + noise * rng.standard_normal((512, D))).astype(np.float32)
        S = ops.cosine_nxm(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)).cpu().numpy().astype(np.float64)
        an = a.astype(np.float64); an /= np.linalg.norm(an, axis=1, keepdims=True)
        bn = b.astype(np.float64); bn /= np.linalg.norm(bn, axis=1, keepdims=True)
        R = an @ bn.T
        d = np.diag(S) - np.diag(R)
        off = (S - R)[~np.eye(512, dtype=bool)]
        pw = ops.cosine_pairwise(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev), always_normalize=True).cpu().numpy().astype(np.float64)
        print(f"D={D} noise={noise}: cos~{np.diag(R).mean():.3f} diag err mean {d.mean():+.2e} max|.| {np.abs(d).max():.2e} | off-diag max {np.abs(off).max():.2e} | pairwise max {np.abs(pw-np.diag(R)).max():.2e}")
