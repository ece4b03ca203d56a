"""DFT-as-GEMM on tcgen05 at recipe W's shape, measured with the repository's own error-compensated tensor-core contraction.

north_star names two engines for the transform: a shared-memory mixed-radix FFT or a DFT-as-GEMM on tcgen05 with
error-compensated split operands, "the choice evidenced".  The cosine kernel (csrc/cosine.cu: bfloat16 hi/lo planes of both
operands, three products, float32 accumulation in tensor memory) IS such a contraction, so the DFT stage of recipe W can be timed
with it directly: frames [F, 400] (already framed and windowed: the GEMM gets this for free here) x basis [402, 400]
(cos rows 0..200, sin rows 0..200) -> [F, 402].  F = 191 872 is cfg2 (64 clips x 2 998 frames).  The row normalisation of the
cosine call is undone on the host for the accuracy figure; its cost (one pass over the operands, fused with the hi/lo split) is
what any DFT-as-GEMM would pay to stage its operands.

    python tools/bench_dft_gemm.py            # one JSON line: ms per cfg2 batch, error vs a float64 DFT, w_frames for comparison
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from speech_transcript_embeddings_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
F, NFFT = 64 * 2998, 400
rng = np.random.default_rng(0)
n = np.arange(NFFT)
win = 0.5 - 0.5 * np.cos(2 * np.pi * n / NFFT)                       # periodic Hann (TF/models/whisper/feature_extraction_whisper.py:135-164)
k = np.arange(201)[:, None]
basis = np.concatenate([np.cos(2 * np.pi * k * n / NFFT), -np.sin(2 * np.pi * k * n / NFFT)], 0)     # [402, 400] float64
basis[201] = basis[0] * 0 + 1e-30                                    # sin rows 0 and 200 are zero: keep their norms finite
basis[401] = basis[201]
frames = (rng.standard_normal((F, NFFT)) * win).astype(np.float32)   # windowed frames, float32 like the recipe's input
A = torch.from_numpy(frames).to(dev)
Bm = torch.from_numpy(basis.astype(np.float32)).to(dev)
out = torch.empty((F, 402), dtype=torch.float32, device=dev)


def step():
    return ops.cosine_nxm(A, Bm, always_normalize=True, out=out)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 10
e0.record()
for _ in range(iters):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters

# accuracy on a sample of frames: undo the normalisation, compare with the float64 DFT of the same float32 frames
idx = rng.choice(F, 512, replace=False)
S = out[torch.from_numpy(idx).to(dev)].double().cpu().numpy()
na = np.linalg.norm(frames[idx].astype(np.float64), axis=1)[:, None]
nb = np.linalg.norm(basis.astype(np.float32).astype(np.float64), axis=1)[None, :]
Y = S * na * nb
ref = frames[idx].astype(np.float64) @ basis.astype(np.float32).astype(np.float64).T
rms = np.sqrt((frames[idx].astype(np.float64) ** 2).sum(1))[:, None]  # the frame's amplitude scale (|x| * |basis row| ~ rms * sqrt(200))
err_rel_frame = float(np.max(np.abs(Y - ref) / (rms * np.sqrt(NFFT / 2))))

res = {"what": "DFT stage of recipe W as a tcgen05 split-bf16 GEMM (repo's cosine contraction), cfg2 frame count",
       "frames": F, "n_fft": NFFT, "basis_rows": 402, "ms_dft_gemm": ms,
       "flop_algorithmic": 2.0 * F * NFFT * 402, "tflops_algorithmic": 2.0 * F * NFFT * 402 / (ms * 1e-3) / 1e12,
       "max_abs_err_relative_to_frame_amplitude": err_rel_frame,
       "output_bytes": F * 402 * 4,
       "note": "transform only: framing / reflect padding / window before it and power / mel / log10 after it are not included; "
               "w_frames (FFT engine, everything included) takes 0.107 ms on the same batch"}
print(json.dumps(res))
