"""stx_fbank_k_projection (PCM -> Linear(LayerNorm(input_features)) without the normalised features going through HBM) on
random batches: ragged lengths incl. clips without a single frame, both padding values, with and without the optional
features output.  Checks: features bit-identical to stx_fbank_k, masks equal, hidden states equal with and without the
features output, <= 2e-5 from the two-step path (extractor, then stx_feature_projection) and from torch's float64 LayerNorm +
Linear.  Run under a timeout by tests/test_fused_projection_gpu.py.

    python tests/scripts/fused_projection_check.py [rounds]
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from speech_transcript_embeddings_b200 import ops, synth  # noqa: E402
from speech_transcript_embeddings_b200.feature_extraction import _layout  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
g = torch.Generator().manual_seed(3)
w = (0.05 * torch.randn(1024, 160, generator=g)).to(dev)
bias = (0.1 * torch.randn(1024, generator=g)).to(dev)
gamma = (1.0 + 0.1 * torch.randn(160, generator=g)).to(dev)
beta = (0.1 * torch.randn(160, generator=g)).to(dev)
t0 = time.time()
worst = worst64 = 0.0
for it in range(rounds):
    B = int(rng.integers(1, 40))
    lens_np = rng.integers(300, 16000 * int(rng.integers(1, 12)) + 1, size=B).astype(np.int32)
    uniform = it % 7 == 0
    if uniform:
        lens_np[:] = 16000 * 5
    clips = [synth.clip("G", int(n), int(rng.integers(0, 1 << 30))) for n in lens_np]
    offsets, total = _layout(lens_np)
    host = np.zeros(total, np.float32)
    for c, o in zip(clips, offsets):
        host[o:o + c.size] = c
    pcm, off, lens = torch.from_numpy(host).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(lens_np).to(dev)
    frames = np.array([ops.k_num_frames(int(n)) for n in lens_np])
    T_pad = int(frames.max() + (frames.max() & 1))
    if T_pad == 0:
        continue
    pv, ml = float(it % 2), int(lens_np.max())
    h1, f1, m1 = ops.fbank_k_projection(pcm, off, lens, ml, T_pad, gamma, beta, w, bias, padding_value=pv, want_features=True,
                                        uniform=uniform)
    h2, f2, m2 = ops.fbank_k_projection(pcm, off, lens, ml, T_pad, gamma, beta, w, None if it % 3 == 0 else bias, padding_value=pv)
    f0, m0 = ops.fbank_k(pcm, off, lens, ml, T_pad, padding_value=pv)
    two, _ = ops.feature_projection(f0, gamma, beta, w, bias)
    ok = torch.isfinite(two)                                           # one-frame clips are NaN, like NumPy's variance
    assert f2 is None and bool(((f1 == f0) | (torch.isnan(f1) & torch.isnan(f0))).all())
    assert torch.equal(m1, m0) and torch.equal(m2, m0)
    if it % 3 != 0:
        assert bool(((h1 == h2) | ~ok).all())
    worst = max(worst, float((h1 - two)[ok].abs().max()))
    ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(f0.double(), (160,), gamma.double(), beta.double()),
                                     w.double(), bias.double())
    worst64 = max(worst64, float((h1.double() - ref)[ok].abs().max()))
torch.cuda.synchronize()
assert worst <= 2e-5 and worst64 <= 2e-5, (worst, worst64)
print(f"fused projection ok: {rounds} rounds, worst vs two-step {worst:.2e}, vs float64 {worst64:.2e}, {time.time() - t0:.1f} s")
