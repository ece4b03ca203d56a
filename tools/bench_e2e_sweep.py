"""End-to-end sweep: chunk size of the H2D | kernels | D2H pipeline, and the raw PCIe copy rates beside it.

    python tools/bench_e2e_sweep.py [K|W]
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200.feature_extraction import (B200SeamlessM4TFeatureExtractor,  # noqa: E402
                                                                  B200WhisperFeatureExtractor, PackedClips)

recipe = sys.argv[1] if len(sys.argv) > 1 else "K"
dev = torch.device("cuda", 0)
B, n = 64, 480000
fe = (B200SeamlessM4TFeatureExtractor if recipe == "K" else B200WhisperFeatureExtractor)(device=dev)
pinned = torch.empty(B * n, dtype=torch.float32, pin_memory=True)
torch.randn(B * n, out=pinned, generator=torch.Generator().manual_seed(0))
pinned.mul_(0.1)
packed = PackedClips(pinned, np.arange(B, dtype=np.int64) * n, np.full(B, n, np.int32))


def timed(f, reps=8):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


d = torch.empty(B * n, dtype=torch.float32, device=dev)
h_out = torch.empty(B * 1499 * 160, dtype=torch.float32, pin_memory=True)
d_out = torch.empty(B * 1499 * 160, dtype=torch.float32, device=dev)
ms = timed(lambda: d.copy_(pinned, non_blocking=True))
print(f"H2D 123 MB pinned: {ms:.3f} ms = {B * n * 4 / ms / 1e6:.1f} GB/s")
ms = timed(lambda: h_out.copy_(d_out, non_blocking=True))
print(f"D2H 61 MB pinned: {ms:.3f} ms = {h_out.numel() * 4 / ms / 1e6:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        d.copy_(pinned, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


ms = timed(both)
print(f"H2D 123 MB || D2H 61 MB: {ms:.3f} ms")
for mb in (2, 4, 6, 8, 12, 16, 24, 32, 64, 128):
    type(fe).CHUNK_BYTES = mb << 20
    ms = timed(lambda: fe(packed, sampling_rate=16000, return_tensors="pt", output="host"))
    print(f"chunk {mb:4d} MB: {ms:.3f} ms/step = {B * 30 / ms * 1e3:,.0f} audio-s/s")
