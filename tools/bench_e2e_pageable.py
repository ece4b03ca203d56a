"""End-to-end sweep of the reference-signature call (a list of pageable float32 arrays in, CPU tensors out): packing threads
and pipeline chunk size.

    python tools/bench_e2e_pageable.py [K|W]
"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200.feature_extraction import (B200SeamlessM4TFeatureExtractor,  # noqa: E402
                                                                  B200WhisperFeatureExtractor)

recipe = sys.argv[1] if len(sys.argv) > 1 else "K"
dev = torch.device("cuda", 0)
B, n = 64, 480000
rng = np.random.default_rng(0)
clips = [(0.1 * rng.standard_normal(n)).astype(np.float32) for _ in range(B)]
fe = (B200SeamlessM4TFeatureExtractor if recipe == "K" else B200WhisperFeatureExtractor)(device=dev)


def timed(f, reps=10):
    for _ in range(4):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print("cores", len(os.sched_getaffinity(0)))
for threads in (2, 4, 6, 8, 12, 16):
    os.environ["STX_PACK_THREADS"] = str(threads)
    for mb in (6, 12, 24):
        type(fe).CHUNK_BYTES = mb << 20
        ms = timed(lambda: fe(clips, sampling_rate=16000, return_tensors="pt", output="host"))
        print(f"pack threads {threads:2d}  chunk {mb:3d} MB: {ms:.3f} ms/step = {B * 30 / ms * 1e3:,.0f} audio-s/s", flush=True)
