"""Host-side overhead of the chunked H2D | kernels | D2H pipeline: 66 one-second clips (the device work is negligible) in 1, 11 and
22 chunks.  Measured on the round-2 box: 0.51 / 1.55 / 2.73 ms per call, i.e. ~0.5 ms per call + ~0.10 ms per chunk of Python,
ctypes and torch stream / event calls.  On the cfg2 batch that overhead hides under the 0.22 ms a 12 MB chunk needs on PCIe: 6 MB
chunks (21 of them) are as fast as 12 MB chunks (tools/bench_e2e_pageable.py), so the pipeline is bound by the host's memory
system (pack + DMA), not by the loop that drives it.

    python tools/bench_e2e_overhead.py
"""
import sys, time, numpy as np, torch
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200.feature_extraction import B200SeamlessM4TFeatureExtractor
dev = torch.device("cuda", 0)
fe = B200SeamlessM4TFeatureExtractor(device=dev)
rng = np.random.default_rng(0)
def timed(f, reps=20):
    for _ in range(5): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
# tiny clips, many chunks: host overhead per chunk
clips = [(0.1 * rng.standard_normal(16000)).astype(np.float32) for _ in range(66)]
for nch, cb in ((1, 1 << 30), (11, 6 * 16000 * 4), (22, 3 * 16000 * 4)):
    type(fe).CHUNK_BYTES = cb; type(fe).FIRST_CHUNK_BYTES = cb; type(fe).PACK_INLINE_BYTES = 0
    ms = timed(lambda: fe(clips, sampling_rate=16000, return_tensors="pt", output="host"))
    print(f"66 x 1 s clips, ~{nch} chunks: {ms:.3f} ms per call")
