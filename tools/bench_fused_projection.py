"""PCM -> encoder input-stage hidden states on the cfg2 batch: extractor + stx_feature_projection against the fused
stx_fbank_k_projection.

    python tools/bench_fused_projection.py
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B, n = 64, 480000
pools = [0.1 * torch.randn(B * n, generator=torch.Generator(device=dev).manual_seed(s), device=dev) for s in range(3)]
off = torch.arange(B, device=dev, dtype=torch.int64) * n
ln = torch.full((B,), n, dtype=torch.int32, device=dev)
T_pad = 2 * ((ops.k_num_frames(n) + 1) // 2)
g = torch.Generator().manual_seed(0)
w = (0.05 * torch.randn(1024, 160, generator=g)).to(dev)
bias = torch.zeros(1024, device=dev)
gamma, beta = torch.ones(160, device=dev), torch.zeros(160, device=dev)


def timed(f, iters=10):
    for i in range(3):
        f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        f(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def two_step(i):
    f, _ = ops.fbank_k(pools[i % 3], off, ln, n, T_pad, uniform=True)
    return ops.feature_projection(f, gamma, beta, w, bias, return_norm=False)


def fused(i):
    return ops.fbank_k_projection(pools[i % 3], off, ln, n, T_pad, gamma, beta, w, bias, uniform=True)


print(json.dumps({"workload": "64 x 30 s PCM -> [64, 1499, 1024] hidden states", "two_step_ms": round(timed(two_step), 4),
                  "fused_ms": round(timed(fused), 4)}))
