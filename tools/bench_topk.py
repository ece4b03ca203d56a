"""Retrieval (stx_cosine_topk) against the matrix + torch.topk, N = M = 4096, D = 768, k = 8.

    python tools/bench_topk.py
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
a, b = synth.embedding_pairs(4096, 768, seed=0)
a, b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)


def timed(f, iters=20):
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


fused = timed(lambda: ops.cosine_topk(a, b, 8))
S = torch.empty((4096, 4096), dtype=torch.float32, device=dev)
two_step = timed(lambda: torch.topk(ops.cosine_nxm(a, b, out=S), 8, dim=1))
matrix = timed(lambda: ops.cosine_nxm(a, b, out=S))
print(json.dumps({"workload": "top-8 of 4096 x 4096 x 768 cosine scores", "fused_ms": round(fused, 4),
                  "matrix_then_torch_topk_ms": round(two_step, 4), "matrix_only_ms": round(matrix, 4)}))
