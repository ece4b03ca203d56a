"""Runs the device-resident hot path a few times (for ncu / compute-sanitizer captures).

    python tools/run_path.py [K|W|C] [clips] [seconds] [iters]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200 import ops, synth  # noqa: E402

recipe = sys.argv[1] if len(sys.argv) > 1 else "K"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda", 0)
if recipe == "C":
    a, b = synth.embedding_pairs(4096, 768, seed=0)
    a, b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    for _ in range(iters):
        S = ops.cosine_nxm(a, b)
    torch.cuda.synchronize()
    print("cosine", tuple(S.shape), float(S.diagonal().mean()))
    sys.exit(0)
n = int(secs * 16000)
g = torch.Generator(device=dev).manual_seed(0)
pcm = 0.1 * torch.randn(B * n, generator=g, device=dev)
off = (torch.arange(B, device=dev, dtype=torch.int64) * n)
ln = torch.full((B,), n, dtype=torch.int32, device=dev)
for _ in range(iters):
    if recipe == "K":
        T_pad = 2 * ((ops.k_num_frames(n) + 1) // 2)
        x, m = ops.fbank_k(pcm, off, ln, n, T_pad, uniform=True)
    else:
        x, m = ops.logmel_w(pcm, off, ln, n)
torch.cuda.synchronize()
print(recipe, tuple(x.shape), float(x.float().abs().mean()))
