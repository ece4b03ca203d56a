"""Per-kernel durations of the device-resident hot path on the cfg2 batch (CUDA events recorded inside the library around
every launch, on the launching stream).  For A/B runs of kernel variants: STX_B200_LIB=<other build> python tools/time_kernels.py

    python tools/time_kernels.py [K|W] [clips] [seconds] [iters]
"""
import json
import os
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200 import _lib, ops  # noqa: E402

recipe = sys.argv[1] if len(sys.argv) > 1 else "K"          # K | W | C (cosine matrix, N = M = clips * 64, D = 768)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
dev = torch.device("cuda", 0)
if recipe == "C":
    N = B * 64
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.nn.functional.normalize(torch.randn(N, 768, generator=g, device=dev), dim=1)
    b = torch.nn.functional.normalize(a + 0.5 * torch.randn(N, 768, generator=g, device=dev), dim=1)
    S = torch.empty((N, N), dtype=torch.float32, device=dev)
    for _ in range(5):
        ops.cosine_nxm(a, b, out=S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.cosine_nxm(a, b, out=S)
    e1.record()
    torch.cuda.synchronize()
    _lib.profile(True)
    for _ in range(iters):
        ops.cosine_nxm(a, b, out=S)
    torch.cuda.synchronize()
    per = {}
    for name, ms in _lib.profile_collect():
        per.setdefault(name, []).append(ms)
    _lib.profile(False)
    ref = (a[:64].double() @ b.double().T)
    print(json.dumps({"lib": os.environ.get("STX_B200_LIB", "in-tree"), "recipe": "C", "N": N, "step_ms": round(e0.elapsed_time(e1) / iters, 5),
                      "kernels_us": {k: round(1e3 * statistics.mean(v), 2) for k, v in per.items()},
                      "max_abs_err_64_rows": float((S[:64].double() - ref).abs().max())}))
    sys.exit(0)
n = int(secs * 16000)
pools = []
for s in range(4):                                             # 4 batches rotated: inputs + outputs larger than L2
    g = torch.Generator(device=dev).manual_seed(s)
    pools.append(0.1 * torch.randn(B * n, generator=g, device=dev))
off = torch.arange(B, device=dev, dtype=torch.int64) * n
ln = torch.full((B,), n, dtype=torch.int32, device=dev)
T_pad = 2 * ((ops.k_num_frames(n) + 1) // 2)
outs = [torch.empty((B, T_pad // 2, 160) if recipe == "K" else (B, 80, n // 160), dtype=torch.float32, device=dev)
        for _ in range(4)]


def step(i):
    if recipe == "K":
        ops.fbank_k(pools[i % 4], off, ln, n, T_pad, out=outs[i % 4], uniform=True)
    else:
        ops.logmel_w(pools[i % 4], off, ln, n, out=outs[i % 4])


for i in range(5):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    step(i)
e1.record()
torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / iters
_lib.profile(True)
for i in range(iters):
    step(i)
torch.cuda.synchronize()
per = {}
for name, ms in _lib.profile_collect():
    per.setdefault(name, []).append(ms)
_lib.profile(False)
print(json.dumps({"lib": os.environ.get("STX_B200_LIB", "in-tree"), "recipe": recipe, "clips": B, "seconds": secs,
                  "step_ms": round(step_ms, 5), "audio_s_per_s": round(B * secs / step_ms * 1e3),
                  "kernels_us": {k: round(1e3 * statistics.mean(v), 2) for k, v in per.items()},
                  "checksum": float(outs[0].double().abs().mean())}))
