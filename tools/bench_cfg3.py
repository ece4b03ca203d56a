"""cfg3: 512 variable-length clips (1-30 s), device-resident, recipe K and W: ms per batch and audio-s/s.

    python tools/bench_cfg3.py
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200 import ops, synth  # noqa: E402
from speech_transcript_embeddings_b200.feature_extraction import _layout  # noqa: E402

dev = torch.device("cuda", 0)
lens = synth.variable_lengths(512, 1234, True).astype(np.int32)
offsets, total = _layout(lens)
g = torch.Generator(device=dev).manual_seed(0)
pcm = 0.1 * torch.randn(total, generator=g, device=dev)
off_d = torch.from_numpy(offsets).to(dev)
len_d = torch.from_numpy(lens).to(dev)
audio_s = float(lens.sum()) / 16000.0
frames = np.array([ops.k_num_frames(int(n)) for n in lens])
T_pad = int(frames.max() + (frames.max() & 1))
out_k = torch.empty((512, T_pad // 2, 160), dtype=torch.float32, device=dev)
out_w = torch.empty((512, 80, 3000), dtype=torch.float32, device=dev)


def timed(f, iters=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms_k = timed(lambda: ops.fbank_k(pcm, off_d, len_d, int(lens.max()), T_pad, out=out_k))
ms_w = timed(lambda: ops.logmel_w(pcm, off_d, len_d, 480000, out=out_w))
print(json.dumps({"workload": "cfg3: 512 clips, 1-30 s", "audio_s": audio_s,
                  "K": {"ms": ms_k, "audio_s_per_s": audio_s / ms_k * 1e3},
                  "W": {"ms": ms_w, "audio_s_per_s_real_audio": audio_s / ms_w * 1e3,
                        "audio_s_per_s_padded_30s": 512 * 30 / ms_w * 1e3}}))
