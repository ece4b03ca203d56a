"""Device resampler: 64 clips of 30 s at the source rate -> 16 kHz, device-resident.

    python tools/bench_resample.py [rates...]

Prints per source rate: ms per batch (CUDA events, 4 rotated input batches so that the stream comes from HBM),
audio-s/s, and the algorithmic HBM rate (4 B per input sample + 4 B per output sample) against the measured copy peak.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from speech_transcript_embeddings_b200 import ops  # noqa: E402

rates = [int(a) for a in sys.argv[1:]] or [48000, 44100, 32000, 22050, 8000]
dev = torch.device("cuda", 0)
B, secs = 64, 30.0
peak_gbs = 6448.0
try:
    peak_gbs = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
for sr in rates:
    n = int(secs * sr)
    g = torch.Generator(device=dev).manual_seed(sr)
    pools = [0.1 * torch.randn(B * n, generator=g, device=dev) for _ in range(4)]
    off = torch.arange(B, device=dev, dtype=torch.int64) * n
    lens_h = np.full(B, n, np.int32)
    lens = torch.from_numpy(lens_h).to(dev)
    for i in range(4):
        out = ops.resample_poly(pools[i], off, lens, lens_h, sr, 16000)
    torch.cuda.synchronize()
    iters = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        out = ops.resample_poly(pools[i % 4], off, lens, lens_h, sr, 16000)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    n_out = int(out[3][0])
    gbs = B * 4.0 * (n + n_out) / (ms * 1e-3) / 1e9
    p = ops.resample_plan(sr, 16000)
    print(json.dumps({"orig_sr": sr, "up": p["up"], "down": p["down"], "taps_per_output": p["taps_per_phase"],
                      "ms": round(ms, 4), "audio_s_per_s": round(B * secs / ms * 1e3), "algorithmic_GBps": round(gbs, 1),
                      "hbm_frac": round(gbs / peak_gbs, 3)}))
