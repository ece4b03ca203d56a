"""Precision study of recipe K (profiles/r02_k_precision.md): error against the CPU oracle over ALL 64 clips of cfg2 and the six
gated signal classes, and the kernel time, for one build of the library.

    STX_K_SINGLE=1 STX_B200_LIB=build/variants/lib_f32.so python tests/scripts/precision_study.py <label>

The float32 / mixed builds exist only for this study (tools/ab_build.py f32=-DSTX_K_P1_F32=1,-DSTX_K_P2_F32=1
mix=-DSTX_K_P2_F32=1; they change k_frames, the single-group kernel STX_K_SINGLE=1 selects).  The oracle is the checker.
"""
import json
import os
import statistics
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from oracle import fbank_k as OK  # noqa: E402  (checker)
from speech_transcript_embeddings_b200 import _lib, ops, synth  # noqa: E402
from speech_transcript_embeddings_b200.feature_extraction import B200SeamlessM4TFeatureExtractor  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else "default"
dev = torch.device("cuda", 0)
fe = B200SeamlessM4TFeatureExtractor(device=dev)


def errors(clips):
    """(max-abs error of the normalised features, of the raw log-mel) per clip, against the oracle."""
    got = fe(clips, sampling_rate=16000, return_tensors="np")["input_features"]
    raw = fe(clips, sampling_rate=16000, return_tensors="np", do_normalize_per_mel_bins=False)["input_features"]

    def one(i):
        with np.errstate(all="ignore"):
            ref = OK.extract([clips[i]])[0][0]
            ref_raw = OK.extract([clips[i]], normalize=False)[0][0]
        return (float(np.abs(got[i, :ref.shape[0]] - ref).max()), float(np.abs(raw[i, :ref.shape[0]] - ref_raw).max()))

    with ThreadPoolExecutor(8) as ex:
        return np.array(list(ex.map(one, range(len(clips)))))


res = {"label": label, "lib": os.environ.get("STX_B200_LIB", "in-tree"), "single_group_kernel": os.environ.get("STX_K_SINGLE") == "1"}
e = errors(synth.batch_fixed(64, 30.0, "G", 0))
res["cfg2_all_64_clips"] = {"max_abs_norm": float(e[:, 0].max()), "median_clip_norm": float(np.median(e[:, 0])),
                            "max_abs_raw_logmel": float(e[:, 1].max()), "clips_over_1e-4": int((e[:, 0] > 1e-4).sum())}
classes = {}
for kind in synth.GATED_CLASSES:
    e = errors([synth.clip(kind, 160000, 100 + s) for s in range(4)])
    classes[kind] = {"max_abs_norm": float(e[:, 0].max()), "max_abs_raw_logmel": float(e[:, 1].max())}
res["gated_classes_4x10s"] = classes
res["worst_norm"] = max([res["cfg2_all_64_clips"]["max_abs_norm"]] + [v["max_abs_norm"] for v in classes.values()])
res["holds_1e-4"] = bool(res["worst_norm"] <= 1e-4)

# kernel time on the device-resident cfg2 batch (4 batches rotated, CUDA events around every launch inside the library)
B, n = 64, 480000
pools = [0.1 * torch.randn(B * n, generator=torch.Generator(device=dev).manual_seed(s), device=dev) for s in range(4)]
off = torch.arange(B, device=dev, dtype=torch.int64) * n
ln = torch.full((B,), n, dtype=torch.int32, device=dev)
T_pad = 2 * ((ops.k_num_frames(n) + 1) // 2)
outs = [torch.empty((B, T_pad // 2, 160), dtype=torch.float32, device=dev) for _ in range(4)]
for i in range(5):
    ops.fbank_k(pools[i % 4], off, ln, n, T_pad, out=outs[i % 4], uniform=True)
torch.cuda.synchronize()
_lib.profile(True)
for i in range(20):
    ops.fbank_k(pools[i % 4], off, ln, n, T_pad, out=outs[i % 4], uniform=True)
torch.cuda.synchronize()
per = {}
for name, ms in _lib.profile_collect():
    per.setdefault(name, []).append(ms)
_lib.profile(False)
res["kernels_us"] = {k: round(1e3 * statistics.mean(v), 2) for k, v in per.items()}
print(json.dumps(res), flush=True)
