"""Generates the golden fixtures in this directory from the THIRD-PARTY reference implementation.

The reference (yuriyvnv/speech_transcript_embeddings) delegates the arithmetic of the hot path to
``transformers`` (R/processor.py:36, 101-105); this script imports the installed transformers
(5.5.0 here; the reference pins 4.50.2) and torch, runs the very calls the reference makes on seeded
synthetic clips, and stores inputs + outputs.  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures pin ``oracle/`` (tests/test_oracle_pinning.py) and the CUDA path (tests/test_*_gpu.py).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from speech_transcript_embeddings_b200 import synth  # noqa: E402


def main():
    import transformers
    from transformers import SeamlessM4TFeatureExtractor, WhisperFeatureExtractor

    meta = f"transformers {transformers.__version__}, numpy {np.__version__}, torch {torch.__version__}"
    print(meta)

    # ---- recipe K: the call at R/processor.py:101-105, one clip per call --------------------------
    spec = [("G", 16000, 0), ("U", 16160, 1), ("AM", 11200, 2), ("HS", 20800, 3), ("small", 8000, 4),
            ("loud", 9600, 5), ("G", 400, 6), ("G", 719, 7)]
    clips = [synth.clip(k, n, s) for k, n, s in spec]
    fe = SeamlessM4TFeatureExtractor()
    out = {"meta": np.array(meta), "n_clips": np.array(len(clips))}
    for i, c in enumerate(clips):
        r = fe(c, sampling_rate=16000, return_tensors="np")
        out[f"pcm_{i}"] = c
        out[f"feat_{i}"] = r["input_features"]
        out[f"mask_{i}"] = r["attention_mask"]
    # batched call (pad to longest, pad_to_multiple_of=2), padding_value 0 and 1
    batch = clips[:6]
    for pv in (0.0, 1.0):
        r = SeamlessM4TFeatureExtractor(padding_value=pv)(batch, sampling_rate=16000, return_tensors="np")
        out[f"batch_feat_pv{int(pv)}"] = r["input_features"]
        out[f"batch_mask_pv{int(pv)}"] = r["attention_mask"]
    # raw (un-normalised) log-mel of clip 0
    r = fe(clips[0], sampling_rate=16000, return_tensors="np", do_normalize_per_mel_bins=False)
    out["raw_feat_0"] = r["input_features"]
    # the reference's start-up probe: 1000 zeros (R/processor.py:39-45)
    r = fe(np.zeros(1000, np.float32), sampling_rate=16000, return_tensors="np")
    out["probe_feat"] = r["input_features"]
    out["probe_mask"] = r["attention_mask"]
    np.savez_compressed(HERE / "fbank_k.npz", **out)

    # ---- recipe W: short max_length keeps the fixture small; plus one stock 30 s call ------------
    wfe = WhisperFeatureExtractor()
    wclips = [synth.clip("G", 16000, 10), synth.clip("AM", 12345, 11), synth.clip("HS", 8000, 12)]
    r = wfe(wclips, sampling_rate=16000, return_tensors="np", max_length=16000, return_attention_mask=True)
    wout = {"meta": np.array(meta), "n_clips": np.array(len(wclips)),
            "feat_ml16000": r["input_features"], "mask_ml16000": r["attention_mask"]}
    for i, c in enumerate(wclips):
        wout[f"pcm_{i}"] = c
    r = wfe(wclips[1], sampling_rate=16000, return_tensors="np")          # stock: pad to 30 s
    wout["feat_stock_1"] = r["input_features"][:, :, ::25].copy()          # every 25th frame (keeps it small)
    wout["feat_stock_1_tail"] = r["input_features"][:, :, -4:].copy()
    np.savez_compressed(HERE / "logmel_w.npz", **wout)

    # ---- cosine: the formulas at R/processor.py:148-159 in torch ---------------------------------
    a, b = synth.embedding_pairs(64, 768, seed=0)
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)

    def compute_similarity(e1, e2):   # R/processor.py:148-159 verbatim semantics, CPU
        if not torch.allclose(torch.norm(e1, p=2, dim=1), torch.ones(1), atol=1e-4):
            e1 = F.normalize(e1, p=2, dim=1)
        if not torch.allclose(torch.norm(e2, p=2, dim=1), torch.ones(1), atol=1e-4):
            e2 = F.normalize(e2, p=2, dim=1)
        return torch.sum(e1 * e2, dim=1).cpu().numpy()

    cout = {"meta": np.array(meta), "a": a, "b": b,
            "pair_unit": compute_similarity(ta, tb),
            "pair_scaled": compute_similarity(ta * 3.0, tb * 0.25),
            "pair_near_unit": compute_similarity(ta * 1.00005, tb),
            "matrix_f64": (F.normalize(ta.double(), dim=1) @ F.normalize(tb.double(), dim=1).T).numpy()}
    np.savez_compressed(HERE / "cosine.npz", **cout)
    for f in ("fbank_k.npz", "logmel_w.npz", "cosine.npz"):
        print(f, (HERE / f).stat().st_size, "bytes")


if __name__ == "__main__":
    main()
