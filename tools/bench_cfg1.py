"""cfg1: one 30 s clip, batch = 1, through the processor call the reference's inference scripts make
(R/inference.py:106 -> R/processor.py:79-126).  Wall-clock latency of process_audio_array (host array in, device
tensors out, synchronised) next to the third-party CPU extractor the reference calls (one thread).

    python tools/bench_cfg1.py
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200 import synth  # noqa: E402
from speech_transcript_embeddings_b200.processor import AudioTextProcessor  # noqa: E402

dev = torch.device("cuda", 0)
clip = synth.clip("G", 480000, 0)
res = {"workload": "cfg1: one 30 s clip, batch 1, process_audio_array"}
for name in ("facebook/w2v-bert-2.0", "openai/whisper-small"):
    proc = AudioTextProcessor(audio_model_name=name, device=dev)
    for _ in range(5):
        out = proc.process_audio_array(clip, 16000)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        out = proc.process_audio_array(clip, 16000)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    res[name] = {"b200_ms_median": 1e3 * float(np.median(ts)), "shape": list(out["input_features"].shape)}
try:
    import transformers
    torch.set_num_threads(1)
    for name, fe in (("facebook/w2v-bert-2.0", transformers.SeamlessM4TFeatureExtractor()),
                     ("openai/whisper-small", transformers.WhisperFeatureExtractor())):
        x = clip.astype(np.float32)
        for _ in range(2):
            fe(x, sampling_rate=16000, return_tensors="pt")
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            r = fe(x, sampling_rate=16000, return_tensors="pt")
            r["input_features"].to(dev)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        res[name]["reference_cpu_ms_median"] = 1e3 * float(np.median(ts))
        res[name]["speedup"] = res[name]["reference_cpu_ms_median"] / res[name]["b200_ms_median"]
except ImportError:
    pass
print(json.dumps(res))
