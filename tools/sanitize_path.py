"""Small invocations of every kernel path for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_path.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200 import ops, synth  # noqa: E402
from speech_transcript_embeddings_b200.feature_extraction import (B200SeamlessM4TFeatureExtractor,  # noqa: E402
                                                                  B200WhisperFeatureExtractor)
from speech_transcript_embeddings_b200.processor import AudioTextProcessor  # noqa: E402

dev = torch.device("cuda", 0)
clips = [synth.clip("G", 16000 * 3 + 77, 1), synth.clip("AM", 719, 2), synth.clip("loud", 40000, 3), synth.clip("U", 300, 4)]
k = B200SeamlessM4TFeatureExtractor(device=dev)
r = k(clips, sampling_rate=16000, return_tensors="pt")
c = k.collate(clips[:3])
h = k(clips, sampling_rate=16000, return_tensors="np")
w = B200WhisperFeatureExtractor(device=dev)
rw = w(clips, sampling_rate=16000, return_tensors="pt", max_length=48000, return_attention_mask=True)
p = AudioTextProcessor(device=dev, max_audio_length=30000)
pa = p.process_audio_array(clips[2], 16000)
a, b = synth.embedding_pairs(200, 100, seed=0)
ad, bd = torch.from_numpy(a).to(dev), torch.from_numpy(b[:150]).to(dev)
S = ops.cosine_nxm(ad, bd)
s = ops.cosine_pairwise(ad[:150], bd, always_normalize=False)
pn = ops.score_pos_neg(ad[:150], bd, ad[50:200].contiguous())
hid, nrm = ops.feature_projection(r["input_features"], torch.ones(160, device=dev), torch.zeros(160, device=dev),
                                  0.05 * torch.randn(96, 160, device=dev), torch.zeros(96, device=dev))
torch.cuda.synchronize()
print("ok", tuple(r["input_features"].shape), tuple(rw["input_features"].shape), tuple(S.shape), float(pn["loss"]), tuple(hid.shape))
