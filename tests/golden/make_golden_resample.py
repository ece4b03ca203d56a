"""Golden vectors for the resampler from the THIRD-PARTY implementation librosa delegates to.

``librosa.resample(y, orig_sr=sr, target_sr=16000, res_type="polyphase")`` is ``scipy.signal.resample_poly(y, 16000 // g,
sr // g)`` + ``fix_length(ceil(n * 16000 / sr))`` (librosa 0.10.1, core/audio.py).  librosa and soxr (the reference's DEFAULT
res_type "soxr_hq", R/processor.py:85) are not installed offline; scipy is, so the fixture pins the polyphase path only.

    python tests/golden/make_golden_resample.py
"""
from __future__ import annotations

import math
import sys
from pathlib import Path

import numpy as np
import scipy
import scipy.signal

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from speech_transcript_embeddings_b200 import synth  # noqa: E402

# (orig_sr, samples, signal class, seed): every branch of the device path (integer decimation by 2, 3, 4, 6; general
# polyphase with the bank in shared memory; up-sampling), odd lengths, a clip shorter than the filter
SPEC = [(48000, 12000, "G", 0), (48000, 7001, "loud", 1), (44100, 9000, "U", 2), (32000, 8191, "AM", 3),
        (22050, 6000, "G", 4), (8000, 3000, "HS", 5), (24000, 5000, "G", 6), (96000, 13000, "U", 7),
        (64000, 9000, "G", 8), (11025, 4000, "small", 9), (48000, 17, "G", 10), (44100, 1, "G", 11)]


def main():
    out = {"meta": np.array(f"scipy {scipy.__version__}, numpy {np.__version__}"),
           "spec": np.array([[sr, n, seed] for sr, n, _, seed in SPEC], np.int64),
           "kinds": np.array([k for _, _, k, _ in SPEC])}
    for i, (sr, n, kind, seed) in enumerate(SPEC):
        x = synth.clip(kind, n, seed)
        g = math.gcd(sr, 16000)
        y = scipy.signal.resample_poly(x, 16000 // g, sr // g)
        n_out = int(np.ceil(n * 16000 / sr))
        assert y.shape == (n_out,) and y.dtype == np.float32, (y.shape, y.dtype, n_out)   # fix_length is a no-op
        out[f"y_{i}"] = y
    np.savez_compressed(HERE / "resample.npz", **out)
    print("resample.npz", (HERE / "resample.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
