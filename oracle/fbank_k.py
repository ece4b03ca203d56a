"""Recipe K oracle: Kaldi-style fbank + per-bin CMVN + stride-2 stacking.

NumPy float64 restatement of what the reference's processor runs for
``audio_model_name="facebook/w2v-bert-2.0"`` (R/processor.py:36, 101-105;
R/training/trainer_unfreeze.py:856-860): ``SeamlessM4TFeatureExtractor.__call__``.
TF = transformers (third-party; reference pins 4.50.2, 5.5.0 installed here).

  tables            TF/models/seamless_m4t/feature_extraction_seamless_m4t.py:73-85
                    TF/audio_utils.py:282-283 (kaldi mel), 371-375 (triangles),
                    516-530 (mel-space triangularisation), 593-602 (povey)
  per-frame chain   TF/audio_utils.py:774-803 (f64 frame, -mean, pre-emphasis,
                    window, rfft-512, rounded to complex64)
  power/mel/log     TF/audio_utils.py:808-830
  CMVN              TF/models/seamless_m4t/feature_extraction_seamless_m4t.py:257-262
  pad / mask        TF/feature_extraction_sequence_utils.py:196-199, 255-291
  stride-2 stack    TF/models/seamless_m4t/feature_extraction_seamless_m4t.py:281-300

The per-frame Python loop of the original is vectorised over frames here (one
batched float64 rfft); each frame sees exactly the same float64 operations.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

FRAME = 400
HOP = 160
NFFT = 512
NBINS = NFFT // 2 + 1
NMEL = 80
STRIDE = 2
PREEMPH = 0.97
MEL_FLOOR = 1.192092955078125e-07
SCALE = 2.0 ** 15


def povey_window() -> np.ndarray:
    """Symmetric Hann(400) ** 0.85 in float64 (TF/audio_utils.py:593, 601-602)."""
    # numpy's own hanning formulation (odd-integer grid), so the table is bit-identical
    i = np.arange(1 - FRAME, FRAME, 2, dtype=np.float64)
    hann = 0.5 + 0.5 * np.cos(np.pi * i / (FRAME - 1))
    return np.power(hann, 0.85)


def kaldi_mel_filters() -> np.ndarray:
    """[257, 80] float64 triangular filters built in mel space, 20 Hz..8 kHz,
    no area normalisation (TF/audio_utils.py:516-530, 371-375)."""
    def mel(f):
        return 1127.0 * np.log(1.0 + f / 700.0)

    centres = np.linspace(mel(20.0), mel(8000.0), NMEL + 2)
    bin_mel = mel((16000.0 / NFFT) * np.arange(NBINS))
    width = np.diff(centres)
    dist = centres[None, :] - bin_mel[:, None]            # [257, 82]
    falling = -dist[:, :-2] / width[:-1]
    rising = dist[:, 2:] / width[1:]
    return np.maximum(0.0, np.minimum(falling, rising))


def num_frames(n: int) -> int:
    """TF/audio_utils.py:778 (center=False)."""
    return int(1 + np.floor((n - FRAME) / HOP))


def raw_log_mel(pcm: np.ndarray) -> np.ndarray:
    """float32 [T, 80] natural-log mel energies of one clip
    (SeamlessM4TFeatureExtractor._extract_fbank_features)."""
    x = np.asarray(pcm)
    if x.ndim == 2:                       # stereo -> channel 0 (…seamless_m4t.py:121-122)
        x = x[0]
    x = (np.squeeze(x) * SCALE).astype(np.float64)
    T = num_frames(x.size)
    if T <= 0:
        return np.empty((0, NMEL), np.float32)
    idx = HOP * np.arange(T)[:, None] + np.arange(FRAME)[None, :]
    fr = x[idx]                                            # [T, 400] f64
    fr = fr - fr.mean(axis=1, keepdims=True)
    em = np.empty_like(fr)
    em[:, 1:] = fr[:, 1:] - PREEMPH * fr[:, :-1]
    em[:, 0] = fr[:, 0] * (1.0 - PREEMPH)
    em *= povey_window()[None, :]
    buf = np.zeros((T, NFFT))
    buf[:, :FRAME] = em
    spec = np.fft.rfft(buf, axis=1).astype(np.complex64)   # rounding point, :781
    power = np.abs(spec, dtype=np.float64) ** 2.0
    # keep the original's memory layout ([80, T] C-order, returned transposed): the
    # float32 CMVN that follows reduces along the contiguous axis (pairwise sums)
    mel = np.maximum(MEL_FLOOR, np.dot(kaldi_mel_filters().T, power.T))
    return np.log(mel).astype(np.float32).T


def cmvn(feat: np.ndarray) -> np.ndarray:
    """Per-clip per-bin normalisation in float32 (…seamless_m4t.py:257-262)."""
    return (feat - np.expand_dims(feat.mean(0), 0)) / np.sqrt(np.expand_dims(feat.var(0, ddof=1), 0) + 1e-7)


def extract(clips, padding_value: float = 0.0, pad_to_multiple_of: int | None = 2,
            normalize: bool = True, max_length: int | None = None, truncation: bool = False,
            padding="longest"):
    """Batched call: returns (input_features f32 [B, T'/2, 160], attention_mask i32 [B, T'/2])."""
    feats = []
    for c in clips:
        f = raw_log_mel(np.asarray(c, dtype=np.float32))
        if normalize:
            with np.errstate(invalid="ignore", divide="ignore"):
                f = cmvn(f)
        feats.append(f.astype(np.float32))
    if truncation and max_length is not None:
        lim = max_length
        if pad_to_multiple_of is not None and lim % pad_to_multiple_of != 0:
            lim = ((lim // pad_to_multiple_of) + 1) * pad_to_multiple_of
        feats = [f[:lim] for f in feats]
    if padding in (True, "longest"):
        tgt = max(f.shape[0] for f in feats)
    elif padding == "max_length":
        tgt = max_length
    else:
        tgt = None
    if tgt is not None and pad_to_multiple_of is not None and tgt % pad_to_multiple_of != 0:
        tgt = ((tgt // pad_to_multiple_of) + 1) * pad_to_multiple_of
    rows, masks = [], []
    for f in feats:
        t = f.shape[0]
        m = np.ones(t, np.int32)
        if tgt is not None and t < tgt:
            f = np.pad(f, ((0, tgt - t), (0, 0)), "constant", constant_values=padding_value)
            m = np.pad(m, (0, tgt - t))
        rows.append(f)
        masks.append(m)
    x = np.stack(rows).astype(np.float32)
    m = np.stack(masks)
    B, T, C = x.shape
    rem = T % STRIDE
    if rem:
        x = x[:, :T - rem]
        m = m[:, :T - rem]
    x = x.reshape(B, T // STRIDE, C * STRIDE)
    m = m[:, np.arange(T - rem) % STRIDE == 1]
    return x, m
