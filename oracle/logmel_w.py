"""Recipe W oracle: Whisper log-mel (the recipe BASELINE.json's north_star lists).

NumPy restatement of ``WhisperFeatureExtractor`` (TF = transformers, third-party;
what R/processor.py:36 resolves to when ``audio_model_name`` is an
``openai/whisper-*`` checkpoint; the processor is generic over the extractor,
R/processor.py:107-113).

  mel table        TF/models/whisper/feature_extraction_whisper.py:95-103,
                   TF/audio_utils.py:285-296 (slaney hz->mel), 338-352 (mel->hz),
                   371-375 (triangles in Hz), 532-535 (slaney area norm)
  pad / truncate   TF/models/whisper/feature_extraction_whisper.py:296-303
  STFT             :141-150 (torch.stft: centre reflect-pad 200, periodic Hann-400,
                   hop 160, 201 bins, last frame dropped, |.|^2)
  mel/log/clamp    :152-161 (log10(max(.,1e-10)); per-clip max-8; (x+4)/4)
  mask             :328-337 (sample mask every 160th sample)

The float64 chain below is the extractor's NumPy path (:105-133); the default
torch path is the same arithmetic in float32 and agrees with it to ~3e-6 on the
parity-gated signal classes (tests/test_oracle_pinning.py checks both).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

NFFT = 400
HOP = 160
NBINS = NFFT // 2 + 1
NMEL = 80
NSAMPLES = 480000
NFRAMES = NSAMPLES // HOP


def hann_periodic() -> np.ndarray:
    i = np.arange(1 - (NFFT + 1), NFFT + 1, 2, dtype=np.float64)
    return (0.5 + 0.5 * np.cos(np.pi * i / NFFT))[:-1]


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mel = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    hi = f >= 1000.0
    return np.where(hi, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep, mel)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    hi = m >= 15.0
    return np.where(hi, 1000.0 * np.exp(logstep * (m - 15.0)), f)


def slaney_mel_filters() -> np.ndarray:
    """[201, 80] float64, Slaney scale, triangles in Hz, Slaney area normalisation."""
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(8000.0), NMEL + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    bins = np.linspace(0, 16000 // 2, NBINS)
    width = np.diff(hz_pts)
    dist = hz_pts[None, :] - bins[:, None]
    falling = -dist[:, :-2] / width[:-1]
    rising = dist[:, 2:] / width[1:]
    fb = np.maximum(0.0, np.minimum(falling, rising))
    fb *= (2.0 / (hz_pts[2:NMEL + 2] - hz_pts[:NMEL]))[None, :]
    return fb


def log_mel(pcm: np.ndarray, n_samples: int = NSAMPLES) -> np.ndarray:
    """float32 [80, n_samples/160] features of one clip (padded / truncated to n_samples)."""
    x = np.asarray(pcm, dtype=np.float32).reshape(-1)[:n_samples]
    x = np.pad(x, (0, n_samples - x.size)).astype(np.float64)
    x = np.pad(x, (NFFT // 2, NFFT // 2), mode="reflect")
    T = 1 + (x.size - NFFT) // HOP
    idx = HOP * np.arange(T)[:, None] + np.arange(NFFT)[None, :]
    fr = x[idx] * hann_periodic()[None, :]
    spec = np.fft.rfft(fr, axis=1).astype(np.complex64)
    power = np.abs(spec, dtype=np.float64) ** 2.0
    mel = np.maximum(1e-10, np.dot(slaney_mel_filters().T, power.T))
    ls = np.log10(mel).astype(np.float32)[:, :-1]
    ls = np.maximum(ls, ls.max() - 8.0)
    return ((ls + 4.0) / 4.0).astype(np.float32)


def extract(clips, n_samples: int = NSAMPLES, return_attention_mask: bool = False):
    """(input_features f32 [B, 80, n_samples/160], attention_mask i32 [B, n_samples/160] or None)."""
    feats = np.stack([log_mel(c, n_samples) for c in clips])
    mask = None
    if return_attention_mask:
        m = np.zeros((len(clips), n_samples), np.int32)
        for i, c in enumerate(clips):
            m[i, :min(np.asarray(c).reshape(-1).size, n_samples)] = 1
        mask = m[:, ::HOP]
        if n_samples % HOP != 0:
            mask = mask[:, :-1]
    return feats, mask
