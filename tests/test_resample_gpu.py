"""GPU parity tests for the device resampler through the C ABI.

Oracle: oracle/resample.py = scipy.signal.resample_poly (what librosa.resample(..., res_type="polyphase") runs),
evaluated in float64 from the same float32 filter.  Bar: max-abs <= 1e-6 * max(1, max|x|) (float32 accumulation of
<= 127 taps; scipy's own float32 loop is that far from the float64 value too), lengths exact.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import fbank_k as OK
from oracle import resample as OR
from speech_transcript_embeddings_b200 import ops, synth
from speech_transcript_embeddings_b200.feature_extraction import _layout
from speech_transcript_embeddings_b200.processor import AudioTextProcessor

pytestmark = pytest.mark.gpu
RATES = (48000, 44100, 32000, 22050, 8000, 24000, 96000, 64000, 11025)


def run(clips, sr, dev):
    lengths = np.array([c.size for c in clips], np.int32)
    offsets, total = _layout(lengths)
    host = np.zeros(max(total, 1), np.float32)
    for c, o in zip(clips, offsets):
        host[o:o + c.size] = c
    pcm = torch.from_numpy(host).to(dev)
    out, off_d, len_d, out_len, peak = ops.resample_poly(pcm, torch.from_numpy(offsets).to(dev),
                                                         torch.from_numpy(lengths).to(dev), lengths, sr, 16000)
    out_h, off_h = out.cpu().numpy(), off_d.cpu().numpy()
    assert np.array_equal(len_d.cpu().numpy(), out_len)
    return [out_h[o:o + n] for o, n in zip(off_h, out_len)], peak.cpu().numpy()


def test_golden_fixtures(cuda_device):
    g = load_golden("resample.npz")
    for i, ((sr, n, seed), kind) in enumerate(zip(g["spec"], g["kinds"])):
        x = synth.clip(str(kind), int(n), int(seed))
        (y,), peak = run([x], int(sr), cuda_device)
        ref = g[f"y_{i}"]
        assert y.shape == ref.shape and y.dtype == np.float32
        scale = max(1.0, float(np.abs(x).max()))
        err = float(np.abs(y - ref).max()) if ref.size else 0.0
        print(f"{sr} Hz, n = {n}: max-abs vs scipy {err:.2e}")
        assert err <= 1e-6 * scale
        assert peak[0] == max(np.float32(1.0), np.abs(y).max() if y.size else np.float32(0.0))


@pytest.mark.parametrize("sr", RATES)
def test_ragged_batch_matches_oracle(cuda_device, sr):
    rng = np.random.default_rng(sr)
    lens = [int(v) for v in rng.integers(1, 3 * sr, size=9)] + [0, 1, 2, sr // 100]
    kinds = ["G", "U", "AM", "HS", "loud", "small"]
    clips = [synth.clip(kinds[i % len(kinds)], n, 100 + i) for i, n in enumerate(lens)]
    ys, peaks = run(clips, sr, cuda_device)
    worst = 0.0
    for x, y, pk in zip(clips, ys, peaks):
        ref = OR.resample_poly(x, sr, 16000)
        assert y.shape == ref.shape
        if ref.size:
            worst = max(worst, float(np.abs(y - ref).max()) / max(1.0, float(np.abs(x).max())))
            assert pk == max(np.float32(1.0), np.abs(y).max())
        else:
            assert pk == 1.0
    print(f"{sr} Hz: worst scaled max-abs {worst:.2e}")
    assert worst <= 1e-6


@pytest.mark.parametrize("sr", (48000, 44100))
def test_full_size_properties(cuda_device, sr):
    """cfg2-sized batch (64 x 30 s at the source rate): oracle on sampled windows, linearity, unit DC gain."""
    n = 30 * sr
    g = torch.Generator(device=cuda_device).manual_seed(5)
    B = 64
    x = 0.1 * torch.randn(B * n, generator=g, device=cuda_device)
    off = torch.arange(B, device=cuda_device, dtype=torch.int64) * n
    lens_h = np.full(B, n, np.int32)
    lens = torch.from_numpy(lens_h).to(cuda_device)
    y, yo, yl, out_len, peak = ops.resample_poly(x, off, lens, lens_h, sr, 16000)
    assert (out_len == 480000).all()
    yo_h = yo.cpu().numpy()
    for b in (0, 17, 63):
        ref = OR.resample_poly(x[b * n:(b + 1) * n].cpu().numpy(), sr, 16000)
        got = y[yo_h[b]:yo_h[b] + 480000].cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-6
    # linearity: R(2 x + c) = 2 R(x) + R(c), and a constant comes out as the same constant away from the clip edges
    c = torch.full_like(x, 0.25)
    yc = ops.resample_poly(c, off, lens, lens_h, sr, 16000)[0]
    y2 = ops.resample_poly(2.0 * x + c, off, lens, lens_h, sr, 16000)[0]
    assert float((y2 - (2.0 * y + yc)).abs().max()) <= 2e-6
    interior = yc[yo_h[3] + 100:yo_h[3] + 480000 - 100]
    assert float((interior - 0.25).abs().max()) <= 1e-4 * 0.25       # the DC ripple of scipy's own polyphase branches (6e-5)
    assert torch.equal(peak, torch.ones_like(peak))                  # |y| stays below 1 for 0.1 N(0, 1)


def test_processor_resamples_on_the_device(cuda_device):
    """process_audio_array(x, 48000) == extractor(oracle-resampled x): R/processor.py:82-126 end to end."""
    proc = AudioTextProcessor(device=cuda_device, resample="device")
    for sr, kind in ((48000, "G"), (44100, "loud")):
        x = synth.clip(kind, int(2.5 * sr) + 7, 3)
        got = proc.process_audio_array(x, sr)
        xr = OR.resample_poly(x, sr, 16000)
        if np.abs(xr).max() > 1.0:
            xr = xr / np.abs(xr).max()
        ref, mask = OK.extract([xr[:480000]], padding_value=proc.feature_extractor.padding_value)
        assert got["input_features"].shape == ref.shape
        assert np.abs(got["input_features"].cpu().numpy() - ref).max() <= 1e-4
        assert np.array_equal(got["attention_mask_audio"].cpu().numpy(), mask)
    with pytest.raises(ValueError):
        AudioTextProcessor(device=cuda_device, resample="nope")
