"""GPU parity tests for recipe W (Whisper log-mel) through the C ABI.  Bar: max-abs <= 1e-4."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import logmel_w as OW
from speech_transcript_embeddings_b200 import synth
from speech_transcript_embeddings_b200.feature_extraction import B200WhisperFeatureExtractor

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def fe(cuda_device):
    return B200WhisperFeatureExtractor(device=cuda_device)


def test_golden_fixtures(fe):
    g = load_golden("logmel_w.npz")
    clips = [g[f"pcm_{i}"] for i in range(int(g["n_clips"]))]
    got = fe(clips, sampling_rate=16000, return_tensors="np", max_length=16000, return_attention_mask=True)
    assert got["input_features"].shape == (3, 80, 100) and got["input_features"].dtype == np.float32
    assert np.abs(got["input_features"] - g["feat_ml16000"]).max() <= TOL
    assert np.array_equal(got["attention_mask"], g["mask_ml16000"]) and got["attention_mask"].dtype == np.int32
    got = fe(clips[1], sampling_rate=16000, return_tensors="np")
    assert got["input_features"].shape == (1, 80, 3000) and "attention_mask" not in got
    assert np.abs(got["input_features"][:, :, ::25] - g["feat_stock_1"]).max() <= TOL
    assert np.abs(got["input_features"][:, :, -4:] - g["feat_stock_1_tail"]).max() <= TOL


@pytest.mark.parametrize("kind", synth.GATED_CLASSES + synth.REPORTED_CLASSES)
def test_signal_classes(fe, kind):
    c = synth.clip(kind, 100000 + 13, seed=21)
    ref, _ = OW.extract([c])
    got = fe(c, sampling_rate=16000, return_tensors="np")["input_features"]
    err = np.abs(got - ref).max()
    print(f"W {kind}: max-abs {err:.2e}")
    assert err <= TOL


def test_truncation_and_batch(fe):
    clips = [synth.clip("G", 500000, 1), synth.clip("U", 480000, 2), synth.clip("AM", 7, 3), synth.clip("G", 479999, 4)]
    ref, rm = OW.extract(clips, return_attention_mask=True)
    got = fe(clips, sampling_rate=16000, return_tensors="np", return_attention_mask=True)
    assert np.abs(got["input_features"] - ref).max() <= TOL
    assert np.array_equal(got["attention_mask"], rm)


def test_full_size_properties(fe):
    clips = synth.batch_fixed(64, 30.0, "G", 100)
    x = fe(clips, sampling_rate=16000, return_tensors="pt")["input_features"]
    assert tuple(x.shape) == (64, 80, 3000) and x.is_cuda
    mx = x.amax(dim=(1, 2))
    mn = x.amin(dim=(1, 2))
    assert bool((mn >= mx - 2.0 - 1e-6).all())                       # max(x, max - 8) then /4
    alone = fe(clips[5], sampling_rate=16000, return_tensors="pt")["input_features"]
    assert torch.equal(alone[0], x[5])
    ref, _ = OW.extract([clips[5]])
    assert np.abs(x[5].cpu().numpy() - ref[0]).max() <= TOL


def test_host_output_pipeline_is_bit_identical(cuda_device):
    fe = B200WhisperFeatureExtractor(device=cuda_device)
    clips = [synth.clip("G", 16000, 1), synth.clip("AM", 9000, 2), synth.clip("U", 20000, 3)]
    dev_out = fe(clips, sampling_rate=16000, return_tensors="pt", max_length=16000, return_attention_mask=True)
    old = fe.CHUNK_BYTES
    try:
        type(fe).CHUNK_BYTES = 1
        host = fe(clips, sampling_rate=16000, return_tensors="pt", max_length=16000, return_attention_mask=True,
                  output="host")
        assert host["input_features"].is_pinned()
        assert torch.equal(host["input_features"], dev_out["input_features"].cpu())
        assert torch.equal(host["attention_mask"], dev_out["attention_mask"].cpu())
    finally:
        type(fe).CHUNK_BYTES = old


def test_unaligned_clip_starts_take_the_plain_load_path(cuda_device):
    from speech_transcript_embeddings_b200 import ops
    clips = [synth.clip("G", 16000, 1), synth.clip("AM", 9001, 2), synth.clip("U", 12345, 3)]
    lens = np.array([c.size for c in clips], np.int32)

    def run(lead, gap):
        offs, pos = [], lead
        for c in clips:
            offs.append(pos)
            pos += c.size + gap
        buf = np.zeros(pos + 8, np.float32)
        for c, o in zip(clips, offs):
            buf[o:o + c.size] = c
        x, _ = ops.logmel_w(torch.from_numpy(buf).to(cuda_device), torch.tensor(offs, dtype=torch.int64, device=cuda_device),
                            torch.from_numpy(lens).to(cuda_device), 16000)
        return x.cpu()

    x0 = run(0, 3)
    for lead, gap in ((1, 0), (2, 1), (3, 7)):
        assert torch.equal(run(lead, gap), x0)
    ref, _ = OW.extract(clips, n_samples=16000)
    assert np.abs(x0.numpy() - ref).max() <= TOL
