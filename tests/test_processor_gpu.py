"""GPU: the AudioTextProcessor drop-in (R/processor.py:79-159) against the reference's own steps
restated with the oracle."""
import numpy as np
import pytest
import torch

from oracle import cosine as OC
from oracle import fbank_k as OK
from oracle import logmel_w as OW
from speech_transcript_embeddings_b200 import synth
from speech_transcript_embeddings_b200.processor import AudioTextProcessor

pytestmark = pytest.mark.gpu


def _reference_prepare(audio, max_audio_length):
    """R/processor.py:88-97."""
    a = audio.astype(np.float32)
    if np.abs(a).max() > 1.0:
        a = a / np.abs(a).max()
    return a[:max_audio_length]


@pytest.mark.parametrize("kind", ["loud", "G", "AM"])
def test_process_audio_array_recipe_k(cuda_device, kind):
    proc = AudioTextProcessor(device=cuda_device, max_audio_length=40000)
    audio = synth.clip(kind, 52345, seed=8)
    out = proc.process_audio_array(audio, 16000)
    assert set(out) == {"input_features", "attention_mask_audio"}
    ref_x, ref_m = OK.extract([_reference_prepare(audio, 40000)], padding_value=proc.feature_extractor.padding_value)
    x, m = out["input_features"], out["attention_mask_audio"]
    assert x.device == cuda_device and x.dtype == torch.float32 and m.dtype == torch.int32
    assert tuple(x.shape) == ref_x.shape and np.array_equal(m.cpu().numpy(), ref_m)
    assert np.abs(x.cpu().numpy() - ref_x).max() <= 1e-4


def test_process_audio_array_recipe_w(cuda_device):
    proc = AudioTextProcessor(audio_model_name="openai/whisper-small", device=cuda_device)
    audio = synth.clip("loud", 30000, seed=9)
    out = proc.process_audio_array(audio, 16000)
    ref, _ = OW.extract([_reference_prepare(audio, 480000)])
    assert out["attention_mask_audio"] is None
    assert np.abs(out["input_features"].cpu().numpy() - ref).max() <= 1e-4


def test_compute_similarity(cuda_device):
    proc = AudioTextProcessor(device=cuda_device)
    a, b = synth.embedding_pairs(8, 1024, seed=5)
    s = proc.compute_similarity(torch.from_numpy(a * 2).to(cuda_device), torch.from_numpy(b).to(cuda_device))
    assert isinstance(s, np.ndarray) and s.dtype == np.float32 and s.shape == (8,)
    assert np.abs(s - OC.pairwise_reference(a * 2, b)).max() <= 1e-5


def test_unsupported_audio_model_is_an_error(cuda_device):
    with pytest.raises(ValueError):
        AudioTextProcessor(audio_model_name="facebook/wav2vec2-base", device=cuda_device)


def test_process_audio_file_wav_48k(cuda_device, tmp_path):
    """R/processor.py:69-77 end to end on a 48 kHz 16-bit WAV: decode, device resampling, peak rule, recipe K."""
    import wave
    from oracle import resample as OR
    from speech_transcript_embeddings_b200.processor import load_audio
    x = (synth.clip("G", 60000, 9) * 3.0).clip(-1.0, 32767.0 / 32768.0)
    pcm = np.round(x * 32768.0).astype(np.int16)
    path = tmp_path / "clip48k.wav"
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(48000)
        w.writeframes(pcm.tobytes())
    audio, sr = load_audio(path)
    assert sr == 48000 and audio.dtype == np.float32
    proc = AudioTextProcessor(device=cuda_device, resample="device")
    got = proc.process_audio_file(str(path))
    xr = OR.resample_poly(audio, 48000, 16000)
    if np.abs(xr).max() > 1.0:
        xr = xr / np.abs(xr).max()
    ref, mask = OK.extract([xr], padding_value=proc.feature_extractor.padding_value)
    assert np.abs(got["input_features"].cpu().numpy() - ref).max() <= 1e-4
    assert np.array_equal(got["attention_mask_audio"].cpu().numpy(), mask)


def test_default_resampling_is_the_reference_host_call(cuda_device):
    """R/processor.py:82-86: non-16 kHz audio goes through librosa.resample (soxr_hq) unless the device path is asked for."""
    proc = AudioTextProcessor(device=cuda_device)
    assert proc.resample == "librosa"
    x = synth.clip("G", 48000, 4)
    try:
        import librosa
    except ImportError:
        with pytest.raises(ImportError, match="resample='device'"):
            proc.process_audio_array(x, 48000)
        return
    got = proc.process_audio_array(x, 48000)
    xr = librosa.resample(x, orig_sr=48000, target_sr=16000)
    ref, mask = OK.extract([_reference_prepare(xr, 480000)], padding_value=proc.feature_extractor.padding_value)
    assert np.abs(got["input_features"].cpu().numpy() - ref).max() <= 1e-4
    assert np.array_equal(got["attention_mask_audio"].cpu().numpy(), mask)


def test_padding_value_follows_the_hub_config_of_the_model_family(cuda_device):
    """The half-frame behind an odd-length clip carries the extractor's padding_value (mask 0 there)."""
    proc = AudioTextProcessor(device=cuda_device)                     # w2v-bert-2.0: 1.0 unless a cached hub config says otherwise
    pv = proc.feature_extractor.padding_value
    audio = synth.clip("G", 400 + 160 * 6, seed=2)                    # 7 frames: odd
    out = proc.process_audio_array(audio, 16000)
    x = out["input_features"].cpu().numpy()
    assert x.shape == (1, 4, 160) and np.all(x[0, 3, 80:] == np.float32(pv))
    assert out["attention_mask_audio"].cpu().numpy().tolist() == [[1, 1, 1, 0]]
    assert AudioTextProcessor(device=cuda_device, padding_value=0.25).feature_extractor.padding_value == 0.25
