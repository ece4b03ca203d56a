"""CPU: host-side logic of the drop-in (padding rules, packing, sharding, error behaviour)."""
import numpy as np
import pytest

from oracle import fbank_k as OK
from speech_transcript_embeddings_b200 import ops, sharding, synth
from speech_transcript_embeddings_b200.feature_extraction import (B200SeamlessM4TFeatureExtractor,
                                                                  B200WhisperFeatureExtractor, BatchFeature,
                                                                  _as_clip_list, _layout)


@pytest.mark.parametrize("lengths", [[16000], [16160, 400, 48000], [719, 720, 721], [399, 16000]])
def test_padded_frames_match_oracle_shapes(lengths):
    clips = [np.zeros(n, np.float32) + 0.01 * np.arange(n, dtype=np.float32) % 1 for n in lengths]
    frames = np.array([ops.k_num_frames(n) for n in lengths])
    assert [max(OK.num_frames(n), 0) for n in lengths] == list(frames)
    T_pad, _ = B200SeamlessM4TFeatureExtractor._padded_frames(frames, True, None, False, 2)
    if min(lengths) >= 400:
        with np.errstate(all="ignore"):
            x, m = OK.extract(clips)
        assert x.shape[1] * 2 == T_pad and m.shape[1] * 2 == T_pad


def test_padding_strategies():
    f = np.array([99, 10, 300])
    pf = B200SeamlessM4TFeatureExtractor._padded_frames
    assert pf(f, True, None, False, 2)[0] == 300
    assert pf(np.array([99]), True, None, False, 2)[0] == 100
    assert pf(np.array([99]), True, None, False, None)[0] == 98       # remainder frame dropped (…seamless_m4t.py:281-285)
    assert pf(f, "max_length", 401, False, 2)[0] == 402
    T_pad, kept = pf(f, "longest", 100, True, 2)
    assert T_pad == 100 and list(kept) == [99, 10, 100]
    with pytest.raises(ValueError):
        pf(f, False, None, False, 2)
    with pytest.raises(ValueError):
        pf(f, "max_length", 100, False, 2)


def test_layout_is_128_byte_aligned_and_disjoint():
    lengths = np.array([1, 31, 32, 33, 480000, 5], np.int32)
    off, total = _layout(lengths)
    assert (off % 32 == 0).all() and off[0] == 0
    assert (off[1:] >= off[:-1] + lengths[:-1]).all() and total >= off[-1] + lengths[-1]


def test_batching_rules():
    one = _as_clip_list(np.zeros(1000, np.float64), 3, "X")
    assert len(one) == 1 and one[0].dtype == np.float32
    assert len(_as_clip_list([0.0] * 500, 3, "X")) == 1
    assert len(_as_clip_list([np.zeros(500), np.zeros(700)], 3, "X")) == 2
    assert len(_as_clip_list(np.zeros((3, 500)), 3, "X")) == 3
    stereo = _as_clip_list([np.stack([np.ones(500), np.zeros(500)])], 3, "X")
    assert stereo[0].shape == (500,) and stereo[0].all()
    with pytest.raises(ValueError):
        _as_clip_list(np.zeros((2, 2, 2, 2)), 3, "X")


def test_sampling_rate_errors_and_ctor_guards():
    fe = B200SeamlessM4TFeatureExtractor(device=None)
    with pytest.raises(ValueError, match="16000"):
        fe(np.zeros(1000, np.float32), sampling_rate=8000)
    with pytest.raises(ValueError):
        B200SeamlessM4TFeatureExtractor(num_mel_bins=40)
    with pytest.raises(ValueError):
        B200WhisperFeatureExtractor(n_fft=512)


def test_batch_feature_protocol():
    bf = BatchFeature({"input_features": 1, "attention_mask": 2})
    assert "input_features" in bf and bf["attention_mask"] == 2 and bf.get("nope") is None
    assert list(bf.keys()) == ["input_features", "attention_mask"] and bf.input_features == 1


def test_shard_clips_partitions_and_balances():
    lens = synth.variable_lengths(512, seed=1234)
    shards = sharding.shard_clips(lens, 8)
    allidx = np.sort(np.concatenate(shards))
    assert np.array_equal(allidx, np.arange(512))
    load = np.array([lens[s].sum() for s in shards])
    assert load.max() / load.mean() < 1.01
    parts = [lens[s][:, None] for s in shards]
    assert np.array_equal(sharding.unshard(parts, shards)[:, 0], lens)


def test_chunk_bounds_cover_every_clip_once():
    lengths = np.array([480000] * 10 + [16000, 400, 32000], np.int32)
    offsets, _ = _layout(lengths)
    fe = B200SeamlessM4TFeatureExtractor
    for chunk_bytes in (1, 4 * 480000, 12 << 20, 1 << 40):
        bounds = fe.chunk_bounds(offsets, lengths, chunk_bytes)
        assert bounds[0][0] == 0 and bounds[-1][1] == lengths.size
        assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:])) and all(b1 > b0 for b0, b1 in bounds)
        for b0, b1 in bounds:
            if b1 - b0 > 1:
                assert (offsets[b1 - 1] + lengths[b1 - 1] - offsets[b0]) * 4 <= chunk_bytes
    assert len(fe.chunk_bounds(offsets, lengths, 1)) == lengths.size
    assert len(fe.chunk_bounds(offsets, lengths, 1 << 40)) == 1


def test_load_audio_reads_pcm_wav_without_librosa(tmp_path):
    """process_audio_file's decode step (R/processor.py:74) for uncompressed WAV when librosa is absent: libsndfile scaling,
    channel average, native sampling rate."""
    import wave
    from speech_transcript_embeddings_b200.processor import load_audio
    try:
        import librosa  # noqa: F401
        pytest.skip("librosa installed: the reference's own decoder is used")
    except ImportError:
        pass
    rng = np.random.default_rng(0)
    pcm = rng.integers(-32768, 32767, size=(4800, 2), dtype=np.int16)
    path = tmp_path / "a.wav"
    with wave.open(str(path), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000)
        w.writeframes(pcm.tobytes())
    x, sr = load_audio(path)
    assert sr == 48000 and x.dtype == np.float32 and x.shape == (4800,)
    assert np.array_equal(x, (pcm.astype(np.float32) / 32768.0).mean(axis=1, dtype=np.float32))
    bad = tmp_path / "b.mp3"
    bad.write_bytes(b"ID3\x03" + bytes(64))
    with pytest.raises(ImportError):
        load_audio(bad)


@pytest.mark.parametrize("width", [1, 2, 3, 4])
def test_load_audio_pcm_widths(tmp_path, width):
    """8/16/24/32-bit PCM WAV with libsndfile's scaling (full scale = 1.0); mono stays mono."""
    import wave
    from speech_transcript_embeddings_b200.processor import load_audio
    try:
        import librosa  # noqa: F401
        pytest.skip("librosa installed: the reference's own decoder is used")
    except ImportError:
        pass
    rng = np.random.default_rng(width)
    n = 1000
    if width == 1:
        v = rng.integers(0, 256, size=n)
        raw, want = v.astype(np.uint8).tobytes(), (v.astype(np.float32) - 128.0) / 128.0
    elif width == 2:
        v = rng.integers(-32768, 32768, size=n)
        raw, want = v.astype("<i2").tobytes(), v.astype(np.float32) / 32768.0
    elif width == 3:
        v = rng.integers(-(1 << 23), 1 << 23, size=n)
        raw = b"".join(int(x & 0xFFFFFF).to_bytes(3, "little") for x in v)
        want = v.astype(np.float32) / 8388608.0
    else:
        v = rng.integers(-(1 << 31), 1 << 31, size=n)
        raw, want = v.astype("<i4").tobytes(), (v.astype(np.float64) / 2147483648.0).astype(np.float32)
    path = tmp_path / f"w{width}.wav"
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(width); w.setframerate(22050)
        w.writeframes(raw)
    x, sr = load_audio(path)
    assert sr == 22050 and x.dtype == np.float32 and np.array_equal(x, want)


def test_padding_value_resolution(tmp_path, monkeypatch):
    """padding_value=None: cached hub preprocessor_config.json first, else 1.0 for the w2v-bert-2.0 family, 0.0 otherwise."""
    import json
    from speech_transcript_embeddings_b200 import processor
    assert processor.resolve_padding_value("facebook/w2v-bert-2.0") in (0.0, 1.0)        # 1.0 unless a cache says otherwise
    assert processor.resolve_padding_value("some/other-seamless-model") == 0.0
    cfg = tmp_path / "preprocessor_config.json"
    cfg.write_text(json.dumps({"padding_value": 0.5}))
    import huggingface_hub
    monkeypatch.setattr(huggingface_hub, "try_to_load_from_cache", lambda repo, name: str(cfg))
    assert processor.resolve_padding_value("facebook/w2v-bert-2.0") == 0.5


def test_default_resample_mode_is_the_reference_call_and_says_so_without_librosa():
    """R/processor.py:82-86: the drop-in never substitutes another filter silently."""
    import numpy as np
    from speech_transcript_embeddings_b200.processor import AudioTextProcessor
    proc = AudioTextProcessor(device="cpu")              # no CUDA probe on a CPU device; host-side preparation only
    assert proc.resample == "librosa"
    x = np.zeros(4800, np.float32)
    assert proc._prepare(x, 16000).dtype == np.float32   # 16 kHz input needs no resampler
    try:
        import librosa  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="resample='device'"):
            proc._prepare(x, 48000)
    with pytest.raises(ValueError):
        AudioTextProcessor(device="cpu", resample="nope")


def test_native_host_pack_gathers_ragged_segments(lib):
    """stx_host_pack: pageable arrays -> one staging buffer, any thread count, work split by bytes (no CUDA involved)."""
    import ctypes as C
    import numpy as np
    rng = np.random.default_rng(0)
    sizes = [0, 1, 7, 1023, 262144 + 3, 5, 700001, 64]
    clips = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    offs = np.zeros(len(sizes), np.int64)
    np.cumsum([(n + 31) // 32 * 32 for n in sizes[:-1]], out=offs[1:])
    total = int(offs[-1] + sizes[-1])
    src = np.array([c.ctypes.data for c in clips], np.uint64)
    nbytes = np.array(sizes, np.int64) * 4
    for threads in (1, 2, 3, 8, 64):
        dst = np.full(total + 16, np.float32(-7.0))
        rc = lib.stx_host_pack(src.ctypes.data, nbytes.ctypes.data, dst.ctypes.data, (offs * 4).ctypes.data, len(sizes), threads)
        assert rc == 0
        want = np.full(total + 16, np.float32(-7.0))
        for c, o in zip(clips, offs):
            want[o:o + c.size] = c
        assert np.array_equal(dst, want)                      # every byte in place, nothing outside the segments touched
    assert lib.stx_host_pack(None, None, None, None, 0, 4) == 0
    assert lib.stx_host_pack(None, None, None, None, 3, 4) != 0 and b"stx_host_pack" in lib.stx_last_error()


def test_pack_threads_shares_the_cores_between_ranks(monkeypatch):
    """Native packing threads per rank: the cores of the process divided by the ranks of the node (torchrun exports
    LOCAL_WORLD_SIZE), at most 16; STX_PACK_THREADS overrides it (tools/bench_e2e_pageable.py sweeps it)."""
    import os
    from speech_transcript_embeddings_b200 import feature_extraction as fx
    cores = len(os.sched_getaffinity(0))
    monkeypatch.delenv("STX_PACK_THREADS", raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert fx.pack_threads() == max(1, min(16, cores))
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert fx.pack_threads() == max(1, min(16, cores // 8))
    monkeypatch.setenv("STX_PACK_THREADS", "3")
    assert fx.pack_threads() == 3
