"""CPU: pins oracle/ against the golden fixtures generated from the third-party reference
(tests/golden/make_golden.py) and, when importable, against the installed transformers itself."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import cosine as OC
from oracle import fbank_k as OK
from oracle import logmel_w as OW
from oracle import resample as OR
from speech_transcript_embeddings_b200 import synth


def _eq_nan(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def test_fbank_k_oracle_matches_golden_bit_for_bit():
    g = load_golden("fbank_k.npz")
    n = int(g["n_clips"])
    clips = [g[f"pcm_{i}"] for i in range(n)]
    for i, c in enumerate(clips):
        with np.errstate(all="ignore"):
            x, m = OK.extract([c])
        assert _eq_nan(x, g[f"feat_{i}"]), f"clip {i}"
        assert np.array_equal(m, g[f"mask_{i}"])
        assert x.dtype == np.float32 and m.dtype == np.int32
    for pv in (0, 1):
        x, m = OK.extract(clips[:6], padding_value=float(pv))
        assert _eq_nan(x, g[f"batch_feat_pv{pv}"])
        assert np.array_equal(m, g[f"batch_mask_pv{pv}"])
    x, _ = OK.extract([clips[0]], normalize=False)
    assert _eq_nan(x, g["raw_feat_0"])
    x, m = OK.extract([np.zeros(1000, np.float32)])
    assert _eq_nan(x, g["probe_feat"]) and np.array_equal(m, g["probe_mask"])
    assert x.shape == (1, 2, 160) and not x.any()


def test_fbank_k_shapes_follow_the_survey():
    # 480000 samples -> T = 2998 -> [1, 1499, 160]; 16160 samples -> T = 99 -> 50 rows, mask sum 49
    assert OK.num_frames(480000) == 2998 and OK.num_frames(16160) == 99 and OK.num_frames(399) <= 0
    x, m = OK.extract([synth.clip("G", 16160, 0)])
    assert x.shape == (1, 50, 160) and m.sum() == 49
    assert not x[0, 49, 80:].any()          # padding half-row of the odd last frame


def test_logmel_w_oracle_matches_golden():
    g = load_golden("logmel_w.npz")
    clips = [g[f"pcm_{i}"] for i in range(int(g["n_clips"]))]
    x, m = OW.extract(clips, n_samples=16000, return_attention_mask=True)
    # golden comes from the extractor's default torch float32 path; the oracle is its float64 NumPy path
    assert x.shape == g["feat_ml16000"].shape == (3, 80, 100)
    assert np.abs(x - g["feat_ml16000"]).max() < 2e-5
    assert np.array_equal(m, g["mask_ml16000"])
    x, _ = OW.extract([clips[1]])
    assert x.shape == (1, 80, 3000)
    assert np.abs(x[:, :, ::25] - g["feat_stock_1"]).max() < 2e-5
    assert np.abs(x[:, :, -4:] - g["feat_stock_1_tail"]).max() < 2e-5


def test_cosine_oracle_matches_golden():
    g = load_golden("cosine.npz")
    a, b = g["a"], g["b"]
    assert np.abs(OC.pairwise_reference(a, b) - g["pair_unit"]).max() < 2e-7
    assert np.abs(OC.pairwise_reference(a * 3.0, b * 0.25) - g["pair_scaled"]).max() < 2e-7
    assert np.abs(OC.pairwise_reference(a * np.float32(1.00005), b) - g["pair_near_unit"]).max() < 2e-7
    S = OC.matrix_f64(a, b)
    assert np.abs(S - g["matrix_f64"]).max() < 1e-14
    assert np.abs(np.diag(S) - OC.pairwise_f64(a, b)).max() < 1e-15
    # the reference's pairwise float32 formula is the diagonal of the matrix to float32 accuracy
    assert np.abs(np.diag(S) - g["pair_unit"]).max() < 1e-6


def test_oracles_match_installed_transformers_live():
    tf = pytest.importorskip("transformers")
    fe = tf.SeamlessM4TFeatureExtractor()
    assert np.array_equal(OK.povey_window(), fe.window)
    assert np.array_equal(OK.kaldi_mel_filters(), fe.mel_filters)
    clips = synth.batch_variable(4, seed=7, whole_seconds=False, max_s=3)
    ref = fe(clips, sampling_rate=16000, return_tensors="np")
    x, m = OK.extract(clips)
    assert np.array_equal(x, ref["input_features"]) and np.array_equal(m, ref["attention_mask"])
    wfe = tf.WhisperFeatureExtractor()
    assert np.array_equal(OW.slaney_mel_filters(), wfe.mel_filters)
    batch = np.stack([np.pad(c, (0, 480000 - c.size)) for c in clips[:2]])
    assert np.array_equal(OW.extract(clips[:2])[0], wfe._np_extract_fbank_features(batch, "cpu"))


def test_library_tables_match_oracle_tables(lib):
    from speech_transcript_embeddings_b200 import ops
    assert np.abs(ops.get_table("k_window") - OK.povey_window()).max() < 1e-15
    assert np.abs(ops.get_table("k_mel").reshape(257, 80) - OK.kaldi_mel_filters()).max() < 1e-14
    assert np.abs(ops.get_table("w_window") - OW.hann_periodic()).max() < 1e-15
    assert np.abs(ops.get_table("w_mel").reshape(201, 80) - OW.slaney_mel_filters()).max() < 1e-14
    k = OK.kaldi_mel_filters()
    assert (k != 0).sum() == 501 and not k[0].any() and not k[256].any()
    assert (OW.slaney_mel_filters() != 0).sum() == 391


def test_pos_neg_oracle_matches_the_references_torch_ops():
    """oracle.cosine.pos_neg_reference vs the statements of R/training/trainer_unfreeze.py:561-563, 716-741, 924-939
    executed with torch on the CPU (float64)."""
    import torch
    import torch.nn.functional as F
    from oracle import cosine as OC
    from speech_transcript_embeddings_b200 import synth
    aud, pos = synth.embedding_pairs(33, 96, seed=2)
    neg = pos[::-1].copy() * np.float32(1.7)
    fac = np.linspace(0.7, 1.0, 33)
    for gamma, factor in ((0.35, None), (0.0, fac)):
        ref = OC.pos_neg_reference(aud, pos, neg, 0.1, gamma, factor)
        ta, tp, tn = (F.normalize(torch.from_numpy(x).double(), p=2, dim=1) for x in (aud, pos, neg))
        s_pos, s_neg = (ta * tp).sum(1), (ta * tn).sum(1)
        logits = torch.stack([s_pos, s_neg], 1) / 0.1
        per = F.cross_entropy(logits, torch.zeros(33, dtype=torch.long), reduction="none")
        if factor is not None:
            per = per * torch.from_numpy(factor)
        loss = per.mean()
        if gamma > 0:
            loss = loss + gamma * F.relu(s_neg).mean()
        assert np.abs(ref["s_pos"] - s_pos.numpy()).max() < 1e-14 and np.abs(ref["per_sample"] - per.numpy()).max() < 1e-12
        assert abs(ref["loss"] - float(loss)) < 1e-12
        assert np.abs(ref["hr_neg"] - torch.sigmoid(s_neg / 0.1).numpy()).max() < 1e-14


def test_resample_oracle_matches_golden_and_scipy_live():
    """oracle/resample.py vs the scipy.signal.resample_poly fixtures (what librosa's res_type="polyphase" runs) and, when
    scipy is importable, vs scipy itself on fresh clips; float32 summation noise only."""
    g = load_golden("resample.npz")
    for i, ((sr, n, seed), kind) in enumerate(zip(g["spec"], g["kinds"])):
        x = synth.clip(str(kind), int(n), int(seed))
        y = OR.resample_poly(x, int(sr), 16000)
        ref = g[f"y_{i}"]
        assert y.shape == ref.shape == (int(np.ceil(int(n) * 16000 / int(sr))),)
        if ref.size:
            assert np.abs(y - ref).max() <= 1e-6 * max(1.0, float(np.abs(x).max()))
    ss = pytest.importorskip("scipy.signal")
    for sr in (48000, 44100, 8000):
        x = synth.clip("U", 5000, sr)
        up, down = OR.plan(sr, 16000)
        assert np.abs(OR.resample_poly(x, sr, 16000) - ss.resample_poly(x, up, down)).max() <= 1e-6
        h_pad, _ = OR.design(up, down)
        mr = max(up, down)
        h = ss.firwin(20 * mr + 1, 1.0 / mr, window=("kaiser", 5.0)).astype(np.float32) * np.float32(up)
        assert np.array_equal(h_pad[-h.size:], h) and not h_pad[:-h.size].any()


def test_library_resample_filter_matches_oracle(lib):
    from speech_transcript_embeddings_b200 import ops
    for sr in (48000, 44100, 32000, 22050, 8000, 24000, 11025, 96000):
        p = ops.resample_plan(sr, 16000)
        h_pad, npr = OR.design(p["up"], p["down"])
        assert (p["up"], p["down"]) == OR.plan(sr, 16000) and p["n_pre_remove"] == npr and p["filter_len"] == h_pad.size
        assert np.abs(ops.resample_filter(sr, 16000) - h_pad).max() <= 1e-9      # at most a last-bit difference
    assert list(ops.resample_out_lengths(np.array([0, 1, 3, 48000, 48001]), 48000, 16000)) == [0, 1, 1, 16000, 16001]
