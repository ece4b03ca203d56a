"""Thin torch-tensor front of the C ABI (include/stx_b200.h).

Every function takes CUDA tensors, allocates outputs/workspaces with torch (so ownership stays with
torch's caching allocator) and launches on torch's current stream.  PyTorch is plumbing here: device
memory and streams.  All arithmetic happens in libstx_b200.so; there is no fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

K_FRAME, K_HOP, K_NMEL = 400, 160, 80
W_HOP, W_NMEL = 160, 80


def _require_cuda(t: torch.Tensor, name: str, dtype=None) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.StxError(f"{name} must be a CUDA tensor (libstx_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def k_num_frames(n: int) -> int:
    """Frames of a clip of n samples, centre=False framing 400/160 (TF/audio_utils.py:778)."""
    return 1 + (n - K_FRAME) // K_HOP if n >= K_FRAME else 0


def get_table(name: str) -> np.ndarray:
    lib = _lib.load()
    sizes = {"k_window": 400, "k_mel": 257 * 80, "w_window": 400, "w_mel": 201 * 80}
    out = np.empty(sizes[name], np.float64)
    n = lib.stx_get_table(name.encode(), out.ctypes.data_as(C.c_void_p), out.size)
    if n != out.size:
        _lib.check(int(n) if n < 0 else -1, f"stx_get_table({name})")
    return out


def peak_abs(pcm: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """max(1, max|x|) per clip as float32 [B] (the divisor of R/processor.py:91-92)."""
    lib = _lib.load()
    _require_cuda(pcm, "pcm", torch.float32)
    _require_cuda(offsets, "offsets", torch.int64)
    _require_cuda(lengths, "lengths", torch.int32)
    B = lengths.numel()
    peak = torch.empty(B, dtype=torch.float32, device=pcm.device)
    with torch.cuda.device(pcm.device):
        _lib.check(lib.stx_peak_abs(pcm.data_ptr(), offsets.data_ptr(), lengths.data_ptr(), B, peak.data_ptr(),
                                    _stream_ptr(pcm.device)), "stx_peak_abs")
    return peak


def resample_plan(orig_sr: int, target_sr: int) -> dict:
    """up / down / taps per polyphase branch / leading outputs removed / padded filter length of the resampler."""
    lib = _lib.load()
    v = [C.c_int(0) for _ in range(5)]
    _lib.check(lib.stx_resample_plan(int(orig_sr), int(target_sr), *[C.byref(x) for x in v]), "stx_resample_plan")
    return dict(zip(("up", "down", "taps_per_phase", "n_pre_remove", "filter_len"), (x.value for x in v)))


def resample_filter(orig_sr: int, target_sr: int) -> np.ndarray:
    """The padded float32 low-pass the resampler applies (what scipy.signal.resample_poly hands to upfirdn)."""
    lib = _lib.load()
    n = resample_plan(orig_sr, target_sr)["filter_len"]
    out = np.empty(n, np.float32)
    got = lib.stx_resample_filter(int(orig_sr), int(target_sr), out.ctypes.data_as(C.c_void_p), n)
    if got != n:
        _lib.check(int(got) if got < 0 else -1, "stx_resample_filter")
    return out


def resample_out_lengths(lengths: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """ceil(n * target_sr / orig_sr) per clip (librosa.resample's fix_length target), int32."""
    g = int(np.gcd(int(orig_sr), int(target_sr)))
    up, down = int(target_sr) // g, int(orig_sr) // g
    return ((np.asarray(lengths, np.int64) * up + down - 1) // down).astype(np.int32)


def resample_poly(pcm: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor, lengths_host: np.ndarray,
                  orig_sr: int, target_sr: int, want_peak: bool = True, align: int = 32):
    """Device-side librosa.resample(..., res_type="polyphase") of packed clips (R/processor.py:82-86).

    ``lengths_host`` is the host copy of ``lengths`` (the output layout is computed on the host).  Returns
    (pcm_out float32 [total], offsets_out int64 [B] CUDA, lengths_out int32 [B] CUDA, lengths_out_host int32 [B],
    peak float32 [B] = max(1, max|y|) or None); output clips start on ``align``-float boundaries.
    """
    lib = _lib.load()
    _require_cuda(pcm, "pcm", torch.float32)
    _require_cuda(offsets, "offsets", torch.int64)
    _require_cuda(lengths, "lengths", torch.int32)
    B = lengths.numel()
    dev = pcm.device
    out_len = resample_out_lengths(lengths_host, orig_sr, target_sr)
    padded = (out_len.astype(np.int64) + (align - 1)) // align * align
    out_off = np.zeros(B, np.int64)
    if B > 1:
        np.cumsum(padded[:-1], out=out_off[1:])
    total = int(padded.sum())
    meta = torch.empty(2 * B, dtype=torch.int64, pin_memory=True)
    mv = meta.numpy()
    mv[:B] = out_off
    mv[B:].view(np.int32)[:B] = out_len
    meta_d = meta.to(dev, non_blocking=True)
    off_d = meta_d[:B]
    len_d = meta_d[B:].view(torch.int32)[:B]
    out = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
    peak = torch.empty(B, dtype=torch.float32, device=dev) if want_peak else None
    with torch.cuda.device(dev):
        _lib.check(lib.stx_resample_poly(pcm.data_ptr(), offsets.data_ptr(), lengths.data_ptr(), B, int(orig_sr),
                                         int(target_sr), out.data_ptr(), off_d.data_ptr(), len_d.data_ptr(),
                                         int(out_len.max()) if B else 0, peak.data_ptr() if peak is not None else None,
                                         _stream_ptr(dev)), "stx_resample_poly")
    return out, off_d, len_d, out_len, peak


def fbank_k(pcm: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor, max_length: int, T_pad: int,
            padding_value: float = 0.0, normalize: bool = True, peak: torch.Tensor | None = None,
            want_mask: bool = True, out: torch.Tensor | None = None, mask: torch.Tensor | None = None,
            uniform: bool = False):
    """Recipe K on device-resident packed PCM.

    pcm float32 [total], offsets int64 [B], lengths int32 [B] (all CUDA); ``max_length`` is the host's
    max over lengths; ``uniform=True`` promises that every clip has exactly ``max_length`` samples (skips the
    device-side compaction of work items that ragged batches need).  Returns (input_features float32
    [B, T_pad/2, 160], attention_mask int32 [B, T_pad/2] or None).
    """
    lib = _lib.load()
    _require_cuda(pcm, "pcm", torch.float32)
    _require_cuda(offsets, "offsets", torch.int64)
    _require_cuda(lengths, "lengths", torch.int32)
    if T_pad < 0 or T_pad % 2:
        raise ValueError("T_pad must be even and >= 0")
    B = lengths.numel()
    dev = pcm.device
    if out is None:
        out = torch.empty((B, T_pad // 2, 2 * K_NMEL), dtype=torch.float32, device=dev)
    else:
        _require_cuda(out, "out", torch.float32)
        if out.numel() != B * T_pad * K_NMEL:
            raise ValueError("out has the wrong size")
    if want_mask and mask is None:
        mask = torch.empty((B, T_pad // 2), dtype=torch.int32, device=dev)
    if mask is not None:
        _require_cuda(mask, "mask", torch.int32)
    if peak is not None:
        _require_cuda(peak, "peak", torch.float32)
    nbytes = C.c_size_t(0)
    _lib.check(lib.stx_fbank_k_workspace(B, int(max_length), C.byref(nbytes)), "stx_fbank_k_workspace")
    ws = torch.empty(max(int(nbytes.value), 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.stx_fbank_k(pcm.data_ptr(), offsets.data_ptr(), lengths.data_ptr(), B,
                                   -int(max_length) if uniform else int(max_length),
                                   peak.data_ptr() if peak is not None else None, int(T_pad),
                                   float(padding_value), int(bool(normalize)), out.data_ptr(),
                                   mask.data_ptr() if mask is not None else None, ws.data_ptr(), ws.numel(),
                                   _stream_ptr(dev)), "stx_fbank_k")
    return out, mask


def fbank_k_projection(pcm: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor, max_length: int, T_pad: int,
                       ln_weight: torch.Tensor, ln_bias: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None,
                       eps: float = 1e-5, padding_value: float = 0.0, peak: torch.Tensor | None = None,
                       want_features: bool = False, want_mask: bool = True, uniform: bool = False):
    """Recipe K fused with the encoder's input stage: packed PCM -> (hidden float32 [B, T_pad/2, out_dim], input_features
    [B, T_pad/2, 160] or None, attention_mask int32 [B, T_pad/2] or None).  ``hidden`` is what
    ``Wav2Vec2BertFeatureProjection`` returns for the extractor's ``input_features`` (LayerNorm(160) + Linear)."""
    lib = _lib.load()
    _require_cuda(pcm, "pcm", torch.float32)
    _require_cuda(offsets, "offsets", torch.int64)
    _require_cuda(lengths, "lengths", torch.int32)
    for name, t_ in (("ln_weight", ln_weight), ("ln_bias", ln_bias), ("weight", weight)):
        _require_cuda(t_, name, torch.float32)
    if bias is not None:
        _require_cuda(bias, "bias", torch.float32)
    if T_pad < 0 or T_pad % 2:
        raise ValueError("T_pad must be even and >= 0")
    out_dim = int(weight.shape[0])
    if tuple(weight.shape) != (out_dim, 2 * K_NMEL) or tuple(ln_weight.shape) != (2 * K_NMEL,) or tuple(ln_bias.shape) != (2 * K_NMEL,):
        raise ValueError("the projection takes the 160 stacked features: weight [out_dim, 160], LayerNorm parameters [160]")
    B = lengths.numel()
    dev = pcm.device
    hidden = torch.empty((B, T_pad // 2, out_dim), dtype=torch.float32, device=dev)
    feats = torch.empty((B, T_pad // 2, 2 * K_NMEL), dtype=torch.float32, device=dev) if want_features else None
    mask = torch.empty((B, T_pad // 2), dtype=torch.int32, device=dev) if want_mask else None
    ml = -int(max_length) if uniform else int(max_length)
    nbytes = C.c_size_t(0)
    _lib.check(lib.stx_fbank_k_projection_workspace(B, ml, int(T_pad), out_dim, int(bool(want_features)), C.byref(nbytes)),
               "stx_fbank_k_projection_workspace")
    ws = torch.empty(max(int(nbytes.value), 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.stx_fbank_k_projection(pcm.data_ptr(), offsets.data_ptr(), lengths.data_ptr(), B, ml,
                                              peak.data_ptr() if peak is not None else None, int(T_pad), float(padding_value),
                                              ln_weight.data_ptr(), ln_bias.data_ptr(), float(eps), weight.data_ptr(),
                                              bias.data_ptr() if bias is not None else None, out_dim, hidden.data_ptr(),
                                              feats.data_ptr() if feats is not None else None,
                                              mask.data_ptr() if mask is not None else None, ws.data_ptr(), ws.numel(),
                                              _stream_ptr(dev)), "stx_fbank_k_projection")
    return hidden, feats, mask


def fbank_k_collate(pcm: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor, max_length: int, T_pad: int,
                    padding_value: float = 0.0, out: torch.Tensor | None = None, mask: torch.Tensor | None = None):
    """Recipe K with the trainer's collate fused in (R/training/trainer_unfreeze.py:855-866, 898-908):
    (input_values float32 [B, T_pad/2, 160] zero-padded, attention_mask_audio int64 [B, T_pad/2])."""
    lib = _lib.load()
    _require_cuda(pcm, "pcm", torch.float32)
    _require_cuda(offsets, "offsets", torch.int64)
    _require_cuda(lengths, "lengths", torch.int32)
    if T_pad < 0 or T_pad % 2:
        raise ValueError("T_pad must be even and >= 0")
    B = lengths.numel()
    dev = pcm.device
    if out is None:
        out = torch.empty((B, T_pad // 2, 2 * K_NMEL), dtype=torch.float32, device=dev)
    else:
        _require_cuda(out, "out", torch.float32)
    if mask is None:
        mask = torch.empty((B, T_pad // 2), dtype=torch.int64, device=dev)
    else:
        _require_cuda(mask, "mask", torch.int64)
    nbytes = C.c_size_t(0)
    _lib.check(lib.stx_fbank_k_workspace(B, int(max_length), C.byref(nbytes)), "stx_fbank_k_workspace")
    ws = torch.empty(max(int(nbytes.value), 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.stx_fbank_k_collate(pcm.data_ptr(), offsets.data_ptr(), lengths.data_ptr(), B, int(max_length),
                                           int(T_pad), float(padding_value), out.data_ptr(), mask.data_ptr(),
                                           ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "stx_fbank_k_collate")
    return out, mask


def logmel_w(pcm: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor, n_samples: int = 480000,
             want_mask: bool = False, peak: torch.Tensor | None = None, out: torch.Tensor | None = None,
             mask: torch.Tensor | None = None):
    """Recipe W on device-resident packed PCM -> (float32 [B, 80, n_samples/160], int32 mask or None)."""
    lib = _lib.load()
    _require_cuda(pcm, "pcm", torch.float32)
    _require_cuda(offsets, "offsets", torch.int64)
    _require_cuda(lengths, "lengths", torch.int32)
    B = lengths.numel()
    dev = pcm.device
    T = n_samples // W_HOP
    if out is None:
        out = torch.empty((B, W_NMEL, T), dtype=torch.float32, device=dev)
    else:
        _require_cuda(out, "out", torch.float32)
    if want_mask and mask is None:
        mask = torch.empty((B, T), dtype=torch.int32, device=dev)
    if mask is not None:
        _require_cuda(mask, "mask", torch.int32)
    if peak is not None:
        _require_cuda(peak, "peak", torch.float32)
    nbytes = C.c_size_t(0)
    _lib.check(lib.stx_logmel_w_workspace(B, int(n_samples), C.byref(nbytes)), "stx_logmel_w_workspace")
    ws = torch.empty(max(int(nbytes.value), 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.stx_logmel_w(pcm.data_ptr(), offsets.data_ptr(), lengths.data_ptr(), B, int(n_samples),
                                    peak.data_ptr() if peak is not None else None, out.data_ptr(), mask.data_ptr() if mask is not None else None,
                                    ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "stx_logmel_w")
    return out, mask


def _cosine_ws(N: int, M: int, D: int, dev) -> torch.Tensor:
    lib = _lib.load()
    nbytes = C.c_size_t(0)
    _lib.check(lib.stx_cosine_workspace(N, M, D, C.byref(nbytes)), "stx_cosine_workspace")
    return torch.empty(max(int(nbytes.value), 256), dtype=torch.uint8, device=dev)


def cosine_pairwise(a: torch.Tensor, b: torch.Tensor, always_normalize: bool = False) -> torch.Tensor:
    """s[i] = <normalize(a_i), normalize(b_i)>, float32 [N] (R/processor.py:148-159)."""
    lib = _lib.load()
    _require_cuda(a, "a", torch.float32)
    _require_cuda(b, "b", torch.float32)
    if a.dim() != 2 or a.shape != b.shape:
        raise ValueError(f"expected two [N, D] tensors of equal shape, got {tuple(a.shape)} and {tuple(b.shape)}")
    N, D = a.shape
    s = torch.empty(N, dtype=torch.float32, device=a.device)
    ws = _cosine_ws(N, N, D, a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.stx_cosine_pairwise(a.data_ptr(), b.data_ptr(), N, D, int(bool(always_normalize)),
                                           s.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(a.device)),
                   "stx_cosine_pairwise")
    return s


def cosine_nxm(a: torch.Tensor, b: torch.Tensor, always_normalize: bool = True,
               out: torch.Tensor | None = None) -> torch.Tensor:
    """S[i, j] = <normalize(a_i), normalize(b_j)>, float32 [N, M]."""
    lib = _lib.load()
    _require_cuda(a, "a", torch.float32)
    _require_cuda(b, "b", torch.float32)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError(f"expected [N, D] and [M, D], got {tuple(a.shape)} and {tuple(b.shape)}")
    N, D = a.shape
    M = b.shape[0]
    if out is None:
        out = torch.empty((N, M), dtype=torch.float32, device=a.device)
    else:
        _require_cuda(out, "out", torch.float32)
        if out.shape != (N, M):
            raise ValueError("out has the wrong shape")
    ws = _cosine_ws(N, M, D, a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.stx_cosine_nxm(a.data_ptr(), b.data_ptr(), N, M, D, int(bool(always_normalize)),
                                      out.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(a.device)),
                   "stx_cosine_nxm")
    return out


def cosine_topk(a: torch.Tensor, b: torch.Tensor, k: int, always_normalize: bool = True):
    """For every row of ``a`` the ``k`` (1..8) best rows of ``b`` under the cosine score: (scores float32 [N, k], indices
    int32 [N, k]), ordered by (score descending, index ascending), without materialising the [N, M] matrix."""
    lib = _lib.load()
    _require_cuda(a, "a", torch.float32)
    _require_cuda(b, "b", torch.float32)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError(f"expected [N, D] and [M, D], got {tuple(a.shape)} and {tuple(b.shape)}")
    N, D = a.shape
    M = b.shape[0]
    val = torch.empty((N, k), dtype=torch.float32, device=a.device)
    idx = torch.empty((N, k), dtype=torch.int32, device=a.device)
    nbytes = C.c_size_t(0)
    _lib.check(lib.stx_cosine_topk_workspace(N, M, D, C.byref(nbytes)), "stx_cosine_topk_workspace")
    ws = torch.empty(max(int(nbytes.value), 256), dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.stx_cosine_topk(a.data_ptr(), b.data_ptr(), N, M, D, int(bool(always_normalize)), int(k),
                                       val.data_ptr(), idx.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(a.device)),
                   "stx_cosine_topk")
    return val, idx


def score_pos_neg(aud: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, temperature: float = 0.1,
                  corrupt_gamma: float = 0.35, alignment_factor: torch.Tensor | None = None) -> dict:
    """Evaluation-time scoring of a batch (R/training/trainer_unfreeze.py:1206-1216, 716-741), forward only:
    ``s_pos``, ``s_neg``, ``hr_pos``, ``hr_neg`` (sigmoid(s / temperature)), ``per_sample`` (2-way InfoNCE) as
    float32 [B] and the scalar ``loss``, all on the GPU."""
    lib = _lib.load()
    for name, t in (("aud", aud), ("pos", pos), ("neg", neg)):
        _require_cuda(t, name, torch.float32)
    if aud.dim() != 2 or aud.shape != pos.shape or aud.shape != neg.shape:
        raise ValueError("expected three [B, D] tensors of equal shape")
    if alignment_factor is not None:
        _require_cuda(alignment_factor, "alignment_factor", torch.float32)
        if alignment_factor.shape != (aud.shape[0],):
            raise ValueError("alignment_factor must be [B]")
    B, D = aud.shape
    dev = aud.device
    out = {k: torch.empty(B, dtype=torch.float32, device=dev) for k in ("s_pos", "s_neg", "hr_pos", "hr_neg", "per_sample")}
    out["loss"] = torch.zeros((), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.stx_score_pos_neg(aud.data_ptr(), pos.data_ptr(), neg.data_ptr(), B, D, float(temperature),
                                         float(corrupt_gamma),
                                         alignment_factor.data_ptr() if alignment_factor is not None else None,
                                         out["s_pos"].data_ptr(), out["s_neg"].data_ptr(), out["hr_pos"].data_ptr(),
                                         out["hr_neg"].data_ptr(), out["per_sample"].data_ptr(), out["loss"].data_ptr(),
                                         _stream_ptr(dev)), "stx_score_pos_neg")
    return out


def feature_projection(x: torch.Tensor, ln_weight: torch.Tensor, ln_bias: torch.Tensor, weight: torch.Tensor,
                       bias: torch.Tensor | None = None, eps: float = 1e-5, return_norm: bool = True):
    """``Wav2Vec2BertFeatureProjection.forward`` in eval mode (TF/models/wav2vec2_bert/modeling_wav2vec2_bert.py:118-130):
    x [..., in_dim] -> (hidden [..., out_dim], norm [..., in_dim] or None); the Linear runs on tcgen05."""
    lib = _lib.load()
    _require_cuda(x, "x", torch.float32)
    for name, t in (("ln_weight", ln_weight), ("ln_bias", ln_bias), ("weight", weight)):
        _require_cuda(t, name, torch.float32)
    if bias is not None:
        _require_cuda(bias, "bias", torch.float32)
    in_dim = x.shape[-1]
    out_dim = weight.shape[0]
    if weight.shape != (out_dim, in_dim) or ln_weight.shape != (in_dim,) or ln_bias.shape != (in_dim,):
        raise ValueError("shape mismatch between x, the LayerNorm parameters and the Linear weight")
    rows = x.numel() // in_dim
    dev = x.device
    hidden = torch.empty(x.shape[:-1] + (out_dim,), dtype=torch.float32, device=dev)
    norm = torch.empty_like(x) if return_norm else None
    nbytes = C.c_size_t(0)
    _lib.check(lib.stx_feature_projection_workspace(rows, in_dim, out_dim, C.byref(nbytes)), "stx_feature_projection_workspace")
    ws = torch.empty(max(int(nbytes.value), 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.stx_feature_projection(x.data_ptr(), ln_weight.data_ptr(), ln_bias.data_ptr(), float(eps),
                                              weight.data_ptr(), bias.data_ptr() if bias is not None else None, rows, in_dim,
                                              out_dim, hidden.data_ptr(), norm.data_ptr() if norm is not None else None,
                                              ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "stx_feature_projection")
    return hidden, norm
