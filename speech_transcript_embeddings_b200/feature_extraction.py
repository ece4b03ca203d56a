"""Drop-in feature extractors with the HuggingFace call signature the reference uses.

``B200SeamlessM4TFeatureExtractor`` mirrors ``SeamlessM4TFeatureExtractor.__call__``
(TF/models/seamless_m4t/feature_extraction_seamless_m4t.py:141-302), i.e. what
``AutoFeatureExtractor.from_pretrained("facebook/w2v-bert-2.0")`` returns at R/processor.py:36 and
R/training/trainer_unfreeze.py:1388 (recipe K).  ``B200WhisperFeatureExtractor`` mirrors
``WhisperFeatureExtractor.__call__`` (TF/models/whisper/feature_extraction_whisper.py:189-342), the
recipe BASELINE.json's north_star lists (recipe W).

Same arguments, same keys (``input_features``, ``attention_mask``), same shapes and dtypes, same
exceptions.  One deliberate difference: ``return_tensors="pt"`` tensors live on the extractor's CUDA
device (the reference moves them there right after the call, R/processor.py:118-121, so ``.to(device)``
becomes a no-op); ``return_tensors="np"`` / ``None`` copies them back to host NumPy arrays.

Host work here is only: pack the clips into one pinned buffer, one H2D copy, one C-ABI call.
"""
from __future__ import annotations

import logging
from collections import UserDict
from typing import Sequence

import numpy as np
import torch

from . import ops

logger = logging.getLogger(__name__)

_ALIGN = 32   # clips start on 128-byte boundaries inside the packed buffer


class BatchFeature(UserDict):
    """Minimal stand-in for transformers.BatchFeature: dict access, ``.keys()``, ``.get``, ``in``,
    attribute access and ``.to(device)`` — what R/processor.py:45, 108-121 and
    R/training/trainer_unfreeze.py:862-866 use."""

    def __getattr__(self, item):
        try:
            return self.data[item]
        except KeyError:
            raise AttributeError(item) from None

    def to(self, *args, **kwargs):
        self.data = {k: (v.to(*args, **kwargs) if isinstance(v, torch.Tensor) else v) for k, v in self.data.items()}
        return self


class PackedClips:
    """Clips packed back to back (128-byte aligned) in one float32 buffer, plus per-clip offsets/lengths."""

    def __init__(self, pcm: torch.Tensor, offsets: np.ndarray, lengths: np.ndarray):
        self.pcm = pcm                      # pinned host tensor or CUDA tensor, float32 [total]
        self.offsets = offsets              # int64 [B] (host)
        self.lengths = lengths              # int32 [B] (host)

    @property
    def batch_size(self) -> int:
        return int(self.lengths.size)

    @property
    def max_length(self) -> int:
        return int(self.lengths.max()) if self.lengths.size else 0


def _layout(lengths: np.ndarray):
    padded = (lengths.astype(np.int64) + (_ALIGN - 1)) // _ALIGN * _ALIGN
    offsets = np.zeros(lengths.size, np.int64)
    if lengths.size > 1:
        np.cumsum(padded[:-1], out=offsets[1:])
    total = int(padded.sum())
    return offsets, total


class _HostStage:
    """Grow-only pinned staging buffers, reused across calls (guarded by an event)."""

    def __init__(self):
        self.pcm = None
        self.meta = None
        self.event = None

    def get(self, n_float: int, n_meta: int):
        if self.event is not None:
            self.event.synchronize()        # the previous call's H2D copies have drained
        if self.pcm is None or self.pcm.numel() < n_float:
            self.pcm = torch.empty(max(n_float, 1), dtype=torch.float32, pin_memory=True)
        if self.meta is None or self.meta.numel() < n_meta:
            self.meta = torch.empty(max(n_meta, 1), dtype=torch.int64, pin_memory=True)
        return self.pcm, self.meta


def _as_clip_list(raw_speech, max_dims: int, cls_name: str):
    """Batching rules of the HF extractors (…seamless_m4t.py:232-251): returns a list of float32 arrays."""
    if isinstance(raw_speech, torch.Tensor) and raw_speech.dim() > 1:
        raw_speech = list(raw_speech)
    is_batched_numpy = isinstance(raw_speech, np.ndarray) and raw_speech.ndim > 1
    if is_batched_numpy and raw_speech.ndim > max_dims:
        raise ValueError(f"Only mono-channel or stereo-channel audio is supported for input to {cls_name}")
    is_batched = is_batched_numpy or (
        isinstance(raw_speech, (list, tuple)) and len(raw_speech) > 0
        and isinstance(raw_speech[0], (torch.Tensor, np.ndarray, tuple, list)))
    if not is_batched:
        raw_speech = [raw_speech]
    clips = []
    for s in raw_speech:
        if isinstance(s, torch.Tensor):
            s = s.detach().cpu().numpy()
        a = np.asarray(s, dtype=np.float32)
        if a.ndim == 2:                     # stereo: keep channel 0 (…seamless_m4t.py:121-122)
            a = a[0]
        clips.append(np.ascontiguousarray(a.reshape(-1)))
    return clips


class _B200ExtractorBase:
    sampling_rate = 16000

    def __init__(self, device=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        self.device = torch.device(device) if device is not None else None
        self._stage = _HostStage()

    def _check_rate(self, sampling_rate):
        if sampling_rate is not None:
            if sampling_rate != self.sampling_rate:
                raise ValueError(
                    f"The model corresponding to this feature extractor: {self.__class__.__name__} was trained using a "
                    f"sampling rate of {self.sampling_rate}. Please make sure that the provided `raw_speech` input "
                    f"was sampled with {self.sampling_rate} and not {sampling_rate}.")
        else:
            logger.warning(
                "It is strongly recommended to pass the `sampling_rate` argument to `%s()`. "
                "Failing to do so can result in silent errors that might be hard to debug.", self.__class__.__name__)

    def _device(self) -> torch.device:
        if self.device is None or self.device.type != "cuda":
            from ._lib import StxError
            raise StxError("this feature extractor needs a CUDA device (sm_100a); there is no CPU fallback")
        return self.device

    # -- host -> device ----------------------------------------------------------------------
    def pack(self, clips: Sequence[np.ndarray]) -> PackedClips:
        """Copy clips into one pinned host buffer (reused across calls)."""
        lengths = np.fromiter((c.size for c in clips), dtype=np.int32, count=len(clips))
        offsets, total = _layout(lengths)
        pcm, _ = self._stage.get(total, 2 * len(clips))
        view = pcm.numpy()
        for c, o in zip(clips, offsets):
            view[o:o + c.size] = c
        return PackedClips(pcm[:total], offsets, lengths)

    def to_device(self, packed: PackedClips):
        """(pcm, offsets, lengths) on the device; one async copy for the PCM, one for the metadata."""
        dev = self._device()
        B = packed.batch_size
        if packed.pcm.is_cuda:
            pcm_d = packed.pcm
        else:
            pcm_d = torch.empty(packed.pcm.numel(), dtype=torch.float32, device=dev)
            pcm_d.copy_(packed.pcm, non_blocking=True)
        _, meta = self._stage.get(0, 2 * B) if not packed.pcm.is_cuda else (None, torch.empty(2 * B, dtype=torch.int64, pin_memory=True))
        mv = meta.numpy()
        mv[:B] = packed.offsets
        mv[B:2 * B].view(np.int32)[:B] = packed.lengths       # lengths live in the low half of the second block
        meta_d = torch.empty(2 * B, dtype=torch.int64, device=dev)
        meta_d.copy_(meta[:2 * B], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        self._stage.event = ev
        offsets_d = meta_d[:B]
        lengths_d = meta_d[B:2 * B].view(torch.int32)[:B]
        return pcm_d, offsets_d, lengths_d

    @staticmethod
    def _finish(data: dict, return_tensors):
        if return_tensors in ("pt", "torch"):
            return BatchFeature(data)
        if return_tensors in (None, "np", "numpy"):
            return BatchFeature({k: v.cpu().numpy() for k, v in data.items()})
        raise ValueError(f"return_tensors={return_tensors!r} is not supported (use 'pt' or 'np')")


class B200SeamlessM4TFeatureExtractor(_B200ExtractorBase):
    """Recipe K: 80-bin Kaldi fbank, per-clip per-bin CMVN, pad to a multiple of 2, stride-2 stacking."""

    model_input_names = ["input_features", "attention_mask"]

    def __init__(self, feature_size=80, sampling_rate=16000, num_mel_bins=80, padding_value=0.0, stride=2,
                 device=None, **kwargs):
        if feature_size != 80 or num_mel_bins != 80 or stride != 2 or sampling_rate != 16000:
            raise ValueError("the sm_100a kernel implements the facebook/w2v-bert-2.0 configuration only: "
                             "80 mel bins, stride 2, 16 kHz")
        super().__init__(device)
        self.feature_size = feature_size
        self.sampling_rate = sampling_rate
        self.num_mel_bins = num_mel_bins
        self.padding_value = float(padding_value)
        self.stride = stride
        self.return_attention_mask = True
        self.padding_side = "right"

    @staticmethod
    def _padded_frames(frames: np.ndarray, padding, max_length, truncation, pad_to_multiple_of):
        """T_pad (raw frames) following SequenceFeatureExtractor.pad (TF/feature_extraction_sequence_utils.py
        :176-219, 255-291, 293-334) and the remainder drop of …seamless_m4t.py:281-285."""
        if padding is True or padding == "longest":
            strategy = "longest"
        elif padding == "max_length":
            strategy = "max_length"
            if max_length is None:
                raise ValueError("padding='max_length' needs max_length")
        elif padding is False or padding == "do_not_pad":
            strategy = "do_not_pad"
        else:
            raise ValueError(f"unsupported padding strategy {padding!r}")
        kept = frames.copy()
        if truncation:
            if max_length is None:
                raise ValueError("When setting ``truncation=True``, make sure that ``max_length`` is defined.")
            lim = max_length
            if pad_to_multiple_of is not None and lim % pad_to_multiple_of != 0:
                lim = ((lim // pad_to_multiple_of) + 1) * pad_to_multiple_of
            kept = np.minimum(kept, lim)
        if strategy == "do_not_pad":
            if kept.size and (kept != kept[0]).any():
                raise ValueError("padding=False with clips of different lengths cannot be stacked into one tensor")
            tgt = int(kept[0]) if kept.size else 0
        else:
            tgt = int(kept.max()) if strategy == "longest" else int(max_length)
            if pad_to_multiple_of is not None and tgt % pad_to_multiple_of != 0:
                tgt = ((tgt // pad_to_multiple_of) + 1) * pad_to_multiple_of
            if kept.size and int(kept.max()) > tgt:
                raise ValueError("clips longer than max_length need truncation=True")
        return tgt - (tgt % 2), kept

    def __call__(self, raw_speech, padding=True, pad_to_multiple_of=2, max_length=None, truncation=False,
                 return_tensors=None, sampling_rate=None, return_attention_mask=None,
                 do_normalize_per_mel_bins=True, **kwargs):
        self._check_rate(sampling_rate)
        self._device()                       # fail loudly before any host work when there is no GPU
        return_attention_mask = self.return_attention_mask if return_attention_mask is None else return_attention_mask
        if isinstance(raw_speech, PackedClips):
            packed = raw_speech
        else:
            packed = self.pack(_as_clip_list(raw_speech, 3, self.__class__.__name__))
        pcm_d, off_d, len_d = self.to_device(packed)
        frames = np.array([ops.k_num_frames(int(n)) for n in packed.lengths], dtype=np.int64)
        T_pad, _ = self._padded_frames(frames, padding, max_length, truncation, pad_to_multiple_of)
        feats, mask = ops.fbank_k(pcm_d, off_d, len_d, packed.max_length, T_pad, self.padding_value,
                                  bool(do_normalize_per_mel_bins), want_mask=bool(return_attention_mask))
        data = {"input_features": feats}
        if return_attention_mask:
            data["attention_mask"] = mask
        return self._finish(data, return_tensors)


class B200WhisperFeatureExtractor(_B200ExtractorBase):
    """Recipe W: Whisper log-mel, [B, 80, 3000] for the stock 30 s chunk."""

    model_input_names = ["input_features"]

    def __init__(self, feature_size=80, sampling_rate=16000, hop_length=160, chunk_length=30, n_fft=400,
                 padding_value=0.0, dither=0.0, return_attention_mask=False, device=None, **kwargs):
        if feature_size != 80 or sampling_rate != 16000 or hop_length != 160 or n_fft != 400:
            raise ValueError("the sm_100a kernel implements the stock Whisper front end only: "
                             "80 mel bins, n_fft 400, hop 160, 16 kHz")
        if dither != 0.0:
            raise ValueError("dither is not supported (the reference's default is 0.0)")
        if padding_value != 0.0:
            raise ValueError("Whisper pads the waveform with 0.0")
        super().__init__(device)
        self.feature_size = feature_size
        self.sampling_rate = sampling_rate
        self.hop_length = hop_length
        self.chunk_length = chunk_length
        self.n_fft = n_fft
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        self.padding_value = 0.0
        self.return_attention_mask = return_attention_mask

    def __call__(self, raw_speech, truncation=True, pad_to_multiple_of=None, return_tensors=None,
                 return_attention_mask=None, padding="max_length", max_length=None, sampling_rate=None,
                 do_normalize=None, device="cpu", **kwargs):
        self._check_rate(sampling_rate)
        self._device()
        if do_normalize:
            raise ValueError("do_normalize (waveform zero-mean/unit-variance) is not part of the hot path")
        if padding != "max_length" or not truncation:
            raise ValueError("only the stock padding='max_length', truncation=True call is supported")
        n_samples = int(max_length) if max_length else self.n_samples
        if pad_to_multiple_of is not None and n_samples % pad_to_multiple_of != 0:
            n_samples = ((n_samples // pad_to_multiple_of) + 1) * pad_to_multiple_of
        if n_samples % self.hop_length != 0:
            raise ValueError("max_length must be a multiple of hop_length (160)")
        if isinstance(raw_speech, PackedClips):
            packed = raw_speech
        else:
            packed = self.pack(_as_clip_list(raw_speech, 2, self.__class__.__name__))
        pcm_d, off_d, len_d = self.to_device(packed)
        want_mask = bool(return_attention_mask if return_attention_mask is not None else self.return_attention_mask)
        feats, mask = ops.logmel_w(pcm_d, off_d, len_d, n_samples, want_mask=want_mask)
        data = {"input_features": feats}
        if want_mask:
            data["attention_mask"] = mask
        return self._finish(data, return_tensors)
