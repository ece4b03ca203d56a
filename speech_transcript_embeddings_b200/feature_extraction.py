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

Host work here is only: pack the clips into one pinned buffer, H2D, C-ABI calls.  When the result is wanted on
the host (``return_tensors="np"``, or ``output="host"`` for pinned CPU tensors like the reference's own CPU
tensors) the batch is cut into chunks of clips and three streams overlap H2D(i+1) | kernels(i) | D2H(i-1), so
the call costs about max(H2D, D2H) over PCIe instead of their sum.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import threading
import weakref
from collections import UserDict
from typing import Sequence

import numpy as np
import torch

from . import _lib, ops

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


def pack_threads() -> int:
    """Native threads one packing job may use: the cores this process may run on, shared between the ranks of the node
    (torchrun exports LOCAL_WORLD_SIZE), at most 16 -- beyond that the copy is bound by host memory bandwidth.
    ``STX_PACK_THREADS`` overrides it."""
    forced = os.environ.get("STX_PACK_THREADS")
    if forced:
        return max(1, int(forced))
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(1, min(16, cores // ranks))


class PackedClips:
    """Clips packed back to back (128-byte aligned) in one float32 buffer, plus per-clip offsets/lengths.

    ``bounds`` (optional): the clip ranges of the pipeline chunks.  ``job`` (optional): the native packing job
    (``stx_host_pack_begin``) that is still filling the buffer chunk by chunk; consumers call ``wait(i)`` right before they
    copy chunk ``i`` to the device.  ``stage``: the pinned staging buffer of the extractor's pool this object owns (it goes
    back to the pool when the object dies and the copies out of it have drained)."""

    def __init__(self, pcm: torch.Tensor, offsets: np.ndarray, lengths: np.ndarray, bounds=None, job=None, stage=None,
                 keep=None):
        self.pcm = pcm                      # pinned host tensor or CUDA tensor, float32 [total]
        self.offsets = offsets              # int64 [B] (host)
        self.lengths = lengths              # int32 [B] (host)
        self.bounds = bounds
        self.stage = stage
        self._job = job                     # native handle, or None once the job has been joined
        self._keep = keep                   # the source arrays and the job's argument arrays stay alive while it runs

    def wait(self, i: int | None = None):
        if self._job is None:
            return
        lib = _lib.load()
        _lib.check(lib.stx_host_pack_wait(self._job, -1 if i is None else int(i)), "stx_host_pack_wait")
        if i is None or (self.bounds is not None and i >= len(self.bounds) - 1):
            self._finish()

    def _finish(self):
        if self._job is not None:
            job, self._job = self._job, None
            _lib.load().stx_host_pack_end(job)
            self._keep = None

    def __del__(self):
        try:
            self._finish()                  # joins the job: nothing writes into the staging buffer after this object is gone
        except Exception:
            pass

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


class _Stage:
    """One pinned staging buffer of the pool."""

    def __init__(self, n_float: int):
        self.pcm = torch.empty(max(n_float, 1), dtype=torch.float32, pin_memory=True)
        self.owner = None                   # weakref to the PackedClips that holds it (its death joins the packing job)
        self.event = None                   # CUDA event after the last copy out of the buffer that was enqueued

    def free(self) -> bool:
        if self.owner is not None and self.owner() is not None:
            return False
        return self.event is None or self.event.query()


class _StagePool:
    """Pinned staging buffers of one extractor.  A buffer is reused only when the PackedClips that held it is gone (which
    joins its packing job) and the H2D copies enqueued from it have completed; otherwise another buffer is allocated, so a
    caller may hold several PackedClips (or pack batch i + 1 while batch i is in flight) safely."""

    def __init__(self):
        self.stages = []
        self.lock = threading.Lock()

    def acquire(self, n_float: int) -> _Stage:
        with self.lock:
            best = None
            for st in self.stages:
                if st.free() and st.pcm.numel() >= n_float and (best is None or st.pcm.numel() < best.pcm.numel()):
                    best = st
            if best is None:
                # drop free buffers that are too small (grow-only in effect), then allocate
                self.stages = [st for st in self.stages if not (st.free() and st.pcm.numel() < n_float)]
                best = _Stage(n_float)
                self.stages.append(best)
            best.event = None
            best.owner = None
            return best


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
        self._stages = _StagePool()

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

    # -- chunked three-stream pipeline (host in -> host out) -----------------------------------
    CHUNK_BYTES = 12 << 20          # PCM bytes per chunk: large enough for PCIe efficiency, small enough to overlap

    def _streams(self):
        if getattr(self, "_pipe_streams", None) is None:
            dev = self._device()
            self._pipe_streams = tuple(torch.cuda.Stream(device=dev) for _ in range(3))
        return self._pipe_streams

    @classmethod
    def chunk_bounds(cls, offsets: np.ndarray, lengths: np.ndarray, chunk_bytes: int | None = None):
        """Clip ranges [b0, b1) whose packed PCM is about ``chunk_bytes`` each (at least one clip per chunk)."""
        chunk_bytes = cls.CHUNK_BYTES if chunk_bytes is None else chunk_bytes
        B = int(lengths.size)
        bounds, b0 = [], 0
        while b0 < B:
            b1 = b0 + 1
            while b1 < B and (int(offsets[b1]) + int(lengths[b1]) - int(offsets[b0])) * 4 <= chunk_bytes:
                b1 += 1
            bounds.append((b0, b1))
            b0 = b1
        return bounds

    def _pipeline(self, packed: "PackedClips", launch, outputs):
        """Runs ``launch(pcm_chunk_d, offsets_d, lengths_d, b0, b1, max_len)`` per chunk of clips.

        ``outputs`` = list of (device tensor [B, ...], pinned host tensor [B, ...]) pairs: rows b0:b1 are copied
        back as soon as the chunk's kernels finish.  Returns after the last D2H copy has completed.
        """
        dev = self._device()
        B = packed.batch_size
        s_in, s_run, s_out = self._streams()
        cur = torch.cuda.current_stream(dev)
        for s_ in (s_in, s_run, s_out):
            s_.wait_stream(cur)                                  # buffers allocated on the caller's stream are ready
        bounds = packed.bounds if packed.bounds is not None else self.chunk_bounds(packed.offsets, packed.lengths)
        # chunk-relative offsets: every chunk is an independent call into the library
        rel = packed.offsets.copy()
        for b0, b1 in bounds:
            rel[b0:b1] -= packed.offsets[b0]
        meta = torch.empty(2 * B, dtype=torch.int64, pin_memory=True)
        mv = meta.numpy()
        mv[:B] = rel
        mv[B:2 * B] = 0
        mv[B:2 * B].view(np.int32)[:B] = packed.lengths
        total = int(packed.offsets[-1]) + int(packed.lengths[-1]) if B else 0
        pcm_d = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
        meta_d = torch.empty(2 * B, dtype=torch.int64, device=dev)
        with torch.cuda.stream(s_in):
            meta_d.copy_(meta, non_blocking=True)
        off_d, len_d = meta_d[:B], meta_d[B:2 * B].view(torch.int32)[:B]
        for ci, (b0, b1) in enumerate(bounds):
            lo = int(packed.offsets[b0])
            hi = int(packed.offsets[b1 - 1]) + int(packed.lengths[b1 - 1])
            packed.wait(ci)                                      # this chunk's clips are in the pinned buffer
            with torch.cuda.stream(s_in):
                pcm_d[lo:hi].copy_(packed.pcm[lo:hi], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            if packed.stage is not None:
                packed.stage.event = ev_in                       # the staging buffer is busy until this copy has drained
            with torch.cuda.stream(s_run):
                s_run.wait_event(ev_in)
                launch(pcm_d[lo:hi], off_d[b0:b1], len_d[b0:b1], b0, b1, int(packed.lengths[b0:b1].max()))
                ev_run = torch.cuda.Event()
                ev_run.record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_run)
                for dev_t, host_t in outputs:
                    host_t[b0:b1].copy_(dev_t[b0:b1], non_blocking=True)
        s_out.synchronize()                                      # the caller holds the result on the host
        cur.wait_stream(s_run)                                   # device buffers may be reused by the caller's stream

    # -- host -> device ----------------------------------------------------------------------
    FIRST_CHUNK_BYTES = 3 << 20          # a small first chunk: the H2D pipeline starts after ~50 us of packing

    def pack(self, clips: Sequence[np.ndarray]) -> PackedClips:
        """Gather the clips (ordinary pageable arrays) into one pinned host buffer of the extractor's pool.

        The copy runs in the library (``stx_host_pack_begin``: a native job, non-temporal stores, all the cores this rank
        may use), one pipeline chunk after the other; the returned object waits per chunk, so the chunked pipeline starts
        the H2D copy of chunk 0 while later chunks are still being packed."""
        lengths = np.fromiter((c.size for c in clips), dtype=np.int32, count=len(clips))
        offsets, total = _layout(lengths)
        stage = self._stages.acquire(total)
        pcm = stage.pcm[:total]
        lib = _lib.load()
        if not len(clips):
            packed = PackedClips(pcm, offsets, lengths, [], None, stage)
            stage.owner = weakref.ref(packed)
            return packed
        bounds = self.chunk_bounds(offsets, lengths)
        if len(bounds) > 1 and self.FIRST_CHUNK_BYTES < self.CHUNK_BYTES:
            first = self.chunk_bounds(offsets[:bounds[0][1]], lengths[:bounds[0][1]], self.FIRST_CHUNK_BYTES)
            bounds = first + bounds[1:]
        src = np.fromiter((c.ctypes.data for c in clips), dtype=np.uint64, count=len(clips))
        nbytes = lengths.astype(np.int64) * 4
        dst_off = offsets * 4
        starts = np.array([b0 for b0, _ in bounds] + [len(clips)], dtype=np.int32)
        if total * 4 < self.PACK_INLINE_BYTES:
            _lib.check(lib.stx_host_pack(src.ctypes.data, nbytes.ctypes.data, pcm.data_ptr(), dst_off.ctypes.data, len(clips), 1),
                       "stx_host_pack")
            job, keep = None, None
        else:
            job = lib.stx_host_pack_begin(src.ctypes.data, nbytes.ctypes.data, pcm.data_ptr(), dst_off.ctypes.data, len(clips),
                                          starts.ctypes.data, len(bounds), pack_threads())
            if not job:
                _lib.check(-1, "stx_host_pack_begin")
            keep = (clips, src, nbytes, dst_off, starts)
        packed = PackedClips(pcm, offsets, lengths, bounds, job, stage, keep)
        stage.owner = weakref.ref(packed)
        return packed

    PACK_INLINE_BYTES = 1 << 20

    def to_device(self, packed: PackedClips):
        """(pcm, offsets, lengths) on the device; one async copy for the PCM, one for the metadata."""
        dev = self._device()
        B = packed.batch_size
        packed.wait()
        if packed.pcm.is_cuda:
            pcm_d = packed.pcm
        else:
            pcm_d = torch.empty(packed.pcm.numel(), dtype=torch.float32, device=dev)
            pcm_d.copy_(packed.pcm, non_blocking=True)
        meta = torch.empty(2 * B, dtype=torch.int64, pin_memory=True)     # (torch caches pinned blocks; freed in stream order)
        mv = meta.numpy()
        mv[:B] = packed.offsets
        mv[B:2 * B] = 0
        mv[B:2 * B].view(np.int32)[:B] = packed.lengths       # lengths live in the low half of the second block
        meta_d = torch.empty(2 * B, dtype=torch.int64, device=dev)
        meta_d.copy_(meta, non_blocking=True)
        if packed.stage is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            packed.stage.event = ev
        offsets_d = meta_d[:B]
        lengths_d = meta_d[B:2 * B].view(torch.int32)[:B]
        return pcm_d, offsets_d, lengths_d

    @staticmethod
    def _check_return(return_tensors, output):
        if return_tensors not in ("pt", "torch", None, "np", "numpy"):
            raise ValueError(f"return_tensors={return_tensors!r} is not supported (use 'pt' or 'np')")
        if output not in (None, "device", "host"):
            raise ValueError("output must be 'device' or 'host'")
        to_numpy = return_tensors in (None, "np", "numpy")
        on_host = to_numpy or output == "host"
        return to_numpy, on_host

    @staticmethod
    def _finish(data: dict, to_numpy: bool):
        if to_numpy:
            return BatchFeature({k: (v.numpy() if not v.is_cuda else v.cpu().numpy()) for k, v in data.items()})
        return BatchFeature(data)


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
                 do_normalize_per_mel_bins=True, output=None, **kwargs):
        """``output="host"`` (implied by ``return_tensors="np"``) returns pinned CPU tensors through the chunked
        H2D | kernels | D2H pipeline; the default for ``"pt"`` keeps the tensors on the CUDA device."""
        self._check_rate(sampling_rate)
        dev = self._device()                 # fail loudly before any host work when there is no GPU
        to_numpy, on_host = self._check_return(return_tensors, output)
        return_attention_mask = self.return_attention_mask if return_attention_mask is None else return_attention_mask
        if isinstance(raw_speech, PackedClips):
            packed = raw_speech
        else:
            packed = self.pack(_as_clip_list(raw_speech, 3, self.__class__.__name__))
        frames = np.array([ops.k_num_frames(int(n)) for n in packed.lengths], dtype=np.int64)
        T_pad, _ = self._padded_frames(frames, padding, max_length, truncation, pad_to_multiple_of)
        B = packed.batch_size
        if on_host and not packed.pcm.is_cuda and B > 0:
            feats = torch.empty((B, T_pad // 2, 160), dtype=torch.float32, device=dev)
            mask = torch.empty((B, T_pad // 2), dtype=torch.int32, device=dev) if return_attention_mask else None
            h_feats = torch.empty((B, T_pad // 2, 160), dtype=torch.float32, pin_memory=True)
            h_mask = torch.empty((B, T_pad // 2), dtype=torch.int32, pin_memory=True) if return_attention_mask else None

            def launch(pcm_c, off_c, len_c, b0, b1, max_len):
                lens_c = packed.lengths[b0:b1]
                ops.fbank_k(pcm_c, off_c, len_c, max_len, T_pad, self.padding_value, bool(do_normalize_per_mel_bins),
                            want_mask=bool(return_attention_mask), out=feats[b0:b1],
                            mask=mask[b0:b1] if mask is not None else None,
                            uniform=bool(lens_c.size and lens_c.min() == lens_c.max()))

            self._pipeline(packed, launch, [(feats, h_feats)] + ([(mask, h_mask)] if mask is not None else []))
            data = {"input_features": h_feats}
            if return_attention_mask:
                data["attention_mask"] = h_mask
            return self._finish(data, to_numpy)
        pcm_d, off_d, len_d = self.to_device(packed)
        feats, mask = ops.fbank_k(pcm_d, off_d, len_d, packed.max_length, T_pad, self.padding_value,
                                  bool(do_normalize_per_mel_bins), want_mask=bool(return_attention_mask),
                                  uniform=bool(B and packed.lengths.min() == packed.lengths.max()))
        data = {"input_features": feats}
        if return_attention_mask:
            data["attention_mask"] = mask
        if on_host:
            data = {k: v.cpu() for k, v in data.items()}
        return self._finish(data, to_numpy)


    def collate(self, speech_arrays, output=None):
        """The audio half of the trainer's batch in one call.

        Replaces one ``feature_extractor(speech_array, sampling_rate, return_tensors="pt")`` call per item in
        ``CommonVoiceDataset.__getitem__`` (R/training/trainer_unfreeze.py:855-866; no peak-normalise, no trim)
        plus the audio padding of ``custom_collate_fn`` (R/training/trainer_unfreeze.py:898-908): returns
        ``{"input_values": float32 [B, max T', 160] zero-padded, "attention_mask_audio": int64 [B, max T']}``,
        every stacked frame a per-clip call returns marked valid (the collate discards the extractor's own mask).
        ``output="host"`` gives pinned CPU tensors through the chunked pipeline (what a DataLoader would hand to
        ``.to(device, non_blocking=True)``, R/training/trainer_unfreeze.py:1059-1061); the default keeps them on the GPU.
        """
        dev = self._device()
        if output not in (None, "device", "host"):
            raise ValueError("output must be 'device' or 'host'")
        packed = speech_arrays if isinstance(speech_arrays, PackedClips) else self.pack(
            _as_clip_list(list(speech_arrays), 3, self.__class__.__name__))
        B = packed.batch_size
        frames = np.array([ops.k_num_frames(int(n)) for n in packed.lengths], dtype=np.int64)
        T_pad = 2 * int(((frames + 1) // 2).max()) if B else 0
        if output == "host" and not packed.pcm.is_cuda and B > 0:
            feats = torch.empty((B, T_pad // 2, 160), dtype=torch.float32, device=dev)
            mask = torch.empty((B, T_pad // 2), dtype=torch.int64, device=dev)
            h_feats = torch.empty((B, T_pad // 2, 160), dtype=torch.float32, pin_memory=True)
            h_mask = torch.empty((B, T_pad // 2), dtype=torch.int64, pin_memory=True)

            def launch(pcm_c, off_c, len_c, b0, b1, max_len):
                ops.fbank_k_collate(pcm_c, off_c, len_c, max_len, T_pad, self.padding_value, out=feats[b0:b1], mask=mask[b0:b1])

            self._pipeline(packed, launch, [(feats, h_feats), (mask, h_mask)])
            return {"input_values": h_feats, "attention_mask_audio": h_mask}
        pcm_d, off_d, len_d = self.to_device(packed)
        feats, mask = ops.fbank_k_collate(pcm_d, off_d, len_d, packed.max_length, T_pad, self.padding_value)
        if output == "host":
            feats, mask = feats.cpu(), mask.cpu()
        return {"input_values": feats, "attention_mask_audio": mask}


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
                 do_normalize=None, device="cpu", output=None, **kwargs):
        self._check_rate(sampling_rate)
        dev = self._device()
        to_numpy, on_host = self._check_return(return_tensors, output)
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
        want_mask = bool(return_attention_mask if return_attention_mask is not None else self.return_attention_mask)
        B, T = packed.batch_size, n_samples // self.hop_length
        if on_host and not packed.pcm.is_cuda and B > 0:
            feats = torch.empty((B, 80, T), dtype=torch.float32, device=dev)
            mask = torch.empty((B, T), dtype=torch.int32, device=dev) if want_mask else None
            h_feats = torch.empty((B, 80, T), dtype=torch.float32, pin_memory=True)
            h_mask = torch.empty((B, T), dtype=torch.int32, pin_memory=True) if want_mask else None

            def launch(pcm_c, off_c, len_c, b0, b1, max_len):
                ops.logmel_w(pcm_c, off_c, len_c, n_samples, want_mask=want_mask, out=feats[b0:b1],
                             mask=mask[b0:b1] if mask is not None else None)

            self._pipeline(packed, launch, [(feats, h_feats)] + ([(mask, h_mask)] if mask is not None else []))
            data = {"input_features": h_feats}
            if want_mask:
                data["attention_mask"] = h_mask
            return self._finish(data, to_numpy)
        pcm_d, off_d, len_d = self.to_device(packed)
        feats, mask = ops.logmel_w(pcm_d, off_d, len_d, n_samples, want_mask=want_mask)
        data = {"input_features": feats}
        if want_mask:
            data["attention_mask"] = mask
        if on_host:
            data = {k: v.cpu() for k, v in data.items()}
        return self._finish(data, to_numpy)
