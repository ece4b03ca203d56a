"""ctypes binding of libstx_b200.so (the C ABI declared in include/stx_b200.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

# STX_B200_LIB: development override, an alternative BUILD OF THE SAME LIBRARY (A/B timing of kernel variants,
# tools/time_kernels.py).  It is still the CUDA library or nothing: a missing file raises like the default path does.
LIB_PATH = Path(os.environ.get("STX_B200_LIB") or Path(__file__).resolve().parent / "libstx_b200.so")

# symbol -> (restype, argtypes); lists every function include/stx_b200.h declares
_SIGNATURES = {
    "stx_abi_version": (C.c_int, []),
    "stx_last_error": (C.c_char_p, []),
    "stx_kernel_launch_count": (C.c_uint64, []),
    "stx_profile_enable": (C.c_int, [C.c_int]),
    "stx_profile_collect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "stx_get_table": (C.c_int64, [C.c_char_p, C.c_void_p, C.c_int64]),
    "stx_host_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "stx_host_pack_begin": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "stx_host_pack_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "stx_host_pack_end": (C.c_int, [C.c_void_p]),
    "stx_fbank_k_workspace": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "stx_fbank_k": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                              C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "stx_fbank_k_collate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "stx_fbank_k_projection_workspace": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "stx_fbank_k_projection": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float,
                                         C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "stx_peak_abs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "stx_logmel_w_workspace": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "stx_logmel_w": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "stx_cosine_workspace": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "stx_cosine_pairwise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "stx_cosine_nxm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "stx_cosine_topk_workspace": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "stx_cosine_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "stx_score_pos_neg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "stx_feature_projection_workspace": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "stx_feature_projection": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "stx_cosine_gather_sizes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t),
                                          C.POINTER(C.c_size_t)]),
    "stx_resample_plan": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "stx_resample_filter": (C.c_longlong, [C.c_int, C.c_int, C.c_void_p, C.c_longlong]),
    "stx_resample_poly": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "stx_cosine_nxm_gathered": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                          C.c_size_t, C.c_void_p]),
}

_lock = threading.Lock()
_lib = None


class StxError(RuntimeError):
    """A libstx_b200 entry point returned non-zero, or the library is unusable."""


def exported_symbols():
    return tuple(_SIGNATURES)


def load() -> C.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise StxError(
                    f"{LIB_PATH} is missing: build it with `python -m speech_transcript_embeddings_b200.build` "
                    "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
            lib = C.CDLL(str(LIB_PATH))
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)     # AttributeError if the .so lacks a declared symbol
                fn.restype = res
                fn.argtypes = args
            if lib.stx_abi_version() != 1:
                raise StxError(f"libstx_b200 ABI {lib.stx_abi_version()} != 1: rebuild the library")
            _lib = lib
        return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().stx_last_error().decode(errors="replace")
        raise StxError(f"{what} failed (rc={rc}): {msg}")


def profile(on: bool) -> None:
    load().stx_profile_enable(int(bool(on)))


def profile_collect(cap: int = 65536):
    """[(kernel name, milliseconds), ...] of the launches since profiling was enabled (synchronises)."""
    lib = load()
    names = C.create_string_buffer(32 * cap)
    ms = (C.c_float * cap)()
    n = lib.stx_profile_collect(names, ms, cap)
    return [(names.raw[32 * i:32 * (i + 1)].split(b"\0", 1)[0].decode(), float(ms[i])) for i in range(n)]


def launch_count() -> int:
    return int(load().stx_kernel_launch_count())
