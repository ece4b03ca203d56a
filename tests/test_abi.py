"""CPU: the C-ABI library loads, exports every symbol include/stx_b200.h declares, and rejects bad
arguments with error codes (no compute calls: there is no GPU here)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "stx_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from speech_transcript_embeddings_b200 import _lib
    declared = _declared()
    assert len(declared) >= 12
    assert sorted(_lib.exported_symbols()) == declared
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_error_channel(lib):
    assert lib.stx_abi_version() == 1
    n = C.c_size_t(0)
    assert lib.stx_fbank_k_workspace(4, 480000, C.byref(n)) == 0 and n.value > 0
    assert lib.stx_fbank_k_workspace(-1, 0, C.byref(n)) == -1
    assert b"bad argument" in lib.stx_last_error()
    # odd T_pad and null pointers are argument errors, reported before any device work
    assert lib.stx_fbank_k(None, None, None, 1, 16000, None, 3, 0.0, 1, None, None, None, 0, None) == -1
    assert lib.stx_fbank_k(None, None, None, 1, 16000, None, 4, 0.0, 1, None, None, None, 0, None) == -1
    assert b"null pointer" in lib.stx_last_error()
    assert lib.stx_logmel_w(None, None, None, 1, 1000, None, None, None, None, 0, None) == -1
    assert lib.stx_cosine_nxm(None, None, 4, 4, 0, 1, None, None, 0, None) == -1
    assert lib.stx_get_table(b"nope", None, 0) == -1


def test_no_silent_cpu_path():
    import numpy as np
    import torch
    from speech_transcript_embeddings_b200 import _lib, ops
    from speech_transcript_embeddings_b200.feature_extraction import B200SeamlessM4TFeatureExtractor
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    x = torch.zeros(1000)
    with pytest.raises(_lib.StxError):
        ops.cosine_pairwise(x.view(1, -1), x.view(1, -1))
    with pytest.raises(_lib.StxError):
        B200SeamlessM4TFeatureExtractor()(np.zeros(1000, np.float32), sampling_rate=16000)


def test_argument_errors_of_the_newer_entry_points(lib):
    """Argument checks run before any device work, so they can be exercised without a GPU."""
    n = C.c_size_t(0)
    assert lib.stx_cosine_topk(None, None, 4, 4, 8, 1, 0, None, None, None, 0, None) == -1          # k outside 1..8
    assert b"outside 1..8" in lib.stx_last_error()
    assert lib.stx_cosine_topk(None, None, 4, 4, 8, 1, 9, None, None, None, 0, None) == -1
    assert lib.stx_cosine_topk(None, None, 4, 4, 8, 1, 3, None, None, None, 0, None) == -1          # null pointers
    assert lib.stx_cosine_topk_workspace(4096, 4096, 768, C.byref(n)) == 0 and n.value > 4096 * 32 * 8 * 8
    assert lib.stx_resample_poly(None, None, None, 1, 16000, 16000, None, None, None, 0, None, None) == -1
    assert lib.stx_resample_poly(None, None, None, 1, 48000, 16000, None, None, None, 0, None, None) == -1
    assert b"null pointer" in lib.stx_last_error()
    assert lib.stx_resample_plan(0, 16000, None, None, None, None, None) == -1
    up, down, taps = C.c_int(0), C.c_int(0), C.c_int(0)
    assert lib.stx_resample_plan(44100, 16000, C.byref(up), C.byref(down), C.byref(taps), None, None) == 0
    assert (up.value, down.value, taps.value) == (160, 441, 58)
    assert lib.stx_fbank_k_projection_workspace(4, 480000, 3, 1024, 0, C.byref(n)) == -1            # odd T_pad
    assert lib.stx_fbank_k_projection_workspace(4, 480000, 3000, 1024, 0, C.byref(n)) == 0 and n.value > 4 * 1500 * 160 * 4 * 3
    assert lib.stx_fbank_k_projection(None, None, None, 4, 480000, None, 3000, 0.0, None, None, 1e-5, None, None, 1024,
                                      None, None, None, None, 0, None) == -1
    assert b"null pointer" in lib.stx_last_error()
    # the "uniform batch" promise is a negated max_length: sizes are those of |max_length|
    a, b = C.c_size_t(0), C.c_size_t(0)
    assert lib.stx_fbank_k_workspace(8, 480000, C.byref(a)) == 0 and lib.stx_fbank_k_workspace(8, -480000, C.byref(b)) == 0
    assert a.value == b.value
