"""CPU: properties of the generated SASS that the measured performance depends on (cuobjdump on the in-tree objects).

* the shipped recipe-K kernel reads its per-warp window / twiddle / DC tables through the UNIFORM datapath (LDCU).  ptxas
  decides that by register pressure: a few more live values at the 128-register cap and it falls back to register-indexed
  LDCs through the MIO, which costs 13-40 % of the kernel with no change in results -- a silent regression
  that only this test (or a benchmark) would catch;
* the TMEM stash (tcgen05.st / tcgen05.ld), the TMA bulk copies and the tcgen05 MMA of the cosine kernel are really there.
"""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
OBJ = ROOT / "speech_transcript_embeddings_b200" / "csrc" / "_obj"


def _sass(obj):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(exe).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", str(OBJ / obj)], capture_output=True, text=True, check=True).stdout
    funcs = {}
    for chunk in re.split(r"\n\s*Function : ", out)[1:]:
        name, _, body = chunk.partition("\n")
        funcs[name.strip()] = body
    return funcs


def _count(body, mnemonic):
    return len(re.findall(r"\s" + re.escape(mnemonic) + r"[.\s]", body))


def test_recipe_k_kernel_uses_uniform_constant_loads_tmem_and_tma(lib):
    funcs = _sass("fbank_k.o")
    duo = {n: b for n, b in funcs.items() if "k_frames_duo" in n}
    assert len(duo) == 2, list(funcs)                       # <false> and <true> (peak-normalise fused)
    for name, body in duo.items():
        ldcu, ldc = _count(body, "LDCU"), _count(body, "LDC")
        assert ldcu >= 150 and ldc <= 40, f"{name}: {ldcu} LDCU vs {ldc} LDC -- the per-warp tables left the uniform datapath"
        assert _count(body, "STTM") >= 2 and _count(body, "LDTM") >= 2, "TMEM stash (tcgen05.st / tcgen05.ld) missing"
        assert _count(body, "UBLKCP") >= 1, "TMA bulk copy (cp.async.bulk) missing"
        assert _count(body, "DFMA") >= 300, "the FFT runs on the FP64 pipe"
        assert "STL" not in body and "LDL" not in body, "register spills in the hot kernel"


def test_recipe_w_and_cosine_kernels(lib):
    w = [b for n, b in _sass("logmel_w.o").items() if "w_frames" in n]
    assert len(w) == 2                                       # the stock 30 s chunk (n_samples at compile time) and the general form
    for body in w:
        assert _count(body, "UBLKCP") >= 1 and _count(body, "FFMA") >= 250
        assert "STL" not in body and "LDL" not in body
    c = [b for n, b in _sass("cosine.o").items() if "c_nxm_tc" in n]
    assert c, "tcgen05 cosine kernel missing"
    for body in c:
        assert _count(body, "UTCHMMA") >= 1 and _count(body, "UTMALDG") >= 1 and _count(body, "LDTM") >= 1
