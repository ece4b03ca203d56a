"""How should pageable user arrays reach the device?  (a) gather into a pinned staging buffer (stx_host_pack), then H2D;
(b) page-lock the user's arrays in place (cudaHostRegister), H2D straight from them, unregister; (c) plain pageable copies.
Times all three on the cfg2 batch (64 pageable float32 arrays of 30 s) -- wall clock, one rank.

    python tools/bench_hostreg.py
"""
import json
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200 import _lib  # noqa: E402
from speech_transcript_embeddings_b200.feature_extraction import pack_threads  # noqa: E402

dev = torch.device("cuda", 0)
rt = torch.cuda.cudart()
B, n = 64, 480000
rng = np.random.default_rng(0)
clips = [rng.standard_normal(n).astype(np.float32) for _ in range(B)]
dst = torch.empty(B * n, dtype=torch.float32, device=dev)
res = {"clips": B, "bytes": B * n * 4, "pack_threads": pack_threads()}


def timed(f, reps=5):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


# (a) staging: native pack (all at once), then one H2D
pinned = torch.empty(B * n, dtype=torch.float32, pin_memory=True)
lib = _lib.load()
src = np.array([c.ctypes.data for c in clips], np.uint64)
nb = np.full(B, n * 4, np.int64)
off = np.arange(B, dtype=np.int64) * n * 4
for th in (1, 4, 8, 16):
    res[f"pack_only_ms_{th}_threads"] = timed(
        lambda: lib.stx_host_pack(src.ctypes.data, nb.ctypes.data, pinned.data_ptr(), off.ctypes.data, B, th))
res["h2d_only_ms"] = timed(lambda: dst.copy_(pinned, non_blocking=True))


# (b) register in place
def reg_all():
    for c in clips:
        rc = rt.cudaHostRegister(c.ctypes.data, c.nbytes, 0)
        assert int(rc) == 0, rc


def unreg_all():
    for c in clips:
        rt.cudaHostUnregister(c.ctypes.data)


for tag in ("", "_2nd"):
    t0 = time.perf_counter()
    reg_all()
    res["register_ms_serial" + tag] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    unreg_all()
    res["unregister_ms_serial" + tag] = (time.perf_counter() - t0) * 1e3
ex = ThreadPoolExecutor(8)
t0 = time.perf_counter()
list(ex.map(lambda c: rt.cudaHostRegister(c.ctypes.data, c.nbytes, 0), clips))
res["register_ms_8_threads"] = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
list(ex.map(lambda c: rt.cudaHostUnregister(c.ctypes.data), clips))
res["unregister_ms_8_threads"] = (time.perf_counter() - t0) * 1e3


def h2d_registered():
    reg_all()
    for i, c in enumerate(clips):
        dst[i * n:(i + 1) * n].copy_(torch.from_numpy(c), non_blocking=True)
    torch.cuda.synchronize()
    unreg_all()


res["register_h2d_unregister_ms"] = timed(h2d_registered, 3)


# (c) plain pageable copies (what .to(device) of a pageable tensor does)
def h2d_pageable():
    for i, c in enumerate(clips):
        dst[i * n:(i + 1) * n].copy_(torch.from_numpy(c), non_blocking=True)


res["pageable_h2d_ms"] = timed(h2d_pageable, 3)
print(json.dumps(res))
