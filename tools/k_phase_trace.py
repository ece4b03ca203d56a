"""Where a group of k_frames_duo spends its time: cycles between its group barriers, per phase, from a trace build of the library.

    python tools/ab_build.py trace=-DSTX_K_TRACE=1 trace2=-DSTX_K_TRACE=2
    STX_B200_LIB=build/variants/lib_trace.so  python tools/k_phase_trace.py [clips] [seconds] [iters] 1
    STX_B200_LIB=build/variants/lib_trace2.so python tools/k_phase_trace.py [clips] [seconds] [iters] 2

Thread 0 of every group adds clock64() differences to per-phase counters after each of its group barriers, so a phase's figure
includes the wait for the slowest warp of the group: it is the group's critical path, not pipe-busy time.  Mode 2: every warp adds its BUSY cycles per phase (release by the previous
barrier -> arrival at the next), so phase time minus busy time is barrier wait, and the spread over the warps shows imbalance.
"""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speech_transcript_embeddings_b200 import _lib, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = torch.device("cuda", 0)
n = int(secs * 16000)
g = torch.Generator(device=dev).manual_seed(0)
pcm = 0.1 * torch.randn(B * n, generator=g, device=dev)
off = torch.arange(B, device=dev, dtype=torch.int64) * n
ln = torch.full((B,), n, dtype=torch.int32, device=dev)
T = ops.k_num_frames(n)
T_pad = 2 * ((T + 1) // 2)
out = torch.empty((B, T_pad // 2, 160), dtype=torch.float32, device=dev)
lib = _lib.load()
fn = lib.stx_debug_ktrace
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]
buf = (C.c_uint64 * 64)()
for _ in range(3):
    ops.fbank_k(pcm, off, ln, n, T_pad, out=out, uniform=True)
fn(buf)                                                        # clear
for _ in range(iters):
    ops.fbank_k(pcm, off, ln, n, T_pad, out=out, uniform=True)
fn(buf)
tiles = B * ((T + 31) // 32) * iters                           # (exact for chunk sizes that are multiples of 32)
# (a trace point between the store and the conversion makes ptxas drop the uniform constant loads, so the two are one phase)
names = ["pass1 x2 + stash", "c + H1 loads", "stash -> H2", "pass 2 (both halves) + DC + power", "mel + ln",
         "(unused)", "store + statistics + convert next"]
if mode == 2:
    res = {names[i]: {"busy_mean": round(sum(buf[8 * i + w] for w in range(8)) / 8 / tiles),
                      "busy_per_warp": [round(buf[8 * i + w] / tiles) for w in range(8)]} for i in range(7) if i != 5}
    res["sum_busy_mean"] = sum(v["busy_mean"] for v in res.values())
else:
    tot = sum(buf[i] for i in range(7))
    res = {names[i]: round(buf[i] / tiles) for i in range(7)}
    res["sum_cycles_per_group_tile"] = round(tot / tiles)
print(json.dumps({"lib": os.environ.get("STX_B200_LIB", "in-tree"), "clips": B, "seconds": secs, "cycles_per_group_tile": res}))
