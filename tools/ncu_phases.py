"""Per-PHASE summary of an ncu report's SASS page: the instruction stream is cut at every BAR.SYNC / BAR.ARV, and for each
segment the executed warp instructions, stall samples, top stall reasons, shared-memory wavefronts and the instruction mix
are printed.  (Segments that never executed at the hot trip count are dropped.)

    python tools/ncu_phases.py report.ncu-rep [min_inst_executed]
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
min_exec = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


segs = []
cur = None


def new():
    return {"n": 0, "exec": 0, "samples": 0, "wf": 0, "stalls": collections.Counter(), "mix": collections.Counter(), "first": None}


cur = new()
for r in rows:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    ins = r[ix["Source"]].strip()
    op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
    op0 = op.split(".")[0]
    e = num(r[ix["Instructions Executed"]])
    cur["n"] += 1
    cur["exec"] += e
    cur["samples"] += num(r[ix["# Samples"]])
    cur["wf"] += num(r[ix["L1 Wavefronts Shared"]])
    cur["mix"][op0] += e
    for s in stall_cols:
        cur["stalls"][s[6:]] += num(r[ix[s]])
    if cur["first"] is None:
        cur["first"] = r[0]
    if op0 == "BAR":
        segs.append(cur)
        cur = new()
segs.append(cur)
tot = sum(s["samples"] for s in segs)
print(f"total samples {tot}")
for k, s in enumerate(segs):
    if s["exec"] < min_exec:
        continue
    st = " ".join(f"{n}={v}" for n, v in s["stalls"].most_common(5) if v)
    mix = " ".join(f"{n}:{v}" for n, v in s["mix"].most_common(8))
    print(f"seg {k:2d} static={s['n']:4d} exec={s['exec']:9d} samples={s['samples']:5d} ({100 * s['samples'] / tot:4.1f}%) smem_wf={s['wf']:9d} "
          f"samples/kinst={1e3 * s['samples'] / max(1, s['exec']):.2f}\n        stalls: {st}\n        mix: {mix}")
