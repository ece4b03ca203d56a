"""Per-source-line summary of an ncu report's source page (stall samples, instructions, smem conflicts).

    python tools/ncu_lines.py report.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
col = {n: i for i, n in enumerate(hdr)}
def c(name):
    return [i for i, n in enumerate(hdr) if n == name][0]


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0
i_samp, i_inst = c("# Samples"), c("Instructions Executed")
i_conf, i_wf = c("L1 Wavefronts Shared Excessive"), c("L1 Wavefronts Shared")
stall_cols = [(n, i) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
lines = []
total = 0
for r in rows:
    if len(r) < len(hdr) or r[0] in ("Line No", "") or not r[0].isdigit():
        continue
    s = num(r[i_samp])
    total += s
    stalls = sorted(((num(r[i]), n) for n, i in stall_cols), reverse=True)[:3]
    lines.append((s, int(r[0]), num(r[i_inst]), num(r[i_conf]), num(r[i_wf]), stalls, r[1].strip()[:90]))
lines.sort(reverse=True)
print(f"total samples {total}")
for s, ln, inst, conf, wf, stalls, src in lines[:top]:
    st = " ".join(f"{n[6:]}={v}" for v, n in stalls if v)
    print(f"{100 * s / total:5.1f}% L{ln:<4d} inst={inst:<9d} smem_wf={wf:<9d} excess={conf:<9d} [{st}] {src}")
