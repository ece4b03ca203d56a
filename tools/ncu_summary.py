"""Key metrics of an ncu report (raw page) as a markdown table.

    python tools/ncu_summary.py report.ncu-rep
"""
import csv
import subprocess
import sys

WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum launch__registers_per_thread launch__grid_size
launch__block_size launch__waves_per_multiprocessor sm__cycles_active.avg sm__warps_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active
sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum sm__throughput.avg.pct_of_peak_sustained_elapsed
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed dram__bytes_read.sum.per_second dram__bytes_write.sum.per_second
lts__t_sector_hit_rate.pct""".split()

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"### {name}\n\n| metric | value | unit |\n|---|---|---|")
    for i, h in enumerate(hdr):
        if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            try:
                if "issue_stalled" in h and float(vals[i]) < 0.2:
                    continue
            except ValueError:
                pass
            print(f"| {h} | {vals[i]} | {units[i]} |")
    print()
