"""Summarise one .ncu-rep (first kernel) into the handful of metrics DESIGN.md / bench.py quote.
usage: python tools/ncu_summary.py <report.ncu-rep> "<header comment>" > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg',
        'sm__cycles_elapsed.avg.per_second', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__cluster_dim_x',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
print("# " + sys.argv[2])
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} = {vals[i]} {units[i]}")
