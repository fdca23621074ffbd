#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + per-source-line instruction / stall shares."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic"]
print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for w in want:
    if w in hdr:
        print(f"{w} = {vals[hdr.index(w)]}")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        try:
            v = float(vals[i])
        except ValueError:
            continue
        if v > 0.1:
            print(f"stall {h.split('stalled_')[1].split('_per_issue')[0]:24s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, agg = None, []
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 10 and r[0].isdigit():
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:100], int(r[6]), int(r[7])))
        except ValueError:
            pass
ts, ti = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
print(f"== source lines by stall samples (total samples {ts}, warp instructions {ti})")
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{a[0]:16s}:{a[1]:4d} samp {100 * a[3] / ts:5.1f}% inst {100 * a[4] / ti:5.1f}%  {a[2]}")
