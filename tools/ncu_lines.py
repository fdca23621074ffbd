#!/usr/bin/env python
"""Per-source-line instruction and stall shares of an .ncu-rep, normalised per unit of work.
usage: ncu_lines.py report.ncu-rep units [top]   (units = e.g. clips * frames * warps)"""
import csv, io, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, agg = None, []
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 10 and r[0].isdigit():
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:105], int(r[6]), int(r[7])))
        except ValueError:
            pass
ti, ts = sum(a[4] for a in agg), sum(a[3] for a in agg) or 1
print(f"total warp instructions {ti}  = {ti / units:.1f} per unit; stall samples {ts}")
for a in sorted(agg, key=lambda a: -a[4])[:top]:
    print(f"{a[0][:12]:12s}:{a[1]:4d} inst {100 * a[4] / ti:5.1f}% ({a[4] / units:6.1f}/unit) samp {100 * a[3] / ts:5.1f}%  {a[2]}")
