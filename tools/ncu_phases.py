#!/usr/bin/env python
"""Per-phase view of a kernel from an .ncu-rep SASS page: segments between synchronisation instructions
(BAR / WARPSYNC / branch targets are not tracked), with executed instructions, stall samples and the dominant stall reasons."""
import csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
segs, cur = [], None
def new(name):
    global cur
    cur = {"name": name, "n": 0, "inst": 0, "samp": 0, "st": {c: 0 for c in stall_cols}, "ops": {}, "wf": 0, "wf_ideal": 0}
    segs.append(cur)
new("start")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    sass = r[ix["Source"]].strip()
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    cur["n"] += 1
    cur["inst"] += int(r[ix["Instructions Executed"]])
    cur["samp"] += int(r[ix["# Samples"]])
    cur["wf"] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    cur["wf_ideal"] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    for c in stall_cols:
        cur["st"][c] += int(r[ix[c]] or 0)
    base = op.split(".")[0]
    cur["ops"][base] = cur["ops"].get(base, 0) + int(r[ix["Instructions Executed"]])
    if base in ("BAR", "WARPSYNC"):
        new(sass[:40])
tot_s = sum(s["samp"] for s in segs) or 1
tot_i = sum(s["inst"] for s in segs) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
for s in segs:
    if s["samp"] < 0.003 * tot_s and s["inst"] < 0.003 * tot_i:
        continue
    top = sorted(s["st"].items(), key=lambda kv: -kv[1])[:4]
    ops = sorted(s["ops"].items(), key=lambda kv: -kv[1])[:6]
    print(f"after [{s['name']:<40s}] static {s['n']:5d} inst {100*s['inst']/tot_i:5.1f}% samp {100*s['samp']/tot_s:5.1f}% wf {s['wf']/1e6:7.1f}M (ideal {s['wf_ideal']/1e6:7.1f}M) | "
          + " ".join(f"{k[6:]}={100*v/max(1,s['samp']):.0f}%" for k, v in top) + " | " + " ".join(f"{k}:{100*v/max(1,s['inst']):.0f}%" for k, v in ops))
