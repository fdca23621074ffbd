#!/usr/bin/env python
"""Turn an .ncu-rep (one kernel launch) into the small JSON bench.py reads (profiles/r2_*.json): per-launch DRAM bytes,
duration, issue utilisation, stall breakdown -- stamped with the hash of the kernel sources the capture was taken from, so
that bench.py refuses it once those sources change.
usage: capture_json.py report.ncu-rep out.json "workload text" source1.cu [source2.cuh ...]"""
import csv, hashlib, io, json, os, subprocess, sys

rep, out, workload, sources = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]


def get(name, scale=1.0):
    if name not in hdr:
        return None
    i = hdr.index(name)
    try:
        v = float(vals[i].replace(",", ""))
    except ValueError:
        return None
    u = units[i].lower()
    mult = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return v * mult * scale


h = hashlib.sha256()
for n in sources:
    with open(os.path.join(ROOT, "spectrogram-midi_b200", "csrc", n), "rb") as f:
        h.update(f.read())
rec = {
    "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None,
    "workload": workload,
    "source": f"ncu --set full --clock-control none ({os.path.basename(rep)})",
    "source_files": sources, "source_sha16": h.hexdigest()[:16],
    "duration_ms": get("gpu__time_duration.sum", 1e3),
    "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
    "smsp_issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "sm_warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "inst_executed": get("smsp__inst_executed.sum"),
    "registers_per_thread": get("launch__registers_per_thread"), "grid": get("launch__grid_size"), "block": get("launch__block_size"),
    "shared_wavefronts": get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "shared_bank_conflicts": get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "stalls_per_issue": {hh.split("stalled_")[1].split("_per_issue")[0]: float(vals[i]) for i, hh in enumerate(hdr)
                         if hh.startswith("smsp__average_warps_issue_stalled") and hh.endswith("_per_issue_active.ratio")
                         and vals[i].replace(".", "").isdigit() and float(vals[i]) > 0.05},
}
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps(rec, indent=1))
