#!/usr/bin/env python
"""profiles/ncu_sweep_512.json from an `ncu --set full` capture of the three sweeps of one RK stage (what bench.py's roofline record
quotes: DRAM bytes, FP64-pipe activity, FP64 thread-instructions per launch).

usage: python tools/ncu_to_json.py gpurun_out/r02l_sweep512.ncu-rep "<source note>" > profiles/ncu_sweep_512.json
The capture holds the x-, y- and last-direction launch in this order; the last one is `sweep_fused` when its kernel is the RKF
instantiation of k_sweep_tma (7th template argument true), else `sweep_z`."""
import csv
import io
import json
import re
import subprocess
import sys

rep, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def num(r, key):
    return float(r[col[key]].replace(",", ""))


launches = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    m = re.search(r"k_sweep_tma<([^>]*)>", name)
    args = [a.strip() for a in m.group(1).split(",")] if m else []
    launches.append({"name": name, "xs": len(args) > 2 and args[2] in ("1", "true", "(bool)1"),
                     "rkf": len(args) > 6 and args[6] in ("1", "true", "(bool)1"),
                     "dram_bytes_read": int(num(r, "dram__bytes_read.sum") * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[rows[1][col["dram__bytes_read.sum"]]]),
                     "dram_bytes_write": int(num(r, "dram__bytes_write.sum") * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[rows[1][col["dram__bytes_write.sum"]]]),
                     "fp64_pipe_active_pct": num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                     "time_ms_under_ncu": num(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[rows[1][col["gpu__time_duration.sum"]]]})
# FP64 thread-instructions per launch from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, check=True).stdout
counts, cur, seen = [], None, None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = 0; seen = set(); counts.append(0); continue
    if r and r[0] == "Address":
        h2 = r; iA, iS, iT = h2.index("Address"), h2.index("Source"), h2.index("Thread Instructions Executed"); continue
    if not counts or len(r) < 10 or r[iA] in seen:
        continue
    seen.add(r[iA])
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
    if m and m.group(2).split(".")[0] in ("DFMA", "DMUL", "DADD"):
        try:
            counts[-1] += int(r[iT])
        except ValueError:
            pass
if len(counts) == 2 * len(launches):       # the source page lists every launch twice
    counts = counts[::2]
out = {"_source": note, "_workload": "C4 NavierStokes3D WENO5(mapped)+Rusanov+viscous, 512^3 points per launch"}
names = ["sweep_x", "sweep_y"]
for k, L in enumerate(launches[:3]):
    key = names[k] if k < 2 else ("sweep_fused" if L["rkf"] else "sweep_z")
    d = {a: L[a] for a in ("dram_bytes_read", "dram_bytes_write", "fp64_pipe_active_pct", "time_ms_under_ncu")}
    if k < len(counts):
        d["fp64_thread_instr"] = counts[k]
    d["kernel"] = L["name"][:120]
    out[key] = d
out["_fp64_count"] = "DFMA + DMUL + DADD thread-instructions executed per launch, from the source page of the same capture"
print(json.dumps(out, indent=2))
