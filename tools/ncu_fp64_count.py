"""FP64 thread-instructions per launch from an ncu source page: ncu -i X.ncu-rep --page source --csv --kernel-name regex:k_sweep
(the CSV lists every SASS instruction of every captured launch; launches are separated by a 'Kernel Name' row).
Prints, per launch: warp instructions, the FP64 share (DFMA + DMUL + DADD) and FP64 thread-instructions."""
import csv, re, sys, json
rows = list(csv.reader(open(sys.argv[1])))
out = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"kernel": r[1], "seen": set(), "warp": 0, "fp64_warp": 0, "fp64_thread": 0, "mix": {}}
        out.append(cur)
        continue
    if r and r[0] == "Address":
        hdr = r
        iA, iS, iE, iT = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if cur is None or len(r) < 10 or r[iA] in cur["seen"]:
        continue
    cur["seen"].add(r[iA])
    try:
        e, t = int(r[iE]), int(r[iT])
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
    op = m.group(2).split(".")[0] if m else "?"
    cur["warp"] += e
    cur["mix"][op] = cur["mix"].get(op, 0) + e
    if op in ("DFMA", "DMUL", "DADD"):
        cur["fp64_warp"] += e
        cur["fp64_thread"] += t
for k in out:
    mix = sorted(k["mix"].items(), key=lambda x: -x[1])[:8]
    print(json.dumps({"kernel": k["kernel"][:60], "warp_instr": k["warp"], "fp64_share": round(k["fp64_warp"] / max(k["warp"], 1), 4),
                      "fp64_thread_instr": k["fp64_thread"], "mix_top": {a: round(b / k["warp"], 4) for a, b in mix}}))
