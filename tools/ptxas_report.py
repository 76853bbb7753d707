"""Registers / spills / shared memory of every kernel in the ptxas logs of the last build (hypar_b200/csrc/build/*.ptxas.log).
usage: python tools/ptxas_report.py [substring of the demangled-ish name]"""
import glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = sys.argv[1] if len(sys.argv) > 1 else ""
rows = []
for f in sorted(glob.glob(os.path.join(ROOT, "hypar_b200", "csrc", "build", "*.ptxas.log"))):
    name = None
    for ln in open(f):
        m = re.search(r"Compiling entry function '(\S+)'", ln)
        if m:
            name = m.group(1); spill = None; continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m and name:
            spill = (int(m.group(2)), int(m.group(3))); continue
        m = re.search(r"Used (\d+) registers", ln)
        if m and name:
            rows.append((name, int(m.group(1)), spill)); name = None
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
for (mangled, regs, spill), dn in zip(rows, names):
    dn = re.sub(r"\(hpbf::SweepArgs.*", "", dn).replace("(int)", "").replace("(bool)", "")
    if pat in dn:
        print(f"{regs:4d} regs  spill st/ld {spill[0]:4d}/{spill[1]:4d}  {dn}")
