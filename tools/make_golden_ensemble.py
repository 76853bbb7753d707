"""Generate tests/golden/ensemble/*.npz from the reference's OWN main (oracle/_ref/hypar_ref_main = src/main.cpp and every
other reference source, unmodified, compiled where they lie by oracle/Makefile) run as an ensemble (simulation.inp).

    python tools/make_golden_ensemble.py [names ...]

Each fixture holds the final solution file of every simulation (op_<n>.bin after n_iter steps) and the CFL / norm columns of
the screen log; the run directory is re-created from hypar_b200.cases.ensemble(name). TEST INFRASTRUCTURE ONLY.
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypar_b200 import cases, hypario  # noqa: E402
from hypar_b200.ensemble import index_string  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "hypar_ref_main")
NAMES = ["vortex3", "burgers2", "linadvvar2", "sod2", "turb12"]


def run_reference_ensemble(name, d=None):
    sims = cases.ensemble(name)
    d = d or tempfile.mkdtemp(prefix="hpb_ens_")
    cases.write_ensemble(d, sims)
    p = subprocess.run([EXE], cwd=d, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=900)
    if p.returncode != 0 or "Finished." not in p.stdout:
        raise RuntimeError(f"hypar_ref_main failed in {d}:\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
    out = {"stdout": p.stdout, "u": [], "x": []}
    for n in range(len(sims)):
        x, u = hypario.read_op_bin(os.path.join(d, f"op_{index_string(n, len(sims))}.bin"))[:2]
        out["x"].append(x)
        out["u"].append(u)
    out["screen"] = [(int(i), float(c), float(nm)) for i, c, nm in
                     re.findall(r"iter=\s*(\d+),.*?CFL=([-+0-9.Ee]+)\s+norm=([-+0-9.Ee]+)", p.stdout, flags=re.S)]
    return out


def main():
    out_dir = os.path.join(ROOT, "tests", "golden", "ensemble")
    os.makedirs(out_dir, exist_ok=True)
    for name in (sys.argv[1:] or NAMES):
        o = run_reference_ensemble(name)
        data = {f"u_{n}": u for n, u in enumerate(o["u"])}
        data["screen"] = np.array(o["screen"], dtype=np.float64)          # rows: iter, CFL, norm (as printed: 4 / 5 digits)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **data)
        print(name, len(o["u"]), "simulations,", len(o["screen"]), "screen rows")


if __name__ == "__main__":
    main()
