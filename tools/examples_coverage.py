"""Which of the reference's own run directories (Examples/**/solver.inp) ask only for features on the B200 path?

    python tools/examples_coverage.py [/root/reference/Examples]          (authoring container only)

Every directory's solver.inp / boundary.inp / physics.inp / weno.inp / muscl.inp are read with the package's own readers
and handed to config_from_inputs + hpb_create (host set-up only, no device needed) on a coarsened grid -- the same
accept / reject decisions a user's run would meet. Ensemble directories (simulation.inp) are read the way
hypar_b200.ensemble does. Directories that need the sparse-grids driver or an immersed body are rejected here, since those
inputs never reach the library. Prints one line per
directory and the totals by reason; DESIGN.md section 1 quotes the totals.
"""
import collections
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypar_b200 import hypario  # noqa: E402
from hypar_b200.solver import MODELS, HyParB200Error, Solver  # noqa: E402


def classify(d):
    nsims = hypario.read_simulation_inp(os.path.join(d, "simulation.inp"))
    # ensembles: per-simulation size / iproc vectors; the simulations differ in size only, the first one decides
    s = hypario.read_ensemble_solver_inp(os.path.join(d, "solver.inp"), nsims)[0]
    model = str(s.get("model", "none"))
    if model not in MODELS:
        return f"model {model}"
    if os.path.exists(os.path.join(d, "sparse_grids.inp")):
        return "sparse-grids driver"
    if str(s.get("immersed_body", "none")) != "none":
        return "immersed boundary"
    nd, nv = int(s["ndims"]), int(s["nvars"])
    try:
        b = hypario.read_boundary_inp(os.path.join(d, "boundary.inp"), nd, nv)
    except Exception as e:                                    # zone types whose extra lines the reader does not know
        return f"boundary.inp: {str(e)[:60]}"
    vk = {"gravity": 3 if model == "navierstokes3d" else 2 if model == "navierstokes2d" else 1,
          "advection": nd * nv, "diffusion": nd * nv}
    pf = os.path.join(d, "physics.inp")
    ph = hypario.read_keyword_file(pf, vector_keys=vk) if os.path.exists(pf) else {}
    for k in ("advection", "diffusion", "gravity"):
        if k in ph:
            ph[k] = [float(v) for v in (ph[k] if isinstance(ph[k], (list, tuple)) else [ph[k]])]
    wf, mf = os.path.join(d, "weno.inp"), os.path.join(d, "muscl.inp")
    w = hypario.read_keyword_file(wf) if os.path.exists(wf) else None
    mu = hypario.read_keyword_file(mf) if os.path.exists(mf) else None
    # coarsen: the decision does not depend on the grid size; keep the set-up cheap
    s = dict(s)
    s["size"] = [min(int(n), 16) for n in s["size"]][:nd]
    s["iproc"] = [1] * nd
    x = [np.linspace(0.0, 1.0, n, endpoint=False) for n in s["size"]]
    af = None
    if str(ph.get("advection_filename", "none")) != "none":
        af = np.ones(tuple(reversed(s["size"])) + (nd * nv,))
    try:
        sv = Solver(s, b, ph, w, x, muscl=mu, advection_field=af)
        sv.close()
    except HyParB200Error as e:
        return str(e)[:80]
    return None


def main():
    top = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/Examples"
    dirs = sorted(os.path.dirname(os.path.join(r, f)) for r, _, fs in os.walk(top) for f in fs if f == "solver.inp")
    why = collections.Counter()
    ok = 0
    for d in dirs:
        try:
            r = classify(d)
        except Exception as e:                                # unreadable input: count, do not stop
            r = f"reader: {type(e).__name__} {str(e)[:60]}"
        print(f"{'ok  ' if r is None else 'NO  '}{os.path.relpath(d, top)}" + ("" if r is None else f"   [{r}]"))
        if r is None:
            ok += 1
        else:
            why[r] += 1
    print(f"\n{ok} of {len(dirs)} run directories are on the B200 path")
    for k, v in why.most_common():
        print(f"  {v:3d}  {k}")


if __name__ == "__main__":
    main()
