"""Debug helper (GPU): serial vs sweep-wise stage sequence, and run-to-run determinism, on a decomposed case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _multirank import LocalRanks, MultiRankOracle
from hypar_b200 import cases

case = cases.ns3d_turbulence((26, 24, 22), "mapped", iproc=(2, 2, 2))
MO = MultiRankOracle(case)
runs = {"serial1": LocalRanks(case, True, sweepwise=False), "serial2": LocalRanks(case, True, sweepwise=False),
        "sweep1": LocalRanks(case, True, sweepwise=True), "sweep2": LocalRanks(case, True, sweepwise=True)}
for R in runs.values():
    R.set_solution(MO.local_u0())
for step in range(3):
    for R in runs.values():
        R.time_step()
    sol = {k: R.get_solution() for k, R in runs.items()}
    ref = sol["serial1"]
    for k in ("serial2", "sweep1", "sweep2"):
        d = [float(np.abs(a - b).max()) for a, b in zip(sol[k], ref)]
        print(f"step {step + 1}: max |{k} - serial1| per rank:", " ".join(f"{x:.1e}" for x in d), flush=True)
# stage-level: compare Udot of each stage after one more step
for R in runs.values():
    R.time_step()
for s in range(4):
    ks = {k: [sv.get_stage_rhs(s) for sv in R.sv] for k, R in runs.items()}
    for k in ("serial2", "sweep1", "sweep2"):
        d = [float(np.abs(a - b).max()) for a, b in zip(ks[k], ks["serial1"])]
        print(f"stage {s}: max |Udot {k} - serial1|:", " ".join(f"{x:.1e}" for x in d), flush=True)
