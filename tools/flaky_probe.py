"""Probe for run-to-run nondeterminism of the exact path (TEST INFRASTRUCTURE)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from hypar_b200 import cases
from hypar_b200.solver import Solver
from oracle import hpo

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for case in (cases.euler1d_sod(101, "js"), cases.ns3d_turbulence((16, 12, 10), "js", viscous=False, upwinding="roe")):
    S = hpo.Setup(case); O = hpo.Oracle(S)
    dt = float(case.solver["dt"]); rk = hpo.rk_type_of(case)
    u_ref = S.local_u0()
    for _ in range(5):
        O.time_step(u_ref, dt, rk)
    outs = []
    for r in range(reps):
        sv = Solver.from_case(case, use_fused=False)
        sv.set_solution(S.local_u0())
        sv.TimeSteps(5)
        u = sv.get_solution()
        outs.append(u)
        sv.close()
    d_ref = [float(np.abs(S.interior(u) - S.interior(u_ref)).max()) for u in outs]
    d_self = [float(np.abs(u - outs[0]).max()) for u in outs]
    print(case.name, "vs oracle:", sorted(set(d_ref)), " vs first run:", sorted(set(d_self)))
