"""Small invocations of every production kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_cases.py [quick]

TMA-fed sweeps (viscous, gravity, Roe), the cp.async sweeps (use_fused = 2), the Q-derivative kernels, the exact kernels,
and one decomposed run (2x2x2 ranks in this process, overlapped schedule: face-layer RK kernel, merged pack / unpack, the
in-process transport's streams and events). Results are checked against the oracle so that a sanitizer-clean run is also a
correct one. TEST INFRASTRUCTURE (imports the oracle)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from hypar_b200 import cases
from hypar_b200.solver import Solver
from oracle import hpo

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def single(case, mode, steps=1):
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u_ref = S.local_u0()
    rhs_ref, hyp_ref, par_ref, src_ref = O.rhs(u_ref, parts=True)
    sv = Solver.from_case(case, use_fused=mode)
    rhs = sv.RHSFunction(S.local_u0())
    scale = max(np.abs(hyp_ref).max(), np.abs(par_ref).max(), np.abs(src_ref).max())
    err = np.abs(rhs - rhs_ref).max() / scale
    sv.set_solution(S.local_u0())
    sv.TimeSteps(steps)
    u = sv.get_solution()
    assert np.isfinite(u).all()
    ur = S.local_u0()
    for _ in range(steps):
        O.time_step(ur, float(case.solver["dt"]), hpo.rk_type_of(case))
    e2 = np.abs(S.interior(u) - S.interior(ur)).max() / np.abs(S.interior(ur)).max()
    print(f"{case.name:40s} use_fused={mode}: rhs err/terms {err:.2e}, {steps} step(s) rel err {e2:.2e}, {sv.kernel_launches} launches "
          f"({sv.tma_launches} TMA sweeps, stage fusion {'on' if sv.stage_fusion_active else 'off'})", flush=True)
    assert err <= (0.0 if mode == 0 else 1e-9) and e2 <= (0.0 if mode == 0 else 1e-11)
    sv.close()


single(cases.ns3d_turbulence((24, 20, 16), "mapped"), 1)
single(cases.ns3d_rising_bubble((16, 20, 12), "yc"), 1)
single(cases.ns3d_turbulence((20, 16, 12), "mapped", upwinding="roe"), 1)
if not quick:
    single(cases.ns3d_turbulence((20, 16, 12), "js"), 2)
    single(cases.ns2d_vortex((40, 28), "z"), 1)
    single(cases.ns3d_turbulence((16, 12, 10), "mapped"), 0)
    single(cases.euler1d_sod(101, "js"), 0)

from _multirank import LocalRanks, MultiRankOracle
case = cases.ns3d_turbulence((26, 24, 22), "mapped", iproc=(2, 2, 2))
MO = MultiRankOracle(case)
LR = LocalRanks(case, use_fused=True, sweepwise=True)
LR.set_solution(MO.local_u0())
rhs = LR.rhs()
rhs_ref = MO.rhs(MO.local_u0())
err = max(np.abs(a - b).max() for a, b in zip(rhs, rhs_ref)) / max(np.abs(b).max() for b in rhs_ref)
LR.time_step(1)
u = LR.get_solution()
assert all(np.isfinite(x).all() for x in u)
ur = MO.local_u0()
MO.time_step(ur, float(case.solver["dt"]), hpo.rk_type_of(case))
e2 = max(np.abs(MO.S[r].interior(u[r]) - MO.S[r].interior(ur[r])).max() for r in range(MO.nranks))
print(f"decomposed 2x2x2, overlapped schedule: rhs rel err {err:.2e}, one step abs err {e2:.2e}", flush=True)
assert err <= 1e-11 and e2 <= 1e-11
LR.close()
print("SANITIZE CASES OK", flush=True)
