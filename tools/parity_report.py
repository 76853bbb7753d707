"""Print the GPU-vs-oracle error of every parity case (both hyperbolic implementations).
TEST INFRASTRUCTURE: imports the oracle. Usage: python tools/parity_report.py [--lib path]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
if "--lib" in sys.argv:
    from hypar_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(sys.argv[sys.argv.index("--lib") + 1])
from hypar_b200.solver import Solver
from oracle import hpo
import test_gpu_parity as T

def rel(a, b):
    s = np.abs(b).max()
    return np.abs(a - b).max() / (s if s > 0 else 1.0)

print(f"{'case':44s} {'mode':8s} {'hyp':>10s} {'par':>10s} {'src':>10s} {'rhs/scale':>10s}  bitexact(hyp)")
for case in T.CASES:
    S = hpo.Setup(case); O = hpo.Oracle(S)
    u_ref = S.local_u0()
    rhs_ref, hyp_ref, par_ref, src_ref = O.rhs(u_ref, parts=True)
    for fused in (0, 1):
        sv = Solver.from_case(case, use_fused=bool(fused))
        u = S.local_u0()
        rhs = sv.RHSFunction(u)
        hyp = sv.HyperbolicFunction(u); par = sv.ParabolicFunction(u); src = sv.SourceFunction(u)
        scale = max(np.abs(hyp_ref).max(), np.abs(par_ref).max(), np.abs(src_ref).max())
        print(f"{case.name:44s} {'fused' if fused else 'generic':8s} {rel(hyp,hyp_ref):10.2e} {rel(par,par_ref):10.2e} "
              f"{rel(src,src_ref):10.2e} {np.abs(rhs-rhs_ref).max()/scale:10.2e}  {np.array_equal(hyp,hyp_ref)}")
        sv.close()
