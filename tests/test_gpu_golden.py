"""GPU parity against the REFERENCE'S OWN OUTPUTS: the committed golden fixtures (tests/golden/*.npz, written by
tools/make_golden.py from the unmodified reference executable) compared with the CUDA path directly -- no oracle in
between. Exact path: bit-identical (NaNs of the never-filled corner ghosts excepted: the device leaves them alone);
production path: <= 1e-12 per RHS evaluation (fused_tolerance), <= 1e-11 after the reference's 3 steps."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import assert_exact, rel_linf
from hypar_b200 import cases
from hypar_b200.solver import Solver

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in meta["kwargs"].items()}
    return z, getattr(cases, meta["builder"])(**kw)


def _local_u0(sv, case):
    return sv.local_from_global(np.asarray(case.u0, dtype=np.float64))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_exact_path_reproduces_the_reference_files(need_gpu, path):
    z, case = _load(path)
    sv = Solver.from_case(case, use_fused=False)
    u = _local_u0(sv, case)
    rhs = sv.RHSFunction(u)
    fin = np.isfinite(z["rhs_u"])
    assert np.array_equal(u[fin], z["rhs_u"][fin]), "u after boundary conditions"
    # viscous channel cases: temperatures on which CUDA's and glibc's exp / log differ by an ulp (conftest.assert_exact)
    ulp = case.name.startswith("chan") and float(case.physics.get("Re", -1.0)) > 0
    for name, got in (("hyp", sv.HyperbolicFunction(u)), ("par", sv.ParabolicFunction(u)),
                      ("source", sv.SourceFunction(u)), ("rhs", rhs)):
        assert_exact(got, z["rhs_" + name], name, libm_ulp=ulp and name in ("par", "rhs"))
    sv.set_solution(_local_u0(sv, case))
    sv.TimeSteps(3)
    got, ref = sv.interior(sv.get_solution()), sv.interior(z["steps3_u"])
    assert_exact(got, ref, "u after 3 steps", libm_ulp=ulp)
    assert sv.kernel_launches > 0
    sv.close()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_production_path_against_the_reference_files(need_gpu, path):
    z, case = _load(path)
    sv = Solver.from_case(case, use_fused=True)
    u = _local_u0(sv, case)
    rhs = sv.RHSFunction(u)
    hyp = sv.HyperbolicFunction(u)
    dt = float(case.solver["dt"])
    # 1e-12 relative + the rounding floor of the Rusanov dissipation (tests/test_gpu_parity.py: fused_tolerance)
    lam = sv.ComputeCFL(u) / dt
    floor = 16 * np.finfo(np.float64).eps * lam * np.abs(u[np.isfinite(u)]).max()
    assert np.abs(hyp - z["rhs_hyp"]).max() <= 1e-12 * np.abs(z["rhs_hyp"]).max() + floor
    scale = max(np.abs(z["rhs_hyp"]).max(), np.abs(z["rhs_par"]).max(), np.abs(z["rhs_source"]).max())
    assert np.abs(rhs - z["rhs_rhs"]).max() <= 1e-12 * scale + floor
    sv.set_solution(_local_u0(sv, case))
    sv.TimeSteps(3)
    got, ref = sv.interior(sv.get_solution()), sv.interior(z["steps3_u"])
    assert rel_linf(got, ref) <= 1e-11, f"u after 3 steps: rel Linf {rel_linf(got, ref):.3e}"
    sv.close()
