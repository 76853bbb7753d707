"""GLM-GEE time integrators on the device (TimeGLMGEE.c through hpb_TimeSteps / hpb_TimeStepsLocal): the solution, the
auxiliary solution and TimeError's numbers against the reference's own files (tests/golden/glmgee) and against the oracle.
Exact path: bit-identical; production path: <= 1e-11 after the steps. Decomposed runs on one GPU, both schedules."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import assert_exact, rel_linf
from hypar_b200 import cases
from hypar_b200.solver import Solver
from oracle import hpo

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "glmgee", "*.npz")))


def _load(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    return z, getattr(cases, meta["builder"])(**meta["kwargs"])


def _local_u0(sv, case):
    return sv.local_from_global(np.asarray(case.u0, dtype=np.float64))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
@pytest.mark.parametrize("fused", [False, True], ids=["exact", "production"])
def test_against_the_reference_files(need_gpu, path, fused):
    z, case = _load(path)
    case.solver["conservation_check"] = "yes"
    sv = Solver.from_case(case, use_fused=fused)
    sv.set_solution(_local_u0(sv, case))
    sbi = []
    for _ in range(3):
        sv.TimeSteps(1)
        sbi.append(sv.dev_StepBoundaryIntegral())
    u, ua = sv.interior(sv.get_solution()), sv.interior(sv.get_aux_solution())
    ru, ra = sv.interior(z["steps3_u"]), sv.interior(z["steps3_uaux"])
    err = sv.glmgee_error()
    if fused:
        assert rel_linf(u, ru) <= 1e-11, f"u after 3 steps: {rel_linf(u, ru):.3e}"
        # the auxiliary solution is an error estimate (yeps) or a second solution (yyt): compare on the solution's scale
        assert np.abs(ua - ra).max() <= 1e-11 * np.abs(ru).max(), f"aux after 3 steps: {np.abs(ua - ra).max():.3e}"
    else:
        viscous = float(case.physics.get("Re", -1.0)) > 0
        assert_exact(u, ru, "u after 3 steps", libm_ulp=viscous)
        assert_exact(ua, ra, "auxiliary solution after 3 steps", libm_ulp=viscous)
        # TimeError's norms: deterministic tree sums on the device, the reference's serial sums (rounding of a re-ordered sum)
        assert np.allclose(err[:3], z["steps3_glmerr"][:3], rtol=1e-12, atol=0.0), (err, z["steps3_glmerr"])
        assert (err[3:] == -1.0).all()
    # StepBoundaryIntegral = sum_i dt B[0][i] BoundaryFlux[i] (TimeGLMGEE.c:131-140)
    for k in range(3):
        ref = z["cons_stepbi"][k]
        assert np.abs(sbi[k] - ref).max() <= 1e-11 * max(np.abs(ref).max(), 1e-300) + 1e-13 * np.abs(ru).max(), f"StepBoundaryIntegral, step {k + 1}"
    sv.close()


def _cases():
    return [
        cases.with_glmgee(cases.linear_advection_sine(64, "mapped"), "24", "yyt"),
        cases.with_glmgee(cases.burgers_nd((32,), "js"), "25i"),
        cases.with_glmgee(cases.euler1d_sod(101, "z"), "35"),
        cases.with_glmgee(cases.euler1d_sod(101, "js", scheme="crweno5"), "23", "yyt"),
        cases.with_glmgee(cases.ns2d_vortex((24, 20), "yc"), "exrk2a"),
        cases.with_glmgee(cases.ns3d_turbulence((12, 10, 8), "mapped"), "rk32g1", "yyt"),
        cases.with_glmgee(cases.ns3d_turbulence((12, 10, 8), "z", upwinding="roe"), "24"),
        cases.with_glmgee(cases.ns3d_rising_bubble((10, 12, 8), "js"), "rk285ex"),
    ]


CASES = _cases()


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
@pytest.mark.parametrize("fused", [False, True], ids=["exact", "production"])
def test_against_the_oracle(need_gpu, case, fused):
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    m, mode = hpo.glmgee_of(case)
    dt = float(case.solver["dt"])
    u = S.local_u0()
    ua = O.glmgee_aux0(u, mode)
    uex = S.local_u0()                  # any field serves as "exact solution" for the error-of-the-estimate norms
    sv = Solver.from_case(case, use_fused=fused)
    sv.set_solution(S.local_u0())
    for _ in range(4):
        O.time_step_glmgee(u, ua, dt, m, mode)
    sv.TimeSteps(4)
    gu, ga = sv.get_solution().reshape(u.shape), sv.get_aux_solution().reshape(u.shape)
    ref_err = O.glmgee_error(u, ua, m, mode, uex)
    err = sv.glmgee_error(uex)
    if fused:
        assert rel_linf(S.interior(gu), S.interior(u)) <= 1e-11
        assert np.abs(S.interior(ga) - S.interior(ua)).max() <= 1e-11 * np.abs(S.interior(u)).max()
        assert np.allclose(err, ref_err, rtol=1e-6)
    else:
        viscous = float(case.physics.get("Re", -1.0)) > 0
        assert_exact(S.interior(gu), S.interior(u), "u after 4 steps", libm_ulp=viscous)
        assert_exact(S.interior(ga), S.interior(ua), "auxiliary solution after 4 steps", libm_ulp=viscous)
        assert np.allclose(err, ref_err, rtol=1e-11, atol=0.0), (err, ref_err)
    # the step norm of TimePostStep.c:44-63 (u - u_prev) comes from the kept previous solution
    n2 = sv.dev_StepNormSumSq()
    up = u.copy()
    ub, uab = S.local_u0(), O.glmgee_aux0(S.local_u0(), mode)
    for _ in range(3):
        O.time_step_glmgee(ub, uab, dt, m, mode)
    ref_n2 = float(((S.interior(up) - S.interior(ub)) ** 2).sum())
    assert abs(n2 - ref_n2) <= 1e-9 * ref_n2 + 1e-30, (n2, ref_n2)
    sv.close()


def _decomposed():
    return [cases.with_glmgee(cases.ns2d_vortex((28, 24), "js", iproc=(2, 2)), "exrk2a", "yyt"),
            cases.with_glmgee(cases.ns3d_turbulence((14, 12, 26), "mapped", iproc=(1, 1, 2)), "23"),
            cases.with_glmgee(cases.ns3d_rising_bubble((14, 26, 12), "yc", iproc=(1, 2, 1)), "rk32g1")]


DEC = _decomposed()


@pytest.mark.parametrize("case", DEC, ids=[c.name for c in DEC])
@pytest.mark.parametrize("overlap", [False, True], ids=["serial", "overlap"])
@pytest.mark.parametrize("fused", [False, True], ids=["exact", "production"])
def test_decomposed(need_gpu, case, overlap, fused):
    from _multirank import LocalRanks, MultiRankOracle
    nr = int(np.prod(case.solver["iproc"]))
    m, mode = hpo.glmgee_of(case)
    MO = MultiRankOracle(case)
    u = MO.local_u0()
    ua = [MO.O[r].glmgee_aux0(u[r], mode) for r in range(nr)]
    for _ in range(2):
        MO.time_step_glmgee(u, ua, float(case.solver["dt"]), m, mode)
    LR = LocalRanks(case, use_fused=fused, sweepwise=overlap)
    LR.set_solution(MO.local_u0())
    LR.time_step(2)
    gu = LR.get_solution()
    ga = [sv.get_aux_solution() for sv in LR.sv]
    for r in range(nr):
        S = MO.S[r]
        a, b = S.interior(gu[r].reshape(u[r].shape)), S.interior(u[r])
        c, d = S.interior(ga[r].reshape(u[r].shape)), S.interior(ua[r])
        if fused:
            assert rel_linf(a, b) <= 1e-11 and np.abs(c - d).max() <= 1e-11 * np.abs(b).max(), f"rank {r}"
        else:
            viscous = float(case.physics.get("Re", -1.0)) > 0
            assert_exact(a, b, f"rank {r}: u after 2 steps", libm_ulp=viscous)
            assert_exact(c, d, f"rank {r}: auxiliary solution after 2 steps", libm_ulp=viscous)
    LR.close()
