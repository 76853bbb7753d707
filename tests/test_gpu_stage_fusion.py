"""Stage fusion (csrc/sweep_tma.cuh, RKF): the last directional sweep of an RK stage completes k in its result tile and
writes the next stage solution U_{s+1} = u + a dt k_s beside it (TimeRK.c:131-141) -- k_rk_combine's two roundings, so the
step must not change BY A BIT against the unfused schedule (hpb_set_stage_fusion 0). Checked on one rank and on
decomposed runs (both exchange schedules), grids whose tiles hang over the block on every side, all weight types,
Rusanov / Roe, gravity, viscous terms, RK4 (every row fusable), SSPRK3 (one row), 2-D and 3-D; and against the oracle."""
import numpy as np
import pytest

from conftest import rel_linf
from hypar_b200 import cases
from hypar_b200.solver import Solver
from oracle import hpo

pytestmark = pytest.mark.gpu


def _cases():
    C = [cases.ns3d_turbulence((24, 20, 16), "mapped"),                                       # C4: viscous, RK4
         cases.ns3d_turbulence((38, 13, 35), "js"),                                           # tiles hang over in x and z
         cases.ns3d_turbulence((16, 12, 70), "z", viscous=False, upwinding="roe"),            # three march steps, Roe
         cases.with_time_scheme(cases.ns3d_turbulence((20, 14, 12), "yc"), "rk", "ssprk3"),   # one fusable row
         cases.ns3d_density_wave((16, 12, 10), "js"),                                         # C5a: SSPRK3, inviscid
         cases.ns3d_rising_bubble((12, 16, 34), "yc"),                                        # C5b: walls + gravity source
         cases.with_time_scheme(cases.ns3d_rising_bubble((14, 12, 10), "mapped", hb=1), "rk", "44"),
         cases.ns2d_vortex((40, 28), "mapped"),                                               # 2-D: the y-sweep is the last one
         cases.with_time_scheme(cases.ns2d_vortex((26, 45), "z", upwinding="roe"), "rk", "44"),
         cases.with_time_scheme(cases.ns3d_turbulence((12, 10, 14), "js"), "rk", "22"),
         cases.with_time_scheme(cases.ns3d_turbulence((12, 10, 14), "js"), "rk", "33")]       # rows with two entries: not fusable
    nl = cases.ns3d_density_wave((12, 10, 8), "js")
    nl.weno["no_limiting"] = 1
    nl.name += "_nolimiting"
    return C + [nl]


CASES = _cases()


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_fused_stage_vector_is_bit_identical(need_gpu, case):
    S = hpo.Setup(case)
    u0 = S.local_u0()
    out = {}
    for on in (True, False):
        sv = Solver.from_case(case, use_fused=True)
        sv.set_stage_fusion(on)
        assert sv.stage_fusion_active == on, "every case here runs the TMA-fed sweeps"
        sv.set_solution(u0.copy())
        n0 = sv.kernel_launches
        sv.TimeSteps(3)
        out[on] = (sv.get_solution(), [sv.get_stage_rhs(s) for s in range(sv.nstages)], sv.kernel_launches - n0)
        sv.close()
    (ua, ka, na), (ub, kb, nb) = out[True], out[False]
    fin = np.isfinite(ub)
    assert np.array_equal(ua[fin], ub[fin]), f"solution differs by {np.abs(ua[fin] - ub[fin]).max():.3e}"
    for s, (a, b) in enumerate(zip(ka, kb)):
        a, b = S.interior(a), S.interior(b)
        assert np.array_equal(a, b), f"stage {s} right-hand side differs by {np.abs(a - b).max():.3e}"
    rk = hpo.rk_type_of(case)
    fusable = {0: 3, 1: 1, 2: 0, 3: 1, 4: 1}[rk]          # rows with the single entry a_{s+1,s}: 44, ssprk3, 1fe, 22, 33
    assert nb - na == 3 * fusable, f"launches: fused {na}, unfused {nb}"
    # and the fused step against the oracle
    O = hpo.Oracle(S)
    ur = S.local_u0()
    for _ in range(3):
        O.time_step(ur, float(case.solver["dt"]), rk)
    assert rel_linf(S.interior(ua), S.interior(ur)) <= 1e-11


def test_not_used_where_it_does_not_apply(need_gpu):
    for case in (cases.with_sponge(cases.ns3d_turbulence((16, 12, 14), "mapped"), 2, -1, 0.0, 3.0, [1.0, 0.1, 0.0, 0.0, 1.8]),
                 cases.ns3d_turbulence((12, 14, 10), "z", viscous=False, interp="characteristic"),
                 cases.linear_advection_nd((24, 20), "mapped"),
                 cases.ns2d_vortex((33, 24), "js"),                      # odd padded row length: no TMA
                 cases.with_glmgee(cases.ns3d_turbulence((12, 10, 8), "mapped"), "23")):
        sv = Solver.from_case(case, use_fused=True)
        assert not sv.stage_fusion_active, case.name
        sv.close()
    sv = Solver.from_case(cases.ns3d_turbulence((12, 10, 8), "mapped"), use_fused=False)
    assert not sv.stage_fusion_active
    sv.close()


def _decomposed():
    # (even padded row length on every rank: the TMA-fed sweeps)
    return [cases.ns3d_turbulence((28, 25, 27), "z", iproc=(2, 2, 2)),
            cases.ns3d_turbulence((14, 26, 40), "mapped", iproc=(1, 2, 2)),
            cases.ns3d_rising_bubble((14, 26, 12), "yc", iproc=(1, 2, 1)),
            cases.ns2d_vortex((40, 28), "mapped", iproc=(2, 2)),
            cases.ns3d_turbulence((16, 14, 26), "mapped", upwinding="roe", iproc=(1, 1, 2))]


DEC = _decomposed()


@pytest.mark.parametrize("case", DEC, ids=[c.name + "_iproc" + "x".join(map(str, c.solver["iproc"])) for c in DEC])
@pytest.mark.parametrize("overlap", [False, True], ids=["serial", "overlap"])
def test_decomposed_fused_stage_vector_is_bit_identical(need_gpu, case, overlap):
    from _multirank import LocalRanks, MultiRankOracle
    MO = MultiRankOracle(case)
    res = {}
    for on in (True, False):
        LR = LocalRanks(case, use_fused=True, sweepwise=overlap)
        for sv in LR.sv:
            sv.set_stage_fusion(on)
            assert sv.stage_fusion_active == on
        LR.set_solution(MO.local_u0())
        LR.time_step(2)
        res[on] = LR.get_solution()
        LR.close()
    for r, (a, b) in enumerate(zip(res[True], res[False])):
        S = MO.S[r]
        a, b = S.interior(a), S.interior(b)
        assert np.array_equal(a, b), f"rank {r}: differs by {np.abs(a - b).max():.3e}"
    u = MO.local_u0()
    for _ in range(2):
        MO.time_step(u, float(case.solver["dt"]), hpo.rk_type_of(case))
    for r in range(MO.nranks):
        S = MO.S[r]
        a = S.interior(res[True][r])
        assert rel_linf(a, S.interior(u[r])) <= 1e-11, f"rank {r} against the oracle"
