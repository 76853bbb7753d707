"""GPU parity of the conservation / error diagnostics through the C ABI:
StageBoundaryIntegral (HyperbolicFunction.c:103-106), StepBoundaryIntegral (TimeRK.c:172-193), VolumeIntegral.c,
BoundaryIntegral.c, CalculateConservationError.c and the norm sums of CalculateError.c, against the oracle (which is
bit-exact against the reference on these: tests/test_oracle_golden.py) and the golden fixtures of the reference.

Bar: the face fluxes, the per-point products and the step combination are computed with the reference's operations
(exact kernels, no FMA); only the ORDER of the sums differs (deterministic parallel tree vs the reference's serial
loop), so the results agree to rounding of the sum: |diff| <= 1e-13 x (sum of |terms|) -- and two runs agree bit for bit.
"""
import os

import numpy as np
import pytest

from _multirank import LocalRanks, MultiRankOracle
from hypar_b200 import cases
from hypar_b200.solver import HyParB200Error, Solver
from oracle import hpo

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cons(case):
    case.solver["conservation_check"] = "yes"
    return case


CASES = [
    _cons(cases.linear_advection_sine(64, "mapped")),
    _cons(cases.euler1d_sod(101, "js")),                                          # extrapolate: open boundaries
    _cons(cases.euler1d_sod(101, "z", interp="components", upwinding="rusanov")),
    _cons(cases.ns2d_vortex((40, 28), "yc")),
    _cons(cases.ns3d_turbulence((20, 14, 12), "mapped")),                         # periodic, viscous
    _cons(cases.ns3d_turbulence((16, 12, 10), "js", viscous=False, upwinding="roe")),
    _cons(cases.ns3d_density_wave((16, 12, 10), "z")),
    _cons(cases.ns3d_rising_bubble((12, 16, 10), "yc")),                          # slip walls + gravity
]


def _sum_tol(face_terms_abs):
    # entries that vanish identically at the probed state (wall-normal mass flux ...) pick up rounding noise
    # ~1e-30 of the dominant flux at later stages: a floor of 1e-20 x the largest face sum covers them
    return 1e-13 * face_terms_abs + 1e-20 * face_terms_abs.max() + 1e-300


def _face_abs_sums(S, O, u):
    """sum over each face of |interface flux| (the magnitude the rounding of a re-ordered sum scales with)"""
    out = np.zeros(2 * S.ndims * S.nvars)
    for d in range(S.ndims):
        f = O.flux(u, d)
        f = np.where(np.isfinite(f), f, 0.0)
        uc = O.modified_solution(u)
        uc = np.where(np.isfinite(uc), uc, 0.0)
        w = O.weno_weights(f, u, d)
        fi = O.upwind(O.interp(f, u, w, 1, d, 0), O.interp(f, u, w, -1, d, 0), O.interp(uc, u, w, 1, d, 1),
                      O.interp(uc, u, w, -1, d, 1), u, d)
        shp = [S.dim[k] + (1 if k == d else 0) for k in reversed(range(S.ndims))] + [S.nvars]
        fi = np.abs(fi.reshape(shp))
        ax = S.ndims - 1 - d
        lo = np.take(fi, 0, axis=ax).reshape(-1, S.nvars).sum(axis=0)
        hi = np.take(fi, S.dim[d], axis=ax).reshape(-1, S.nvars).sum(axis=0)
        out[(2 * d) * S.nvars:(2 * d + 1) * S.nvars] = lo
        out[(2 * d + 1) * S.nvars:(2 * d + 2) * S.nvars] = hi
    return out


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
@pytest.mark.parametrize("fused", [False, True], ids=["exact", "fused"])
def test_stage_boundary_integral(need_gpu, case, fused):
    """StageBoundaryIntegral left by one TimeRHSFunctionExplicit / HyperbolicFunction call, both device paths."""
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u_ref = S.local_u0()
    sbi_ref = O.stage_boundary_integral(u_ref)            # applies the BCs to u_ref
    mag = _face_abs_sums(S, O, u_ref)
    sv = Solver.from_case(case, use_fused=fused)
    u = S.local_u0()
    sv.RHSFunction(u)
    sbi = sv.dev_StageBoundaryIntegral(-1)
    assert np.all(np.abs(sbi - sbi_ref) <= _sum_tol(mag)), f"RHSFunction: {np.abs(sbi - sbi_ref).max():.3e}"
    sv.HyperbolicFunction(u)
    sbi2 = sv.dev_StageBoundaryIntegral(-1)
    assert np.array_equal(sbi, sbi2), "HyperbolicFunction and RHSFunction leave different StageBoundaryIntegral"
    # deterministic: a second solver gives the same bits
    sv2 = Solver.from_case(case, use_fused=fused)
    sv2.RHSFunction(S.local_u0())
    assert np.array_equal(sv2.dev_StageBoundaryIntegral(-1), sbi)
    sv.close()
    sv2.close()


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_conservation_over_steps(need_gpu, case):
    """TimePostStep.c:81-93 over 3 steps on the exact path (u bit-identical to the oracle's): StepBoundaryIntegral,
    VolumeIntegral, the local BoundaryIntegral and the conservation error of every step."""
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    dt, rk = float(case.solver["dt"]), hpo.rk_type_of(case)
    sv = Solver.from_case(case, use_fused=False)
    u_ref = S.local_u0()
    sv.set_solution(S.local_u0())
    vol0_ref = O.volume_integral(u_ref)
    vol0 = sv.dev_VolumeIntegral()
    volmag = O.volume_integral(np.abs(u_ref))
    assert np.all(np.abs(vol0 - vol0_ref) <= 1e-13 * volmag)
    tbi_ref, tbi = np.zeros(S.nvars), np.zeros(S.nvars)
    for step in range(3):
        mag = _face_abs_sums(S, O, O.apply_bc(u_ref.copy()))
        sbi_ref = O.time_step_cons(u_ref, dt, rk)
        sv.TimeStep()
        sbi = sv.dev_StepBoundaryIntegral()
        assert np.all(np.abs(sbi - sbi_ref) <= 2 * dt * _sum_tol(mag)), f"step {step}: {np.abs(sbi - sbi_ref).max():.3e}"
        # host pieces: bit-identical to the oracle on identical inputs
        assert np.array_equal(sv.BoundaryIntegral(sbi_ref), O.boundary_integral(sbi_ref))
        tbi_ref += O.boundary_integral(sbi_ref)
        tbi += sv.BoundaryIntegral(sbi)
        vol_ref, vol = O.volume_integral(u_ref), sv.dev_VolumeIntegral()
        volmag = np.maximum(volmag, O.volume_integral(np.abs(u_ref)))
        assert np.all(np.abs(vol - vol_ref) <= 1e-13 * volmag), f"step {step}: VolumeIntegral"
        assert np.array_equal(sv.CalculateConservationError(vol_ref, vol0_ref, tbi_ref),
                              O.conservation_error(vol_ref, vol0_ref, tbi_ref))
        # the device's own conservation error is rounding noise of the same size as the reference's
        err = sv.CalculateConservationError(vol, vol0, tbi)
        err_ref = O.conservation_error(vol_ref, vol0_ref, tbi_ref)
        base = np.maximum(np.abs(vol0_ref), 1.0)
        assert np.all(np.abs(err - err_ref) <= 1e-12 * volmag / base + 1e-12 * np.abs(tbi_ref) / base + 1e-15)
    u = sv.get_solution()
    assert np.array_equal(S.interior(u), S.interior(u_ref)), "conservation bookkeeping changed the solution"
    sv.close()


GOLD = [("c3_vortex_yc", lambda: cases.ns2d_vortex((20, 16), "yc")),
        ("c5a_denswave_js", lambda: cases.ns3d_density_wave((12, 10, 8), "js")),
        ("c2_sod_js_char_roe", lambda: cases.euler1d_sod(101, "js"))]


@pytest.mark.parametrize("name,make", GOLD, ids=[g[0] for g in GOLD])
def test_conservation_against_reference_golden(need_gpu, name, make):
    """the reference's own numbers (tests/golden, written by tools/make_golden.py from the unmodified reference):
    VolumeIntegralInitial, per-step VolumeIntegral | TotalBoundaryIntegral | ConservationError, StepBoundaryIntegral"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert "cons_vol0" in z, "fixture without conservation data: regenerate with tools/make_golden.py"
    case = _cons(make())
    S = hpo.Setup(case)
    assert S.local_u0().shape == z["steps3_u"].reshape(-1).shape, "fixture of another grid size"
    sv = Solver.from_case(case, use_fused=False)
    sv.set_solution(S.local_u0())
    vol0 = sv.dev_VolumeIntegral()
    scale = np.maximum(np.abs(z["cons_vol0"]), 1.0)
    assert np.all(np.abs(vol0 - z["cons_vol0"]) <= 1e-12 * scale)
    tbi = np.zeros(S.nvars)
    for k in range(len(z["cons_steps"])):
        sv.TimeStep()
        sbi = sv.dev_StepBoundaryIntegral()
        ref = z["cons_stepbi"][k]
        assert np.all(np.abs(sbi - ref) <= 1e-12 * max(np.abs(ref).max(), 1e-300)), f"StepBoundaryIntegral step {k + 1}"
        tbi += sv.BoundaryIntegral(sbi)
        vol = sv.dev_VolumeIntegral()
        nv = S.nvars
        assert np.all(np.abs(vol - z["cons_steps"][k][:nv]) <= 1e-12 * scale)
        assert np.all(np.abs(tbi - z["cons_steps"][k][nv:2 * nv]) <= 1e-12 * scale)
        err = sv.CalculateConservationError(vol, vol0, tbi)
        assert np.all(np.abs(err - z["cons_steps"][k][2 * nv:]) <= 1e-12)
    sv.close()


@pytest.mark.parametrize("case", [CASES[0], CASES[3], CASES[7]], ids=lambda c: c.name)
def test_error_sums(need_gpu, case):
    """CalculateError.c: the six local sums (norms of uex, norms of u - uex) on the device solution"""
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u = S.local_u0()
    dt, rk = float(case.solver["dt"]), hpo.rk_type_of(case)
    for _ in range(2):
        O.time_step(u, dt, rk)
    uex = S.local_u0()
    sv = Solver.from_case(case, use_fused=False)
    sv.set_solution(S.local_u0())
    sv.TimeSteps(2)
    sums = sv.dev_ErrorSums(uex)
    ref = np.concatenate([O.norm_sums(uex), O.norm_sums(uex, u)])
    assert sums[2] == ref[2] and sums[5] == ref[5], "Linf sums are order-independent: must be identical"
    for k in (0, 1, 3, 4):
        assert abs(sums[k] - ref[k]) <= 1e-13 * abs(ref[k]) + 1e-300, f"sum {k}: {sums[k]!r} vs {ref[k]!r}"
    assert np.array_equal(sv.dev_ErrorSums(uex), sums), "reductions are not deterministic"
    sv.close()


def test_bookkeeping_off_fails_loudly(need_gpu):
    case = cases.ns2d_vortex((24, 20), "js")
    sv = Solver.from_case(case)
    sv.set_solution(hpo.Setup(case).local_u0())
    sv.TimeStep()
    with pytest.raises(HyParB200Error):
        sv.dev_StepBoundaryIntegral()
    sv.L.hpb_clear_error()
    assert sv.dev_VolumeIntegral().shape == (4,)          # needs no bookkeeping
    sv.close()


DECOMP = [_cons(cases.ns3d_turbulence((26, 25, 27), "z", iproc=(2, 2, 2))),
          _cons(cases.ns3d_rising_bubble((14, 26, 12), "yc", iproc=(1, 2, 1))),
          _cons(cases.ns2d_vortex((40, 27), "mapped", iproc=(2, 2)))]


@pytest.mark.parametrize("case", DECOMP, ids=lambda c: c.name + "_" + "x".join(str(v) for v in c.solver["iproc"]))
@pytest.mark.parametrize("sweepwise", [False, True], ids=["serial", "overlapped"])
def test_decomposed_boundary_integral(need_gpu, case, sweepwise):
    """Every rank keeps the flux integrals of its OWN block faces (HyperbolicFunction.c:103-106 uses local indices);
    BoundaryIntegral.c sums the ranks' parts and the internal faces cancel. Per rank against an oracle of the same
    decomposition; the sum over ranks against the single-rank run."""
    if sweepwise and case.solver["model"] != "navierstokes3d":
        pytest.skip("sweep-wise schedule: NavierStokes3D production path")
    MO = MultiRankOracle(case)
    dt, rk = float(case.solver["dt"]), hpo.rk_type_of(case)
    LR = LocalRanks(case, use_fused=True, sweepwise=sweepwise)
    LR.set_solution(MO.local_u0())
    LR.time_step()
    u_ref = MO.local_u0()
    sbi_ref = MO.time_step_cons(u_ref, dt, rk)
    total, total_ref = 0.0, 0.0
    for r, sv in enumerate(LR.sv):
        sbi = sv.dev_StepBoundaryIntegral()
        scale = np.abs(sbi_ref[r]).max()
        assert np.all(np.abs(sbi - sbi_ref[r]) <= 1e-11 * scale), f"rank {r}: {np.abs(sbi - sbi_ref[r]).max():.3e} / {scale:.3e}"
        total = total + sv.BoundaryIntegral(sbi)
        total_ref = total_ref + MO.O[r].boundary_integral(sbi_ref[r])
    mag = sum(np.abs(MO.O[r].boundary_integral(np.abs(sbi_ref[r]))) for r in range(MO.nranks))
    assert np.all(np.abs(total - total_ref) <= 1e-11 * mag)
    LR.close()
