"""Fused Roe upwinding (NavierStokes3DUpwind.c:40-125, NavierStokes2DUpwind.c Roe: R |Lambda| L (uR - uL) with Harten's
entropy fix) inside the production sweeps: the closed-form wave decomposition of sweep_fused.cuh::roe_dissipation against
the oracle's matrix products, per RHS evaluation (<= 1e-12, tests/test_gpu_parity.py::fused_tolerance) and over 5 steps
(<= 1e-11), on the TMA-fed kernel and on the cp.async fallback (use_fused = 2), with viscous terms and walls (gravity needs
Rusanov upwinding: NavierStokes3DInitialize.c:370-377 refuses anything else, and so does hpb_create). The
configuration of the reference's own flagship CUDA run (Examples/3D/NavierStokes3D/DNS_IsotropicTurbulenceDecay_CUDA:
weno5 mapped + Roe + viscous, SSPRK3) is the first case."""
import numpy as np
import pytest

from conftest import rel_linf
from hypar_b200 import cases
from hypar_b200.solver import Solver
from oracle import hpo
from test_gpu_parity import fused_tolerance

pytestmark = pytest.mark.gpu

ROE = [
    cases.with_time_scheme(cases.ns3d_turbulence((20, 14, 12), "mapped", upwinding="roe"), "rk", "ssprk3"),
    cases.ns3d_turbulence((16, 12, 14), "js", viscous=False, upwinding="roe"),
    cases.ns3d_turbulence((14, 16, 12), "z", upwinding="roe"),
    cases.ns3d_turbulence((12, 14, 16), "yc", upwinding="roe"),
    cases.ns3d_density_wave((16, 12, 10), "js", upwinding="roe"),
    cases.ns2d_vortex((40, 28), "mapped", upwinding="roe"),
    cases.ns2d_vortex((28, 40), "z", upwinding="roe"),
    cases.ns_channel((28, 24), "js", bcs="amb2", upwinding="roe"),
    cases.ns_channel((14, 12, 12), "mapped", viscous=True, bcs="amb3", upwinding="roe"),
]
nl = cases.ns3d_density_wave((12, 10, 8), "js", upwinding="roe")
nl.weno["no_limiting"] = 1
nl.name += "_nolimiting"
ROE.append(nl)
for c in ROE:
    c.name += "_roe"


@pytest.mark.parametrize("case", ROE, ids=[c.name for c in ROE])
@pytest.mark.parametrize("mode", [1, 2], ids=["tma", "cpasync"])
def test_fused_roe_rhs_and_steps(need_gpu, case, mode):
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u_ref = S.local_u0()
    rhs_ref, hyp_ref, par_ref, src_ref = O.rhs(u_ref, parts=True)
    dt = float(case.solver["dt"])
    sv = Solver.from_case(case, use_fused=mode)
    u = S.local_u0()
    rhs = sv.RHSFunction(u)
    hyp = sv.HyperbolicFunction(u)
    assert np.isfinite(hyp).all() and np.isfinite(rhs).all()
    tol = fused_tolerance(O, u_ref, dt, hyp_ref)
    assert np.abs(hyp - hyp_ref).max() <= tol, \
        f"HyperbolicFunction: abs err {np.abs(hyp - hyp_ref).max():.3e} > {tol:.3e} (rel {rel_linf(hyp, hyp_ref):.3e})"
    scale = max(np.abs(hyp_ref).max(), np.abs(par_ref).max(), np.abs(src_ref).max())
    tol = fused_tolerance(O, u_ref, dt, np.array([scale]))
    assert np.abs(rhs - rhs_ref).max() <= tol, f"rhs: abs err {np.abs(rhs - rhs_ref).max():.3e} > {tol:.3e}"
    # the production kernels ran (not the per-interface exact ones)
    if mode == 1 and S.shape_g()[-2] % 2 == 0:        # even padded row length: the TMA-fed sweep
        assert sv.tma_launches > 0
    rk = hpo.rk_type_of(case)
    for _ in range(5):
        O.time_step(u_ref, dt, rk)
    sv.set_solution(S.local_u0())
    sv.TimeSteps(5)
    u = sv.get_solution()
    a, b = S.interior(u), S.interior(u_ref)
    assert rel_linf(a, b) <= 1e-11, f"u after 5 steps: rel err {rel_linf(a, b):.3e}"
    sv.close()


def test_fused_roe_matches_exact_roe_on_a_shock(need_gpu):
    """a Riemann problem (strong jumps: every wave family and the entropy fix's |lambda| ~ 0 at rest) -- fused vs exact path"""
    case = cases.ns2d_vortex((48, 32), "js", upwinding="roe")
    x, y = np.meshgrid(case.x[0], case.x[1])
    u0 = case.u0
    left = x < 0.5 * (case.x[0][0] + case.x[0][-1])
    rho = np.where(left, 1.0, 0.125)
    p = np.where(left, 1.0, 0.1)
    u0[..., 0] = rho
    u0[..., 1] = 0.0
    u0[..., 2] = 0.0
    u0[..., 3] = p / 0.4
    S = hpo.Setup(case)
    A, B = Solver.from_case(case, use_fused=False), Solver.from_case(case, use_fused=True)
    u = S.local_u0()
    A.ApplyBoundaryConditions(u)                  # HyperbolicFunction does not fill the ghost cells itself
    ha, hb = A.HyperbolicFunction(u), B.HyperbolicFunction(u)
    assert np.isfinite(ha).all() and np.abs(ha).max() > 0
    assert np.abs(ha - hb).max() <= 1e-12 * np.abs(ha).max()
    A.close()
    B.close()


def test_gravity_needs_rusanov_like_the_reference(need_gpu):
    """NavierStokes3DInitialize.c:370-377: "rusanov upwinding is needed for flows with gravitational forces" """
    from hypar_b200.solver import HyParB200Error
    with pytest.raises(HyParB200Error, match="rusanov"):
        Solver.from_case(cases.ns3d_rising_bubble((12, 16, 10), "yc", upwinding="roe"))


# ------------------------------------------------------------------------------------------------------------------
# characteristic-wise WENO5 inside the production sweep (Interp1PrimFifthOrderWENOChar.c:86-198 with the weights of
# WENOFifthOrderCalculateWeights.c:760-1440): NavierStokes2D / 3D without gravity, Roe or Rusanov
CHAR = [
    cases.ns3d_turbulence((20, 14, 12), "mapped", interp="characteristic", upwinding="roe"),          # the reference's default interp type
    cases.ns3d_turbulence((16, 12, 14), "js", viscous=False, interp="characteristic", upwinding="roe"),
    cases.ns3d_turbulence((14, 16, 12), "z", interp="characteristic"),
    cases.ns3d_turbulence((12, 14, 16), "yc", viscous=False, interp="characteristic"),
    cases.ns3d_density_wave((16, 12, 10), "mapped", upwinding="roe"),
    cases.ns2d_vortex((40, 28), "mapped", upwinding="roe", interp="characteristic"),
    cases.ns2d_vortex((28, 40), "js", interp="characteristic"),
    cases.ns_channel((28, 24), "z", bcs="amb2", upwinding="roe"),
    cases.ns_channel((14, 12, 12), "yc", viscous=True, bcs="amb3", upwinding="roe"),
]
for c in CHAR[4:5] + CHAR[7:]:
    c.solver["hyp_interp_type"] = "characteristic"
for c in CHAR:
    c.name += "_charfused"


@pytest.mark.parametrize("case", CHAR, ids=[c.name for c in CHAR])
def test_fused_characteristic_rhs_and_steps(need_gpu, case):
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u_ref = S.local_u0()
    rhs_ref, hyp_ref, par_ref, src_ref = O.rhs(u_ref, parts=True)
    dt = float(case.solver["dt"])
    sv = Solver.from_case(case, use_fused=True)
    u = S.local_u0()
    rhs = sv.RHSFunction(u)
    hyp = sv.HyperbolicFunction(u)
    assert sv.tma_launches > 0, "the characteristic-wise production kernel did not run"
    assert np.isfinite(hyp).all() and np.isfinite(rhs).all()
    tol = fused_tolerance(O, u_ref, dt, hyp_ref)
    assert np.abs(hyp - hyp_ref).max() <= tol, \
        f"HyperbolicFunction: abs err {np.abs(hyp - hyp_ref).max():.3e} > {tol:.3e} (rel {rel_linf(hyp, hyp_ref):.3e})"
    scale = max(np.abs(hyp_ref).max(), np.abs(par_ref).max(), np.abs(src_ref).max())
    tol = fused_tolerance(O, u_ref, dt, np.array([scale]))
    assert np.abs(rhs - rhs_ref).max() <= tol, f"rhs: abs err {np.abs(rhs - rhs_ref).max():.3e} > {tol:.3e}"
    rk = hpo.rk_type_of(case)
    for _ in range(5):
        O.time_step(u_ref, dt, rk)
    sv.set_solution(S.local_u0())
    sv.TimeSteps(5)
    a, b = S.interior(sv.get_solution()), S.interior(u_ref)
    assert rel_linf(a, b) <= 1e-11, f"u after 5 steps: rel err {rel_linf(a, b):.3e}"
    sv.close()


def test_fused_characteristic_on_a_shock(need_gpu):
    """a 2-D Riemann problem: the characteristic projection matters where the jumps are -- fused vs exact path"""
    case = cases.ns2d_vortex((48, 32), "js", upwinding="roe", interp="characteristic")
    x, y = np.meshgrid(case.x[0], case.x[1])
    u0 = case.u0
    left = x < 0.5 * (case.x[0][0] + case.x[0][-1])
    low = y < 0.5 * (case.x[1][0] + case.x[1][-1])
    rho = np.where(left, 1.0, 0.125) * np.where(low, 1.0, 0.7)
    p = np.where(left, 1.0, 0.1) * np.where(low, 1.0, 0.8)
    u0[..., 0] = rho
    u0[..., 1] = 0.1 * rho
    u0[..., 2] = -0.05 * rho
    u0[..., 3] = p / 0.4 + 0.5 * rho * (0.1 ** 2 + 0.05 ** 2)
    S = hpo.Setup(case)
    A, B = Solver.from_case(case, use_fused=False), Solver.from_case(case, use_fused=True)
    u = S.local_u0()
    A.ApplyBoundaryConditions(u)
    ha, hb = A.HyperbolicFunction(u), B.HyperbolicFunction(u)
    assert B.tma_launches > 0 and np.isfinite(ha).all() and np.abs(ha).max() > 0
    assert np.abs(ha - hb).max() <= 1e-12 * np.abs(ha).max(), f"{np.abs(ha - hb).max() / np.abs(ha).max():.3e}"
    A.close()
    B.close()
