"""GPU parity of the TMA-fed fused sweep (hypar_b200/csrc/sweep_tma.cuh) -- the production kernel of
NavierStokes2D / NavierStokes3D whenever the padded row length is even.

  * against the oracle (= the reference), <= 1e-12 relative per RHS evaluation (fused_tolerance);
  * against the cp.async variant of the same sweep (use_fused = 2), which differs only in the data movement
    and in re-using the momentum reconstruction for the mass flux;
  * shapes that exercise the tiling: lines that are not a multiple of 8 per CTA tile, line lengths that are not
    a multiple of the 32-cell march step, several march steps, the idle ghost line of the first y/z tile,
    2-D grids, gravity + slip walls, and odd row lengths (which must fall back to the cp.async kernel).
"""
import numpy as np
import pytest

from conftest import rel_linf
from hypar_b200 import cases
from hypar_b200.solver import Solver
from oracle import hpo
from test_gpu_parity import fused_tolerance

pytestmark = pytest.mark.gpu


def _cases():
    C = [
        cases.ns3d_turbulence((70, 44, 36), "mapped"),              # 3 march steps in x, partial tiles everywhere
        cases.ns3d_turbulence((34, 66, 20), "js"),
        cases.ns3d_turbulence((18, 12, 68), "z"),
        cases.ns3d_turbulence((32, 32, 32), "yc"),                  # exact multiples: the start-up step only
        cases.ns3d_turbulence((36, 20, 14), "mapped", viscous=False),
        cases.ns3d_density_wave((40, 22, 18), "mapped"),
        cases.ns3d_rising_bubble((24, 38, 12), "yc"),               # gravity source inside the sweep, slip walls
        cases.ns3d_rising_bubble((14, 34, 10), "mapped", hb=1),
        cases.ns2d_vortex((72, 40), "mapped"),
        cases.ns2d_vortex((36, 70), "z"),
    ]
    nl = cases.ns3d_density_wave((34, 12, 10), "js")
    nl.weno["no_limiting"] = 1
    nl.name += "_nolimiting"
    C.append(nl)
    return C


CASES = _cases()


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_tma_sweep_rhs_parity(need_gpu, case):
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u_ref = S.local_u0()
    rhs_ref, hyp_ref, par_ref, src_ref = O.rhs(u_ref, parts=True)
    dt = float(case.solver["dt"])

    sv = Solver.from_case(case, use_fused=1)
    u = S.local_u0()
    n0 = sv.tma_launches
    rhs = sv.RHSFunction(u)
    assert sv.tma_launches - n0 == S.ndims, "the TMA sweep did not run for every direction"
    assert np.isfinite(rhs).all()
    scale = max(np.abs(hyp_ref).max(), np.abs(par_ref).max(), np.abs(src_ref).max())
    tol = fused_tolerance(O, u_ref, dt, np.array([scale]))
    err = np.abs(rhs - rhs_ref)
    assert err.max() <= tol, f"rhs vs oracle: abs err {err.max():.3e} > {tol:.3e} at {np.unravel_index(err.argmax(), err.shape)}"
    # ghost entries of the right-hand side stay exactly zero (the tiles add +0.0 there)
    gh = rhs.reshape(S.shape_g()).copy()
    g = S.ghosts
    gh[tuple(slice(g, g + n) for n in reversed(S.dim))] = 0.0
    assert not gh.any(), "ghost entries of rhs were modified"
    hyp = sv.HyperbolicFunction(u)
    tolh = fused_tolerance(O, u_ref, dt, hyp_ref)
    assert np.abs(hyp - hyp_ref).max() <= tolh
    sv.close()

    # same sweep without the TMA (cp.async staging): differences are rounding only
    sv2 = Solver.from_case(case, use_fused=2)
    u2 = S.local_u0()
    rhs2 = sv2.RHSFunction(u2)
    assert sv2.tma_launches == 0
    assert np.abs(rhs2 - rhs_ref).max() <= tol
    assert np.abs(rhs - rhs2).max() <= tol
    sv2.close()


@pytest.mark.parametrize("case", [CASES[0], CASES[6], CASES[8]], ids=lambda c: c.name)
def test_tma_sweep_time_steps(need_gpu, case):
    """3 steps of the device-resident RK loop through the TMA sweeps vs the oracle (<= 1e-11)."""
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u_ref = S.local_u0()
    dt = float(case.solver["dt"])
    rk = hpo.rk_type_of(case)
    for _ in range(3):
        O.time_step(u_ref, dt, rk)
    sv = Solver.from_case(case)
    sv.set_solution(S.local_u0())
    sv.TimeSteps(3)
    assert sv.tma_launches > 0
    u = sv.get_solution()
    assert rel_linf(S.interior(u), S.interior(u_ref)) <= 1e-11
    sv.close()


def test_odd_row_length_falls_back(need_gpu):
    """P0 = N0 + 6 odd: row strides are not multiples of 16 bytes, the TMA cannot address the array; the cp.async
    sweep must serve the call with the same parity."""
    case = cases.ns3d_turbulence((21, 14, 12), "mapped")
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u_ref = S.local_u0()
    rhs_ref, hyp_ref, par_ref, _ = O.rhs(u_ref, parts=True)
    sv = Solver.from_case(case)
    u = S.local_u0()
    rhs = sv.RHSFunction(u)
    assert sv.tma_launches == 0 and sv.kernel_launches > 0
    scale = max(np.abs(hyp_ref).max(), np.abs(par_ref).max())
    assert np.abs(rhs - rhs_ref).max() <= 1e-12 * scale
    sv.close()
