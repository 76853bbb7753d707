"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Two bars (DESIGN.md "Parity"):
  * exact path (use_fused = 0; generic kernels compiled without FMA contraction): BIT-IDENTICAL to the oracle,
    which is itself bit-identical to the reference CPU build;
  * production path (fused kernels, FMA, division-free weights): relative Linf and L2 <= 1e-12 per RHS
    evaluation in double precision (BASELINE.json), <= 1e-11 after 5 time steps.
Cases cover the five configurations of BASELINE.json at sizes the oracle finishes in seconds, and
the reference's edge cases: every WENO weight type, characteristic and component-wise
reconstruction, Roe and Rusanov, periodic / extrapolate / slip-wall boundaries, gravity,
viscous terms, non-cubic grids (axis mix-ups), no_limiting.
"""
import numpy as np
import pytest

from conftest import RHS_TOL, assert_close, assert_exact, rel_l2, rel_linf
from hypar_b200 import cases
from hypar_b200.solver import Solver
from oracle import hpo

pytestmark = pytest.mark.gpu


def _cases():
    C = []
    for w in ("js", "mapped", "z", "yc"):
        C.append(cases.linear_advection_sine(64, w))
    C.append(cases.linear_advection_sine(1024, "mapped"))                       # C1 at full size
    C.append(cases.linear_advection_sine(80, "js", diffusion=0.01, par_scheme="2"))
    C.append(cases.linear_advection_sine(80, "z", diffusion=0.02, par_scheme="4"))
    for w in ("js", "mapped", "z", "yc"):
        C.append(cases.euler1d_sod(101, w))                                     # C2: char + Roe
    C.append(cases.euler1d_sod(201, "js"))
    C.append(cases.euler1d_sod(101, "z", interp="components", upwinding="rusanov"))
    C.append(cases.euler1d_sod(101, "js", interp="components", upwinding="roe"))
    C.append(cases.euler1d_sod(101, "mapped", interp="characteristic", upwinding="rusanov"))
    for w in ("js", "mapped", "z", "yc"):
        C.append(cases.ns2d_vortex((40, 28), w))                                # C3
    for w in ("js", "mapped", "z", "yc"):
        C.append(cases.ns3d_turbulence((20, 14, 12), w))                        # C4: viscous
    C.append(cases.ns3d_turbulence((16, 12, 10), "js", viscous=False, upwinding="roe"))
    C.append(cases.ns3d_turbulence((12, 14, 10), "z", viscous=False, interp="characteristic"))
    C.append(cases.ns3d_turbulence((12, 10, 14), "mapped", viscous=True, interp="characteristic", upwinding="roe"))
    C.append(cases.ns3d_density_wave((16, 12, 10), "js"))                       # C5a
    C.append(cases.ns3d_rising_bubble((12, 16, 10), "yc"))                      # C5b: slip walls + gravity
    C.append(cases.ns3d_rising_bubble((10, 14, 12), "mapped", hb=1))
    # characteristic-based Roe-fixed / local Lax-Friedrichs upwinding; appended so that the
    # indices used below (STEP_CASES ...) stay put
    nl = cases.ns3d_density_wave((12, 10, 8), "js")
    nl.weno["no_limiting"] = 1
    nl.name += "_nolimiting"
    C.append(nl)
    C.append(cases.euler1d_sod(101, "js", upwinding="rf-char"))
    C.append(cases.euler1d_sod(101, "z", interp="components", upwinding="llf-char"))
    C.append(cases.ns3d_turbulence((12, 10, 14), "mapped", viscous=False, upwinding="rf-char"))
    C.append(cases.ns3d_turbulence((10, 12, 8), "yc", viscous=True, interp="characteristic", upwinding="llf-char"))
    # compact schemes (tridiagonal solve per line and component) and fifth-order upwind
    C.append(cases.linear_advection_sine(96, "mapped", scheme="crweno5"))
    C.append(cases.euler1d_sod(101, "js", interp="components", upwinding="roe", scheme="crweno5"))
    C.append(cases.euler1d_sod(151, "yc", interp="components", upwinding="rusanov", scheme="crweno5"))
    C.append(cases.ns2d_vortex((40, 28), "z", scheme="crweno5"))
    C.append(cases.ns3d_turbulence((20, 14, 12), "mapped", scheme="crweno5"))
    C.append(cases.ns3d_turbulence((12, 14, 10), "js", viscous=False, upwinding="roe", scheme="crweno5"))
    C.append(cases.ns3d_rising_bubble((12, 16, 10), "yc", scheme="crweno5"))
    nlc = cases.ns3d_density_wave((12, 10, 8), "js", scheme="crweno5")
    nlc.weno["no_limiting"] = 1
    nlc.name += "_nolimiting"
    C.append(nlc)
    C.append(cases.ns2d_vortex((28, 40), "js", scheme="cupw5"))
    C.append(cases.ns3d_rising_bubble((10, 14, 12), "js", scheme="cupw5"))
    C.append(cases.euler1d_sod(101, "js", interp="components", upwinding="llf-char", scheme="upw5"))
    C.append(cases.ns3d_turbulence((14, 10, 12), "js", scheme="upw5"))
    # the other explicit RK tableaux and forward Euler
    C.append(cases.with_time_scheme(cases.linear_advection_sine(96, "js"), "rk", "1fe"))
    C.append(cases.with_time_scheme(cases.euler1d_sod(101, "mapped"), "rk", "22"))
    C.append(cases.with_time_scheme(cases.ns2d_vortex((40, 28), "yc"), "rk", "33"))
    C.append(cases.with_time_scheme(cases.ns3d_rising_bubble((12, 16, 10), "z"), "rk", "tvdrk3"))
    C.append(cases.with_time_scheme(cases.ns3d_turbulence((16, 12, 10), "mapped"), "euler"))
    # NavierStokes2D Roe / rf-char / llf-char upwinding and characteristic reconstruction
    C.append(cases.ns2d_vortex((40, 28), "yc", upwinding="roe"))
    C.append(cases.ns2d_vortex((28, 40), "js", upwinding="rf-char"))
    C.append(cases.ns2d_vortex((32, 24), "mapped", upwinding="llf-char", interp="characteristic"))
    C.append(cases.ns2d_vortex((24, 32), "z", upwinding="rusanov", interp="characteristic"))
    C.append(cases.ns2d_vortex((24, 20), "js", upwinding="roe", scheme="crweno5"))
    # NavierStokes2D with gravity (exact kernels only)
    C.append(cases.ns2d_rising_bubble((28, 24), "mapped"))
    C.append(cases.ns2d_rising_bubble((24, 28), "z", hb=1, upwinding="roe"))
    hb3 = cases.ns2d_rising_bubble((24, 24), "js", hb=3, upwinding="llf-char")
    hb3.physics["N_bv"] = 0.01
    hb3.name += "_nbv"
    C.append(hb3)
    C.append(cases.ns2d_rising_bubble((20, 24), "yc", scheme="crweno5"))
    # inflow / outflow / wall / Dirichlet boundary zones
    C.append(cases.ns_channel((32, 24), "mapped"))
    C.append(cases.ns_channel((24, 28), "js", bcs="sup", mach=1.6))
    C.append(cases.ns_channel((28, 24), "z", bcs="amb2", upwinding="roe"))
    C.append(cases.ns_channel((24, 24), "yc", viscous=True))
    C.append(cases.ns_channel((16, 12, 14), "js"))
    C.append(cases.ns_channel((12, 14, 12), "mapped", bcs="sup3", mach=1.4))
    C.append(cases.ns_channel((14, 12, 12), "z", viscous=True, bcs="amb3"))
    # the low-order linear schemes "1", "2", "4"
    C.append(cases.linear_advection_sine(96, "js", diffusion=0.01, scheme="1"))
    C.append(cases.ns2d_vortex((28, 24), "js", scheme="2"))
    C.append(cases.ns3d_rising_bubble((12, 14, 10), "js", scheme="4"))
    C.append(cases.ns_channel((16, 12, 14), "js", scheme="1"))
    # MUSCL reconstructions
    C.append(cases.with_muscl(cases.euler1d_sod(101, "js", interp="components", upwinding="rusanov"), "muscl2", limiter="vanleer"))
    C.append(cases.with_muscl(cases.ns2d_vortex((28, 24), "js", upwinding="roe"), "muscl2", limiter="superbee"))
    C.append(cases.with_muscl(cases.ns2d_vortex((24, 28), "js"), "muscl2", limiter="minmod"))
    C.append(cases.with_muscl(cases.ns3d_rising_bubble((12, 14, 10), "js"), "muscl3", epsilon=1e-6))
    C.append(cases.with_muscl(cases.ns_channel((16, 12, 14), "js"), "muscl2"))
    # Euler1D with gravity
    C.append(cases.euler1d_sod(101, "js", gravity=1.0))
    C.append(cases.euler1d_sod(101, "mapped", interp="components", upwinding="llf-char", gravity=1.0))
    C.append(cases.euler1d_sod(101, "z", upwinding="roe", gravity=1.0, gravity_type=1))
    C.append(cases.euler1d_sod(101, "yc", interp="components", upwinding="llf-char", gravity=0.5, scheme="crweno5"))
    # gravity source reconstructed characteristic-wise
    C.append(cases.with_characteristic(cases.ns3d_rising_bubble((12, 14, 10), "js")))
    C.append(cases.with_characteristic(cases.ns2d_rising_bubble((24, 20), "mapped", upwinding="roe")))
    # the linear / MUSCL schemes characteristic-wise
    C.append(cases.euler1d_sod(101, "js", scheme="upw5"))
    # (central scheme: a smooth flow -- on the Sod tube it produces NaNs within a few steps, in the reference too)
    C.append(cases.ns2d_vortex((28, 24), "js", upwinding="llf-char", interp="characteristic", scheme="2"))
    C.append(cases.with_muscl(cases.euler1d_sod(101, "js", gravity=1.0), "muscl3"))
    C.append(cases.with_muscl(cases.ns2d_vortex((24, 28), "js", upwinding="rf-char", interp="characteristic"), "muscl2", limiter="vanleer"))
    C.append(cases.with_characteristic(cases.ns3d_turbulence((12, 14, 10), "js", viscous=False, upwinding="roe", scheme="4")))
    C.append(cases.with_characteristic(cases.ns_channel((14, 12, 12), "js", scheme="upw5")))
    # LinearADR in 2-D / 3-D
    C.append(cases.linear_advection_nd((32, 24), "mapped"))
    C.append(cases.linear_advection_nd((24, 28), "z", diffusion=[0.01, 0.02], par_scheme="4"))
    C.append(cases.linear_advection_nd((16, 12, 14), "js", diffusion=[0.01, 0.0, 0.02]))
    C.append(cases.linear_advection_nd((12, 14, 10), "yc", scheme="crweno5", advection=[-1.0, 0.5, 0.3]))
    C.append(cases.linear_advection_nd((33, 24), "js"))          # odd row length
    # sponge zones
    C.append(cases.with_sponge(cases.linear_advection_sine(96, "js"), 0, 1, 0.6, 0.9, [0.1]))
    C.append(cases.with_sponge(cases.linear_advection_nd((24, 20), "mapped"), 1, -1, 0.1, 0.5, [0.5]))
    C.append(cases.with_sponge(cases.ns_channel((28, 24), "js"), 0, 1, 0.7, 1.0, [1.0, 0.5, 0.0, 2.0]))
    C.append(cases.with_sponge(cases.ns3d_rising_bubble((12, 14, 10), "yc"), 1, 1, 700.0, 1000.0, [1.0, 0.0, 0.0, 0.0, 2.0e5]))
    C.append(cases.with_sponge(cases.ns3d_turbulence((16, 12, 14), "mapped"), 2, -1, 0.0, 3.0, [1.0, 0.1, 0.0, 0.0, 1.8]))
    # compact schemes on characteristic variables (block tridiagonal systems)
    C.append(cases.euler1d_sod(101, "js", scheme="crweno5"))
    C.append(cases.euler1d_sod(101, "mapped", scheme="crweno5", upwinding="llf-char", gravity=1.0))
    C.append(cases.ns2d_vortex((28, 24), "z", upwinding="roe", interp="characteristic", scheme="crweno5"))
    C.append(cases.ns2d_vortex((24, 28), "js", upwinding="rf-char", interp="characteristic", scheme="cupw5"))
    C.append(cases.with_characteristic(cases.ns3d_turbulence((14, 12, 10), "mapped", viscous=False, upwinding="roe", scheme="crweno5")))
    C.append(cases.with_characteristic(cases.ns_channel((14, 12, 12), "js", scheme="cupw5")))
    # inviscid Burgers equation
    C.append(cases.burgers_nd((96,), "mapped"))
    C.append(cases.burgers_nd((28, 24), "z"))
    C.append(cases.burgers_nd((14, 12, 16), "js"))
    C.append(cases.burgers_nd((24, 28), "yc", scheme="crweno5"))
    C.append(cases.with_muscl(cases.burgers_nd((80,), "js"), "muscl3"))
    # LinearADR with a spatially varying advection field (sign changes: all three branches of LinearADRUpwind.c:56-82)
    C.append(cases.linear_advection_varying((96,), "js"))
    C.append(cases.linear_advection_varying((28, 24), "z"))
    C.append(cases.linear_advection_varying((24, 28), "mapped", periodic=False))
    C.append(cases.linear_advection_varying((14, 12, 16), "yc"))
    C.append(cases.linear_advection_varying((24, 28), "js", scheme="crweno5"))
    C.append(cases.linear_advection_varying((80,), "js", scheme="muscl3", periodic=False))
    # quasi-1-D grids: 3 cells = the number of ghost layers along one dimension (Examples/2D/NavierStokes2D/1DHydrostaticBalance,
    # 1DSodShockTubeWithGravity, 3D/NavierStokes3D/2D_RisingThermalBubble): ghosts filled from ghosts-wide interiors
    C.append(cases.ns2d_vortex((32, 3), "js"))
    C.append(cases.ns2d_rising_bubble((3, 28), "yc"))
    C.append(cases.ns2d_rising_bubble((24, 3), "z", hb=1, upwinding="roe"))
    C.append(cases.ns3d_rising_bubble((12, 16, 3), "mapped"))
    C.append(cases.ns3d_turbulence((3, 14, 12), "js"))
    C.append(cases.ns_channel((28, 3), "js"))
    C.append(cases.ns2d_vortex((3, 24), "mapped", scheme="crweno5"))
    C.append(cases.linear_advection_nd((3, 20), "js", diffusion=[0.01, 0.02]))
    # hybrid compact-WENO5 (Interp1PrimFifthOrderHCWENO.c / ...HCWENOChar.c): component-wise and characteristic
    C.append(cases.linear_advection_sine(96, "z", scheme="hcweno5"))
    C.append(cases.euler1d_sod(151, "mapped+rc0.5", interp="components", upwinding="rusanov", scheme="hcweno5"))
    C.append(cases.euler1d_sod(101, "js", scheme="hcweno5"))
    C.append(cases.euler1d_sod(101, "yc+rc0.5+xi0.01", upwinding="llf-char", gravity=1.0, scheme="hcweno5"))
    C.append(cases.ns2d_vortex((28, 40), "js", scheme="hcweno5"))
    C.append(cases.ns2d_vortex((24, 20), "z", upwinding="roe", interp="characteristic", scheme="hcweno5"))
    C.append(cases.ns3d_turbulence((14, 10, 12), "z", scheme="hcweno5"))
    C.append(cases.ns3d_rising_bubble((10, 14, 12), "mapped+rc0.2", scheme="hcweno5"))
    C.append(cases.with_characteristic(cases.ns3d_turbulence((12, 10, 14), "mapped", viscous=True, upwinding="roe", scheme="hcweno5")))
    C.append(cases.burgers_nd((20, 24), "yc", scheme="hcweno5"))
    return C


CASES = _cases()


def _uses_roe(case):
    return case.physics.get("upwinding") == "roe"


def fused_tolerance(O, u_ref, dt, ref):
    """Absolute tolerance of the FUSED (FMA, restructured weights) kernels for one RHS term.

    1e-12 x max|ref| (BASELINE.json's relative Linf bound) plus a rounding floor of 16 ulp of the largest
    quantity that is summed to form the term: the Rusanov dissipation alpha * u * dxinv, whose magnitude
    is (CFL/dt) * max|u|. The floor only matters in dimensional units (rising bubble: E = 2.5e5, c = 347 m/s,
    hydrostatic balance): there alpha*u*dxinv = 1.3e6 while the flux divergence is 12, so one ulp of the
    reconstructed energy already moves the result by 2e-11 relative -- the reference's own value carries that
    rounding noise (the exact path below reproduces it bit for bit; no other evaluation order can)."""
    lam_dxinv = O.cfl(u_ref, dt) / dt
    return 1e-12 * np.abs(ref).max() + 16 * np.finfo(np.float64).eps * lam_dxinv * np.abs(u_ref).max()


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_rhs_exact_path_bit_identical(need_gpu, case):
    """use_fused = 0: generic per-interface kernels compiled without FMA contraction. Every operation
    rounds as in the reference's CPU build, so TimeRHSFunctionExplicit and its pieces are BIT-IDENTICAL
    to the oracle (= the reference) -- boundary conditions, hyperbolic, parabolic, source, rhs."""
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u_ref = S.local_u0()
    rhs_ref, hyp_ref, par_ref, src_ref = O.rhs(u_ref, parts=True)
    sv = Solver.from_case(case, use_fused=False)
    u = S.local_u0()
    rhs = sv.RHSFunction(u)
    assert np.array_equal(u, u_ref), "u after boundary conditions"
    hyp = sv.HyperbolicFunction(u)
    par = sv.ParabolicFunction(u)
    src = sv.SourceFunction(u)
    for name, a, b in (("HyperbolicFunction", hyp, hyp_ref), ("ParabolicFunction", par, par_ref),
                       ("SourceFunction", src, src_ref), ("RHSFunction", rhs, rhs_ref)):
        assert np.isfinite(a).all(), name
        assert np.array_equal(a, b), f"{name}: not bit-identical, max abs diff {np.abs(a - b).max():.3e} " \
                                     f"(rel {rel_linf(a, b):.3e})"
    assert sv.kernel_launches > 0
    sv.close()


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_rhs_parity(need_gpu, case):
    """Production path (fused sweeps where the configuration has them): one TimeRHSFunctionExplicit,
    <= 1e-12 relative (fused_tolerance)."""
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u_ref = S.local_u0()
    rhs_ref, hyp_ref, par_ref, src_ref = O.rhs(u_ref, parts=True)
    dt = float(case.solver["dt"])

    sv = Solver.from_case(case, use_fused=True)
    u = S.local_u0()
    rhs = sv.RHSFunction(u)
    assert np.array_equal(u, u_ref), "u after boundary conditions"
    hyp = sv.HyperbolicFunction(u)
    assert np.isfinite(hyp).all() and np.isfinite(rhs).all()
    tol = fused_tolerance(O, u_ref, dt, hyp_ref)
    assert np.abs(hyp - hyp_ref).max() <= tol, \
        f"HyperbolicFunction: abs err {np.abs(hyp - hyp_ref).max():.3e} > {tol:.3e} (rel {rel_linf(hyp, hyp_ref):.3e})"
    assert rel_l2(hyp, hyp_ref) <= max(RHS_TOL, tol / np.abs(hyp_ref).max())
    # rhs = -hyp + par + source may cancel strongly (hydrostatic balance): measured against the terms summed
    scale = max(np.abs(hyp_ref).max(), np.abs(par_ref).max(), np.abs(src_ref).max())
    tol = fused_tolerance(O, u_ref, dt, np.array([scale]))
    assert np.abs(rhs - rhs_ref).max() <= tol, f"rhs: abs err {np.abs(rhs - rhs_ref).max():.3e} > {tol:.3e}"
    # the viscous / source contributions alone: rhs + hyp = par + source
    if np.abs(par_ref).max() > 0 or np.abs(src_ref).max() > 0:
        ps, ps_ref = rhs + hyp, rhs_ref + hyp_ref
        assert np.abs(ps - ps_ref).max() <= 2 * tol
    assert sv.kernel_launches > 0
    sv.close()


STEP_CASES = [CASES[0], CASES[7], CASES[15], CASES[19], CASES[25], CASES[26], CASES[30], CASES[32],
              CASES[34], CASES[36], CASES[37], CASES[38], CASES[40], CASES[41], CASES[43], CASES[45],
              CASES[46], CASES[47], CASES[48], CASES[49], CASES[50], CASES[51], CASES[53],
              CASES[56], CASES[57], CASES[58], CASES[59], CASES[60], CASES[61], CASES[62], CASES[63], CASES[64],
              CASES[65], CASES[66], CASES[67], CASES[68], CASES[69], CASES[70],
              CASES[71], CASES[72], CASES[73], CASES[74], CASES[75], CASES[76], CASES[77], CASES[78], CASES[79],
              CASES[80], CASES[81], CASES[82], CASES[83], CASES[84], CASES[85], CASES[86], CASES[87],
              CASES[88], CASES[89], CASES[90], CASES[91], CASES[92], CASES[93], CASES[94], CASES[95], CASES[96], CASES[97],
              CASES[98], CASES[99], CASES[100], CASES[101], CASES[102], CASES[103],
              CASES[104], CASES[105], CASES[106], CASES[107], CASES[108],
              CASES[109], CASES[110], CASES[111], CASES[112], CASES[113], CASES[114],
              CASES[115], CASES[116], CASES[117], CASES[118], CASES[119], CASES[120], CASES[121], CASES[122]]


@pytest.mark.parametrize("case", STEP_CASES, ids=lambda c: c.name)
def test_time_steps_exact_path_bit_identical(need_gpu, case):
    """TimeRK (RK4 / SSPRK3) over 5 steps with the exact path: the device-resident loop, the host-array
    TimeIntegrate and the reference agree to the last bit (Roe cases included)."""
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u_ref = S.local_u0()
    dt = float(case.solver["dt"])
    rk = hpo.rk_type_of(case)
    for _ in range(5):
        O.time_step(u_ref, dt, rk)
    sv = Solver.from_case(case, use_fused=False)
    sv.set_solution(S.local_u0())
    sv.TimeSteps(5)
    u = sv.get_solution()
    # viscous channel cases: temperatures on which CUDA's and glibc's exp / log differ by an ulp (conftest.assert_exact)
    ulp = case.name.startswith("chan") and float(case.physics.get("Re", -1.0)) > 0
    assert_exact(S.interior(u), S.interior(u_ref), "u after 5 steps", libm_ulp=ulp)
    u2 = S.local_u0()
    sv.TimeIntegrate(u2, 5)
    assert np.array_equal(u, u2), "host-array TimeIntegrate differs from the device-resident loop"
    assert abs(sv.time - 5 * dt) <= 1e-12 * max(1.0, 5 * dt)      # TimeIntegrate restarts the clock at t0
    sv.close()


@pytest.mark.parametrize("case", STEP_CASES, ids=lambda c: c.name)
def test_time_steps_parity(need_gpu, case):
    """Production path over 5 steps. Documented final-time agreement: relative Linf/L2 <= 1e-11 (the per-RHS
    1e-12 bound accumulated over 15-20 RHS evaluations); the step norm and CFL reductions agree likewise."""
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u_ref = S.local_u0()
    dt = float(case.solver["dt"])
    rk = hpo.rk_type_of(case)
    for _ in range(5):
        u_prev = u_ref.copy()
        O.time_step(u_ref, dt, rk)
    sv = Solver.from_case(case)
    sv.set_solution(S.local_u0())
    sv.TimeSteps(5)
    u = sv.get_solution()
    assert_close(S.interior(u), S.interior(u_ref), 1e-11, "u after 5 steps (device loop)")
    u2 = S.local_u0()
    sv.TimeIntegrate(u2, 5)
    assert np.array_equal(u, u2), "host-array TimeIntegrate differs from the device-resident loop"
    # TimePostStep.c:44-63 norm of the last step (local sum of squares) and TimePreStep.c CFL
    ss_ref = float(((S.interior(u_ref) - S.interior(u_prev)) ** 2).sum())
    ss = sv.dev_StepNormSumSq()
    assert abs(ss - ss_ref) <= 1e-9 * ss_ref + 1e-300
    cfl_ref = O.cfl(u_ref, dt)
    assert abs(sv.dev_ComputeCFL() - cfl_ref) <= 1e-10 * cfl_ref
    sv.close()


@pytest.mark.parametrize("case", [CASES[4], CASES[12], CASES[16], CASES[20], CASES[26],
                                  CASES[35], CASES[37], CASES[40], CASES[42], CASES[44], CASES[51], CASES[52], CASES[53],
                                  CASES[56], CASES[58], CASES[60], CASES[65], CASES[69], CASES[72], CASES[74], CASES[76], CASES[79], CASES[82], CASES[85], CASES[87], CASES[98], CASES[100], CASES[102], CASES[104], CASES[105], CASES[109], CASES[110], CASES[111], CASES[115], CASES[118]],
                         ids=lambda c: c.name)
def test_function_pointer_pieces(need_gpu, case):
    """FFunction, UFunction, SetInterpLimiterVar, InterpolateInterfacesHyp, Upwind,
    FirstDerivativePar, SecondDerivativePar, ComputeCFL -- one by one against the oracle."""
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    sv = Solver.from_case(case)
    u = S.local_u0()
    O.apply_bc(u)
    u2 = S.local_u0()
    sv.ApplyBoundaryConditions(u2)
    assert np.array_equal(u, u2), "ApplyBoundaryConditions"
    assert sv.ComputeCFL(u) == O.cfl(u, float(case.solver["dt"])), "ComputeCFL"
    for d in range(S.ndims):
        f_ref = O.flux(u, d)
        f = sv.FFunction(u, d)
        # corners of the ghost-padded array are never filled (zeros -> 0/0): compare where finite in the oracle
        m = np.isfinite(f_ref)
        assert np.array_equal(f[m], f_ref[m]), f"FFunction dir {d}"
        uc_ref = O.modified_solution(u)
        uc = sv.UFunction(u, d)
        m = np.isfinite(uc_ref)
        assert np.array_equal(uc[m], uc_ref[m]), "UFunction"
        f_in = np.where(np.isfinite(f_ref), f_ref, 0.0)
        uc_in = np.where(np.isfinite(uc_ref), uc_ref, 0.0)
        w_ref = O.weno_weights(f_in, u, d)
        sv.SetInterpLimiterVar(f_in, u, d)
        if case.solver["hyp_space_scheme"] in ("weno5", "crweno5", "hcweno5"):     # the linear schemes keep no weights
            w = sv.GetInterpWeights(d)
            assert np.array_equal(w, w_ref), f"weights dir {d}: {np.abs(w - w_ref).max():.3e}"
        outs = {}
        for name, arr, upw, uflag in (("uL", uc_in, 1, 1), ("uR", uc_in, -1, 1), ("fL", f_in, 1, 0), ("fR", f_in, -1, 0)):
            ref = O.interp(arr, u, w_ref, upw, d, uflag)
            got = sv.InterpolateInterfacesHyp(arr, u, upw, d, uflag)
            assert np.array_equal(got, ref), f"InterpolateInterfacesHyp {name} dir {d}: {np.abs(got - ref).max():.3e}"
            outs[name] = ref
        fi_ref = O.upwind(outs["fL"], outs["fR"], outs["uL"], outs["uR"], u, d)
        fi = sv.Upwind(outs["fL"], outs["fR"], outs["uL"], outs["uR"], u, d)
        assert np.array_equal(fi, fi_ref), f"Upwind dir {d}: {np.abs(fi - fi_ref).max():.3e}"
        # derivative operators applied to a smooth, finite field
        rng = np.random.RandomState(7 + d)
        fld = rng.standard_normal(u.shape)
        d1_ref = O.first_derivative(fld, d)
        d1 = sv.FirstDerivativePar(fld, d)
        assert np.array_equal(d1, d1_ref), f"FirstDerivativePar dir {d}"
        order = int(case.solver["par_space_scheme"])
        d2_ref = O.second_derivative(fld, d, order)
        d2 = sv.SecondDerivativePar(fld, d)
        assert np.array_equal(d2, d2_ref), f"SecondDerivativePar dir {d}"
    sv.close()
