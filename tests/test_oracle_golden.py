"""The CPU oracle (oracle/hypar_oracle.c) pinned against golden vectors produced by the UNMODIFIED
reference (tests/golden/*.npz, generator: tools/make_golden.py), and -- where oracle/_ref is present --
against the reference executable run live on further cases.

The reference's own tests hold no value-pinning vectors for this path: its
WENO5 test only bounds the output and the baselines of its regression suite live in external
repositories. The pin is therefore the reference executable itself; the comparison is BIT-EXACT
(the oracle restates the same arithmetic in the same order, compiled by the same gcc -O3).
"""
import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from hypar_b200 import cases  # noqa: E402
from oracle import hpo  # noqa: E402

GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def load(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in meta["kwargs"].items()}
    return z, getattr(cases, meta["builder"])(**kw), meta


def same(a, b, what):
    """bit-exact where both are finite; NaNs (never-filled corner ghosts, 0/0) must coincide"""
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    assert a.shape == b.shape, what
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN pattern differs"
    assert np.array_equal(a[~na], b[~nb]), f"{what}: max abs diff {np.abs(a[~na] - b[~nb]).max():.3e}"


def test_golden_present():
    assert len(GOLDEN) >= 45


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    z, case, meta = load(path)
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    same(S.x, z["rhs_x"], "x with ghosts")
    same(S.dxinv, z["rhs_dxinv"], "dxinv")
    u = S.local_u0()
    rhs, hyp, par, src = O.rhs(u, parts=True)
    same(u, z["rhs_u"], "u after boundary conditions")
    same(hyp, z["rhs_hyp"], "hyp")
    same(par, z["rhs_par"], "par")
    same(src, z["rhs_source"], "source")
    same(rhs, z["rhs_rhs"], "rhs")
    # three steps of the reference's own time loop
    u = S.local_u0()
    rk = hpo.rk_type_of(case)
    for _ in range(3):
        O.time_step(u, float(case.solver["dt"]), rk)
    same(S.interior(u), S.interior(z["steps3_u"]), "u after 3 steps")
    # conservation diagnostics of the same three steps (TimePostStep.c:81-93)
    u = S.local_u0()
    vol0 = O.volume_integral(u)
    same(vol0, z["cons_vol0"], "VolumeIntegralInitial")
    tbi = np.zeros(S.nvars)
    for k in range(3):
        sbi = O.time_step_cons(u, float(case.solver["dt"]), rk)
        same(sbi, z["cons_stepbi"][k], f"StepBoundaryIntegral, step {k + 1}")
        tbi = tbi + O.boundary_integral(sbi)
        vol = O.volume_integral(u)
        same(np.concatenate([vol, tbi, O.conservation_error(vol, vol0, tbi)]), z["cons_steps"][k],
             f"VolumeIntegral | TotalBoundaryIntegral | ConservationError, step {k + 1}")
    same(S.interior(u), S.interior(z["steps3_u"]), "u after 3 steps (conservation bookkeeping on)")
    if "pieces_u" not in z:
        return
    # the function-pointer pieces, per direction
    u = S.local_u0()
    O.apply_bc(u)
    same(u, z["pieces_u"], "pieces: u")
    assert O.cfl(u, float(case.solver["dt"])) == float(z["pieces_cfl"][0])
    for d in range(S.ndims):
        f = O.flux(u, d)
        same(f, z[f"pieces_fluxC_{d}"], f"FFunction {d}")
        uc = O.modified_solution(u)
        same(uc, z[f"pieces_uC_{d}"], f"UFunction {d}")
        w = O.weno_weights(f, u, d)
        if f"pieces_weights_{d}" in z:          # the linear schemes (cupw5, upw5) have no SetInterpLimiterVar
            same(w, z[f"pieces_weights_{d}"], f"WENO weights {d}")
        outs = {}
        for name, arr, upw, uflag in (("uL", uc, 1, 1), ("uR", uc, -1, 1), ("fL", f, 1, 0), ("fR", f, -1, 0)):
            outs[name] = O.interp(arr, u, w, upw, d, uflag)
            same(outs[name], z[f"pieces_{name}_{d}"], f"{name} {d}")
        fi = O.upwind(outs["fL"], outs["fR"], outs["uL"], outs["uR"], u, d)
        same(fi, z[f"pieces_fluxI_{d}"], f"Upwind {d}")
        if f"pieces_D1_{d}" in z:
            same(O.first_derivative(u, d), z[f"pieces_D1_{d}"], f"FirstDerivativePar {d}")
        if f"pieces_D2_{d}" in z:
            same(O.second_derivative(u, d, int(case.solver["par_space_scheme"])), z[f"pieces_D2_{d}"],
                 f"SecondDerivativePar {d}")


# ---- live reference (only where oracle/_ref was built: this container, or the GPU box via gpurun)
def _hb3(case, vertical=1):
    """stratified atmosphere (HB 3 <N_bv>): gravity along the vertical coordinate only"""
    case.physics["N_bv"] = 0.01
    g = [0.0] * case.ndims
    g[vertical] = 9.8
    case.physics["gravity"] = g
    case.name += "_hb3"
    return case


def _live_cases():
    return [
        cases.linear_advection_sine(96, "z"),
        cases.linear_advection_sine(80, "yc", diffusion=0.02, par_scheme="2"),
        cases.euler1d_sod(151, "yc"),
        cases.euler1d_sod(101, "js", interp="components", upwinding="roe"),
        cases.euler1d_sod(101, "mapped", interp="characteristic", upwinding="rusanov"),
        cases.ns2d_vortex((24, 40), "z"),
        cases.ns3d_turbulence((20, 14, 12), "yc"),
        cases.ns3d_turbulence((12, 10, 14), "mapped", viscous=True, interp="characteristic", upwinding="roe"),
        cases.ns3d_rising_bubble((10, 14, 12), "z", hb=3) if False else cases.ns3d_rising_bubble((10, 14, 12), "z"),
        cases.euler1d_sod(101, "mapped", upwinding="rf-char"),
        cases.euler1d_sod(101, "js", interp="components", upwinding="llf-char"),
        cases.ns3d_turbulence((12, 10, 14), "js", viscous=False, upwinding="llf-char"),
        cases.ns3d_turbulence((10, 12, 8), "z", viscous=True, interp="characteristic", upwinding="rf-char"),
        # compact schemes / fifth-order upwind
        cases.linear_advection_sine(96, "yc", scheme="crweno5"),
        cases.euler1d_sod(151, "mapped", interp="components", upwinding="rusanov", scheme="crweno5"),
        cases.ns2d_vortex((24, 40), "js", scheme="crweno5"),
        cases.ns3d_turbulence((14, 10, 12), "z", scheme="crweno5"),
        cases.ns3d_rising_bubble((10, 14, 12), "mapped", scheme="crweno5"),
        cases.ns3d_rising_bubble((10, 14, 12), "js", scheme="cupw5"),
        cases.euler1d_sod(101, "js", interp="components", upwinding="llf-char", scheme="cupw5"),
        cases.ns3d_turbulence((12, 10, 14), "js", viscous=False, upwinding="roe", scheme="upw5"),
        # NavierStokes2D Roe / rf-char / llf-char upwinding and characteristic reconstruction
        cases.ns2d_vortex((24, 40), "yc", upwinding="roe"),
        cases.ns2d_vortex((24, 20), "js", upwinding="rf-char"),
        cases.ns2d_vortex((20, 24), "mapped", upwinding="llf-char", interp="characteristic"),
        cases.ns2d_vortex((24, 20), "z", upwinding="rusanov", interp="characteristic"),
        cases.ns2d_vortex((24, 20), "js", upwinding="roe", scheme="crweno5"),
        # NavierStokes2D with gravity (well-balanced source term; HB 1, 2, 3)
        cases.ns2d_rising_bubble((20, 24), "mapped"),
        cases.ns2d_rising_bubble((24, 20), "z", hb=1, upwinding="roe"),
        _hb3(cases.ns2d_rising_bubble((20, 24), "js", hb=3, upwinding="llf-char")),
        _hb3(cases.ns3d_rising_bubble((10, 12, 14), "js", hb=3), vertical=2),
        cases.ns2d_rising_bubble((20, 20), "yc", upwinding="llf-char", interp="characteristic"),
        # inflow / outflow / wall / Dirichlet boundary zones (BCNoslipWall.c, BCDirichlet.c, BCSub*/BCSuper*.c)
        cases.ns_channel((24, 20), "mapped", upwinding="roe"),
        cases.ns_channel((20, 24), "js", bcs="sup", mach=1.6),
        cases.ns_channel((24, 20), "z", bcs="amb2", upwinding="llf-char"),
        cases.ns_channel((20, 20), "yc", viscous=True),
        cases.ns_channel((12, 10, 14), "js"),
        cases.ns_channel((10, 12, 10), "mapped", bcs="sup3", mach=1.4),
        cases.ns_channel((12, 10, 10), "z", viscous=True, bcs="amb3"),
        # the low-order linear schemes "1", "2", "4"
        cases.linear_advection_sine(96, "js", diffusion=0.01, scheme="1"),
        cases.euler1d_sod(101, "js", interp="components", upwinding="rusanov", scheme="1"),
        cases.ns2d_vortex((24, 20), "js", scheme="2"),
        cases.ns3d_rising_bubble((10, 14, 12), "js", scheme="4"),
        cases.ns_channel((12, 10, 14), "js", viscous=True, scheme="1"),
        # MUSCL reconstructions (muscl.inp: limiter of muscl2, epsilon of muscl3)
        cases.with_muscl(cases.linear_advection_sine(96, "js"), "muscl3"),
        cases.with_muscl(cases.euler1d_sod(101, "js", interp="components", upwinding="rusanov"), "muscl2", limiter="vanleer"),
        cases.with_muscl(cases.ns2d_vortex((24, 20), "js", upwinding="roe"), "muscl2", limiter="superbee"),
        cases.with_muscl(cases.ns2d_vortex((20, 24), "js"), "muscl2", limiter="minmod"),
        cases.with_muscl(cases.ns3d_rising_bubble((10, 14, 12), "js"), "muscl3", epsilon=1e-6),
        cases.with_muscl(cases.ns_channel((12, 10, 14), "js", viscous=True), "muscl2"),
        # gravity source reconstructed characteristic-wise (the source function goes through the same
        # InterpolateInterfacesHyp as the flux: NavierStokes3DSource.c:77-78)
        cases.with_characteristic(cases.ns3d_rising_bubble((10, 14, 12), "js")),
        cases.with_characteristic(cases.ns2d_rising_bubble((20, 24), "mapped", upwinding="roe")),
        # the linear / MUSCL schemes characteristic-wise (Interp1Prim...Char.c)
        cases.euler1d_sod(101, "js", scheme="upw5"),
        cases.ns2d_vortex((24, 20), "js", upwinding="llf-char", interp="characteristic", scheme="2"),
        cases.with_muscl(cases.euler1d_sod(101, "js", gravity=1.0), "muscl3"),
        cases.with_muscl(cases.ns2d_vortex((20, 24), "js", upwinding="rf-char", interp="characteristic"), "muscl2", limiter="vanleer"),
        cases.with_characteristic(cases.ns3d_turbulence((12, 10, 14), "js", viscous=False, upwinding="roe", scheme="4")),
        cases.with_characteristic(cases.ns_channel((12, 10, 14), "js", viscous=True, scheme="upw5")),
        # LinearADR in 2-D / 3-D (Examples/2D/LinearAdvection, Examples/3D/LinearAdvection)
        cases.linear_advection_nd((24, 20), "mapped"),
        cases.linear_advection_nd((20, 24), "z", diffusion=[0.01, 0.02], par_scheme="4"),
        cases.linear_advection_nd((12, 10, 14), "js", diffusion=[0.01, 0.0, 0.02]),
        cases.linear_advection_nd((10, 12, 10), "yc", scheme="crweno5", advection=[-1.0, 0.5, 0.3]),
        # sponge zones (BCSponge.c through SourceFunction.c)
        cases.with_sponge(cases.linear_advection_sine(96, "js"), 0, 1, 0.6, 0.9, [0.1]),
        cases.with_sponge(cases.linear_advection_nd((24, 20), "mapped"), 1, -1, 0.1, 0.5, [0.5]),
        cases.with_sponge(cases.ns_channel((24, 20), "js"), 0, 1, 0.7, 1.0, [1.0, 0.5, 0.0, 2.0]),
        cases.with_sponge(cases.ns3d_rising_bubble((10, 14, 12), "yc"), 1, 1, 700.0, 1000.0, [1.0, 0.0, 0.0, 0.0, 2.0e5]),
        # compact schemes on characteristic variables (block tridiagonal systems; flows that are not at rest: the
        # reference's un-pivoted block inversion is singular for a fluid at rest)
        cases.euler1d_sod(101, "js", scheme="crweno5"),
        cases.euler1d_sod(101, "mapped", scheme="crweno5", upwinding="llf-char", gravity=1.0),
        cases.ns2d_vortex((24, 20), "z", upwinding="roe", interp="characteristic", scheme="crweno5"),
        cases.ns2d_vortex((20, 24), "js", upwinding="rf-char", interp="characteristic", scheme="cupw5"),
        cases.with_characteristic(cases.ns3d_turbulence((12, 10, 14), "mapped", viscous=True, upwinding="roe", scheme="crweno5")),
        cases.with_characteristic(cases.ns_channel((12, 10, 14), "js", scheme="cupw5")),
        # inviscid Burgers equation (src/PhysicalModels/Burgers)
        cases.burgers_nd((96,), "mapped"),
        cases.burgers_nd((24, 20), "z"),
        cases.burgers_nd((12, 10, 14), "js"),
        cases.burgers_nd((20, 24), "yc", scheme="crweno5"),
        cases.with_muscl(cases.burgers_nd((80,), "js"), "muscl3"),
        # LinearADR with a spatially varying advection field read from a file (LinearADRAdvectionField.c, LinearADRUpwind.c:56-82)
        cases.linear_advection_varying((96,), "js"),
        cases.linear_advection_varying((80,), "mapped", periodic=False, tstype="ssprk3"),
        cases.linear_advection_varying((24, 20), "z"),
        cases.linear_advection_varying((20, 24), "js", periodic=False),
        cases.linear_advection_varying((12, 10, 14), "yc"),
        cases.linear_advection_varying((20, 24), "js", scheme="crweno5"),
        cases.linear_advection_varying((64,), "js", scheme="muscl3"),
        # quasi-1-D grids (3 cells = ghosts along one dimension, as the reference's 1DHydrostaticBalance / 2D_RisingThermalBubble)
        cases.ns2d_vortex((32, 3), "js"),
        cases.ns2d_rising_bubble((3, 28), "yc"),
        cases.ns3d_rising_bubble((12, 16, 3), "mapped"),
        cases.ns3d_turbulence((3, 14, 12), "js"),
        cases.ns_channel((28, 3), "js"),
        # Euler1D with gravity (Euler1DGravityField.c, Euler1DSource.c)
        cases.euler1d_sod(101, "js", gravity=1.0),
        cases.euler1d_sod(101, "mapped", interp="components", upwinding="llf-char", gravity=1.0),
        cases.euler1d_sod(101, "z", upwinding="roe", gravity=1.0, gravity_type=1),
        cases.euler1d_sod(101, "yc", interp="components", upwinding="llf-char", gravity=0.5, scheme="crweno5"),
        # hybrid compact-WENO5 (second half), component-wise and characteristic, default and other rc / xi
        cases.linear_advection_sine(96, "mapped", scheme="hcweno5"),
        cases.euler1d_sod(151, "mapped+rc0.5", interp="components", upwinding="rusanov", scheme="hcweno5"),
        cases.euler1d_sod(101, "js", scheme="hcweno5"),
        cases.ns2d_vortex((24, 40), "js", scheme="hcweno5"),
        cases.ns3d_turbulence((14, 10, 12), "z+xi0.01", scheme="hcweno5"),
        cases.with_characteristic(cases.ns3d_turbulence((12, 10, 14), "mapped", viscous=True, upwinding="roe", scheme="hcweno5")),
        cases.ns3d_rising_bubble((10, 14, 12), "yc", scheme="hcweno5"),
        # the other explicit RK tableaux (TimeExplicitRKInitialize.c:27-79) and forward Euler (TimeForwardEuler.c)
        cases.with_time_scheme(cases.linear_advection_sine(96, "js"), "rk", "1fe"),
        cases.with_time_scheme(cases.euler1d_sod(101, "mapped"), "rk", "22"),
        cases.with_time_scheme(cases.ns2d_vortex((24, 40), "yc"), "rk", "33"),
        cases.with_time_scheme(cases.ns3d_rising_bubble((10, 14, 12), "z"), "rk", "tvdrk3"),
        cases.with_time_scheme(cases.ns3d_turbulence((12, 10, 14), "mapped"), "euler"),
    ]


LIVE = _live_cases()


@pytest.mark.parametrize("case", LIVE, ids=[c.name for c in LIVE])
@pytest.mark.parametrize("exe", ["hypar_ref", "hypar_ref_mpi1"])
def test_oracle_matches_reference_live(case, exe):
    from refrun import ref_available, run_reference
    if not ref_available(exe):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    o = run_reference(case, "rhs", exe=exe)
    S = hpo.Setup(case, mpi_semantics=exe.endswith("mpi1"))
    O = hpo.Oracle(S)
    u = S.local_u0()
    rhs, hyp, par, src = O.rhs(u, parts=True)
    for k, a in (("u", u), ("hyp", hyp), ("par", par), ("source", src), ("rhs", rhs)):
        same(a, o[k]["data"], f"{exe} {k}")
    o = run_reference(case, "steps", [2], exe=exe)
    u = S.local_u0()
    for _ in range(2):
        O.time_step(u, float(case.solver["dt"]), hpo.rk_type_of(case))
    same(S.interior(u), S.interior(o["ufinal"]["data"]), f"{exe} u after 2 steps")


@pytest.mark.parametrize("case", [LIVE[0], LIVE[5], LIVE[8]], ids=lambda c: c.name)
def test_error_norms_match_reference_live(case, tmp_path):
    """CalculateError.c:26-124 (errors.dat): with exact.inp = the initial solution, the reference's L1/L2/Linf
    errors after 2 steps equal the ones formed from the oracle's norm sums (same serial summation order)."""
    import shutil
    from refrun import ref_available, run_reference
    if not ref_available("hypar_ref"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    d = str(tmp_path / "run")
    case.write(d)
    shutil.copy(os.path.join(d, "initial.inp"), os.path.join(d, "exact.inp"))
    o = run_reference(case, "steps", [2], keep=d)
    ref = [float(x) for x in [l for l in o["stdout"].splitlines() if l.startswith("ERRORS")][0].split()[1:]]
    S = hpo.Setup(case, mpi_semantics=False)
    O = hpo.Oracle(S)
    u = S.local_u0()
    for _ in range(2):
        O.time_step(u, float(case.solver["dt"]), hpo.rk_type_of(case))
    uex = S.local_u0()
    npts = float(np.prod(S.dim))
    n, e = O.norm_sums(uex), O.norm_sums(uex, u)
    sol = [n[0] / npts, np.sqrt(n[1] / npts), n[2]]
    err = [e[0] / npts, np.sqrt(e[1] / npts), e[2]]
    if all(v > 1e-15 for v in sol):
        err = [a / b for a, b in zip(err, sol)]
    assert err == ref, f"{err} vs {ref}"


# ---- the reference's own unit tests for this path, re-run against the oracle
def test_first_derivative_polynomial_exactness():
    """tests/FirstDerivative/test_first_derivative.c: the 4th-order central operator differentiates
    polynomials up to degree 4 exactly (interior points), tolerance 1e-10 as in the reference's harness."""
    n = 40
    case = cases.linear_advection_sine(n, "js")
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    g = S.ghosts
    xi = np.arange(-g, n + g, dtype=np.float64)
    for deg in range(5):
        f = np.ascontiguousarray(xi ** deg)
        d = O.first_derivative(f, 0)
        exact = deg * xi ** (deg - 1) if deg > 0 else np.zeros_like(xi)
        assert np.abs(d - exact).max() <= 1e-10 * max(1.0, np.abs(exact).max()), f"degree {deg}"


def test_second_derivative_polynomial_exactness():
    """tests/SecondDerivative/test_second_derivative.c: 2nd order exact to degree 3, 4th order to degree 5."""
    n = 40
    case = cases.linear_advection_sine(n, "js")
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    g = S.ghosts
    xi = np.arange(-g, n + g, dtype=np.float64) / 8.0
    h = 1.0 / 8.0
    for order, maxdeg in ((2, 3), (4, 5)):
        for deg in range(maxdeg + 1):
            f = np.ascontiguousarray(xi ** deg)
            d = O.second_derivative(f, 0, order)[g:g + n] / (h * h)
            exact = (deg * (deg - 1) * xi ** (deg - 2) if deg > 1 else np.zeros_like(xi))[g:g + n]
            assert np.abs(d - exact).max() <= 1e-10 * max(1.0, np.abs(exact).max()), f"order {order} degree {deg}"


def test_weno5_smoke_bounded_like_reference():
    """tests/InterpolationFunctions/test_interpolation.c:310-403: with weights frozen at (0.1,0.6,0.3)
    (no_limiting) the interface values of a smooth function stay bounded (|fI| < 1.5) -- and are the
    5th-order polynomial interpolant."""
    case = cases.linear_advection_sine(64, "js")
    case.weno["no_limiting"] = 1
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    # cell averages of sin(2 pi x) over [x_i - h/2, x_i + h/2]: the scheme reconstructs interface POINT values
    h, g = 1.0 / 64, S.ghosts
    xc = np.arange(64) * h
    u = np.zeros(64 + 2 * g)
    u[g:g + 64] = (np.cos(2 * np.pi * (xc - h / 2)) - np.cos(2 * np.pi * (xc + h / 2))) / (2 * np.pi * h)
    O.apply_bc(u)
    w = O.weno_weights(u, u, 0)
    assert np.allclose(w.reshape(4, 3, -1)[:, 0], 0.1) and np.allclose(w.reshape(4, 3, -1)[:, 1], 0.6)
    fI = O.interp(u, u, w, 1, 0, 0)
    assert np.abs(fI).max() < 1.5
    xh = (np.arange(65) - 0.5) / 64.0
    assert np.abs(fI - np.sin(2 * np.pi * xh)).max() < 1e-6
