"""C ABI: the library loads, exports every symbol include/hypar_b200.h declares, validates
configurations with the reference's error behaviour, refuses to compute without a GPU, and its host
set-up (partition, ghost coordinates, dxinv, neighbours, zone extents, gravity field) agrees with the
independent numpy/C restatement in oracle/hpo.py. No compute calls: runs on a CPU-only box."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hypar_b200 import _lib, cases
from hypar_b200.solver import HyParB200Error, Solver
from oracle import hpo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "hypar_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hpb_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 55
    for name in sorted(declared):
        assert hasattr(L, name), f"libhypar_b200.so does not export {name}"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert b"sm_100a" in L.hpb_version()


def test_config_struct_layout_matches_header():
    """the ctypes mirror must have the C struct's size: hpb_config_defaults writes the whole struct"""
    L = _lib.load()
    c = _lib.Config()
    guard = (C.c_char * 64)()
    L.hpb_config_defaults(C.byref(c))
    assert c.ghosts == 1 and c.gamma == 1.4 and c.weno_eps == 1e-6 and c.use_fused == 1 and c.device == -1
    assert bytes(guard) == b"\0" * 64


def test_partition_and_rank_maps():
    L = _lib.load()
    for ng, npr in ((512, 2), (50, 3), (201, 4), (64, 1)):
        sizes = [L.hpb_partition1d(ng, npr, r) for r in range(npr)]
        assert sum(sizes) == ng and sizes[:-1] == [ng // npr] * (npr - 1)          # MPIPartition1D.c
        assert sizes == [hpo.partition1d(ng, npr, r) for r in range(npr)]
    iproc = (C.c_int * 3)(2, 3, 2)
    for r in range(12):
        ip = (C.c_int * 3)()
        L.hpb_ranknd(3, r, iproc, ip)
        assert list(ip) == hpo.rank_nd([2, 3, 2], r)
        assert L.hpb_rank1d(3, iproc, ip) == r


ALL_CASES = [
    cases.linear_advection_sine(64, "js"),
    cases.euler1d_sod(101, "js"),
    cases.ns2d_vortex((40, 28), "yc"),
    cases.ns3d_turbulence((20, 14, 12), "mapped"),
    cases.ns3d_rising_bubble((12, 16, 10), "yc"),
    cases.ns3d_rising_bubble((10, 14, 12), "mapped", hb=1),
    cases.ns2d_rising_bubble((20, 24), "js"),                      # 2-D gravity field (HB 2) and slip walls
    cases.ns2d_rising_bubble((24, 20), "z", hb=1, upwinding="roe"),
    cases.linear_advection_nd((24, 20), "js"),
    cases.burgers_nd((24, 20), "js"),
    cases.linear_advection_varying((24, 20), "js"),                 # advection field from a file: periodic copies
    cases.linear_advection_varying((12, 10, 14), "z", periodic=False),   # ... and mirrored at open faces
    cases.with_sponge(cases.ns_channel((24, 20), "js"), 0, 1, 0.7, 1.0, [1.0, 0.5, 0.0, 2.0]),    # interior sponge box
    cases.linear_advection_nd((12, 10, 14), "js", diffusion=[0.01, 0.0, 0.02]),
    cases.euler1d_sod(101, "js", gravity=1.0),                      # 1-D gravity field, mirrored at the physical faces
    cases.euler1d_sod(101, "z", gravity=1.0, gravity_type=1),
    cases.ns_channel((24, 20), "js"),                               # inflow / outflow / no-slip / slip zones
    cases.ns_channel((10, 12, 10), "js", bcs={(1, 1): "dirichlet", (1, -1): "subsonic-ambivalent"}),
]


@pytest.mark.parametrize("case", ALL_CASES, ids=[c.name for c in ALL_CASES])
def test_host_setup_matches_oracle_setup(case):
    S = hpo.Setup(case)
    sv = Solver.from_case(case)
    assert sv.dim_local == S.dim and sv.npoints_local_wghosts == S.npoints_g
    x, dxinv = sv.grid()
    assert np.array_equal(x, S.x) and np.array_equal(dxinv, S.dxinv)
    for n, z in enumerate(S.zones):
        a, b, on = sv.zone_extent(n)
        assert on == z["on"]
        if on:
            assert a == z["is"] and b == z["ie"]
    f, g = sv.gravity_field()
    assert np.array_equal(f, S.grav_f) and np.array_equal(g, S.grav_g)
    if S.adv_field is not None:
        assert np.array_equal(sv.advection_field(), S.adv_field)
    assert sv.neighbors == [-1] * (2 * S.ndims)
    assert sv.nstages == (4 if case.solver["time_scheme_type"] == "44" else 3)
    if case is ALL_CASES[0]:
        for ts, tst, ns in (("rk", "1fe", 1), ("rk", "22", 2), ("rk", "33", 3), ("rk", "tvdrk3", 3), ("euler", " ", 1)):
            c2 = cases.with_time_scheme(cases.linear_advection_sine(64, "js"), ts, tst)
            s2 = Solver.from_case(c2)
            assert s2.nstages == ns
            s2.close()
    for d in range(S.ndims):
        assert sv.ninterfaces(d) == S.ninterfaces(d)
    sv.close()


def test_decomposed_setup_neighbors_and_remainders():
    """50^3 over iproc 2x2x2 and 50x40x36 over 2x1x3: remainder on the last rank, periodic wrap,
    same-peer left/right neighbours when iproc = 2 (MPIExchangeBoundariesnD.c:65-76)."""
    for n, iproc in (((50, 50, 50), (2, 2, 2)), ((50, 40, 37), (2, 1, 3))):
        case = cases.ns3d_turbulence(n, "js", iproc=iproc)
        nranks = int(np.prod(iproc))
        for r in range(nranks):
            S = hpo.Setup(case, rank=r)
            sv = Solver.from_case(case, rank=r)
            assert sv.dim_local == S.dim and sv.is_global == S.is_
            x, dxinv = sv.grid()
            assert np.array_equal(x, S.x) and np.array_equal(dxinv, S.dxinv)
            ip = hpo.rank_nd(iproc, r)
            for d in range(3):
                lo, hi = sv.neighbors[2 * d], sv.neighbors[2 * d + 1]
                if iproc[d] == 1:
                    assert lo == -1 and hi == -1
                else:
                    ipl, iph = list(ip), list(ip)
                    ipl[d] = (ip[d] - 1) % iproc[d]
                    iph[d] = (ip[d] + 1) % iproc[d]
                    assert lo == hpo.rank_1d(iproc, ipl) and hi == hpo.rank_1d(iproc, iph)
                    if iproc[d] == 2:
                        assert lo == hi
            sv.close()
    # non-periodic: physical faces have no neighbour
    case = cases.ns3d_rising_bubble((24, 24, 24), "yc", iproc=(2, 2, 1))
    sv = Solver.from_case(case, rank=0)
    assert sv.neighbors == [-1, 1, -1, 2, -1, -1]
    sv.close()


def test_decomposed_advection_field_blocks():
    """LinearADRAdvectionField.c:122-190 on every rank of a split domain: neighbours' interiors on internal faces, the
    periodic copy only where one rank spans the dimension, the mirror image at the end ranks otherwise (sic: also in a
    periodic dimension split among ranks), zero edges / corners."""
    for n, iproc, per in (((26, 21), (2, 3), True), ((26, 21), (3, 1), True), ((13, 11, 14), (2, 1, 2), False),
                          ((50,), (3,), True), ((50,), (2,), False)):
        case = cases.linear_advection_varying(n, "js", iproc=iproc, periodic=per)
        for r in range(int(np.prod(iproc))):
            S = hpo.Setup(case, rank=r)
            sv = Solver.from_case(case, rank=r)
            a = sv.advection_field()
            assert np.array_equal(a, S.adv_field), (n, iproc, per, r)
            sv.close()
    # the interior of every block is the file's field
    case = cases.linear_advection_varying((26, 21), "js", iproc=(2, 3))
    for r in range(6):
        sv = Solver.from_case(case, rank=r)
        a = sv.advection_field().reshape(sv.shape_g()[:-1] + (2,))
        g = sv.ghosts
        blk = a[g:-g, g:-g]
        i0, j0 = sv.is_global
        assert np.array_equal(blk, case.advection_field[j0:j0 + blk.shape[0], i0:i0 + blk.shape[1]])
        assert not a[:g, :g].any() and not a[-g:, -g:].any()
        sv.close()


def test_decomposed_zone_extents_of_open_boundaries_and_sponges():
    """every rank's zone boxes (edge zones: only on the ranks that own the face; sponge: an interior box on every rank it
    overlaps, InitializeBoundaries.c:381-440) against the independent numpy set-up"""
    for case in (cases.with_sponge(cases.ns_channel((24, 20), "js", iproc=(2, 2)), 0, 1, 0.4, 1.0, [1.0, 0.5, 0.0, 2.0]),
                 cases.with_sponge(cases.ns_channel((12, 10, 14), "js", iproc=(2, 1, 2), bcs="sup3"), 2, -1, 0.0, 0.6,
                                   [1, 0, 0, 0, 2]),
                 cases.with_sponge(cases.linear_advection_nd((24, 21), "js", iproc=(3, 2)), 1, 1, 0.2, 0.8, [0.5])):
        for r in range(int(np.prod(case.solver["iproc"]))):
            S = hpo.Setup(case, rank=r)
            sv = Solver.from_case(case, rank=r)
            for n, z in enumerate(S.zones):
                a, b, on = sv.zone_extent(n)
                assert on == z["on"], (case.name, r, n)
                if on:
                    assert a == z["is"] and b == z["ie"], (case.name, r, n)
            sv.close()


@pytest.mark.parametrize("mutate, msg", [
    (lambda c: c.solver.__setitem__("hyp_space_scheme", "weno7"), "weno5"),
    # (compact schemes split among ranks run, component-wise and characteristic: tests/test_gpu_decomposed.py)
    (lambda c: (c.solver.__setitem__("hyp_space_scheme", "cupw5"), c.solver.__setitem__("iproc", [1, 80, 1])), "at most 64 ranks"),
    (lambda c: (c.solver.__setitem__("hyp_space_scheme", "crweno5"), c.solver.__setitem__("iproc", [1, 2, 1]),
                setattr(c, "lusolver", {"reducedsolvetype": "gather-and-solve"})), "gather-and-solve"),
    (lambda c: c.solver.__setitem__("time_scheme", "arkimex"), "rk"),
    (lambda c: (c.solver.__setitem__("time_scheme", "glm-gee"), c.solver.__setitem__("time_scheme_type", "44")), "glm-gee method"),
    (lambda c: (c.solver.__setitem__("time_scheme", "glm-gee"), c.solver.__setitem__("time_scheme_type", "23"),
                setattr(c, "glm_gee", {"ee_mode": "both"})), "ee_mode"),
    (lambda c: c.solver.__setitem__("time_scheme_type", "ssprk2"), "ssprk3"),
    (lambda c: c.solver.__setitem__("ghost", 2), "ghost"),
    (lambda c: c.solver.__setitem__("model", "shallow-water-2d"), "model"),
    (lambda c: c.boundary[0].__setitem__("type", "thermal-slip-wall"), "boundary type"),
    (lambda c: c.physics.__setitem__("upwinding", "steger-warming"), "upwinding"),
    (lambda c: c.solver.__setitem__("par_space_type", "conservative-1stage"), "nonconservative-2stage"),
    (lambda c: c.solver.__setitem__("par_space_scheme", "2"), "par_space_scheme 4 only"),
    (lambda c: c.solver.__setitem__("immersed_body", "sphere.stl"), "immersed"),
])
def test_unsupported_configurations_fail_loudly(mutate, msg):
    case = cases.ns3d_turbulence((12, 12, 12), "js")
    mutate(case)
    with pytest.raises(HyParB200Error) as e:
        Solver.from_case(case)
    assert msg in str(e.value)
    _lib.load().hpb_clear_error()


@pytest.mark.parametrize("pst", ["nonconservative-2stage", "nonconservative-1.5stage", "conservative-1stage"])
def test_linear_diffusion_in_another_form_fails_loudly(pst):
    """LinearADR installs GFunction and HFunction (LinearADRInitialize.c:190-191), so every par_space_type is a different
    discretisation of the diffusion term (InitializeSolvers.c:107-176); the device has nonconservative-1stage. Without
    diffusion every form is identically zero and is accepted."""
    case = cases.linear_advection_sine(64, "js", diffusion=0.02)
    case.solver["par_space_type"] = pst
    with pytest.raises(HyParB200Error, match="nonconservative-1stage only"):
        Solver.from_case(case)
    _lib.load().hpb_clear_error()
    case = cases.linear_advection_sine(64, "js")
    case.solver["par_space_type"] = pst
    Solver.from_case(case).close()
    case.solver["par_space_type"] = "conservative-2stage"
    with pytest.raises(HyParB200Error, match="not a supported spatial discretization type"):
        Solver.from_case(case)


@pytest.mark.parametrize("scheme", ["crweno5", "hcweno5", "cupw5", "upw5", "1", "2", "4", "muscl2", "muscl3"])
def test_compact_and_linear_schemes_are_accepted(scheme):
    """crweno5 / cupw5 (one rank per line) and upw5 (any decomposition) set up like weno5"""
    case = cases.ns3d_rising_bubble((12, 12, 12), "yc", scheme=scheme)
    sv = Solver.from_case(case)
    assert sv.dim_local == [12, 12, 12]
    sv.close()
    if scheme in ("upw5", "1", "2", "4", "muscl2", "muscl3"):
        sv = Solver.from_case(cases.ns3d_turbulence((12, 12, 12), "js", iproc=(1, 2, 1), scheme=scheme), rank=1)
        assert sv.dim_local == [12, 6, 12]
        sv.close()


def test_gravity_needs_rusanov_like_reference():
    case = cases.ns3d_rising_bubble((12, 12, 12), "yc")
    case.physics["upwinding"] = "roe"               # NavierStokes3DInitialize.c:371-378
    with pytest.raises(HyParB200Error) as e:
        Solver.from_case(case)
    assert "rusanov" in str(e.value).lower()
    _lib.load().hpb_clear_error()


def test_no_cpu_fallback():
    """without a CUDA device every compute entry point fails with HPB_ERR_NO_DEVICE and the error is sticky"""
    L = _lib.load()
    if L.hpb_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    case = cases.linear_advection_sine(64, "js")
    sv = Solver.from_case(case)                       # host set-up works without a GPU
    u = hpo.Setup(case).local_u0()
    for call in (lambda: sv.RHSFunction(u), lambda: sv.HyperbolicFunction(u), lambda: sv.TimeIntegrate(u, 1),
                 lambda: sv.TimeStep(), lambda: sv.ApplyBoundaryConditions(u), lambda: sv.set_solution(u)):
        with pytest.raises(HyParB200Error) as e:
            call()
        assert "no CUDA device" in str(e.value)
    assert L.hpb_error_state() == 2
    L.hpb_clear_error()
    assert L.hpb_error_state() == 0
    sv.close()


def test_product_never_touches_the_oracle():
    """the shipped package must not import, link or execute anything under oracle/"""
    import subprocess
    pkg = os.path.join(ROOT, "hypar_b200")
    banned = ("import oracle", "from oracle", "hpo.", "libhypar_oracle", "hypar_oracle", "oracle/_ref", "hypar_ref")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f)).read()
                for b in banned:
                    assert b not in txt, f"hypar_b200/{f} references the oracle ({b})"
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
