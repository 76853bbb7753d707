"""The drop-in claim, end to end: HyPar's own executable with libhypar_b200.so attached
(oracle/_ref/hypar_b200_dropin = the unmodified reference objects + integration/hyparb200_attach.c +
integration/hypar_b200_main.cpp, one added call) against the plain reference executable
(oracle/_ref/hypar_main_mpi1, same main without the call) on the same run directory: HyPar's input files in,
HyPar's own output files out (op_*.bin through its WriteArray/WriteBinary, conservation.dat, errors.dat, the
screen log's CFL / norm / cons_err columns).

  exact path (HYPARB200_USE_FUSED=0)  -> every solution file BYTE-IDENTICAL to the reference's
  production path                      -> relative Linf <= 1e-11 (the documented final-time bound)
  both modes of the glue: device-resident (default) and host-array (HYPARB200_MODE=host)

Both executables are built in the authoring container (integration/Makefile needs the reference tree) and travel to
the GPU box with the repository snapshot; nothing here reads /root/reference at run time.
"""
import filecmp
import glob
import os
import re
import subprocess

import numpy as np
import pytest

from hypar_b200 import cases, hypario

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "hypar_main_mpi1")
B200_EXE = os.path.join(ROOT, "oracle", "_ref", "hypar_b200_dropin")


def _prep(case, n_iter=5, cons=True, screen=2, fileop=3):
    case.solver.update({"n_iter": n_iter, "screen_op_iter": screen, "file_op_iter": fileop,
                        "op_overwrite": "no", "conservation_check": "yes" if cons else "no"})
    return case


CASES = [
    _prep(cases.linear_advection_sine(128, "mapped")),                                   # C1
    _prep(cases.euler1d_sod(101, "js")),                                                 # C2: char + Roe (exact path only)
    _prep(cases.ns2d_vortex((32, 24), "yc")),                                            # C3
    _prep(cases.ns3d_turbulence((16, 14, 12), "mapped"), n_iter=4, cons=False),          # C4: viscous
    _prep(cases.ns3d_density_wave((14, 12, 10), "z"), n_iter=4),                         # C5a
    _prep(cases.ns3d_rising_bubble((12, 16, 10), "yc"), n_iter=4, screen=1),             # C5b: walls + gravity
    # compact schemes through HyPar's own executable (exact kernels whatever HYPARB200_USE_FUSED says)
    _prep(cases.ns2d_vortex((32, 24), "mapped", scheme="crweno5")),
    _prep(cases.ns3d_rising_bubble((12, 16, 10), "yc", scheme="crweno5"), n_iter=4, screen=1),
    _prep(cases.euler1d_sod(101, "js", interp="components", scheme="cupw5")),
    _prep(cases.ns2d_vortex((32, 24), "z+rc0.5", scheme="hcweno5")),                    # weno.inp rc read by HyPar, handed over by the glue
    _prep(cases.ns2d_vortex((32, 24), "js", upwinding="roe", interp="characteristic")),       # NavierStokes2D char + Roe
    _prep(cases.ns2d_rising_bubble((24, 28), "yc"), n_iter=4, screen=1),                      # NavierStokes2D + gravity
    _prep(cases.ns_channel((28, 24), "js"), n_iter=4),                                        # inflow / outflow / walls
    _prep(cases.ns_channel((12, 14, 12), "mapped", bcs="sup3", mach=1.4), n_iter=4),          # supersonic / Dirichlet / ambivalent
    _prep(cases.with_muscl(cases.ns2d_vortex((32, 24), "js", upwinding="roe"), "muscl3")),    # muscl.inp through HyPar's reader
    _prep(cases.euler1d_sod(101, "mapped", gravity=1.0)),                                     # Euler1D + gravity
    _prep(cases.with_sponge(cases.linear_advection_nd((28, 24), "js"), 0, -1, 0.0, 0.4, [0.5])),   # sponge zone
    _prep(cases.ns2d_vortex((32, 24), "mapped", upwinding="roe", interp="characteristic", scheme="crweno5")),   # char CRWENO5
    _prep(cases.burgers_nd((32, 24), "js")),                                                  # model burgers
    _prep(cases.linear_advection_varying((32, 24), "js")),                                    # advection.inp through HyPar's reader
]


def _run(exe, d, env=None):
    e = dict(os.environ, OMP_NUM_THREADS="1")
    e.update(env or {})
    p = subprocess.run([exe], cwd=d, env=e, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, f"{os.path.basename(exe)} failed ({p.returncode}):\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}"
    return p.stdout


def _screen_rows(stdout):
    rows = []
    for line in stdout.splitlines():
        if line.startswith("iter="):
            rows.append({k: float(v) for k, v in re.findall(r"(\w+)[=:]\s*([-+0-9.Ee]+)", line) if k != "wctime"})
    return rows


def _dat(path):
    return [float(x) for x in open(path).read().split()] if os.path.exists(path) else None


@pytest.fixture(scope="module")
def need_exes(need_gpu):
    for exe in (REF_EXE, B200_EXE):
        if not os.access(exe, os.X_OK):
            pytest.fail(f"{exe} is missing: build it with `make -C oracle ref && make -C integration` where the "
                        "reference tree is available (it ships to the GPU box with the repository snapshot)")


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
@pytest.mark.parametrize("variant", ["exact-resident", "exact-host", "fused-resident"])
def test_hypar_executable_with_library_attached(need_exes, case, variant, tmp_path):
    path, mode = variant.split("-")
    dref, dnew = str(tmp_path / "ref"), str(tmp_path / "b200")
    case.write(dref)
    case.write(dnew)
    # an exact solution file so that CalculateError has something to do (errors.dat): the initial solution
    for d in (dref, dnew):
        with open(os.path.join(d, "initial.inp"), "rb") as f, open(os.path.join(d, "exact.inp"), "wb") as g:
            g.write(f.read())
    out_ref = _run(REF_EXE, dref)
    out_new = _run(B200_EXE, dnew, {"HYPARB200_USE_FUSED": "0" if path == "exact" else "1", "HYPARB200_MODE": mode})
    assert "hypar_b200 attached" in out_new and f"{mode} mode" in out_new
    m = re.search(r"(\d+) steps, (\d+) kernel launches", out_new)
    assert m and int(m.group(1)) == int(case.solver["n_iter"]) and int(m.group(2)) > 0, "no CUDA kernels were launched"

    files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(dref, "op_*.bin")))
    assert len(files) >= 3, files                       # initial, one intermediate, final
    assert files == sorted(os.path.basename(f) for f in glob.glob(os.path.join(dnew, "op_*.bin")))
    for f in files:
        a, b = os.path.join(dref, f), os.path.join(dnew, f)
        if path == "exact":
            assert filecmp.cmp(a, b, shallow=False), f"{f}: not byte-identical to the reference's file"
        else:
            xa, ua = hypario.read_op_bin(a)[:2]
            xb, ub = hypario.read_op_bin(b)[:2]
            assert all(np.array_equal(p, q) for p, q in zip(xa, xb)), f"{f}: grid differs"
            err = np.abs(ua - ub).max() / np.abs(ua).max()
            assert err <= 1e-11, f"{f}: rel Linf {err:.3e}"

    # the screen log: CFL and norm every screen_op_iter steps, conservation error
    ra, rb = _screen_rows(out_ref), _screen_rows(out_new)
    assert len(ra) == len(rb) and len(ra) >= 1
    for x, y in zip(ra, rb):
        assert x["iter"] == y["iter"] and x["t"] == y["t"]
        assert abs(x["CFL"] - y["CFL"]) <= 2e-3 * x["CFL"]          # printed with 4 significant digits
        assert abs(x["norm"] - y["norm"]) <= 2e-4 * x["norm"] + 1e-300
        if "cons_err" in x:
            # rounding noise (1e-15) where the scheme conserves; with gravity the source term makes it O(1) and the
            # two runs must agree on it
            assert abs(y["cons_err"] - x["cons_err"]) <= 2e-4 * x["cons_err"] + 1e-10, \
                f"conservation error {y['cons_err']} vs {x['cons_err']}"
    # errors.dat: dims, iproc, dt, L1, L2, Linf errors, runtimes -- compare the three error norms
    ea, eb = _dat(os.path.join(dref, "errors.dat")), _dat(os.path.join(dnew, "errors.dat"))
    nd = int(case.solver["ndims"])
    for k in range(2 * nd + 1, 2 * nd + 4):
        tol = 0.0 if path == "exact" else 1e-9 * abs(ea[k])
        assert abs(ea[k] - eb[k]) <= tol, f"errors.dat column {k}: {ea[k]!r} vs {eb[k]!r}"
    if case.solver["conservation_check"] == "yes":
        ca, cb = _dat(os.path.join(dref, "conservation.dat")), _dat(os.path.join(dnew, "conservation.dat"))
        assert len(ca) == len(cb)
        nv = int(case.solver["nvars"])
        # the error is |vol + boundary integral - vol0| / max(|vol0|, 1): a difference of integrals of size
        # (flux) x (face area) x (time), e.g. 1e5 Pa x 1e6 m^2 x 0.04 s for the bubble's wall-normal momentum. Two
        # evaluation orders agree to rounding OF THOSE INTEGRALS, not of their (cancelling) difference.
        xs, u0 = hypario.read_op_bin(os.path.join(dref, files[0]))
        ext = [float(x[-1] - x[0]) * len(x) / max(len(x) - 1, 1) for x in xs]
        area = max(np.prod(ext) / e for e in ext)
        scale = float(np.abs(u0).max()) * area * float(case.solver["dt"]) * int(case.solver["n_iter"])
        for k in range(len(ca) - nv, len(ca)):
            assert abs(ca[k] - cb[k]) <= 1e-9 * abs(ca[k]) + 1e-14 * scale + 1e-12, \
                f"conservation.dat column {k}: {ca[k]!r} vs {cb[k]!r} (integral scale {scale:.2e})"


# ------------------------------------------------------------------------------------------------------------------
# GLM-GEE through the glue (HyPar::TimeIntegrate = TimeGLMGEE): the solution files, the auxiliary-solution files
# OutputSolution.cpp:42-79 writes through TimeGetAuxSolutions (ts0_*.bin)
GLM = [_prep(cases.with_glmgee(cases.ns2d_vortex((32, 24), "yc"), "exrk2a", "yyt"), cons=False),
       _prep(cases.with_glmgee(cases.ns3d_turbulence((14, 12, 10), "mapped"), "rk32g1"), n_iter=4, cons=False),
       _prep(cases.with_glmgee(cases.euler1d_sod(101, "js"), "23"), cons=True)]


@pytest.mark.parametrize("case", GLM, ids=[c.name for c in GLM])
@pytest.mark.parametrize("variant", ["exact-resident", "exact-host", "fused-resident"])
def test_glmgee_through_the_glue(need_exes, case, variant, tmp_path):
    path, mode = variant.split("-")
    dref, dnew = str(tmp_path / "ref"), str(tmp_path / "b200")
    case.write(dref)
    case.write(dnew)
    _run(REF_EXE, dref)
    out_new = _run(B200_EXE, dnew, {"HYPARB200_USE_FUSED": "0" if path == "exact" else "1", "HYPARB200_MODE": mode})
    assert "hypar_b200 attached" in out_new
    for pat, least in (("op_*.bin", 3), ("ts0_*.bin", 2)):       # no auxiliary file with the initial solution (no integrator yet)
        files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(dref, pat)))
        assert len(files) >= least, (pat, files)
        assert files == sorted(os.path.basename(f) for f in glob.glob(os.path.join(dnew, pat)))
        scale = np.abs(hypario.read_op_bin(os.path.join(dref, sorted(glob.glob(os.path.join(dref, "op_*.bin")))[-1]))[1]).max()
        for f in files:
            a, b = os.path.join(dref, f), os.path.join(dnew, f)
            viscous = float(case.physics.get("Re", -1.0)) > 0
            if path == "exact" and not viscous:
                assert filecmp.cmp(a, b, shallow=False), f"{f}: not byte-identical to the reference's file"
            else:
                ua, ub = hypario.read_op_bin(a)[1], hypario.read_op_bin(b)[1]
                assert np.abs(ua - ub).max() <= (1e-11 if path == "fused" else 1e-14) * scale, f"{f}: {np.abs(ua - ub).max():.3e}"
    # glm_err.dat is not compared: HyPar's main calls CalculateError after TimeCleanup, where TimeError returns at once
    # (TimeError.c:41: the integrator is gone) -- the file is never written; tests/test_gpu_glmgee.py checks those norms
    assert not os.path.exists(os.path.join(dref, "glm_err.dat")) and not os.path.exists(os.path.join(dnew, "glm_err.dat"))


# ------------------------------------------------------------------------------------------------------------------
# ensembles (nsims > 1, TimeRK.c:50-93) through the glue: one library solver per SimulationObject
REF_MAIN = os.path.join(ROOT, "oracle", "_ref", "hypar_ref_main")


@pytest.mark.parametrize("name", ["vortex3", "sod2", "turb12"])
@pytest.mark.parametrize("path", ["exact", "fused"])
def test_ensemble_through_the_glue(need_exes, name, path, tmp_path):
    if not os.access(REF_MAIN, os.X_OK):
        pytest.fail(f"{REF_MAIN} is missing (make -C oracle ref)")
    sims = cases.ensemble(name, n_iter=4)
    dref, dnew = str(tmp_path / "ref"), str(tmp_path / "b200")
    cases.write_ensemble(dref, sims)
    cases.write_ensemble(dnew, sims)
    _run(REF_MAIN, dref)
    out_new = _run(B200_EXE, dnew, {"HYPARB200_USE_FUSED": "0" if path == "exact" else "1"})
    assert "one solver per simulation" in out_new
    files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(dref, "op_*.bin")))
    assert len(files) == len(sims), files
    assert files == sorted(os.path.basename(f) for f in glob.glob(os.path.join(dnew, "op_*.bin")))
    for f in files:
        a, b = os.path.join(dref, f), os.path.join(dnew, f)
        if path == "exact":
            assert filecmp.cmp(a, b, shallow=False), f"{f}: not byte-identical to the reference's file"
        else:
            ua, ub = hypario.read_op_bin(a)[1], hypario.read_op_bin(b)[1]
            assert np.abs(ua - ub).max() <= 1e-11 * np.abs(ua).max()


# ------------------------------------------------------------------------------------------------------------------
# several ranks: HyPar's executable, one GPU per rank, the halo exchange inside the library over NCCL; the ranks are
# forked by the multi-process MPI shim (oracle/mpishim/mpishim_mp.c). Reference = the unmodified reference with the same
# number of ranks on the CPU. Needs as many GPUs as ranks (NCCL refuses two ranks on one device).
REF_MP = os.path.join(ROOT, "oracle", "_ref", "hypar_main_mp")
B200_MP = os.path.join(ROOT, "oracle", "_ref", "hypar_b200_dropin_mp")
MP_CASES = [
    _prep(cases.ns3d_turbulence((26, 25, 27), "mapped", iproc=(1, 1, 2)), n_iter=4, cons=False),
    _prep(cases.ns3d_rising_bubble((14, 26, 12), "yc", iproc=(1, 2, 1)), n_iter=4, screen=1),
    _prep(cases.ns2d_vortex((40, 27), "mapped", iproc=(2, 1))),
]


@pytest.mark.parametrize("case", MP_CASES, ids=[c.name for c in MP_CASES])
@pytest.mark.parametrize("path", ["exact", "fused"])
def test_multirank_executable_with_library_attached(need_exes, case, path, tmp_path):
    import ctypes
    from hypar_b200 import _lib
    nr = int(np.prod(case.solver["iproc"]))
    if _lib.load().hpb_device_count() < nr:
        pytest.skip(f"needs {nr} GPUs (one rank per GPU)")
    for exe in (REF_MP, B200_MP):
        if not os.access(exe, os.X_OK):
            pytest.fail(f"{exe} is missing (make -C oracle refmp && make -C integration)")
    dref, dnew = str(tmp_path / "ref"), str(tmp_path / "b200")
    case.write(dref)
    case.write(dnew)
    env = {"HPB_MPI_NP": str(nr)}
    out_ref = _run(REF_MP, dref, env)
    out_new = _run(B200_MP, dnew, dict(env, HYPARB200_USE_FUSED="0" if path == "exact" else "1"))
    assert "in-library NCCL halo exchange" in out_new
    files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(dref, "op_*.bin")))
    assert len(files) >= 2 and files == sorted(os.path.basename(f) for f in glob.glob(os.path.join(dnew, "op_*.bin")))
    for f in files:
        a, b = os.path.join(dref, f), os.path.join(dnew, f)
        if path == "exact":
            assert filecmp.cmp(a, b, shallow=False), f"{f}: not byte-identical to the {nr}-rank reference's file"
        else:
            ua, ub = hypario.read_op_bin(a)[1], hypario.read_op_bin(b)[1]
            assert np.abs(ua - ub).max() <= 1e-11 * np.abs(ua).max()
    ra, rb = _screen_rows(out_ref), _screen_rows(out_new)
    assert len(ra) == len(rb) and len(ra) >= 1
    for x, y in zip(ra, rb):
        assert abs(x["CFL"] - y["CFL"]) <= 2e-3 * x["CFL"] and abs(x["norm"] - y["norm"]) <= 2e-4 * x["norm"] + 1e-300
