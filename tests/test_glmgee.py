"""GLM-GEE time integrators (TimeGLMGEE.c, TimeGLMGEEInitialize.c, TimeError.c) -- CPU part.

The oracle's restatement (hpo_time_step_glmgee, hpo_glmgee_error; coefficient tables oracle/glmgee_tables.h, generated
by tools/make_glmgee_tables.py from the reference's own initialisation) pinned against the UNMODIFIED reference:
  * the committed fixtures tests/golden/glmgee/*.npz (tools/make_golden.py): solution, auxiliary solution and the six
    numbers of glm_err.dat after three steps -- bit for bit;
  * live, where oracle/_ref is present: all seven methods, both error-estimation modes, both reference builds;
  * the decomposed oracle (tests/_multirank.py) against the reference running with the same ranks.
"""
import glob
import json
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from hypar_b200 import cases  # noqa: E402
from oracle import hpo  # noqa: E402

GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "glmgee", "*.npz")))


def load(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    return z, getattr(cases, meta["builder"])(**meta["kwargs"])


def oracle_steps(case, n, mpi_semantics=True):
    S = hpo.Setup(case, mpi_semantics=mpi_semantics)
    O = hpo.Oracle(S)
    m, mode = hpo.glmgee_of(case)
    u = S.local_u0()
    ua = O.glmgee_aux0(u, mode)
    for _ in range(n):
        O.time_step_glmgee(u, ua, float(case.solver["dt"]), m, mode)
    return S, O, u, ua, O.glmgee_error(u, ua, m, mode)


def test_fixtures_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    z, case = load(path)
    S, O, u, ua, err = oracle_steps(case, 3)
    assert np.array_equal(S.interior(u), S.interior(z["steps3_u"])), "u after 3 steps"
    assert np.array_equal(S.interior(ua), S.interior(z["steps3_uaux"])), "auxiliary solution after 3 steps"
    assert np.array_equal(err, z["steps3_glmerr"]), "glm_err.dat"


def test_tables_are_consistent():
    """c = row sums of A (TimeGLMGEEInitialize.c:415-420); the first step-completion row reproduces a consistent method:
    D[0][0] + D[0][1] = 1 in yyt mode, D = I in both; B row 0 sums to 1 in yeps mode"""
    for name, m in hpo.GLMGEE_METHODS.items():
        for mode in (0, 1):
            T = hpo.glmgee_table(m, mode)
            s = T["s"]
            assert np.array_equal(T["D"], np.array([1.0, 0.0, 0.0, 1.0])), name
            A = T["A"].reshape(s, s)
            assert np.allclose(np.triu(A), 0.0), f"{name}: A is not strictly lower triangular"
            if mode == 0:
                assert abs(T["B"][:s].sum() - 1.0) < 1e-14, f"{name}: sum b = {T['B'][:s].sum()}"
                assert abs(T["B"][s:].sum()) < 1e-14, f"{name}: the error row sums to {T['B'][s:].sum()}"


def _live_cases():
    return [
        cases.with_glmgee(cases.linear_advection_sine(64, "mapped"), "24", "yyt"),
        cases.with_glmgee(cases.burgers_nd((32,), "js"), "25i"),
        cases.with_glmgee(cases.burgers_nd((32,), "js"), "25i", "yyt"),
        cases.with_glmgee(cases.euler1d_sod(101, "z"), "35"),
        cases.with_glmgee(cases.euler1d_sod(101, "js"), "23", "yyt"),
        cases.with_glmgee(cases.ns2d_vortex((24, 20), "yc"), "exrk2a"),
        cases.with_glmgee(cases.linear_advection_sine(64, "js"), "rk32g1"),
        cases.with_glmgee(cases.ns3d_turbulence((12, 10, 8), "mapped"), "rk32g1", "yyt"),
        cases.with_glmgee(cases.ns3d_rising_bubble((10, 12, 8), "js"), "rk285ex"),
        cases.with_glmgee(cases.burgers_nd((20, 16), "z"), "24"),
    ]


LIVE = _live_cases()


@pytest.mark.parametrize("case", LIVE, ids=[c.name for c in LIVE])
@pytest.mark.parametrize("exe", ["hypar_ref", "hypar_ref_mpi1"])
def test_oracle_matches_reference_live(case, exe):
    from refrun import ref_available, run_reference
    if not ref_available(exe):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    o = run_reference(case, "steps", [3], exe=exe)
    S, O, u, ua, err = oracle_steps(case, 3, mpi_semantics=exe.endswith("mpi1"))
    assert np.array_equal(S.interior(u), S.interior(o["ufinal"]["data"])), "u after 3 steps"
    assert np.array_equal(S.interior(ua), S.interior(o["uaux"]["data"])), "auxiliary solution after 3 steps"
    ref = np.array([float(v) for v in re.search(r"GLMERR (.*)", o["stdout"]).group(1).split()])
    assert np.array_equal(err, ref), "glm_err.dat"


def test_decomposed_oracle_matches_the_multirank_reference():
    from _multirank import MultiRankOracle
    from test_oracle_multirank_ref import EXE, run_ref_mp
    if not os.access(EXE, os.X_OK):
        pytest.skip("oracle/_ref/hypar_ref_mp not built (needs /root/reference)")
    for case in (cases.with_glmgee(cases.ns2d_vortex((28, 24), "js", iproc=(2, 2)), "exrk2a", "yyt"),
                 cases.with_glmgee(cases.ns3d_turbulence((14, 12, 26), "mapped", iproc=(1, 1, 2)), "23")):
        nr = int(np.prod(case.solver["iproc"]))
        m, mode = hpo.glmgee_of(case)
        MO = MultiRankOracle(case)
        u = MO.local_u0()
        ua = [MO.O[r].glmgee_aux0(u[r], mode) for r in range(nr)]
        for _ in range(2):
            MO.time_step_glmgee(u, ua, float(case.solver["dt"]), m, mode)
        out = run_ref_mp(case, "steps", [2], nranks=nr)
        for r in range(nr):
            S = MO.S[r]
            a = S.interior(out[f"ufinal.r{r:04d}"]["data"].reshape(S.shape_g()))
            assert np.array_equal(a, S.interior(u[r])), f"{case.name} rank {r}: u after 2 steps"
            b = S.interior(out[f"uaux.r{r:04d}"]["data"].reshape(S.shape_g()))
            assert np.array_equal(b, S.interior(ua[r])), f"{case.name} rank {r}: auxiliary solution after 2 steps"
