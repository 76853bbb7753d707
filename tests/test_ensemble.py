"""HyPar's ensemble driver (simulation.inp, nsims simulations in lock-step): host-side pieces on the CPU.

  * the readers / the directory writer agree with the reference's conventions (per-simulation size / iproc vectors,
    initial_<n>.inp with (int) log10(nsims) + 1 digits)
  * the committed fixtures (tests/golden/ensemble/*.npz, written by tools/make_golden_ensemble.py from the reference's OWN
    main, unmodified) are reproduced bit for bit by the oracle stepping every simulation on its own with simulation 0's
    time step -- i.e. "ensemble = independent simulations" is a fact of the reference, which is what hypar_b200.ensemble
    relies on; the screen log's norm (all simulations together) and CFL (the last simulation's) likewise
  * where oracle/_ref/hypar_ref_main is present, the fixtures are regenerated live and must not have moved
"""
import math
import os

import numpy as np
import pytest

from hypar_b200 import cases, hypario
from hypar_b200.ensemble import Ensemble, index_string
from oracle import hpo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ensemble")
NAMES = ["vortex3", "burgers2", "linadvvar2", "sod2", "turb12"]


def test_index_string_width():
    assert index_string(0, 1) == "0" and index_string(3, 4) == "3" and index_string(9, 10) == "09"
    assert index_string(7, 12) == "07" and index_string(11, 12) == "11" and index_string(5, 100) == "005"


def test_fixtures_present():
    for n in NAMES:
        assert os.path.exists(os.path.join(GOLD, n + ".npz")), n


@pytest.mark.parametrize("name", NAMES)
def test_directory_round_trip(name, tmp_path):
    sims = cases.ensemble(name)
    d = str(tmp_path / "run")
    cases.write_ensemble(d, sims)
    assert hypario.read_simulation_inp(os.path.join(d, "simulation.inp")) == len(sims)
    cfgs = hypario.read_ensemble_solver_inp(os.path.join(d, "solver.inp"), len(sims))
    assert len(cfgs) == len(sims)
    for c, s in zip(cfgs, sims):
        assert c["size"] == [int(v) for v in s.solver["size"]] and c["model"] == s.solver["model"]
        assert c["dt"] == float(sims[0].solver["dt"]) and c["iproc"] == [1] * s.ndims
    for n, s in enumerate(sims):
        x, u = hypario.read_initial_bin(os.path.join(d, f"initial_{index_string(n, len(sims))}.inp"), s.solver["size"], s.nvars)
        assert np.array_equal(u, s.u0) and all(np.array_equal(a, b) for a, b in zip(x, s.x))
    # host set-up of every simulation through the C ABI (no device needed)
    E = Ensemble.from_directory(d)
    assert len(E) == len(sims)
    for sv, s in zip(E.sims, sims):
        assert sv.dim_local == [int(v) for v in s.solver["size"]] and sv.dt == float(sims[0].solver["dt"])
    assert E.npoints_global == [int(np.prod(s.solver["size"])) for s in sims]
    assert all(np.array_equal(a, s.u0) for a, s in zip(E.u0_global, sims))
    if name == "linadvvar2":          # advection_<n>.inp: every simulation its own field
        for sv, s in zip(E.sims, sims):
            S = hpo.Setup(s)
            assert np.array_equal(sv.advection_field(), S.adv_field)
    E.close()


def test_single_simulation_directory_is_an_ensemble_of_one(tmp_path):
    c = cases.linear_advection_sine(64, "js")
    d = str(tmp_path / "run")
    c.write(d)
    E = Ensemble.from_directory(d)
    assert len(E) == 1 and E.sims[0].dim_local == [64]
    E.close()


def oracle_ensemble(sims, n_iter):
    """Every simulation on its own, simulation 0's dt; returns the final interiors and the screen rows (iter, CFL, norm)."""
    dt = float(sims[0].solver["dt"])
    every = int(sims[0].solver["screen_op_iter"])
    S = [hpo.Setup(c, mpi_semantics=True) for c in sims]
    O = [hpo.Oracle(s) for s in S]
    U = [s.local_u0() for s in S]
    npts = sum(int(np.prod(c.solver["size"])) for c in sims)
    rows = []
    for it in range(n_iter):
        screen = (it + 1) % every == 0
        if screen:
            for o, u in zip(O, U):
                o.apply_bc(u)
            cfl = O[-1].cfl(U[-1], dt)                       # TimePreStep.c:84-92: the last simulation's value survives
            prev = [u.copy() for u in U]
        for c, o, u in zip(sims, O, U):
            o.time_step(u, dt, hpo.rk_type_of(c))
        if screen:
            ss = sum(float(((s.interior(u) - s.interior(p)) ** 2).sum()) for s, u, p in zip(S, U, prev))
            rows.append((it + 1, cfl, math.sqrt(ss / npts)))
    return [s.interior(u) for s, u in zip(S, U)], rows


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_the_reference_ensemble(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    sims = cases.ensemble(name)
    U, rows = oracle_ensemble(sims, int(sims[0].solver["n_iter"]))
    for n, u in enumerate(U):
        ref = z[f"u_{n}"]
        assert np.array_equal(u.reshape(ref.shape), ref), f"simulation {n}: not bit-identical to the reference's op file"
    scr = z["screen"]
    assert len(scr) == len(rows) >= 1
    for (it, cfl, norm), r in zip(rows, scr):
        assert it == int(r[0])
        assert abs(cfl - r[1]) <= 6e-4 * r[1]                # printed with 4 significant digits
        assert abs(norm - r[2]) <= 6e-5 * r[2] + 1e-300      # ... 5


@pytest.mark.parametrize("name", NAMES)
def test_fixture_matches_reference_live(name):
    import make_golden_ensemble as mg
    if not os.access(mg.EXE, os.X_OK):
        pytest.skip("oracle/_ref/hypar_ref_main not built (needs /root/reference)")
    z = np.load(os.path.join(GOLD, name + ".npz"))
    o = mg.run_reference_ensemble(name)
    for n, u in enumerate(o["u"]):
        assert np.array_equal(u, z[f"u_{n}"])
    assert np.array_equal(np.array(o["screen"]), z["screen"])
