"""HyPar's ensemble driver on the device: hypar_b200.ensemble.Ensemble attached to an ensemble run directory
(simulation.inp + per-simulation sizes and initial_<n>.inp) against the REFERENCE'S OWN OUTPUT of the same directory
(tests/golden/ensemble/*.npz, written by the reference's unmodified main; tools/make_golden_ensemble.py).

  exact path       every simulation's final solution BIT-IDENTICAL to the reference's op_<n>.bin
  production path  relative Linf <= 1e-11 (the documented final-time bound)
  screen columns   norm over all simulations together, CFL of the last simulation (as the reference prints them)
  output files     op_<n>.bin written by the class read back equal to the reference's arrays
"""
import os

import numpy as np
import pytest

from conftest import assert_exact, rel_linf
from hypar_b200 import cases, hypario
from hypar_b200.ensemble import Ensemble, index_string

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ensemble")
NAMES = ["vortex3", "burgers2", "linadvvar2", "sod2", "turb12"]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("fused", [False, True], ids=["exact", "fused"])
def test_ensemble_directory_against_the_reference_files(need_gpu, name, fused, tmp_path):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    sims = cases.ensemble(name)
    d = str(tmp_path / "run")
    cases.write_ensemble(d, sims)
    n_iter = int(sims[0].solver["n_iter"])
    every = int(sims[0].solver["screen_op_iter"])
    E = Ensemble.from_directory(d, use_fused=fused)
    assert len(E) == len(sims)
    E.set_initial_solutions()
    rows = []
    for it in range(n_iter):
        screen = (it + 1) % every == 0
        if screen:
            cfl = E.dev_ComputeCFL()
        E.TimeSteps(1)
        if screen:
            rows.append((it + 1, cfl, E.dev_StepNorm()))
    U = E.get_solutions()
    for n, (sv, u) in enumerate(zip(E.sims, U)):
        got, ref = sv.interior(u), z[f"u_{n}"]
        assert got.shape == ref.shape
        if not fused:
            # viscous simulations: CUDA's exp / log vs glibc's in the viscosity law (conftest.assert_exact)
            assert_exact(got, ref, f"simulation {n} vs the reference's op_{n}.bin", libm_ulp=(name == "turb12"))
        else:
            assert rel_linf(got, ref) <= 1e-11, f"simulation {n}: rel Linf {rel_linf(got, ref):.3e}"
        assert sv.kernel_launches > 0
    scr = z["screen"]
    assert len(rows) == len(scr)
    for (it, cfl, norm), r in zip(rows, scr):
        assert it == int(r[0])
        assert abs(cfl - r[1]) <= 6e-4 * r[1]                # printed with 4 significant digits
        assert abs(norm - r[2]) <= 6e-5 * r[2] + 1e-300      # ... 5
    # the output files in the reference's naming
    names = E.write_solutions(d)
    assert names == [f"op_{index_string(n, len(sims))}.bin" for n in range(len(sims))]
    for n, (nm, s) in enumerate(zip(names, sims)):
        x, u = hypario.read_op_bin(os.path.join(d, nm))[:2]
        assert all(np.array_equal(a, b) for a, b in zip(x, s.x)), "grid"
        if not fused:
            assert_exact(u, z[f"u_{n}"], f"op file of simulation {n}", libm_ulp=(name == "turb12"))
    E.close()


def test_ensemble_streams_are_independent(need_gpu, tmp_path):
    """Stepping the simulations one after the other, or all together with nothing synchronising in between, gives the
    same bits: every simulation owns its stream and its buffers."""
    sims = cases.ensemble("vortex3")
    d = str(tmp_path / "run")
    cases.write_ensemble(d, sims)
    E = Ensemble.from_directory(d, use_fused=True)
    E.set_initial_solutions()
    E.TimeSteps(3)
    together = E.get_solutions()
    E.set_initial_solutions()
    for sv in E.sims:
        sv.TimeSteps(3)
        sv.synchronize()
    apart = E.get_solutions()
    for a, b in zip(together, apart):
        assert np.array_equal(a, b)
    E.close()
