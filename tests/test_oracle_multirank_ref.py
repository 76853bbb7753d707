"""Decomposed runs pinned against the REAL reference running with several ranks (CPU, no GPU needed).

oracle/_ref/hypar_ref_mp = the unmodified reference sources + oracle/ref_harness.cpp linked against the multi-process MPI shim
(oracle/mpishim/mpishim_mp.c: MPI_Init forks HPB_MPI_NP - 1 children, shared-memory point-to-point with tag / communicator
matching). Every rank dumps its own block; tests/_multirank.py::MultiRankOracle -- the checker of the decomposed GPU tests
(tests/test_gpu_decomposed.py, tools/multigpu_check.py) -- must reproduce them bit for bit: TimeRHSFunctionExplicit (with the
reference's MPIExchangeBoundariesnD, its QDerivX / QDerivY exchanges and its iproc-dependent quirks Q1 / Q2) and two full
time steps."""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from _multirank import MultiRankOracle
from hypar_b200 import cases, hypario
from oracle import hpo

EXE = os.path.join(ROOT, "oracle", "_ref", "hypar_ref_mp")
pytestmark = pytest.mark.skipif(not os.access(EXE, os.X_OK), reason="oracle/_ref/hypar_ref_mp not built (make -C oracle refmp)")


def run_ref_mp(case, mode, args=(), nranks=1):
    d = tempfile.mkdtemp(prefix="hpbmp_")
    try:
        case.write(d)
        env = dict(os.environ, OMP_NUM_THREADS="1", HPB_MPI_NP=str(nranks))
        p = subprocess.run([EXE, mode, *map(str, args)], cwd=d, env=env, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, f"reference failed ({p.returncode}):\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}"
        return {os.path.basename(f)[4:-4]: hypario.read_ref_dump(f) for f in glob.glob(os.path.join(d, "ref_*.bin"))}
    finally:
        shutil.rmtree(d, ignore_errors=True)


CASES = [
    cases.ns3d_turbulence((25, 14, 13), "mapped", iproc=(2, 1, 1)),                     # viscous, remainder in x
    cases.ns3d_turbulence((14, 26, 27), "js", iproc=(1, 2, 2)),                          # viscous, y/z split (Q1: z)
    cases.ns3d_turbulence((26, 25, 27), "z", iproc=(2, 2, 2)),                           # 8 ranks, remainders
    cases.ns3d_turbulence((14, 13, 38), "yc", viscous=False, iproc=(1, 1, 3)),           # inviscid, 3 ranks
    cases.ns3d_turbulence((16, 14, 26), "mapped", upwinding="roe", iproc=(1, 1, 2)),     # same-peer periodic pair, Roe
    cases.ns3d_rising_bubble((14, 26, 12), "yc", iproc=(1, 2, 1)),                       # walls + gravity
    cases.ns2d_vortex((40, 27), "mapped", iproc=(2, 2)),
    cases.ns_channel((14, 12, 16), "mapped", viscous=True, bcs="amb3", iproc=(2, 1, 2)),
    cases.ns2d_rising_bubble((24, 28), "yc", iproc=(2, 2)),
    cases.ns2d_vortex((28, 24), "z", upwinding="roe", interp="characteristic", iproc=(2, 1)),
    cases.linear_advection_varying((26, 21), "js", iproc=(2, 3)),
]
c1 = cases.euler1d_sod(101, "js", interp="characteristic", upwinding="roe")
c1.solver["iproc"] = [2]
c2 = cases.linear_advection_sine(96, "z")
c2.solver["iproc"] = [3]
CASES += [c1, c2]
for c in CASES:
    c.name += "_iproc" + "x".join(str(v) for v in c.solver["iproc"])


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_multirank_oracle_equals_the_multirank_reference(case):
    nr = int(np.prod(case.solver["iproc"]))
    MO = MultiRankOracle(case)
    out = run_ref_mp(case, "rhs", nranks=nr)
    u = MO.local_u0()
    rhs = MO.rhs(u)
    for r in range(nr):
        a = out[f"rhs.r{r:04d}"]["data"]
        assert np.array_equal(a, rhs[r].ravel()), f"rank {r}: rhs differs from the reference by {np.abs(a - rhs[r].ravel()).max():.3e}"
        ub = out[f"u.r{r:04d}"]["data"]
        assert np.array_equal(ub, u[r].ravel()), f"rank {r}: u after BCs + halo differs from the reference"
    out = run_ref_mp(case, "steps", [2], nranks=nr)
    u = MO.local_u0()
    for _ in range(2):
        MO.time_step(u, float(case.solver["dt"]), hpo.rk_type_of(case))
    for r in range(nr):
        S = MO.S[r]
        a = S.interior(out[f"ufinal.r{r:04d}"]["data"].reshape(S.shape_g()))
        assert np.array_equal(a, S.interior(u[r])), f"rank {r}: u after 2 steps differs from the reference"
