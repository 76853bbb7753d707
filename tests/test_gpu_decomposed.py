"""Domain-decomposed runs: every rank's CUDA result against an oracle run WITH THE SAME iproc (the viscous
term depends on the decomposition: SURVEY.md Q1/Q2). All ranks live in one process on one GPU
(tests/_multirank.py) and are advanced by the library's own distributed step (hpb_TimeStepsLocal): schedule, pack /
unpack / face-layer RK kernels and event ordering are the ones the NCCL run uses; only the transport (device-to-device
copy vs ncclSend/Recv) differs, and that one is covered by tests/test_multigpu_gloo.py (the library's message plan over
gloo) and the multi-GPU check in tools/multigpu_check.py."""
import numpy as np
import pytest

from _multirank import LocalRanks, MultiRankOracle
from conftest import assert_exact, rel_linf
from hypar_b200 import cases
from oracle import hpo

pytestmark = pytest.mark.gpu

DECOMP = [
    cases.ns3d_turbulence((25, 14, 13), "mapped", iproc=(2, 1, 1)),                     # viscous, remainder in x
    cases.ns3d_turbulence((14, 26, 27), "js", iproc=(1, 2, 2)),                          # viscous, y/z split (Q1: z)
    cases.ns3d_turbulence((26, 25, 27), "z", iproc=(2, 2, 2)),                           # 8 ranks, remainders
    cases.ns3d_turbulence((14, 13, 38), "yc", viscous=False, iproc=(1, 1, 3)),           # inviscid, 3 ranks
    cases.ns3d_rising_bubble((14, 26, 12), "yc", iproc=(1, 2, 1)),                       # walls + gravity
    cases.ns2d_vortex((40, 27), "mapped", iproc=(2, 2)),
    cases.euler1d_sod(101, "js", interp="characteristic", upwinding="roe"),              # iproc 1 (generic path)
    cases.linear_advection_sine(96, "z"),
]
DECOMP[-2].solver["iproc"] = [2]
DECOMP[-1].solver["iproc"] = [3]
DECOMP += [
    # open / wall boundary zones and a sponge box across ranks; NavierStokes2D gravity; characteristic + Roe in 2-D
    cases.with_sponge(cases.ns_channel((28, 24), "js", iproc=(2, 2)), 0, 1, 0.4, 1.0, [1.0, 0.5, 0.0, 2.0]),
    cases.ns_channel((14, 12, 16), "mapped", viscous=True, bcs="amb3", iproc=(2, 1, 2)),
    cases.ns2d_rising_bubble((24, 28), "yc", iproc=(2, 2)),
    cases.ns2d_vortex((28, 24), "z", upwinding="roe", interp="characteristic", iproc=(2, 1)),
    cases.with_muscl(cases.linear_advection_nd((24, 21), "js", iproc=(3, 2)), "muscl3"),
    # spatially varying advection field: every rank's block with the neighbours' values / mirror images in its ghosts
    cases.linear_advection_varying((26, 21), "js", iproc=(2, 3)),
    cases.linear_advection_varying((24, 20), "z", iproc=(2, 1), periodic=False),
]
for c in DECOMP:
    c.name += "_iproc" + "x".join(str(v) for v in c.solver["iproc"])


@pytest.mark.parametrize("case", DECOMP, ids=[c.name for c in DECOMP])
@pytest.mark.parametrize("fused", [False, True], ids=["exact", "fused"])
@pytest.mark.parametrize("overlap", [False, True], ids=["serial", "overlap"])
def test_decomposed_rhs_and_steps(need_gpu, case, fused, overlap):
    MO = MultiRankOracle(case)
    u_ref = MO.local_u0()
    rhs_ref = MO.rhs(u_ref)
    LR = LocalRanks(case, use_fused=fused, sweepwise=overlap)
    LR.set_solution(MO.local_u0())
    rhs = LR.rhs()
    scale = max(np.abs(r).max() for r in rhs_ref)
    for r in range(MO.nranks):
        assert np.isfinite(rhs[r]).all()
        if not fused:
            ulp = case.name.startswith("chan") and float(case.physics.get("Re", -1.0)) > 0
            assert_exact(rhs[r], rhs_ref[r], f"rank {r}: rhs", libm_ulp=ulp)
        else:
            lam = MO.O[r].cfl(u_ref[r], float(case.solver["dt"])) / float(case.solver["dt"])
            tol = 1e-12 * scale + 16 * np.finfo(np.float64).eps * lam * np.abs(u_ref[r]).max()
            assert np.abs(rhs[r] - rhs_ref[r]).max() <= tol, \
                f"rank {r}: rhs abs err {np.abs(rhs[r] - rhs_ref[r]).max():.3e} > {tol:.3e}"
    # two full time steps
    dt = float(case.solver["dt"])
    rk = hpo.rk_type_of(case)
    u_ref = MO.local_u0()
    for _ in range(2):
        MO.time_step(u_ref, dt, rk)
    LR.set_solution(MO.local_u0())
    LR.time_step(2)
    u = LR.get_solution()
    for r in range(MO.nranks):
        a, b = MO.S[r].interior(u[r]), MO.S[r].interior(u_ref[r])
        if not fused:
            # viscous channel case: temperatures on which CUDA's and glibc's exp / log differ by an ulp (conftest.assert_exact)
            ulp = case.name.startswith("chan") and float(case.physics.get("Re", -1.0)) > 0
            assert_exact(a, b, f"rank {r}: u after 2 steps", libm_ulp=ulp)
        else:
            assert rel_linf(a, b) <= 1e-11, f"rank {r}: u after 2 steps rel err {rel_linf(a, b):.3e}"
    LR.close()


def test_inviscid_rhs_is_decomposition_invariant(need_gpu):
    """hyperbolic term: the 8-rank result equals the single-rank result on every block (exact path: bit for bit)"""
    kw = dict(n=(26, 25, 27), weno="mapped", viscous=False)
    c8, c1 = cases.ns3d_turbulence(iproc=(2, 2, 2), **kw), cases.ns3d_turbulence(**kw)
    S1 = hpo.Setup(c1)
    r1 = hpo.Oracle(S1).rhs(S1.local_u0()).reshape(S1.shape_g())
    for fused in (False, True):
        LR = LocalRanks(c8, use_fused=fused)
        MO = MultiRankOracle(c8)
        LR.set_solution(MO.local_u0())
        rhs = LR.rhs()
        g = S1.ghosts
        for r, S in enumerate(MO.S):
            sl = tuple(slice(g + S.is_[k], g + S.is_[k] + S.dim[k]) for k in reversed(range(3)))
            a, b = S.interior(rhs[r]), r1[sl]
            if not fused:
                assert np.array_equal(a, b)
            else:
                assert rel_linf(a, b) <= 1e-12
        LR.close()


SWEEPWISE = [cases.ns3d_turbulence((26, 24, 22), "mapped", iproc=(2, 2, 2)),
             cases.ns3d_turbulence((25, 14, 13), "js", iproc=(2, 1, 1)),
             cases.ns3d_rising_bubble((20, 24, 14), "yc", iproc=(1, 2, 1)),
             cases.ns3d_density_wave((16, 12, 20), "mapped", iproc=(1, 1, 2))]


@pytest.mark.parametrize("case", SWEEPWISE, ids=lambda c: c.name + "_" + "x".join(str(v) for v in c.solver["iproc"]))
def test_overlapped_call_sequence_is_identical(need_gpu, case):
    """The overlapped multi-GPU schedule (face layers of the stage vector first, their exchange under the full-array RK
    update; the step completion likewise, which serves the next step's TimePreStep; one sweep per dimension as soon as
    that dimension's Q-derivative halos are unpacked) gives bit-identical results to the serial sequence, and both agree
    with the multi-rank oracle."""
    MO = MultiRankOracle(case)
    A = LocalRanks(case, use_fused=True, sweepwise=False)
    B = LocalRanks(case, use_fused=True, sweepwise=True)
    A.set_solution(MO.local_u0())
    B.set_solution(MO.local_u0())
    ra, rb = A.rhs(), B.rhs()
    rhs_ref = MO.rhs(MO.local_u0())
    scale = max(np.abs(r).max() for r in rhs_ref)
    for r in range(MO.nranks):
        assert np.array_equal(ra[r], rb[r]), f"rank {r}: rhs differs between the schedules ({np.abs(ra[r] - rb[r]).max():.3e})"
        lam = MO.O[r].cfl(MO.local_u0()[r], float(case.solver["dt"])) / float(case.solver["dt"])
        tol = 1e-12 * scale + 16 * np.finfo(np.float64).eps * lam * np.abs(MO.local_u0()[r]).max()
        assert np.abs(rb[r] - rhs_ref[r]).max() <= tol
    for _ in range(2):
        A.time_step()
        B.time_step()
    ua, ub = A.get_solution(), B.get_solution()
    for r in range(MO.nranks):
        # interiors: the overlapped schedule has already exchanged the face ghosts of u for the NEXT step's TimePreStep
        # (they hold the neighbours' new values), the serial one leaves the previous step's there, as the reference does
        a, b = MO.S[r].interior(ua[r]), MO.S[r].interior(ub[r])
        assert np.array_equal(a, b), f"rank {r}: u after 2 steps differs between the schedules"
    A.close()
    B.close()


@pytest.mark.parametrize("case", [DECOMP[2], DECOMP[4], DECOMP[5]], ids=lambda c: c.name)
def test_exchange_boundaries_on_the_device_solution(need_gpu, case):
    """hpb_ExchangeBoundariesnD / ...Local = MPIExchangeBoundariesnD (MPIExchangeBoundariesnD.c:42-173) on the device solution:
    every face ghost layer of every rank equals the multi-rank oracle's after its exchange (the oracle is pinned to the
    multi-rank reference, tests/test_oracle_multirank_ref.py); interiors, physical faces, edges and corners are untouched."""
    MO = MultiRankOracle(case)
    u0 = MO.local_u0()
    LR = LocalRanks(case, use_fused=True)
    LR.set_solution(u0)
    LR.exchange_boundaries()
    got = LR.get_solution()
    ref = [u.copy() for u in u0]
    MO.exchange(ref)
    for r in range(MO.nranks):
        assert np.array_equal(got[r], ref[r]), f"rank {r}: ghost layers differ after the exchange"
    # the plan the library reports moved what it says: messages and bytes sent by rank 0
    msgs, nbytes = LR.sv[0].comm_stats()
    plan = LR.sv[0].exchange_plan(0)
    sends = [p for p in plan if p[0] == "send"]
    assert msgs == len(sends) and nbytes == 8 * sum(p[3] for p in sends)
    LR.close()


# ------------------------------------------------------------------------------------------------------------------
# compact schemes with the grid lines split among ranks: TridiagLU/tridiagLU.c:84-274 with all four stages,
# the reduced system solved by tridiagIterJacobi.c as the reference does by default, and the hand-over of the shared
# interface (Interp1PrimFifthOrderCRWENO.c:206-223). The checker is the REAL reference running with the same number of ranks
# (oracle/_ref/hypar_ref_mp on the multi-process MPI shim): the multi-rank solve is decomposition-dependent by design (the
# reduced system is iterated to 1e-10), so only the same algorithm on the same decomposition can match -- bit for bit.
def _compact_cases():
    C = [cases.ns2d_vortex((40, 27), "mapped", iproc=(2, 2), scheme="crweno5"),
         cases.ns3d_turbulence((26, 25, 27), "z", iproc=(2, 2, 2), scheme="crweno5"),          # viscous, 8 ranks
         cases.ns3d_rising_bubble((14, 26, 12), "yc", iproc=(1, 2, 1), scheme="crweno5"),       # gravity source reconstructions
         cases.ns2d_vortex((28, 40), "js", scheme="cupw5", iproc=(1, 3)),
         cases.ns3d_density_wave((16, 12, 26), "js", iproc=(1, 1, 4), scheme="crweno5"),        # 4 ranks on one line
         cases.ns2d_vortex((40, 27), "z", iproc=(2, 2), scheme="hcweno5"),                      # hybrid compact-WENO5 across ranks
         cases.ns3d_rising_bubble((14, 26, 12), "mapped+rc0.2", iproc=(1, 3, 1), scheme="hcweno5"),
         # characteristic compact schemes: BLOCK tridiagonal systems across ranks (blocktridiagLU.c stages 1-4, block Jacobi)
         cases.ns2d_vortex((40, 27), "z", upwinding="roe", interp="characteristic", iproc=(2, 2), scheme="crweno5"),
         cases.with_characteristic(cases.ns3d_turbulence((14, 12, 26), "mapped", viscous=False, upwinding="roe", iproc=(1, 1, 2), scheme="crweno5")),
         cases.with_characteristic(cases.ns3d_turbulence((26, 12, 13), "js", upwinding="rf-char", iproc=(2, 1, 1), scheme="cupw5")),
         cases.with_characteristic(cases.ns3d_density_wave((12, 14, 39), "yc", iproc=(1, 1, 3), scheme="hcweno5"))]
    a = cases.linear_advection_sine(96, "z", scheme="crweno5")
    a.solver["iproc"] = [3]
    b = cases.euler1d_sod(101, "js", interp="components", upwinding="roe", scheme="cupw5")
    b.solver["iproc"] = [2]
    c = cases.euler1d_sod(121, "mapped", upwinding="roe", scheme="crweno5")              # characteristic, 1-D, three ranks
    c.solver["iproc"] = [3]
    C += [a, b, c]
    # lusolver.inp (tridiagLUInit.c:54-90): the Jacobi iteration on the reduced system with other limits -- four ranks on a
    # line, so that the iteration count matters (two ranks converge at the initial guess): a fixed count without the norm,
    # a loose relative tolerance (stops on the norm), the block systems with a fixed count
    for lu, tag in (({"maxiter": 1, "evaluate_norm": 0}, "lu1"), ({"maxiter": 10, "rtol": 1e-3}, "lurtol")):
        d = cases.ns2d_vortex((28, 40), "js", scheme="cupw5", iproc=(1, 4))
        d.lusolver = lu
        d.name += "_" + tag
        C.append(d)
    e = cases.with_characteristic(cases.ns3d_density_wave((12, 14, 39), "yc", iproc=(1, 1, 3), scheme="crweno5"))
    e.lusolver = {"maxiter": 2, "evaluate_norm": 0}
    e.name += "_lu2"
    C.append(e)
    for c in C:
        c.name += "_iproc" + "x".join(str(v) for v in c.solver["iproc"])
    return C


COMPACT_MR = _compact_cases()


@pytest.mark.parametrize("case", COMPACT_MR, ids=[c.name for c in COMPACT_MR])
@pytest.mark.parametrize("overlap", [False, True], ids=["serial", "overlap"])
def test_compact_schemes_across_ranks(need_gpu, case, overlap):
    from test_oracle_multirank_ref import EXE, run_ref_mp
    import os
    if not os.access(EXE, os.X_OK):
        pytest.fail(f"{EXE} is missing: build it where the reference tree is available (make -C oracle refmp)")
    nr = int(np.prod(case.solver["iproc"]))
    ref = run_ref_mp(case, "rhs", nranks=nr)
    MO = MultiRankOracle(case)                 # only for the initial blocks and their shapes
    LR = LocalRanks(case, use_fused=False, sweepwise=overlap)
    LR.set_solution(MO.local_u0())
    rhs = LR.rhs()
    for r in range(nr):
        a = ref[f"rhs.r{r:04d}"]["data"]
        assert np.isfinite(rhs[r]).all()
        assert np.array_equal(rhs[r].ravel(), a), \
            f"rank {r}: rhs differs from the {nr}-rank reference by {np.abs(rhs[r].ravel() - a).max():.3e} (max {np.abs(a).max():.3e})"
    ref = run_ref_mp(case, "steps", [2], nranks=nr)
    LR.set_solution(MO.local_u0())
    LR.time_step(2)
    u = LR.get_solution()
    for r in range(nr):
        S = MO.S[r]
        a = S.interior(ref[f"ufinal.r{r:04d}"]["data"].reshape(S.shape_g()))
        b = S.interior(u[r])
        assert np.array_equal(a, b), f"rank {r}: u after 2 steps differs from the {nr}-rank reference by {np.abs(a - b).max():.3e}"
    LR.close()
