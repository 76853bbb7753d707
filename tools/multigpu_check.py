"""2..8-GPU parity check over NCCL (TEST INFRASTRUCTURE; run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py

Every rank drives hypar_b200.multigpu.DistributedSolver on its block of a small decomposed case and compares
its RHS and its solution after 2 steps with the multi-rank oracle (all ranks evaluated in-process)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

from _multirank import MultiRankOracle
from hypar_b200 import cases
from hypar_b200.multigpu import DistributedSolver
from oracle import hpo


def diagnostics_and_io(rank, world, local, iproc):
    """conservation bookkeeping summed over the ranks, error norms, partitioned files -- against a single-rank
    oracle run of the same (inviscid, decomposition-invariant) case"""
    import tempfile
    from hypar_b200 import hypario as H
    case = cases.ns3d_density_wave((26, 24, 28), "js", iproc=iproc)
    case.solver["conservation_check"] = "yes"
    one = cases.ns3d_density_wave((26, 24, 28), "js")
    S1 = hpo.Setup(one)
    O1 = hpo.Oracle(S1)
    dt, rk = float(case.solver["dt"]), hpo.rk_type_of(case)
    d = [tempfile.mkdtemp(prefix="hpb_mg_") if rank == 0 else None]
    dist.broadcast_object_list(d, src=0)
    os.chdir(d[0])
    dg, nv = list(case.solver["size"]), 5
    if rank == 0:
        H.write_initial_bin("initial.inp", case.x, case.u0)
        H.serial_to_parallel("initial.inp", "initial", dg, list(iproc), nv, 2 if world % 2 == 0 else 1)
    dist.barrier()
    ds = DistributedSolver(case.solver, case.boundary, case.physics, case.weno, case.x, rank=rank, device=local,
                           use_fused=False)
    ds.solver.load_solution_parallel("initial", 2 if world % 2 == 0 else 1)
    vol0 = ds.volume_integral()
    u1 = S1.local_u0()
    vol0_ref = O1.volume_integral(u1)
    ok = bool(np.all(np.abs(vol0 - vol0_ref) <= 1e-13 * np.abs(vol0_ref).max()))
    tbi, tbi_ref = np.zeros(nv), np.zeros(nv)
    for step in range(2):
        sbi_ref = O1.time_step_cons(u1, dt, rk)
        tbi_ref += O1.boundary_integral(sbi_ref)
        ds.time_step()
        tbi += ds.boundary_integral()
        ds.solver.write_solution_parallel("op.bin", 1, record=step)
    vol = ds.volume_integral()
    err = ds.solver.CalculateConservationError(vol, vol0, tbi)
    uex = MultiRankOracle(case).local_u0()[rank]
    norms = ds.error_norms(uex)
    n0, e0 = O1.norm_sums(S1.local_u0()), O1.norm_sums(S1.local_u0(), u1)
    npts = float(np.prod(dg))
    norms_ref = [e0[0] / npts / (n0[0] / npts), np.sqrt(e0[1] / npts) / np.sqrt(n0[1] / npts), e0[2] / n0[2]]
    ok = ok and bool(np.all(err <= 1e-13)) and all(abs(a - b) <= 1e-10 * abs(b) for a, b in zip(norms, norms_ref))
    dist.barrier()
    if rank == 0:
        xs, us = H.parallel_to_serial("op.bin", dg, list(iproc), nv, 1, record=1)
        same = np.array_equal(us, S1.interior(u1))          # exact path, inviscid: bit-identical to the single-rank run
        ok = ok and same
        print(f"[diagnostics, {world} ranks, iproc {iproc}] conservation error {err.max():.2e}, total boundary integral "
              f"{np.abs(tbi).max():.2e} (ref {np.abs(tbi_ref).max():.2e}), error norms {['%.6e' % v for v in norms]} "
              f"(ref {['%.6e' % v for v in norms_ref]}), partitioned output == single-rank solution: {same} "
              f"{'ok' if ok else 'FAIL'}", flush=True)
    ds.solver.close()
    return ok


def compact_across_ranks(rank, world, local, iproc):
    """compact scheme with its grid lines split among the ranks (tridiagLU's four stages + Jacobi reduced solve over the
    library's NCCL line exchanges) against the REAL reference running with the same ranks (oracle/_ref/hypar_ref_mp)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_multirank_ref import EXE, run_ref_mp
    if not os.access(EXE, os.X_OK):
        if rank == 0:
            print("[compact] oracle/_ref/hypar_ref_mp missing: skipped", flush=True)
        return True
    ok = True
    for name, case in (("crweno5 visc", cases.ns3d_turbulence((26, 25, 27), "z", iproc=iproc, scheme="crweno5")),
                       ("cupw5 bubble", cases.ns3d_rising_bubble((26, 24, 28), "yc", iproc=iproc, scheme="cupw5")),
                       # characteristic: block tridiagonal systems across the ranks (blocktridiagLU + block Jacobi)
                       ("crweno5 char roe", cases.with_characteristic(cases.ns3d_turbulence((26, 25, 27), "mapped", viscous=False,
                                                                                          upwinding="roe", iproc=iproc, scheme="crweno5"))),
                       ("hcweno5", cases.ns3d_density_wave((26, 24, 28), "js", iproc=iproc, scheme="hcweno5"))):
        ref = run_ref_mp(case, "steps", [2], nranks=world)
        ds = DistributedSolver(case.solver, case.boundary, case.physics, case.weno, case.x, rank=rank, device=local, use_fused=False)
        MO = MultiRankOracle(case)
        ds.solver.set_solution(MO.local_u0()[rank])
        ds.time_steps(2)
        u = ds.solver.get_solution()
        S = MO.S[rank]
        a = S.interior(ref[f"ufinal.r{rank:04d}"]["data"].reshape(S.shape_g()))
        b = S.interior(u)
        good = bool(np.array_equal(a, b))
        ok = ok and good
        print(f"[rank {rank}/{world}] iproc {iproc} {name}: u(2 steps) vs the {world}-rank reference: max diff {np.abs(a - b).max():.2e} "
              f"{'bit-identical ok' if good else 'FAIL'}", flush=True)
        ds.solver.close()
    return ok


def glmgee_over_nccl(rank, world, local, iproc):
    """GLM-GEE (TimeGLMGEE.c) through hpb_TimeStepsDistributed: solution and auxiliary solution against the decomposed oracle"""
    ok = True
    case = cases.with_glmgee(cases.ns3d_turbulence((26, 25, 27), "mapped", iproc=iproc), "exrk2a", "yyt")
    m, mode = hpo.glmgee_of(case)
    MO = MultiRankOracle(case)
    u = MO.local_u0()
    ua = [MO.O[r].glmgee_aux0(u[r], mode) for r in range(world)]
    for _ in range(2):
        MO.time_step_glmgee(u, ua, float(case.solver["dt"]), m, mode)
    S = MO.S[rank]
    for fused in (False, True):
        ds = DistributedSolver(case.solver, case.boundary, case.physics, case.weno, case.x, rank=rank, device=local,
                               use_fused=fused, glm_gee=case.glm_gee)
        ds.solver.set_solution(MO.local_u0()[rank])
        ds.time_steps(2)
        a, b = S.interior(ds.solver.get_solution()), S.interior(ds.solver.get_aux_solution())
        ra, rb = S.interior(u[rank]), S.interior(ua[rank])
        e = max(np.abs(a - ra).max(), np.abs(b - rb).max()) / np.abs(ra).max()
        good = (e <= 1e-11) if fused else (e <= 1e-14)       # exact path: bit-identical but for the viscosity law's libm ulp
        ok = ok and good
        print(f"[rank {rank}/{world}] iproc {iproc} glm-gee exrk2a yyt {'fused' if fused else 'exact'}: u and aux (2 steps) rel err "
              f"{e:.2e} {'ok' if good else 'FAIL'}", flush=True)
        ds.solver.close()
    return ok


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    iprocs = {2: [(1, 1, 2), (2, 1, 1)], 4: [(1, 2, 2), (2, 2, 1)], 8: [(2, 2, 2)]}[world]
    ok = True
    for iproc in iprocs:
        for name, case in (("visc", cases.ns3d_turbulence((26, 25, 27), "mapped", iproc=iproc)),
                           ("bubble", cases.ns3d_rising_bubble((26, 24, 28), "yc", iproc=iproc))):
            keep = {}
            for fused, overlap in ((False, False), (False, True), (True, False), (True, True)):
                MO = MultiRankOracle(case)
                u_ref = MO.local_u0()
                rhs_ref = MO.rhs(u_ref)
                ds = DistributedSolver(case.solver, case.boundary, case.physics, case.weno, case.x, rank=rank,
                                       device=local, use_fused=fused, overlap=overlap)
                ds.solver.set_solution(MO.local_u0()[rank])
                rhs = ds.rhs()
                scale = max(np.abs(r).max() for r in rhs_ref)
                e_rhs = np.abs(rhs - rhs_ref[rank]).max() / scale
                dt = float(case.solver["dt"])
                rk = hpo.rk_type_of(case)
                u_ref = MO.local_u0()
                for _ in range(2):
                    MO.time_step(u_ref, dt, rk)
                ds.solver.set_solution(MO.local_u0()[rank])
                ds.time_steps(2)
                u = ds.solver.get_solution()
                S = MO.S[rank]
                a, b = S.interior(u), S.interior(u_ref[rank])
                e_u = np.abs(a - b).max() / np.abs(b).max()
                cfl = ds.max_cfl()
                # exact path: bit-identical. Fused path: 1e-12 of the largest RHS plus 16 ulp of the dissipation term
                # alpha*u*dxinv (tests/test_gpu_parity.py::fused_tolerance); the hydrostatically balanced bubble has
                # |rhs| << |terms|, so its error is measured against that floor
                lam = MO.O[rank].cfl(MO.local_u0()[rank], dt) / dt
                tol = 1e-12 + 16 * np.finfo(np.float64).eps * lam * np.abs(MO.local_u0()[rank]).max() / scale
                good = (e_rhs == 0 and e_u == 0) if not fused else (e_rhs <= tol and e_u <= 1e-11)
                # the two schedules of the library's step must agree bit for bit
                # (interiors of u: the overlapped schedule has already exchanged the face ghosts for the next step)
                if overlap:
                    good = good and np.array_equal(keep[fused][0], rhs) and np.array_equal(keep[fused][1], a)
                else:
                    keep[fused] = (rhs.copy(), a.copy())
                ok = ok and good
                print(f"[rank {rank}/{world}] iproc {iproc} {name:6s} {'fused' if fused else 'exact'} "
                      f"{'overlapped' if overlap else 'serial    '}: "
                      f"rhs err/scale {e_rhs:.2e}, u(2 steps) rel err {e_u:.2e}, max CFL {cfl:.4f} {'ok' if good else 'FAIL'}",
                      flush=True)
                ds.solver.close()
    ok = diagnostics_and_io(rank, world, local, iprocs[0]) and ok
    ok = compact_across_ranks(rank, world, local, iprocs[0]) and ok
    ok = glmgee_over_nccl(rank, world, local, iprocs[0]) and ok
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    if int(t.item()):
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU CHECK PASSED")


if __name__ == "__main__":
    main()
