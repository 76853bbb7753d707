"""N > 1 path on CPU: world_size-2 (and 4) gloo process groups execute the message plan the library issues through
NCCL (hpb_exchange_plan, csrc/comm.cu) on host buffers packed/unpacked by the oracle.
Checked: every rank's ghost faces equal the neighbour's interior layers of the GLOBAL array, including the
same-peer case (iproc = 2 with periodic boundaries: two messages each way between one pair of ranks, matched
by issue order -- the reference needs tags 1630/1631 for this, MPIExchangeBoundariesnD.c:95-137), remainder
partitions and non-periodic physical faces; and the decomposed oracle RHS equals the single-rank oracle RHS
on the inviscid path (decomposition invariance of the hyperbolic term).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_with_periodic_ghost(u0, lo, n, g, periodic):
    """slice [lo-g, lo+n+g) of a global axis with periodic wrap (or zeros outside when not periodic)"""
    N = u0.shape[0]
    idx = np.arange(lo - g, lo + n + g)
    if periodic:
        return u0[idx % N]
    out = np.zeros((len(idx),) + u0.shape[1:])
    ok = (idx >= 0) & (idx < N)
    out[ok] = u0[idx[ok]]
    return out


def _worker(rank, world, port, builder, kwargs, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from hypar_b200 import cases
        from hypar_b200.multigpu import HaloExchanger, exchange_ops
        from hypar_b200.solver import Solver
        from oracle import hpo
        case = getattr(cases, builder)(**kwargs)
        S = hpo.Setup(case, rank=rank)
        O = hpo.Oracle(S)
        sv = Solver.from_case(case, rank=rank)          # host set-up only (no GPU needed)
        nd, g = S.ndims, S.ghosts
        u = S.local_u0()
        O.apply_bc(u)                                    # physical faces (periodic only when iproc == 1)
        send, recv = [None] * (2 * nd), [None] * (2 * nd)
        for d in range(nd):
            for side in (0, 1):
                if sv.neighbors[2 * d + side] >= 0:
                    send[2 * d + side] = torch.from_numpy(O.pack(u, d, side))
                    recv[2 * d + side] = torch.zeros_like(send[2 * d + side])
        # the plan under test is the LIBRARY's (hpb_exchange_plan: what hpb_TimeStepsDistributed issues through NCCL)
        plan = sv.exchange_plan(0)
        ops = exchange_ops(sv.neighbors)
        assert [p[:3] for p in plan] == ops, f"library plan {plan} != {ops}"
        for kind, face, peer, count in plan:
            assert count == send[face].numel(), f"face {face}: the library sends {count} doubles, the oracle packs {send[face].numel()}"
        ex = HaloExchanger(sv.neighbors, send, recv, plan=plan)
        ex.exchange()
        for d in range(nd):
            for side in (0, 1):
                if recv[2 * d + side] is not None:
                    O.unpack(u, d, side, recv[2 * d + side].numpy())
        # expected: face ghosts = the global array (periodic wrap), interior = own block
        ug = case.u0
        per = [any(z["type"] == "periodic" and z["dim"] == d for z in case.boundary) for d in range(nd)]
        got = u.reshape(S.shape_g())
        err = 0.0
        for d in range(nd):
            ax = nd - 1 - d                              # numpy axis of dimension d
            # take the block's interior range on the other axes, ghost range on this axis
            sl_g = [slice(S.is_[k], S.is_[k] + S.dim[k]) for k in reversed(range(nd))]
            blk = ug
            for k in range(nd):
                a = nd - 1 - k
                if k == d:
                    continue
                blk = np.take(blk, np.arange(S.is_[k], S.is_[k] + S.dim[k]), axis=a)
            blk = np.moveaxis(blk, ax, 0)
            exp = _global_with_periodic_ghost(blk, S.is_[d], S.dim[d], g, per[d])
            exp = np.moveaxis(exp, 0, ax)
            sl = [slice(g, g + S.dim[k]) for k in reversed(range(nd))]
            sl[ax] = slice(0, S.dim[d] + 2 * g)
            have = got[tuple(sl)]
            if not per[d]:
                # physical non-periodic faces are filled by the boundary condition, not by the exchange
                lo_phys, hi_phys = S.ip[d] == 0, S.ip[d] == S.iproc[d] - 1
                cut = [slice(None)] * have.ndim
                cut[ax] = slice(g if lo_phys else 0, S.dim[d] + (g if hi_phys else 2 * g))
                have, exp = have[tuple(cut)], exp[tuple(cut)]
            err = max(err, float(np.abs(have - exp).max()))
        # decomposed RHS (inviscid cases): equals the single-rank RHS on this block, bit for bit
        rhs_err = None
        if float(case.physics.get("Re", -1.0)) <= 0:
            rhs = O.rhs(u.copy())                        # u already carries BC + halo; rhs re-applies the BCs
            case1 = getattr(cases, builder)(**{**kwargs, "iproc": None})
            S1 = hpo.Setup(case1)
            r1 = hpo.Oracle(S1).rhs(S1.local_u0()).reshape(S1.shape_g())
            sl1 = tuple(slice(g + S.is_[k], g + S.is_[k] + S.dim[k]) for k in reversed(range(nd)))
            rhs_err = float(np.abs(S.interior(rhs) - r1[sl1]).max())
        q.put((rank, err, rhs_err, len(ops), None))
        dist.destroy_process_group()
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((rank, None, None, None, traceback.format_exc()))


def _run(world, builder, kwargs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, builder, kwargs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, err, rhs_err, nops, tb in res:
        assert tb is None, f"rank {rank} failed:\n{tb}"
        assert err == 0.0, f"rank {rank}: ghost faces differ from the global array by {err}"
        if rhs_err is not None:
            assert rhs_err == 0.0, f"rank {rank}: decomposed RHS differs from the single-rank RHS by {rhs_err}"
    return res


def test_exchange_plan_orders_same_peer_messages():
    from hypar_b200.multigpu import exchange_ops
    # iproc = 2, periodic: both neighbours are rank 1
    ops = exchange_ops([1, 1])
    assert ops == [("send", 0, 1), ("send", 1, 1), ("recv", 1, 1), ("recv", 0, 1)]
    # no neighbour on the low side (physical boundary)
    assert exchange_ops([-1, 3]) == [("send", 1, 3), ("recv", 1, 3)]
    assert exchange_ops([2, -1, -1, -1]) == [("send", 0, 2), ("recv", 0, 2)]
    assert exchange_ops([0, 1, 2, 3], dims=[1]) == [("send", 2, 2), ("send", 3, 3), ("recv", 3, 3), ("recv", 2, 2)]


def test_world2_periodic_same_peer_3d():
    # 2 ranks along z, periodic: left and right neighbour coincide; remainder partition (25 = 12 + 13)
    _run(2, "ns3d_turbulence", dict(n=(12, 10, 25), weno="js", viscous=False, iproc=(1, 1, 2)))


def test_world2_periodic_2d_split_x():
    _run(2, "ns2d_vortex", dict(n=(26, 16), weno="mapped", iproc=(2, 1)))


def test_world2_slip_walls_with_gravity():
    # non-periodic: physical faces are not exchanged; one internal face only
    _run(2, "ns3d_rising_bubble", dict(n=(12, 16, 10), weno="yc", iproc=(1, 2, 1)))


def test_world4_periodic_2x2():
    _run(4, "ns3d_turbulence", dict(n=(14, 13, 12), weno="z", viscous=False, iproc=(2, 2, 1)))
