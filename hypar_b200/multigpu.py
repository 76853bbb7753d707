"""One rank of a domain-decomposed run (one process per GPU).

The halo exchange of MPIExchangeBoundariesnD (reference src/MPIFunctions/MPIExchangeBoundariesnD.c:42-173) and the
whole time step around it run INSIDE the library (csrc/comm.cu, csrc/capi.cu: hpb_comm_init_nccl,
hpb_TimeStepsDistributed): pack kernels, ncclSend / ncclRecv on the library's communication stream, unpack kernels,
ordered by CUDA events -- no Python and no host synchronisation between the stages. This module only brings the
communicator up (the ncclUniqueId travels over torch.distributed, as it would over MPI_Bcast in HyPar) and offers the
scalar reductions of TimePreStep / TimePostStep.

Decomposition = HyPar's: ``iproc[d]`` blocks per dimension, remainder on the last block, rank =
ip0 + iproc0*(ip1 + iproc1*ip2); faces only (edges/corners are never exchanged, as in the reference).

Message matching. NCCL point-to-point has no tags: between one pair of ranks, sends and receives match in issue
order. The reference distinguishes the two messages of a pair with tags 1630/1631 (:95-100, :132-137); they matter
when iproc[d] == 2 with periodic boundaries, where the left and the right neighbour are the same peer. The library
issues, per dimension, ``send(low face)``, ``send(high face)``, ``recv(high ghost)``, ``recv(low ghost)``
(``hpb_exchange_plan``; tests/test_multigpu_gloo.py runs that plan over gloo on the CPU).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .solver import Solver, comm_unique_id


def bind_to_gpu_numa(device: int):
    """Pin this process to the CPUs NVML reports as local to CUDA device `device` (one process per GPU), so that the
    pinned host buffers it allocates afterwards are placed on that GPU's NUMA node: with 8 ranks staging 11 GB per step
    each, buffers that all sit on one socket share its memory controllers and the inter-socket link. Returns the CPU
    set used, or None when the topology is not visible (containers with a restricted CPU set, no NVML)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device), "pci_domain_id", 0)
        dev_id = torch.cuda.get_device_properties(device).pci_device_id
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev_id:02x}.0".encode())
        ncpu = os.cpu_count() or 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i * 64 + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        use = cpus & os.sched_getaffinity(0)
        if not use or use == os.sched_getaffinity(0):
            return None
        os.sched_setaffinity(0, use)
        return sorted(use)
    except Exception:
        return None


def exchange_ops(neighbors: Sequence[int], dims: Optional[Sequence[int]] = None):
    """The ordered list of point-to-point operations of one face exchange, as csrc/comm.cu issues them
    (``Solver.exchange_plan`` returns the library's own list for a solver; this is the same rule for a bare
    neighbour table): [("send" | "recv", face index 2*d + side, peer rank)], side 0 = low, 1 = high."""
    nd = len(neighbors) // 2
    ops = []
    for d in (range(nd) if dims is None else dims):
        lo, hi = neighbors[2 * d], neighbors[2 * d + 1]
        if lo >= 0:
            ops.append(("send", 2 * d, lo))
        if hi >= 0:
            ops.append(("send", 2 * d + 1, hi))
        if hi >= 0:
            ops.append(("recv", 2 * d + 1, hi))
        if lo >= 0:
            ops.append(("recv", 2 * d, lo))
    return ops


class HaloExchanger:
    """A face exchange over torch.distributed between host (gloo) or device (NCCL) tensors, following a plan of
    (kind, face, peer) operations. TEST HELPER of the N > 1 path on the CPU (tests/test_multigpu_gloo.py): the product
    exchanges inside the library (hpb_TimeStepsDistributed)."""

    def __init__(self, neighbors: Sequence[int], send: Sequence, recv: Sequence, group=None, plan=None):
        self.neighbors, self.send, self.recv, self.group = list(neighbors), list(send), list(recv), group
        self.plan = plan

    def start(self, dims: Optional[Sequence[int]] = None):
        import torch.distributed as dist
        p2p = []
        plan = self.plan if (self.plan is not None and dims is None) else exchange_ops(self.neighbors, dims)
        for op in plan:
            kind, face, peer = op[0], op[1], op[2]
            if kind == "send":
                p2p.append(dist.P2POp(dist.isend, self.send[face], peer, self.group))
            else:
                p2p.append(dist.P2POp(dist.irecv, self.recv[face], peer, self.group))
        return dist.batch_isend_irecv(p2p) if p2p else []

    @staticmethod
    def finish(works) -> None:
        for w in works:
            w.wait()

    def exchange(self, dims: Optional[Sequence[int]] = None) -> None:
        self.finish(self.start(dims))


class DistributedSolver:
    """One rank of a decomposed run: a Solver whose NCCL transport is up (``hpb_comm_init_nccl``); the steps are
    ``hpb_TimeStepsDistributed`` -- the stage loop, the pack / unpack kernels and the ncclSend / ncclRecv calls all
    live in the library.

    ``overlap`` selects the library's schedule (``hpb_set_overlap``): True (default) = the exchange of u travels under
    the full-array RK update (face layers of the stage vector first, straight into the send buffers), the Q-derivative
    exchange of dimensions 1.. under the x-sweep; False = pack - exchange - unpack in sequence. Bit-identical results.
    """

    def __init__(self, solver_inp, boundary, physics, weno, x, rank: int, device: int, group=None,
                 use_fused: bool = True, overlap: bool = True, muscl=None, advection_field=None, glm_gee=None, lusolver=None):
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.solver = Solver(solver_inp, boundary, physics, weno, x, rank=rank, device=device, use_fused=use_fused,
                             muscl=muscl, advection_field=advection_field, glm_gee=glm_gee, lusolver=lusolver)
        sv = self.solver
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.ExternalStream(sv.stream, device=self.device)
        self.viscous = bool(sv.L.hpb_needs_viscous_exchange(sv.h))
        self.group = group
        nranks = int(np.prod(sv.iproc))
        if dist.get_world_size(group) != nranks:
            raise RuntimeError(f"iproc {sv.iproc} needs {nranks} ranks, the process group has {dist.get_world_size(group)}")
        # ncclUniqueId: made by rank 0 of the group through the library, broadcast as it would be over MPI_Bcast
        uid = [comm_unique_id() if dist.get_rank(group) == 0 else None]
        dist.broadcast_object_list(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        sv.comm_init_nccl(uid[0], nranks)
        sv.set_overlap(overlap)
        self.overlap = bool(overlap)
        self.sweepwise = self.overlap and bool(sv.L.hpb_stage_overlap_supported(sv.h))

    def time_step(self) -> None:
        """TimePreStep (BCs + halo on u) and TimeRK (TimeRK.c:126-195), one step; only enqueues."""
        self.solver.TimeStepsDistributed(1)

    def time_steps(self, n: int) -> None:
        self.solver.TimeStepsDistributed(n)

    def time_integrate_host(self, u_host: np.ndarray, nsteps: int = 1) -> np.ndarray:
        """TimeIntegrate on a host array in HyPar's layout (this rank's block with ghosts): H2D, steps, D2H."""
        self.solver.set_solution(u_host)
        self.time_steps(nsteps)
        self.solver._ck(self.solver.L.hpb_dev_get_solution(self.solver.h,
                                                           u_host.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double))))
        return u_host

    def time_integrate_host_async(self, u_in: np.ndarray, u_out: np.ndarray, nsteps: int = 1) -> None:
        """The same for a sequence of independent fields (ensembles): only enqueues -- the H2D copy of this field
        overlaps the steps of the previous one and the D2H copy of the one before (hpb_pipe_*). Finish with
        ``solver.pipe_wait()``."""
        self.solver.pipe_upload(u_in, self.solver.time)
        self.time_steps(nsteps)
        self.solver.pipe_download(u_out)

    def rhs(self, want: bool = True):
        """One TimeRHSFunctionExplicit of the device solution (stage 0 buffers); returns this rank's rhs."""
        self.solver.RHSFunctionDistributed()
        return self.solver.get_stage_rhs(0) if want else None

    # scalar reductions of TimePreStep.c:81-107 / TimePostStep.c:44-63
    def max_cfl(self) -> float:
        import torch.distributed as dist
        t = self.torch.tensor([self.solver.dev_ComputeCFL()], device=self.device, dtype=self.torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    # conservation diagnostics of TimePostStep.c:81-93 across the ranks: each rank reduces its own block on the device,
    # the nvars-long partial results are summed where the reference calls MPISum_double (VolumeIntegral.c:44,
    # BoundaryIntegral.c:52); the flux integrals of internal faces cancel between neighbours
    def _allreduce_sum(self, a: np.ndarray) -> np.ndarray:
        import torch.distributed as dist
        t = self.torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def volume_integral(self) -> np.ndarray:
        return self._allreduce_sum(self.solver.dev_VolumeIntegral())

    def boundary_integral(self) -> np.ndarray:
        """global boundary-flux integral of the LAST step (to be added to TotalBoundaryIntegral)"""
        return self._allreduce_sum(self.solver.BoundaryIntegral(self.solver.dev_StepBoundaryIntegral()))

    def error_norms(self, uex_local: np.ndarray):
        """CalculateError.c:26-124: (L1, L2, Linf) of u - uex, relative to the norms of uex when those are not tiny"""
        import torch.distributed as dist
        s = self.solver.dev_ErrorSums(uex_local)
        sums = self._allreduce_sum(s[[0, 1, 3, 4]])
        t = self.torch.as_tensor(s[[2, 5]], device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        mx = t.cpu().numpy()
        npts = float(np.prod(self.solver.dim_global))
        sol = [sums[0] / npts, np.sqrt(sums[1] / npts), mx[0]]
        err = [sums[2] / npts, np.sqrt(sums[3] / npts), mx[1]]
        if all(v > 1e-15 for v in sol):
            err = [e / n for e, n in zip(err, sol)]
        return err

    def step_norm(self) -> float:
        import torch.distributed as dist
        t = self.torch.tensor([self.solver.dev_StepNormSumSq()], device=self.device, dtype=self.torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        npts = float(np.prod(self.solver.dim_global))
        return float(np.sqrt(t.item() / npts))
