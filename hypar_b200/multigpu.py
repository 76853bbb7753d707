"""One rank of a domain-decomposed run: the halo exchange of MPIExchangeBoundariesnD
(reference src/MPIFunctions/MPIExchangeBoundariesnD.c:42-173) over torch.distributed (NCCL on the
GPUs of one NVSwitch box; gloo in the CPU tests), wrapped around the staged C-ABI step
(``hpb_step_begin`` ... ``hpb_step_finish``, include/hypar_b200.h).

Decomposition = HyPar's: ``iproc[d]`` blocks per dimension, remainder on the last block, rank =
ip0 + iproc0*(ip1 + iproc1*ip2); faces only (edges/corners are never exchanged, as in the reference).
The path has exactly one real exchange step per field -- no other collective is on the data path;
CFL / norm reductions are scalar all-reduces outside the hot loop.

Message matching. NCCL point-to-point has no tags: between one pair of ranks, sends and receives
match in issue order. The reference distinguishes the two messages of a pair with tags 1630/1631
(:95-100, :132-137); they matter when iproc[d] == 2 with periodic boundaries, where the left and the
right neighbour are the same peer. Here every rank issues, per dimension, ``send(low face)``,
``send(high face)`` and ``recv(high ghost)``, ``recv(low ghost)`` -- the peer's low-face send is the
first message it sends us and lands in our high ghost, its high-face send is the second.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .solver import FIELD_QDERIVX, FIELD_QDERIVY, FIELD_U, Solver


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
    """The ordered list of point-to-point operations of one face exchange:
    [("send" | "recv", face index 2*d + side, peer rank)], side 0 = low, 1 = high."""
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
    """Face exchange between persistent send/receive buffers (torch tensors: CUDA for NCCL, CPU for gloo)."""

    def __init__(self, neighbors: Sequence[int], send: Sequence, recv: Sequence, group=None):
        self.neighbors, self.send, self.recv, self.group = list(neighbors), list(send), list(recv), group

    def start(self, dims: Optional[Sequence[int]] = None):
        import torch.distributed as dist
        p2p = []
        for kind, face, peer in exchange_ops(self.neighbors, dims):
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


class _DevBuf:
    """a raw device allocation seen through __cuda_array_interface__ (no copy, no ownership)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class DistributedSolver:
    """Drives the staged step of one rank and exchanges its halo buffers.

    The library's kernels run on the solver's own CUDA stream. Serial schedule (``overlap=False``, the default, and
    every configuration the library cannot drive sweep by sweep): that stream is made torch's current stream
    around every exchange, so NCCL's send/recv are ordered after the pack kernels and before the unpack kernels
    without any host synchronisation. Overlapped schedule (``overlap=True``): the exchanges are issued on a second
    (communication) stream, dimension by dimension, ordered against the compute stream with CUDA events only:

      u halos (all dimensions)            ||  Q-derivatives of the deep interior        (viscous)
      Q-derivative halos of dimension d+1 ||  sweep d
      u halos of dimension d+1            ||  sweep d                                   (inviscid)

    Sweep d reads the halos of dimension d only (faces, never edges/corners: MPIExchangeBoundariesnD.c:60-76),
    so the result is identical to the serial schedule (tests/test_gpu_decomposed.py: bit for bit).

    Measured on 8 B200 (C4, 512^3 per GPU, 2x2x2; profiles/r01f_bench8*.json): serial 193.0 ms/step (95.5 % of the
    single-GPU rate), overlapped 201.6 ms/step. The exchange itself is ~1 ms of a 48 ms stage (NVLink 5), the
    pack/unpack kernels another ~0.9 ms; splitting the derivative kernel into deep interior + six thin shell
    boxes and running NCCL's copy kernels next to FP64-saturated SMs costs more than the ~1 ms it hides. The
    overlapped schedule is kept for slower links / smaller blocks; the serial one is the default because it is
    faster here.
    """

    def __init__(self, solver_inp, boundary, physics, weno, x, rank: int, device: int, group=None,
                 use_fused: bool = True, overlap: bool = False, muscl=None, advection_field=None):
        import torch
        self.torch = torch
        self.solver = Solver(solver_inp, boundary, physics, weno, x, rank=rank, device=device, use_fused=use_fused,
                             muscl=muscl, advection_field=advection_field)
        sv = self.solver
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.ExternalStream(sv.stream, device=self.device)
        self.viscous = bool(sv.L.hpb_needs_viscous_exchange(sv.h))
        self.ex = {}
        fields = [FIELD_U] + ([FIELD_QDERIVX, FIELD_QDERIVY] if self.viscous else [])
        for f in fields:
            send, recv, nbytes = sv.halo_buffers(f)
            st = [torch.as_tensor(_DevBuf(p, n), device=self.device) if (p and n) else None for p, n in zip(send, nbytes)]
            rt = [torch.as_tensor(_DevBuf(p, n), device=self.device) if (p and n) else None for p, n in zip(recv, nbytes)]
            self.ex[f] = HaloExchanger(sv.neighbors, st, rt, group)
        self.group = group
        self.overlap = bool(overlap) and bool(sv.L.hpb_stage_overlap_supported(sv.h))
        if self.overlap:
            self.comm_stream = torch.cuda.Stream(device=self.device)
            self._ev = [torch.cuda.Event() for _ in range(8)]

    # ---- overlapped schedule -------------------------------------------------------------------------------
    def _exchange_async(self, fields, dims, ev_ready, ev_done) -> None:
        """issue the exchange of `fields` restricted to `dims` on the communication stream: it starts when the compute
        stream has reached `ev_ready` (already recorded) and records ev_done[d] after the messages of dimension d"""
        torch = self.torch
        self.comm_stream.wait_event(ev_ready)
        with torch.cuda.stream(self.comm_stream):
            for d in dims:
                works = []
                for f in fields:
                    works += self.ex[f].start([d])
                HaloExchanger.finish(works)
                ev_done[d].record(self.comm_stream)

    def _stage_overlapped(self, s: int) -> None:
        sv, L = self.solver, self.solver.L
        nd = sv.ndims
        dims = list(range(nd))
        ev_pack, ev_pack2, ev_dim = self._ev[0], self._ev[1], self._ev[2:2 + nd]
        sv._ck(L.hpb_stage_begin(sv.h, s))                       # stage vector, BCs, pack u
        ev_pack.record(self.stream)
        if self.viscous:
            self._exchange_async([FIELD_U], dims, ev_pack, ev_dim)
            sv._ck(L.hpb_stage_interior(sv.h, s))                # || exchange of u
            self.stream.wait_event(ev_dim[nd - 1])
            sv._ck(L.hpb_stage_halo_done(sv.h, FIELD_U))
            sv._ck(L.hpb_stage_rhs_a(sv.h, s))                   # shell + ghost slabs of the Q-derivatives, pack
            ev_pack2.record(self.stream)
            self._exchange_async([FIELD_QDERIVX, FIELD_QDERIVY], dims, ev_pack2, ev_dim)
            for d in dims:
                self.stream.wait_event(ev_dim[d])
                sv._ck(L.hpb_stage_halo_done_dim(sv.h, FIELD_QDERIVX, d))
                sv._ck(L.hpb_stage_halo_done_dim(sv.h, FIELD_QDERIVY, d))
                sv._ck(L.hpb_stage_sweep(sv.h, s, d))            # || exchange of dimensions d+1..
        else:
            self._exchange_async([FIELD_U], dims, ev_pack, ev_dim)
            for d in dims:
                self.stream.wait_event(ev_dim[d])
                sv._ck(L.hpb_stage_halo_done_dim(sv.h, FIELD_U, d))
                sv._ck(L.hpb_stage_sweep(sv.h, s, d))

    def _exchange(self, fields) -> None:
        with self.torch.cuda.stream(self.stream):
            works = []
            for f in fields:
                works += self.ex[f].start()
            HaloExchanger.finish(works)

    def time_step(self) -> None:
        """TimePreStep (BCs + halo on u) and TimeRK (TimeRK.c:126-195), one step."""
        sv, L = self.solver, self.solver.L
        sv._ck(L.hpb_step_begin(sv.h))
        self._exchange([FIELD_U])
        sv._ck(L.hpb_step_halo_done(sv.h))
        for s in range(sv.nstages):
            if self.overlap:
                self._stage_overlapped(s)
                continue
            sv._ck(L.hpb_stage_begin(sv.h, s))
            self._exchange([FIELD_U])
            sv._ck(L.hpb_stage_halo_done(sv.h, FIELD_U))
            sv._ck(L.hpb_stage_rhs_a(sv.h, s))
            if self.viscous:
                self._exchange([FIELD_QDERIVX, FIELD_QDERIVY])
                sv._ck(L.hpb_stage_halo_done(sv.h, FIELD_QDERIVX))
                sv._ck(L.hpb_stage_halo_done(sv.h, FIELD_QDERIVY))
            sv._ck(L.hpb_stage_rhs_b(sv.h, s))
        sv._ck(L.hpb_step_finish(sv.h))

    def time_steps(self, n: int) -> None:
        for _ in range(n):
            self.time_step()

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
        sv, L = self.solver, self.solver.L
        if self.overlap:
            self._stage_overlapped(0)
            return sv.get_stage_rhs(0) if want else None
        sv._ck(L.hpb_stage_begin(sv.h, 0))
        self._exchange([FIELD_U])
        sv._ck(L.hpb_stage_halo_done(sv.h, FIELD_U))
        sv._ck(L.hpb_stage_rhs_a(sv.h, 0))
        if self.viscous:
            self._exchange([FIELD_QDERIVX, FIELD_QDERIVY])
            sv._ck(L.hpb_stage_halo_done(sv.h, FIELD_QDERIVX))
            sv._ck(L.hpb_stage_halo_done(sv.h, FIELD_QDERIVY))
        sv._ck(L.hpb_stage_rhs_b(sv.h, 0))
        return sv.get_stage_rhs(0) if want else None

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
