"""All ranks of a decomposed run inside ONE process (TEST INFRASTRUCTURE).

MultiRankOracle : one oracle context per rank; the halo exchange of MPIExchangeBoundariesnD is a copy
                  of packed faces between the ranks' arrays. Gives the reference result of a run with
                  the SAME iproc -- needed because the viscous term depends on the decomposition
                  (QDerivZ is never exchanged, un-exchanged ghost derivatives
                  are one-sided).
LocalRanks      : one hpb_solver per rank on ONE GPU, driven through the staged C-ABI step; the halo
                  buffers are copied device-to-device instead of sent over NCCL. Exercises exactly the
                  pack / unpack / staged-step code the multi-GPU run uses.
"""
from __future__ import annotations

import numpy as np

from hypar_b200.solver import FIELD_QDERIVX, FIELD_QDERIVY, FIELD_U, Solver
from oracle import hpo


def _neighbors(S):
    """[2*nd] neighbour ranks of hpo.Setup S (-1: none), MPIExchangeBoundariesnD.c:65-76."""
    nb = []
    for d in range(S.ndims):
        for side in (0, 1):
            ip = list(S.ip)
            edge = (S.ip[d] == 0) if side == 0 else (S.ip[d] == S.iproc[d] - 1)
            if S.iproc[d] == 1 or (edge and not S.periodic[d]):
                nb.append(-1)
                continue
            ip[d] = (S.ip[d] + (-1 if side == 0 else 1)) % S.iproc[d]
            nb.append(hpo.rank_1d(S.iproc, ip))
    return nb


class MultiRankOracle:
    def __init__(self, case):
        self.case = case
        self.nranks = int(np.prod(case.solver["iproc"]))
        self.S = [hpo.Setup(case, rank=r) for r in range(self.nranks)]
        self.O = [hpo.Oracle(s) for s in self.S]
        self.nb = [_neighbors(s) for s in self.S]
        self.nd = self.S[0].ndims
        self.viscous = case.solver["model"] in ("navierstokes3d", "navierstokes2d") and float(case.physics.get("Re", -1)) > 0

    def exchange(self, arrs):
        bufs = {}
        for r in range(self.nranks):
            for d in range(self.nd):
                for side in (0, 1):
                    if self.nb[r][2 * d + side] >= 0:
                        bufs[(r, d, side)] = self.O[r].pack(arrs[r], d, side)
        for r in range(self.nranks):
            for d in range(self.nd):
                for side in (0, 1):
                    peer = self.nb[r][2 * d + side]
                    if peer >= 0:
                        # my ghost on `side` = the peer's interior layers on its opposite face
                        self.O[r].unpack(arrs[r], d, side, bufs[(peer, d, 1 - side)])

    def local_u0(self):
        return [s.local_u0() for s in self.S]

    def rhs(self, u, sbi=None):
        """TimeRHSFunctionExplicit on every rank (u modified: BCs + halos). Returns [rhs_r].
        sbi: optional list of per-rank arrays receiving StageBoundaryIntegral (HyperbolicFunction.c:103-106)"""
        for r in range(self.nranks):
            self.O[r].apply_bc(u[r])
        self.exchange(u)
        out = []
        pieces = []
        for r in range(self.nranks):
            if sbi is not None:
                hpo.lib().hpo_set_boundary_flux_sink(hpo._p(sbi[r]))
            try:
                hyp, w = self.O[r].hyperbolic(u[r], want_weights=True)
            finally:
                hpo.lib().hpo_set_boundary_flux_sink(None)
            pieces.append((hyp, w))
        par = [o.zeros() for o in self.O]
        if self.viscous:
            P1 = [self.O[r].parabolic_p1(u[r]) for r in range(self.nranks)]
            self.exchange([p[1] for p in P1])          # QDerivX
            self.exchange([p[2] for p in P1])          # QDerivY (the reference exchanges it twice, QDerivZ never)
            par = [self.O[r].parabolic_p2(*P1[r]) for r in range(self.nranks)]
        elif self.case.solver["model"] == "linear-advection-diffusion-reaction":
            par = [self.O[r].parabolic(u[r]) for r in range(self.nranks)]
        for r in range(self.nranks):
            hyp, w = pieces[r]
            src = self.O[r].source(u[r], w)
            rhs = np.zeros_like(hyp)
            rhs += -1.0 * hyp
            rhs += par[r]
            rhs += src
            out.append(rhs)
        return out

    def time_step(self, u, dt, rk_type):
        """TimePreStep BC/halo + TimeRK (TimeRK.c:126-195) on every rank, in place."""
        A, b, c = np.zeros(16), np.zeros(4), np.zeros(4)
        ns = hpo.lib().hpo_rk_tableau(rk_type, hpo._p(A), hpo._p(b), hpo._p(c))
        for r in range(self.nranks):
            self.O[r].apply_bc(u[r])
        self.exchange(u)
        k = []
        for s in range(ns):
            U = [x.copy() for x in u]
            for i in range(s):
                for r in range(self.nranks):
                    U[r] += (dt * A[s * ns + i]) * k[i][r]
            k.append(self.rhs(U))
        for s in range(ns):
            for r in range(self.nranks):
                u[r] += (dt * b[s]) * k[s][r]
        return u


    def time_step_glmgee(self, u, uaux, dt, method, mode):
        """TimePreStep BC/halo + TimeGLMGEE (TimeGLMGEE.c:66-151) on every rank, in place (u and uaux)."""
        T = hpo.glmgee_table(method, mode)
        s, A, B, Cm, D = T["s"], T["A"], T["B"], T["C"], T["D"]
        for r in range(self.nranks):
            self.O[r].apply_bc(u[r])
        self.exchange(u)
        k = []
        for j in range(s):
            U = [Cm[2 * j] * x for x in u]
            for r in range(self.nranks):
                U[r] += Cm[2 * j + 1] * uaux[r]
                for i in range(j):
                    U[r] += (dt * A[j * s + i]) * k[i][r]
            k.append(self.rhs(U))
        new = []
        for j in range(2):
            V = [D[2 * j] * x for x in u]
            for r in range(self.nranks):
                V[r] += D[2 * j + 1] * uaux[r]
                for i in range(s):
                    V[r] += (dt * B[j * s + i]) * k[i][r]
            new.append(V)
        for r in range(self.nranks):
            u[r][...] = new[0][r]
            uaux[r][...] = new[1][r]
        return u, uaux

    def time_step_cons(self, u, dt, rk_type):
        """time_step with the boundary-flux bookkeeping of TimeRK.c:172-193; returns [StepBoundaryIntegral_r]"""
        A, b, c = np.zeros(16), np.zeros(4), np.zeros(4)
        ns = hpo.lib().hpo_rk_tableau(rk_type, hpo._p(A), hpo._p(b), hpo._p(c))
        nbf = 2 * self.nd * self.S[0].nvars
        for r in range(self.nranks):
            self.O[r].apply_bc(u[r])
        self.exchange(u)
        k, bf = [], []
        for s in range(ns):
            U = [x.copy() for x in u]
            for i in range(s):
                for r in range(self.nranks):
                    U[r] += (dt * A[s * ns + i]) * k[i][r]
            sbi = [np.zeros(nbf) for _ in range(self.nranks)]
            k.append(self.rhs(U, sbi))
            bf.append(sbi)
        step = [np.zeros(nbf) for _ in range(self.nranks)]
        for s in range(ns):
            for r in range(self.nranks):
                u[r] += (dt * b[s]) * k[s][r]
                step[r] += (dt * b[s]) * bf[s][r]
        return step


class LocalRanks:
    def __init__(self, case, use_fused=True, device=0, sweepwise=False):
        """All ranks of a decomposed case in this process on one GPU, advanced in lock step by the library's own
        distributed step (hpb_TimeStepsLocal / hpb_RHSFunctionLocal) over the in-process transport: pack / unpack kernels,
        face-layer RK kernels, event ordering and schedule are the ones the NCCL run uses; only the copy between the ranks'
        buffers differs (cudaMemcpyAsync instead of ncclSend / ncclRecv).
        sweepwise: the overlapped schedule (hpb_set_overlap 1); otherwise the serial one."""
        import ctypes as C
        from hypar_b200.solver import comm_init_local
        self.nranks = int(np.prod(case.solver["iproc"]))
        self.sv = [Solver.from_case(case, rank=r, device=device, use_fused=use_fused) for r in range(self.nranks)]
        self.viscous = bool(self.sv[0].L.hpb_needs_viscous_exchange(self.sv[0].h))
        self.sweepwise = bool(sweepwise)
        comm_init_local(self.sv)
        for sv in self.sv:
            sv.set_overlap(self.sweepwise)
        self._arr = (C.c_void_p * self.nranks)(*[sv.h for sv in self.sv])

    def _sync(self):
        for sv in self.sv:
            sv.synchronize()

    def set_solution(self, u):
        for sv, x in zip(self.sv, u):
            sv.set_solution(x)

    def get_solution(self):
        self._sync()
        return [sv.get_solution() for sv in self.sv]

    def rhs(self):
        sv = self.sv[0]
        sv._ck(sv.L.hpb_RHSFunctionLocal(self._arr, self.nranks))
        self._sync()
        return [sv.get_stage_rhs(0) for sv in self.sv]

    def exchange_boundaries(self):
        """MPIExchangeBoundariesnD on the device solutions (hpb_ExchangeBoundariesLocal)"""
        sv = self.sv[0]
        sv._ck(sv.L.hpb_ExchangeBoundariesLocal(self._arr, self.nranks))

    def time_step(self, n=1):
        sv = self.sv[0]
        sv._ck(sv.L.hpb_TimeStepsLocal(self._arr, self.nranks, n))

    def close(self):
        for sv in self.sv:
            sv.close()
