"""HyPar's ensemble driver on the B200 path: ``nsims`` independent simulations of the same equations on grids of their own,
advanced in lock-step with the same time step (reference: src/Simulation/EnsembleSimulations*.cpp, selected by the
presence of ``simulation.inp``, src/main.cpp:268-320).

What the reference shares and what it keeps per simulation (and so does this class):

  solver.inp     shared, except ``size`` / ``iproc`` / ``size_exact``: one vector per simulation (ReadInputs.c:186-250)
  boundary.inp   ``boundary_<n>.inp`` if it exists, else the shared ``boundary.inp`` (InitializeBoundaries.c:52-80)
  physics.inp, weno.inp, muscl.inp   shared
  initial.inp    ``initial_<n>.inp`` (InitialSolution.c:36-43); LinearADR ``advection_filename``: ``<name>_<n>.inp``
                 (LinearADRAdvectionField.c:66-76)
  output         ``op_<n>.bin`` / ``op_<n>_<iter>.bin`` (OutputSolution.cpp:52-60)

``<n>`` is written with ``(int) log10(nsims) + 1`` digits (CommonFunctions.c GetStringFromInteger). With an explicit
Runge-Kutta integrator the simulations never exchange data: the reference's TimeRK walks one concatenated vector, sim
after sim (TimeRK.c:60-160). Here every simulation is one ``Solver`` (one ``hpb_solver`` with its own stream), so their
kernels may overlap on the device. The screen diagnostics follow the reference's loops: the step norm is taken over all
simulations together (TimePostStep.c:44-63), the reported CFL is that of the LAST simulation (TimePreStep.c:84-92 overwrites
``max_cfl`` in every pass of the loop).
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence

import numpy as np

from . import hypario
from .solver import HyParB200Error, Solver


def index_string(n: int, nsims: int) -> str:
    """GetStringFromInteger(n, ., (int) log10(nsims) + 1): zero-padded decimal."""
    width = int(math.log10(nsims)) + 1
    return f"{n:0{width}d}"


class Ensemble:
    """``nsims`` solvers stepped together (one rank per simulation: ``iproc`` = 1 everywhere, or pass ``rank``)."""

    def __init__(self, sims: Sequence[Solver], n_global: Optional[Sequence[int]] = None):
        if not sims:
            raise HyParB200Error("an ensemble needs at least one simulation")
        self.sims: List[Solver] = list(sims)
        dts = {s.dt for s in self.sims}
        if len(dts) != 1:
            raise HyParB200Error("the simulations of an ensemble share the time step")
        self.dt = self.sims[0].dt
        self.npoints_global = [int(np.prod(s.dim_global)) for s in self.sims] if n_global is None else list(n_global)
        self.u0_global: List[Optional[np.ndarray]] = [None] * len(self.sims)

    # -- construction
    @classmethod
    def from_directory(cls, path: str, rank: int = 0, device: int = -1, use_fused: bool = True) -> "Ensemble":
        nsims = hypario.read_simulation_inp(os.path.join(path, "simulation.inp"))
        if nsims < 1:
            raise HyParB200Error("simulation.inp: nsims must be at least 1")
        cfgs = hypario.read_ensemble_solver_inp(os.path.join(path, "solver.inp"), nsims)
        s0 = cfgs[0]
        nd, nv = int(s0["ndims"]), int(s0["nvars"])
        vk = {"gravity": 3 if s0["model"] == "navierstokes3d" else 2 if s0["model"] == "navierstokes2d" else 1,
              "advection": nd * nv, "diffusion": nd * nv}
        pf = os.path.join(path, "physics.inp")
        ph = hypario.read_keyword_file(pf, vector_keys=vk) if os.path.exists(pf) else {}
        for k in ("advection", "diffusion", "gravity"):
            if k in ph:
                ph[k] = [float(v) for v in (ph[k] if isinstance(ph[k], (list, tuple)) else [ph[k]])]
        wf, mf = os.path.join(path, "weno.inp"), os.path.join(path, "muscl.inp")
        w = hypario.read_keyword_file(wf) if os.path.exists(wf) else None
        mu = hypario.read_keyword_file(mf) if os.path.exists(mf) else None
        ipt = str(s0.get("ip_file_type", "ascii"))
        if ipt not in ("ascii", "binary", "bin"):
            raise HyParB200Error(f"ip_file_type '{ipt}' is neither ascii nor binary")
        sims, u0s = [], []
        try:
            for n, s in enumerate(cfgs):
                tag = "_" + index_string(n, nsims) if nsims > 1 else ""
                bf = os.path.join(path, f"boundary{tag}.inp")
                if not os.path.exists(bf):
                    bf = os.path.join(path, "boundary.inp")
                b = hypario.read_boundary_inp(bf, nd, nv)
                x, u0 = hypario.read_initial(os.path.join(path, f"initial{tag}.inp"), s["size"], nv, ipt)
                af = None
                if str(ph.get("advection_filename", "none")) != "none":
                    fn = os.path.join(path, f"{ph['advection_filename']}{tag}.inp")
                    if os.path.exists(fn):
                        af = hypario.read_initial(fn, s["size"], nd * nv, ipt)[1]
                sims.append(Solver(s, b, ph, w, x, rank, device, use_fused, muscl=mu, advection_field=af))
                u0s.append(u0)
        except Exception:
            for sv in sims:
                sv.close()
            raise
        obj = cls(sims)
        obj.u0_global = u0s
        return obj

    def close(self) -> None:
        for s in self.sims:
            s.close()
        self.sims = []

    def __len__(self) -> int:
        return len(self.sims)

    # -- stepping (device-resident)
    def set_initial_solutions(self) -> None:
        """Every simulation's block of its ``initial_<n>.inp`` onto the device."""
        for s, u0 in zip(self.sims, self.u0_global):
            if u0 is None:
                raise HyParB200Error("no initial solution was read for this ensemble")
            s.set_solution(s.local_from_global(u0))

    def TimeSteps(self, n: int) -> None:
        """``n`` steps of every simulation. Each solver enqueues on its own stream; nothing synchronises in between."""
        for s in self.sims:
            s.TimeSteps(n)

    def synchronize(self) -> None:
        for s in self.sims:
            s.synchronize()

    def get_solutions(self) -> List[np.ndarray]:
        return [s.get_solution() for s in self.sims]

    # -- screen diagnostics of the reference's loop
    def dev_ComputeCFL(self) -> float:
        """TimePreStep.c:84-92: each simulation overwrites ``max_cfl``; the last one is what HyPar prints."""
        cfl = -1.0
        for s in self.sims:
            cfl = s.dev_ComputeCFL()
        return cfl

    def dev_StepNorm(self) -> float:
        """TimePostStep.c:44-63: sqrt(sum over simulations of |u^{n+1} - u^n|^2 / sum of their global point counts)."""
        return math.sqrt(self.StepNormSumSq() / float(sum(self.npoints_global)))

    def StepNormSumSq(self) -> float:
        return float(sum(s.dev_StepNormSumSq() for s in self.sims))

    # -- output in the reference's naming
    def write_solutions(self, path: str, index: Optional[int] = None) -> List[str]:
        """Every simulation's solution file in the reference's naming (OutputSolution.cpp:52-60: root ``op_<n>``), format and
        ``op_overwrite`` from solver.inp; ``index`` overrides the running file index of ``op_overwrite no``."""
        nsims = len(self.sims)
        names = []
        for n, s in enumerate(self.sims):
            root = "op" + ("_" + index_string(n, nsims) if nsims > 1 else "")
            names.append(s.write_solution(path, 0 if index is None else index, root))
        return names
