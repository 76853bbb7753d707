"""Parity table per BASELINE.json configuration (TEST INFRASTRUCTURE: imports the oracle, runs oracle/_ref).

For C1 .. C5b at sizes the CPU reference finishes in seconds, one TimeRHSFunctionExplicit of the same input:

  ref-noise : the UNMODIFIED reference against ITSELF built another way -- oracle/_ref/hypar_ref_fma (gcc -O3 -march=native
              -ffp-contract=fast: x86 FMA contraction, still IEEE operations) and oracle/_ref/hypar_ref_fast (-ffast-math) --
              versus the default build (gcc -O3, no contraction; = the oracle, bit for bit). This is the rounding noise the
              reference's own right-hand side carries: no implementation that evaluates the same formulas in another order /
              with another contraction can agree with it better than the reference agrees with itself.
  exact     : the library's exact path (use_fused = 0) vs the oracle                      [needs a GPU]
  fused     : the production path (fused sweeps, FMA, division-free weights) vs the oracle [needs a GPU]

Errors are absolute Linf differences of rhs, divided by (a) max|rhs|, (b) max(|hyp|, |par|, |source|) = the size of the terms
that are summed, (c) the rounding floor 16 ulp (CFL/dt) max|u| of tests/test_gpu_parity.py::fused_tolerance.

    python tools/parity_table.py [--out profiles/parity_table.txt] [--no-gpu]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

from hypar_b200 import cases
from oracle import hpo
from refrun import ref_available, run_reference


def configs():
    c2 = cases.euler1d_sod(201, "js")
    return [
        ("C1  LinearADR sine, WENO5-JS, RK4, 1024", cases.linear_advection_sine(1024, "js")),
        ("C2  Euler1D Sod, char WENO5 + Roe, SSPRK3, 201", c2),
        ("C3  NS2D vortex, WENO5-JS + Rusanov, 128x128", cases.ns2d_vortex((128, 128), "js")),
        ("C4  NS3D turbulence, mapped + Rusanov + viscous, 48^3", cases.ns3d_turbulence((48, 48, 48), "mapped")),
        ("C4r NS3D turbulence, mapped + Roe + viscous, 48^3", cases.ns3d_turbulence((48, 48, 48), "mapped", upwinding="roe")),
        ("C5a NS3D density wave, JS + Rusanov, 48^3", cases.ns3d_density_wave((48, 48, 48), "js")),
        ("C5b NS3D rising bubble + gravity, YC + Rusanov, 48^3", cases.ns3d_rising_bubble((48, 48, 48), "yc")),
        ("C5b' rising bubble after 20 steps (moving flow), 32^3", "bubble_moving"),
    ]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-gpu", action="store_true")
    args = ap.parse_args()
    gpu = False
    if not args.no_gpu:
        try:
            from hypar_b200 import _lib
            gpu = _lib.load().hpb_device_count() > 0
        except Exception:
            gpu = False
    lines = []

    def emit(s=""):
        print(s, flush=True)
        lines.append(s)

    emit("# parity table: one TimeRHSFunctionExplicit per BASELINE.json configuration (tools/parity_table.py)")
    emit(f"# GPU columns: {'measured on ' + os.popen('nvidia-smi --query-gpu=name --format=csv,noheader').read().strip() if gpu else 'not available (no device)'}")
    emit("# err = max|rhs - rhs_ref| ;  /rhs = err / max|rhs_ref| ;  /terms = err / max(|hyp|,|par|,|src|) ;  /floor = err / (16 ulp (CFL/dt) max|u|)")
    emit(f"{'configuration':58s} {'who':12s} {'err':>10s} {'/rhs':>10s} {'/terms':>10s} {'/floor':>10s}")
    for label, case in configs():
        if case == "bubble_moving":
            case = cases.ns3d_rising_bubble((32, 32, 32), "yc")
            S = hpo.Setup(case)
            O = hpo.Oracle(S)
            u = S.local_u0()
            for _ in range(20):
                O.time_step(u, float(case.solver["dt"]), hpo.rk_type_of(case))
            case.u0 = S.interior(u).copy()
        S = hpo.Setup(case)
        O = hpo.Oracle(S)
        u_ref = S.local_u0()
        rhs_ref, hyp_ref, par_ref, src_ref = O.rhs(u_ref, parts=True)
        dt = float(case.solver["dt"])
        terms = max(np.abs(hyp_ref).max(), np.abs(par_ref).max(), np.abs(src_ref).max())
        floor = 16 * np.finfo(np.float64).eps * (O.cfl(u_ref, dt) / dt) * np.abs(u_ref).max()
        rmax = np.abs(rhs_ref).max()

        def row(who, rhs):
            err = float(np.abs(np.asarray(rhs).ravel() - rhs_ref.ravel()).max())
            emit(f"{label:58s} {who:12s} {err:10.2e} {err / rmax:10.2e} {err / terms:10.2e} {err / floor:10.2e}")

        for exe, who in (("hypar_ref_mpi1", "ref(default)"), ("hypar_ref_fma", "ref(FMA)"), ("hypar_ref_fast", "ref(fast)")):
            if not ref_available(exe):
                continue
            try:
                out = run_reference(case, "rhs", exe=exe, threads=1)
                row(who, out["rhs"]["data"])
            except Exception as ex:
                emit(f"{label:58s} {who:12s} failed: {str(ex)[:60]}")
        if gpu:
            from hypar_b200.solver import Solver
            for fused, who in ((False, "b200 exact"), (True, "b200 fused")):
                sv = Solver.from_case(case, use_fused=fused)
                u = S.local_u0()
                row(who, sv.RHSFunction(u))
                sv.close()
        emit()
    if args.out:
        open(args.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
