"""HyPar's own executable with the library attached, as a PERFORMANCE path: configuration C4 at 128^3 /
256^3 through oracle/_ref/hypar_b200_dropin (resident mode) -- HyPar's main, Solve(), TimePreStep / TimeStep / TimePostStep,
its own wall clock -- against the same steps through the Python driver (Solver.TimeSteps). TEST / MEASUREMENT INFRASTRUCTURE.

    python tools/dropin_bench.py [--n 128 256] [--steps 20] [--out gpurun_out/dropin.json]
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypar_b200 import cases

EXE = os.path.join(ROOT, "oracle", "_ref", "hypar_b200_dropin")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[128, 256])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    recs = []
    for n in args.n:
        case = cases.ns3d_turbulence((n, n, n), "mapped")
        # no screen / file output inside the timed steps: on a screen-output step HyPar itself copies u, forms u - u_prev and sums
        # its squares on the host (TimePreStep.c:84, TimePostStep.c:44-63: ~0.5 s at 256^3, single-threaded index arithmetic) --
        # the reference's own use_gpu path skips that norm altogether (TimePostStep.c:38-40 sets it to -1)
        case.solver.update({"n_iter": args.steps, "screen_op_iter": 1000 * args.steps, "file_op_iter": 1000 * args.steps, "op_overwrite": "yes"})
        d = tempfile.mkdtemp(prefix="hpbdropin_")
        try:
            case.write(d)
            t0 = time.time()
            p = subprocess.run([EXE], cwd=d, capture_output=True, text=True, timeout=1800,
                               env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1)))
            wall = time.time() - t0
            if p.returncode:
                raise RuntimeError(p.stdout[-2000:] + p.stderr[-2000:])
            m = re.search(r"total wctime: ([0-9.Ee+-]+)", p.stdout)
            sec = float(m.group(1)) / args.steps
            rec = {"grid": f"{n}^3", "steps": args.steps, "dropin_s_per_step": sec,
                   "dropin_mpoint_rk_stage_per_s": n ** 3 * 4 / sec / 1e6, "dropin_wall_s": wall}
        finally:
            shutil.rmtree(d, ignore_errors=True)
        from hypar_b200.solver import Solver
        from oracle import hpo
        sv = Solver.from_case(case)
        S = hpo.Setup(case)
        sv.set_solution(S.local_u0())
        sv.TimeSteps(2)
        sv.synchronize()
        t0 = time.perf_counter()
        sv.TimeSteps(args.steps)
        sv.synchronize()
        sec_py = (time.perf_counter() - t0) / args.steps
        sv.close()
        rec.update({"python_s_per_step": sec_py, "python_mpoint_rk_stage_per_s": n ** 3 * 4 / sec_py / 1e6,
                    "dropin_over_python": sec_py / sec})
        recs.append(rec)
        print(json.dumps(rec), flush=True)
    if args.out:
        json.dump({"what": "HyPar's executable + libhypar_b200 (resident mode, HyPar's own total wctime / steps) vs the Python driver, C4",
                   "records": recs}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
