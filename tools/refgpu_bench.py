"""The reference's OWN CUDA path as a second baseline (TEST / MEASUREMENT INFRASTRUCTURE).

oracle/_ref/hypar_ref_gpu = the unmodified reference sources compiled with -DHAVE_CUDA for sm_100a instead of the sm_70
they are pinned to (recipe: oracle/Makefile `refgpu`; src/CMakeLists.txt:38). This script writes a HyPar run directory of
configuration C4 (NavierStokes3D, WENO5 mapped, Rusanov or Roe, viscous, RK4 or SSPRK3) with `use_gpu yes`, runs the
reference's own main in it on GPU 0 and reports HyPar's own per-iteration wall-clock time as Mpoint-RK-stage/s, next to
the B200 library on the same grid (device-resident loop) and the largest difference between the two final solutions.

    python tools/refgpu_bench.py [--n 64 128] [--steps 10] [--upwinding rusanov] [--out gpurun_out/refgpu.json]
"""
from __future__ import annotations

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
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np

from hypar_b200 import cases, hypario

EXE = os.path.join(ROOT, "oracle", "_ref", "hypar_ref_gpu")


def run_ref_gpu(case, steps, keep=None, exe=EXE, timeout=1800):
    d = keep or tempfile.mkdtemp(prefix="hpbrefgpu_")
    try:
        c = cases.Case(**{**case.__dict__})
        c.solver = dict(case.solver)
        c.solver.update({"n_iter": steps, "screen_op_iter": 1, "file_op_iter": steps, "use_gpu": "yes", "gpu_device_no": 0,
                         "op_overwrite": "yes"})
        c.write(d)
        t0 = time.time()
        p = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=timeout)
        wall = time.time() - t0
        if p.returncode != 0:
            raise RuntimeError(f"hypar_ref_gpu failed ({p.returncode}):\n{p.stdout[-3000:]}\n{p.stderr[-2000:]}")
        wct = [float(m.group(1)) for m in re.finditer(r"wctime:\s*([0-9.Ee+-]+)", p.stdout)]
        op = hypario.read_op_bin(os.path.join(d, "op.bin")) if os.path.exists(os.path.join(d, "op.bin")) else None
        return {"wctime": wct, "wall_s": wall, "stdout_tail": p.stdout[-1500:], "op": op}
    finally:
        if keep is None:
            shutil.rmtree(d, ignore_errors=True)


def run_b200(case, steps):
    import ctypes as C
    from hypar_b200.solver import Solver
    from oracle import hpo
    sv = Solver.from_case(case)
    S = hpo.Setup(case)
    sv.set_solution(S.local_u0())
    sv.TimeSteps(2)                    # warm-up (also pays the first-launch costs)
    sv.set_solution(S.local_u0())
    sv.synchronize()
    t0 = time.perf_counter()
    sv.TimeSteps(steps)
    sv.synchronize()
    sec = (time.perf_counter() - t0) / steps
    u = S.interior(sv.get_solution())
    sv.close()
    return sec, u


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[64, 128])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--upwinding", default="rusanov")
    ap.add_argument("--out", default=None)
    ap.add_argument("--prepare", default=None, help="only write the run directory of the first --n here (for `ncu ... hypar_ref_gpu`)")
    ap.add_argument("--tstype", default="44", help="RK type: 44 (C4) or ssprk3 (the reference's DNS example)")
    args = ap.parse_args()
    if args.prepare:
        n = args.n[0]
        c = cases.ns3d_turbulence((n, n, n), "mapped", upwinding=args.upwinding, tstype=args.tstype)
        c.solver.update({"n_iter": args.steps, "screen_op_iter": 1, "file_op_iter": args.steps, "use_gpu": "yes", "gpu_device_no": 0,
                         "op_overwrite": "yes"})
        c.write(args.prepare)
        return
    if not os.access(EXE, os.X_OK):
        raise SystemExit(f"{EXE} is missing (make -C oracle refgpu, where /root/reference exists)")
    recs = []
    for n in args.n:
        case = cases.ns3d_turbulence((n, n, n), "mapped", upwinding=args.upwinding, tstype=args.tstype)
        nst = 4 if args.tstype == "44" else 3
        r = run_ref_gpu(case, args.steps)
        w = r["wctime"][2:] if len(r["wctime"]) > 4 else r["wctime"]       # HyPar's own per-iteration wall clock, first two dropped
        sec_ref = float(np.median(w))
        rec = {"grid": f"{n}^3", "upwinding": args.upwinding, "rk": args.tstype, "steps": args.steps, "rk_stages": nst,
               "ref_gpu_s_per_step": sec_ref, "ref_gpu_mpoint_rk_stage_per_s": n ** 3 * nst / sec_ref / 1e6,
               "ref_gpu_wall_s": r["wall_s"]}
        try:
            sec, u = run_b200(case, args.steps)
            rec.update({"b200_s_per_step": sec, "b200_mpoint_rk_stage_per_s": n ** 3 * nst / sec / 1e6,
                        "speedup": sec_ref / sec})
            if r["op"] is not None:
                uo = r["op"][1] if isinstance(r["op"], tuple) else r["op"]
                uo = np.asarray(uo).reshape(u.shape)
                rec["max_rel_diff_final_solution"] = float(np.abs(uo - u).max() / np.abs(u).max())
        except Exception as ex:     # the library's arm is a report here, the reference's number is the point
            rec["b200_error"] = str(ex)
        recs.append(rec)
        print(json.dumps(rec), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump({"what": "reference's own CUDA path (HAVE_CUDA, unmodified sources, sm_100a) vs hypar_b200, configuration C4",
                   "records": recs}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
